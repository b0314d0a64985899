#!/bin/bash
# attention kernels: parity tests + isolated timing at the LatentAction shape, new vs first backward version
set -u
TAG=${1:-r02r}
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x > $O/${TAG}_attn_tests.log 2>&1; tail -5 $O/${TAG}_attn_tests.log
echo "== pipelined"; timeout 120 python scripts/ncu_attention.py
echo "== v1"; OG_FLASH_BWD_V1=1 timeout 120 python scripts/ncu_attention.py
