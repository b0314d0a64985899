#!/bin/bash
# attention kernels: parity tests + isolated timing at the LatentAction shape, new vs first backward version
set -u
TAG=${1:-r02r}
O=gpurun_out
for v in ${VARIANTS:-OG_FLASH_BWD_WARPS=8}; do
  echo "== tests $v"; env $v timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x > $O/${TAG}_attn_tests_${v}.log 2>&1; tail -2 $O/${TAG}_attn_tests_${v}.log
  echo "== timing $v"; env $v timeout 120 python scripts/ncu_attention.py | tail -1
done
echo "== v1"; OG_FLASH_BWD_V1=1 timeout 120 python scripts/ncu_attention.py | tail -1
