#!/bin/bash
set -u
TAG=${1:-r02h}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${TAG}_gputest.log 2>&1; tail -8 $O/${TAG}_gputest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 300 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
python scripts/bench_configs.py action dynamics genie > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err; cut -c1-200 $O/${TAG}_configs.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"og_rope|og_temporal" -o $O/${TAG}_attn_rows python scripts/ncu_hbm_kernels.py > $O/${TAG}_ncu_attn_rows.log 2>&1; tail -3 $O/${TAG}_ncu_attn_rows.log
