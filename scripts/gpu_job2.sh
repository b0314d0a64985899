#!/bin/bash
# tests (all, no -x) + bench N=1 + config timings
set -u
TAG=${1:-r02g}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${TAG}_gputest.log 2>&1; tail -15 $O/${TAG}_gputest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 300 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
python scripts/bench_configs.py action dynamics > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err; cut -c1-200 $O/${TAG}_configs.jsonl
