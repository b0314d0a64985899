#!/bin/bash
# One GPU-box job of round 2 (kept for reproducibility of the files under profiles/): tests, bench, config timings, ncu passes.
set -u
TAG=${1:-r02f}
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/${TAG}_gputest.log 2>&1; tail -6 $O/${TAG}_gputest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 400 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
python scripts/bench_configs.py tokenize action dynamics genie > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err; cut -c1-300 $O/${TAG}_configs.jsonl; tail -3 $O/${TAG}_configs.err
OG_TEMPORAL_MMA=1 python scripts/bench_configs.py action dynamics > $O/${TAG}_configs_tmma.jsonl 2> $O/${TAG}_configs_tmma.err; cut -c1-300 $O/${TAG}_configs_tmma.jsonl; tail -3 $O/${TAG}_configs_tmma.err
# every launch of ~2 eager tokenizer steps with its device time and DRAM bytes (cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3300 -c 1900 --csv \
  --log-file $O/${TAG}_step_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/${TAG}_ncu_step.log 2>&1; tail -2 $O/${TAG}_ncu_step.log
# full sections for the bandwidth / SFU bound kernels at their largest shapes
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^og_ -o $O/${TAG}_hbm python scripts/ncu_hbm_kernels.py > $O/${TAG}_ncu_hbm.log 2>&1; tail -3 $O/${TAG}_ncu_hbm.log
ls -la $O | grep ${TAG}
