#!/bin/bash
# Final single-GPU evidence of the round: GPU test suite + smoke, bench (both arms), config timings, ncu launch list.
set -u
TAG=${1:-r02m}
O=gpurun_out
python -m pytest tests -m gpu -q --junitxml=$O/${TAG}_gputest.xml > $O/${TAG}_gputest.log 2>&1; tail -5 $O/${TAG}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; tail -c 300 $O/${TAG}_bench_reference.json
python scripts/bench_configs.py tokenize action dynamics genie > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err; cut -c1-160 $O/${TAG}_configs.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3300 -c 1800 --csv \
  --log-file $O/${TAG}_step_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/${TAG}_ncu_step.log 2>&1; tail -1 $O/${TAG}_ncu_step.log | cut -c1-200
