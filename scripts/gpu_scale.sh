#!/bin/bash
# weak-scaling check on one 8-GPU box: N = 1, 2, 4, 8 back to back (what the driver does at round end)
set -u
TAG=${1:-r02i}
O=gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_scale_n1.json 2> $O/${TAG}_scale_n1.err
python -c "import json;d=json.loads(open('$O/${TAG}_scale_n1.json').read().strip().splitlines()[-1]);print('N=1',d['value'],d['ms_per_step'],d['clocks'])"
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_scale_n$N.json 2> $O/${TAG}_scale_n$N.err
  echo "N=$N rc=$?"
  python -c "import json;d=json.loads(open('$O/${TAG}_scale_n$N.json').read().strip().splitlines()[-1]);print('N=$N',d['value'],d['ms_per_step'],d['config']['launch'][:40],d['clocks'])" || tail -5 $O/${TAG}_scale_n$N.err
done
