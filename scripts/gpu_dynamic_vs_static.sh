#!/bin/bash
set -u
TAG=${1:-r02k}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${TAG}_gputest.log 2>&1; tail -6 $O/${TAG}_gputest.log
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);k=d['roofline']['kernels'];print('$2',round(d['value'],1),round(d['ms_per_step'],2),d['clocks']['sm_mhz'],'igemm',round(k['og_conv_igemm_kernel']['ms_per_step'],2),'wgrad',round(k['og_conv_wgrad_kernel']['ms_per_step'],2))" || tail -3 ${1%.json}.err; }
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_n1_dyn.json 2> $O/${TAG}_n1_dyn.err; show $O/${TAG}_n1_dyn.json N1-dynamic
OG_IGEMM_DYNAMIC=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_n1_static.json 2> $O/${TAG}_n1_static.err; show $O/${TAG}_n1_static.json N1-static
for mode in 1 0; do
  OG_IGEMM_DYNAMIC=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2966$mode bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_n2_dyn$mode.json 2> $O/${TAG}_n2_dyn$mode.err
  echo rc=$?; show $O/${TAG}_n2_dyn$mode.json N2-dynamic=$mode
done
