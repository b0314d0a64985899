#!/bin/bash
# GroupNorm backward over sample groups (OG_GN_BWD_GROUPS): parity with the grouping forced on small tensors, then step time per setting
set -u
TAG=${1:-r02cc}
O=gpurun_out
OG_GN_BWD_GROUPS=2 OG_GN_BWD_MIN_MB=0 python -m pytest tests/test_gpu_layers.py tests/test_gpu_tokenizer.py tests/test_gpu_full_configs.py -m gpu -q > $O/${TAG}_tests_groups2.log 2>&1; tail -2 $O/${TAG}_tests_groups2.log
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);k=d['roofline']['kernels'];print('$2',round(d['value'],1),round(d['ms_per_step'],2),d['clocks']['sm_mhz'],'launches/step',d['gpu_launches']//d['steps'],'gn_bwd',round(k['og_gn_act_bwd']['ms_per_step'],2),'reduce',round(k['og_affine_act_bwd_reduce']['ms_per_step'],2))" || tail -3 ${1%.json}.err; }
for g in 1 4 8 2 1; do
  OG_GN_BWD_GROUPS=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_g$g.json 2> $O/${TAG}_bench_g$g.err; show $O/${TAG}_bench_g$g.json groups=$g
done
