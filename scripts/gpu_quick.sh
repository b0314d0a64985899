#!/bin/bash
# quick single-GPU check: full GPU suite + bench (graph) with and without a switch given as $2 (e.g. OG_SPLITK_FUSED=1)
set -u
TAG=${1:-r02o}
ALT=${2:-OG_SPLITK_FUSED=1}
O=gpurun_out
if [ -n "${NATIVE:-}" ]; then   # low-resolution split-K shapes, isolated (tests/native/test_conv.cu): NATIVE="VAR=1 VAR2=0 ..."
  for v in _ $NATIVE; do
    echo "== native $v"; env ${v/_/X_=0} tests/native/bin/test_conv benchonly 50 2 | grep bench; env ${v/_/X_=0} tests/native/bin/test_conv benchonly 50 3 | grep bench
  done
fi
python -m pytest tests -m gpu -q > $O/${TAG}_gputest.log 2>&1; tail -4 $O/${TAG}_gputest.log
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);k=d['roofline']['kernels'];print('$2',round(d['value'],1),round(d['ms_per_step'],2),d['clocks']['sm_mhz'],'launches/step',d['gpu_launches']//d['steps'],'igemm',round(k['og_conv_igemm_kernel']['ms_per_step'],2),'wgrad',round(k['og_conv_wgrad_kernel']['ms_per_step'],2))" || tail -3 ${1%.json}.err; }
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; show $O/${TAG}_bench.json default
env $ALT python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_alt.json 2> $O/${TAG}_bench_alt.err; show $O/${TAG}_bench_alt.json "$ALT"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench2.json 2> $O/${TAG}_bench2.err; show $O/${TAG}_bench2.json default-again
