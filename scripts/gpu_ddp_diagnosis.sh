#!/bin/bash
set -u
TAG=${1:-r02l}
O=gpurun_out
python -m pytest tests/test_gpu_layers.py tests/test_gpu_tokenizer.py tests/test_gpu_full_configs.py -m gpu -q > $O/${TAG}_gputest.log 2>&1; tail -4 $O/${TAG}_gputest.log
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);k=d['roofline']['kernels'];print('$2',round(d['value'],1),round(d['ms_per_step'],2),d['clocks']['sm_mhz'],'igemm',round(k['og_conv_igemm_kernel']['ms_per_step'],2),'wgrad',round(k['og_conv_wgrad_kernel']['ms_per_step'],2))" || tail -3 ${1%.json}.err; }
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err; show $O/${TAG}_n1.json N1
i=0
for mode in "OG_DDP_MODE=none" "OG_BUCKET_MB=100000" "OG_BUCKET_MB=64" "OG_BUCKET_MB=16"; do
  i=$((i+1))
  env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2967$i bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_n2_$i.json 2> $O/${TAG}_n2_$i.err
  echo rc=$?; show $O/${TAG}_n2_$i.json "N2 $mode"
done
