#!/bin/bash
# N = 2: effect of NCCL's CTA cap and of the all-reduce range size on the graphed step
set -u
TAG=${1:-r02j}
O=gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  python -c "import json;d=json.loads(open('$O/${TAG}_$name.json').read().strip().splitlines()[-1]);print('$name',round(d['value'],1),round(d['ms_per_step'],2),d['clocks']['sm_mhz'])" || tail -3 $O/${TAG}_$name.err
}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_n1.json 2>/dev/null
python -c "import json;d=json.loads(open('$O/${TAG}_n1.json').read().strip().splitlines()[-1]);print('N=1',round(d['value'],1),round(d['ms_per_step'],2),d['clocks']['sm_mhz'])"
run ctas_default_b128 NCCL_MAX_CTAS=64 OG_BUCKET_MB=128
run ctas8_b64 NCCL_MAX_CTAS=8 OG_BUCKET_MB=64
run ctas4_b64 NCCL_MAX_CTAS=4 OG_BUCKET_MB=64
run ctas2_b64 NCCL_MAX_CTAS=2 OG_BUCKET_MB=64
run ctas4_b32 NCCL_MAX_CTAS=4 OG_BUCKET_MB=32
run ctas1_b32 NCCL_MAX_CTAS=1 OG_BUCKET_MB=32
OG_WGRAD_PLAIN_STORE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_n1_plainstore.json 2>/dev/null
python -c "import json;d=json.loads(open('$O/${TAG}_n1_plainstore.json').read().strip().splitlines()[-1]);print('N=1 plain-store wgrad',round(d['value'],1),round(d['ms_per_step'],2),d['roofline']['kernels']['og_conv_wgrad_kernel']['ms_per_step'])"
