"""Per-layer timing of one eager training step (CUDA events around every tensor-core launch)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import open_genie_b200 as og
from open_genie_b200 import ops
torch.manual_seed(0)
m = og.VideoTokenizer(og.MAGVIT2_ENC_DESC, og.MAGVIT2_DEC_DESC, gan_loss_weight=0, perc_loss_weight=0).cuda()
opt = m.configure_optimizers()
v = torch.randn(int(os.environ.get('B', 8)), 3, 16, 64, 64, device='cuda')
def step():
    loss = m.training_step(v, 0); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(3): step()
torch.cuda.synchronize()
ops.PROFILE = []
step(); torch.cuda.synchronize()
agg = collections.OrderedDict()
for kind, flops, a, b, shape in ops.PROFILE:
    k = (kind, shape)
    d = agg.setdefault(k, [0, 0.0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b); d[2] += flops
tot = sum(d[1] for d in agg.values())
print(f'total conv ms {tot:.2f}')
for (kind, shape), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'{ms:7.3f} ms {100*ms/tot:5.1f}% n={n:3d} {fl/ms*1e-9:7.1f} TF  {kind:5s} {shape}')
