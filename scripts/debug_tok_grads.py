import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from helpers import det_weights, rel_l2
from oracle import fixtures as fx, genie_oracle as O
import open_genie_b200 as og
g = torch.load('tests/golden/tokenizer_mini.pt', weights_only=False)
tok = og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=6, gan_loss_weight=0, perc_loss_weight=0)
det_weights(tok); tok.cuda().train()
video = O.det_uniform('tokenizer.video', fx.MINI_VIDEO_SHAPE).cuda()
loss, aux = tok(video); loss.backward()
print('loss', loss.item(), g['loss'].item())
for k, p in tok.named_parameters():
    n = g['grads']['norm'][k]; gn = p.grad.float().norm().item()
    full = g['grads']['full'].get(k)
    r = rel_l2(p.grad.float().cpu(), full) if full is not None else float('nan')
    print(f'{k:45s} ref_norm {n:.4e} got {gn:.4e} ratio {gn/max(n,1e-30):.3f} rel_l2 {r:.3f}')
