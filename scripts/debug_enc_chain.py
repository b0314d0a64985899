import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from helpers import det_weights, rel_l2, bf16_round, round_conv_weights
from oracle import fixtures as fx, genie_oracle as O
import open_genie_b200 as og
from open_genie_b200 import ops
tok = og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=6, gan_loss_weight=0, perc_loss_weight=0)
sd = det_weights(tok); tok.cuda()
video = bf16_round(O.det_uniform('tokenizer.video', fx.MINI_VIDEO_SHAPE))
enc = tok.encode(video.cuda())
gup = O.det_uniform('enc.upstream', tuple(enc.shape))
enc.backward(gup.cuda().to(enc.dtype).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3))
sdr = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in round_conv_weights(sd).items()}
enc_o = O.tokenizer_encode(sdr, fx.MINI_ENC, video); enc_o.backward(gup)
print('enc rel', rel_l2(ops.to_reference(enc).cpu(), enc_o.detach()))
for k, p in tok.named_parameters():
    if k.startswith('enc_layers'):
        print(f'{k:40s} {rel_l2(p.grad.float().cpu(), sdr[k].grad):.4f}  norm {p.grad.float().norm().item():.4e} ref {sdr[k].grad.norm().item():.4e}')
