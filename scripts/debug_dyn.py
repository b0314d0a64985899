import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, open_genie_b200 as og
torch.manual_seed(0)
B, L = 8, 8
dm = og.DynamicsModel((('space-time_attn', {'n_rep': L, 'n_head': 8, 'd_head': 64, 'transpose': False}),), tok_vocab=1024, act_vocab=8, embed_dim=512).cuda()
tok = torch.randint(0, 1024, (B, 16, 16, 16), device='cuda'); act = torch.randint(0, 8, (B, 16), device='cuda')
mask = torch.rand(B, 16, 16, 16, device='cuda') < 0.75
opt = og.FusedAdamW(dm.parameters())
for i in range(4):
    logits = dm._logits(torch.masked_fill(tok, mask, 0), act)
    print('step', i, 'logits absmax', logits.float().abs().max().item(), 'mean', logits.float().mean().item(), 'finite', torch.isfinite(logits.float()).all().item())
    loss = dm.compute_loss(tok, act, mask=mask); print('  loss', loss.item())
    loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
x = og.ops.embed_add(tok, act, dm.tok_emb.weight, dm.act_emb[0].weight)
for j, dec in enumerate(dm.dec_layers):
    x = dec(x); print('block', j, 'absmax', x.float().abs().max().item(), 'std', x.float().std().item())
