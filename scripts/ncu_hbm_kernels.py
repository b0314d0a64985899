"""One launch of every bandwidth- / SFU-bound kernel north_star names, at its largest shape in BASELINE's configs, for
`ncu --set full` (dram__bytes vs algorithmic bytes, achieved GB/s). Prints the algorithmic bytes of each launch so the
summary in profiles/ can put them next to ncu's DRAM counters.

    ncu --set full --clock-control none --import-source on -k regex:^og_ -o gpurun_out/r02_hbm python scripts/ncu_hbm_kernels.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_genie_b200 import _lib, ops
import open_genie_b200 as og

dev = 'cuda'
s = torch.cuda.current_stream().cuda_stream
alg = {}
# ---- GroupNorm family at C = 128, 16x64x64, B = 8 (134 MB per tensor)
N, V, C, G = 8, 16 * 64 * 64, 128, 1
x = torch.randn(N, V, C, device=dev).bfloat16(); dy = torch.randn(N, V, C, device=dev).bfloat16()
add = torch.randn(N, V, C, device=dev).bfloat16(); out = torch.empty_like(x)
A = torch.randn(N, C, device=dev); Bc = torch.randn(N, C, device=dev)
gamma = torch.randn(C, device=dev); beta = torch.randn(C, device=dev); mr = torch.rand(N, G, 2, device=dev) + 0.5
sums = torch.zeros(N, G, 2, dtype=torch.float64, device=dev); S = torch.zeros(N, C, 2, device=dev)
dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev); cs = torch.zeros(C, device=dev)
T = N * V * C * 2
_lib.call('og_gn_stats', x.data_ptr(), N, V, C, G, sums.data_ptr(), s); alg['og_gn_stats_kernel'] = T
_lib.call('og_gn_act_fwd', x.data_ptr(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, None, 1e-5, G, 1, out.data_ptr(), A.data_ptr(), Bc.data_ptr(), mr.data_ptr(), N, V, C, s); alg['og_gn_act_fwd_kernel'] = 2 * T
A.normal_(); Bc.normal_(); mr.uniform_(0.5, 1.5)
_lib.call('og_affine_act_bwd_reduce', dy.data_ptr(), x.data_ptr(), A.data_ptr(), Bc.data_ptr(), 1, S.data_ptr(), N, V, C, s); alg['og_affine_act_bwd_reduce_kernel'] = 2 * T
_lib.call('og_gn_act_bwd', dy.data_ptr(), x.data_ptr(), A.data_ptr(), Bc.data_ptr(), S.data_ptr(), mr.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, G, 1, add.data_ptr(), out.data_ptr(), dg.data_ptr(), db.data_ptr(), None, None, cs.data_ptr(), N, V, C, s); alg['og_gn_act_bwd_kernel'] = 4 * T
_lib.call('og_colsum', dy.data_ptr(), N * V, C, C, cs.data_ptr(), s); alg['og_colsum_vec_kernel'] = T
del x, dy, add, out
# ---- fused AdamW on 128 M parameters (+ bf16 operand copy): 4 fp32 reads + 3 fp32 writes + 2 B
p = torch.randn(128 << 20, device=dev, requires_grad=True); p.grad = torch.randn_like(p)
opt = og.FusedAdamW([p]); opt.step(); alg['og_adamw_kernel'] = 28 * p.numel()
del p, opt
# ---- LFQ at the tokenizer's shape: N = 2048 tokens, D = 18 (the reference materialises 2.1 GB of logits here)
xq = (torch.randn(2048, 18, device=dev) * 0.5).requires_grad_(True)
o, idx, loss = ops.lfq(xq, 18, 100.0, True, 0.25, 0.1, 1.0)
(loss + o.sum()).backward(); alg['og_lfq_fwd_kernel'] = 2048 * 18 * 14; alg['og_lfq_bwd_kernel'] = 2048 * 18 * 14
# ---- RoPE + LayerNorm and temporal attention at the LatentAction shape (B = 4 clips here: 16 x 64 x 64 x 256)
B, Tt, H, W, Cc = 4, 16, 64, 64, 256
xa = torch.randn(B, Tt, H, W, Cc, device=dev).bfloat16().requires_grad_(True)
freq = torch.rand(Cc // 2, device=dev); g1 = torch.ones(Cc, device=dev, requires_grad=True); b1 = torch.zeros(Cc, device=dev, requires_grad=True)
y = ops.time_attention_res(xa, freq, g1, b1, 4, 4 * 64 ** -0.5)
y.backward(torch.randn_like(y))
rows = B * Tt * H * W
alg["og_rope_ln_fwd_vec_kernel"] = rows * Cc * 4; alg["og_rope_ln_bwd_vec_kernel"] = rows * Cc * 2 * 6
alg['og_temporal_attn_fwd_kernel'] = rows * Cc * 2 * 3; alg['og_temporal_attn_bwd_kernel'] = rows * Cc * 2 * 6
torch.cuda.synchronize()
print(json.dumps({'algorithmic_bytes_per_launch': alg}))
