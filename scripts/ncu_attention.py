"""One forward + backward of the tcgen05 attention kernels at the LatentAction shape (S = 4096, d = 64), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_genie_b200 import _lib
nseq, S, nh = int(os.environ.get('NSEQ', 32)), int(os.environ.get('SEQ', 4096)), 4
C = nh * 64
s = torch.cuda.current_stream().cuda_stream
q, k, v, do = (torch.randn(nseq, S, C, device='cuda').mul_(0.5).bfloat16() for _ in range(4))
out = torch.empty_like(q); lse = torch.empty(nseq, nh, S, device='cuda'); delta = torch.empty_like(lse)
dq, dk, dv = (torch.empty_like(q) for _ in range(3))
scale = nh * 64 ** -0.5
for _ in range(2):
    _lib.call('og_flash_attn_fwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), None, None, lse.data_ptr(), nseq, S, C, nh, scale, s)
    _lib.call('og_flash_attn_bwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), nseq, S, C, nh, scale, s)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
_lib.call('og_flash_attn_fwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), None, None, lse.data_ptr(), nseq, S, C, nh, scale, s)
e1.record()
_lib.call('og_flash_attn_bwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), nseq, S, C, nh, scale, s)
e2.record(); torch.cuda.synchronize()
fl = 4.0 * S * S * 64 * nseq * nh
print(f'fwd {e0.elapsed_time(e1):.3f} ms {fl / e0.elapsed_time(e1) * 1e-9:.0f} TFLOP/s ; bwd {e1.elapsed_time(e2):.3f} ms {2.5 * fl / e1.elapsed_time(e2) * 1e-9:.0f} TFLOP/s')
# steady-state timing (5 launches each) and checksums for A/B runs (OG_FLASH_BWD_V1=1 selects the first backward version)
e0.record()
for _ in range(5):
    _lib.call('og_flash_attn_fwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), None, None, lse.data_ptr(), nseq, S, C, nh, scale, s)
e1.record()
for _ in range(5):
    _lib.call('og_flash_attn_bwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), nseq, S, C, nh, scale, s)
e2.record(); torch.cuda.synchronize()
tf, tb = e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / 5
print(f'x5: fwd {tf:.3f} ms {fl / tf * 1e-9:.0f} TFLOP/s ; bwd {tb:.3f} ms {2.5 * fl / tb * 1e-9:.0f} TFLOP/s ; '
      f'|dq| {dq.double().abs().sum().item():.6e} |dk| {dk.double().abs().sum().item():.6e} |dv| {dv.double().abs().sum().item():.6e}')
