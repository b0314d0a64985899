"""2-D image modules of the reference's registry (genie/module/image.py) that the hot path can reach by blueprint
name. A (B, C, H, W) image is handled as an NHWC bf16 tensor, i.e. the T = 1 case of the NDHWC kernels."""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib, ops


class _BlurPool2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k: int, stride: Tuple[int, int], pad: int):
        ops._require_cuda(x, 'blur_pool input')
        B, C, H, W = x.shape
        xi = x.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()            # NHWC rows
        sh, sw = stride
        Ho, Wo = (H + 2 * pad - k) // sh + 1, (W + 2 * pad - k) // sw + 1
        y = torch.empty((B, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
        scratch = torch.empty(B * H * W, dtype=torch.float32, device=x.device)
        _lib.call('og_blurpool2d', xi.data_ptr(), y.data_ptr(), scratch.data_ptr(), 0, B, H, W, C, C, k, sh, sw, pad,
                  ops._stream())
        ctx.cfg = (B, C, H, W, k, sh, sw, pad, Ho, Wo, x.dtype)
        return y.permute(0, 3, 1, 2).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        B, C, H, W, k, sh, sw, pad, Ho, Wo, dtype = ctx.cfg
        dyi = dy.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        dx = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dy.device)
        scratch = torch.empty(B * Ho * Wo, dtype=torch.float32, device=dy.device)
        _lib.call('og_blurpool2d', dyi.data_ptr(), dx.data_ptr(), scratch.data_ptr(), 1, B, H, W, C, C, k, sh, sw, pad,
                  ops._stream())
        return dx.permute(0, 3, 1, 2).to(dtype), None, None, None


class BlurPooling2d(nn.Module):
    """Anti-aliased strided pooling of images — genie/module/image.py:43-85, registry name 'blur_pool'
    (genie/module/__init__.py:33). Same constructor, same 'blur' buffer, same quirky padding `(k-1) // stride`.
    With num_groups == 1 the reference's repeated kernel makes every output channel blur(sum_c x_c);
    that is what og_blurpool2d computes. Returns a (B, C, Ho, Wo) tensor of the input's dtype."""

    def __init__(self, kernel_size, stride=2, num_groups: int = 1, **kwargs) -> None:
        super().__init__()
        ker = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        strd = (stride, stride) if isinstance(stride, int) else tuple(stride)
        if ker[0] != ker[1] or num_groups != 1 or kwargs:
            raise NotImplementedError('BlurPooling2d: square kernels and num_groups == 1 only')
        row = torch.tensor([math.comb(ker[0] - 1, i) for i in range(ker[0])], dtype=torch.float32)
        k2 = row[:, None] * row[None, :]
        self.register_buffer('blur', k2 / k2.sum())
        self.k, self.stride, self.num_groups = ker[0], strd, num_groups
        pad = ((ker[0] - 1) // strd[0], (ker[1] - 1) // strd[1])
        if pad[0] != pad[1]:
            raise NotImplementedError('BlurPooling2d: equal strides only')
        self.padding = pad

    def forward(self, inp: Tensor) -> Tensor:
        if inp.shape[1] % 8 != 0:
            raise NotImplementedError('BlurPooling2d kernels need a channel count that is a multiple of 8')
        return _BlurPool2dFn.apply(inp, self.k, self.stride, self.padding[0])
