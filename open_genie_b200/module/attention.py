"""Space-time attention — same class names, constructor kwargs and state_dict keys as the reference's
genie/module/attention.py, executed by the fused B200 kernels (csrc/attention_rows.cu, flash_attn.cu,
conv3d_*.cu).

Only the HEAD-valid configuration is implemented: d_inp == d_out == n_head*d_head, so
to_q / to_k / to_v / to_out are Identity, q = k = v = LayerNorm(RoPE(x)); the one live conditioning path is
the temporal one (latent action -> K, V through `time_attn_kw={'key_dim': k}`). Anything else raises.
state_dict keys: {space,temp}_attn.norm.{weight,bias}, {space,temp}_attn.embed.freq,
temp_attn.to_qkv.to_{k,v}.weight (with key_dim), ffn.1.net.0.{weight,bias}, ffn.1.net.1.0.weight.
"""
from __future__ import annotations

from math import pi
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .. import ops
from ..utils import default, exists
from .video import Conv3dParams


class RotaryEmbedding(nn.Module):
    """Holds the rotary frequencies exactly as the reference does (attention.py:17-46): '1d' =
    1/theta^(2i/dim), '2d' = linspace(1, max_freq/2, dim/2)*pi; `freq` is a (non-trainable) Parameter that
    lives in the state_dict. The rotation itself is fused into the RoPE+LayerNorm kernel."""

    def __init__(self, dim: int, kind: str = '1d', theta=10000, max_freq=10, learned_freq=False) -> None:
        super().__init__()
        if learned_freq:
            raise NotImplementedError('learned rotary frequencies are outside the B200 hot-path scope')
        match kind:
            case '1d':
                freq = 1. / (theta ** (torch.arange(0, dim, 2)[:(dim // 2)].float() / dim))
            case '2d':
                freq = torch.linspace(1., max_freq / 2, dim // 2) * pi
            case _:
                raise NotImplementedError(f"RotaryEmbedding kind {kind!r} is not used on the hot path")
        self.freq = nn.Parameter(freq, requires_grad=False)


class Adapter(nn.Module):
    """Q/K/V adapter (attention.py:105-149): Identity where dims match, Linear(bias=False) for the K/V of a
    conditioning input of width key_dim."""

    def __init__(self, qry_dim: int, n_head: int, d_head: int, key_dim: int | None = None,
                 val_dim: int | None = None, bias: bool = False) -> None:
        super().__init__()
        key_dim = default(key_dim, qry_dim)
        val_dim = default(val_dim, key_dim)
        hid = n_head * d_head
        if qry_dim != hid:
            raise NotImplementedError('d_inp != n_head*d_head does not run in the reference either '
                                      '(LayerNorm dim mismatch, attention.py:179,220)')
        self.to_q = nn.Identity()
        self.to_k = nn.Linear(key_dim, hid, bias=bias) if key_dim != hid else nn.Identity()
        self.to_v = nn.Linear(val_dim, hid, bias=bias) if val_dim != hid else nn.Identity()
        self.n_head = n_head


class Attention(nn.Module):
    """Parameter layout of the reference's Attention (attention.py:154-239)."""
    rope_kind = None
    causal_default = False

    def __init__(self, n_head: int, d_head: int, d_inp: int | None = None, d_out: int | None = None, bias: bool = False,
                 embed: bool = True, scale: float | None = None, causal: bool = False, dropout: float = 0.0,
                 transpose: bool = False, **kwargs) -> None:
        super().__init__()
        hid = n_head * d_head
        self.d_inp = default(d_inp, hid)
        self.d_out = default(d_out, self.d_inp)
        if self.d_inp != hid or self.d_out != hid:
            raise NotImplementedError('only d_inp == d_out == n_head*d_head is valid at the reference HEAD')
        if dropout != 0.0:
            raise NotImplementedError('attention dropout is not used by any shipped blueprint')
        if not embed:
            raise NotImplementedError('embed=False is not used by any shipped blueprint')
        if d_head != 64:
            raise NotImplementedError('the tcgen05 attention kernels are specialised for d_head = 64 '
                                      '(every shipped blueprint uses 64)')
        self.norm = nn.LayerNorm(hid)
        self.embed = RotaryEmbedding(self.d_inp, kind=self.rope_kind)
        self.to_qkv = Adapter(qry_dim=self.d_inp, n_head=n_head, d_head=d_head, bias=bias, **kwargs)
        self.n_head, self.d_head = n_head, d_head
        # reference precedence: default(scale, n_head * d_head ** -0.5)   (attention.py:195)
        self.scale = default(scale, n_head * d_head ** -0.5)
        self.causal = causal
        self.transpose = transpose

    def _rows(self, video: Tensor, transpose: bool | None) -> Tuple[Tensor, bool]:
        t = default(transpose, self.transpose)
        x = video.permute(0, 2, 3, 4, 1) if t else video            # -> (B, T, H, W, C), free for internal tensors
        return x, t

    @staticmethod
    def _back(y: Tensor, t: bool) -> Tensor:
        return y.permute(0, 4, 1, 2, 3) if t else y


class SpatialAttention(Attention):
    """Self-attention over (h w) within each frame, 2-D rotary embedding (attention.py:241-307).
    forward returns attention(video) WITHOUT the skip, like the reference class."""
    rope_kind = '2d'

    def forward(self, video: Tensor, cond: Tensor | None = None, mask: Tensor | None = None,
                transpose: bool | None = None, _residual: bool = False) -> Tensor:
        if exists(cond) or exists(mask):
            raise NotImplementedError('spatial cond / mask are dead code at the reference HEAD (attention.py:290,296)')
        x, t = self._rows(video, transpose)
        y = ops.space_attention_res(x, self.embed.freq, self.norm.weight, self.norm.bias, self.n_head, self.scale,
                                    self.norm.eps)
        if not _residual:
            y = y - ops._rows_bf16(x)          # rarely used stand-alone path
        return self._back(y, t)


class TemporalAttention(Attention):
    """Causal self-attention over t for every pixel, 1-D rotary embedding; optional (B, T, key_dim) conditioning
    that provides K and V (attention.py:309-371)."""
    rope_kind = '1d'

    def forward(self, video: Tensor, cond: Tensor | None = None, mask: Tensor | None = None,
                transpose: bool | None = None, _residual: bool = False) -> Tensor:
        if exists(mask):
            raise NotImplementedError('SDPA forbids attn_mask together with is_causal (attention.py:229-234)')
        x, t = self._rows(video, transpose)
        kc = vc = None
        if exists(cond):
            c = cond.float()
            kc = self.to_qkv.to_k(c)            # (B, T, C): a few kFLOP of host-side plumbing (torch, autograd)
            vc = self.to_qkv.to_v(c)
        y = ops.time_attention_res(x, self.embed.freq, self.norm.weight, self.norm.bias, self.n_head, self.scale,
                                   kc, vc, self.norm.eps)
        if not _residual:
            y = y - ops._rows_bf16(x)
        return self._back(y, t)


class _FfnNet(nn.Module):
    """Mirror of ForwardBlock(in_dim, block=nn.Conv3d, hid_dim=None) -> net = Sequential(GroupNorm,
    Sequential(Conv3d, Identity)) (genie/module/misc.py:71-104) for state_dict keys net.0.*, net.1.0.weight."""

    def __init__(self, dim: int, num_groups: int, kernel_size: int, bias: bool) -> None:
        super().__init__()
        if bias:
            raise NotImplementedError('SpaceTimeAttention(bias=True) is not used by any shipped blueprint')
        self.net = nn.Sequential(
            nn.GroupNorm(num_groups, dim),
            nn.Sequential(Conv3dParams(dim, dim, kernel_size, causal=False, bias=False), nn.Identity()),
        )


class SpaceTimeAttention(nn.Module):
    """x = space(x)+x ; x = time(x, cond)+x ; x = ffn(x)+x  (attention.py:373-474), three fused launches groups."""

    def __init__(self, n_head, d_head, d_inp: int | None = None, d_out: int | None = None, hid_dim=None,
                 bias: bool = False, embed=True, scale: float | None = None, dropout: float = 0.0,
                 kernel_size: int = 3, transpose: bool = False, time_attn_kw: dict = {}, space_attn_kw: dict = {}) -> None:
        super().__init__()
        if isinstance(n_head, int):
            n_head = (n_head, n_head)
        if isinstance(d_head, int):
            d_head = (d_head, d_head)
        if isinstance(embed, bool):
            embed = (embed, embed)
        if exists(hid_dim) or exists(d_inp) or exists(d_out):
            raise NotImplementedError('d_inp / d_out / hid_dim variants do not run at the reference HEAD')
        if n_head[0] * d_head[0] != n_head[1] * d_head[1]:
            raise NotImplementedError('space and time widths must match')
        self.space_attn = SpatialAttention(n_head=n_head[0], d_head=d_head[0], bias=bias, scale=scale, embed=embed[0],
                                           causal=False, dropout=dropout, transpose=transpose, **space_attn_kw)
        self.temp_attn = TemporalAttention(n_head=n_head[1], d_head=d_head[1], bias=bias, scale=scale, embed=embed[1],
                                           causal=True, dropout=dropout, transpose=transpose, **time_attn_kw)
        dim = n_head[1] * d_head[1]
        # nn.Sequential(Rearrange, ForwardBlock, Rearrange): ForwardBlock sits at index 1 -> keys 'ffn.1.net...'
        self.ffn = nn.Sequential(nn.Identity(), _FfnNet(dim, n_head[1], kernel_size, bias), nn.Identity())
        self.in_channels = self.out_channels = dim
        self.transpose = transpose
        self.time_skip = nn.Identity()
        self.space_skip = nn.Identity()
        self.ffn_skip = nn.Identity()

    def forward(self, video: Tensor, cond=None, mask: Tensor | None = None) -> Tensor:
        if not isinstance(cond, tuple):
            cond = (cond, cond)
        space_cond, time_cond = cond
        if exists(space_cond) and self.transpose is not None and not isinstance(space_cond, Tensor):
            space_cond = None
        if exists(mask):
            raise NotImplementedError('mask must stay None (SDPA forbids attn_mask with is_causal)')
        x = video.permute(0, 2, 3, 4, 1) if self.transpose else video
        x = self.space_attn(x, transpose=False, _residual=True)
        # LatentAction passes cond=(None, q_act); a bare tensor cond is the temporal one too
        tc = time_cond if isinstance(self.temp_attn.to_qkv.to_k, nn.Linear) else None
        x = self.temp_attn(x, cond=tc, transpose=False, _residual=True)
        gn = self.ffn[1].net[0]
        conv = self.ffn[1].net[1][0]
        x = ops.ffn_res(x, gn.weight, gn.bias, conv.weight, conv.packed(), conv.geom, gn.num_groups, gn.eps)
        return x.permute(0, 4, 1, 2, 3) if self.transpose else x
