"""DynamicsModel — mirrors genie/dynamics.py:14-194 (MaskGIT dynamics over space-time transformer blocks) on the
B200 kernels: fused token+action embedding, SpaceTimeAttention blocks, the vocabulary head on the tcgen05 GEMM
kernel, a fused masked cross entropy. Same constructor, forward / compute_loss / generate / get_schedule
signatures, return values and state_dict keys (tok_emb.weight, act_emb.0.weight, head.{weight,bias},
dec_layers.N.*)."""
from __future__ import annotations

from math import inf, pi, prod
from typing import Literal

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .module import parse_blueprint
from .module.linear import LinearRows
from .utils import Blueprint, default


class DynamicsModel(nn.Module):
    def __init__(self, desc: Blueprint, tok_vocab: int, act_vocab: int, embed_dim: int) -> None:
        super().__init__()
        self.dec_layers, self.ext_kw = parse_blueprint(desc)
        self.head = LinearRows(embed_dim, tok_vocab)
        self.tok_emb = nn.Embedding(tok_vocab, embed_dim)
        self.act_emb = nn.Sequential(nn.Embedding(act_vocab, embed_dim), nn.Identity())
        self.tok_vocab, self.act_vocab, self.embed_dim = tok_vocab, act_vocab, embed_dim

    def _logits(self, tokens: Tensor, act_id: Tensor) -> Tensor:
        x = ops.embed_add(tokens, act_id, self.tok_emb.weight, self.act_emb[0].weight)      # (B,T,H,W,C) bf16
        for dec in self.dec_layers:
            x = dec(x)
        return self.head(x)                                                                  # (B,T,H,W,V) bf16

    def forward(self, tokens: Tensor, act_id: Tensor):
        """genie/dynamics.py:44-64: returns (logits (B,T,H,W,V), logits[:, -1])."""
        logits = self._logits(tokens, act_id).float()
        return logits, logits[:, -1]

    def compute_loss(self, tokens: Tensor, act_id: Tensor, mask: Tensor | None = None, fill: float = 0.) -> Tensor:
        """genie/dynamics.py:66-99 — including its quirk that the target is read from the ALREADY-masked tokens."""
        b, t, h, w = tokens.shape
        if mask is None:
            rate = torch.empty(1).uniform_(0.5, 1).item()
            mask = torch.distributions.Bernoulli(rate).sample((b, t, h, w)).bool()
        mask = mask.to(tokens.device)
        tokens = torch.masked_fill(tokens, mask, fill)
        logits = self._logits(tokens, act_id.detach())
        return ops.masked_cross_entropy(logits.reshape(-1, logits.shape[-1]), tokens, mask)

    @torch.no_grad()
    def generate(self, tokens: Tensor, act_id: Tensor, steps: int = 10,
                 which: Literal['linear', 'cosine', 'arccos'] = 'linear', temp: float = 1., topk: int = 50,
                 masked_tok: int = 0) -> Tensor:
        """MaskGIT iterative sampling — genie/dynamics.py:101-165. The transformer evaluations run on the B200
        kernels; the per-step sampling bookkeeping (softmax / multinomial / top-k over h*w tokens) is host-side
        torch plumbing exactly as in the reference."""
        b, t, h, w = tokens.shape
        schedule = self.get_schedule(steps, shape=(h, w), which=which)
        mask = torch.ones(b, h, w, dtype=torch.bool, device=tokens.device)
        code = torch.full((b, h, w), masked_tok, device=tokens.device, dtype=tokens.dtype)
        mock = torch.zeros(b, 1, dtype=act_id.dtype, device=tokens.device)
        act_all = torch.cat([act_id, mock], dim=1)
        pred_tok = torch.cat([tokens, code[:, None]], dim=1)
        for num_tokens in schedule.tolist():
            if mask.sum() == 0:
                break
            tok_id = torch.cat([tokens, code[:, None]], dim=1)
            _, logits = self(tok_id, act_all)
            prob = torch.softmax(logits / temp, dim=-1).reshape(-1, logits.shape[-1])
            pred = torch.multinomial(prob, num_samples=1)
            conf = torch.gather(prob, -1, pred).reshape(b, h, w)
            conf[~mask] = -inf
            idxs = torch.topk(conf.view(b, -1), k=num_tokens, dim=-1).indices
            pred = pred.view(b, -1)
            code = code.view(b, -1).scatter(1, idxs, torch.gather(pred, -1, idxs).to(code.dtype)).view(b, h, w)
            mask = mask.view(b, -1).scatter(1, idxs, False).view(b, h, w)
            pred_tok = torch.cat([tokens, code[:, None]], dim=1)
        assert mask.sum() == 0, f'Not all tokens were predicted. {mask.sum()} tokens left.'
        return pred_tok

    def get_schedule(self, steps: int, shape: tuple[int, int],
                     which: Literal['linear', 'cosine', 'arccos'] = 'linear') -> Tensor:
        """genie/dynamics.py:167-194."""
        n = prod(shape)
        t = torch.linspace(1, 0, steps)
        match which:
            case 'linear':
                s = 1 - t
            case 'cosine':
                s = torch.cos(t * pi * .5)
            case 'arccos':
                s = torch.acos(t) / (pi * .5)
            case _:
                raise ValueError(f'Unknown schedule type: {which}')
        schedule = (s / s.sum()) * n
        schedule = schedule.round().int().clamp(min=1)
        schedule[-1] += n - schedule.sum()
        return schedule
