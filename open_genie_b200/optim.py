"""FusedAdamW — torch.optim.AdamW semantics (the reference's default optimizer, genie/tokenizer.py:250,437-442)
as ONE multi-tensor CUDA launch per step (csrc/optim.cu), which also refreshes the bf16 operand copies the
conv kernels read so no separate weight-cast pass exists."""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional

import torch
from torch.optim import Optimizer

from . import _lib


class FusedAdamW(Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, bf16_targets: Optional[dict] = None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0                                   # host mirror of the step count (eager launches)
        self._cache_key = None
        self._size_key = None
        self._dev_table = self._dev_ct = self._dev_ci = None
        self._num_chunks = 0
        self.grad_scale: Optional[torch.Tensor] = None   # device scalar multiplied into every gradient
        self._step_dev: Optional[torch.Tensor] = None    # device-side step counter (CUDA-graph safe, authoritative)
        self._lr_dev = {}                                # id(group) -> (device scalar, host value it holds)
        self._pinned = []                                # host staging buffers referenced by captured copies

    # ---- step count / learning rate live on the device so a captured step stays exact -------------------
    def step_count(self) -> int:
        """Number of optimizer steps taken (reads the device counter: replays of a captured step advance it
        without running any Python)."""
        if self._step_dev is not None:
            self._step = int(self._step_dev.item())
        return self._step

    def sync_lr(self):
        """Copy each group's current `lr` (as set by an LR scheduler) into the device scalar the kernel reads.
        Called by step() when launching eagerly and by GraphedTrainStep before every replay; never captured."""
        for group in self.param_groups:
            ent = self._lr_dev.get(id(group))
            if ent is None:
                continue
            t, held = ent
            lr = float(group['lr'])
            if lr != held:
                t.fill_(lr)
                self._lr_dev[id(group)] = (t, lr)

    def state_dict(self):
        """torch.optim.AdamW layout: every state entry carries 'step' (fp32 scalar tensor) next to exp_avg /
        exp_avg_sq, so checkpoints interchange with the reference's optimizer (genie/tokenizer.py:250,437-442)."""
        n = self.step_count()
        for st in self.state.values():
            if st:
                st['step'] = torch.tensor(float(n))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [int(st['step']) for st in self.state.values() if st and 'step' in st]
        self._step = max(steps) if steps else 0
        if self._step_dev is not None:
            self._step_dev.fill_(self._step)
        self._cache_key = None

    def _bf16_target(self, p):
        """Packed bf16 operand (tensor, column offset) of the conv that owns parameter p, if any."""
        from .module.video import CONV_REGISTRY
        m = CONV_REGISTRY.get(id(p))
        if m is None or m.weight is not p:
            return None
        return m.bf16_target()

    @staticmethod
    def _dense_like(p: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        """Gradient with exactly the parameter's memory order (elementwise kernel works on raw storage)."""
        if g.dtype != torch.float32:
            g = g.float()
        if g.stride() == p.stride():
            return g
        out = torch.empty_like(p)          # preserves p's strides
        out.copy_(g)
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():       # like torch.optim: Lightning's closure runs training_step + backward
                loss = closure()
        entries = []
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError('FusedAdamW: parameters must live on a CUDA device (no CPU path)')
                st = self.state[p]
                if not st:
                    st['exp_avg'] = torch.zeros_like(p)      # same strides as p
                    st['exp_avg_sq'] = torch.zeros_like(p)
                g = self._dense_like(p, p.grad)
                if g is not p.grad:
                    p.grad = g
                entries.append((p, g, st['exp_avg'], st['exp_avg_sq'], group))
        if not entries:
            return loss
        self._step += 1
        if self._step_dev is None:
            self._step_dev = torch.full((1,), self._step - 1, dtype=torch.int32, device=entries[0][0].device)
        _lib.call('og_adamw_tick', self._step_dev.data_ptr(), torch.cuda.current_stream().cuda_stream)
        # one launch per param group (hyper-parameters are launch arguments)
        by_group = {}
        for e in entries:
            by_group.setdefault(id(e[4]), []).append(e)
        for ents in by_group.values():
            self._launch(ents)
        from . import ops
        ops.mark_step()                     # gradients consumed: the step-scoped zero arena may be recycled
        return loss

    def _launch(self, ents):
        group = ents[0][4]
        chunk = _lib.load().og_adamw_chunk_elems()
        key = tuple((p.data_ptr(), g.data_ptr()) for p, g, _, _, _ in ents) + tuple(
            (t[0].data_ptr() if t else 0) for t in (self._bf16_target(p) for p, _, _, _, _ in ents))
        if key != self._cache_key or len(self.param_groups) > 1:
            n = len(ents)
            table = (_lib.og_adamw_tensor * n)()
            dev = ents[0][0].device
            size_key = tuple(p.numel() for p, _, _, _, _ in ents)
            if size_key != self._size_key:          # chunk map depends on sizes only: built once
                import numpy as np
                nchs = [(s + chunk - 1) // chunk for s in size_key]
                ct = np.repeat(np.arange(n, dtype=np.int32), nchs)
                ci = np.concatenate([np.arange(c, dtype=np.int32) for c in nchs])
                self._dev_ct = torch.from_numpy(ct).pin_memory().to(dev, non_blocking=True)
                self._dev_ci = torch.from_numpy(ci).pin_memory().to(dev, non_blocking=True)
                self._num_chunks = int(ct.shape[0])
                self._size_key = size_key
            for i, (p, g, m, v, _) in enumerate(ents):
                t = table[i]
                t.p, t.g, t.m, t.v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                t.n = p.numel()
                tgt = self._bf16_target(p)
                if tgt is not None:
                    packed, col_off = tgt[0], tgt[1]
                    t.p_bf16 = packed.data_ptr() + 2 * col_off
                    t.row_len = tgt[2] if len(tgt) > 2 else p.numel() // p.shape[0]
                    t.dst_ld = tgt[3] if len(tgt) > 2 else packed.shape[1]
                else:
                    t.p_bf16, t.row_len, t.dst_ld = None, 1, 1
            raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).pin_memory()
            self._pinned.append(raw)        # keep alive: a captured graph replays this H2D copy
            if len(self._pinned) > 8 and not torch.cuda.is_current_stream_capturing():
                del self._pinned[:-8]
            self._dev_table = raw.to(dev, non_blocking=True)
            self._cache_key = key
        b1, b2 = group['betas']
        stream = torch.cuda.current_stream().cuda_stream
        if id(group) not in self._lr_dev:
            lr0 = float(group['lr'])
            self._lr_dev[id(group)] = (torch.full((1,), lr0, dtype=torch.float32, device=ents[0][0].device), lr0)
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        _lib.call('og_adamw_step', self._dev_table.data_ptr(), self._dev_ct.data_ptr(), self._dev_ci.data_ptr(),
                  self._num_chunks, float(group['lr']), float(b1), float(b2), float(group['eps']),
                  float(group['weight_decay']), self._step, self._step_dev.data_ptr(),
                  self._lr_dev[id(group)][0].data_ptr(),
                  None if self.grad_scale is None else self.grad_scale.data_ptr(), stream)
