"""Batch-axis data parallelism: one process per GPU, ONE bucketed gradient all-reduce per step over
NCCL (NVLink 5 / NVSwitch), launched bucket by bucket from autograd hooks so it overlaps the rest of
backward. This is what Lightning's `strategy: ddp_find_unused_parameters_false`
(config/tokenize.yaml:76-77) asks for; the reference itself contains no torch.distributed call.

Per-rank semantics preserved from DDP-of-the-reference: GroupNorm statistics are per sample and the LFQ
batch statistics are over the LOCAL batch (genie/module/quantization.py:120) — no extra collectives.
Works with the gloo backend on CPU tensors too (used by the world_size-2 CPU tests).

Two reducers: ArenaGradAllReducer (the hot path: all-reduce in place on the step's zero arena, capturable in the
step's CUDA graph) and GradBucketAllReducer (generic flat buckets for arbitrary modules / no arena)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class GradBucketAllReducer:
    def __init__(self, params, bucket_bytes: int = 256 << 20, process_group=None, average: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        # buckets in REVERSE registration order ~ the order gradients become ready in backward
        self.buckets = []      # list of dict(flat, params[(p, offset, numel)], pending, work)
        cur, cur_n = [], 0
        cap = bucket_bytes // 4
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._close(cur, cur_n)
                cur, cur_n = [], 0
            cur.append((p, cur_n, p.numel()))
            cur_n += p.numel()
        if cur:
            self._close(cur, cur_n)
        self._where = {}
        for b in self.buckets:
            for p, off, n in b['params']:
                self._where[id(p)] = (b, off, n)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _close(self, plist, n):
        dev = plist[0][0].device
        self.buckets.append({'flat': torch.zeros(n, dtype=torch.float32, device=dev), 'params': plist,
                             'pending': len(plist), 'work': None})

    def _view(self, b, off, n, p):
        # same memory order as the parameter (conv weights are channels_last_3d): elementwise-safe aliasing
        return b['flat'].as_strided(p.size(), p.stride(), off) if _dense(p) else b['flat'][off:off + n].view_as(p)

    @torch.no_grad()
    def _on_grad(self, p):
        b, off, n = self._where[id(p)]
        view = self._view(b, off, n, p)
        view.copy_(p.grad)
        p.grad = view
        b['pending'] -= 1
        if b['pending'] == 0 and self.world > 1:
            op = dist.ReduceOp.AVG if (self.average and dist.get_backend(self.pg) == 'nccl') else dist.ReduceOp.SUM
            b['work'] = (dist.all_reduce(b['flat'], op=op, group=self.pg, async_op=True), op)

    def finish(self):
        """Wait for the in-flight bucket reductions (call after backward, before the optimizer step)."""
        if self.world > 1:
            late = [i for i, b in enumerate(self.buckets) if b['pending'] != 0]
            if late:     # DDP(find_unused_parameters=False) raises here too; silently skipping would let ranks diverge
                raise RuntimeError(f'GradBucketAllReducer: {len(late)} bucket(s) did not receive all their gradients in '
                                   f'this backward pass (unused parameters?) — first: bucket {late[0]}')
        for b in self.buckets:
            if b['work'] is not None:
                work, op = b['work']
                work.wait()
                if self.average and op == dist.ReduceOp.SUM:
                    b['flat'].div_(self.world)
                b['work'] = None
            b['pending'] = len(b['params'])

    def grad_bytes(self) -> int:
        return sum(b['flat'].numel() * 4 for b in self.buckets)


class ArenaGradAllReducer:
    """Data-parallel gradient exchange directly on the step's zero arena (ops.ZeroArena).

    Every parameter gradient of a step is carved from the arena in the order backward produces it, so the gradients
    already ARE one contiguous fp32 range: as backward advances, each completed `bucket_bytes` of that range is
    all-reduced in place (NCCL, async, overlapping the rest of backward); `finish()` reduces the tail and waits.
    No flat bucket, no 260 per-parameter copy kernels, and FusedAdamW reads the averaged gradients where they lie.
    Ranks allocate in the same order (same model, same shapes), so the ranges line up across ranks.

    A gradient that autograd produced outside the arena (torch-autograd leaves such as the AdaGN / LFQ projection
    Linears, cloned bias gradients) is moved into it by the hook — a few KB per step. What else sits in the range
    (GroupNorm reduction scratch, already consumed) is averaged along and never read again.
    `average`: NCCL averages inside the collective; other backends (gloo, CPU tests) sum and divide afterwards."""

    def __init__(self, params, arena=None, bucket_bytes: int = 128 << 20, process_group=None, average: bool = True):
        from . import ops
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket = int(bucket_bytes)
        self.average = average
        self.arena = arena
        self._ops = ops
        self._done = None          # (segment, offset) up to which the range has been handed to the collective
        self._works = []
        self._seen = 0
        self._nccl = dist.is_initialized() and dist.get_backend(process_group) == 'nccl'
        self.reduced_bytes = 0     # bytes handed to all-reduce in the last step (bench / tests)
        self._step_bytes = 0
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def bind_arena(self, arena):
        self.arena = arena

    def _arena(self):
        a = self.arena if self.arena is not None else self._ops.current_arena()
        if a is None:
            raise RuntimeError('ArenaGradAllReducer needs the zero arena: call open_genie_b200.enable_zero_arena(True) '
                               'or run the step through GraphedTrainStep')
        return a

    def _reduce(self, seg: torch.Tensor, lo: int, hi: int):
        if hi <= lo or self.world <= 1:
            return
        view = seg[lo:hi].view(torch.float32)
        op = dist.ReduceOp.AVG if (self.average and self._nccl) else dist.ReduceOp.SUM
        self._works.append((dist.all_reduce(view, op=op, group=self.pg, async_op=True), view, op))
        self._step_bytes += hi - lo

    def _flush(self, dev, final: bool):
        a = self._arena()
        start = a.bwd0.get(dev)
        if start is None:
            return
        if self._done is None:
            self._done = start
        si, off = self._done
        csi, coff = a.position(dev)
        segs = a.segs[dev]
        while si < csi:                                    # segments backward has moved past are complete
            self._reduce(segs[si], off, a.hi[dev][si])
            si, off = si + 1, 0
        while coff - off >= self.bucket:
            self._reduce(segs[si], off, off + self.bucket)
            off += self.bucket
        if final and coff > off:
            self._reduce(segs[si], off, coff)
            off = coff
        self._done = (si, off)

    @torch.no_grad()
    def _on_grad(self, p):
        a = self._arena()
        g = p.grad
        if a.locate(g) is None or g.dtype != torch.float32:    # produced outside the arena: move it in
            inside = a.zeros(tuple(p.shape), torch.float32, p.device)
            if _dense(p) and p.stride() != inside.stride():
                inside = inside.view(-1).as_strided(p.size(), p.stride())
            inside.copy_(g)
            p.grad = inside
        self._seen += 1
        self._flush(p.device, final=False)

    def finish(self):
        """Reduce the tail, wait for every in-flight range (call after backward, before the optimizer step)."""
        if self.world > 1 and self._seen != len(self.params):
            missing = len(self.params) - self._seen
            self._seen = 0
            raise RuntimeError(f'ArenaGradAllReducer: {missing} parameter(s) received no gradient in this backward pass; '
                               f'ranks would diverge (freeze them with requires_grad_(False) before building the reducer)')
        devs = {p.device for p in self.params}
        for dev in devs:
            self._flush(dev, final=True)
        for w, view, op in self._works:
            w.wait()
            if self.average and op == dist.ReduceOp.SUM:
                view.div_(self.world)
        self._works = []
        self.reduced_bytes, self._step_bytes = self._step_bytes, 0
        self._done = None
        self._seen = 0

    def grad_bytes(self) -> int:
        return self.reduced_bytes


def _dense(p: torch.Tensor) -> bool:
    """Non-overlapping and dense (any permutation of a contiguous layout)."""
    sizes_strides = sorted(((st, sz) for sz, st in zip(p.size(), p.stride()) if sz > 1))
    expect = 1
    for st, sz in sizes_strides:
        if st != expect:
            return False
        expect *= sz
    return True


def shard_batch(global_batch: int, rank: int, world: int):
    """Clips [lo, hi) of the global batch owned by `rank` (independent units = clips)."""
    per = global_batch // world
    rem = global_batch % world
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)
