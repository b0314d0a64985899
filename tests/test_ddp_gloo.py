"""CPU, world_size 2, gloo: the bucketed gradient all-reduce used for batch-axis data parallelism."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from open_genie_b200.ddp import GradBucketAllReducer, shard_batch
    torch.manual_seed(0)
    w1 = torch.nn.Parameter(torch.randn(8, 4, 1, 3, 3).contiguous(memory_format=torch.channels_last_3d))
    w2 = torch.nn.Parameter(torch.randn(33))
    w3 = torch.nn.Parameter(torch.randn(5, 7))
    red = GradBucketAllReducer([w1, w2, w3], bucket_bytes=256)       # tiny buckets -> several of them
    x = torch.arange(8, dtype=torch.float32)
    lo, hi = shard_batch(8, rank, world)
    loss = sum((p * (1 + x[lo:hi].sum())).sum() for p in (w1, w2, w3))
    loss.backward()
    red.finish()
    want = 1 + 0.5 * (x[:4].sum() + x[4:].sum())                     # mean over the two ranks' scalars
    ok = all(torch.allclose(p.grad, torch.full_like(p, float(want))) for p in (w1, w2, w3))
    ok = ok and w1.grad.stride() == w1.stride() and len(red.buckets) >= 2
    # second step reuses the buckets
    for p in (w1, w2, w3):
        p.grad = None
    loss = sum((p * float(rank + 1)).sum() for p in (w1, w2, w3))
    loss.backward()
    red.finish()
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in (w1, w2, w3))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out[0] and out[1]


class _ArenaGradFn(torch.autograd.Function):
    """Stand-in for the product's backward functions: the parameter gradient is carved from the zero arena."""

    @staticmethod
    def forward(ctx, w, scale):
        ctx.scale = scale
        ctx.shape = tuple(w.shape)
        ctx.strides = tuple(w.stride())
        return (w * scale).sum()

    @staticmethod
    def backward(ctx, g):
        from open_genie_b200 import ops
        gw = ops._zeros(ctx.shape, torch.float32, 'cpu')
        if gw.stride() != ctx.strides:
            gw = gw.view(-1).as_strided(ctx.shape, ctx.strides)
        gw += g * ctx.scale
        return gw, None


def _arena_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from open_genie_b200 import ops
    from open_genie_b200.ddp import ArenaGradAllReducer
    ops.enable_zero_arena(True)
    try:
        torch.manual_seed(0)
        w1 = torch.nn.Parameter(torch.randn(8, 4, 1, 3, 3).contiguous(memory_format=torch.channels_last_3d))
        w2 = torch.nn.Parameter(torch.randn(700))
        w3 = torch.nn.Parameter(torch.randn(5, 7))          # its gradient comes from plain torch autograd (outside the arena)
        red = ArenaGradAllReducer([w1, w2, w3], bucket_bytes=1024)   # tiny buckets -> several in-flight ranges
        ok = True
        for it in range(3):                                  # step 0 grows the arena, step 1 consolidates, step 2 recycles
            s = float(rank + 1 + it)
            loss = _ArenaGradFn.apply(w1, s) + _ArenaGradFn.apply(w2, 2 * s) + (w3 * s).sum()
            loss.backward()
            red.finish()
            mean = sum(float(r + 1 + it) for r in range(world)) / world
            ok = ok and torch.allclose(w1.grad, torch.full_like(w1, mean)) and w1.grad.stride() == w1.stride()
            ok = ok and torch.allclose(w2.grad, torch.full_like(w2, 2 * mean))
            ok = ok and torch.allclose(w3.grad, torch.full_like(w3, mean))
            arena = ops.current_arena()
            ok = ok and all(arena.locate(p.grad) is not None for p in (w1, w2, w3))     # reduced in place, in the arena
            ok = ok and red.grad_bytes() >= 4 * (w1.numel() + w2.numel() + w3.numel())
            for p in (w1, w2, w3):
                p.grad = None
            ops.mark_step()
        # a rank whose backward skips a parameter must fail loudly instead of letting the ranks diverge
        loss = _ArenaGradFn.apply(w1, 1.0)
        loss.backward()
        try:
            red.finish()
            ok = False
        except RuntimeError as e:
            ok = ok and 'received no gradient' in str(e)
        out[rank] = bool(ok)
    finally:
        ops.enable_zero_arena(False)
        dist.destroy_process_group()


def test_arena_allreduce_world2_gloo():
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_arena_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out[0] and out[1]
