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
