"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: the flat gradient bucket used by the
batch-parallel mode reduces to the sum of the per-rank gradients, and batch sharding covers the batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bsms_gnn_b200.dist import GradBucket, shard_batch
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5)),
                  torch.nn.Parameter(torch.randn(2, 2, 2))]
        bucket = GradBucket(params)
        g = torch.Generator().manual_seed(100 + rank)
        grads = [torch.randn(p.shape, generator=g) for p in params]
        for p, gr in zip(params[:2], grads[:2]):  # third parameter got no gradient on this rank
            p.grad = gr.clone()
        bucket.step_sync()
        expect = []
        for k, p in enumerate(params):
            tot = torch.zeros_like(p)
            for r in range(world):
                gg = torch.Generator().manual_seed(100 + r)
                gs = [torch.randn(q_.shape, generator=gg) for q_ in params]
                if k < 2:
                    tot += gs[k]
            expect.append(tot)
        ok = all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(params, expect))
        ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))
        lo, hi = shard_batch(7, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        ok = ok and spans == [(0, 4), (4, 7)]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_grad_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_batch_covers():
    from bsms_gnn_b200.dist import shard_batch
    for n in (1, 7, 48):
        for w in (1, 2, 4, 8):
            spans = [shard_batch(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
