"""CPU, world_size 2 over gloo: the host logic of the data-parallel gradient exchange (bucket planning, skipping of
tensors without a gradient, 1/world scaling, shard split).  The CUDA pack / unpack kernels are replaced by a test-only
packer defined HERE (the product has no CPU path); NCCL is replaced by gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from affganwriting_b200.parallel import GradientReducer, broadcast_module, plan_buckets, shard_batch


class _CpuPacker:                      # test infrastructure only
    def pack(self, grads, bucket):
        o = 0
        for g in grads:
            bucket[o:o + g.numel()].copy_(g.reshape(-1))
            o += g.numel()

    def unpack(self, grads, bucket, scale):
        o = 0
        for g in grads:
            g.copy_((bucket[o:o + g.numel()] * scale).view_as(g))
            o += g.numel()


def test_plan_buckets():
    assert plan_buckets([10, 10, 10], 25) == [[0, 1], [2]]
    assert plan_buckets([100, 1, 1], 25) == [[0], [1, 2]]
    assert plan_buckets([], 25) == []
    assert plan_buckets([5], 1) == [[0]]


def test_shard_batch_requires_equal_shards():
    a, b = torch.arange(8).view(8, 1), torch.arange(8)
    s0, s1 = shard_batch((a, b, "meta"), 0, 2), shard_batch((a, b, "meta"), 1, 2)
    assert s0[0].tolist() == [[0], [1], [2], [3]] and s1[1].tolist() == [4, 5, 6, 7] and s0[2] == "meta"
    with pytest.raises(ValueError):
        shard_batch((torch.arange(7),), 0, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 4), (5,), (2, 2, 2), (7,))]
        for i, p in enumerate(params):
            if i != 1:                                   # params[1] never receives a gradient
                p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        red = GradientReducer(params, bucket_bytes=4 * 13, packer=_CpuPacker())
        nb = red.reduce()
        ok = params[1].grad is None and nb == 3
        for i, p in enumerate(params):
            if i != 1:
                ok = ok and torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))   # mean over ranks of (rank+1)*(i+1)
        # broadcast_module: rank 1 starts from other weights AND other BatchNorm buffers; afterwards both hold rank 0's
        torch.manual_seed(100 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 1))
        net[1].running_mean.normal_()
        broadcast_module(net)
        torch.manual_seed(100)
        want = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 1))
        want[1].running_mean.normal_()
        ok = ok and all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), want.state_dict().values()))
        # equal shards of dim 0 + mean-reduced loss per rank + mean all-reduce == the full-batch gradient;
        # layers without cross-sample coupling only (BatchNorm statistics stay per rank by design)
        lin = torch.nn.Linear(6, 3)
        broadcast_module(lin)
        g = torch.Generator().manual_seed(7)
        xs, ys = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
        x, y = shard_batch((xs, ys), rank, world)
        torch.nn.functional.mse_loss(lin(x), y).backward()
        GradientReducer(lin.parameters(), packer=_CpuPacker()).reduce()
        full = torch.nn.Linear(6, 3)
        full.load_state_dict(lin.state_dict())
        torch.nn.functional.mse_loss(full(xs), ys).backward()
        ok = ok and all(torch.allclose(a.grad, b.grad, atol=1e-6) for a, b in zip(lin.parameters(), full.parameters()))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradient_reducer_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
