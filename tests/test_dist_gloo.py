"""world_size-2 gloo checks of the process-group plumbing (CPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, W, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=W)
    from hmmc_b200 import parallel
    from hmmc_b200.modeling import dist_collect
    try:
        b = 3
        g = torch.Generator().manual_seed(rank)
        x = torch.randn(b, 5, generator=g, requires_grad=True)
        full = dist_collect(x)
        assert full.shape == (W * b, 5)
        for r in range(W):
            ref = torch.randn(b, 5, generator=torch.Generator().manual_seed(r))
            assert torch.equal(full[r * b:(r + 1) * b], ref)
        # every rank evaluates the same global loss; backward = SUM reduce-scatter, so rank r sees
        # W x the gradient of its own slice (the reference's gradient-scale contract)
        coef = torch.arange(W * b * 5, dtype=torch.float32).reshape(W * b, 5)
        (full * coef).sum().backward()
        assert torch.allclose(x.grad, W * coef[rank * b:(rank + 1) * b])
        # rank-dependent upstream gradients are summed per slice
        x2 = torch.ones(b, 5, requires_grad=True)
        f2 = dist_collect(x2)
        (f2 * float(rank + 1)).sum().backward()
        assert torch.allclose(x2.grad, torch.full((b, 5), float(sum(range(1, W + 1)))))
        # a consumer replicated on all ranks (the fine-tune head): the communication-free backward gives
        # the reduce-scatter's result bit for bit
        coef2 = torch.randn(W * b, 5, generator=torch.Generator().manual_seed(7))
        xa = torch.randn(b, 5, generator=torch.Generator().manual_seed(10 + rank), requires_grad=True)
        xb = xa.detach().clone().requires_grad_(True)
        (parallel.all_gather_cat(xa) * coef2).pow(2).sum().backward()
        (parallel.all_gather_cat_replicated(xb) * coef2).pow(2).sum().backward()
        assert torch.equal(xa.grad, xb.grad)
        # key gather, sharding and count merge
        rows = parallel.all_gather_rows(torch.full((2, 4), float(rank)))
        assert rows.shape == (2 * W, 4) and rows[2 * (W - 1), 0] == W - 1
        lo, hi = parallel.shard_range(11)
        assert (lo, hi) == ((0, 6) if rank == 0 else (6, 11))
        cnt = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int32)
        parallel.all_reduce_sum_(cnt)
        assert cnt.tolist() == [3, 30]
        v = torch.arange(hi - lo, dtype=torch.int32) + 100 * rank
        allv = parallel.all_gather_varlen(v, [6, 5])
        assert allv.tolist() == list(range(6)) + [100 + i for i in range(5)]
        # the peer-memory key exchange is for NCCL ranks of one node: on gloo every rank gets None (and agrees)
        assert parallel.same_node()
        assert not parallel.PeerExchange.usable()
        assert parallel.make_peer_exchange(4, 8, torch.device("cpu")) is None
        assert parallel.overlap_group() is dist.group.WORLD
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        ret[rank] = repr(e)
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    W = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(W, port, ret), nprocs=W, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_single_process_is_identity():
    from hmmc_b200 import parallel
    x = torch.randn(4, 3, requires_grad=True)
    y = parallel.all_gather_cat(x)
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    assert parallel.shard_range(10, 4, 3) == (9, 10)
    assert parallel.shard_range(10, 4, 0) == (0, 3)
