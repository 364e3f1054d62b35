"""GPU, >= 2 devices (skipped otherwise): NCCL source-sharded huge-pair mode vs the single-GPU result, bit for bit,
and batch-sharded independence.  One process per GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from rma_net_b200 import parallel
        from rma_net_b200.model.loss import chamfer_dist_with_idx
        g = torch.Generator().manual_seed(4)
        N = M = 40000
        x = (torch.rand(1, N, 3, generator=g) * 2 - 1).to(dev)
        y = (torch.rand(1, M, 3, generator=g) * 2 - 1).to(dev)
        x[:, 30001] = x[:, 7]                                     # cross-shard tie -> lowest global index must win
        xr = x.clone().requires_grad_(True)
        yr = y.clone().requires_grad_(True)
        d1, d2, i1, i2 = chamfer_dist_with_idx(xr, yr)            # single-GPU reference result
        w1 = torch.rand(1, N, generator=g).to(dev)
        w2 = torch.rand(1, M, generator=g).to(dev)
        ((d1 * w1).sum() + (d2 * w2).sum()).backward()
        s, e = parallel.shard_range(N, rank, world)
        xl = x[:, s:e].clone().requires_grad_(True)
        yl = y.clone().requires_grad_(True)
        sd1, sd2, si1, si2 = parallel.SourceShardedChamfer.apply(xl, yl, s, None)
        assert torch.equal(sd1, d1[:, s:e]) and torch.equal(si1, i1[:, s:e])
        assert torch.equal(sd2, d2) and torch.equal(si2, i2)
        ((sd1 * w1[:, s:e]).sum() + (sd2 * w2).sum()).backward()
        gx_ref, gy_ref = xr.grad[:, s:e], yr.grad
        assert (xl.grad - gx_ref).abs().max() <= 1e-5 * gx_ref.abs().max()
        assert (yl.grad - gy_ref).abs().max() <= 1e-5 * gy_ref.abs().max()
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_source_sharded_nccl_matches_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results
