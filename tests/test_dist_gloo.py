"""CPU, world_size = 2, gloo: the N>1 host logic of rma_net_b200/parallel.py (sharding, packed-key all-reduce-MIN
merge, owner masking, grad all-reduce-SUM), with the oracle standing in for the CUDA ops."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle, chamfer_oracle as co
from rma_net_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ---- CPU stand-ins for the CUDA ops (test infrastructure) ------------------------------------------------------
def _cpu_forward(x, y):
    r = c_oracle.chamfer_fwd(x.numpy(), y.numpy(), "f32")
    return (torch.from_numpy(r["dist1"]), torch.from_numpy(r["dist2"]),
            torch.from_numpy(r["idx1"]), torch.from_numpy(r["idx2"]))


def _cpu_backward(x, y, idx1, idx2, g1, g2, need_x=True, need_y=True):
    gx, gy = c_oracle.chamfer_bwd(x.numpy(), y.numpy(), idx1.numpy(), idx2.numpy(), g1.numpy(), g2.numpy())
    return torch.from_numpy(gx).float(), torch.from_numpy(gy).float()


def _cpu_pack(dist_, idx, offset):
    bits = dist_.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return (bits << 32) | (idx.to(torch.int64) + offset)


def _cpu_unpack(keys):
    d = (keys >> 32).to(torch.int32).view(torch.float32)
    return d, (keys & 0xFFFFFFFF).to(torch.int32)


CPU_OPS = parallel.ShardOps(_cpu_forward, _cpu_backward, _cpu_pack, _cpu_unpack)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        B, N, M = 2, 600, 700
        x = torch.rand(B, N, 3, generator=g) * 2 - 1
        y = torch.rand(B, M, 3, generator=g) * 2 - 1
        x[:, 400] = x[:, 100]                      # duplicate source in the OTHER shard: tie -> lowest global index
        full = c_oracle.chamfer_fwd(x.numpy(), y.numpy(), "f32")

        # ---- source-sharded huge-pair mode --------------------------------------------------------------------
        s, e = parallel.shard_range(N, rank, world)
        xl = x[:, s:e].contiguous()
        d1, d2, i1, i2 = parallel.source_sharded_forward(xl, y, s, ops=CPU_OPS)
        assert np.array_equal(d1.numpy(), full["dist1"][:, s:e]) and np.array_equal(i1.numpy(), full["idx1"][:, s:e])
        assert np.array_equal(d2.numpy(), full["dist2"]) and np.array_equal(i2.numpy(), full["idx2"])
        g1 = torch.rand(B, N, generator=g)
        g2 = torch.rand(B, M, generator=g)
        gx, gy = parallel.source_sharded_backward(xl, y, s, i1, i2, g1[:, s:e].contiguous(), g2, ops=CPU_OPS)
        rx, ry = c_oracle.chamfer_bwd(x.numpy(), y.numpy(), full["idx1"], full["idx2"], g1.numpy(), g2.numpy())
        np.testing.assert_allclose(gx.numpy(), rx[:, s:e], atol=1e-5 * np.abs(rx).max())
        np.testing.assert_allclose(gy.numpy(), ry, atol=1e-5 * np.abs(ry).max())

        # ---- batch-sharded mode: no data-path collective; global loss from local sums ---------------------------
        Bg = 5
        xb = torch.rand(Bg, 50, 3, generator=g)
        yb = torch.rand(Bg, 60, 3, generator=g)
        xs, ys = parallel.shard_batch(xb, rank, world), parallel.shard_batch(yb, rank, world)
        r = co.chamfer_exact(xs.numpy(), ys.numpy())
        local_sum = torch.tensor(0.5 * (r["dist1"].sum() + r["dist2"].sum()))
        loss = parallel.global_loss(local_sum, xs.shape[0])
        ref, _, _, _ = co.loss_on_cd_exact(xb.numpy(), yb.numpy())
        assert abs(loss.item() - ref) < 1e-5 * ref
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partition():
    for n in (0, 1, 7, 16, 512, 1000003):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


def test_key_order_matches_distance_then_index():
    d = torch.tensor([0.5, 0.5, 0.25, 0.0, 3e38], dtype=torch.float32)
    i = torch.tensor([7, 3, 9, 2, 1], dtype=torch.int32)
    k = _cpu_pack(d, i, 0)
    order = torch.argsort(k).tolist()
    assert order == [3, 2, 1, 0, 4] and bool((k >= 0).all())
    dd, ii = _cpu_unpack(k)
    assert torch.equal(dd, d) and torch.equal(ii, i)


def test_world2_gloo_sharded_paths():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
