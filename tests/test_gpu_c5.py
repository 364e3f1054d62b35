"""GPU parity at BASELINE.json configs[4] size (one pair, N = M = 2^20 = 1.1e12 point pairs):
filter path == exact path bit for bit, sampled rows / columns against the fp64 C oracle, and the source-sharded
merge (packed (distance, index) keys, MIN) of parallel.py emulated shard by shard on one GPU against the unsharded
result.  (The NCCL all-reduce itself runs in tests/test_gpu_multi.py on >= 2 GPUs and in bench.py's `secondary.C5`
on every multi-GPU run.)"""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu
N = M = 1 << 20


@pytest.fixture(scope="module")
def clouds(cuda):
    g = torch.Generator().manual_seed(4)                      # the seed of SURVEY.md section 8d for C5
    x = torch.rand(1, N, 3, generator=g) * 2 - 1
    y = torch.rand(1, M, 3, generator=g) * 2 - 1
    return x, y, x.to(cuda), y.to(cuda)


@pytest.fixture(scope="module")
def exact_result(clouds):
    from rma_net_b200 import ext
    _, _, xd, yd = clouds
    with ext.chamfer_options(path="exact"):
        assert ext.chamfer_path(1, N, M) == "exact"
        return ext.chamfer_forward(xd, yd)


def test_c5_filter_equals_exact(clouds, exact_result):
    """2^40 pairs through both search implementations: identical distances and indices.  The filter's group records are
    B*N*M/64 = 16 GiB here, exactly the default cap: the automatic choice is the filter path (40 GiB workspace); a
    lower per-call cap sends the same call to the exact path."""
    from rma_net_b200 import ext
    _, _, xd, yd = clouds
    with ext.chamfer_options(path="auto", record_cap_bytes=4 << 30):
        assert ext.chamfer_path(1, N, M) == "exact"
    assert ext.chamfer_path(1, N, M) == "filter"              # automatic choice at this size on one GPU
    got = ext.chamfer_forward(xd, yd)
    for a, b in zip(got, exact_result):
        assert torch.equal(a, b)


def test_c5_sampled_rows_and_columns_vs_fp64_oracle(clouds, exact_result):
    """256 sampled sources against all 2^20 targets and 256 sampled targets against all sources in the fp64 C oracle:
    distances within 1e-5 relative, indices equal except documented near-ties (fp64 distances within 2^-20)."""
    x, y, _, _ = clouds
    d1, d2, i1, i2 = [t.cpu().numpy()[0] for t in exact_result]
    rng = np.random.default_rng(5)
    rows, cols = rng.choice(N, 256, replace=False), rng.choice(M, 256, replace=False)
    er = c_oracle.chamfer_fwd(x[:, rows].numpy(), y.numpy(), "f64")
    ec = c_oracle.chamfer_fwd(x.numpy(), y[:, cols].numpy(), "f64")
    x64, y64 = x[0].double().numpy(), y[0].double().numpy()
    for got_d, got_i, ref_d, ref_i, own, other in (
            (d1[rows], i1[rows], er["dist1"][0], er["idx1"][0], x64[rows], y64),
            (d2[cols], i2[cols], ec["dist2"][0], ec["idx2"][0], y64[cols], x64)):
        np.testing.assert_allclose(got_d, ref_d, rtol=1e-5, atol=0)
        d_at = ((own - other[got_i]) ** 2).sum(-1)                # fp64 distance to the index the GPU returned
        bad = (got_i != ref_i) & (np.abs(d_at - ref_d) > 2.0 ** -20 * ref_d)
        assert not bad.any()
        assert got_i.min() >= 0 and got_i.max() < N


@pytest.mark.parametrize("world", [4, 8])
def test_c5_source_sharded_merge_emulated(clouds, exact_result, world):
    """parallel.py's huge-pair mode with the ranks emulated one after the other on this GPU: every shard searches its
    sources against the replicated target, packs (distance, global index) keys, the keys are merged with an elementwise
    MIN (what ncclAllReduce(ncclMin) computes) - the result equals the unsharded search bit for bit."""
    from rma_net_b200 import ext, parallel
    _, _, xd, yd = clouds
    d1, d2, i1, i2 = exact_result
    merged = None
    for rank in range(world):
        s, e = parallel.shard_range(N, rank, world)
        ld1, ld2, li1, li2 = ext.chamfer_forward(xd[:, s:e], yd)          # strided slice, automatic path (filter)
        assert torch.equal(ld1, d1[:, s:e]) and torch.equal(li1, i1[:, s:e])
        keys = ext.pack_min_keys(ld2, li2, s)
        merged = keys if merged is None else torch.minimum(merged, keys)
    md2, mi2 = ext.unpack_min_keys(merged)
    assert torch.equal(md2, d2) and torch.equal(mi2, i2)
    owned, local = parallel.owned_winners(mi2, *parallel.shard_range(N, world - 1, world)[:1],
                                          N - parallel.shard_range(N, world - 1, world)[0])
    assert int(owned.sum()) == int((mi2 >= parallel.shard_range(N, world - 1, world)[0]).sum())
