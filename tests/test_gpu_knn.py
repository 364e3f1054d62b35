"""GPU parity tests of rma_knn through the C ABI and the reference-named wrappers:
indices equal to the exact (float64, (distance, index)-ordered) oracle except near-ties below fp32 resolution,
distances 1e-5 relative; against the reference's own stored outputs at its expanded-form error bound."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import chamfer_oracle as co

pytestmark = pytest.mark.gpu
CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "knn_*.npz")))


def _near_tie(pts, got_row, ref_row, i):
    p = pts.astype(np.float64)
    d = ((p[i] - p) ** 2).sum(-1)
    bound = 8 * 2.0 ** -24 * 2 * (p ** 2).sum(-1).max()
    a, b = set(got_row.tolist()), set(ref_row.tolist())
    only = list(a ^ b)
    if only:
        kth = max(d[list(a)].max(), d[list(b)].max())
        return all(abs(d[j] - kth) <= bound for j in only)
    return all(abs(d[g_] - d[r_]) <= bound for g_, r_ in zip(got_row, ref_row))


def _check_vs_exact(q, c, k, skip, idx, dist):
    ei, ed = co.knn_exact(q, c, k, skip)
    np.testing.assert_allclose(dist, ed, rtol=1e-5, atol=1e-37)
    bad = np.nonzero((idx != ei).any(-1))
    q64, c64 = q.astype(np.float64), c.astype(np.float64)
    for b, i in zip(*bad):
        d = ((q64[b, i] - c64[b]) ** 2).sum(-1)
        # same distances position by position up to fp32 difference-form resolution (4 ulp)
        assert np.all(np.abs(d[idx[b, i]] - d[ei[b, i]]) <= 4 * 2.0 ** -23 * d[ei[b, i]] + 1e-37), (b, i)
    assert len(bad[0]) <= max(2, idx.shape[0] * idx.shape[1] // 2000)


@pytest.mark.parametrize("case", CASES)
def test_knn_wrappers_against_reference_golden(cuda, golden_dir, case):
    from rma_net_b200.model.loss import get_neighbor_index, indexing_neighbor
    from rma_net_b200.model.network import get_graph_feature, knn
    g = dict(np.load(os.path.join(golden_dir, f"knn_{case}.npz")))
    pts = g["pts"]
    t = torch.from_numpy(pts).to(cuda)
    x = t.transpose(2, 1).contiguous()                                   # [B,3,N] as network.py holds it
    idx5 = knn(x, 5)
    nb4 = get_neighbor_index(t, 4)
    assert idx5.dtype == torch.int64 and idx5.shape == (pts.shape[0], pts.shape[1], 5)
    assert nb4.shape == (pts.shape[0], pts.shape[1], 4)
    _check_vs_exact(pts, pts, 5, 0, idx5.cpu().numpy(), co.knn_exact(pts, pts, 5)[1])
    assert torch.equal(nb4, idx5[..., 1:])
    # vs the reference's own output: identical except near-ties inside its expanded-form error bound
    for got, ref in ((idx5.cpu().numpy(), g["ref_knn5"]), (nb4.cpu().numpy(), g["ref_neighbor4"])):
        for b, i in zip(*np.nonzero((got != ref).any(-1))):
            assert _near_tie(pts[b], got[b, i], ref[b, i], i)
    f = get_graph_feature(x, 5)
    assert f.shape == (pts.shape[0], 6, pts.shape[1], 5)
    assert torch.equal(f[:, 3:, :, 0], x) and torch.equal(f[:, :3, :, 0], x)          # nearest neighbour = self
    nbp = indexing_neighbor(t, nb4)
    assert nbp.shape == (pts.shape[0], pts.shape[1], 4, 3)


@pytest.mark.parametrize("B,N,M,k,skip", [(1, 1, 1, 1, 0), (2, 33, 70, 5, 0), (1, 700, 700, 5, 1), (3, 130, 2500, 8, 0),
                                          (1, 2049, 1025, 16, 2), (2, 300, 64, 32, 0), (1, 50, 7, 7, 0)])
def test_knn_shapes_and_k(cuda, B, N, M, k, skip):
    from rma_net_b200 import ext
    rng = np.random.default_rng(B * 1000 + N + M + k)
    q = rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)
    c = rng.uniform(-1, 1, (B, M, 3)).astype(np.float32)
    idx, dist = ext.knn_indices(torch.from_numpy(q).to(cuda), torch.from_numpy(c).to(cuda), k, skip, want_dist=True)
    _check_vs_exact(q, c, k, skip, idx.cpu().numpy(), dist.cpu().numpy())


def test_knn_ties_duplicates_strides_and_errors(cuda):
    from rma_net_b200 import ext
    base = np.random.default_rng(1).uniform(-1, 1, (1, 200, 3)).astype(np.float32)
    pts = np.concatenate([base, base, base], 1)                            # every point three times
    t = torch.from_numpy(pts).to(cuda)
    idx = ext.knn_indices(t, t, 3).cpu().numpy()
    exp = np.stack([np.arange(200), np.arange(200) + 200, np.arange(200) + 400], -1)
    assert np.array_equal(idx[0], np.tile(exp, (3, 1))), "equal distances -> ascending index"
    # strided views: a batch-strided slice (train_view.py:369-370)
    pair = torch.rand(2, 600, 3, device=cuda)
    a = pair[:, :300]
    ia = ext.knn_indices(a, a, 5)
    assert torch.equal(ia, ext.knn_indices(a.contiguous(), a.contiguous(), 5))
    with pytest.raises(RuntimeError):
        ext.knn_indices(a, a, 0)
    with pytest.raises(RuntimeError):
        ext.knn_indices(a, a, 301)
    with pytest.raises(RuntimeError):
        ext.knn_indices(a.cpu(), a.cpu(), 5)


def test_knn_full_size_properties(cuda):
    """DGCNN / ARAP sizes (B=16, N=8192): too large for the numpy oracle, checked through properties:
    self first at distance 0, distances ascending, every listed neighbour at the listed distance, and no closer
    point outside the list for a sample of rows (brute force on the GPU with torch in float64)."""
    from rma_net_b200 import ext
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(16, 8192, 3, generator=g) * 2 - 1).to(cuda)
    idx, dist = ext.knn_indices(pts, pts, 5, want_dist=True)
    ar = torch.arange(8192, device=cuda, dtype=torch.int32)[None].expand(16, -1)
    assert torch.equal(idx[..., 0], ar) and torch.all(dist[..., 0] == 0)
    assert torch.all(dist[..., 1:] >= dist[..., :-1])
    nb = pts[torch.arange(16, device=cuda).view(-1, 1, 1), idx.long()]
    dchk = ((pts[:, :, None].double() - nb.double()) ** 2).sum(-1)
    assert torch.allclose(dist.double(), dchk, rtol=1e-6, atol=0)
    rows = torch.arange(0, 8192, 97, device=cuda)
    full = ((pts[:, rows, None].double() - pts[:, None].double()) ** 2).sum(-1)         # [16, 85, 8192]
    kth = full.topk(5, dim=-1, largest=False)[0][..., -1]
    assert torch.allclose(dist[:, rows, -1].double(), kth, rtol=1e-6, atol=0)
