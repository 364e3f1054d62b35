"""GPU parity tests: CUDA chamfer path (through the C ABI) vs the oracle, on golden fixtures, seeded synthetic
inputs, edge cases, and — at BASELINE.json's full sizes — the bit-exact fp32 C oracle and size-independent properties.

Bars: argmin indices equal to the fp64 difference-form oracle except documented near-ties
(two candidates whose fp64 distances differ by <= 2^-20 relative — fp32 difference-form resolution is ~3*2^-24);
distances within 1e-5 relative; gradients within 1e-5 of |grad|_inf per tensor.  Against the fp32 C oracle that
uses the kernels' operation order the bar is BIT-exact distances and indices.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, chamfer_oracle as co

pytestmark = pytest.mark.gpu

REL = 1e-5


@pytest.fixture(autouse=True, params=["filter", "exact"])
def chamfer_path(request):
    """Every test runs on both search implementations: the filter + exact-resolve path (default for these sizes) and
    the exact difference-form search (default for huge pairs).  They must agree bit for bit."""
    import torch as _t
    if not _t.cuda.is_available():
        pytest.skip("no CUDA device")
    from rma_net_b200 import ext
    with ext.chamfer_options(path=request.param):      # per-call option of the C ABI (rma_chamfer_opts), no library state
        yield request.param

CASES = sorted(os.path.basename(p)[8:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "chamfer_*.npz")))


def _run(x_np, y_np, dev, grads=True):
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    x = torch.from_numpy(np.ascontiguousarray(x_np)).to(dev).requires_grad_(grads)
    y = torch.from_numpy(np.ascontiguousarray(y_np)).to(dev).requires_grad_(grads)
    d1, d2, i1, i2 = chamfer_dist_with_idx(x, y)
    out = dict(dist1=d1.detach().cpu().numpy(), dist2=d2.detach().cpu().numpy(),
               idx1=i1.cpu().numpy(), idx2=i2.cpu().numpy())
    if grads:
        B = x.shape[0]
        loss = (d1.sum() + d2.sum()) * 0.5 / B
        loss.backward()
        out.update(loss=loss.item(), grad_x=x.grad.cpu().numpy(), grad_y=y.grad.cpu().numpy())
    return out


def _run_fused(x_np, y_np, dev):
    from rma_net_b200.model.loss import loss_on_cd
    x = torch.from_numpy(np.ascontiguousarray(x_np)).to(dev).requires_grad_(True)
    y = torch.from_numpy(np.ascontiguousarray(y_np)).to(dev).requires_grad_(True)
    loss = loss_on_cd(x, y)
    (loss * 1.75).backward()          # non-trivial upstream scalar
    return loss.item(), x.grad.cpu().numpy() / 1.75, y.grad.cpu().numpy() / 1.75


def _check_fused(x, y, dev, got):
    """fused loss_on_cd == unfused composition (same indices -> gradients equal up to fp32 atomics order)."""
    loss, gx, gy = _run_fused(x, y, dev)
    assert abs(loss - got["loss"]) <= 2e-6 * abs(got["loss"]) + 1e-30
    np.testing.assert_allclose(gx, got["grad_x"], rtol=0, atol=2e-6 * np.abs(got["grad_x"]).max() + 1e-30)
    np.testing.assert_allclose(gy, got["grad_y"], rtol=0, atol=2e-6 * np.abs(got["grad_y"]).max() + 1e-30)


def _check_vs_exact(x, y, got, exact=None):
    exact = exact or c_oracle.chamfer_fwd(x, y, "f64")
    np.testing.assert_allclose(got["dist1"], exact["dist1"], rtol=REL, atol=1e-37)
    np.testing.assert_allclose(got["dist2"], exact["dist2"], rtol=REL, atol=1e-37)
    n1, near1, bad1 = co.classify_index_mismatches(x, y, got["idx1"], exact["idx1"])
    n2, near2, bad2 = co.classify_index_mismatches(y, x, got["idx2"], exact["idx2"])
    assert bad1 == 0 and bad2 == 0, f"index mismatches that are not near-ties: {bad1}, {bad2}"
    assert n1 + n2 <= max(2, (got["idx1"].size + got["idx2"].size) // 20000), "too many near-ties to be chance"
    return exact


def _check_bit_exact_f32(x, y, got):
    ref = c_oracle.chamfer_fwd(x, y, "f32")
    assert np.array_equal(got["idx1"], ref["idx1"]) and np.array_equal(got["idx2"], ref["idx2"])
    assert np.array_equal(got["dist1"].view(np.uint32), ref["dist1"].view(np.uint32))
    assert np.array_equal(got["dist2"].view(np.uint32), ref["dist2"].view(np.uint32))


def _check_grads(x, y, got):
    B = x.shape[0]
    # gradients evaluated at the GPU's own indices (identical to the oracle's except at documented near-ties)
    gx, gy = c_oracle.chamfer_bwd(x, y, got["idx1"], got["idx2"], 0.5 / B, 0.5 / B)
    np.testing.assert_allclose(got["grad_x"], gx, rtol=0, atol=REL * np.abs(gx).max() + 1e-30)
    np.testing.assert_allclose(got["grad_y"], gy, rtol=0, atol=REL * np.abs(gy).max() + 1e-30)


@pytest.mark.parametrize("case", CASES)
def test_golden_fixtures(cuda, golden_dir, case):
    g = dict(np.load(os.path.join(golden_dir, f"chamfer_{case}.npz")))
    x, y = g["x"], g["y"]
    got = _run(x, y, cuda)
    _check_vs_exact(x, y, got)
    _check_bit_exact_f32(x, y, got)
    _check_grads(x, y, got)
    _check_fused(x, y, cuda, got)
    # against the reference's own fp32 outputs, at the reference's error bound
    scale = float((x.astype(np.float64) ** 2).sum(-1).max() + (y.astype(np.float64) ** 2).sum(-1).max())
    atol = 8 * 2.0 ** -24 * scale
    np.testing.assert_allclose(got["dist1"], g["ref_dist1"], rtol=0, atol=atol)
    np.testing.assert_allclose(got["dist2"], g["ref_dist2"], rtol=0, atol=atol)
    assert abs(got["loss"] - float(g["ref_loss"])) <= 1e-5 * abs(got["loss"]) + atol
    if np.array_equal(got["idx1"], g["ref_idx1"]) and np.array_equal(got["idx2"], g["ref_idx2"]):
        np.testing.assert_allclose(got["grad_x"], g["ref_grad_x"], rtol=0, atol=3e-6 * np.abs(g["ref_grad_x"]).max())
        np.testing.assert_allclose(got["grad_y"], g["ref_grad_y"], rtol=0, atol=3e-6 * np.abs(g["ref_grad_y"]).max())


def test_survey_known_answers_on_gpu(cuda, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "chamfer_samples1.npz")))
    got = _run(g["x"], g["y"], cuda)
    assert int(got["idx1"].sum()) == 2122746 and int(got["idx2"].sum()) == 2157379
    assert abs(got["dist1"].astype(np.float64).sum() - 8.03275674327) < 1e-5
    assert abs(got["dist2"].astype(np.float64).sum() - 9.98332504104) < 1e-5
    assert abs(np.linalg.norm(got["grad_x"].astype(np.float64)) - 13.2095995) < 2e-4


SHAPES = [(1, 1, 1), (1, 1, 7), (2, 5, 1), (1, 15, 17), (3, 16, 32), (2, 31, 33), (1, 511, 64), (1, 512, 65),
          (2, 513, 100), (1, 2047, 2049), (2, 2048, 640), (1, 2049, 31), (5, 100, 3000), (1, 4096, 4096),
          (3, 1000, 1300), (1, 6000, 200)]


@pytest.mark.parametrize("B,N,M", SHAPES)
@pytest.mark.parametrize("dist", ["uniform", "normal"])
def test_ragged_shapes(cuda, B, N, M, dist):
    g = torch.Generator().manual_seed(1000 + B * 7 + N * 3 + M)
    if dist == "uniform":
        x = (torch.rand(B, N, 3, generator=g) * 2 - 1).numpy()
        y = (torch.rand(B, M, 3, generator=g) * 2 - 1).numpy()
    else:
        x = (torch.randn(B, N, 3, generator=g) * 0.3).numpy()
        y = (torch.randn(B, M, 3, generator=g) * 0.3 + 0.1).numpy()
    got = _run(x, y, cuda)
    _check_vs_exact(x, y, got)
    _check_bit_exact_f32(x, y, got)
    _check_grads(x, y, got)
    _check_fused(x, y, cuda, got)


def test_exact_ties_lowest_index(cuda):
    """torch.min(dim) tie-break: first index wins and gets all the gradient (SURVEY.md §8c tie test)."""
    x = np.zeros((1, 1, 3), np.float32)
    y = np.array([[[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 1, 0]]], np.float32)
    got = _run(x, y, cuda)
    assert got["idx1"][0, 0] == 0 and list(got["idx2"][0]) == [0, 0, 0, 0]
    # duplicated points across tile / row-group / split boundaries
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, (1, 700, 3)).astype(np.float32)
    xs = np.concatenate([base, base, base[:, :300]], 1)            # N = 1700, every point duplicated
    ys = np.concatenate([base[:, ::-1], base, base], 1)            # M = 2100
    got = _run(xs, ys, cuda)
    ref = c_oracle.chamfer_fwd(xs, ys, "f64")
    assert np.array_equal(got["idx1"], ref["idx1"]) and np.array_equal(got["idx2"], ref["idx2"])
    assert np.all(got["dist1"] == 0) and np.all(got["dist2"] == 0)
    _check_grads(xs, ys, got)


def test_strided_views_as_the_callers_pass_them(cuda):
    """train_view.py:369-370 slices [B,2N,3] into source/target (batch stride 6N); loss.py:264 passes
    deformation.transpose(1,2) of a [B,3,N] tensor."""
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    g = torch.Generator().manual_seed(7)
    B, N = 3, 777
    pair = (torch.rand(B, 2 * N, 3, generator=g) * 2 - 1).to(cuda)
    src, tgt = pair[:, :N], pair[:, N:]
    assert not src.is_contiguous()
    d1, d2, i1, i2 = chamfer_dist_with_idx(src, tgt)
    e = c_oracle.chamfer_fwd(src.cpu().numpy(), tgt.cpu().numpy(), "f32")
    assert np.array_equal(i1.cpu().numpy(), e["idx1"]) and np.array_equal(i2.cpu().numpy(), e["idx2"])
    assert np.array_equal(d1.cpu().numpy(), e["dist1"])
    d3n = src.transpose(1, 2).contiguous()                         # [B,3,N] as the network holds it
    d1b, d2b, i1b, i2b = chamfer_dist_with_idx(d3n.transpose(1, 2), tgt)
    assert torch.equal(d1b, d1) and torch.equal(i2b, i2)
    # gradient flows back into the strided leaf layout
    d3n.requires_grad_(True)
    a, b_, _, _ = chamfer_dist_with_idx(d3n.transpose(1, 2), tgt)
    (a.sum() + b_.sum()).backward()
    assert d3n.grad.shape == d3n.shape and torch.isfinite(d3n.grad).all()


def test_metric_call_no_grad_and_argument_errors(cuda):
    """train_view.py:285-286, 379-380: no_grad forward-only calls; bad arguments raise RuntimeError."""
    from rma_net_b200.model.loss import chamfer_dist, loss_on_cd
    x = torch.rand(2, 100, 3, device=cuda)
    y = torch.rand(2, 90, 3, device=cuda)
    with torch.no_grad():
        d1, d2 = chamfer_dist(y, x)
        assert d1.shape == (2, 90) and d2.shape == (2, 100)
        err = (torch.mean(d1 + d2[:, :90]) * 0.5).item()
        assert np.isfinite(err)
    assert loss_on_cd(x, y).dim() == 0
    with pytest.raises(RuntimeError):
        chamfer_dist(x, y[:1])
    with pytest.raises(RuntimeError):
        chamfer_dist(x.double(), y.double())
    with pytest.raises(RuntimeError):
        chamfer_dist(x[:, :, :2], y)
    with pytest.raises(RuntimeError):
        chamfer_dist(x[:, :0], y)
    with pytest.raises(RuntimeError):
        chamfer_dist(x.cpu(), y.cpu())


def test_upstream_gradients_and_partial_requires_grad(cuda):
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    rng = np.random.default_rng(11)
    x = rng.normal(size=(2, 300, 3)).astype(np.float32)
    y = rng.normal(size=(2, 350, 3)).astype(np.float32)
    g1 = rng.normal(size=(2, 300)).astype(np.float32)
    g2 = rng.normal(size=(2, 350)).astype(np.float32)
    xt = torch.from_numpy(x).to(cuda).requires_grad_(True)
    yt = torch.from_numpy(y).to(cuda).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_dist_with_idx(xt, yt)
    (d1 * torch.from_numpy(g1).to(cuda)).sum().backward(retain_graph=True)        # only g1 flows
    gx, gy = c_oracle.chamfer_bwd(x, y, i1.cpu().numpy(), i2.cpu().numpy(), g1, np.zeros_like(g2))
    np.testing.assert_allclose(xt.grad.cpu().numpy(), gx, atol=REL * np.abs(gx).max())
    np.testing.assert_allclose(yt.grad.cpu().numpy(), gy, atol=REL * np.abs(gy).max())
    xt.grad = None; yt.grad = None
    ((d1 * torch.from_numpy(g1).to(cuda)).sum() + (d2 * torch.from_numpy(g2).to(cuda)).sum()).backward()
    gx, gy = c_oracle.chamfer_bwd(x, y, i1.cpu().numpy(), i2.cpu().numpy(), g1, g2)
    np.testing.assert_allclose(xt.grad.cpu().numpy(), gx, atol=REL * np.abs(gx).max())
    np.testing.assert_allclose(yt.grad.cpu().numpy(), gy, atol=REL * np.abs(gy).max())
    # only x requires grad
    xt2 = torch.from_numpy(x).to(cuda).requires_grad_(True)
    d1, d2, _, _ = chamfer_dist_with_idx(xt2, torch.from_numpy(y).to(cuda))
    (d1.sum() + d2.sum()).backward()
    gx, _ = c_oracle.chamfer_bwd(x, y, i1.cpu().numpy(), i2.cpu().numpy(), 1.0, 1.0)
    np.testing.assert_allclose(xt2.grad.cpu().numpy(), gx, atol=REL * np.abs(gx).max())


def test_heavy_scatter_contention(cuda, golden_dir):
    """samples/4: 35 distinct targets are nearest to 1024 sources -> same-address atomics; also a degenerate cloud."""
    x = np.random.default_rng(3).normal(size=(2, 4096, 3)).astype(np.float32)
    y = np.zeros((2, 64, 3), np.float32)
    y[:, :, 0] = 100.0 + np.arange(64)[None] * 1e-3                # everything maps to target 0
    got = _run(x, y, cuda)
    assert np.all(got["idx1"] == 0)
    _check_grads(x, y, got)
    _check_fused(x, y, cuda, got)
    # fused op with only one side requiring grad, and under no_grad
    from rma_net_b200.model.loss import loss_on_cd
    xt = torch.from_numpy(x).to(cuda).requires_grad_(True)
    l = loss_on_cd(xt, torch.from_numpy(y).to(cuda))
    l.backward()
    np.testing.assert_allclose(xt.grad.cpu().numpy(), got["grad_x"], atol=2e-6 * np.abs(got["grad_x"]).max())
    with torch.no_grad():
        assert abs(loss_on_cd(xt, torch.from_numpy(y).to(cuda)).item() - got["loss"]) <= 2e-6 * got["loss"]


def test_sqrdis_map_and_host_entry(cuda):
    import ctypes
    from rma_net_b200 import _lib
    from rma_net_b200.model.loss import compute_sqrdis_map
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (2, 130, 3)).astype(np.float32)
    y = rng.uniform(-1, 1, (2, 70, 3)).astype(np.float32)
    sq = compute_sqrdis_map(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)).cpu().numpy()
    ref = ((x[:, :, None].astype(np.float64) - y[:, None].astype(np.float64)) ** 2).sum(-1)
    np.testing.assert_allclose(sq, ref, rtol=1e-6, atol=1e-12)
    # C-ABI host-buffer entry (what a non-torch binding calls)
    lib = _lib.load()
    loss = ctypes.c_float()
    d1 = np.empty((2, 130), np.float32); d2 = np.empty((2, 70), np.float32)
    i1 = np.empty((2, 130), np.int32); i2 = np.empty((2, 70), np.int32)
    gx = np.empty_like(x); gy = np.empty_like(y)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.rma_chamfer_loss_host(p(x), p(y), 2, 130, 70, ctypes.byref(loss), p(d1), p(d2), p(i1), p(i2), p(gx), p(gy))
    assert rc == 0
    e = c_oracle.chamfer_fwd(x, y, "f32")
    assert np.array_equal(i1, e["idx1"]) and np.array_equal(d2, e["dist2"])
    l64, gx64, gy64, _ = co.loss_on_cd_exact(x, y)
    assert abs(loss.value - l64) < 1e-5 * l64
    np.testing.assert_allclose(gx, gx64, atol=REL * np.abs(gx64).max())


@pytest.mark.parametrize("cfg", ["C2_uniform", "C2_normal", "C3_shape", "C4_slice"])
def test_baseline_sizes_bit_exact_and_properties(cuda, cfg):
    """BASELINE.json sizes (SURVEY.md §8d seeds): bit-exact vs the fp32 C oracle, fp64 oracle bars, and
    size-independent properties: swap symmetry, self-distance, translation of the index set."""
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    if cfg == "C2_uniform":
        g = torch.Generator().manual_seed(0)
        x = torch.rand(16, 4096, 3, generator=g) * 2 - 1
        y = torch.rand(16, 4096, 3, generator=g) * 2 - 1
    elif cfg == "C2_normal":
        g = torch.Generator().manual_seed(1)
        x = torch.randn(16, 4096, 3, generator=g) * 0.3
        y = torch.randn(16, 4096, 3, generator=g) * 0.3
    elif cfg == "C3_shape":
        g = torch.Generator().manual_seed(2)
        x = torch.rand(4, 8192, 3, generator=g) * 2 - 1
        y = torch.rand(4, 8192, 3, generator=g) * 2 - 1
    else:
        g = torch.Generator().manual_seed(3)
        x = torch.rand(64, 2048, 3, generator=g) * 2 - 1
        y = torch.rand(64, 2048, 3, generator=g) * 2 - 1
    xn, yn = x.numpy(), y.numpy()
    got = _run(xn, yn, cuda)
    _check_bit_exact_f32(xn, yn, got)
    _check_vs_exact(xn, yn, got)
    _check_grads(xn, yn, got)
    _check_fused(xn, yn, cuda, got)
    xd, yd = x.to(cuda), y.to(cuda)
    # swap symmetry: chamfer(y, x) is chamfer(x, y) with the outputs exchanged, bit for bit
    d1s, d2s, i1s, i2s = chamfer_dist_with_idx(yd, xd)
    assert np.array_equal(d1s.cpu().numpy(), got["dist2"]) and np.array_equal(d2s.cpu().numpy(), got["dist1"])
    assert np.array_equal(i1s.cpu().numpy(), got["idx2"]) and np.array_equal(i2s.cpu().numpy(), got["idx1"])
    # self distance: zero, identity indices (no duplicate points in these clouds)
    d1, d2, i1, i2 = chamfer_dist_with_idx(xd, xd)
    ar = torch.arange(x.shape[1], device=cuda, dtype=torch.int32)[None].expand(x.shape[0], -1)
    assert torch.all(d1 == 0) and torch.all(d2 == 0) and torch.equal(i1, ar) and torch.equal(i2, ar)
    # permuting the targets permutes the indices (ties aside: none at distance level here)
    perm = torch.randperm(y.shape[1], generator=g).to(cuda)
    d1p, d2p, i1p, i2p = chamfer_dist_with_idx(xd, yd[:, perm])
    assert torch.equal(d1p.cpu(), torch.from_numpy(got["dist1"]))
    assert torch.equal(perm[i1p.long()].cpu().int(), torch.from_numpy(got["idx1"]))


def _path_stats(B, N, M):
    """(runs_filter_path, [extra row candidates, extra column candidates, row-block rescans]) of the last call."""
    from rma_net_b200 import ext
    return 1 if ext.chamfer_path(B, N, M) == "filter" else 0


@pytest.mark.parametrize("case", ["offset", "all_equal", "duplicates", "tiny", "huge", "planar_grid", "two_clusters"])
def test_filter_guard_adversarial(cuda, case, chamfer_path):
    """Inputs built to stress the expanded-form filter: heavy cancellation (clouds far from the origin), massive exact
    ties, duplicated points in different tiles / row groups, extreme scales, lattice points (many equal distances).
    The resolve stage must still return the exact difference-form minimum with the lowest index."""
    rng = np.random.default_rng(42)
    B, N, M = 2, 1500, 1700
    if case == "offset":
        x = (rng.uniform(-1, 1, (B, N, 3)) * 0.05 + 300.0).astype(np.float32)
        y = (rng.uniform(-1, 1, (B, M, 3)) * 0.05 + 300.0).astype(np.float32)
    elif case == "all_equal":
        x = np.full((B, N, 3), 0.37, np.float32)
        y = np.full((B, M, 3), 0.37, np.float32)
        y[:, 5::7] += 1.0
    elif case == "duplicates":
        base = rng.uniform(-1, 1, (B, 300, 3)).astype(np.float32)
        x = np.concatenate([base] * 5, 1)
        y = np.concatenate([base[:, ::-1]] * 5 + [base[:, :200]], 1)
    elif case == "tiny":
        x = (rng.uniform(-1, 1, (B, N, 3)) * 1e-12).astype(np.float32)
        y = (rng.uniform(-1, 1, (B, M, 3)) * 1e-12).astype(np.float32)
    elif case == "huge":
        x = (rng.uniform(-1, 1, (B, N, 3)) * 1e12).astype(np.float32)
        y = (rng.uniform(-1, 1, (B, M, 3)) * 1e12).astype(np.float32)
    elif case == "planar_grid":
        gx_, gy_ = np.meshgrid(np.arange(40), np.arange(40))
        grid = np.stack([gx_.ravel(), gy_.ravel(), np.zeros(1600)], -1).astype(np.float32) * 0.125
        x = np.stack([grid[rng.permutation(1600)][:N] for _ in range(B)])
        y = np.stack([(grid + np.float32(0.0625))[rng.permutation(1600)] for _ in range(B)])
    else:
        c = rng.uniform(-1, 1, (B, 2, 3)) * 50
        x = (c[:, rng.integers(0, 2, N)][0][None] + rng.normal(size=(B, N, 3)) * 1e-3).astype(np.float32)
        y = (c[:, rng.integers(0, 2, M)][0][None] + rng.normal(size=(B, M, 3)) * 1e-3).astype(np.float32)
    got = _run(x, y, cuda)
    _check_bit_exact_f32(x, y, got)
    _check_grads(x, y, got)
    _check_fused(x, y, cuda, got)


def test_paths_agree_and_auto_dispatch(cuda):
    """The automatic choice (filter below the record cap, exact above) is invisible in the results."""
    from rma_net_b200 import ext
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    g = torch.Generator().manual_seed(77)
    x = (torch.rand(3, 2500, 3, generator=g) * 2 - 1).to(cuda)
    y = (torch.rand(3, 1900, 3, generator=g) * 2 - 1).to(cuda)
    outs = []
    for path, cap in ((2, 0), (1, 0), (0, 1 << 30), (0, 1)):         # cap = 1 byte forces the exact path in auto mode
        with ext.chamfer_options(path=path, record_cap_bytes=cap):
            assert ext.chamfer_path(3, 2500, 1900) == ("exact" if path == 1 or cap == 1 else "filter")
            outs.append([t.cpu() for t in chamfer_dist_with_idx(x, y)])
    for o in outs[1:]:
        for a, b_ in zip(outs[0], o):
            assert torch.equal(a, b_)


def test_stage_loop_fused_equals_loop(cuda, chamfer_path):
    """loss.py:260-267 as one batched call (loss_cd_stages_fused) == the per-stage loop, value and gradients."""
    from rma_net_b200.model.loss import loss_cd_stages, loss_cd_stages_fused
    g = torch.Generator().manual_seed(31)
    T, B, N, M = 5, 3, 700, 650
    stages = [(torch.rand(B, 3, N, generator=g) * 2 - 1).to(cuda).requires_grad_(True) for _ in range(T)]
    target = (torch.rand(B, M, 3, generator=g) * 2 - 1).to(cuda).requires_grad_(True)
    ref = loss_cd_stages(stages, target, gamma=0.8, weight_cd=1.3)
    ref.backward()
    gs = [s.grad.clone() for s in stages]
    gt = target.grad.clone()
    for s in stages:
        s.grad = None
    target.grad = None
    got = loss_cd_stages_fused(stages, target, gamma=0.8, weight_cd=1.3)
    got.backward()
    assert abs(got.item() - ref.item()) <= 2e-6 * abs(ref.item())
    for s, r in zip(stages, gs):
        assert (s.grad - r).abs().max() <= 2e-6 * r.abs().max()
    assert (target.grad - gt).abs().max() <= 4e-6 * gt.abs().max()


def test_filter_path_fuzz_against_exact_path(cuda):
    """Randomised stress of the filter + exact-resolve search: 60 seeded configurations across coordinate scales
    (1e-6 .. 1e6), offsets of up to 1000 cloud diameters, anisotropy, clusters, lattices (many exactly equal
    distances), duplicated points and ragged sizes must reproduce the exact difference-form search bit for bit
    (distances AND indices; the exact path itself is pinned to the fp32 C oracle by the tests above)."""
    from rma_net_b200 import ext
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    rng = np.random.default_rng(20260101)
    kinds = ["uniform", "normal", "offset", "aniso", "clusters", "lattice", "dups", "surface"]
    try:
        for trial in range(60):
            kind = kinds[trial % len(kinds)]
            B = int(rng.integers(1, 4))
            N = int(rng.integers(1, 2600))
            M = int(rng.integers(1, 2600))
            scale = float(10.0 ** rng.uniform(-6, 6))

            def cloud(n):
                if kind == "uniform":
                    p = rng.uniform(-1, 1, (B, n, 3))
                elif kind == "normal":
                    p = rng.normal(size=(B, n, 3)) * 0.3
                elif kind == "offset":
                    p = rng.uniform(-1, 1, (B, n, 3)) + rng.uniform(-1000, 1000, (B, 1, 3))
                elif kind == "aniso":
                    p = rng.normal(size=(B, n, 3)) * np.array([1.0, 1e-3, 30.0])
                elif kind == "clusters":
                    c = rng.uniform(-5, 5, (B, 6, 3))
                    p = c[:, rng.integers(0, 6, n)] + rng.normal(size=(B, n, 3)) * 1e-4
                elif kind == "lattice":
                    p = rng.integers(-6, 7, (B, n, 3)).astype(np.float64) * 0.25
                elif kind == "dups":
                    base = rng.uniform(-1, 1, (B, max(1, n // 7), 3))
                    p = base[:, rng.integers(0, base.shape[1], n)]
                else:                                            # points on a sphere: 2-D manifold, near-equal norms
                    v = rng.normal(size=(B, n, 3))
                    p = v / np.linalg.norm(v, axis=-1, keepdims=True)
                return (p * scale).astype(np.float32)
            x, y = cloud(N), cloud(M)
            if kind == "offset":                                 # both clouds around the SAME far-away centre
                y = (y - y.mean(1, keepdims=True) + x.mean(1, keepdims=True)).astype(np.float32)
            xt, yt = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
            outs = []
            for path in (2, 1):
                with ext.chamfer_options(path=path):
                    outs.append([t.cpu() for t in chamfer_dist_with_idx(xt, yt)])
            for a, b_ in zip(*outs):
                assert torch.equal(a, b_), (trial, kind, B, N, M, scale)
    finally:
        pass


@pytest.mark.parametrize("kind", ["uniform", "lattice", "dups", "all_equal"])
def test_filter_long_unit_runs_and_slow_path(cuda, kind):
    """Forced unit runs of 1 .. 200 32-column units per search CTA: segments with several row-direction entries (one per 16 tiles),
    partially filled and empty entries, runs that cross row blocks and batch elements.  Tie-heavy clouds (lattice,
    duplicated points, one repeated point) send most items through the slow path, which re-evaluates whole candidate
    entries; the result must still equal the exact path's bit for bit (distances and indices)."""
    from rma_net_b200 import ext
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    rng = np.random.default_rng(7)
    B, N, M = 2, 3000, 5003                                      # nT = 79 tiles, 2 CTA row blocks per batch element

    def cloud(n):
        if kind == "uniform":
            p = rng.uniform(-1, 1, (B, n, 3))
        elif kind == "lattice":
            p = rng.integers(-5, 6, (B, n, 3)).astype(np.float64) * 0.5
        elif kind == "dups":
            base = rng.uniform(-1, 1, (B, 40, 3))
            p = base[:, rng.integers(0, 40, n)]
        else:
            p = np.tile(np.array([0.3, -0.2, 0.9]), (B, n, 1))
        return p.astype(np.float32)
    xt, yt = torch.from_numpy(cloud(N)).to(cuda), torch.from_numpy(cloud(M)).to(cuda)
    with ext.chamfer_options(path="exact"):
        ref = [t.cpu() for t in chamfer_dist_with_idx(xt, yt)]
    for run in (1, 2, 15, 17, 31, 32, 33, 65, 79, 100, 158, 200, 400):   # odd starts split tiles between segments
        with ext.chamfer_options(path="filter", unit_run=run):
            out = [t.cpu() for t in chamfer_dist_with_idx(xt, yt)]
        for a, b_ in zip(out, ref):
            assert torch.equal(a, b_), (kind, run)


def test_nan_point_does_not_corrupt_anything(cuda, chamfer_path):
    """NaN coordinates are outside the contract, but they must not send an index out of range (the
    gradient scatter would write there): every index stays inside [0, M) / [0, N), and on both paths the NaN point's own
    distance is NaN (the other points skip it as a candidate)."""
    from rma_net_b200.model.loss import chamfer_dist_with_idx, loss_on_cd
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(2, 700, 3, generator=g) * 2 - 1)
    y = (torch.rand(2, 900, 3, generator=g) * 2 - 1)
    x[0, 13, 1] = float("nan")
    y[1, 77, 0] = float("nan")
    xd, yd = x.to(cuda).requires_grad_(True), y.to(cuda).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_dist_with_idx(xd, yd)
    torch.cuda.synchronize()
    assert torch.isnan(d1[0, 13]) and torch.isnan(d2[1, 77])
    assert int(torch.isnan(d1).sum()) == 1 and int(torch.isnan(d2).sum()) == 1
    assert int(i1.min()) >= 0 and int(i1.max()) < 900 and int(i2.min()) >= 0 and int(i2.max()) < 700
    loss_on_cd(xd, yd).backward()
    torch.cuda.synchronize()
    assert xd.grad.shape == x.shape and yd.grad.shape == y.shape


@pytest.mark.parametrize("shape", [(4, 16384, 12000), (1, 65536, 40001)])
@pytest.mark.parametrize("kind", ["uniform", "lattice"])
def test_filter_equals_exact_at_large_sizes(cuda, shape, kind):
    """Sizes at which a search CTA's run (chosen automatically) spans several row-direction entries and the resolve reads
    dozens of entries per row; the lattice cloud (41^3 positions, many duplicates) sends a large share of the items through
    the slow path.  Filter path == exact path bit for bit (tools/parity_large.py runs the longer list)."""
    from rma_net_b200 import ext
    from rma_net_b200.model.loss import chamfer_dist_with_idx
    B, N, M = shape
    rng = np.random.default_rng(11)

    def cloud(n):
        if kind == "uniform":
            p = rng.uniform(-1, 1, (B, n, 3))
        else:
            p = rng.integers(-20, 21, (B, n, 3)).astype(np.float64) * 0.05
        return torch.from_numpy(p.astype(np.float32)).to(cuda)
    x, y = cloud(N), cloud(M)
    outs = []
    for path in (2, 1):
        with ext.chamfer_options(path=path):
            outs.append([t.cpu() for t in chamfer_dist_with_idx(x, y)])
    for a, b_ in zip(*outs):
        assert torch.equal(a, b_)


def test_fused_loss_is_reproducible_and_matches_the_sum(cuda):
    """The fused loss is accumulated in double and published as the largest rounded prefix sum (csrc/chamfer_common.cuh,
    loss_commit), so repeated runs on the same inputs agree bit for bit and the value equals the fp64 sum of the returned
    distances to fp32 rounding; the metric variant returns exactly the distances chamfer_dist returns."""
    from rma_net_b200.model.exfunc.functions import ChamferCDMetricsFunction
    from rma_net_b200.model.loss import chamfer_dist, loss_on_cd
    g = torch.Generator().manual_seed(21)
    B, N, M = 16, 4096, 4096
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(cuda)
    y = (torch.rand(B, M, 3, generator=g) * 2 - 1).to(cuda)
    vals = {float(loss_on_cd(x, y)) for _ in range(8)}
    assert len(vals) == 1, vals
    d1, d2 = chamfer_dist(x, y)
    ref = (d1.double().sum() + d2.double().sum()).item() * 0.5 / B
    assert abs(vals.pop() - ref) <= 1.2e-7 * abs(ref)
    loss, m1, m2 = ChamferCDMetricsFunction.apply(x, y)
    assert torch.equal(m1, d1) and torch.equal(m2, d2)


def test_fused_loss_weight_contract(cuda):
    """include/rma_b200.h: coef and the per-batch weights of the fused loss are non-negative.  Zero weights are fine, a
    negative weight poisons the value (NaN) instead of returning a wrong prefix, a negative coef is a bad argument; NaN
    coordinates give a NaN loss as the reference's mean does."""
    from rma_net_b200 import ext
    g = torch.Generator().manual_seed(22)
    B, N, M = 4, 700, 900
    x = torch.randn(B, N, 3, generator=g).to(cuda)
    y = torch.randn(B, M, 3, generator=g).to(cuda)
    w = torch.tensor([0.0, 1.0, 0.5, 2.0], device=cuda)
    loss, _, _, d1, d2 = ext.chamfer_cd_forward(x, y, False, False, w, 0.25, True)
    ref = 0.25 * (w.double() * (d1.double().sum(1) + d2.double().sum(1))).sum().item()
    assert abs(float(loss) - ref) <= 1.2e-7 * abs(ref)
    zero = ext.chamfer_cd_forward(x, y, False, False, torch.zeros(B, device=cuda), 0.25)[0]
    assert float(zero) == 0.0
    neg = ext.chamfer_cd_forward(x, y, False, False, torch.tensor([1.0, -1.0, 1.0, 1.0], device=cuda), 0.25)[0]
    assert torch.isnan(neg)
    with pytest.raises(RuntimeError):
        ext.chamfer_cd_forward(x, y, False, False, w, -0.25)
    xn = x.clone()
    xn[1, 5, 0] = float("nan")
    assert torch.isnan(ext.chamfer_cd_forward(xn, y, False, False, w, 0.25)[0])


def test_alternating_inputs_through_one_workspace(cuda):
    """The resolve reads its item's pack entry before its dependency wait (csrc/chamfer_filter.cu): results must follow
    the inputs when two different pairs alternate through the same workspace and output buffers, step after step."""
    from rma_net_b200.model.loss import chamfer_dist_with_idx, loss_on_cd
    g = torch.Generator().manual_seed(33)
    B, N, M = 8, 2048, 1536
    pairs = [((torch.rand(B, N, 3, generator=g) * 2 - 1).to(cuda), (torch.rand(B, M, 3, generator=g) * 2 - 1).to(cuda))
             for _ in range(2)]
    ref = [[t.clone() for t in chamfer_dist_with_idx(x, y)] for x, y in pairs]
    ref_loss = [float(loss_on_cd(x, y)) for x, y in pairs]
    assert not torch.equal(ref[0][2], ref[1][2])
    graphs = []
    outs = []
    for x, y in pairs:                    # the caching allocator hands both captures the same workspace block
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            chamfer_dist_with_idx(x, y)
        torch.cuda.synchronize()
    pool = torch.cuda.graph_pool_handle()
    for x, y in pairs:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, pool=pool):
            o = chamfer_dist_with_idx(x, y)
            l = loss_on_cd(x, y)
        graphs.append(gr)
        outs.append((o, l))
    for it in range(40):
        k = it & 1
        graphs[k].replay()
        graphs[k ^ 1].replay()
        graphs[k].replay()
        torch.cuda.synchronize()
        o, l = outs[k]
        for a, b_ in zip(o, ref[k]):
            assert torch.equal(a, b_), it
        assert float(l) == ref_loss[k]


def test_shared_target_period_equals_repeated_target(cuda, chamfer_path):
    """rma_chamfer_opts.y_batch_period: batch element b reads y[b % period]; everything the call returns is bit-identical
    to the call on the target repeated T times, and a batch that the period does not divide is a bad argument."""
    from rma_net_b200 import ext
    g = torch.Generator().manual_seed(41)
    T, B, N, M = 3, 4, 1100, 900
    x = (torch.rand(T * B, N, 3, generator=g) * 2 - 1).to(cuda)
    y = (torch.rand(B, M, 3, generator=g) * 2 - 1).to(cuda)
    w = torch.rand(T * B, generator=g).to(cuda)
    rep = y.repeat(T, 1, 1)
    a = ext.chamfer_cd_forward(x, rep, True, True, w, 0.125, True)
    b_ = ext.chamfer_cd_forward(x, y, True, True, w, 0.125, True, y_period=B)
    assert float(a[0]) == float(b_[0])
    for pa, pb in zip((a[1][0], a[2][0]) + a[3:], (b_[1][0], b_[2][0]) + b_[3:]):   # own terms, dist1, dist2
        assert torch.equal(pa, pb)
    for pa, pb in zip((a[1][1], a[2][1]), (b_[1][1], b_[2][1])):                    # scattered terms: float atomics,
        assert (pa - pb).abs().max() <= 1e-6 * pa.abs().max()                        # the order of the adds is not fixed
    with pytest.raises(RuntimeError):
        ext.chamfer_cd_forward(x[:-1], y, True, False, None, None, False, y_period=B)
