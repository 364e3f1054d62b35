"""GPU parity tests of the point_render / point_masker kernels (the reference's model/exfunc extension, SURVEY.md
section 8f rank 1), through the C ABI:
  * against the committed outputs of the REFERENCE's own CUDA kernels (tests/golden/exfunc_*.npz,
    oracle/gen_golden_exfunc.py),
  * against the reference kernels themselves when oracle/_ref/libexfunc_ref.so travelled to the box, at full size,
  * against the numpy restatement (oracle/exfunc_oracle.py).
Bars: pixel_index / mask / render backward (given the same is_visible): bit-exact; weight_points: 3e-7 relative
(expf); depth / weight images and the mask backward (atomicAdd order is unspecified in the reference): 2e-6 of the
tensor's max magnitude (1e-5 for the mask backward: fp32 sums of >1000 terms).  front / back / is_visible: exact against the intended semantics (numpy oracle); against
the reference only one-sided, because its float atomics lose updates under contention (oracle/exfunc_oracle.py
header): reference front >= ours, reference back <= ours, and visibility may differ only at such pixels."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import exfunc_oracle as eo, exfunc_ref

pytestmark = pytest.mark.gpu
CASES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "exfunc_*.npz")))


def _ours_render(proj, H, W, vpw, cpw, eps, thr, gd, gw, dev):
    from rma_net_b200 import ext
    pd = torch.from_numpy(proj).to(dev)
    pix, vis, wp, depth, weight = ext.point_render_forward(pd, H, W, vpw, cpw, eps, thr, want_weight_points=True)
    gp = ext.point_render_backward(torch.from_numpy(gd).to(dev), torch.from_numpy(gw).to(dev), pd, pix, vis, cpw)[0]
    return dict(pixel_index=pix.cpu().numpy(), is_visible=vis.cpu().numpy(), weight_points=wp.cpu().numpy(),
                depth_image=depth.cpu().numpy(), weight_image=weight.cpu().numpy(), grad_proj=gp.cpu().numpy())


def _close_sum(a, b, what, rel=2e-6):
    scale = max(float(np.abs(b).max()), 1e-30)
    assert np.abs(a.astype(np.float64) - b.astype(np.float64)).max() <= rel * scale, what


def _check_render_vs_reference(got, ref, proj, H, W, cpw, front=None, back=None):
    """`ref` = arrays produced by the reference kernels (golden file or live)."""
    B, N = proj.shape[:2]
    assert np.array_equal(got["pixel_index"], ref["pixel_index"])
    pix = ref["pixel_index"].reshape(B, N, 2)
    if front is not None:
        rf, rb = ref["front"].reshape(B, H, W), ref["back"].reshape(B, H, W)
        assert np.all(rf >= front) and np.all(rb <= back), "reference min/max can only MISS updates"
    diff = got["is_visible"] != ref["is_visible"]
    assert diff.sum() <= max(3, diff.size // 50), f"{int(diff.sum())} visibility flags differ"
    if front is not None and diff.any():
        bi, ni = np.nonzero(diff)
        y, x = pix[bi, ni, 1], pix[bi, ni, 0]
        assert np.all((rf[bi, y, x] != front[bi, y, x]) | (rb[bi, y, x] != back[bi, y, x])), \
            "visibility differs where the reference did not lose an update"
    same = ~diff
    np.testing.assert_allclose(got["weight_points"][same], ref["weight_points"][same], rtol=3e-7, atol=1e-37)
    # images: compare away from the colour patches of the points whose visibility differs
    keep = np.ones((B, H, W), bool)
    r = (cpw - 1) // 2
    for b_, n_ in zip(*np.nonzero(diff)):
        x, y = pix[b_, n_]
        keep[b_, max(0, y - r - 1):y + r + 2, max(0, x - r - 1):x + r + 2] = False
    for name in ("depth_image", "weight_image"):
        a, b = got[name].reshape(B, H, W), ref[name].reshape(B, H, W)
        assert np.abs(a - b)[keep].max() <= 2e-6 * np.abs(b).max(), name


@pytest.mark.parametrize("case", CASES)
def test_render_against_reference_golden(cuda, golden_dir, case):
    from rma_net_b200 import ext
    g = dict(np.load(os.path.join(golden_dir, f"exfunc_{case}.npz")))
    B, N, H, W, vpw, cpw, mpw = (int(v) for v in g["params"])
    eps, thr, thr2 = (float(v) for v in g["fparams"])
    got = _ours_render(g["proj"], H, W, vpw, cpw, eps, thr, g["grad_depth"], g["grad_weight"], cuda)
    o = eo.render_forward(g["proj"], H, W, vpw, cpw, eps, thr)
    # intended semantics: exact
    assert np.array_equal(got["pixel_index"].reshape(B, N, 2), o["pixel_index"])
    assert np.array_equal(got["is_visible"], o["is_visible"])
    _close_sum(got["depth_image"][..., 0], o["depth_image"], "depth vs oracle")
    _close_sum(got["weight_image"][..., 0], o["weight_image"], "weight vs oracle")
    # the reference's own output
    _check_render_vs_reference(got, g, g["proj"], H, W, cpw, o["front"], o["back"])
    # backward, decoupled from the visibility race: feed the reference's saved tensors -> its gradient bit for bit
    pd = torch.from_numpy(g["proj"]).to(cuda)
    gp = ext.point_render_backward(torch.from_numpy(g["grad_depth"]).to(cuda), torch.from_numpy(g["grad_weight"]).to(cuda),
                                   pd, torch.from_numpy(g["pixel_index"]).to(cuda),
                                   torch.from_numpy(g["is_visible"]).to(cuda), cpw)[0].cpu().numpy()
    assert np.array_equal(gp.view(np.uint32), g["grad_proj"].view(np.uint32)), "render backward bits"


@pytest.mark.parametrize("case", CASES)
def test_masker_against_reference_golden(cuda, golden_dir, case):
    from rma_net_b200 import ext
    g = dict(np.load(os.path.join(golden_dir, f"exfunc_{case}.npz")))
    B, N, H, W, vpw, cpw, mpw = (int(v) for v in g["params"])
    eps, thr, thr2 = (float(v) for v in g["fparams"])
    pd = torch.from_numpy(g["proj"]).to(cuda)
    mask = ext.point_masker_forward(pd, H, W, mpw)[0]
    assert np.array_equal(mask.cpu().numpy(), g["mask"])
    gm = ext.point_masker_backward(torch.from_numpy(g["grad_mask"]).to(cuda), pd, eps, thr2)[0].cpu().numpy()
    assert np.all(gm[..., :2] == 0)
    _close_sum(gm, g["grad_proj_mask"], "mask backward", rel=1e-5)   # fp32 sums of >1000 terms, order differs
    # deterministic: two runs agree bit for bit (the reference's atomics do not guarantee that)
    gm2 = ext.point_masker_backward(torch.from_numpy(g["grad_mask"]).to(cuda), pd, eps, thr2)[0].cpu().numpy()
    assert np.array_equal(gm, gm2)


def _cloud(B, N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    p = (torch.randn(B, N, 3, generator=g) * 0.35).clamp(-0.99, 0.99)
    xy = (p[..., :2] + 1.0) * (torch.tensor([W, H]) - 1) / 2
    return torch.cat([xy, p[..., 2:] - 1.0], -1).contiguous()


@pytest.mark.parametrize("B,N,res", [(4, 2048, 256), (2, 8192, 256), (3, 777, 100)])
def test_full_size_against_reference_kernels(cuda, B, N, res):
    """loss.py:33-37 configuration: 256x256, Render(7, 9, 1e-5), threshold 0.1, Masker(5), threshold 1e-10."""
    if not exfunc_ref.available():
        pytest.skip("oracle/_ref/libexfunc_ref.so not built (needs /root/reference at build time)")
    from rma_net_b200 import ext
    proj = _cloud(B, N, res, res, 5)
    pd = proj.to(cuda)
    ref = exfunc_ref.render_forward(pd, res, res, 7, 9, 1e-5, 0.1)
    g = torch.Generator().manual_seed(6)
    gd = torch.randn(B, res, res, 1, generator=g).to(cuda)
    gw = torch.randn(B, res, res, 1, generator=g).to(cuda)
    ref_gp = exfunc_ref.render_backward(gd, gw, pd, ref["pixel_index"], ref["is_visible"], ref["weight_points"])
    got = _ours_render(proj.numpy(), res, res, 7, 9, 1e-5, 0.1, gd.cpu().numpy(), gw.cpu().numpy(), cuda)
    refn = {k: v.cpu().numpy() for k, v in ref.items()}
    o = eo.render_forward(proj.numpy(), res, res, 7, 9, 1e-5, 0.1)
    assert np.array_equal(got["is_visible"], o["is_visible"])
    _check_render_vs_reference(got, refn, proj.numpy(), res, res, 9, o["front"], o["back"])
    gp = ext.point_render_backward(gd, gw, pd, ref["pixel_index"], ref["is_visible"], 9)[0]
    assert torch.equal(gp.view(torch.int32), ref_gp.view(torch.int32))
    # masker
    mask = ext.point_masker_forward(pd, res, res, 5)[0]
    assert torch.equal(mask, exfunc_ref.masker_forward(pd, res, res, 5))
    other = ext.point_masker_forward(_cloud(B, N, res, res, 9).to(cuda), res, res, 5)[0]
    gm = torch.sign(mask - other) / mask.numel()                        # d mean|a-b| / da  (loss.py:195)
    ours = ext.point_masker_backward(gm, pd, 1e-5, 1e-10)[0]
    theirs = exfunc_ref.masker_backward(gm.contiguous(), pd, 1e-5, 1e-10)
    _close_sum(ours.cpu().numpy(), theirs.cpu().numpy(), "mask backward", rel=1e-5)


def test_autograd_modules_like_loss_py(cuda):
    """Render / Masker used the way loss_on_depth / loss_on_mask use them (loss.py:169-177, 193-195)."""
    from rma_net_b200.model.exfunc.functions import Masker, Render
    res = 128
    render, masker = Render(res, res, 7, 9, 1e-5), Masker(res, res, 5, 1e-5)
    a = _cloud(2, 1500, res, res, 11).to(cuda).requires_grad_(True)
    b = _cloud(2, 1500, res, res, 12).to(cuda)
    da, wa, va = render(a, 0.1)
    db, wb, _ = render(b, 0.1)
    assert da.shape == (2, res, res, 1) and va.dtype == torch.int32 and not va.requires_grad
    ia, ib = da / wa, db / wb
    loss = ((ia - ib) ** 2).sum()
    loss.backward()
    g_render = a.grad.clone()
    assert torch.isfinite(g_render).all() and g_render.abs().max() > 0
    # numerical check of d loss / d z on a few visible points (z enters depth_image linearly)
    o = eo.render_forward(a.detach().cpu().numpy(), res, res, 7, 9, 1e-5, 0.1)
    assert np.array_equal(o["is_visible"], va.cpu().numpy())
    np.testing.assert_allclose(da.detach().cpu().numpy()[..., 0], o["depth_image"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(wa.detach().cpu().numpy()[..., 0], o["weight_image"], rtol=0, atol=2e-5)
    a.grad = None
    lm = (masker(a, 1e-10) - masker(b, 1e-10)).abs().mean()
    lm.backward()
    assert torch.all(a.grad[..., :2] == 0) and a.grad[..., 2].abs().max() > 0
    with pytest.raises(RuntimeError):
        render(a.detach().cpu(), 0.1)
    with pytest.raises(RuntimeError):
        render(a.detach().transpose(0, 1), 0.1)                       # CHECK_CONTIGUOUS (point_render.cpp:24-25)


def test_numpy_oracle_agrees_on_gpu_case(cuda):
    from rma_net_b200 import ext
    proj = _cloud(1, 400, 48, 48, 21)
    o = eo.render_forward(proj.numpy(), 48, 48, 5, 7, 1e-5, 0.1)
    rng = np.random.default_rng(0)
    gd = rng.normal(size=(1, 48, 48, 1)).astype(np.float32)
    gw = rng.normal(size=(1, 48, 48, 1)).astype(np.float32)
    got = _ours_render(proj.numpy(), 48, 48, 5, 7, 1e-5, 0.1, gd, gw, cuda)
    assert np.array_equal(got["pixel_index"].reshape(1, -1, 2), o["pixel_index"])
    assert np.array_equal(got["is_visible"], o["is_visible"])
    _close_sum(got["depth_image"][..., 0], o["depth_image"], "depth")
    ob = eo.render_backward(gd, gw, proj.numpy(), o["pixel_index"], o["is_visible"], got["weight_points"], 7)
    np.testing.assert_allclose(got["grad_proj"], ob, rtol=0, atol=2e-6 * np.abs(ob).max())
    gm = np.where(rng.random((1, 48, 48, 1)) < 0.2, rng.normal(size=(1, 48, 48, 1)), 0).astype(np.float32)
    ours = ext.point_masker_backward(torch.from_numpy(gm).to(cuda), proj.to(cuda), 1e-5, 1e-10)[0].cpu().numpy()
    _close_sum(ours, eo.masker_backward(gm, proj.numpy(), 1e-5, 1e-10), "mask backward vs numpy", rel=1e-5)


def test_multi_view_losses_batched_equals_per_view_loop(cuda):
    """loss_on_depth / loss_on_mask (loss.py:147-196): the views folded into the batch (2 kernel calls) against the
    reference's literal per-view loop (2V calls) on the same kernels, and euler2rot against the reference formula."""
    from rma_net_b200.model.loss import euler2rot, loss_on_depth, loss_on_mask, view_rotations
    g = torch.Generator().manual_seed(5)
    B, N, vd = 2, 1200, 3
    src = (torch.randn(B, N, 3, generator=g) * 0.3).clamp(-0.9, 0.9).to(cuda).requires_grad_(True)
    tgt = (torch.randn(B, N, 3, generator=g) * 0.3).clamp(-0.9, 0.9).to(cuda)
    rot = view_rotations(vd, cuda, jitter=torch.rand(vd * vd, 2, generator=g))
    assert rot.shape == (3 * vd * vd, 3)
    r0 = rot[:3]
    assert torch.allclose(r0 @ r0.T, torch.eye(3, device=cuda), atol=1e-6)
    e = euler2rot(torch.tensor([[0.3, -0.2, 0.1]], device=cuda))[0]
    assert torch.allclose(e @ e.T, torch.eye(3, device=cuda), atol=1e-6) and abs(torch.det(e).item() - 1) < 1e-5
    for fn, rel in ((loss_on_depth, 2e-5), (loss_on_mask, 1e-5)):
        src.grad = None
        a = fn(src, tgt, rot, resolution=128, batched=True)
        a.backward()
        ga = src.grad.clone()
        src.grad = None
        b = fn(src, tgt, rot, resolution=128, batched=False)
        b.backward()
        gb = src.grad.clone()
        assert abs(a.item() - b.item()) <= rel * abs(b.item()) + 1e-12, (fn.__name__, a.item(), b.item())
        assert (ga - gb).abs().max() <= 2e-5 * gb.abs().max() + 1e-12, fn.__name__
        assert gb.abs().max() > 0
