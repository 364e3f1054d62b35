"""GPU parity tests: se(3) exp-map, transform and the fused recurrent-loop warp vs the fp64 oracle and the golden
outputs of the reference's LieAlgebra package.  Tolerance: 1e-5 relative to the output tensor's |.|_inf (fp32
kernels vs fp64 oracle); against the reference's own fp32 outputs 3e-5 (its sinc2/sinc3 cancel near t = 0.01)."""
import os

import numpy as np
import pytest
import torch

from oracle import chamfer_oracle as co, se3_oracle as so

pytestmark = pytest.mark.gpu
REL = 1e-5


def _close(a, b, rel=REL):
    b = np.asarray(b)
    np.testing.assert_allclose(np.asarray(a), b, rtol=0, atol=rel * max(np.abs(b).max(), 1e-30))


def test_exp_forward_backward_golden(cuda, golden_dir):
    from rma_net_b200.model.LieAlgebra import se3
    g = dict(np.load(os.path.join(golden_dir, "se3_exp.npz")))
    x = torch.from_numpy(g["x"]).to(cuda).requires_grad_(True)
    G = se3.Exp(x)
    assert G.shape == (g["x"].shape[0], 4, 4)
    _close(G.detach().cpu().numpy(), so.se3_exp(g["x"].astype(np.float64)))
    _close(G.detach().cpu().numpy(), g["ref_g"], 3e-5)
    G.backward(torch.from_numpy(g["grad_out"]).to(cuda))
    _close(x.grad.cpu().numpy(), so.se3_expmap_backward(g["x"].astype(np.float64), g["grad_out"]))
    _close(x.grad.cpu().numpy(), g["ref_grad_x"], 3e-5)


def test_exp_known_answers_and_shapes(cuda):
    from rma_net_b200.model.LieAlgebra import se3
    x = torch.tensor([[0.1, -0.2, 0.3, 0.5, -0.4, 0.25], [0, 0, 0, 1, 2, 3], [1e-4, 2e-4, -1e-4, .01, .02, .03]],
                     device=cuda)
    G = se3.Exp(x).cpu().numpy()
    np.testing.assert_allclose(G[0, 0], [0.935754836, -0.302932739, -0.180540085, 0.526401162], atol=3e-7)
    np.testing.assert_allclose(G[1], [[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]], atol=0)
    np.testing.assert_allclose(G[2, :3, 3], [0.0100040, 0.0199980, 0.0300000], atol=1e-7)
    # leading dims: [B,1,6] -> [B,1,4,4]; per-point twists [B,N,6] -> [B,N,4,4]
    assert se3.Exp(x.view(3, 1, 6)).shape == (3, 1, 4, 4)
    tw = torch.randn(2, 1000, 6, device=cuda) * 0.7
    Gp = se3.Exp(tw)
    _close(Gp.cpu().numpy(), so.se3_exp(tw.cpu().numpy().astype(np.float64)))
    # sweep |t| across the Taylor threshold and up to ~pi
    t = torch.logspace(-6, 0.5, 400, device=cuda)
    ax = torch.nn.functional.normalize(torch.randn(400, 3, device=cuda), dim=1)
    tw = torch.cat([ax * t[:, None], torch.randn(400, 3, device=cuda)], 1)
    _close(se3.Exp(tw).cpu().numpy(), so.se3_exp(tw.cpu().numpy().astype(np.float64)), 2e-6)
    assert se3.exp(tw).requires_grad is False and se3.genmat().shape == (6, 4, 4)


def test_transform_both_branches_golden(cuda, golden_dir):
    from rma_net_b200.model.LieAlgebra import se3
    g = dict(np.load(os.path.join(golden_dir, "se3_transform.npz")))
    phi = torch.from_numpy(g["phi"]).to(cuda).requires_grad_(True)
    a = torch.from_numpy(g["a"]).to(cuda).requires_grad_(True)
    B = phi.shape[0]
    gm = se3.Exp(phi).view(B, 1, 4, 4)
    gm.retain_grad()
    b = se3.transform(gm, a)                                    # g [B,1,4,4] x a [B,N,3]
    _close(b.detach().cpu().numpy(), g["ref_b"], 3e-6)
    b.backward(torch.from_numpy(g["grad_out"]).to(cuda))
    _close(a.grad.cpu().numpy(), g["ref_grad_a"], 3e-6)
    _close(gm.grad.cpu().numpy()[..., :3, :], g["ref_grad_g"][..., :3, :])
    _close(phi.grad.cpu().numpy(), g["ref_grad_phi"], 3e-5)
    gg, ga = so.se3_transform_backward(so.se3_exp(g["phi"].astype(np.float64))[:, None], g["a"].astype(np.float64),
                                       g["grad_out"])
    _close(gm.grad.cpu().numpy(), gg)
    _close(phi.grad.cpu().numpy(), so.se3_expmap_backward_closed(g["phi"].astype(np.float64), gg[:, 0]))   # 1e-5 vs fp64
    # same-rank branch: g [B,4,4] x a [B,3,N], and the strided [B,3,N]->transpose view the network passes
    a3n = a.detach().transpose(1, 2).contiguous()
    b2 = se3.transform(gm.detach().view(B, 4, 4), a3n)
    _close(b2.cpu().numpy(), g["ref_b_3n"], 3e-6)
    b3 = se3.transform(gm.detach(), a3n.transpose(1, 2))        # network.py:654 call shape
    _close(b3.cpu().numpy(), g["ref_b"], 3e-6)
    # per-point transforms: g [B,N,4,4] x a [B,N,3]
    tw = (torch.randn(B, a.shape[1], 6, device=cuda) * 0.4).requires_grad_(True)
    ad = a.detach().clone().requires_grad_(True)
    out = se3.transform(se3.Exp(tw), ad)
    ref = so.se3_transform(so.se3_exp(tw.detach().cpu().numpy().astype(np.float64)), g["a"].astype(np.float64))
    _close(out.detach().cpu().numpy(), ref)
    out.sum().backward()
    assert tw.grad.shape == tw.shape and ad.grad.shape == ad.shape
    Rt = np.swapaxes(so.se3_exp(tw.detach().cpu().numpy().astype(np.float64))[..., :3, :3], -1, -2)
    _close(ad.grad.cpu().numpy(), Rt.sum(-1))


def test_fused_warp_vs_composed_ops_and_golden_loop(cuda, golden_dir):
    """The recurrent loop (network.py:567-570, 606, 652-655) + gamma-weighted chamfer loss (loss.py:260-267):
    fused Warp op, and the drop-in composition Exp -> transform -> blend, both against the reference's outputs."""
    from rma_net_b200.model.LieAlgebra import se3
    from rma_net_b200.model.exfunc.functions import Warp
    from rma_net_b200.model.loss import loss_cd_stages
    g = dict(np.load(os.path.join(golden_dir, "warp_loop.npz")))
    T, B = g["phis"].shape[0], g["src"].shape[0]
    src = torch.from_numpy(g["src"]).to(cuda)
    tgt = torch.from_numpy(g["tgt"]).to(cuda)
    src_3n = src.transpose(1, 2).contiguous()
    warp = Warp()
    for mode in ("fused", "composed"):
        phis = [torch.from_numpy(p).to(cuda).requires_grad_(True) for p in g["phis"]]
        ws = [torch.from_numpy(w).to(cuda).requires_grad_(True) for w in g["ws"]]
        defs, deformation = [], None
        for k in range(T):
            if mode == "fused":
                prev = None if k == 0 else deformation.transpose(1, 2)
                out, rigid, gmat = warp(phis[k], src_3n.transpose(1, 2), None if k == 0 else ws[k], prev)
                deformation = out.transpose(1, 2)
                assert gmat.shape == (B, 4, 4) and rigid.shape == out.shape
            else:
                if k > 0:
                    deformation = deformation.detach()
                rigid_matrix = se3.Exp(phis[k]).view(B, 1, 4, 4)
                rigid = se3.transform(rigid_matrix, src_3n.transpose(1, 2)).transpose(1, 2)
                deformation = rigid if k == 0 else torch.mul(ws[k], rigid) + torch.mul((1 - ws[k]), deformation)
            defs.append(deformation)
        for k in range(T):
            _close(defs[k].detach().transpose(1, 2).cpu().numpy(), g["ref_defs"][k], 3e-6)
        loss = loss_cd_stages(defs, tgt, gamma=float(g["gamma"]))
        loss.backward()
        assert abs(loss.item() - float(g["ref_loss"])) < 2e-5 * abs(loss.item())
        for k in range(T):
            _close(phis[k].grad.cpu().numpy(), g["ref_grad_phis"][k], 2e-4)
            if k > 0:
                _close(ws[k].grad.cpu().numpy(), g["ref_grad_ws"][k], 2e-4)


def test_warp_backward_vs_oracle_large(cuda):
    from rma_net_b200.model.exfunc.functions import Warp
    rng = np.random.default_rng(4)
    B, N = 4, 5000
    src = rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)
    prev = rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)
    phi = (rng.normal(size=(B, 6)) * 0.3).astype(np.float32)
    w = rng.uniform(0, 1, (B, N)).astype(np.float32)
    go = rng.normal(size=(B, N, 3)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    phit, wt, srct = t(phi).requires_grad_(True), t(w).requires_grad_(True), t(src).requires_grad_(True)
    out, rigid, gm = Warp()(phit, srct, wt, t(prev))
    out.backward(t(go))
    gphi, gw = so.warp_stage_backward(phi.astype(np.float64), src.astype(np.float64), w.astype(np.float64),
                                      prev.astype(np.float64), go.astype(np.float64))
    _close(phit.grad.cpu().numpy(), gphi)           # 1e-5: the 12 sums over N and the cross products run in double
    _close(wt.grad.cpu().numpy(), gw)
    R = so.se3_exp(phi.astype(np.float64))[:, :3, :3]
    gsrc = np.einsum("bji,bnj->bni", R, go.astype(np.float64) * w[..., None])
    _close(srct.grad.cpu().numpy(), gsrc)
    ref = so.warp_nonrigid_loop(src.astype(np.float64), [phi, phi], [None, w])
    rigid64 = ref[0]
    _close(rigid.detach().cpu().numpy(), rigid64, 3e-6)
    _close(out.detach().cpu().numpy(), w[..., None] * rigid64 + (1 - w[..., None]) * prev, 3e-6)


def test_warp_matrix_and_rigid_outputs_are_differentiable(cuda):
    """The rigid matrix and the rigid points returned by Warp carry gradient back to phi, as in the reference
    (network.py:653-654: rigid_matrix = Exp(phi) feeds rigid_matrix_list -> Loss.loss_on_tran, loss.py:224-230,
    311-319; train_sample.py trains with --weight_tran 0.1).  Checked against the composed drop-in ops and fp64."""
    from rma_net_b200.model.LieAlgebra import se3
    from rma_net_b200.model.exfunc.functions import Warp
    from rma_net_b200.model.loss import loss_on_tran
    rng = np.random.default_rng(11)
    B, N = 3, 777
    src = torch.from_numpy(rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)).to(cuda)
    prev = torch.from_numpy(rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)).to(cuda)
    phi_np = (rng.normal(size=(B, 6)) * 0.3).astype(np.float32)
    w_np = rng.uniform(0, 1, (B, 1, N)).astype(np.float32)
    go_out = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda)
    go_rig = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda)
    for blend in (True, False):
        grads = {}
        for mode in ("fused", "composed"):
            phi = torch.from_numpy(phi_np).to(cuda).requires_grad_(True)
            w = torch.from_numpy(w_np).to(cuda).requires_grad_(True)
            if mode == "fused":
                out, rigid, gm = Warp()(phi, src, w if blend else None, prev if blend else None)
            else:
                gm4 = se3.Exp(phi).view(B, 1, 4, 4)
                rigid = se3.transform(gm4, src)
                out = torch.mul(w.transpose(1, 2), rigid) + torch.mul(1 - w.transpose(1, 2), prev) if blend else rigid
                gm = gm4.view(B, 4, 4)
            assert gm.requires_grad and rigid.requires_grad and out.requires_grad
            loss = 0.1 * loss_on_tran(gm) + (out * go_out).sum() + 0.5 * (rigid * go_rig).sum()
            loss.backward()
            grads[mode] = (phi.grad.cpu().numpy(), w.grad.cpu().numpy() if blend else None)
        _close(grads["fused"][0], grads["composed"][0])
        if blend:
            _close(grads["fused"][1], grads["composed"][1])
        # fp64: d/dphi of 0.1 * |p|^2 / B alone (only the matrix output carries gradient)
        phi = torch.from_numpy(phi_np).to(cuda).requires_grad_(True)
        _, _, gm = Warp()(phi, src, torch.from_numpy(w_np).to(cuda) if blend else None, prev if blend else None)
        (0.1 * loss_on_tran(gm)).backward()
        g64 = so.se3_exp(phi_np.astype(np.float64))
        gg = np.zeros_like(g64)
        gg[:, :3, 3] = 0.2 * g64[:, :3, 3] / B
        _close(phi.grad.cpu().numpy(), so.se3_expmap_backward_closed(phi_np.astype(np.float64), gg))


def test_recurrent_step_c3_size_vs_oracle(cuda):
    """BASELINE.json configs[2] at full size: 8 stages of warp + chamfer, B=16, N=M=8192, gamma = 0.8
    (network.py:567-570, 606, 652-655 + loss.py:260-267).  Deformations against se3_oracle.warp_nonrigid_loop (fp64);
    gradients w.r.t. every phi_k and w_k within 1e-5 of the fp64 chain chamfer backward (C oracle, on the indices the
    GPU found - those are pinned bit-exact elsewhere) -> warp_stage_backward."""
    from oracle import c_oracle
    from rma_net_b200.model.exfunc.functions import Warp
    from rma_net_b200.model.loss import chamfer_dist_with_idx, loss_cd_stages_fused, loss_on_cd
    B, N, T, gamma = 16, 8192, 8, 0.8
    g = torch.Generator().manual_seed(2)
    src = torch.rand(B, N, 3, generator=g) * 2 - 1
    tgt = torch.rand(B, N, 3, generator=g) * 2 - 1
    phis = [torch.randn(B, 6, generator=g) * 0.1 for _ in range(T)]
    ws = [torch.rand(B, 1, N, generator=g) for _ in range(T)]
    srcd, tgtd = src.to(cuda), tgt.to(cuda)
    ref_defs = so.warp_nonrigid_loop(src.numpy().astype(np.float64), [p.numpy() for p in phis],
                                     [w.numpy()[:, 0] for w in ws])
    warp = Warp()
    for fused in (False, True):
        phid = [p.to(cuda).requires_grad_(True) for p in phis]
        wd = [w.to(cuda).requires_grad_(True) for w in ws]
        outs, prev = [], None
        for k in range(T):
            out_k, _, _ = warp(phid[k], srcd, None if k == 0 else wd[k], prev)
            outs.append(out_k)
            prev = out_k
        if fused:
            loss = loss_cd_stages_fused([o.transpose(1, 2) for o in outs], tgtd, gamma=gamma)
        else:
            loss = sum(gamma ** (T - k - 1) * loss_on_cd(outs[k], tgtd) for k in range(T))
        loss.backward()
        for k in range(T):
            o_np = outs[k].detach().cpu().numpy()
            _close(o_np, ref_defs[k], 3e-6)
            with torch.no_grad():
                _, _, i1, i2 = chamfer_dist_with_idx(outs[k].detach(), tgtd)
            coef = gamma ** (T - k - 1) * 0.5 / B
            gx, _ = c_oracle.chamfer_bwd(o_np, tgt.numpy(), i1.cpu().numpy(), i2.cpu().numpy(), coef, coef)
            gphi, gw = so.warp_stage_backward(phis[k].numpy().astype(np.float64), src.numpy().astype(np.float64),
                                              None if k == 0 else ws[k].numpy()[:, 0].astype(np.float64),
                                              None if k == 0 else ref_defs[k - 1], gx)
            _close(phid[k].grad.cpu().numpy(), gphi)
            if k > 0:
                _close(wd[k].grad.cpu().numpy()[:, 0], gw)
