"""Generate tests/golden/*.npz by running the REFERENCE's own code in the build container.

Run from the repo root:   python oracle/gen_golden.py
Needs /root/reference (read-only).  The outputs are committed; the GPU box never reads /root/reference.

* chamfer: lines 45-70 of /root/reference/model/loss.py (compute_sqrdis_map, chamfer_dist) are exec'd verbatim —
  the enclosing module cannot be imported (trimesh, the compiled point_render extension, cuda:0 allocations;
  SURVEY.md §8c) — on the four sample OBJ pairs and on seeded synthetic clouds; loss_on_cd (loss.py:200-205) is
  applied on top and differentiated with torch autograd.
* se3: /root/reference/model/LieAlgebra is imported as a package; Exp forward/backward, transform (both branches),
  and the non-rigid warp loop of network.py:652-655 composed from those reference functions, differentiated with
  torch autograd (which routes through the reference's ExpMap.backward).
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference_chamfer():
    with open(os.path.join(REF, "model", "loss.py")) as f:
        lines = f.readlines()
    ns = {"torch": torch}
    exec("".join(lines[44:70]), ns)          # loss.py:45-70
    return ns["compute_sqrdis_map"], ns["chamfer_dist"]


def read_obj(path):
    pts = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) >= 4 and t[0] == "v":
                pts.append([float(t[1]), float(t[2]), float(t[3])])
    return np.asarray(pts, np.float32)


def chamfer_case(name, x, y, compute_sqrdis_map, chamfer_dist):
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    yt = torch.from_numpy(y).clone().requires_grad_(True)
    d1, d2 = chamfer_dist(xt, yt)
    loss = (torch.sum(d1) + torch.sum(d2)) * 0.5 / xt.shape[0]          # loss.py:200-205
    loss.backward()
    with torch.no_grad():
        sq = compute_sqrdis_map(xt, yt)
        i1 = sq.min(dim=2)[1]
        i2 = sq.min(dim=1)[1]
    np.savez_compressed(
        os.path.join(OUT, f"chamfer_{name}.npz"), x=x, y=y,
        ref_dist1=d1.detach().numpy(), ref_dist2=d2.detach().numpy(),
        ref_idx1=i1.numpy().astype(np.int32), ref_idx2=i2.numpy().astype(np.int32),
        ref_loss=np.float32(loss.item()), ref_grad_x=xt.grad.numpy(), ref_grad_y=yt.grad.numpy())
    print(name, x.shape, y.shape, "loss", loss.item(), "sum d1", d1.sum().item(), "sum d2", d2.sum().item())


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(4)
    compute_sqrdis_map, chamfer_dist = load_reference_chamfer()

    for k in (1, 2, 3, 4):
        x = read_obj(os.path.join(REF, "samples", str(k), "source.obj"))[None]
        y = read_obj(os.path.join(REF, "samples", str(k), "target.obj"))[None]
        chamfer_case(f"samples{k}", x, y, compute_sqrdis_map, chamfer_dist)

    g = torch.Generator().manual_seed(10)
    x = (torch.rand(3, 700, 3, generator=g) * 2 - 1).numpy()
    y = (torch.rand(3, 900, 3, generator=g) * 2 - 1).numpy()
    chamfer_case("uniform_3x700x900", x, y, compute_sqrdis_map, chamfer_dist)
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(2, 1024, 3, generator=g) * 0.3).numpy()
    y = (torch.randn(2, 513, 3, generator=g) * 0.3).numpy()
    chamfer_case("normal_2x1024x513", x, y, compute_sqrdis_map, chamfer_dist)
    # exact ties (SURVEY.md §8c tie test) + duplicated points
    x = np.array([[[0, 0, 0], [1, 1, 1], [0, 0, 0]]], np.float32)
    y = np.array([[[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 1, 0], [1, 1, 1], [1, 1, 1]]], np.float32)
    chamfer_case("ties", x, y, compute_sqrdis_map, chamfer_dist)

    # ---- se3 -------------------------------------------------------------------------------------------------
    sys.path.insert(0, REF)
    import model.LieAlgebra as LA            # namespace package import of the reference

    se3 = LA.se3
    g = torch.Generator().manual_seed(20)
    tw = torch.cat([
        torch.tensor([[0.1, -0.2, 0.3, 0.5, -0.4, 0.25], [0, 0, 0, 1, 2, 3],
                      [1e-4, 2e-4, -1e-4, .01, .02, .03], [0.005, 0.003, -0.004, 1., -1., 2.]]),
        torch.randn(24, 6, generator=g) * 0.1,
        torch.randn(24, 6, generator=g) * 0.8,
        torch.randn(12, 6, generator=g) * 2.5,
    ]).float()
    x = tw.clone().requires_grad_(True)
    G = se3.Exp(x)
    go = torch.randn(G.shape, generator=g)
    G.backward(go)
    np.savez_compressed(os.path.join(OUT, "se3_exp.npz"), x=tw.numpy(), ref_g=G.detach().numpy(),
                        grad_out=go.numpy(), ref_grad_x=x.grad.numpy())
    print("se3_exp", tuple(G.shape))

    # transform, hot branch: g [B,1,4,4], a [B,N,3]
    B, N = 3, 257
    phi = (torch.randn(B, 6, generator=g) * 0.5).requires_grad_(True)
    a = torch.randn(B, N, 3, generator=g).requires_grad_(True)
    gm = se3.Exp(phi).view(B, 1, 4, 4)
    gm.retain_grad()
    b = se3.transform(gm, a)
    gb = torch.randn(b.shape, generator=g)
    b.backward(gb)
    # other branch: g [B,4,4], a [B,3,N]
    a2 = a.detach().transpose(1, 2).contiguous()
    b2 = se3.transform(gm.detach().view(B, 4, 4), a2)
    np.savez_compressed(os.path.join(OUT, "se3_transform.npz"), phi=phi.detach().numpy(), a=a.detach().numpy(),
                        g=gm.detach().numpy(), ref_b=b.detach().numpy(), grad_out=gb.numpy(),
                        ref_grad_g=gm.grad.numpy(), ref_grad_a=a.grad.numpy(), ref_grad_phi=phi.grad.numpy(),
                        ref_b_3n=b2.numpy())
    print("se3_transform", tuple(b.shape))

    # non-rigid warp loop, 4 stages (network.py:567-570, 606, 652-655) + chamfer loss on each stage (loss.py:260-267)
    B, N, M, T = 2, 300, 280, 4
    src = (torch.rand(B, N, 3, generator=g) * 2 - 1)
    tgt = (torch.rand(B, M, 3, generator=g) * 2 - 1)
    phis = [(torch.randn(B, 6, generator=g) * 0.1).requires_grad_(True) for _ in range(T)]
    ws = [torch.rand(B, 1, N, generator=g).requires_grad_(True) for _ in range(T)]
    src_3n = src.transpose(1, 2).contiguous()                          # network holds [B,3,N]
    defs = []
    deformation = None
    for k in range(T):
        if k > 0:
            deformation = deformation.detach()
        rigid_matrix = se3.Exp(phis[k]).view(B, 1, 4, 4)
        rigid = se3.transform(rigid_matrix, src_3n.transpose(1, 2)).transpose(1, 2)
        deformation = rigid if k == 0 else torch.mul(ws[k], rigid) + torch.mul((1 - ws[k]), deformation)
        defs.append(deformation)
    gamma = 0.8
    loss = 0
    for i in range(T):
        d1, d2 = chamfer_dist(defs[i].transpose(1, 2), tgt)
        loss = loss + gamma ** (T - i - 1) * (torch.sum(d1) + torch.sum(d2)) * 0.5 / B
    loss.backward()
    np.savez_compressed(
        os.path.join(OUT, "warp_loop.npz"), src=src.numpy(), tgt=tgt.numpy(),
        phis=np.stack([p.detach().numpy() for p in phis]), ws=np.stack([w.detach().numpy() for w in ws]),
        ref_defs=np.stack([d.detach().transpose(1, 2).numpy() for d in defs]), ref_loss=np.float32(loss.item()),
        ref_grad_phis=np.stack([p.grad.numpy() for p in phis]),
        ref_grad_ws=np.stack([(w.grad if w.grad is not None else torch.zeros_like(w)).numpy() for w in ws]),
        gamma=np.float32(gamma))
    print("warp_loop loss", loss.item())


if __name__ == "__main__":
    main()
