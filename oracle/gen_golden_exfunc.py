"""Generate tests/golden/exfunc_*.npz from the REFERENCE's own point_render / point_masker kernels.

Needs a GPU and oracle/_ref/libexfunc_ref.so (make -C oracle _ref/libexfunc_ref.so, in the build container where
/root/reference exists).  Run on the GPU box:   python oracle/gen_golden_exfunc.py gpurun_out/golden_exfunc
then copy the .npz files into tests/golden/.  Inputs are seeded; every array the tests compare is stored."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import exfunc_ref  # noqa: E402

CASES = {
    # name: (B, N, H, W, vpw, cpw, mpw, epsilon, threshold, threshold2, kind)
    "small": (2, 300, 64, 64, 7, 9, 5, 1e-5, 0.1, 1e-10, "cloud"),
    "edges": (1, 120, 48, 40, 6, 4, 3, 1e-5, 0.05, 1e-10, "outside"),
    "layers": (1, 400, 32, 32, 5, 7, 5, 1e-3, 0.1, 1e-10, "layers"),
}


def make_points(B, N, H, W, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "cloud":        # loss.py:166 mapping of a unit-ball cloud: xy -> pixels, z - 1
        p = torch.randn(B, N, 3, generator=g) * 0.35
        p = p.clamp(-0.99, 0.99)
        xy = (p[..., :2] + 1.0) * (torch.tensor([W, H]) - 1) / 2
        return torch.cat([xy, p[..., 2:] - 1.0], -1).contiguous()
    if kind == "outside":      # points off the image on every side, positive z, exact pixel centres and borders
        p = torch.rand(B, N, 3, generator=g)
        xy = p[..., :2] * torch.tensor([W + 20.0, H + 20.0]) - 10.0
        z = p[..., 2:] * 3.0 - 2.5
        out = torch.cat([xy, z], -1)
        out[:, :8, 0] = torch.tensor([0.0, 0.5, W - 1.0, W - 0.5, float(W), -0.0, 3.0, 7.5])
        return out.contiguous()
    # two depth layers over the same pixels: the far one must be hidden
    xy = torch.rand(B, N, 2, generator=g) * torch.tensor([W - 1.0, H - 1.0])
    z = torch.where(torch.arange(N)[None, :, None] % 2 == 0, torch.tensor(-1.6), torch.tensor(-0.4))
    z = z + torch.rand(B, N, 1, generator=g) * 0.05
    return torch.cat([xy, z.expand(B, N, 1)], -1).contiguous()


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_exfunc"
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    for k, (name, (B, N, H, W, vpw, cpw, mpw, eps, thr, thr2, kind)) in enumerate(CASES.items()):
        proj = make_points(B, N, H, W, kind, 100 + k)
        pd = proj.to(dev)
        f = exfunc_ref.render_forward(pd, H, W, vpw, cpw, eps, thr)
        g = torch.Generator().manual_seed(200 + k)
        gd = torch.randn(B, H, W, 1, generator=g).to(dev)
        gw = torch.randn(B, H, W, 1, generator=g).to(dev)
        gp = exfunc_ref.render_backward(gd, gw, pd, f["pixel_index"], f["is_visible"], f["weight_points"])
        mask = exfunc_ref.masker_forward(pd, H, W, mpw)
        gm = torch.randn(B, H, W, 1, generator=g)
        gm = torch.where(torch.rand(B, H, W, 1, generator=g) < 0.3, gm, torch.zeros(()))     # sparse, like |a-b| grads
        gmask = exfunc_ref.masker_backward(gm.to(dev), pd, eps, thr2)
        np.savez_compressed(
            os.path.join(out_dir, f"exfunc_{name}.npz"),
            params=np.array([B, N, H, W, vpw, cpw, mpw], np.int64), fparams=np.array([eps, thr, thr2], np.float64),
            proj=proj.numpy(), pixel_index=f["pixel_index"].cpu().numpy(), is_visible=f["is_visible"].cpu().numpy(),
            weight_points=f["weight_points"].cpu().numpy(), depth_image=f["depth_image"].cpu().numpy(),
            weight_image=f["weight_image"].cpu().numpy(), front=f["front"].cpu().numpy(), back=f["back"].cpu().numpy(),
            grad_depth=gd.cpu().numpy(), grad_weight=gw.cpu().numpy(), grad_proj=gp.cpu().numpy(),
            mask=mask.cpu().numpy(), grad_mask=gm.numpy(), grad_proj_mask=gmask.cpu().numpy())
        print(name, "ok", int(f["is_visible"].sum()), "visible of", B * N, "mask px", int((mask != 0).sum()))


if __name__ == "__main__":
    main()
