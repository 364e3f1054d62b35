"""point_render / point_masker: B200 kernels vs the REFERENCE's own kernels (oracle/_ref/libexfunc_ref.so), at the
loss.py:33-37 configuration (256x256, Render(7, 9), Masker(5)), forward + backward per call, CUDA events.

The reference leg is timed twice: `kernels` = its launchers alone on pre-filled buffers, and `wrapper` = with what
point_render.cpp / point_masker.cpp do around them on every call (6 torch fills and 3 cudaDeviceSynchronize for the
render forward, point_render.cpp:54-96), which is what a user of the reference actually pays.
    python tests/bench_exfunc_vs_reference.py [B] [N]   (lives under tests/: it runs the reference kernels of oracle/_ref)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import exfunc_ref  # noqa: E402
from rma_net_b200 import ext  # noqa: E402


def cloud(B, N, res, seed):
    g = torch.Generator().manual_seed(seed)
    p = (torch.randn(B, N, 3, generator=g) * 0.35).clamp(-0.99, 0.99)
    xy = (p[..., :2] + 1.0) * (res - 1) / 2
    return torch.cat([xy, p[..., 2:] - 1.0], -1).contiguous()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    res, vpw, cpw, mpw, eps, thr, thr2 = 256, 7, 9, 5, 1e-5, 0.1, 1e-10
    dev = torch.device("cuda:0")
    pd = cloud(B, N, res, 1).to(dev)
    other = cloud(B, N, res, 2).to(dev)
    g = torch.Generator().manual_seed(3)
    gd = torch.randn(B, res, res, 1, generator=g).to(dev)
    gw = torch.randn(B, res, res, 1, generator=g).to(dev)
    out = {"config": f"B={B} N={N} {res}x{res} Render({vpw},{cpw}) Masker({mpw})"}

    def ours_render():
        pix, vis, _, depth, weight = ext.point_render_forward(pd, res, res, vpw, cpw, eps, thr)
        ext.point_render_backward(gd, gw, pd, pix, vis, cpw)
    out["render_fwd_bwd_us_ours"] = timed(ours_render)
    mask_a = ext.point_masker_forward(pd, res, res, mpw)[0]
    mask_b = ext.point_masker_forward(other, res, res, mpw)[0]
    gm = (torch.sign(mask_a - mask_b) / mask_a.numel()).contiguous()
    out["mask_active_pixels_per_image"] = int((gm != 0).sum()) // B

    def ours_mask():
        ext.point_masker_forward(pd, res, res, mpw)
        ext.point_masker_backward(gm, pd, eps, thr2)
    out["mask_fwd_bwd_us_ours"] = timed(ours_mask)

    if exfunc_ref.available():
        import ctypes
        lib = exfunc_ref._load()
        P = exfunc_ref._p
        pixel_index = torch.empty((B, 2 * N), dtype=torch.int32, device=dev)
        is_visible = torch.empty((B, N), dtype=torch.int32, device=dev)
        front = torch.empty((B, res, res, 1), device=dev); back = torch.empty_like(front)
        wp = torch.empty((B, N, cpw * cpw), device=dev)
        depth = torch.empty_like(front); weight = torch.empty_like(front)
        gwp = torch.empty_like(wp); gproj = torch.empty_like(pd)

        def fills():
            pixel_index.fill_(-1); is_visible.fill_(1); front.zero_(); back.fill_(-2.0); wp.zero_(); depth.zero_()
            weight.fill_(eps)

        def ref_kernels():
            fills(); gwp.zero_(); gproj.zero_()
            lib.ref_point_render_forward(P(pd), P(pixel_index), P(is_visible), P(front), P(back), P(wp), P(depth),
                                         P(weight), B, N, res, res, vpw, cpw, ctypes.c_float(thr))
            lib.ref_point_render_backward(P(gd), P(gw), P(pd), P(pixel_index), P(is_visible), P(wp), P(gwp), P(gproj),
                                          B, N, res, res, cpw)

        def ref_wrapper():      # point_render.cpp:54-96, 116-130: allocate + fill + device syncs around the launches
            torch.cuda.synchronize(); torch.cuda.synchronize()
            f = exfunc_ref.render_forward(pd, res, res, vpw, cpw, eps, thr)
            exfunc_ref.render_backward(gd, gw, pd, f["pixel_index"], f["is_visible"], f["weight_points"])
        out["render_fwd_bwd_us_reference_kernels"] = timed(ref_kernels)
        out["render_fwd_bwd_us_reference_wrapper"] = timed(ref_wrapper)

        def ref_mask():
            exfunc_ref.masker_forward(pd, res, res, mpw)
            exfunc_ref.masker_backward(gm, pd, eps, thr2)
        out["mask_fwd_bwd_us_reference_wrapper"] = timed(ref_mask, reps=5)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
