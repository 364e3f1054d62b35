"""HBM-bound kernels of the path at a size where bytes dominate launch latency (2^22 points per call):
warp forward / backward (network.py:652-655 semantics) and the generic chamfer backward scatter.
Reports achieved ALGORITHMIC GB/s (warp fwd 40 B/point, 24 at stage 0; chamfer backward
56 B/point) against MEASURED_PEAKS.json's HBM copy bandwidth.   python tools/bench_hbm_kernels.py [log2_points]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200 import ext  # noqa: E402
from rma_net_b200.model.exfunc.functions import Warp  # noqa: E402


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
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    B, N = 16, (1 << lg) // 16
    dev = torch.device("cuda:0")
    peak = 6552.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    g = torch.Generator().manual_seed(0)
    src = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    prev = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    phi = (torch.randn(B, 6, generator=g) * 0.1).to(dev)
    w = torch.rand(B, 1, N, generator=g).to(dev)
    out = {"points": B * N, "hbm_peak_GBps": peak}
    lib = __import__("rma_net_b200._lib", fromlist=["load"]).load()
    import ctypes
    P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    o = torch.empty_like(src)
    go = torch.randn(B, N, 3, generator=g).to(dev)
    gphi = torch.empty(B, 6, device=dev)
    gw = torch.empty(B, N, device=dev)
    scratch = torch.empty(B * 12, dtype=torch.float64, device=dev)
    ss = src.stride()

    def warp_fwd():
        lib.rma_warp_forward(P(phi), P(src), ss[0], ss[1], ss[2], P(w), P(prev), ss[0], ss[1], ss[2], B, N, P(o),
                             ss[0], ss[1], ss[2], None, None, st())

    def warp_fwd0():
        lib.rma_warp_forward(P(phi), P(src), ss[0], ss[1], ss[2], None, None, 0, 0, 0, B, N, P(o), ss[0], ss[1], ss[2],
                             None, None, st())

    def warp_bwd():
        lib.rma_warp_backward(P(phi), P(src), ss[0], ss[1], ss[2], P(w), P(prev), ss[0], ss[1], ss[2], P(go), None,
                              ss[0], ss[1], ss[2], None, B, N, P(gphi), P(gw), None, 0, 0, 0, P(scratch), st())
    t = timed(warp_fwd); out["warp_fwd_GBps"] = 40 * B * N / t / 1e9
    t = timed(warp_fwd0); out["warp_fwd_stage0_GBps"] = 24 * B * N / t / 1e9
    t = timed(warp_bwd); out["warp_bwd_GBps"] = (12 + 12 + 4 + 12 + 4) * B * N / t / 1e9    # src, prev, w, go in; grad_w out
    # generic chamfer backward scatter: 56 B/point (g 4 + idx 4 + own xyz 12 + partner gather 12 + own grad 12 + atomic 12)
    y = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    idx1 = torch.randint(0, N, (B, N), generator=g, dtype=torch.int32).to(dev)
    idx2 = torch.randint(0, N, (B, N), generator=g, dtype=torch.int32).to(dev)
    g1 = torch.rand(B, N, generator=g).to(dev)
    g2 = torch.rand(B, N, generator=g).to(dev)
    t = timed(lambda: ext.chamfer_backward(src, y, idx1, idx2, g1, g2))
    out["chamfer_backward_random_idx_GBps"] = 56 * 2 * B * N / t / 1e9
    near = (torch.arange(N, dtype=torch.int32)[None].expand(B, -1).contiguous()).to(dev)   # local partners: coalesced
    t = timed(lambda: ext.chamfer_backward(src, y, near, near, g1, g2))
    out["chamfer_backward_local_idx_GBps"] = 56 * 2 * B * N / t / 1e9
    # ---- kernels of the Loss module (csrc/loss_terms.cu), algorithmic bytes as stated in DESIGN.md section 12 ----
    # depth image comparison: 4 images read (16 B/pixel) + 2 gradient images written (8 B/pixel); 200 images 256x256
    npx = 200 * 256 * 256
    imgs = [torch.rand(npx, generator=g).to(dev) + 0.5 for _ in range(4)]
    t = timed(lambda: ext.depth_view_term(imgs[0], imgs[1], imgs[2], imgs[3], 0.05, 1.0, True))
    out["depth_view_term_GBps"] = 24 * npx / t / 1e9          # + two torch.empty and one 4-byte memset per call
    # multi-view projection, V = 25: forward 12 B read + 12 V B written per point; backward the reverse
    Bp, Np, V = 8, 1 << 15, 25
    pts = (torch.rand(Bp, Np, 3, generator=g) - 0.5).to(dev)
    rot = torch.linalg.qr(torch.randn(V, 3, 3, generator=g))[0].reshape(-1, 3).contiguous().to(dev)
    add3, mul3 = (1., 1., -1.), (127.5, 127.5, 1.)
    t = timed(lambda: ext.view_project_forward(pts, rot, add3, mul3))
    out["view_project_fwd_GBps"] = 12 * (V + 1) * Bp * Np / t / 1e9
    gproj = torch.randn(V * Bp, Np, 3, generator=g).to(dev)
    t = timed(lambda: ext.view_project_backward(gproj, rot, Bp, Np, mul3))
    out["view_project_bwd_GBps"] = 12 * (V + 1) * Bp * Np / t / 1e9
    # ARAP, n = 4, T*B = 56 groups: per (group, vertex) 60 B deformation + 60 B source + 16 B indices + 60 B gradient
    # updates = 196 B (the source and the indices are shared by the T stages and the neighbour reads hit L2: the
    # figure is algorithmic, not DRAM traffic)
    Ba, Va, T = 8, 1 << 15, 7
    srca = (torch.rand(Ba, Va, 3, generator=g) - 0.5).to(dev)
    defo = (srca.repeat(T, 1, 1) + 0.01 * torch.randn(T * Ba, Va, 3, generator=g).to(dev)).contiguous()
    from rma_net_b200.model.loss import get_neighbor_index
    nb = get_neighbor_index(srca, 4).int()
    wa = torch.ones(T * Ba, device=dev)
    t = timed(lambda: ext.arap_forward(defo, srca, nb, wa, 1.0 / Ba, True))
    out["arap_fwd_with_grad_GBps"] = 196 * T * Ba * Va / t / 1e9     # includes the zero-fill of the gradient buffer
    for k in list(out):
        if k.endswith("GBps") and k != "hbm_peak_GBps":
            out[k.replace("GBps", "frac_of_peak")] = out[k] / peak
    print(json.dumps(out))


if __name__ == "__main__":
    main()
