"""Warm per-stage timings (CUDA events, L2-warm, as inside a real step) of the chamfer pipeline at a BASELINE config.
    python tools/stage_timing.py [C2|C3|C4]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200 import _lib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    B, N, M, _, _ = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    x, y = make_workload(wl)
    xd, yd = x.to(dev), y.to(dev)
    lib = _lib.load()
    P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    wsb = lib.rma_chamfer_workspace_bytes(B, N, M, None)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    gx = torch.empty(B, N, 3, device=dev); gy = torch.empty(B, M, 3, device=dev)
    g1 = torch.full((B, N), 0.5 / B, device=dev); g2 = torch.full((B, M), 0.5 / B, device=dev)
    loss = torch.empty((), device=dev)
    one = torch.ones((), device=dev)
    ox = torch.empty_like(gx); oy = torch.empty_like(gy)
    sx = torch.empty(B, N, 4, device=dev); sy = torch.empty(B, M, 4, device=dev)
    xs, ys = xd.stride(), yd.stride()
    stages = {
        "prep": lambda: lib.rma_chamfer_prep(P(xd), *xs, P(yd), *ys, B, N, M, P(ws), wsb, None, st),
        "search": lambda: lib.rma_chamfer_search(B, N, M, P(ws), wsb, None, st),
        "finalize": lambda: lib.rma_chamfer_finalize(B, N, M, P(d1), P(d2), P(i1), P(i2), P(ws), wsb, None, st),
        "backward_generic(memset+scatter)": lambda: lib.rma_chamfer_backward(
            P(xd), *xs, P(yd), *ys, B, N, M, P(g1), P(g2), P(i1), P(i2), P(gx), P(gy), st),
        "cd_forward(fused prep+search+finalize+loss+grads)": lambda: lib.rma_chamfer_cd_forward(
            P(xd), *xs, P(yd), *ys, B, N, M, P(loss), None, None, None, None, P(gx), P(sx), P(gy), P(sy), P(ws), wsb, None, st),
        "cd_backward": lambda: lib.rma_chamfer_cd_backward(P(gx), P(sx), B * N, P(gy), P(sy), B * M, P(one), P(ox), P(oy), st),
    }
    res = {}
    for name, fn in stages.items():
        for _ in range(3):
            assert fn() == 0
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        res[name] = {"us_min": round(ts[0], 2), "us_med": round(ts[len(ts) // 2], 2)}
        print(f"{name:55s} min {ts[0]:8.2f} us   median {ts[len(ts)//2]:8.2f} us", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"stage_timing_{wl}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
