"""Device timeline of one chamfer forward (prep -> search -> resolve) inside a replayed CUDA graph, from %globaltimer
marks the kernels leave in the workspace when the library is built with -DRMA_B200_TIMELINE (tools/ build only):
per phase the earliest and latest CTA.      python tools/step_timeline.py [C2]
Build:  cd rma_net_b200/csrc && nvcc <build.NVCC_FLAGS> -DRMA_B200_TIMELINE -o ../lib/librma_b200_tl.so <sources>"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200 import _lib  # noqa: E402

SLOTS = ["prep start", "prep end", "search start", "search end", "resolve entry (before dependency wait)",
         "resolve start", "resolve end"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    path = os.environ.get("RMA_B200_TL_LIB") or os.path.join(ROOT, "rma_net_b200", "lib", "librma_b200_tl.so")
    lib = _lib.bind(ctypes.CDLL(path))
    lib.rma_debug_timeline_read.restype = ctypes.c_int
    lib.rma_debug_timeline_read.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p]
    B, N, M, _, _ = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    x, y = make_workload(wl)
    xd, yd = x.to(dev), y.to(dev)
    P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    wsb = lib.rma_chamfer_workspace_bytes(B, N, M, None)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    gx = torch.empty(B, N, 3, device=dev); gy = torch.empty(B, M, 3, device=dev)
    sx = torch.empty(B, N, 4, device=dev); sy = torch.empty(B, M, 4, device=dev)
    loss = torch.empty((), device=dev)
    xs, ys = xd.stride(), yd.stride()

    def st():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fwd():
        assert lib.rma_chamfer_cd_forward(P(xd), *xs, P(yd), *ys, B, N, M, P(loss), None, None, None, None,
                                          P(gx), P(sx), P(gy), P(sy), P(ws), wsb, None, st()) == 0

    fwd()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd()
    out = (ctypes.c_ulonglong * 16)()
    rows = []
    for it in range(12):
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        assert lib.rma_debug_timeline_read(P(ws), B, N, M, out, st()) == 0      # read + reset
        g.replay()        # ONE replay between a reset and the read: the marks are min / max over the CTAs of this replay
        torch.cuda.synchronize()
        assert lib.rma_debug_timeline_read(P(ws), B, N, M, out, st()) == 0
        v = [int(t) for t in out]
        t0 = v[0]
        rows.append([((v[2 * k] - t0) / 1e3, (v[2 * k + 1] - t0) / 1e3) for k in range(len(SLOTS))])
    med = []
    for k, name in enumerate(SLOTS):
        a = sorted(r[k][0] for r in rows)[len(rows) // 2]
        b = sorted(r[k][1] for r in rows)[len(rows) // 2]
        med.append((name, a, b))
        print(f"{name:42s} first CTA {a:8.2f} us   last CTA {b:8.2f} us")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"step_timeline_{wl}.json"), "w") as f:
        json.dump({n: {"first_cta_us": a, "last_cta_us": b} for n, a, b in med}, f, indent=1)


if __name__ == "__main__":
    main()
