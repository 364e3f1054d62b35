"""Experiment harness (GPU box): build -D variants of the library and time the fused resolve (cd_forward minus
prep+search, L2-warm as inside a real step) at a BASELINE config.  Not part of the product path.
    python tools/resolve_variants.py [C2] [variant ...]"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200 import _lib, build as b  # noqa: E402

VARIANTS = {
    "base": [],
    "noslow": ["-DRMA_B200_EXPERIMENTS", "-DRMA_EXP_NO_SLOW"],
    "noslow_rows": ["-DRMA_B200_EXPERIMENTS", "-DRMA_EXP_NO_SLOW_ROWS"],
    "noslow_cols": ["-DRMA_B200_EXPERIMENTS", "-DRMA_EXP_NO_SLOW_COLS"],
}


def build_variant(name, flags):
    out = os.path.join(ROOT, "gpurun_out", f"librma_{name}.so")
    cmd = [b.nvcc_path()] + b.NVCC_FLAGS + flags + ["-o", out] + b.SOURCES
    subprocess.run(cmd, cwd=b.CSRC, check=True, capture_output=True, text=True)
    return out


def med(fn, reps=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    which = sys.argv[2:] or list(VARIANTS)
    B, N, M, _, _ = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    x, y = make_workload(wl)
    xd, yd = x.to(dev), y.to(dev)
    res = {}
    for name in which:
        lib = _lib.bind(ctypes.CDLL(build_variant(name, VARIANTS[name])))
        P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        xs, ys = xd.stride(), yd.stride()
        wsb = lib.rma_chamfer_workspace_bytes(B, N, M, None)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        gx = torch.empty(B, N, 3, device=dev); gy = torch.empty(B, M, 3, device=dev)
        sx = torch.empty(B, N, 4, device=dev); sy = torch.empty(B, M, 4, device=dev)
        loss = torch.empty((), device=dev)

        def ps():
            assert lib.rma_chamfer_prep(P(xd), *xs, P(yd), *ys, B, N, M, P(ws), wsb, None, st) == 0
            assert lib.rma_chamfer_search(B, N, M, P(ws), wsb, None, st) == 0

        def full():
            assert lib.rma_chamfer_cd_forward(P(xd), *xs, P(yd), *ys, B, N, M, P(loss), None, None, None, None,
                                              P(gx), P(sx), P(gy), P(sy), P(ws), wsb, None, st) == 0

        a = med(ps); f = med(full)
        res[name] = {"prep_search_us": [round(v, 2) for v in a], "cd_forward_us": [round(v, 2) for v in f],
                     "resolve_us_med": round(f[1] - a[1], 2), "loss": float(loss)}
        print(name, json.dumps(res[name]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"resolve_variants_{wl}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
