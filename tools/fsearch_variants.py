"""Experiment harness (GPU box): build -D variants of the filter search kernel and time the search stage alone at
several unit runs.  Not part of the product path.   python tools/fsearch_variants.py [C2] [variant ...]"""
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
    "nofence": ["-DRMA_F_NO_FENCE"],              # the loop as ptxas schedules it by itself (FFMA2 / FMNMX3 alternating)
    "occ3": ["-DRMA_F_MINCTAS=3", "-DRMA_F_CHUNK=512"],
    "chunk512": ["-DRMA_F_CHUNK=512"],
}


def build_variant(name, flags):
    out = os.path.join(ROOT, "gpurun_out", f"librma_{name}.so")
    cmd = [b.nvcc_path()] + b.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", out] + b.SOURCES
    p = subprocess.run(cmd, cwd=b.CSRC, check=True, capture_output=True, text=True)
    regs = [l for l in p.stderr.splitlines() if "registers" in l]
    names = [l for l in p.stderr.splitlines() if "Compiling entry" in l]
    for n, r in zip(names, regs):
        if "fsearch" in n:
            print(name, r.strip(), flush=True)
    return out


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
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        xs, ys = xd.stride(), yd.stride()
        for run in (0, 14, 28):
            o = ctypes.byref(_lib.ChamferOpts(0, run, 0))
            wsb = lib.rma_chamfer_workspace_bytes(B, N, M, o)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            times = []
            for it in range(14):
                lib.rma_chamfer_prep(P(xd), xs[0], xs[1], xs[2], P(yd), ys[0], ys[1], ys[2], B, N, M, P(ws), wsb, o, st)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.rma_chamfer_search(B, N, M, P(ws), wsb, o, st)
                e1.record()
                torch.cuda.synchronize()
                assert rc == 0, rc
                if it >= 3:
                    times.append(e0.elapsed_time(e1) * 1e3)
            key = f"{name}/U={run or 'auto'}"
            res[key] = {"us_min": round(min(times), 2), "us_med": round(sorted(times)[len(times) // 2], 2)}
            print(key, json.dumps(res[key]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"fsearch_variants_{wl}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
