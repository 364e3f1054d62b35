"""Experiment harness (GPU box): build variants of the search kernel with -D flags, time the search stage alone,
and dump per-CTA clock traces.  Not part of the product path.   python tools/search_variants.py [C2]"""
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
}


def build_variant(name, flags):
    out = os.path.join(ROOT, "gpurun_out", f"librma_{name}.so")
    cmd = [b.nvcc_path()] + b.NVCC_FLAGS + flags + ["-o", out] + b.SOURCES
    subprocess.run(cmd, cwd=b.CSRC, check=True, capture_output=True)
    return out


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    B, N, M, _, _ = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    x, y = make_workload(wl)
    xd, yd = x.to(dev), y.to(dev)
    res = {}
    for name, flags in VARIANTS.items():
        lib = ctypes.CDLL(build_variant(name, flags))
        for fn, (rt, at) in _lib.SIGNATURES.items():
            f = getattr(lib, fn); f.restype = rt; f.argtypes = at
        lib.rma_exp_set_search_debug.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        wsb = lib.rma_chamfer_workspace_bytes(B, N, M)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        xs, ys = xd.stride(), yd.stride()
        for splits, stagger in [(0, 0), (0, 300), (0, 600), (0, 1000), (0, 1500), (0, 2500), (8, 0), (8, 1000), (16, 0), (16, 1000)]:
            trace = torch.zeros(4096 * 5, dtype=torch.int64, device=dev)
            lib.rma_exp_set_search_debug(P(trace) if splits == 0 else None, splits, stagger)
            times = []
            for it in range(12):
                lib.rma_chamfer_prep(P(xd), xs[0], xs[1], xs[2], P(yd), ys[0], ys[1], ys[2], B, N, M, P(ws), wsb, st)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.rma_chamfer_search(B, N, M, P(ws), wsb, st)
                e1.record()
                torch.cuda.synchronize()
                assert rc == 0, rc
                if it >= 2:
                    times.append(e0.elapsed_time(e1) * 1e3)
            key = f"{name}/S={splits or 'auto'}/stagger={stagger}"
            res[key] = {"us_min": min(times), "us_med": sorted(times)[len(times) // 2]}
            if splits == 0:
                t = trace.cpu().view(-1, 5)
                n = int((t[:, 2] != 0).sum())
                t = t[:n]
                pro = (t[:, 1] - t[:, 0]).float()
                tot = (t[:, 2] - t[:, 0]).float()
                gt = t[:, 4]
                res[key].update(ctas=n, prologue_cyc_mean=pro.mean().item(), prologue_cyc_max=pro.max().item(),
                                total_cyc_mean=tot.mean().item(), total_cyc_min=tot.min().item(),
                                total_cyc_max=tot.max().item(),
                                end_spread_us=(gt.max() - gt.min()).item() / 1e3,
                                ctas_per_sm_max=int(torch.bincount(t[:, 3]).max()))
            print(key, json.dumps(res[key]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"search_variants_{wl}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
