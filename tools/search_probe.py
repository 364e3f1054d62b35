"""Search kernel alone, n launches per CUDA graph (so that neither the host's replay rate nor the graph-to-graph gap
is in the figure), plus the dealing the library chose.    python tools/search_probe.py S16x2048 [n]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200 import _lib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    B, N, M, _, _ = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    x, y = make_workload(wl)
    xd, yd = x.to(dev), y.to(dev)
    lib = _lib.load()
    P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    wsb = lib.rma_chamfer_workspace_bytes(B, N, M, None)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    c = [ctypes.c_int() for _ in range(4)]
    lib.rma_chamfer_search_config(B, N, M, None, *[ctypes.byref(v) for v in c])
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    xs, ys = xd.stride(), yd.stride()
    assert lib.rma_chamfer_prep(P(xd), *xs, P(yd), *ys, B, N, M, P(ws), wsb, None, st()) == 0
    out = {"workload": wl, "ctas": c[0].value, "threads": c[1].value, "longest_run_units": c[2].value, "ctas_per_sm": c[3].value}
    for reps_in_graph in (1, n):
        for _ in range(2):
            assert lib.rma_chamfer_search(B, N, M, P(ws), wsb, None, st()) == 0
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps_in_graph):
                assert lib.rma_chamfer_search(B, N, M, P(ws), wsb, None, st()) == 0
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        best = 1e9
        reps = 50
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(int(reps * 40e-6 * 2.0e9))
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps / reps_in_graph)
        out[f"search_us_{reps_in_graph}_per_graph"] = round(best, 2)
    pairs = B * N * M
    out["pct_fp32_peak_n_per_graph"] = round(100 * 12 * pairs / (out[f"search_us_{n}_per_graph"] * 1e-6) / 74.45e12, 1)
    print(out, flush=True)


if __name__ == "__main__":
    main()
