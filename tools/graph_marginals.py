"""Marginal in-graph cost of each kernel of the C2 step: CUDA graphs of growing prefixes of the pipeline, replayed
back to back (L2-warm), time per replay.    python tools/graph_marginals.py [C2]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200 import _lib  # noqa: E402
from rma_net_b200.model.loss import loss_on_cd  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    B, N, M, _, _ = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    x, y = make_workload(wl)
    xd, yd = x.to(dev), y.to(dev)
    lib = _lib.load()
    P = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    wsb = lib.rma_chamfer_workspace_bytes(B, N, M, None)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    gx = torch.empty(B, N, 3, device=dev); gy = torch.empty(B, M, 3, device=dev)
    sx = torch.empty(B, N, 4, device=dev); sy = torch.empty(B, M, 4, device=dev)
    ox = torch.empty_like(gx); oy = torch.empty_like(gy)
    loss = torch.empty((), device=dev)
    one = torch.ones((), device=dev)
    xs, ys = xd.stride(), yd.stride()
    xr, yr = xd.clone().requires_grad_(True), yd.clone().requires_grad_(True)

    def st():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def k_prep():
        assert lib.rma_chamfer_prep(P(xd), *xs, P(yd), *ys, B, N, M, P(ws), wsb, None, st()) == 0

    def k_search():
        assert lib.rma_chamfer_search(B, N, M, P(ws), wsb, None, st()) == 0

    def k_fwd():
        assert lib.rma_chamfer_cd_forward(P(xd), *xs, P(yd), *ys, B, N, M, P(loss), None, None, None, None,
                                          P(gx), P(sx), P(gy), P(sy), P(ws), wsb, None, st()) == 0

    d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)

    def k_fin():
        assert lib.rma_chamfer_finalize(B, N, M, P(d1), P(d2), P(i1), P(i2), P(ws), wsb, None, st()) == 0

    def k_bwd():
        assert lib.rma_chamfer_cd_backward(P(gx), P(sx), B * N, P(gy), P(sy), B * M, P(one), P(ox), P(oy), st()) == 0

    def k_step():
        xr.grad = None; yr.grad = None
        loss_on_cd(xr, yr).backward()

    variants = {
        "search": [k_search],
        "prep+search": [k_prep, k_search],
        "prep+search+resolve (cd_forward)": [k_fwd],
        "cd_forward+cd_backward": [k_fwd, k_bwd],
        "public API step (adds autograd's seed fill)": [k_step],
        "prep+search+finalize (dist, idx; no gradients)": [k_prep, k_search, k_fin],
        "finalize x8 (same records)": [k_fin] * 8,
        "prep": [k_prep],
        "cd_backward": [k_bwd],
    }
    k_prep(); k_search(); k_fwd(); k_bwd(); k_step()
    torch.cuda.synchronize()
    res = {}
    for name, fns in variants.items():
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for f in fns:
                f()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        reps = 200
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # gate: a spin kernel in front of the start event, so that the host has enqueued all replays before the
            # first one runs (a replay costs the host ~20 us: without the gate every graph shorter than that reads 20.5 us)
            torch.cuda._sleep(int(reps * 30e-6 * 2.0e9))
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        res[name] = round(best, 2)
        print(f"{name:50s} {best:8.2f} us per replay", flush=True)
        if name.startswith("public API"):
            # the bench's protocol on the same graph: one event pair per replay, without / with the 256 MiB L2 flush
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            for label, do_flush in (("per-replay events, no flush", False), ("per-replay events, L2 flush", True)):
                ts = []
                for _ in range(60):
                    if do_flush:
                        flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record()
                    ts.append((e0, e1))
                torch.cuda.synchronize()
                v = sorted(a.elapsed_time(b) * 1e3 for a, b in ts[10:])
                res[label] = round(v[len(v) // 2], 2)
                print(f"{label:50s} {v[len(v) // 2]:8.2f} us (median)", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"graph_marginals_{wl}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
