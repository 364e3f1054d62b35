"""Public-API step (loss_on_cd(x, y).backward()) captured n times per CUDA graph: per-step device time without the
graph-to-graph launch gap.    python tools/step_probe.py S16x2048 [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200.model.loss import loss_on_cd  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    B, N, M, _, _ = WORKLOADS[wl]
    dev = torch.device("cuda:0")
    x, y = make_workload(wl)
    xr, yr = x.to(dev).requires_grad_(True), y.to(dev).requires_grad_(True)

    def step():
        xr.grad = None; yr.grad = None
        loss_on_cd(xr, yr).backward()

    out = {"workload": wl}
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    for per in (1, n):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(per):
                step()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        best = 1e9
        reps = 50
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(int(reps * 60e-6 * 2.0e9))
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps / per)
        out[f"step_us_{per}_per_graph"] = round(best, 2)
    pairs = B * N * M
    out["pct_fp32_peak_1_per_graph"] = round(100 * 12 * pairs / (out["step_us_1_per_graph"] * 1e-6) / 74.45e12, 1)
    out[f"pct_fp32_peak_{n}_per_graph"] = round(100 * 12 * pairs / (out[f"step_us_{n}_per_graph"] * 1e-6) / 74.45e12, 1)
    print(out, flush=True)


if __name__ == "__main__":
    main()
