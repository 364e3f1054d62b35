"""Stage loop of Loss.forward (loss.py:260-267): per-stage calls vs one batched call (eager), and the batched call
replayed as a CUDA graph (rma_net_b200.graphs.GraphedStep), fwd+bwd.
    python tools/bench_stage_fusion.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200 import ext  # noqa: E402
from rma_net_b200.graphs import GraphedStep  # noqa: E402
from rma_net_b200.model.loss import loss_cd_stages, loss_cd_stages_fused  # noqa: E402

dev = torch.device("cuda:0")


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
    return e0.elapsed_time(e1) / reps


info = ext.device_info()
peak_fma = info["sm_count"] * 128 * 1965e6          # FMA/s (bench.py: peak_source)

for (T, B, N) in ((7, 4, 2048), (7, 16, 2048), (7, 16, 4096), (8, 16, 8192)):
    g = torch.Generator().manual_seed(N)
    stages = [(torch.rand(B, 3, N, generator=g) * 2 - 1).to(dev).requires_grad_(True) for _ in range(T)]
    target = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)

    def run(fn):
        for s in stages:
            s.grad = None
        fn(stages, target, gamma=0.8).backward()
    a = timed(lambda: run(loss_cd_stages))
    b = timed(lambda: run(loss_cd_stages_fused))
    step = GraphedStep(lambda *st: loss_cd_stages_fused(list(st), target, gamma=0.8), stages)
    c = timed(step.replay, reps=100)
    pairs = T * B * N * N
    print(json.dumps({"config": f"T={T} B={B} N=M={N}", "ms_loop": a, "ms_fused": b, "speedup": a / b,
                      "ms_fused_graph": c, "G_pairs_per_s_fused": pairs / b / 1e6,
                      "G_pairs_per_s_fused_graph": pairs / c / 1e6,
                      "fp32_fma_peak_frac_fused": 6 * pairs / (b * 1e-3) / peak_fma,
                      "fp32_fma_peak_frac_fused_graph": 6 * pairs / (c * 1e-3) / peak_fma}), flush=True)
