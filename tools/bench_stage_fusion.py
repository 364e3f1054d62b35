"""Stage loop of Loss.forward (loss.py:260-267): per-stage calls vs one batched call, fwd+bwd, eager.
    python tools/bench_stage_fusion.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

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
    print(json.dumps({"config": f"T={T} B={B} N=M={N}", "ms_loop": a, "ms_fused": b, "speedup": a / b,
                      "G_pairs_per_s_fused": T * B * N * N / b / 1e6}))
