"""loss_on_depth / loss_on_mask (loss.py:147-196), 25 views, 256x256: per-view loop (2 x 25 kernel-call groups, as the
reference structures it) vs views folded into the batch (2 calls), fwd+bwd, on the same B200 kernels.
    python tools/bench_view_losses.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200.model.loss import loss_on_depth, loss_on_mask, view_rotations  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (B, N) in ((4, 2048), (16, 2048)):
    g = torch.Generator().manual_seed(1)
    src = (torch.randn(B, N, 3, generator=g) * 0.3).clamp(-0.9, 0.9).to(dev).requires_grad_(True)
    tgt = (torch.randn(B, N, 3, generator=g) * 0.3).clamp(-0.9, 0.9).to(dev)
    rot = view_rotations(5, dev, jitter=torch.rand(25, 2, generator=g))
    row = {"config": f"B={B} N={N} 25 views 256x256"}
    for name, fn in (("depth", loss_on_depth), ("mask", loss_on_mask)):
        for mode in (False, True):
            def run():
                src.grad = None
                fn(src, tgt, rot, batched=mode).backward()
            row[f"{name}_ms_{'batched' if mode else 'per_view_loop'}"] = round(timed(run), 3)
    print(json.dumps(row), flush=True)
