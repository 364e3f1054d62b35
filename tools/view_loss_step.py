"""One loss_on_depth + one loss_on_mask forward+backward (25 views folded into the batch), for ncu launch lists.
    python tools/view_loss_step.py [B N reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rma_net_b200.model.loss import loss_on_depth, loss_on_mask, view_rotations  # noqa: E402

B, N, reps = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 8192, 1)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
source = (torch.rand(B, N, 3, generator=g) - 0.5)
target = (source + 0.02 * torch.randn(B, N, 3, generator=g)).to(dev)
x = (source + 0.01 * torch.randn(B, N, 3, generator=g)).to(dev).requires_grad_(True)
rot = view_rotations(5, dev, jitter=torch.rand(25, 2, generator=g))
for name, fn in (("depth", loss_on_depth), ("mask", loss_on_mask)):
    for _ in range(reps):
        x.grad = None
        fn(x, target, rot).backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x.grad = None
    fn(x, target, rot).backward()
    e1.record()
    torch.cuda.synchronize()
    print(name, "fwd+bwd ms", e0.elapsed_time(e1), flush=True)
