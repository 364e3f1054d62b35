"""Small driver for ncu: a few eager chamfer fwd+bwd steps (and one warp stage) at a BASELINE.json config.
    python tools/profile_step.py [C2|C3|C4] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS, make_workload  # noqa: E402
from rma_net_b200.model.exfunc.functions import Warp  # noqa: E402
from rma_net_b200.model.loss import loss_on_cd  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
x, y = make_workload(name)
xd, yd = x.to(dev).requires_grad_(True), y.to(dev).requires_grad_(True)
B, N = xd.shape[0], xd.shape[1]
phi = (torch.randn(B, 6) * 0.1).to(dev).requires_grad_(True)
w = torch.rand(B, 1, N).to(dev).requires_grad_(True)
warp = Warp()
for _ in range(steps):
    xd.grad = None
    yd.grad = None
    out, _, _ = warp(phi, xd.detach(), w, yd.detach()[:, :N])
    loss = loss_on_cd(xd, yd) + loss_on_cd(out, yd)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
