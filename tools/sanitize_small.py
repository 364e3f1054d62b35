"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel once on ragged small shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rma_net_b200.model.exfunc.functions import Warp  # noqa: E402
from rma_net_b200.model.LieAlgebra import se3  # noqa: E402
from rma_net_b200.model.loss import chamfer_dist_with_idx, compute_sqrdis_map, loss_on_cd  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (B, N, M) in [(2, 700, 1300), (1, 2100, 95), (3, 33, 2050)]:
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev).requires_grad_(True)
    y = (torch.rand(B, M, 3, generator=g) * 2 - 1).to(dev).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_dist_with_idx(x, y)
    (d1.sum() + d2.sum()).backward()
    loss_on_cd(x, y).backward()
    compute_sqrdis_map(x.detach(), y.detach())
    phi = (torch.randn(B, 6, generator=g) * 0.2).to(dev).requires_grad_(True)
    w = torch.rand(B, 1, N, generator=g).to(dev).requires_grad_(True)
    out, rigid, gm = Warp()(phi, x.detach(), w, x.detach() * 0.5)
    out.sum().backward()
    gmat = se3.Exp(phi).view(B, 1, 4, 4)
    se3.transform(gmat, x.detach()).sum().backward()
torch.cuda.synchronize()
print("sanitize run ok")
