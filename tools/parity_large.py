"""Filter path vs exact path, bit for bit, at sizes where a search CTA's run spans several row-direction entries
(auto unit run; uniform and tie-heavy clouds).    python tools/parity_large.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rma_net_b200 import _lib, ext  # noqa: E402
from rma_net_b200.model.loss import chamfer_dist_with_idx  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
rng = np.random.default_rng(11)
ok = True
for (B, N, M) in [(16, 8192, 8192), (4, 16384, 12000), (1, 65536, 65536), (2, 30000, 50001)]:
    for kind in ("uniform", "lattice", "surface"):
        def cloud(n):
            if kind == "uniform":
                p = rng.uniform(-1, 1, (B, n, 3))
            elif kind == "lattice":
                p = rng.integers(-20, 21, (B, n, 3)).astype(np.float64) * 0.05
            else:
                v = rng.normal(size=(B, n, 3)); p = v / np.linalg.norm(v, axis=-1, keepdims=True)
            return torch.from_numpy(p.astype(np.float32)).to(dev)
        x, y = cloud(N), cloud(M)
        outs = []
        for path in (2, 1):
            with ext.chamfer_options(path=path):
                outs.append(chamfer_dist_with_idx(x, y))
        same = all(torch.equal(a, b) for a, b in zip(*outs))
        print(B, N, M, kind, "identical" if same else "MISMATCH", flush=True)
        ok &= same
print("ALL IDENTICAL" if ok else "FAILED")
sys.exit(0 if ok else 1)
