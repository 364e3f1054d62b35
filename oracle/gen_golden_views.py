"""tests/golden/view_grid.npz: the view rotations of the REFERENCE's Loss.__init__ (regular grid, loss.py:136-143) and of
loss_on_depth's per-call re-sampling (loss.py:151-160), computed by exec'ing those lines VERBATIM on the CPU
("cuda:0" -> "cpu"), together with the reference's `torch.pi` (loss.py:29) and euler2rot (loss.py:73-95).
   python oracle/gen_golden_views.py"""
import os
import random
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference/model/loss.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _lines(first, last, dedent=0):
    with open(REF) as f:
        ls = f.readlines()[first - 1:last]
    return "".join(l[dedent:] if l.strip() else l for l in ls).replace('"cuda:0"', '"cpu"')


def main():
    out = {}
    for vd in (2, 5):
        ns = {"torch": torch, "random": random}
        exec(compile(_lines(29, 29) + _lines(73, 95), REF, "exec"), ns)
        self = SimpleNamespace(args=SimpleNamespace(view_divide=vd))
        ns["self"], ns["args"] = self, self.args
        exec(compile(_lines(136, 143, dedent=8), REF, "exec"), ns)             # Loss.__init__: the regular grid
        out[f"grid_{vd}"] = self.all_rot_matrix_for_lfd.numpy().copy()
        random.seed(7 + vd)
        exec(compile(_lines(151, 160, dedent=8), REF, "exec"), ns)             # loss_on_depth: jittered re-sampling
        out[f"jitter_{vd}_seed{7 + vd}"] = self.all_rot_matrix_for_lfd.numpy().copy()
    out["torch_pi"] = np.float64(ns["torch"].pi)
    np.savez_compressed(os.path.join(GOLD, "view_grid.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
