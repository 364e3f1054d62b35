"""rma_net_b200 — B200-native (sm_100a) implementation of RMA-Net's data-parallel hot path:
batched chamfer nearest-neighbour search + backward, and the se(3) exp-map warp of the recurrent alignment loop.

Drop-in names (same as the reference):
    rma_net_b200.model.loss.chamfer_dist / compute_sqrdis_map / loss_on_cd      (reference model/loss.py)
    rma_net_b200.model.LieAlgebra.se3.Exp / exp / transform / mat / genmat      (reference model/LieAlgebra/se3.py)
    rma_net_b200.model.exfunc.functions.chamfer_func.Chamfer                    (exfunc-style autograd.Function)

The directory is spelled `rma_net_b200` (the hyphen of "rma-net" is not importable in Python).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
