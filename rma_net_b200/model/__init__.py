"""Host-side mirror of the reference's `model` package for the hot path only (loss.chamfer_dist, LieAlgebra.se3)."""
from . import LieAlgebra, loss  # noqa: F401
