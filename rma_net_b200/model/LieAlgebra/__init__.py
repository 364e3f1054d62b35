"""Hot-path subset of the reference's model/LieAlgebra package (se3.Exp, se3.exp, se3.transform and helpers)."""
from . import se3  # noqa: F401
