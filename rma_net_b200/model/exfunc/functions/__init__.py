from .chamfer_func import Chamfer, ChamferCDFunction, ChamferFunction  # noqa: F401
from .warp_func import Warp, WarpFunction  # noqa: F401
