from .chamfer_func import (ArapFunction, Chamfer, ChamferCDFunction, ChamferCDMetricsFunction,  # noqa: F401
                           ChamferFunction, DepthViewTermFunction, ViewProjectFunction)
from .warp_func import Warp, WarpFunction  # noqa: F401
from .point_masker_func import Masker, Masker_Point  # noqa: F401
from .point_render_func import Render, Render_Point  # noqa: F401
