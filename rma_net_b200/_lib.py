"""ctypes binding of include/rma_b200.h.  No fallback: if the CUDA library is missing or fails, this raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librma_b200.so")

_P = c_void_p
_I64 = c_int64


class ChamferOpts(ctypes.Structure):
    """`rma_chamfer_opts` of include/rma_b200.h: per-call choice of the search implementation (no library state)."""
    _fields_ = [("path", c_int32), ("unit_run", c_int32), ("record_cap_bytes", c_int64),
                ("y_batch_period", c_int32), ("reserved_", c_int32)]


PATH_AUTO, PATH_EXACT, PATH_FILTER = 0, 1, 2
_O = ctypes.POINTER(ChamferOpts)

# name -> (restype, argtypes); mirrors include/rma_b200.h declaration by declaration.
SIGNATURES = {
    "rma_b200_abi_version": (c_int, []),
    "rma_b200_error_string": (c_char_p, [c_int]),
    "rma_b200_device_info": (c_int, [_P, _P, _P, _P]),
    "rma_chamfer_workspace_bytes": (c_size_t, [c_int, c_int, c_int, _O]),
    "rma_chamfer_path": (c_int, [c_int, c_int, c_int, _O]),
    "rma_chamfer_dealing_selfcheck": (c_int, [c_int, c_int, c_int, c_int, _O, _P, _P]),
    "rma_chamfer_guard_stats": (c_int, [c_int, c_int, c_int, _P, _O, _P, _P]),
    "rma_chamfer_forward": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, c_int, c_int, c_int,
                                    _P, _P, _P, _P, _P, c_size_t, _O, _P]),
    "rma_chamfer_prep": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, c_int, c_int, c_int, _P, c_size_t, _O,
                                 _P]),
    "rma_chamfer_search": (c_int, [c_int, c_int, c_int, _P, c_size_t, _O, _P]),
    "rma_chamfer_finalize": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _O, _P]),
    "rma_chamfer_search_config": (c_int, [c_int, c_int, c_int, _O, _P, _P, _P, _P]),
    "rma_chamfer_cd_forward": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, c_int, c_int, c_int,
                                       _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _O, _P]),
    "rma_chamfer_cd_forward_weighted": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, c_int, c_int, c_int,
                                                _P, c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _O,
                                                _P]),
    "rma_chamfer_cd_backward": (c_int, [_P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P]),
    "rma_chamfer_backward": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, c_int, c_int, c_int,
                                     _P, _P, _P, _P, _P, _P, _P]),
    "rma_sqrdis_map": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, c_int, c_int, c_int, _P, _P]),
    "rma_pack_min_keys": (c_int, [_P, _P, _I64, c_int32, _P, _P]),
    "rma_unpack_min_keys": (c_int, [_P, _I64, _P, _P, _P]),
    "rma_se3_exp_forward": (c_int, [_P, _I64, _P, _P]),
    "rma_se3_exp_backward": (c_int, [_P, _P, _I64, _P, _P]),
    "rma_se3_transform_forward": (c_int, [_P, _P, _I64, _I64, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P]),
    "rma_se3_transform_backward": (c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I64, _I64, _P,
                                           _P, _I64, _I64, _I64, _P, _P]),
    "rma_warp_forward": (c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64, c_int, c_int,
                                 _P, _I64, _I64, _I64, _P, _P, _P]),
    "rma_warp_backward": (c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I64,
                                  _P, c_int, c_int, _P, _P, _P, _I64, _I64, _I64, _P, _P]),
    "rma_point_render_forward": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                         _P, _P, _P, _P, _P, _P, _P, _P]),
    "rma_point_render_backward": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "rma_point_masker_forward": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "rma_point_masker_backward_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rma_point_masker_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_float, c_float, _P, _P, c_size_t, _P]),
    "rma_knn": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "rma_arap_forward": (c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P, c_int, c_int, c_int, c_int,
                                 _P, c_float, _P, _P, _P]),
    "rma_depth_view_term": (c_int, [_P, _P, _P, _P, _I64, c_float, c_float, _P, _P, _P, _P]),
    "rma_view_project_forward": (c_int, [_P, _I64, _I64, _I64, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "rma_view_project_backward": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "rma_chamfer_loss_host": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


class RmaB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load librma_b200.so (built by `python -m rma_net_b200.build`).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RMA_B200_LIB") or LIB_PATH     # override: A/B runs of two builds of this same library
    if not os.path.exists(path):
        raise RmaB200Error(
            f"{path} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m rma_net_b200.build` (needs nvcc, sm_100a).")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    """Declare the header's signatures on another build of this library (tools/: -D variants of one kernel)."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().rma_b200_error_string(code)
        raise RmaB200Error(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")
