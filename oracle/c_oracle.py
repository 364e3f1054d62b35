"""ctypes wrapper of oracle/chamfer_ref.c (TEST INFRASTRUCTURE).  `make -C oracle` builds the library."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libchamfer_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "chamfer_ref.c")
    if force or not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libchamfer_oracle.so"], check=True, capture_output=True)
    return _PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def chamfer_fwd(x: np.ndarray, y: np.ndarray, precision: str = "f64"):
    """x [B,N,3], y [B,M,3] float32 -> dict(dist1, dist2, idx1, idx2); 'f64' or 'f32' (GPU operation order)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    B, N, M = x.shape[0], x.shape[1], y.shape[1]
    dt = np.float64 if precision == "f64" else np.float32
    d1, d2 = np.empty((B, N), dt), np.empty((B, M), dt)
    i1, i2 = np.empty((B, N), np.int32), np.empty((B, M), np.int32)
    fn = load().chamfer_fwd_f64 if precision == "f64" else load().chamfer_fwd_f32
    fn(_p(x), _p(y), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(M), _p(d1), _p(d2), _p(i1), _p(i2))
    return dict(dist1=d1, dist2=d2, idx1=i1, idx2=i2)


def chamfer_bwd(x, y, idx1, idx2, g1, g2):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    B, N, M = x.shape[0], x.shape[1], y.shape[1]
    g1 = np.ascontiguousarray(np.broadcast_to(g1, (B, N)), np.float64)
    g2 = np.ascontiguousarray(np.broadcast_to(g2, (B, M)), np.float64)
    idx1 = np.ascontiguousarray(idx1, np.int32)
    idx2 = np.ascontiguousarray(idx2, np.int32)
    gx, gy = np.empty((B, N, 3), np.float64), np.empty((B, M, 3), np.float64)
    load().chamfer_bwd_f64(_p(x), _p(y), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(M), _p(g1), _p(g2),
                           _p(idx1), _p(idx2), _p(gx), _p(gy))
    return gx, gy
