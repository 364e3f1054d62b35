"""ctypes harness over oracle/_ref/libexfunc_ref.so = the REFERENCE's own point_render / point_masker CUDA kernels,
compiled where they lie under /root/reference by oracle/Makefile (TEST INFRASTRUCTURE; GPU only).

The pybind/torch wrappers of the reference (point_render.cpp, point_masker.cpp) need the torch C++ API and are not
compiled; what they do around the launchers is restated here: output allocation and initial values
(point_render.cpp:80-88, 119-120; point_masker.cpp:55, 81)."""
from __future__ import annotations

import ctypes
import os

import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libexfunc_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def render_forward(proj, H, W, vpw, cpw, epsilon, threshold):
    B, N = proj.shape[:2]
    dev = proj.device
    pixel_index = -torch.ones((B, 2 * N), dtype=torch.int32, device=dev)
    is_visible = torch.ones((B, N), dtype=torch.int32, device=dev)
    front = torch.zeros((B, H, W, 1), device=dev)
    back = -2.0 * torch.ones((B, H, W, 1), device=dev)
    weight_points = torch.zeros((B, N, cpw * cpw), device=dev)
    depth = torch.zeros((B, H, W, 1), device=dev)
    weight = epsilon * torch.ones((B, H, W, 1), device=dev)
    torch.cuda.synchronize()
    _load().ref_point_render_forward(_p(proj), _p(pixel_index), _p(is_visible), _p(front), _p(back), _p(weight_points),
                                     _p(depth), _p(weight), B, N, H, W, vpw, cpw, ctypes.c_float(threshold))
    torch.cuda.synchronize()
    return dict(pixel_index=pixel_index, is_visible=is_visible, weight_points=weight_points, depth_image=depth,
                weight_image=weight, front=front, back=back)


def render_backward(grad_depth, grad_weight, proj, pixel_index, is_visible, weight_points):
    B, N = proj.shape[:2]
    H, W = grad_depth.shape[1:3]
    cpw = int(round(weight_points.shape[2] ** 0.5))
    grad_proj = torch.zeros_like(proj)
    grad_wp = torch.zeros_like(weight_points)
    torch.cuda.synchronize()
    _load().ref_point_render_backward(_p(grad_depth), _p(grad_weight), _p(proj), _p(pixel_index), _p(is_visible),
                                      _p(weight_points), _p(grad_wp), _p(grad_proj), B, N, H, W, cpw)
    torch.cuda.synchronize()
    return grad_proj


def masker_forward(proj, H, W, mpw):
    B, N = proj.shape[:2]
    mask = torch.zeros((B, H, W, 1), device=proj.device)
    torch.cuda.synchronize()
    _load().ref_point_masker_forward(_p(proj), _p(mask), B, N, H, W, mpw)
    torch.cuda.synchronize()
    return mask


def masker_backward(grad_mask, proj, epsilon, threshold):
    B, N = proj.shape[:2]
    H, W = grad_mask.shape[1:3]
    grad_proj = torch.zeros_like(proj)
    torch.cuda.synchronize()
    _load().ref_point_masker_backward(_p(grad_mask), _p(proj), _p(grad_proj), B, N, H, W, ctypes.c_float(epsilon),
                                      ctypes.c_float(threshold))
    torch.cuda.synchronize()
    return grad_proj
