"""Soft point rasteriser on the B200 kernels (rma_point_render_forward / _backward, csrc/exfunc.cu).

Keeps the public names and call signatures of the reference module of the same name, so `loss.py:34, 169-172` work
unchanged:

    Render(image_height, image_width, visible_patch_width, color_path_width, epsilon)(proj_points, threshold)
        -> depth_image [B,H,W,1], weight_image [B,H,W,1], is_visible [B,N] (int32, no gradient)
    Render_Point.apply(proj_points, H, W, visible_patch_width, color_path_width, epsilon, threshold)
    rendered depth = depth_image / weight_image

What a caller can observe beyond speed: calls are asynchronous on the current stream (the reference wrapper
synchronises the device around every launch group, point_render.cpp:54-55, 78-79, 95-96); the per-point Gaussian
weights [B,N,cpw^2] are neither materialised nor saved (the backward recomputes them, same bits); invalid arguments
raise RuntimeError instead of ending the process.
"""
from typing import NamedTuple

import torch

from .... import ext


class _RenderSetup(NamedTuple):
    height: int
    width: int
    visible_patch: int
    color_patch: int
    epsilon: float


class Render_Point(torch.autograd.Function):
    """Positional arguments as in the reference: (points, H, W, visibility patch, colour patch, epsilon, threshold)."""

    @staticmethod
    def forward(ctx, proj_points, *setup_and_threshold):
        setup, depth_gap = _RenderSetup(*setup_and_threshold[:5]), setup_and_threshold[5]
        pixel_of_point, visible, _unused_weights, depth_sum, weight_sum = ext.point_render_forward(
            proj_points, setup.height, setup.width, setup.visible_patch, setup.color_patch, setup.epsilon, depth_gap)
        ctx.color_patch = setup.color_patch
        ctx.save_for_backward(proj_points, pixel_of_point, visible)
        ctx.mark_non_differentiable(visible)
        return depth_sum, weight_sum, visible

    @staticmethod
    def backward(ctx, g_depth, g_weight, _g_visible):
        points, pixel_of_point, visible = ctx.saved_tensors
        if g_depth is None and g_weight is None:
            return (None,) * 7
        # an output that took no part in the loss arrives as None: its gradient image is all zeros
        like = g_depth if g_depth is not None else g_weight
        g_depth = torch.zeros_like(like) if g_depth is None else g_depth
        g_weight = torch.zeros_like(like) if g_weight is None else g_weight
        (g_points,) = ext.point_render_backward(g_depth, g_weight, points, pixel_of_point, visible, ctx.color_patch)
        return (g_points,) + (None,) * 6


class Render(torch.nn.Module):
    def __init__(self, image_height=256, image_width=256, visible_patch_width=5, color_path_width=7, epsilon=1e-5):
        super().__init__()
        self.setup = _RenderSetup(int(image_height), int(image_width), int(visible_patch_width),
                                  int(color_path_width), float(epsilon))

    # the attributes the reference module exposes
    image_height = property(lambda self: self.setup.height)
    image_width = property(lambda self: self.setup.width)
    visible_patch_width = property(lambda self: self.setup.visible_patch)
    color_path_width = property(lambda self: self.setup.color_patch)
    epsilon = property(lambda self: self.setup.epsilon)

    def forward(self, proj_points, threshold):
        return Render_Point.apply(proj_points, *self.setup, threshold)
