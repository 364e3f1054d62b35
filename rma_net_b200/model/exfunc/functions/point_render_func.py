"""Drop-in for the reference's model/exfunc/functions/point_render_func.py (Render_Point / Render), on the B200
kernels of rma_net_b200/csrc/exfunc.cu.  Same names, argument order and outputs:

    Render(image_height, image_width, visible_patch_width, color_path_width, epsilon)(proj_points, threshold)
        -> depth_image [B,H,W,1], weight_image [B,H,W,1], is_visible [B,N] (int32, non-differentiable)
    rendered depth = depth_image / weight_image                      (point_render_func.py:5-17)

Differences a caller can observe: asynchronous on the current stream (the reference wrapper synchronises the device
around every launch group, point_render.cpp:54-55, 78-79, 95-96); the [B,N,cpw^2] `weight_points` tensor is neither
materialised nor saved (the backward recomputes it); bad arguments raise RuntimeError instead of exiting.
"""
import torch

from .... import ext


class Render_Point(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj_points, image_height, image_width, visible_patch_width, color_path_width, epsilon, threshold):
        pixel_index, is_visible, _, depth_image, weight_image = ext.point_render_forward(
            proj_points, image_height, image_width, visible_patch_width, color_path_width, epsilon, threshold)
        ctx.save_for_backward(proj_points, pixel_index, is_visible)
        ctx.color_path_width = color_path_width
        ctx.mark_non_differentiable(is_visible)
        return depth_image, weight_image, is_visible

    @staticmethod
    def backward(ctx, grad_depth_image, grad_weight_image, _):
        proj_points, pixel_index, is_visible = ctx.saved_tensors
        if grad_depth_image is None:
            grad_depth_image = torch.zeros((proj_points.shape[0],) + tuple(grad_weight_image.shape[1:]),
                                           device=proj_points.device)
        if grad_weight_image is None:
            grad_weight_image = torch.zeros_like(grad_depth_image)
        grad_proj_points = ext.point_render_backward(grad_depth_image, grad_weight_image, proj_points, pixel_index,
                                                     is_visible, ctx.color_path_width)[0]
        return grad_proj_points, None, None, None, None, None, None


class Render(torch.nn.Module):
    def __init__(self, image_height=256, image_width=256, visible_patch_width=5, color_path_width=7, epsilon=1e-5):
        super(Render, self).__init__()
        self.image_height = image_height
        self.image_width = image_width
        self.visible_patch_width = visible_patch_width
        self.color_path_width = color_path_width
        self.epsilon = epsilon  # avoid division by zero

    def forward(self, proj_points, threshold):
        return Render_Point.apply(proj_points, self.image_height, self.image_width, self.visible_patch_width,
                                  self.color_path_width, self.epsilon, threshold)
