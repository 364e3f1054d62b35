"""Silhouette mask of a projected point cloud on the B200 kernels (rma_point_masker_forward / _backward).

Keeps the public names and call signatures of the reference module of the same name, so `loss.py:35, 193-194` work
unchanged:

    Masker(image_height, image_width, mask_patch_width, epsilon)(proj_points, threshold) -> mask [B,H,W,1]
    Masker_Point.apply(proj_points, image_height, image_width, mask_patch_width, epsilon, threshold)

The mask is -1 inside every point's patch and 0 elsewhere.  Its backward spreads each pixel's gradient over the z of
all points with normalised Gaussian weights exp(-d^2/1000), skipping pixels with |grad| <= threshold
(point_masker_cuda.cu:39-88); here that is deterministic and atomic-free (active pixels are compacted, then gathered
per point).
"""
from typing import NamedTuple

import torch

from .... import ext


class _MaskSetup(NamedTuple):
    height: int
    width: int
    patch: int
    epsilon: float


class Masker_Point(torch.autograd.Function):
    """Positional arguments as in the reference: (points, H, W, patch width, epsilon, threshold)."""

    @staticmethod
    def forward(ctx, proj_points, *setup_and_threshold):
        setup, cutoff = _MaskSetup(*setup_and_threshold[:4]), setup_and_threshold[4]
        (mask,) = ext.point_masker_forward(proj_points, setup.height, setup.width, setup.patch)
        ctx.save_for_backward(proj_points)
        ctx.backward_scalars = (setup.epsilon, cutoff)
        return mask

    @staticmethod
    def backward(ctx, upstream):
        (points,) = ctx.saved_tensors
        (grad_points,) = ext.point_masker_backward(upstream, points, *ctx.backward_scalars)
        return (grad_points,) + (None,) * 5


class Masker(torch.nn.Module):
    def __init__(self, image_height=256, image_width=256, mask_patch_width=5, epsilon=1e-5):
        super().__init__()
        self.setup = _MaskSetup(int(image_height), int(image_width), int(mask_patch_width), float(epsilon))

    # the attributes the reference module exposes
    image_height = property(lambda self: self.setup.height)
    image_width = property(lambda self: self.setup.width)
    mask_patch_width = property(lambda self: self.setup.patch)
    epsilon = property(lambda self: self.setup.epsilon)

    def forward(self, proj_points, threshold):
        return Masker_Point.apply(proj_points, *self.setup, threshold)
