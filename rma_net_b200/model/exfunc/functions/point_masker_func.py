"""Drop-in for the reference's model/exfunc/functions/point_masker_func.py (Masker_Point / Masker):

    Masker(image_height, image_width, mask_patch_width, epsilon)(proj_points, threshold) -> mask_image [B,H,W,1]

-1 inside every point's patch, 0 elsewhere; the backward spreads each pixel's gradient over the z of all points with
normalised Gaussian weights exp(-d^2/1000), skipping pixels with |grad| <= threshold (point_masker_cuda.cu:39-88).
Here it is deterministic and atomic-free (active pixels are compacted, points are gathered).
"""
import torch

from .... import ext


class Masker_Point(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj_points, image_height, image_width, mask_patch_width, epsilon, threshold):
        ctx.threshold = threshold
        ctx.epsilon = epsilon
        mask_image = ext.point_masker_forward(proj_points, image_height, image_width, mask_patch_width)[0]
        ctx.save_for_backward(proj_points)
        return mask_image

    @staticmethod
    def backward(ctx, grad_mask_image):
        proj_points = ctx.saved_tensors[0]
        grad_proj_points = ext.point_masker_backward(grad_mask_image, proj_points, ctx.epsilon, ctx.threshold)[0]
        return grad_proj_points, None, None, None, None, None


class Masker(torch.nn.Module):
    def __init__(self, image_height=256, image_width=256, mask_patch_width=5, epsilon=1e-5):
        super(Masker, self).__init__()
        self.image_height = image_height
        self.image_width = image_width
        self.mask_patch_width = mask_patch_width
        self.epsilon = epsilon  # avoid division by zero

    def forward(self, proj_points, threshold):
        return Masker_Point.apply(proj_points, self.image_height, self.image_width, self.mask_patch_width,
                                  self.epsilon, threshold)
