"""Fused recurrent-loop warp (reference model/network.py:652-655, 567-570, 840-843) as an exfunc-style op:
    rigid_matrix = Exp(phi);  rigid = transform(rigid_matrix.view(B,1,4,4), source)
    deformation  = w * rigid + (1 - w) * prev.detach()
All three outputs are differentiable, as in the reference: the deformation (chamfer / arap / ... terms), the rigid
points (deform_rigid_points_list) and the rigid matrix (rigid_matrix_list -> Loss.loss_on_tran, loss.py:224-230,
311-319, trained with --weight_tran 0.1 in train_sample.py).  Gradients flow to phi (through ExpMap.backward
semantics, se3.py:133-152), to w, and optionally to source; `prev` is detached in the reference (network.py:606, 798)
and gets no gradient here either."""
import torch

from .... import ext


class WarpFunction(torch.autograd.Function):
    """forward(phi, source, w, prev) -> (deformation, rigid, g); w = prev = None: stage 0 / rigid net, where the
    deformation IS the rigid points and `rigid` is not returned separately (the module hands out the same tensor)."""

    @staticmethod
    def forward(ctx, phi, source_points, point_weight, prev_points):
        out, rigid, g = ext.warp_forward(phi, source_points, point_weight, prev_points,
                                         want_rigid=point_weight is not None)
        ctx.blend = point_weight is not None
        ctx.w_shape = None if point_weight is None else point_weight.shape
        if ctx.blend:
            ctx.save_for_backward(phi, source_points, point_weight, prev_points)
            return out, rigid, g
        ctx.save_for_backward(phi, source_points)
        return out, g

    @staticmethod
    def backward(ctx, grad_out, *rest):
        if ctx.blend:
            phi, src, w, prev = ctx.saved_tensors
            grad_rigid, grad_g = rest
        else:
            (phi, src), w, prev = ctx.saved_tensors, None, None
            grad_rigid, (grad_g,) = None, rest
        need_w = ctx.blend and ctx.needs_input_grad[2]
        need_src = ctx.needs_input_grad[1]
        grad_phi, grad_w, grad_src = ext.warp_backward(phi, src, w, prev, grad_out, need_w, need_src,
                                                       grad_rigid=grad_rigid, grad_g=grad_g)
        if grad_w is not None:
            grad_w = grad_w.view(ctx.w_shape)
        return (grad_phi if ctx.needs_input_grad[0] else None), grad_src, grad_w, None


class Warp(torch.nn.Module):
    """forward(phi [B,6], source [B,N,3], w [B,1,N] | [B,N] | None, prev [B,N,3] | None)
    -> (deformation [B,N,3], rigid_points [B,N,3], rigid_matrix [B,4,4]), all differentiable."""

    def forward(self, phi, source_points, point_weight=None, prev_points=None):
        if prev_points is not None:
            prev_points = prev_points.detach()
        if point_weight is None:
            out, g = WarpFunction.apply(phi, source_points, None, None)
            return out, out, g          # stage 0 / rigid net: the rigid points ARE the deformation (network.py:570)
        return WarpFunction.apply(phi, source_points, point_weight, prev_points)
