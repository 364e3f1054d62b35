"""Fused recurrent-loop warp (reference model/network.py:652-655, 567-570, 840-843) as an exfunc-style op:
    deformation = w * transform(Exp(phi).view(B,1,4,4), source) + (1 - w) * prev.detach()
Gradients flow to phi (through ExpMap.backward semantics, se3.py:133-152), to w, and optionally to source;
`prev` is detached in the reference (network.py:606, 798) and gets no gradient here either."""
import torch

from .... import ext


class WarpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, phi, source_points, point_weight, prev_points):
        out, rigid, g = ext.warp_forward(phi, source_points, point_weight, prev_points)
        ctx.blend = point_weight is not None
        ctx.w_shape = None if point_weight is None else point_weight.shape
        if ctx.blend:
            ctx.save_for_backward(phi, source_points, point_weight, prev_points)
        else:
            ctx.save_for_backward(phi, source_points)
        ctx.mark_non_differentiable(rigid, g)
        return out, rigid, g

    @staticmethod
    def backward(ctx, grad_out, _grad_rigid, _grad_g):
        if ctx.blend:
            phi, src, w, prev = ctx.saved_tensors
        else:
            (phi, src), w, prev = ctx.saved_tensors, None, None
        need_w = ctx.blend and ctx.needs_input_grad[2]
        need_src = ctx.needs_input_grad[1]
        grad_phi, grad_w, grad_src = ext.warp_backward(phi, src, w, prev, grad_out, need_w, need_src)
        if grad_w is not None:
            grad_w = grad_w.view(ctx.w_shape)
        return (grad_phi if ctx.needs_input_grad[0] else None), grad_src, grad_w, None


class Warp(torch.nn.Module):
    """forward(phi [B,6], source [B,N,3], w [B,1,N] | [B,N] | None, prev [B,N,3] | None)
    -> (deformation [B,N,3], rigid_points [B,N,3], rigid_matrix [B,4,4])."""

    def forward(self, phi, source_points, point_weight=None, prev_points=None):
        if prev_points is not None:
            prev_points = prev_points.detach()
        return WarpFunction.apply(phi, source_points, point_weight, prev_points)
