"""Chamfer nearest-neighbour op packaged the way the reference packages its native ops
(model/exfunc/functions/point_render_func.py:19-45): an autograd.Function whose forward calls `ext.forward`,
saves what backward needs, returns extra non-differentiable outputs whose grads are ignored, plus a thin nn.Module.
"""
import torch

from .... import ext


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points_x, points_y):
        dist1, dist2, idx1, idx2 = ext.chamfer_forward(points_x, points_y, want_idx=True)
        ctx.save_for_backward(points_x, points_y, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2, _i1, _i2):
        points_x, points_y, idx1, idx2 = ctx.saved_tensors
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_y):
            return None, None
        grad_x, grad_y = ext.chamfer_backward(points_x, points_y, idx1, idx2, grad_dist1, grad_dist2,
                                              need_x=need_x, need_y=need_y)
        return grad_x, grad_y


class ChamferCDFunction(torch.autograd.Function):
    """loss_on_cd (reference model/loss.py:200-205) as ONE op: 0.5 * (sum dist1 + sum dist2) / B and its gradients
    w.r.t. both clouds come out of the forward's three launches; backward is one combine-and-scale launch."""

    @staticmethod
    def forward(ctx, deformation_p, p1, batch_weight=None, coef=None, y_period=0):
        """y_period > 0: p1 holds y_period clouds shared by the batch elements b, b + y_period, ... of deformation_p
        (the stage loop's target, not repeated); its gradient is the sum over the batch elements that used it."""
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        loss, gx, gy = ext.chamfer_cd_forward(deformation_p, p1, need_x, need_y, batch_weight, coef, y_period=y_period)
        ctx.has = (gx is not None, gy is not None)
        ctx.y_period = int(y_period or 0)
        ctx.save_for_backward(*[t for pair in (gx, gy) if pair is not None for t in pair])
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        saved = list(ctx.saved_tensors)
        gx = (saved.pop(0), saved.pop(0)) if ctx.has[0] else None
        gy = (saved.pop(0), saved.pop(0)) if ctx.has[1] else None
        if gx is None and gy is None:
            return None, None, None, None, None
        ox, oy = ext.chamfer_cd_backward(gx, gy, grad_loss)
        if oy is not None and ctx.y_period:
            oy = oy.view(-1, ctx.y_period, oy.shape[1], 3).sum(0)
        return ox, oy, None, None, None


class ChamferCDMetricsFunction(torch.autograd.Function):
    """ChamferCDFunction that also hands back the unweighted dist1 [B,N] / dist2 [B,M] of the same forward
    (non-differentiable), so that the metric callers (reference train_view.py:376-380) do not repeat the search."""

    @staticmethod
    def forward(ctx, deformation_p, p1, batch_weight=None, coef=None, y_period=0):
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        loss, gx, gy, d1, d2 = ext.chamfer_cd_forward(deformation_p, p1, need_x, need_y, batch_weight, coef,
                                                      want_dist=True, y_period=y_period)
        ctx.has = (gx is not None, gy is not None)
        ctx.y_period = int(y_period or 0)
        ctx.save_for_backward(*[t for pair in (gx, gy) if pair is not None for t in pair])
        ctx.mark_non_differentiable(d1, d2)
        return loss, d1, d2

    @staticmethod
    def backward(ctx, grad_loss, _d1, _d2):
        return ChamferCDFunction.backward(ctx, grad_loss)


class ArapFunction(torch.autograd.Function):
    """Edge-length (ARAP) term of Loss.forward for any number of stacked stages (reference model/loss.py:233-243,
    334-345): forward(deformation [G,V,3], source [B,V,3], neighbour_idx [B,V,n] int32, group_weight [G] | None,
    coef) -> scalar; the gradient w.r.t. the deformation comes out of the same launch."""

    @staticmethod
    def forward(ctx, deformation, source_points, neighbour_idx, group_weight=None, coef=1.0):
        if ctx.needs_input_grad[1]:
            raise RuntimeError("ArapFunction: no gradient w.r.t. source_points (it is input data in the reference)")
        loss, grad = ext.arap_forward(deformation, source_points, neighbour_idx, group_weight, float(coef),
                                      ctx.needs_input_grad[0])
        ctx.save_for_backward(*([grad] if grad is not None else []))
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        if not ctx.saved_tensors:
            return None, None, None, None, None
        return ctx.saved_tensors[0] * grad_loss, None, None, None, None


class DepthViewTermFunction(torch.autograd.Function):
    """sum(mask * dis^2) of loss_on_depth (reference model/loss.py:169-181) from the four rendered images, with the
    gradients w.r.t. the deformation's depth / weight images out of the same pass; the target images get none."""

    @staticmethod
    def forward(ctx, ori_depth, ori_weight, tar_depth, tar_weight, coef=1.0, dis_threshold=0.05):
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            raise RuntimeError("DepthViewTermFunction: no gradient w.r.t. the target images")
        loss, grads = ext.depth_view_term(ori_depth, ori_weight, tar_depth, tar_weight, dis_threshold, coef,
                                          ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        ctx.save_for_backward(*(grads or ()))
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        if not ctx.saved_tensors:
            return None, None, None, None, None, None
        g_od, g_ow = ctx.saved_tensors
        return g_od * grad_loss, g_ow * grad_loss, None, None, None, None


class ViewProjectFunction(torch.autograd.Function):
    """points [B,N,3], stacked view rotations [3V,3] -> [V*B,N,3] = (R_v p + add3) * mul3: the rotate-and-map-to-
    pixels step of loss_on_depth / loss_on_mask (reference model/loss.py:161-166, 187-191) with the views folded into
    the batch; one launch forward, one backward (gradient w.r.t. the points only)."""

    @staticmethod
    def forward(ctx, points, rotations, add3, mul3):
        ctx.save_for_backward(rotations)
        ctx.shape = (points.shape[0], points.shape[1])
        ctx.mul3 = tuple(float(m) for m in mul3)
        return ext.view_project_forward(points, rotations, add3, mul3)

    @staticmethod
    def backward(ctx, grad_out):
        (rotations,) = ctx.saved_tensors
        return ext.view_project_backward(grad_out, rotations, ctx.shape[0], ctx.shape[1], ctx.mul3), None, None, None


class Chamfer(torch.nn.Module):
    """forward(points_x [B,M,3], points_y [B,N,3]) -> (dist1 [B,M], dist2 [B,N], idx1, idx2)."""

    def forward(self, points_x, points_y):
        return ChamferFunction.apply(points_x, points_y)
