"""se(3) ops with the reference's names and signatures (model/LieAlgebra/se3.py), CUDA-only.

    exp(x)            se3.py:51-74     [*,6] -> [*,4,4]                      (no autograd, like the reference's
                                                                             plain `exp` is only ever used via Exp)
    ExpMap / Exp      se3.py:120-154   autograd.Function; backward = sum_ij go_ij (gen_k g)_ij, NOT d exp
    transform(g, a)   se3.py:102-112   both broadcast branches, differentiable w.r.t. g and a
    mat / genvec / genmat / group_prod  se3.py:24-49, 114-117  tiny helpers kept for importability (torch ops)
"""
import torch

from ... import ext


def mat(x):
    # [*,6] -> [*,4,4] twist matrix (se3.py:24-37)
    x_ = x.reshape(-1, 6)
    X = x_.new_zeros((x_.shape[0], 4, 4))
    X[:, 0, 1], X[:, 0, 2], X[:, 0, 3] = -x_[:, 2], x_[:, 1], x_[:, 3]
    X[:, 1, 0], X[:, 1, 2], X[:, 1, 3] = x_[:, 2], -x_[:, 0], x_[:, 4]
    X[:, 2, 0], X[:, 2, 1], X[:, 2, 3] = -x_[:, 1], x_[:, 0], x_[:, 5]
    return X.view(*x.shape[:-1], 4, 4)


def genvec():
    return torch.eye(6)


def genmat():
    return mat(genvec())


def group_prod(g, h):
    return g.matmul(h)


def exp(x):
    return ext.se3_exp_forward(x.detach())


class ExpMap(torch.autograd.Function):
    """Exp: se(3) -> SE(3), [B,6] -> [B,4,4] or [B,1,6] -> [B,1,4,4] (any leading dims)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ext.se3_exp_forward(x)

    @staticmethod
    def backward(ctx, grad_output):
        x, = ctx.saved_tensors
        return ext.se3_exp_backward(x, grad_output)


Exp = ExpMap.apply


class _Transform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, a):
        out, _ = ext.se3_transform_forward(g, a)
        # saved through autograd (version-counter check, saved-tensor hooks, freed with the graph); the launch plan
        # is a few integers and is rebuilt in backward
        ctx.save_for_backward(g, a)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        g, a = ctx.saved_tensors
        need_g, need_a = ctx.needs_input_grad
        plan = ext.TransformPlan(g, a)
        grad_g, grad_a = ext.se3_transform_backward(plan, grad_out, g.shape, a.shape, need_g, need_a)
        return grad_g, grad_a


def transform(g, a):
    # g : SE(3),  * x 4 x 4
    # a : R^3,    * x 3[x N]
    return _Transform.apply(g, a)
