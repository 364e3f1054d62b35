"""CPU restatement of the reference's se(3) exp-map / transform / warp semantics (TEST INFRASTRUCTURE).

Follows /root/reference/model/LieAlgebra/se3.py:24-49 (mat/genmat), :51-74 (exp), :102-112 (transform),
:133-152 (ExpMap.backward); so3.py:14-24 (mat); sinc.py:5-17, 91-103, 120-132 (sinc1/2/3 with the |t| < 0.01
Taylor branch); model/network.py:567-570, 606, 652-655 (non-rigid warp loop) and :840-844 (rigid loop).
numpy, dtype-generic: run it in float64 for the tight target, in float32 to see the reference's own rounding.

Pinned against outputs of the reference's own LieAlgebra package imported in the build container
(oracle/gen_golden.py -> tests/golden/se3_*.npz) and the known-answer values of SURVEY.md §8c.
"""
from __future__ import annotations

import numpy as np


def sinc1(t):
    t = np.asarray(t)
    small = np.abs(t) < 0.01
    t2 = t * t
    ts = np.where(small, 1.0, t)            # avoid 0/0 in the unused branch
    taylor = 1 - t2 / 6 * (1 - t2 / 20 * (1 - t2 / 42))
    return np.where(small, taylor, np.sin(ts) / ts)


def sinc2(t):
    t = np.asarray(t)
    small = np.abs(t) < 0.01
    t2 = t * t
    t2s = np.where(small, 1.0, t2)
    taylor = 0.5 * (1 - t2 / 12 * (1 - t2 / 30 * (1 - t2 / 56)))
    return np.where(small, taylor, (1 - np.cos(t)) / t2s)


def sinc3(t):
    t = np.asarray(t)
    small = np.abs(t) < 0.01
    t2 = t * t
    ts = np.where(small, 1.0, t)
    taylor = (1.0 / 6) * (1 - t2 / 20 * (1 - t2 / 42 * (1 - t2 / 72)))
    return np.where(small, taylor, (t - np.sin(t)) / (ts ** 3))


def so3_mat(w):
    """[*,3] -> [*,3,3] cross-product matrix (so3.py:14-24)."""
    w = np.asarray(w)
    W = np.zeros(w.shape[:-1] + (3, 3), w.dtype)
    W[..., 0, 1], W[..., 0, 2] = -w[..., 2], w[..., 1]
    W[..., 1, 0], W[..., 1, 2] = w[..., 2], -w[..., 0]
    W[..., 2, 0], W[..., 2, 1] = -w[..., 1], w[..., 0]
    return W


def se3_mat(x):
    """[*,6] -> [*,4,4] twist matrix (se3.py:24-37)."""
    x = np.asarray(x)
    X = np.zeros(x.shape[:-1] + (4, 4), x.dtype)
    X[..., :3, :3] = so3_mat(x[..., :3])
    X[..., :3, 3] = x[..., 3:]
    return X


def se3_genmat(dtype=np.float64):
    return se3_mat(np.eye(6, dtype=dtype))


def se3_exp(x):
    """[*,6] -> [*,4,4]  (se3.py:51-74)."""
    x = np.asarray(x)
    w, v = x[..., :3], x[..., 3:]
    t = np.sqrt((w * w).sum(-1))[..., None, None]
    W = so3_mat(w)
    S = W @ W
    I = np.eye(3, dtype=x.dtype)
    R = I + sinc1(t) * W + sinc2(t) * S
    V = I + sinc2(t) * W + sinc3(t) * S
    p = (V @ v[..., None])[..., 0]
    g = np.zeros(x.shape[:-1] + (4, 4), x.dtype)
    g[..., :3, :3] = R
    g[..., :3, 3] = p
    g[..., 3, 3] = 1
    return g


def se3_expmap_backward(x, grad_output):
    """ExpMap.backward (se3.py:133-152): grad_x[k] = sum_ij go_ij (gen_k @ exp(x))_ij — generator form."""
    x = np.asarray(x)
    g = se3_exp(x)
    gen = se3_genmat(x.dtype)                                   # [6,4,4]
    dg = gen @ g[..., None, :, :]                               # [*,6,4,4]
    go = np.asarray(grad_output, x.dtype)[..., None, :, :]
    return (go * dg).sum(-1).sum(-1)


def se3_expmap_backward_closed(x, grad_output):
    """Closed form of the above: grad_w = sum_c g[:3,c] x go[:3,c], grad_v = go[:3,3]."""
    x = np.asarray(x)
    g = se3_exp(x)
    go = np.asarray(grad_output, x.dtype)
    gw = np.cross(np.swapaxes(g[..., :3, :], -1, -2), np.swapaxes(go[..., :3, :], -1, -2)).sum(-2)
    return np.concatenate([gw, go[..., :3, 3]], -1)


def se3_transform(g, a):
    """se3.py:102-112, both branches."""
    g = np.asarray(g)
    a = np.asarray(a)
    R = g[..., :3, :3]
    p = g[..., :3, 3]
    if g.ndim == a.ndim:
        return R @ a + p[..., None]
    return (R @ a[..., None])[..., 0] + p


def se3_transform_backward(g, a, grad_out):
    """Analytic backward of the `else` branch with g [B,1,4,4], a [B,N,3] (the hot call): returns
    (grad_g [B,1,4,4], grad_a [B,N,3])."""
    g = np.asarray(g)
    a = np.asarray(a)
    go = np.asarray(grad_out, a.dtype)
    R = g[..., :3, :3]
    grad_a = (np.swapaxes(R, -1, -2) @ go[..., None])[..., 0]
    grad_g = np.zeros_like(g)
    grad_g[..., 0, :3, :3] = np.einsum("bni,bnj->bij", go, a)
    grad_g[..., 0, :3, 3] = go.sum(1)
    return grad_g, grad_a


def warp_nonrigid_loop(source, phis, weights):
    """Non-rigid recurrent loop (network.py:567-570, 606, 652-655).
    source [B,N,3]; phis list of [B,6]; weights list of [B,N] (entry 0 ignored: stage 0 has no blend).
    Every stage warps the ORIGINAL source; stage k>=1 blends with the previous (detached) deformation.
    Returns list of deformations [B,N,3]."""
    source = np.asarray(source)
    out = []
    prev = None
    for k, phi in enumerate(phis):
        g = se3_exp(np.asarray(phi, source.dtype))[:, None]          # [B,1,4,4]
        rigid = se3_transform(g, source)
        if k == 0:
            deformation = rigid
        else:
            w = np.asarray(weights[k], source.dtype)[..., None]
            deformation = w * rigid + (1 - w) * prev
        out.append(deformation)
        prev = deformation
    return out


def warp_rigid_loop(source, phis):
    """Rigid recurrent loop (network.py:840-844): warps the previous deformation, accumulates g_i @ acc."""
    cur = np.asarray(source)
    acc = np.broadcast_to(np.eye(4, dtype=cur.dtype), (cur.shape[0], 4, 4)).copy()
    outs, accs = [], []
    for phi in phis:
        g = se3_exp(np.asarray(phi, cur.dtype))
        cur = se3_transform(g[:, None], cur)
        acc = g @ acc
        outs.append(cur)
        accs.append(acc)
    return outs, accs


def warp_stage_backward(phi, source, w, prev, grad_out):
    """Backward of one blended stage w.r.t. phi (through ExpMap.backward semantics) and w.
    w may be None (stage 0).  Returns (grad_phi [B,6], grad_w [B,N] or None)."""
    source = np.asarray(source)
    dt = source.dtype
    g = se3_exp(np.asarray(phi, dt))[:, None]
    rigid = se3_transform(g, source)
    go = np.asarray(grad_out, dt)
    grad_w = None
    if w is not None:
        grad_w = (go * (rigid - np.asarray(prev, dt))).sum(-1)
        go = go * np.asarray(w, dt)[..., None]
    grad_g, _ = se3_transform_backward(g, source, go)
    grad_phi = se3_expmap_backward_closed(np.asarray(phi, dt), grad_g[:, 0])
    return grad_phi, grad_w
