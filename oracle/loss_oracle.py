"""CPU restatement (numpy, float64) of the reference's Loss.forward and its non-render terms, value AND gradient.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg may import this; the product
path under rma_net_b200/ never does).  Pinned against the reference's own code: tests/golden/loss_forward.npz holds
the outputs of /root/reference/model/loss.py's Loss.forward / loss_on_* exec'd verbatim on CPU
(oracle/gen_golden_loss.py); tests/test_oracle_golden.py checks this restatement against it.

  loss_on_pd      model/loss.py:207-212        loss_on_sparse  :217-220        loss_on_tran  :224-230
  loss_on_arap    model/loss.py:233-243        stage loops, gamma, arap factors, return value   :249-355
The chamfer term comes from oracle/chamfer_oracle.py (fp64 difference form).  Depth / mask terms are the exfunc
oracle's business (oracle/exfunc_oracle.py) and are not restated here.
"""
from __future__ import annotations

import numpy as np

from .chamfer_oracle import knn_exact, loss_on_cd_exact


def loss_on_pd(deformation_p, p1):
    """:207-212  sum((d - p)^2) / B   -> (value, d value / d deformation_p)"""
    d = np.asarray(deformation_p, np.float64) - np.asarray(p1, np.float64)
    B = d.shape[0]
    return (d * d).sum() / B, 2.0 * d / B


def loss_on_sparse(point_weight):
    """:217-220  mean |w|   -> (value, grad)"""
    w = np.asarray(point_weight, np.float64)
    return np.abs(w).mean(), np.sign(w) / w.size


def loss_on_tran(rigid_matrix):
    """:224-230  sum((mask * g)^2) / B with mask = 1 on [0:3, 3] (:134-135)   -> (value, grad)"""
    g = np.asarray(rigid_matrix, np.float64)
    B = g.shape[0]
    m = np.zeros((1, 4, 4))
    m[:, 0:3, 3] = 1.0
    t = m * g
    return (t * t).sum() / B, 2.0 * t * m / B


def loss_on_arap(neighbour_indexes, deformation_p, source_points):
    """:233-243  sum_{b,v,k} (sqrt(|d_j - d_v|^2 + 1e-5) - sqrt(|s_j - s_v|^2 + 1e-5))^2 / B, j = idx[b,v,k]
    -> (value, d value / d deformation_p)"""
    idx = np.asarray(neighbour_indexes, np.int64)
    d = np.asarray(deformation_p, np.float64)
    s = np.asarray(source_points, np.float64)
    B = d.shape[0]
    grad = np.zeros_like(d)
    total = 0.0
    for b in range(B):
        ed = d[b][idx[b]] - d[b][:, None, :]                  # indexing_neighbor (:113-120) - centre   [V,n,3]
        es = s[b][idx[b]] - s[b][:, None, :]
        ld = np.sqrt((ed * ed).sum(-1) + 0.00001)
        ls = np.sqrt((es * es).sum(-1) + 0.00001)
        diff = ld - ls
        total += (diff * diff).sum()
        gj = (2.0 * diff / ld)[..., None] * ed / B            # d / d(d_j); the centre gets the opposite
        np.add.at(grad[b], idx[b].reshape(-1), gj.reshape(-1, 3))
        grad[b] -= gj.sum(1)
    return total / B, grad


def arap_stage_factor(i):
    """:337-343"""
    return 10000.0 if i in (0, 1) else (1000.0 if i in (2, 3) else 100.0)


def get_neighbor_index(vertices, neighbor_num):
    """:99-110 with exact distances: the neighbor_num + 1 nearest by (distance, index), first dropped"""
    return knn_exact(vertices, vertices, neighbor_num + 1, skip_first=1)[0]


def loss_forward(args, point_weight_list, deformation_points_list, rigid_matrix_list, source_points, target_points):
    """Loss.forward (:249-355) without the depth / mask terms.  deformation_points_list[i] is [B,3,N].
    -> dict(loss, stages[7], grad_deform [T,B,3,N], grad_weights, grad_rigid)"""
    T = len(deformation_points_list)
    src = np.asarray(source_points, np.float32)
    tgt = np.asarray(target_points, np.float32)
    dl = [np.asarray(d, np.float32).transpose(0, 2, 1) for d in deformation_points_list]       # [B,N,3]
    gd = [np.zeros(d.shape, np.float64) for d in dl]
    gw = [np.zeros(np.asarray(w).shape, np.float64) for w in point_weight_list]
    gr = [np.zeros(np.asarray(r).shape, np.float64) for r in rigid_matrix_list]
    stages = [0.0] * 7
    loss = 0.0
    gam = [args.gamma ** (T - i - 1) for i in range(T)]
    if args.weight_cd > 0:
        for i in range(T):
            v, gx, _, _ = loss_on_cd_exact(dl[i], tgt)
            stages[0] += gam[i] * v
            gd[i] += args.weight_cd * gam[i] * gx
        loss += stages[0] * args.weight_cd
    if args.weight_pd > 0:
        for i in range(T):
            v, g = loss_on_pd(dl[i], tgt)
            stages[1] += gam[i] * v
            gd[i] += args.weight_pd * gam[i] * g
        loss += stages[1] * args.weight_pd
    if args.weight_sparse > 0:
        for i in range(T):
            v, g = loss_on_sparse(point_weight_list[i])
            stages[2] += gam[i] * v
            gw[i] += args.weight_sparse * gam[i] * g
        loss += stages[2] * args.weight_sparse
    if args.weight_tran > 0:
        for i in range(T):
            v, g = loss_on_tran(rigid_matrix_list[i])
            stages[4] += gam[i] * v
            gr[i] += args.weight_tran * gam[i] * g
        loss += stages[4] * args.weight_tran
    if args.weight_arap > 0:
        nb = get_neighbor_index(src, args.neighbour_num)
        for i in range(T):
            v, g = loss_on_arap(nb, dl[i], src)
            c = arap_stage_factor(i) * gam[i]
            stages[6] += c * v
            gd[i] += args.weight_arap * c * g
        loss += stages[6] * args.weight_arap
    return dict(loss=loss, stages=np.array(stages), grad_deform=np.stack([g.transpose(0, 2, 1) for g in gd]),
                grad_weights=np.stack(gw), grad_rigid=np.stack(gr))


def depth_view_term(ori_depth, ori_weight, tar_depth, tar_weight):
    """The image comparison of loss_on_depth (:169-181) in fp32 as torch evaluates it, and its gradient w.r.t. the
    deformation's depth / weight images (both masks are detached there).  -> (value f64, g_ori_depth, g_ori_weight)"""
    f = np.float32
    od, ow, td, tw = (np.asarray(a, f) for a in (ori_depth, ori_weight, tar_depth, tar_weight))
    ori = (od / ow).astype(f)                                                   # :170
    tar = (td / tw).astype(f)                                                   # :172
    mask = np.sign((ori * tar).astype(f))                                       # :173
    dis = np.abs((ori - tar).astype(f))                                         # :174
    thres = (f(-1.0) * (dis - f(0.05)).astype(f)).astype(f)                     # :175
    mask2 = (np.sign(np.sign(thres) + f(0.5)) + f(1.0)) * f(0.5)                # :176
    mask = (mask2 * mask).astype(f)                                             # :177
    value = (mask.astype(np.float64) * (dis * dis).astype(f).astype(np.float64)).sum()     # :178-179
    g_ori = 2.0 * mask.astype(np.float64) * (ori.astype(np.float64) - tar.astype(np.float64))
    return value, g_ori / ow.astype(np.float64), -g_ori * od.astype(np.float64) / ow.astype(np.float64) ** 2
