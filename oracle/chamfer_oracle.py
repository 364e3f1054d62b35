"""CPU restatement of the reference's chamfer semantics (TEST INFRASTRUCTURE — see oracle/__init__.py).

Two oracles, as SURVEY.md §8c recommends:

* **O1** (`ref_compute_sqrdis_map`, `ref_chamfer_dist`, `ref_loss_on_cd`): the reference's own formulation,
  restated: expanded form ``|x|^2 + |y|^2 - 2 x.y`` with one batched matmul, full ``[B,M,N]`` matrix, two
  ``min`` reductions, implicit torch autograd.  Follows /root/reference/model/loss.py:45-58 (sqrdis map),
  :60-70 (chamfer_dist), :200-205 (loss_on_cd).  fp32, reference rounding (cancellation error up to ~1e-2 rel).
* **O2** (`chamfer_exact`, `chamfer_backward_exact`): the same *semantics* (squared distance, no sqrt, first-minimum
  tie break, gradient to the single winning index — the behaviour of torch ``min(dim)`` + autograd the reference
  relies on implicitly) evaluated in fp64 difference form ``sum((x-y)^2)`` on the fp32-rounded inputs.  This is the
  tight target the CUDA path is compared against (distances/gradients 1e-5 relative, indices equal except
  near-ties at fp32 resolution).

Pinning: the reference ships no tests or golden outputs for this path.  O1/O2 are pinned instead
against outputs of the reference's own code run in the build container (`oracle/gen_golden.py` exec's
loss.py:45-70 verbatim from /root/reference and stores inputs + outputs under tests/golden/), and against the
known-answer values of SURVEY.md §8c.  See tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch


# --------------------------------------------------------------------------------------------------------------
# O1 — reference formulation (loss.py:45-70, 200-205), torch fp32 on CPU
# --------------------------------------------------------------------------------------------------------------
def ref_compute_sqrdis_map(points_x: torch.Tensor, points_y: torch.Tensor) -> torch.Tensor:
    """[B,M,3] x [B,N,3] -> [B,M,N] squared distances, expanded form (loss.py:45-58)."""
    b, m, n = points_x.shape[0], points_x.shape[1], points_y.shape[1]
    xx = (points_x * points_x).sum(-1).view(b, m, 1).expand(-1, -1, n)
    yy = (points_y * points_y).sum(-1).view(b, 1, n).expand(-1, m, -1)
    xy = torch.bmm(points_x, points_y.transpose(1, 2))
    return xx + yy - 2 * xy


def ref_chamfer_dist(points_x: torch.Tensor, points_y: torch.Tensor):
    """(dist1 [B,M], dist2 [B,N]) = row / column minima of the sqrdis map (loss.py:60-70)."""
    b = points_x.shape[0]
    sq = ref_compute_sqrdis_map(points_x, points_y)
    return sq.min(dim=2)[0].view(b, -1), sq.min(dim=1)[0].view(b, -1)


def ref_chamfer_dist_with_idx(points_x: torch.Tensor, points_y: torch.Tensor):
    """As ref_chamfer_dist, plus the argmin indices torch computes and the reference discards."""
    sq = ref_compute_sqrdis_map(points_x, points_y)
    d1, i1 = sq.min(dim=2)
    d2, i2 = sq.min(dim=1)
    return d1, d2, i1, i2


def ref_loss_on_cd(deformation_p: torch.Tensor, p1: torch.Tensor) -> torch.Tensor:
    """0.5 * (sum dist1 + sum dist2) / B (loss.py:200-205)."""
    d1, d2 = ref_chamfer_dist(deformation_p, p1)
    return (d1.sum() + d2.sum()) * 0.5 / deformation_p.shape[0]


def ref_fwd_bwd(points_x: torch.Tensor, points_y: torch.Tensor):
    """One forward+backward of loss_on_cd through torch autograd, the way train_view.py:375,385 exercises it.
    Returns (loss, grad_x, grad_y)."""
    x = points_x.detach().clone().requires_grad_(True)
    y = points_y.detach().clone().requires_grad_(True)
    loss = ref_loss_on_cd(x, y)
    loss.backward()
    return loss.detach(), x.grad, y.grad


# --------------------------------------------------------------------------------------------------------------
# O2 — same semantics, fp64 difference form, explicit argmin + analytic backward
# --------------------------------------------------------------------------------------------------------------
def _sqdist64(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """x [m,3], y [n,3] (any float dtype) -> [m,n] float64 sum((x-y)^2), evaluated in difference form."""
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    d = x[:, None, 0] - y[None, :, 0]
    out = d * d
    d = x[:, None, 1] - y[None, :, 1]
    out += d * d
    d = x[:, None, 2] - y[None, :, 2]
    out += d * d
    return out


def chamfer_exact(points_x, points_y, chunk: int = 1024):
    """fp64 difference-form chamfer with first-index tie break.

    points_x [B,M,3], points_y [B,N,3] (numpy or torch, fp32 values are used as-is, exactly).
    Returns dict(dist1 [B,M] f64, dist2 [B,N] f64, idx1 [B,M] i64, idx2 [B,N] i64).
    np.argmin returns the first occurrence of the minimum == torch.min(dim) behaviour (SURVEY.md §7 tie-break).
    """
    x = np.asarray(points_x.detach().cpu().numpy() if isinstance(points_x, torch.Tensor) else points_x)
    y = np.asarray(points_y.detach().cpu().numpy() if isinstance(points_y, torch.Tensor) else points_y)
    B, M, N = x.shape[0], x.shape[1], y.shape[1]
    dist1 = np.empty((B, M), np.float64)
    idx1 = np.empty((B, M), np.int64)
    dist2 = np.full((B, N), np.inf, np.float64)
    idx2 = np.zeros((B, N), np.int64)
    for b in range(B):
        for s in range(0, M, chunk):
            d = _sqdist64(x[b, s:s + chunk], y[b])          # [c, N]
            j = d.argmin(axis=1)
            idx1[b, s:s + chunk] = j
            dist1[b, s:s + chunk] = d[np.arange(d.shape[0]), j]
            i = d.argmin(axis=0)                              # first row inside this chunk
            dm = d[i, np.arange(N)]
            better = dm < dist2[b]                            # strict: earlier chunk (lower index) wins ties
            dist2[b][better] = dm[better]
            idx2[b][better] = i[better] + s
    return dict(dist1=dist1, dist2=dist2, idx1=idx1, idx2=idx2)


def chamfer_backward_exact(points_x, points_y, idx1, idx2, g1, g2):
    """Analytic backward of (dist1, dist2) w.r.t. both clouds in fp64.

    grad_x[b,i] = 2 g1[b,i] (x_i - y_{idx1[b,i]}) + sum_{j: idx2[b,j]=i} 2 g2[b,j] (x_i - y_j)
    grad_y[b,j] = 2 g2[b,j] (y_j - x_{idx2[b,j]}) + sum_{i: idx1[b,i]=j} 2 g1[b,i] (y_j - x_i)
    """
    x = np.asarray(points_x, np.float64)
    y = np.asarray(points_y, np.float64)
    g1 = np.asarray(g1, np.float64)
    g2 = np.asarray(g2, np.float64)
    gx = np.zeros_like(x)
    gy = np.zeros_like(y)
    for b in range(x.shape[0]):
        d1 = 2.0 * g1[b][:, None] * (x[b] - y[b][idx1[b]])     # [M,3]
        gx[b] += d1
        np.add.at(gy[b], idx1[b], -d1)
        d2 = 2.0 * g2[b][:, None] * (y[b] - x[b][idx2[b]])     # [N,3]
        gy[b] += d2
        np.add.at(gx[b], idx2[b], -d2)
    return gx, gy


def loss_on_cd_exact(points_x, points_y):
    """fp64 loss_on_cd plus its analytic gradients (upstream g1 = g2 = 0.5 / B)."""
    r = chamfer_exact(points_x, points_y)
    B = r["dist1"].shape[0]
    loss = 0.5 * (r["dist1"].sum() + r["dist2"].sum()) / B
    g1 = np.full_like(r["dist1"], 0.5 / B)
    g2 = np.full_like(r["dist2"], 0.5 / B)
    x = np.asarray(points_x.detach().cpu().numpy() if isinstance(points_x, torch.Tensor) else points_x)
    y = np.asarray(points_y.detach().cpu().numpy() if isinstance(points_y, torch.Tensor) else points_y)
    gx, gy = chamfer_backward_exact(x, y, r["idx1"], r["idx2"], g1, g2)
    return loss, gx, gy, r


def classify_index_mismatches(points_q, points_t, idx_test, idx_oracle, rel_tol: float = 2.0 ** -20):
    """Compare argmin indices of one direction.  A mismatch is a *documented near-tie* when the fp64 distances of
    the two candidates differ by <= rel_tol * distance (fp32 difference-form resolution is ~3 * 2^-24 relative).
    Returns (n_mismatch, n_near_tie, n_bad)."""
    q = np.asarray(points_q, np.float64)
    t = np.asarray(points_t, np.float64)
    idx_test = np.asarray(idx_test, np.int64)
    idx_oracle = np.asarray(idx_oracle, np.int64)
    mism = np.argwhere(idx_test != idx_oracle)
    near = 0
    for b, i in mism:
        da = ((q[b, i] - t[b, idx_test[b, i]]) ** 2).sum()
        do = ((q[b, i] - t[b, idx_oracle[b, i]]) ** 2).sum()
        if abs(da - do) <= rel_tol * max(do, 1e-300):
            near += 1
    return len(mism), near, len(mism) - near


def read_obj_vertices(path: str) -> np.ndarray:
    """OBJ parse rule of the fixtures: lines ``v x y z [r g b]``, tokens 1-3, cast to fp32."""
    pts = []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if len(tok) >= 4 and tok[0] == "v":
                pts.append((float(tok[1]), float(tok[2]), float(tok[3])))
    return np.asarray(pts, np.float32)


def knn_exact(queries, candidates, k, skip_first=0):
    """Brute-force kNN in float64 difference form, order (distance, index) — the semantics of rma_knn.
    Returns (idx [B,N,k-skip_first] int32, dist float64).  For sizes the oracle finishes in seconds."""
    q = np.asarray(queries, np.float64)
    c = np.asarray(candidates, np.float64)
    B, N = q.shape[:2]
    M = c.shape[1]
    idx = np.empty((B, N, k - skip_first), np.int32)
    dist = np.empty((B, N, k - skip_first), np.float64)
    for b in range(B):
        for s in range(0, N, 512):
            d = ((q[b, s:s + 512, None, :] - c[b, None, :, :]) ** 2).sum(-1)             # [n, M]
            order = np.lexsort((np.broadcast_to(np.arange(M), d.shape), d), axis=-1)[:, :k]
            idx[b, s:s + 512] = order[:, skip_first:]
            dist[b, s:s + 512] = np.take_along_axis(d, order, -1)[:, skip_first:]
    return idx, dist


def knn_reference_formulation(x_b3n, k):
    """network.py:52-58 restated with torch on CPU (expanded form + topk): idx [B,N,k] int64."""
    import torch
    x = torch.as_tensor(x_b3n)
    inner = -2 * torch.matmul(x.transpose(2, 1).contiguous(), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pairwise_distance = -xx - inner - xx.transpose(2, 1).contiguous()
    return pairwise_distance.topk(k=k, dim=-1)[1].numpy()
