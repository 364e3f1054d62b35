"""Chamfer part of the reference's model/loss.py, same names and argument meaning, CUDA-only.

    compute_sqrdis_map   loss.py:45-58     (kept callable; materialises [B,M,N]; NOT used by the hot path)
    chamfer_dist         loss.py:60-70     -> (dist1 [B,M], dist2 [B,N])
    loss_on_cd           loss.py:200-205   (a method of Loss in the reference; a function here)
    loss_cd_stages       loss.py:260-267   (the gamma-weighted stage loop of Loss.forward)
    loss_cd_stages_fused the same loop as ONE batched call
    euler2rot, view_rotations, loss_on_depth, loss_on_mask
                         loss.py:73-95, 135-141, 147-196  the multi-view depth / mask losses on top of Render / Masker
    get_neighbor_index   loss.py:99-110    ARAP neighbour graph: the n nearest OTHER vertices -> [B,V,n] (int64)
    indexing_neighbor    loss.py:113-120   gather by that index (torch)

Differences from the reference that a caller can observe: distances are evaluated in fp32
difference form (accurate to ~2e-7 relative) instead of the expanded |x|^2+|y|^2-2x.y form (1e-2..1e-5), so they
are never negative; CPU tensors raise instead of running.
"""
import torch

from .. import ext
from .exfunc.functions.chamfer_func import ChamferCDFunction, ChamferFunction


def compute_sqrdis_map(points_x, points_y):
    """points_x [B,M,3], points_y [B,N,3] -> [B,M,N] squared distances (forward only)."""
    if points_x.requires_grad or points_y.requires_grad:
        if torch.is_grad_enabled():
            raise RuntimeError("compute_sqrdis_map is forward-only here; use chamfer_dist for the differentiable path")
    return ext.sqrdis_map(points_x, points_y)


def chamfer_dist(points_x, points_y):
    """points_x [B,M,3], points_y [B,N,3] -> (dist1 [B,M], dist2 [B,N]); differentiable w.r.t. both."""
    dist1, dist2, _, _ = ChamferFunction.apply(points_x, points_y)
    return dist1, dist2


def chamfer_dist_with_idx(points_x, points_y):
    """As chamfer_dist, plus the int32 argmin indices (first minimum, torch.min tie-break)."""
    return ChamferFunction.apply(points_x, points_y)


def loss_on_cd(deformation_p, p1):
    """(sum dist1 + sum dist2) * 0.5 / B, fused with chamfer_dist and its backward (ChamferCDFunction)."""
    return ChamferCDFunction.apply(deformation_p, p1)


def loss_on_cd_unfused(deformation_p, p1):
    """The reference's literal composition (chamfer_dist + torch sums); same value, ~10 more small kernels."""
    thisbatchsize = deformation_p.size()[0]
    dist1, dist2 = chamfer_dist(deformation_p, p1)
    return (torch.sum(dist1) + torch.sum(dist2)) * 0.5 / thisbatchsize


def loss_cd_stages(deformation_points_list, target_points, gamma=0.8, weight_cd=1.0):
    """sum_i gamma^(T-i-1) * loss_on_cd(deformation_i^T, target) * weight_cd; deformation_i is [B,3,N]."""
    iter_time = len(deformation_points_list)
    loss_cd = 0
    for i in range(iter_time):
        loss_cd = loss_cd + gamma ** (iter_time - i - 1) * loss_on_cd(
            deformation_points_list[i].transpose(1, 2), target_points)
    return loss_cd * weight_cd


def loss_cd_stages_fused(deformation_points_list, target_points, gamma=0.8, weight_cd=1.0):
    """The same value and gradients as loss_cd_stages, with all T stages in ONE chamfer call: the stage outputs are
    stacked into a [T*B,N,3] batch, the target is repeated, and the stage weights gamma^(T-i-1) * weight_cd enter the
    fused kernel as per-batch-element weights (rma_chamfer_cd_forward_weighted).  4 launches instead of 4*T."""
    iter_time = len(deformation_points_list)
    B = target_points.shape[0]
    x_all = torch.stack([d.transpose(1, 2) for d in deformation_points_list], 0)
    x_all = x_all.reshape(iter_time * B, x_all.shape[2], 3)
    y_all = target_points.unsqueeze(0).expand(iter_time, -1, -1, -1).reshape(iter_time * B, target_points.shape[1], 3)
    return ChamferCDFunction.apply(x_all, y_all, _stage_weights(iter_time, B, float(gamma), float(weight_cd),
                                                              target_points.device), 0.5 / B)


_STAGE_WEIGHTS = {}


def _stage_weights(iter_time, B, gamma, weight_cd, device):
    """[T*B] device vector weight_cd * gamma^(T-i-1), cached: building it costs a synchronous H2D copy."""
    key = (iter_time, B, gamma, weight_cd, str(device))
    w = _STAGE_WEIGHTS.get(key)
    if w is None:
        w = torch.tensor([weight_cd * gamma ** (iter_time - i - 1) for i in range(iter_time)],
                         dtype=torch.float32).repeat_interleave(B).to(device)
        if len(_STAGE_WEIGHTS) > 64:
            _STAGE_WEIGHTS.clear()
        _STAGE_WEIGHTS[key] = w
    return w


def get_neighbor_index(vertices, neighbor_num):
    """vertices [B,V,3] -> [B,V,neighbor_num] (int64): the reference takes the neighbor_num + 1 smallest entries of
    the expanded-form distance matrix and drops the first (loss.py:107-108).  Here: rma_knn with k = neighbor_num + 1,
    skip_first = 1, ordered by (exact distance, index); the dropped entry is the vertex itself unless an identical
    vertex with a lower index exists (the reference's choice in that case is whatever torch.topk returns first)."""
    return ext.knn_indices(vertices.detach(), vertices.detach(), neighbor_num + 1, skip_first=1).long()


def indexing_neighbor(tensor, index):
    """tensor [B,V,dim], index [B,V,n] -> [B,V,n,dim]   (loss.py:113-120)"""
    bs = index.size(0)
    id_0 = torch.arange(bs, device=tensor.device).view(-1, 1, 1)
    return tensor[id_0, index]


# ------------------------------------------------------------------------------------------------------------
# multi-view depth / mask losses (loss.py:33-37, 73-95, 135-196) on top of the Render / Masker kernels
# ------------------------------------------------------------------------------------------------------------
resolution_ = 256            # loss.py:33
threshold_ = 0.1             # loss.py:36
threshold_2_ = 1e-10         # loss.py:37


def euler2rot(euler_angle):
    """[n,3] Euler angles -> [n,3,3] = Rz(psi) Ry(phi) Rx(theta)   (loss.py:73-95)"""
    n = euler_angle.shape[0]
    one = torch.ones(n, 1, 1, device=euler_angle.device)
    zero = torch.zeros(n, 1, 1, device=euler_angle.device)
    theta = euler_angle[:, 0].reshape(-1, 1, 1)
    phi = euler_angle[:, 1].reshape(-1, 1, 1)
    psi = euler_angle[:, 2].reshape(-1, 1, 1)
    rot_x = torch.cat((torch.cat((one, zero, zero), 1), torch.cat((zero, theta.cos(), theta.sin()), 1),
                       torch.cat((zero, -theta.sin(), theta.cos()), 1)), 2)
    rot_y = torch.cat((torch.cat((phi.cos(), zero, -phi.sin()), 1), torch.cat((zero, one, zero), 1),
                       torch.cat((phi.sin(), zero, phi.cos()), 1)), 2)
    rot_z = torch.cat((torch.cat((psi.cos(), -psi.sin(), zero), 1), torch.cat((psi.sin(), psi.cos(), zero), 1),
                       torch.cat((zero, zero, one), 1)), 2)
    return torch.bmm(rot_z, torch.bmm(rot_y, rot_x))


def view_rotations(view_divide, device, jitter=None):
    """The view_divide^2 view rotations as the reference stacks them, [3*V, 3] (`all_rot_matrix_for_lfd.view(-1,3)`):
    the regular grid of Loss.__init__ (loss.py:135-141) or, with jitter [V,2] in [0,1), the re-sampled grid of
    loss_on_depth (loss.py:151-159, where the offsets come from random.random())."""
    v = torch.arange(view_divide ** 2, device=device)
    ex = (v // view_divide).float()
    ey = (v % view_divide).float()
    if jitter is not None:
        ex = ex + jitter[:, 0].to(device)
        ey = ey + jitter[:, 1].to(device)
    pi = torch.acos(torch.zeros(1)).item() * 2
    ang = torch.stack([-pi + 2 * pi * ex / view_divide, -pi + 2 * pi * ey / view_divide, torch.zeros_like(ex)], 1)
    return euler2rot(ang).reshape(-1, 3)


def _views_to_batch(points, all_rot):
    """[B,N,3] -> [V*B, N, 3]: bmm with the stacked rotations (loss.py:161) and the views folded into the batch."""
    B, N = points.shape[0], points.shape[1]
    V = all_rot.shape[0] // 3
    pv = torch.bmm(all_rot.view(1, -1, 3).expand(B, -1, -1), points.transpose(1, 2)).transpose(1, 2)   # [B,N,3V]
    return pv.reshape(B, N, V, 3).permute(2, 0, 1, 3).reshape(V * B, N, 3), V


def loss_on_depth(deformation_p, p1, all_rot, renderer=None, threshold=threshold_, resolution=resolution_,
                  batched=True):
    """loss.py:147-183 without the in-place re-sampling of the views (pass them in: view_rotations(..., jitter)).
    batched=True folds the V views into the batch dimension: 2 Render calls instead of 2V, same value and gradient
    (up to fp32 summation order); batched=False is the reference's literal per-view loop."""
    from .exfunc.functions import Render
    renderer = renderer or Render(resolution, resolution, 7, 9, 1e-5)            # loss.py:34

    def project(pv):           # loss.py:166: xy -> pixels, z - 1
        return torch.cat(((pv[..., :2] + 1.) * (resolution - 1) / 2, pv[..., 2].unsqueeze(-1) - 1), -1).contiguous()

    def view_term(ori_proj, tar_proj):                                           # loss.py:169-181
        od, ow, _ = renderer(ori_proj, threshold)
        td, tw, _ = renderer(tar_proj, threshold)
        ori_img, tar_img = od / ow, td / tw
        mask = torch.sign(ori_img * tar_img).detach()
        dis = torch.abs(ori_img - tar_img)
        thres = -1.0 * (dis.detach() - 0.05)
        mask2 = (torch.sign(torch.sign(thres).detach() + 0.5) + 1.0) * 0.5
        mask = torch.mul(mask2, mask).detach()
        return torch.sum(torch.mul(mask, torch.mul(dis, dis)))

    if batched:
        a, V = _views_to_batch(deformation_p, all_rot)
        b, _ = _views_to_batch(p1, all_rot)
        return view_term(project(a), project(b)) / V
    V = all_rot.shape[0] // 3
    B = deformation_p.shape[0]
    dv = torch.bmm(all_rot.view(1, -1, 3).expand(B, -1, -1), deformation_p.transpose(1, 2)).transpose(1, 2)
    tv = torch.bmm(all_rot.view(1, -1, 3).expand(B, -1, -1), p1.transpose(1, 2)).transpose(1, 2)
    out = 0
    for v in range(V):
        out = out + view_term(project(dv[:, :, 3 * v:3 * v + 3]), project(tv[:, :, 3 * v:3 * v + 3])) / V
    return out


def loss_on_mask(deformation_p, p1, all_rot, masker=None, threshold=threshold_2_, resolution=resolution_,
                 batched=True):
    """loss.py:185-196: mean |mask(deformation) - mask(target)| averaged over the views; the projection scales all
    three coordinates by resolution/2 after adding (1, 1, -1), as the reference does (:190)."""
    from .exfunc.functions import Masker
    masker = masker or Masker(resolution, resolution, 5, 1e-5)                   # loss.py:35
    shift = torch.tensor([[[1., 1., -1.]]], device=deformation_p.device)

    def project(pv):
        return ((pv + shift) * resolution / 2.).contiguous()

    if batched:
        a, V = _views_to_batch(deformation_p, all_rot)
        b, _ = _views_to_batch(p1, all_rot)
        return torch.mean(torch.abs(masker(project(a), threshold) - masker(project(b), threshold)))
    V = all_rot.shape[0] // 3
    B = deformation_p.shape[0]
    dv = torch.bmm(all_rot.view(1, -1, 3).expand(B, -1, -1), deformation_p.transpose(1, 2)).transpose(1, 2)
    tv = torch.bmm(all_rot.view(1, -1, 3).expand(B, -1, -1), p1.transpose(1, 2)).transpose(1, 2)
    out = 0
    for v in range(V):
        out = out + torch.mean(torch.abs(masker(project(dv[:, :, 3 * v:3 * v + 3]), threshold)
                                         - masker(project(tv[:, :, 3 * v:3 * v + 3]), threshold))) / V
    return out
