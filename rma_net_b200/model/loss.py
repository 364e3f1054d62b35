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
    loss_on_pd / loss_on_sparse / loss_on_tran / loss_on_arap
                         loss.py:207-243   the other terms (arap = one fused CUDA launch, the rest torch one-liners)
    Loss                 loss.py:122-355   the module: same constructor argument (args), same forward signature and
                                           return value (loss, loss_stages); fused=True runs every term once over
                                           the stacked [T*B,N,3] stage outputs instead of T times
    stage_chamfer_errors train_view.py:376-380  per-stage metric from the dists of the training forward

Differences from the reference that a caller can observe: distances are evaluated in fp32
difference form (accurate to ~2e-7 relative) instead of the expanded |x|^2+|y|^2-2x.y form (1e-2..1e-5), so they
are never negative; CPU tensors raise instead of running.
"""
import math
import random

import torch

from .. import ext
from .exfunc.functions.chamfer_func import (ArapFunction, ChamferCDFunction, ChamferCDMetricsFunction,
                                            ChamferFunction)


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


def stack_stages(deformation_points_list):
    """T stage outputs [B,3,N] (as the network returns them) -> one contiguous [T,B,N,3] tensor."""
    return torch.stack([d.transpose(1, 2) for d in deformation_points_list], 0)


def loss_cd_stages_fused(deformation_points_list, target_points, gamma=0.8, weight_cd=1.0, return_dist=False,
                         stacked=None):
    """The same value and gradients as loss_cd_stages, with all T stages in ONE chamfer call: the stage outputs are
    stacked into a [T*B,N,3] batch against the ONE target (y_batch_period = B: no copies of it), and the stage weights gamma^(T-i-1) * weight_cd enter the
    fused kernel as per-batch-element weights (rma_chamfer_cd_forward_weighted).  4 launches instead of 4*T.
    return_dist=True -> (loss, dist1 [T,B,N], dist2 [T,B,M]): the unweighted distances of the same forward, for
    stage_chamfer_errors.  stacked: the [T,B,N,3] tensor of stack_stages, if the caller already has it."""
    iter_time = len(deformation_points_list) if stacked is None else stacked.shape[0]
    B = target_points.shape[0]
    x_all = stack_stages(deformation_points_list) if stacked is None else stacked
    x_all = x_all.reshape(iter_time * B, x_all.shape[2], 3)
    # the target is NOT repeated T times: batch element t * B + b of the stack reads target_points[b] (y_period = B)
    w = _stage_weights(iter_time, B, float(gamma), float(weight_cd), target_points.device)
    if not return_dist:
        return ChamferCDFunction.apply(x_all, target_points, w, 0.5 / B, B)
    loss, d1, d2 = ChamferCDMetricsFunction.apply(x_all, target_points, w, 0.5 / B, B)
    return loss, d1.view(iter_time, B, -1), d2.view(iter_time, B, -1)


def stage_chamfer_errors(dist1, dist2):
    """[T,B,N], [T,B,N] -> [T] device tensor: mean(dist1 + dist2) * 0.5 per stage — the `cd_list` metric of
    train_view.py:376-380 (which calls chamfer_dist(target, deformation_i) again under no_grad; the sum is symmetric
    in the two directions, so the dists of the training forward give the same numbers).  One .tolist() reads all T."""
    return (dist1 + dist2).mean(dim=(1, 2)) * 0.5


_STAGE_WEIGHTS = {}


def _stage_weights(iter_time, B, gamma, weight_cd, device, factors=None):
    """[T*B] device vector weight_cd * gamma^(T-i-1) (* factors[i]), cached: building it costs a synchronous H2D copy."""
    key = (iter_time, B, gamma, weight_cd, str(device), factors)
    w = _STAGE_WEIGHTS.get(key)
    if w is None:
        w = torch.tensor([weight_cd * gamma ** (iter_time - i - 1) * (factors[i] if factors else 1.0)
                          for i in range(iter_time)], dtype=torch.float32).repeat_interleave(B).to(device)
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


def loss_on_pd(deformation_p, p1):
    """point-to-point L2 term, sum((deformation - target)^2) / B   (loss.py:207-212)"""
    dist = deformation_p - p1
    return torch.sum(torch.mul(dist, dist)) / deformation_p.size()[0]


def loss_on_sparse(point_weight):
    """mean |point_weight|   (loss.py:217-220)"""
    return torch.mean(torch.abs(point_weight))


def loss_on_tran(rigid_matrix, trans_mask=None):
    """sum of the squared translation column of the [B,4,4] matrices / B   (loss.py:224-230, mask of loss.py:134-135)"""
    t = rigid_matrix[:, 0:3, 3] if trans_mask is None else torch.mul(trans_mask.expand(rigid_matrix.size()[0], -1, -1),
                                                                     rigid_matrix)
    return torch.sum(torch.mul(t, t)) / rigid_matrix.size()[0]


def loss_on_arap(neighbour_indexes, deformation_p, source_points):
    """sum over vertices and neighbours of (|edge| after - |edge| before)^2 / B, |e| = sqrt(e.e + 1e-5)
    (loss.py:233-243) as one fused launch (ArapFunction); neighbour_indexes [B,V,n] int32 or int64."""
    return ArapFunction.apply(deformation_p, source_points, neighbour_indexes, None,
                              1.0 / deformation_p.size()[0])


def arap_stage_factor(i):
    """the hard-wired stage factors of loss.py:337-343"""
    return 10000.0 if i <= 1 else (1000.0 if i <= 3 else 100.0)


def loss_arap_stages(neighbour_indexes, deformation_points_list, source_points, gamma=0.8):
    """the stage loop of loss.py:334-345, one loss_on_arap per stage"""
    iter_time = len(deformation_points_list)
    out = 0
    for i in range(iter_time):
        out = out + arap_stage_factor(i) * gamma ** (iter_time - i - 1) * loss_on_arap(
            neighbour_indexes, deformation_points_list[i].transpose(1, 2), source_points)
    return out


def loss_arap_stages_fused(neighbour_indexes, deformation_points_list, source_points, gamma=0.8, stacked=None):
    """the same value and gradients as loss_arap_stages in ONE launch over the stacked [T*B,V,3] stage outputs"""
    x_all = stack_stages(deformation_points_list) if stacked is None else stacked
    T, B = x_all.shape[0], x_all.shape[1]
    w = _stage_weights(T, B, float(gamma), 1.0, x_all.device, tuple(arap_stage_factor(i) for i in range(T)))
    return ArapFunction.apply(x_all.reshape(T * B, x_all.shape[2], 3), source_points, neighbour_indexes, w, 1.0 / B)


# ------------------------------------------------------------------------------------------------------------
# multi-view depth / mask losses (loss.py:33-37, 73-95, 135-196) on top of the Render / Masker kernels
# ------------------------------------------------------------------------------------------------------------
resolution_ = 256            # loss.py:33
threshold_ = 0.1             # loss.py:36
threshold_2_ = 1e-10         # loss.py:37


def _plane_rotation(angle, i, j):
    """[n] angles -> [n,3,3] rotations acting in the coordinate plane (i, j): m[i,i] = m[j,j] = cos, m[i,j] = -sin,
    m[j,i] = sin, identity elsewhere."""
    m = torch.eye(3, dtype=angle.dtype, device=angle.device).repeat(angle.shape[0], 1, 1)
    c, s = angle.cos(), angle.sin()
    m[:, i, i] = c
    m[:, j, j] = c
    m[:, i, j] = -s
    m[:, j, i] = s
    return m


def euler2rot(euler_angle):
    """[n,3] Euler angles (theta, phi, psi) -> [n,3,3] = Rz(psi) Ry(phi) Rx(theta) with the reference's sign
    conventions (loss.py:73-95): Rx turns the (y,z) plane, Ry the (z,x) plane, Rz the (y,x) plane; the product is
    formed by two bmm calls in the reference's order, so the fp32 entries are the same."""
    rot_x = _plane_rotation(euler_angle[:, 0], 1, 2)
    rot_y = _plane_rotation(euler_angle[:, 1], 2, 0)
    rot_z = _plane_rotation(euler_angle[:, 2], 1, 0)
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
                  batched=True, fused_compare=True, fused_project=True):
    """loss.py:147-183 without the in-place re-sampling of the views (pass them in: view_rotations(..., jitter)).
    batched=True folds the V views into the batch dimension: 2 Render calls instead of 2V, same value and gradient
    (up to fp32 summation order); batched=False is the reference's literal per-view loop.  fused_compare=True runs
    the image comparison (loss.py:169-181) as one kernel (DepthViewTermFunction) instead of ~14 torch kernels;
    fused_project=True the rotate-and-map-to-pixels step (loss.py:161-166) as one kernel (ViewProjectFunction)."""
    from .exfunc.functions import DepthViewTermFunction, Render
    renderer = renderer or Render(resolution, resolution, 7, 9, 1e-5)            # loss.py:34

    def project(pv):           # loss.py:166: xy -> pixels, z - 1
        return torch.cat(((pv[..., :2] + 1.) * (resolution - 1) / 2, pv[..., 2].unsqueeze(-1) - 1), -1).contiguous()

    def view_term(ori_proj, tar_proj):                                           # loss.py:169-181
        od, ow, _ = renderer(ori_proj, threshold)
        td, tw, _ = renderer(tar_proj, threshold)
        if fused_compare and not ((td.requires_grad or tw.requires_grad) and torch.is_grad_enabled()):
            return DepthViewTermFunction.apply(od, ow, td, tw, 1.0, 0.05)
        ori_img, tar_img = od / ow, td / tw
        mask = torch.sign(ori_img * tar_img).detach()
        dis = torch.abs(ori_img - tar_img)
        thres = -1.0 * (dis.detach() - 0.05)
        mask2 = (torch.sign(torch.sign(thres).detach() + 0.5) + 1.0) * 0.5
        mask = torch.mul(mask2, mask).detach()
        return torch.sum(torch.mul(mask, torch.mul(dis, dis)))

    if batched:
        V = all_rot.shape[0] // 3
        if fused_project:
            from .exfunc.functions import ViewProjectFunction
            add3, mul3 = (1., 1., -1.), ((resolution - 1) / 2, (resolution - 1) / 2, 1.)
            return view_term(ViewProjectFunction.apply(deformation_p, all_rot, add3, mul3),
                             ViewProjectFunction.apply(p1, all_rot, add3, mul3)) / V
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


def target_view_masks(p1, all_rot, masker=None, threshold=threshold_2_, resolution=resolution_):
    """The target's masks for all views, [V*B,H,W,1] (no gradient): what loss_on_mask computes for `p1`.  The views of
    the mask term do not change between the stages of one Loss.forward, so it can be computed once per step."""
    from .exfunc.functions import Masker, ViewProjectFunction
    masker = masker or Masker(resolution, resolution, 5, 1e-5)
    with torch.no_grad():
        return masker(ViewProjectFunction.apply(p1, all_rot, (1., 1., -1.), (resolution / 2.,) * 3), threshold)


def loss_on_mask(deformation_p, p1, all_rot, masker=None, threshold=threshold_2_, resolution=resolution_,
                 batched=True, fused_project=True, target_masks=None):
    """loss.py:185-196: mean |mask(deformation) - mask(target)| averaged over the views; the projection scales all
    three coordinates by resolution/2 after adding (1, 1, -1), as the reference does (:190).  target_masks: the
    result of target_view_masks(p1, all_rot, ...) if the caller already has it (batched + fused_project path)."""
    from .exfunc.functions import Masker
    masker = masker or Masker(resolution, resolution, 5, 1e-5)                   # loss.py:35
    shift = torch.tensor([[[1., 1., -1.]]], device=deformation_p.device)

    def project(pv):
        return ((pv + shift) * resolution / 2.).contiguous()

    if batched and fused_project:
        from .exfunc.functions import ViewProjectFunction
        add3, mul3 = (1., 1., -1.), (resolution / 2.,) * 3
        if target_masks is None:
            target_masks = masker(ViewProjectFunction.apply(p1, all_rot, add3, mul3), threshold)
        return torch.mean(torch.abs(masker(ViewProjectFunction.apply(deformation_p, all_rot, add3, mul3), threshold)
                                    - target_masks))
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


# ------------------------------------------------------------------------------------------------------------
# the Loss module (loss.py:122-355): same constructor argument, forward signature and return value
# ------------------------------------------------------------------------------------------------------------
def sample_view_angles(view_divide, jitter=True, rng=random):
    """The view_divide^2 Euler angle triples in host double precision, exactly as the reference computes them:
    the regular grid of Loss.__init__ (loss.py:137-140) or, with jitter, the re-sampled grid of loss_on_depth
    (loss.py:152-158), drawing rng.random() twice per view in the reference's order (x, then y)."""
    # the reference's pi (loss.py:29: torch.pi = torch.acos(torch.zeros(1)).item() * 2) is the float32-rounded value
    # 3.1415927410125732, not math.pi: after the float32 cast of the angles the two can differ by 1 ulp, which moves
    # points across pixel boundaries in Render / Masker
    pi = torch.acos(torch.zeros(1)).item() * 2
    out = []
    for view_id in range(view_divide ** 2):
        euler_x = view_id // view_divide
        euler_y = view_id - euler_x * view_divide
        rx = rng.random() if jitter else 0.0
        ry = rng.random() if jitter else 0.0
        out.append([-pi + 2 * pi * (euler_x + rx) / view_divide,
                    -pi + 2 * pi * (euler_y + ry) / view_divide, 0.0])
    return out


def rotations_from_angles(angles, device):
    """[V][3] host angles -> [3V,3] stacked rotations on the device (one H2D copy, one batched euler2rot)."""
    return euler2rot(torch.tensor(angles, dtype=torch.float32).to(device)).reshape(-1, 3)


class Loss(torch.nn.Module):
    """Drop-in for the reference's `Loss(args)` (model/loss.py:122-355).  args needs weight_cd, weight_pd,
    weight_sparse, weight_depth, weight_tran, weight_mask, weight_arap, gamma, neighbour_num, view_divide
    (utils/config.py:41-92, 128).  forward(phi_list, point_weight_list, deform_rigid_points_list,
    deformation_points_list, rigid_matrix_list, source_points, target_points) -> (loss, loss_stages) with
    loss_stages = [cd, pd, sparse, depth, tran, mask, arap] (a zero tensor where the weight is not positive).

    fused=True (default): the T stage outputs are stacked once into [T,B,N,3]; cd and arap each run as ONE weighted
    call over the T*B batch, pd / sparse / tran as one stacked torch expression, depth / mask per stage with the V
    views folded into the batch.  fused=False is the reference's literal per-stage loop on the same kernels.
    After forward, `self.last_stage_dists` holds (dist1 [T,B,N], dist2 [T,B,M]) of the cd term (fused path, cd on) for
    stage_chamfer_errors.  Unlike the reference the module does not pin itself to cuda:0 (loss.py:134-141): buffers
    follow `device` (default: the current CUDA device)."""

    def __init__(self, args, device=None, fused=True, rng=random):
        super().__init__()
        self.args = args
        self.fused = fused
        self.rng = rng
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        trans_mask = torch.zeros([1, 4, 4], dtype=torch.float32, device=device)
        trans_mask[:, 0:3, 3] += 1.0
        self.register_buffer("trans_mask", trans_mask, persistent=False)
        self.all_rot_matrix_for_lfd = rotations_from_angles(sample_view_angles(args.view_divide, jitter=False), device)
        self.last_stage_dists = None
        self._renderer = None
        self._masker = None

    # ---- per-term methods, reference names ------------------------------------------------------------------
    def loss_on_cd(self, deformation_p, p1):
        return loss_on_cd(deformation_p, p1)

    def loss_on_pd(self, deformation_p, p1):
        return loss_on_pd(deformation_p, p1)

    def loss_on_sparse(self, point_weight):
        return loss_on_sparse(point_weight)

    def loss_on_tran(self, rigid_matrix):
        return loss_on_tran(rigid_matrix, self.trans_mask)

    def loss_on_arap(self, neighbour_indexes, deformation_p, source_points):
        return loss_on_arap(neighbour_indexes, deformation_p, source_points)

    def loss_on_depth(self, deformation_p, p1):
        """re-samples the views (loss.py:151-159), then the batched-view depth term"""
        from .exfunc.functions import Render
        self.all_rot_matrix_for_lfd = rotations_from_angles(
            sample_view_angles(self.args.view_divide, jitter=True, rng=self.rng), deformation_p.device)
        if self._renderer is None:
            self._renderer = Render(resolution_, resolution_, 7, 9, 1e-5)
        return loss_on_depth(deformation_p, p1, self.all_rot_matrix_for_lfd, self._renderer)

    def loss_on_mask(self, deformation_p, p1, target_masks=None):
        from .exfunc.functions import Masker
        if self._masker is None:
            self._masker = Masker(resolution_, resolution_, 5, 1e-5)
        return loss_on_mask(deformation_p, p1, self.all_rot_matrix_for_lfd.to(deformation_p.device), self._masker,
                            target_masks=target_masks)

    # ---- forward ---------------------------------------------------------------------------------------------
    def forward(self, phi_list, point_weight_list, deform_rigid_points_list, deformation_points_list,
                rigid_matrix_list, source_points, target_points):
        a = self.args
        iter_time = len(phi_list)
        dev = phi_list[0].device
        B = phi_list[0].size()[0]
        gam = [a.gamma ** (iter_time - i - 1) for i in range(iter_time)]
        zero_tensor = torch.zeros((), dtype=torch.float, device=dev)
        loss = torch.zeros((), dtype=torch.float, device=dev)
        loss_stages = []
        fused = self.fused
        stacked = stack_stages(deformation_points_list) if fused else None          # [T,B,N,3]
        stage = [d.transpose(1, 2) for d in deformation_points_list]
        self.last_stage_dists = None

        def stage_vec():
            return _stage_weights(iter_time, 1, float(a.gamma), 1.0, dev)              # [T]

        neighbour_indexes_ = None
        if a.weight_arap > 0:                                                        # loss.py:253 (always, there)
            neighbour_indexes_ = ext.knn_indices(source_points.detach(), source_points.detach(),
                                                 a.neighbour_num + 1, skip_first=1)

        if a.weight_cd > 0:                                                          # loss.py:260-268
            if fused:
                loss_cd, d1, d2 = loss_cd_stages_fused(None, target_points, a.gamma, 1.0, return_dist=True,
                                                       stacked=stacked)
                self.last_stage_dists = (d1, d2)
            else:
                loss_cd = sum(gam[i] * self.loss_on_cd(stage[i], target_points) for i in range(iter_time))
            loss = loss + loss_cd * a.weight_cd
            loss_stages.append(loss_cd)
        else:
            loss_stages.append(zero_tensor)

        if a.weight_pd > 0:                                                          # loss.py:272-284
            if fused:
                diff = stacked - target_points.unsqueeze(0)
                loss_pd = torch.mul(diff, diff).sum(dim=(1, 2, 3)).dot(stage_vec()) / B
            else:
                loss_pd = sum(gam[i] * self.loss_on_pd(stage[i], target_points) for i in range(iter_time))
            loss = loss + loss_pd * a.weight_pd
            loss_stages.append(loss_pd)
        else:
            loss_stages.append(zero_tensor)

        if a.weight_sparse > 0:                                                      # loss.py:289-298
            if fused:
                loss_sparse = torch.stack(list(point_weight_list), 0).abs().flatten(1).mean(dim=1).dot(stage_vec())
            else:
                loss_sparse = sum(gam[i] * self.loss_on_sparse(point_weight_list[i]) for i in range(iter_time))
            loss = loss + loss_sparse * a.weight_sparse
            loss_stages.append(loss_sparse)
        else:
            loss_stages.append(zero_tensor)

        if a.weight_depth > 0:                                                       # loss.py:300-308
            loss_depth = sum(gam[i] * self.loss_on_depth(stage[i], target_points) for i in range(iter_time))
            loss = loss + loss_depth * a.weight_depth
            loss_stages.append(loss_depth)
        else:
            loss_stages.append(zero_tensor)

        if a.weight_tran > 0:                                                        # loss.py:311-319
            if fused:
                t = torch.stack(list(rigid_matrix_list), 0)[:, :, 0:3, 3]
                loss_tran = torch.mul(t, t).sum(dim=(1, 2)).dot(stage_vec()) / B
            else:
                loss_tran = sum(gam[i] * self.loss_on_tran(rigid_matrix_list[i]) for i in range(iter_time))
            loss = loss + loss_tran * a.weight_tran
            loss_stages.append(loss_tran)
        else:
            loss_stages.append(zero_tensor)

        if a.weight_mask > 0:                                                        # loss.py:321-330
            tmasks = None
            if fused and not (target_points.requires_grad and torch.is_grad_enabled()):
                # every stage of this term uses the views of the last depth call: one target render instead of T
                from .exfunc.functions import Masker
                if self._masker is None:
                    self._masker = Masker(resolution_, resolution_, 5, 1e-5)
                tmasks = target_view_masks(target_points, self.all_rot_matrix_for_lfd.to(dev), self._masker)
            loss_mask = sum(gam[i] * self.loss_on_mask(stage[i], target_points, tmasks) for i in range(iter_time))
            loss = loss + loss_mask * a.weight_mask
            loss_stages.append(loss_mask)
        else:
            loss_stages.append(zero_tensor)

        if a.weight_arap > 0:                                                        # loss.py:333-347
            if fused:
                loss_arap = loss_arap_stages_fused(neighbour_indexes_, None, source_points, a.gamma, stacked=stacked)
            else:
                loss_arap = loss_arap_stages(neighbour_indexes_, deformation_points_list, source_points, a.gamma)
            loss = loss + loss_arap * a.weight_arap
            loss_stages.append(loss_arap)
        else:
            loss_stages.append(zero_tensor)

        return loss, loss_stages
