"""Tensor-level `forward` / `backward` entry points over the C ABI (include/rma_b200.h).

This plays the role of the pybind11 modules of the reference's extension (`point_render.forward/backward`,
model/exfunc/expackages/point_render/point_render.cpp:40-139): validate tensors the way CHECK_CUDA /
CHECK_CONTIGUOUS do (point_render.cpp:21-27) — but raising instead of exiting, and accepting strided views —
allocate outputs and scratch with torch, and hand raw device pointers + the current stream to the kernels.
PyTorch is used for memory and streams only; there is no CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes
import threading
from typing import Optional, Tuple

import torch

from . import _lib


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _check_points(name: str, t: torch.Tensor) -> None:
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (rma_net_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 3 or t.shape[-1] != 3:
        raise RuntimeError(f"{name} must have shape [B, N, 3], got {tuple(t.shape)}")


def _same_device(*ts: torch.Tensor) -> None:
    dev = ts[0].device
    for t in ts[1:]:
        if t is not None and t.device != dev:
            raise RuntimeError("all tensors must be on the same CUDA device")


# ------------------------------------------------------------------------------------------------------------
# chamfer
# ------------------------------------------------------------------------------------------------------------
# `rma_chamfer_opts` is a per-call argument of the C ABI (the library keeps no state).  On the Python side the
# options of the enclosing `chamfer_options(...)` block (thread-local) are passed with every call; outside such a
# block NULL is passed = the library's defaults.
_tls = threading.local()
_PATHS = {None: _lib.PATH_AUTO, "auto": _lib.PATH_AUTO, "exact": _lib.PATH_EXACT, "filter": _lib.PATH_FILTER,
          0: _lib.PATH_AUTO, 1: _lib.PATH_EXACT, 2: _lib.PATH_FILTER}


def _opts():
    o = getattr(_tls, "opts", None)
    return ctypes.byref(o) if o is not None else None


@contextlib.contextmanager
def chamfer_options(path=None, unit_run: int = 0, record_cap_bytes: int = 0):
    """Run the chamfer calls of this block (and this thread) with explicit options: path "auto" | "exact" | "filter"
    (both give identical results; parity tests and A/B timing force one), unit_run (32-column work units per search CTA) and the
    record cap of the automatic choice.  Backward passes of ops created inside the block do not depend on it."""
    prev = getattr(_tls, "opts", None)
    _tls.opts = _lib.ChamferOpts(_PATHS[path], int(unit_run), int(record_cap_bytes))
    try:
        yield
    finally:
        _tls.opts = prev


def chamfer_workspace_bytes(B: int, N: int, M: int) -> int:
    return int(_lib.load().rma_chamfer_workspace_bytes(B, N, M, _opts()))


def chamfer_path(B: int, N: int, M: int) -> str:
    """Which search implementation (B,N,M) runs under the current options: "filter" or "exact"."""
    rc = _lib.load().rma_chamfer_path(B, N, M, _opts())
    if rc < 0:
        raise RuntimeError(f"problem size not supported: B={B}, N={N}, M={M}")
    return "filter" if rc == _lib.PATH_FILTER else "exact"


def _alloc_workspace(lib, B: int, N: int, M: int, dev, y_period: int = 0):
    """(workspace tensor, bytes, opts pointer, keep-alive) for one chamfer call under the current options.  The filter
    path of a single huge pair wants tens of GB (40 GiB for 2^20 x 2^20: include/rma_b200.h); if torch cannot give that,
    the same call runs on the exact path - identical results, a few MiB of workspace - instead of failing."""
    cur = getattr(_tls, "opts", None)
    o = None
    if cur is not None or y_period:
        o = _lib.ChamferOpts(cur.path if cur else _lib.PATH_AUTO, cur.unit_run if cur else 0,
                             cur.record_cap_bytes if cur else 0, int(y_period), 0)
    opts = ctypes.byref(o) if o is not None else None
    wsb = lib.rma_chamfer_workspace_bytes(B, N, M, opts)
    if wsb == 0:
        raise RuntimeError(f"problem size not supported: B={B}, N={N}, M={M}")
    try:
        return torch.empty(wsb, dtype=torch.uint8, device=dev), wsb, opts, o
    except torch.OutOfMemoryError:
        if wsb < (1 << 30) or (cur is not None and cur.path == _lib.PATH_FILTER):
            raise
        import warnings
        warnings.warn(f"rma_net_b200: {wsb / 2**30:.1f} GiB of filter-path workspace for B={B}, N={N}, M={M} do not fit; "
                      "running the exact search path (same results)")
        o = _lib.ChamferOpts(_lib.PATH_EXACT, 0, 0, int(y_period), 0)
        opts = ctypes.byref(o)
        wsb = lib.rma_chamfer_workspace_bytes(B, N, M, opts)
        return torch.empty(wsb, dtype=torch.uint8, device=dev), wsb, opts, o


def chamfer_forward(points_x: torch.Tensor, points_y: torch.Tensor, want_idx: bool = True):
    """[B,N,3], [B,M,3] (any strides) -> dist1 [B,N], dist2 [B,M], idx1 [B,N] int32, idx2 [B,M] int32."""
    _check_points("points_x", points_x)
    _check_points("points_y", points_y)
    _same_device(points_x, points_y)
    B, N, M = points_x.shape[0], points_x.shape[1], points_y.shape[1]
    if points_y.shape[0] != B:
        raise RuntimeError(f"batch sizes differ: {B} vs {points_y.shape[0]}")
    if B == 0 or N == 0 or M == 0:
        raise RuntimeError("chamfer_dist needs non-empty point sets (min over an empty dimension)")
    lib = _lib.load()
    with torch.cuda.device(points_x.device):
        dev = points_x.device
        ws, wsb, opts, _keep = _alloc_workspace(lib, B, N, M, dev)
        dist1 = torch.empty((B, N), dtype=torch.float32, device=dev)
        dist2 = torch.empty((B, M), dtype=torch.float32, device=dev)
        idx1 = torch.empty((B, N), dtype=torch.int32, device=dev) if want_idx else None
        idx2 = torch.empty((B, M), dtype=torch.int32, device=dev) if want_idx else None
        xs, ys = points_x.stride(), points_y.stride()
        rc = lib.rma_chamfer_forward(_ptr(points_x), xs[0], xs[1], xs[2], _ptr(points_y), ys[0], ys[1], ys[2],
                                     B, N, M, _ptr(dist1), _ptr(dist2), _ptr(idx1), _ptr(idx2),
                                     _ptr(ws), wsb, opts, _stream())
        _lib.check(rc, "rma_chamfer_forward")
    return dist1, dist2, idx1, idx2


def chamfer_backward(points_x, points_y, idx1, idx2, g1, g2, need_x: bool = True, need_y: bool = True):
    """Returns (grad_x [B,N,3] or None, grad_y [B,M,3] or None)."""
    _check_points("points_x", points_x)
    _check_points("points_y", points_y)
    B, N, M = points_x.shape[0], points_x.shape[1], points_y.shape[1]
    dev = points_x.device
    if g1 is not None:
        g1 = g1.to(torch.float32).contiguous()
    if g2 is not None:
        g2 = g2.to(torch.float32).contiguous()
    lib = _lib.load()
    with torch.cuda.device(dev):
        gx = torch.empty((B, N, 3), dtype=torch.float32, device=dev) if need_x else None
        gy = torch.empty((B, M, 3), dtype=torch.float32, device=dev) if need_y else None
        xs, ys = points_x.stride(), points_y.stride()
        rc = lib.rma_chamfer_backward(_ptr(points_x), xs[0], xs[1], xs[2], _ptr(points_y), ys[0], ys[1], ys[2],
                                      B, N, M, _ptr(g1), _ptr(g2), _ptr(idx1), _ptr(idx2), _ptr(gx), _ptr(gy),
                                      _stream())
        _lib.check(rc, "rma_chamfer_backward")
    return gx, gy


def chamfer_cd_forward(points_x: torch.Tensor, points_y: torch.Tensor, need_x: bool, need_y: bool,
                       batch_weight: Optional[torch.Tensor] = None, coef: Optional[float] = None,
                       want_dist: bool = False, y_period: int = 0):
    """Fused loss_on_cd: returns (loss 0-dim, gx_parts, gy_parts); g*_parts = (own [.,.,3], scat [.,.,4]) or None,
    with d loss/d cloud = own + scat[..., :3] (combined and scaled by chamfer_cd_backward).
    loss = coef * sum_b batch_weight[b] * (sum dist1[b] + sum dist2[b]); defaults: weights 1, coef 0.5 / B.
    want_dist=True appends (dist1 [B,N], dist2 [B,M]) — the unweighted distances, for the metric callers
    (train_view.py:376-380) that would otherwise repeat the forward.
    y_period > 0: points_y holds y_period clouds and batch element b of points_x is matched against
    points_y[b % y_period] (the stage loop's shared target, rma_chamfer_opts.y_batch_period); every output, the
    gradient halves of points_y included, stays per batch element of points_x ([B,M,.])."""
    _check_points("points_x", points_x)
    _check_points("points_y", points_y)
    _same_device(points_x, points_y)
    B, N, M = points_x.shape[0], points_x.shape[1], points_y.shape[1]
    y_period = int(y_period or 0)
    if y_period:
        if points_y.shape[0] != y_period or B % y_period != 0:
            raise RuntimeError(f"y_period={y_period}: points_y must hold {y_period} clouds and divide the batch of {B}")
    elif points_y.shape[0] != B:
        raise RuntimeError(f"batch sizes differ: {B} vs {points_y.shape[0]}")
    if B == 0 or N == 0 or M == 0:
        raise RuntimeError("chamfer_dist needs non-empty point sets (min over an empty dimension)")
    lib = _lib.load()
    dev = points_x.device
    with torch.cuda.device(dev):
        ws, wsb, opts, _keep = _alloc_workspace(lib, B, N, M, dev, y_period)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        gx = (torch.empty((B, N, 3), **f32), torch.empty((B, N, 4), **f32)) if need_x else None
        gy = (torch.empty((B, M, 3), **f32), torch.empty((B, M, 4), **f32)) if need_y else None
        d1 = torch.empty((B, N), **f32) if want_dist else None
        d2 = torch.empty((B, M), **f32) if want_dist else None
        xs, ys = points_x.stride(), points_y.stride()
        if batch_weight is not None:
            if batch_weight.numel() != B or not batch_weight.is_cuda:
                raise RuntimeError(f"batch_weight must be a CUDA tensor with {B} elements")
            batch_weight = batch_weight.to(torch.float32).contiguous()
        rc = lib.rma_chamfer_cd_forward_weighted(_ptr(points_x), xs[0], xs[1], xs[2], _ptr(points_y), ys[0], ys[1], ys[2],
                                        B, N, M, _ptr(batch_weight), float(0.5 / B if coef is None else coef),
                                        _ptr(loss), _ptr(d1), _ptr(d2), None, None,
                                        _ptr(gx[0] if gx else None), _ptr(gx[1] if gx else None),
                                        _ptr(gy[0] if gy else None), _ptr(gy[1] if gy else None),
                                        _ptr(ws), wsb, opts, _stream())
        _lib.check(rc, "rma_chamfer_cd_forward_weighted")
    if want_dist:
        return loss, gx, gy, d1, d2
    return loss, gx, gy


def chamfer_cd_backward(gx, gy, scalar: torch.Tensor):
    """(scalar * (own_x + scat_x), scalar * (own_y + scat_y)) in one launch; scalar is read on the device."""
    ref = (gx or gy)[0]
    sc = scalar.to(torch.float32).reshape(1).contiguous()
    ox = torch.empty_like(gx[0]) if gx else None
    oy = torch.empty_like(gy[0]) if gy else None
    with torch.cuda.device(ref.device):
        rc = _lib.load().rma_chamfer_cd_backward(
            _ptr(gx[0] if gx else None), _ptr(gx[1] if gx else None), gx[0].numel() // 3 if gx else 0,
            _ptr(gy[0] if gy else None), _ptr(gy[1] if gy else None), gy[0].numel() // 3 if gy else 0,
            _ptr(sc), _ptr(ox), _ptr(oy), _stream())
        _lib.check(rc, "rma_chamfer_cd_backward")
    return ox, oy


def sqrdis_map(points_x: torch.Tensor, points_y: torch.Tensor) -> torch.Tensor:
    _check_points("points_x", points_x)
    _check_points("points_y", points_y)
    B, N, M = points_x.shape[0], points_x.shape[1], points_y.shape[1]
    if points_y.shape[0] != B:
        raise RuntimeError(f"batch sizes differ: {B} vs {points_y.shape[0]}")
    out = torch.empty((B, N, M), dtype=torch.float32, device=points_x.device)
    if out.numel() == 0:
        return out
    xs, ys = points_x.stride(), points_y.stride()
    with torch.cuda.device(points_x.device):
        rc = _lib.load().rma_sqrdis_map(_ptr(points_x), xs[0], xs[1], xs[2], _ptr(points_y), ys[0], ys[1], ys[2],
                                        B, N, M, _ptr(out), _stream())
        _lib.check(rc, "rma_sqrdis_map")
    return out


def pack_min_keys(dist: torch.Tensor, idx: torch.Tensor, idx_offset: int = 0) -> torch.Tensor:
    """(float32 d >= 0, int32 idx) -> int64 keys whose signed order is (d, idx) order."""
    dist = dist.contiguous()
    idx = idx.contiguous()
    keys = torch.empty(dist.shape, dtype=torch.int64, device=dist.device)
    with torch.cuda.device(dist.device):
        rc = _lib.load().rma_pack_min_keys(_ptr(dist), _ptr(idx), dist.numel(), idx_offset, _ptr(keys), _stream())
        _lib.check(rc, "rma_pack_min_keys")
    return keys


def unpack_min_keys(keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    keys = keys.contiguous()
    dist = torch.empty(keys.shape, dtype=torch.float32, device=keys.device)
    idx = torch.empty(keys.shape, dtype=torch.int32, device=keys.device)
    with torch.cuda.device(keys.device):
        rc = _lib.load().rma_unpack_min_keys(_ptr(keys), keys.numel(), _ptr(dist), _ptr(idx), _stream())
        _lib.check(rc, "rma_unpack_min_keys")
    return dist, idx


# ------------------------------------------------------------------------------------------------------------
# se(3)
# ------------------------------------------------------------------------------------------------------------
def _check_f32_cuda(name: str, t: torch.Tensor, last: int) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (rma_net_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() < 1 or t.shape[-1] != last:
        raise RuntimeError(f"{name} must have last dimension {last}, got {tuple(t.shape)}")


def se3_exp_forward(x: torch.Tensor) -> torch.Tensor:
    """[*,6] -> [*,4,4]"""
    _check_f32_cuda("x", x, 6)
    xc = x.contiguous().view(-1, 6)
    n = xc.shape[0]
    g = torch.empty((n, 4, 4), dtype=torch.float32, device=x.device)
    if n:
        with torch.cuda.device(x.device):
            rc = _lib.load().rma_se3_exp_forward(_ptr(xc), n, _ptr(g), _stream())
            _lib.check(rc, "rma_se3_exp_forward")
    return g.view(*x.shape[:-1], 4, 4)


def se3_exp_backward(x: torch.Tensor, grad_g: torch.Tensor) -> torch.Tensor:
    _check_f32_cuda("x", x, 6)
    xc = x.contiguous().view(-1, 6)
    n = xc.shape[0]
    go = grad_g.to(torch.float32).contiguous().view(-1, 4, 4)
    if go.shape[0] != n:
        raise RuntimeError("grad_output does not match x")
    gx = torch.empty((n, 6), dtype=torch.float32, device=x.device)
    if n:
        with torch.cuda.device(x.device):
            rc = _lib.load().rma_se3_exp_backward(_ptr(xc), _ptr(go), n, _ptr(gx), _stream())
            _lib.check(rc, "rma_se3_exp_backward")
    return gx.view(x.shape)


def _collapse(shape, strides):
    """Stride of the flattened index over dims (shape, strides), or None if they do not collapse to one stride."""
    dims = [(d, s) for d, s in zip(shape, strides) if d != 1]
    if not dims:
        return 0
    for (d0, s0), (d1, s1) in zip(dims[:-1], dims[1:]):
        if s0 != s1 * d1:
            return None
    return dims[-1][1]


class TransformPlan:
    """How a `transform(g, a)` call maps onto G transforms x P points (see rma_se3_transform_forward)."""

    def __init__(self, g: torch.Tensor, a: torch.Tensor):
        _check_f32_cuda("g", g, 4)
        if g.dim() < 2 or g.shape[-2] != 4:
            raise RuntimeError(f"g must have shape [*,4,4], got {tuple(g.shape)}")
        if not a.is_cuda or a.dtype != torch.float32:
            raise RuntimeError("a must be a float32 CUDA tensor (rma_net_b200 has no CPU path)")
        self.same_rank = g.dim() == a.dim()
        if self.same_rank:                       # a: [*,3,N]   (se3.py:108-109)
            if a.dim() < 2 or a.shape[-2] != 3:
                raise RuntimeError(f"a must have shape [*,3,N], got {tuple(a.shape)}")
            lead = torch.broadcast_shapes(g.shape[:-2], a.shape[:-2])
            npts = a.shape[-1]
            self.out_shape = tuple(lead) + (3, npts)
            ge = g.expand(*lead, 4, 4)
            ae = a.expand(*lead, 3, npts)
            self.G = 1
            for d in lead:
                self.G *= d
            self.P = npts
            st = _collapse(lead, ae.stride()[:-2])
            if st is None:
                ae = ae.contiguous()
                st = 3 * npts
            self.a = ae
            self.a_strides = (st, ae.stride(-1), ae.stride(-2))      # (s_t, s_k, s_c)
            self.g_flat = ge.reshape(-1, 4, 4).contiguous()
            self.o_strides = (3 * npts, 1, npts)
            self.lead = tuple(lead)
            self.split = len(lead)
        else:                                    # a: [*,3]     (se3.py:110-111)
            if a.shape[-1] != 3:
                raise RuntimeError(f"a must have shape [*,3], got {tuple(a.shape)}")
            lead = torch.broadcast_shapes(g.shape[:-2], a.shape[:-1])
            self.out_shape = tuple(lead) + (3,)
            ge = g.expand(*lead, 4, 4)
            ae = a.expand(*lead, 3)
            # trailing leading-dims over which g is constant -> points per transform
            gs = ge.stride()[:-2]
            split = len(lead)
            while split > 0 and (gs[split - 1] == 0 or lead[split - 1] == 1):
                split -= 1
            self.G = 1
            for d in lead[:split]:
                self.G *= d
            self.P = 1
            for d in lead[split:]:
                self.P *= d
            st = _collapse(lead[:split], ae.stride()[:split])
            sk = _collapse(lead[split:], ae.stride()[split:-1])
            if st is None or sk is None:
                ae = ae.contiguous()
                st, sk = 3 * self.P, 3
            self.a = ae
            self.a_strides = (st, sk, ae.stride(-1))
            idx = (slice(None),) * split + (0,) * (len(lead) - split)
            self.g_flat = ge[idx].reshape(-1, 4, 4).contiguous()
            self.o_strides = (3 * self.P, 3, 1)
            self.lead = tuple(lead)
            self.split = split


def se3_transform_forward(g: torch.Tensor, a: torch.Tensor):
    plan = TransformPlan(g, a)
    out = torch.empty(plan.out_shape, dtype=torch.float32, device=a.device)
    if out.numel():
        with torch.cuda.device(a.device):
            rc = _lib.load().rma_se3_transform_forward(
                _ptr(plan.g_flat), _ptr(plan.a), *plan.a_strides, plan.G, plan.P,
                _ptr(out), *plan.o_strides, _stream())
            _lib.check(rc, "rma_se3_transform_forward")
    return out, plan


def se3_transform_backward(plan: TransformPlan, grad_out: torch.Tensor, g_shape, a_shape,
                           need_g: bool, need_a: bool):
    go = grad_out.to(torch.float32).contiguous()
    dev = go.device
    # zeros, not empty: with no points (go.numel() == 0) nothing is launched and the sums over points are 0
    grad_g = torch.zeros((plan.G, 4, 4), dtype=torch.float32, device=dev) if need_g else None
    grad_a = torch.zeros(plan.out_shape, dtype=torch.float32, device=dev) if need_a else None
    if go.numel():
        scratch = torch.empty((plan.G, 12), dtype=torch.float64, device=dev) if need_g else None
        with torch.cuda.device(dev):
            rc = _lib.load().rma_se3_transform_backward(
                _ptr(plan.g_flat), _ptr(plan.a), *plan.a_strides, _ptr(go), *plan.o_strides, plan.G, plan.P,
                _ptr(grad_g), _ptr(grad_a), *plan.o_strides, _ptr(scratch), _stream())
            _lib.check(rc, "rma_se3_transform_backward")
    if grad_g is not None:
        lead_g = plan.lead[:plan.split] + (1,) * (len(plan.lead) - plan.split)
        grad_g = grad_g.view(*lead_g, 4, 4).sum_to_size(g_shape) if tuple(lead_g) + (4, 4) != tuple(g_shape) \
            else grad_g.view(g_shape)
    if grad_a is not None and tuple(grad_a.shape) != tuple(a_shape):
        grad_a = grad_a.sum_to_size(a_shape)
    return grad_g, grad_a


# ------------------------------------------------------------------------------------------------------------
# fused recurrent-loop warp
# ------------------------------------------------------------------------------------------------------------
def warp_forward(phi, src, w=None, prev=None, want_rigid: bool = True, want_g: bool = True):
    """phi [B,6]; src [B,N,3] (any strides); w [B,N] or [B,1,N] or None; prev [B,N,3] or None.
    Returns (out [B,N,3], rigid [B,N,3] or None, g [B,4,4] or None)."""
    _check_f32_cuda("phi", phi, 6)
    _check_points("src", src)
    B, N = src.shape[0], src.shape[1]
    phic = phi.contiguous().view(-1, 6)
    if phic.shape[0] != B:
        raise RuntimeError(f"phi must be [B,6] with B={B}, got {tuple(phi.shape)}")
    wc = None
    if w is not None:
        if prev is None:
            raise RuntimeError("prev is required when w is given")
        _check_points("prev", prev)
        wc = w.to(torch.float32).contiguous().view(B, N)
    dev = src.device
    out = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
    blend = wc is not None
    rigid = torch.empty((B, N, 3), dtype=torch.float32, device=dev) if (want_rigid and blend) else None
    g = torch.empty((B, 4, 4), dtype=torch.float32, device=dev) if want_g else None
    ss = src.stride()
    ps = prev.stride() if blend else (0, 0, 0)
    if B and N:
        with torch.cuda.device(dev):
            rc = _lib.load().rma_warp_forward(_ptr(phic), _ptr(src), ss[0], ss[1], ss[2], _ptr(wc),
                                              _ptr(prev if blend else None), ps[0], ps[1], ps[2], B, N,
                                              _ptr(out), 3 * N, 3, 1, _ptr(g), _ptr(rigid), _stream())
            _lib.check(rc, "rma_warp_forward")
    if want_rigid and not blend:
        rigid = out.detach()          # stage 0 / rigid net: the rigid points ARE the deformation (network.py:570)
    return out, rigid, g


def warp_backward(phi, src, w, prev, grad_out, need_w: bool, need_src: bool, grad_rigid=None, grad_g=None):
    """grad_out / grad_rigid [B,N,3] / grad_g [B,4,4] = gradients w.r.t. the three outputs of warp_forward (any may be
    None).  Returns (grad_phi [B,6], grad_w [B,N] or None, grad_src [B,N,3] or None)."""
    B, N = src.shape[0], src.shape[1]
    dev = src.device
    phic = phi.contiguous().view(-1, 6)
    blend = w is not None
    wc = w.to(torch.float32).contiguous().view(B, N) if blend else None
    go = grad_out.to(torch.float32).contiguous() if grad_out is not None else None
    gr = grad_rigid.to(torch.float32).contiguous() if grad_rigid is not None else None
    gg = grad_g.to(torch.float32).contiguous().view(B, 4, 4) if grad_g is not None else None
    grad_phi = torch.empty((B, 6), dtype=torch.float32, device=dev)
    grad_w = torch.empty((B, N), dtype=torch.float32, device=dev) if (need_w and blend) else None
    grad_src = torch.empty((B, N, 3), dtype=torch.float32, device=dev) if need_src else None
    scratch = torch.empty((B, 12), dtype=torch.float64, device=dev)
    ss = src.stride()
    ps = prev.stride() if blend else (0, 0, 0)
    with torch.cuda.device(dev):
        rc = _lib.load().rma_warp_backward(_ptr(phic), _ptr(src), ss[0], ss[1], ss[2], _ptr(wc),
                                           _ptr(prev if blend else None), ps[0], ps[1], ps[2],
                                           _ptr(go), _ptr(gr), 3 * N, 3, 1, _ptr(gg), B, N, _ptr(grad_phi),
                                           _ptr(grad_w), _ptr(grad_src), 3 * N, 3, 1, _ptr(scratch), _stream())
        _lib.check(rc, "rma_warp_backward")
    return grad_phi.view(phi.shape), grad_w, grad_src


def device_info():
    sm, khz, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = _lib.load().rma_b200_device_info(ctypes.byref(sm), ctypes.byref(khz), ctypes.byref(maj), ctypes.byref(mnr))
    _lib.check(rc, "rma_b200_device_info")
    return dict(sm_count=sm.value, clock_khz=khz.value, cc=(maj.value, mnr.value))


# ------------------------------------------------------------------------------------------------------------
# point_render / point_masker (the reference's model/exfunc extension)
# ------------------------------------------------------------------------------------------------------------
def _check_proj(name: str, t: torch.Tensor) -> None:
    """CHECK_CUDA + CHECK_CONTIGUOUS of the reference wrapper (point_render.cpp:21-27), raising RuntimeError."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")


def point_render_forward(proj_points: torch.Tensor, image_height: int, image_width: int, visible_patch_width: int,
                         color_patch_width: int, epsilon: float, threshold: float, want_weight_points: bool = False):
    """Same call as the reference's `point_render.forward` (point_render.cpp:40-99).  Returns
    [pixel_index [B,2N] i32, is_visible [B,N] i32, weight_points [B,N,cpw^2] or None, depth_image [B,H,W,1],
    weight_image [B,H,W,1]]."""
    _check_proj("proj_points", proj_points)
    if proj_points.dim() != 3 or proj_points.shape[-1] != 3:
        raise RuntimeError(f"proj_points must have shape [B, N, 3], got {tuple(proj_points.shape)}")
    B, N = proj_points.shape[0], proj_points.shape[1]
    H, W = int(image_height), int(image_width)
    dev = proj_points.device
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    pixel_index = torch.empty((B, 2 * N), **i32)
    is_visible = torch.empty((B, N), **i32)
    wp = torch.empty((B, N, color_patch_width * color_patch_width), **f32) if want_weight_points else None
    depth = torch.empty((B, H, W, 1), **f32)
    weight = torch.empty((B, H, W, 1), **f32)
    front_back = torch.empty((2, B, H, W), **f32)
    if B * N:
        with torch.cuda.device(dev):
            rc = _lib.load().rma_point_render_forward(
                _ptr(proj_points), B, N, H, W, int(visible_patch_width), int(color_patch_width), float(epsilon),
                float(threshold), _ptr(pixel_index), _ptr(is_visible), _ptr(wp), _ptr(depth), _ptr(weight),
                _ptr(front_back[0]), _ptr(front_back[1]), _stream())
            _lib.check(rc, "rma_point_render_forward")
    else:
        depth.zero_()
        weight.fill_(float(epsilon))
    return [pixel_index, is_visible, wp, depth, weight]


def point_render_backward(grad_depth_image, grad_weight_image, proj_points, pixel_index, is_visible,
                          color_patch_width: int):
    """The reference's `point_render.backward` (point_render.cpp:101-133) -> [grad_proj_points]."""
    _check_proj("proj_points", proj_points)
    B, N = proj_points.shape[0], proj_points.shape[1]
    H, W = grad_depth_image.shape[1], grad_depth_image.shape[2]
    gd = grad_depth_image.to(torch.float32).contiguous()
    gw = grad_weight_image.to(torch.float32).contiguous()
    out = torch.empty_like(proj_points)
    if B * N:
        with torch.cuda.device(proj_points.device):
            rc = _lib.load().rma_point_render_backward(_ptr(gd), _ptr(gw), _ptr(proj_points), _ptr(pixel_index),
                                                       _ptr(is_visible), B, N, H, W, int(color_patch_width),
                                                       _ptr(out), _stream())
            _lib.check(rc, "rma_point_render_backward")
    return [out]


def point_masker_forward(proj_points: torch.Tensor, image_height: int, image_width: int, mask_patch_width: int):
    """The reference's `point_masker.forward` (point_masker.cpp:35-66) -> [mask_image [B,H,W,1]]."""
    _check_proj("proj_points", proj_points)
    if proj_points.dim() != 3 or proj_points.shape[-1] != 3:
        raise RuntimeError(f"proj_points must have shape [B, N, 3], got {tuple(proj_points.shape)}")
    B, N = proj_points.shape[0], proj_points.shape[1]
    H, W = int(image_height), int(image_width)
    mask = torch.empty((B, H, W, 1), dtype=torch.float32, device=proj_points.device)
    if B * N:
        with torch.cuda.device(proj_points.device):
            rc = _lib.load().rma_point_masker_forward(_ptr(proj_points), B, N, H, W, int(mask_patch_width), _ptr(mask),
                                                      _stream())
            _lib.check(rc, "rma_point_masker_forward")
    else:
        mask.zero_()
    return [mask]


def point_masker_backward(grad_mask_image, proj_points, epsilon: float, threshold: float):
    """The reference's `point_masker.backward` (point_masker.cpp:68-97) -> [grad_proj_points]."""
    _check_proj("proj_points", proj_points)
    B, N = proj_points.shape[0], proj_points.shape[1]
    H, W = grad_mask_image.shape[1], grad_mask_image.shape[2]
    g = grad_mask_image.to(torch.float32).contiguous()
    out = torch.empty_like(proj_points)
    if B * N == 0:
        return [out]
    lib = _lib.load()
    with torch.cuda.device(proj_points.device):
        nbytes = lib.rma_point_masker_backward_scratch_bytes(B, N, H, W)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=proj_points.device)
        rc = lib.rma_point_masker_backward(_ptr(g), _ptr(proj_points), B, N, H, W, float(epsilon), float(threshold),
                                           _ptr(out), _ptr(scratch), nbytes, _stream())
        _lib.check(rc, "rma_point_masker_backward")
    return [out]


# ------------------------------------------------------------------------------------------------------------
# brute-force kNN
# ------------------------------------------------------------------------------------------------------------
def knn_indices(queries: torch.Tensor, candidates: torch.Tensor, k: int, skip_first: int = 0,
                want_dist: bool = False):
    """queries [B,N,3], candidates [B,M,3] (any strides) -> idx [B,N,k-skip_first] int32 (and squared distances),
    ascending by (distance, index)."""
    _check_points("queries", queries)
    _check_points("candidates", candidates)
    _same_device(queries, candidates)
    B, N, M = queries.shape[0], queries.shape[1], candidates.shape[1]
    if candidates.shape[0] != B:
        raise RuntimeError(f"batch sizes differ: {B} vs {candidates.shape[0]}")
    k = int(k)
    if not (0 < k <= 32) or not (0 <= skip_first < k):
        raise RuntimeError(f"k must be in 1..32 and skip_first in 0..k-1, got k={k}, skip_first={skip_first}")
    if k > M:
        raise RuntimeError(f"k={k} exceeds the number of candidates {M} (torch.topk raises as well)")
    dev = queries.device
    idx = torch.empty((B, N, k - skip_first), dtype=torch.int32, device=dev)
    dist = torch.empty((B, N, k - skip_first), dtype=torch.float32, device=dev) if want_dist else None
    if B * N:
        qs, cs = queries.stride(), candidates.stride()
        with torch.cuda.device(dev):
            rc = _lib.load().rma_knn(_ptr(queries), qs[0], qs[1], qs[2], _ptr(candidates), cs[0], cs[1], cs[2],
                                     B, N, M, k, int(skip_first), _ptr(idx), _ptr(dist), _stream())
            _lib.check(rc, "rma_knn")
    return (idx, dist) if want_dist else idx


def arap_forward(deformation: torch.Tensor, source: torch.Tensor, neighbour_idx: torch.Tensor,
                 group_weight: Optional[torch.Tensor], coef: float, need_grad: bool):
    """deformation [G,V,3] (G a multiple of B), source [B,V,3] (any strides), neighbour_idx [B,V,n] int32 ->
    (loss 0-dim, d loss / d deformation [G,V,3] or None); see rma_arap_forward in include/rma_b200.h."""
    _check_points("deformation", deformation)
    _check_points("source", source)
    _same_device(deformation, source)
    G, V = deformation.shape[0], deformation.shape[1]
    B = source.shape[0]
    if B == 0 or G % B != 0 or source.shape[1] != V:
        raise RuntimeError(f"deformation {tuple(deformation.shape)} does not stack over source {tuple(source.shape)}")
    if neighbour_idx.dim() != 3 or neighbour_idx.shape[0] != B or neighbour_idx.shape[1] != V:
        raise RuntimeError(f"neighbour_idx must be [B,V,n], got {tuple(neighbour_idx.shape)}")
    if not neighbour_idx.is_cuda or neighbour_idx.device != deformation.device:
        raise RuntimeError("neighbour_idx must be on the same CUDA device as the points")
    idx = neighbour_idx.to(torch.int32).contiguous()
    n = idx.shape[2]
    dev = deformation.device
    if group_weight is not None:
        if group_weight.numel() != G or not group_weight.is_cuda:
            raise RuntimeError(f"group_weight must be a CUDA tensor with {G} elements")
        group_weight = group_weight.to(torch.float32).contiguous()
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    grad = torch.zeros((G, V, 3), dtype=torch.float32, device=dev) if need_grad else None
    if n and V:
        ds, ss = deformation.stride(), source.stride()
        with torch.cuda.device(dev):
            rc = _lib.load().rma_arap_forward(_ptr(deformation), ds[0], ds[1], ds[2], _ptr(source), ss[0], ss[1], ss[2],
                                              _ptr(idx), G, B, V, n, _ptr(group_weight), float(coef), _ptr(loss),
                                              _ptr(grad), _stream())
            _lib.check(rc, "rma_arap_forward")
    return loss, grad


def depth_view_term(ori_depth: torch.Tensor, ori_weight: torch.Tensor, tar_depth: torch.Tensor,
                    tar_weight: torch.Tensor, dis_threshold: float, coef: float, need_grad: bool):
    """Rendered images [*,H,W,1] x 4 -> (loss 0-dim, (d loss / d ori_depth, d loss / d ori_weight) or None);
    see rma_depth_view_term in include/rma_b200.h."""
    imgs = [ori_depth, ori_weight, tar_depth, tar_weight]
    for t in imgs:
        if not t.is_cuda or t.dtype != torch.float32 or t.shape != ori_depth.shape:
            raise RuntimeError("depth_view_term: four float32 CUDA images of one shape expected")
    imgs = [t.contiguous() for t in imgs]
    dev = ori_depth.device
    n = ori_depth.numel()
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    grads = (torch.empty_like(imgs[0]), torch.empty_like(imgs[1])) if need_grad else None
    if n:
        with torch.cuda.device(dev):
            rc = _lib.load().rma_depth_view_term(_ptr(imgs[0]), _ptr(imgs[1]), _ptr(imgs[2]), _ptr(imgs[3]), n,
                                                 float(dis_threshold), float(coef), _ptr(loss),
                                                 _ptr(grads[0] if grads else None), _ptr(grads[1] if grads else None),
                                                 _stream())
            _lib.check(rc, "rma_depth_view_term")
    return loss, grads


def _float3(v):
    import ctypes
    return (ctypes.c_float * 3)(*[float(a) for a in v])


def view_project_forward(points: torch.Tensor, rotations: torch.Tensor, add3, mul3) -> torch.Tensor:
    """points [B,N,3] (any strides), rotations [3V,3] or [V,3,3] -> [V*B, N, 3] = (R_v p + add3) * mul3."""
    _check_points("points", points)
    rot = rotations.to(torch.float32).contiguous()
    if rot.numel() % 9 != 0 or rot.device != points.device:
        raise RuntimeError("rotations must be [V,3,3] / [3V,3] on the points' device")
    V = rot.numel() // 9
    B, N = points.shape[0], points.shape[1]
    out = torch.empty((V * B, N, 3), dtype=torch.float32, device=points.device)
    if B * N:
        st = points.stride()
        a, m = _float3(add3), _float3(mul3)
        with torch.cuda.device(points.device):
            rc = _lib.load().rma_view_project_forward(_ptr(points), st[0], st[1], st[2], _ptr(rot), V, B, N, a, m,
                                                      _ptr(out), _stream())
            _lib.check(rc, "rma_view_project_forward")
    return out


def view_project_backward(grad_out: torch.Tensor, rotations: torch.Tensor, B: int, N: int, mul3) -> torch.Tensor:
    rot = rotations.to(torch.float32).contiguous()
    V = rot.numel() // 9
    g = grad_out.to(torch.float32).contiguous()
    out = torch.empty((B, N, 3), dtype=torch.float32, device=g.device)
    if B * N:
        with torch.cuda.device(g.device):
            rc = _lib.load().rma_view_project_backward(_ptr(g), _ptr(rot), V, B, N, _float3(mul3), _ptr(out), _stream())
            _lib.check(rc, "rma_view_project_backward")
    return out
