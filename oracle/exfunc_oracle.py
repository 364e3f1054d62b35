"""CPU restatement (numpy) of the reference's `model/exfunc` extension: point_render and point_masker.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg may import this; the product
path under rma_net_b200/ never does).  Pinned against outputs of the REFERENCE's own CUDA kernels (compiled from
/root/reference by oracle/Makefile into oracle/_ref/libexfunc_ref.so and run on a B200 by oracle/gen_golden_exfunc.py;
fixtures tests/golden/exfunc_*.npz).

Each function cites the reference lines it follows:
  model/exfunc/expackages/point_render/point_render_cuda.cu   (kernels)      point_render.cpp   (allocation / init)
  model/exfunc/expackages/point_masker/point_masker_cuda.cu   (kernels)      point_masker.cpp   (allocation / init)

KNOWN REFERENCE DEFECT (documented, not reproduced): fatomicMin / fatomicMax (point_render_cuda.cu:4-28) assign the
`unsigned int` returned by atomicCAS to a `float` (numeric conversion, not a bit cast), so when the first CAS of a
thread fails under contention the loop exits WITHOUT applying its value.  The reference's front / back images can
therefore miss updates (front too large, back too small — never the other way; seen on every golden case), and a few
`is_visible` flags flip with them, run to run.  This restatement and the B200 kernels implement the intended min / max.

Arithmetic is fp32 where the reference is fp32.  Two places of the reference promote to double through a `2.0`
literal (point_render_cuda.cu:318) — restated as written.  Sums that the reference accumulates with atomicAdd from
many threads (weight_image, depth_image, the mask backward) have no defined order there; they are restated as exact
(float64-accumulated) sums and compared with a tolerance.  Sums a single reference thread accumulates in program
order (the render backward, point_render_cuda.cu:296-331) are restated in that order, in fp32.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def pixel_index(proj_points: np.ndarray, height: int, width: int) -> np.ndarray:
    """point_render_cuda.cu:44-49 — id = max(min(int(p), size-1), 0), C truncation toward zero.  -> [B,N,2] (x, y)"""
    p = np.asarray(proj_points, F)
    ix = np.clip(np.trunc(p[..., 0]).astype(np.int64), 0, width - 1)
    iy = np.clip(np.trunc(p[..., 1]).astype(np.int64), 0, height - 1)
    return np.stack([ix, iy], -1).astype(np.int32)


def _patch_offsets(pw: int):
    half = (pw - 1) // 2
    jj, ii = np.meshgrid(np.arange(pw), np.arange(pw), indexing="ij")      # j = row (y), i = column (x)
    return (ii - half).ravel(), (jj - half).ravel()                         # order j*pw + i, as in the reference loops


def render_forward(proj_points, height, width, visible_patch_width, color_patch_width, epsilon, threshold):
    """point_render.cpp:80-88 (init) + visibility_kernel1/2, Rasterize_kernel1/2 (point_render_cuda.cu:30-187).

    Returns dict(pixel_index [B,N,2] i32, is_visible [B,N] i32, weight_points [B,N,cpw^2] f32,
                 depth_image [B,H,W] f64-accumulated, weight_image [B,H,W] f64-accumulated, front, back)."""
    p = np.asarray(proj_points, F)
    B, N, _ = p.shape
    pix = pixel_index(p, height, width)
    front = np.zeros((B, height, width), F)            # "z-value max is 0"      point_render.cpp:84
    back = np.full((B, height, width), -2.0, F)        # "z-value min is -2"     point_render.cpp:85
    ox, oy = _patch_offsets(visible_patch_width)
    for b in range(B):
        cx = pix[b, :, 0:1] + ox[None]                 # [N, vp^2]
        cy = pix[b, :, 1:2] + oy[None]
        ok = (cx >= 0) & (cx < width) & (cy >= 0) & (cy < height)
        z = np.broadcast_to(p[b, :, 2:3], cx.shape)
        np.minimum.at(front[b], (cy[ok], cx[ok]), z[ok])       # fatomicMin  :78
        np.maximum.at(back[b], (cy[ok], cx[ok]), z[ok])        # fatomicMax  :79
    # visibility_kernel2 :86-117 — fp32 arithmetic as written
    bi = np.arange(B)[:, None]
    f = front[bi, pix[..., 1], pix[..., 0]]
    k = back[bi, pix[..., 1], pix[..., 0]]
    z = p[..., 2]
    hidden = (z > (f + k) * F(0.5)) & (z > f + F(threshold))
    vis = np.where(hidden, 0, 1).astype(np.int32)

    cp2 = color_patch_width * color_patch_width
    wp = np.zeros((B, N, cp2), F)                                # point_render.cpp:86
    depth = np.zeros((B, height, width), np.float64)
    weight = np.full((B, height, width), float(F(epsilon)), np.float64)      # epsilon * ones, fp32 value
    ox, oy = _patch_offsets(color_patch_width)
    divisor = F(((color_patch_width + 1) // 2) ** 2)             # :141
    for b in range(B):
        cx = pix[b, :, 0:1] + ox[None]
        cy = pix[b, :, 1:2] + oy[None]
        ok = (cx >= 0) & (cx < width) & (cy >= 0) & (cy < height) & (vis[b, :, None] == 1)
        ctr_x = (2 * cx + 1).astype(F) / F(2)                    # float(2*cx+1)/2.0  :159 (exact)
        ctr_y = (2 * cy + 1).astype(F) / F(2)
        dx = p[b, :, 0:1] - ctr_x
        dy = p[b, :, 1:2] - ctr_y
        dist = _fma_dist(dx, dy)
        g = np.exp((-dist / divisor).astype(F)).astype(F)        # expf(-distance / divisor)  :162
        wp[b][ok] = g[ok]
        np.add.at(weight[b], (cy[ok], cx[ok]), g[ok].astype(np.float64))                         # :164
        gz = (g * p[b, :, 2:3]).astype(F)                        # gaussian * z in fp32  :222
        np.add.at(depth[b], (cy[ok], cx[ok]), gz[ok].astype(np.float64))
    return dict(pixel_index=pix, is_visible=vis, weight_points=wp, depth_image=depth, weight_image=weight,
                front=front, back=back)


def _fma_dist(dx, dy):
    """(px-cx)*(px-cx) + (py-cy)*(py-cy) as nvcc compiles it by default (-fmad=true): fma(dx, dx, dy*dy) —
    the right-hand product is rounded, the left one is fused into the add."""
    dx64, dy64 = dx.astype(np.float64), dy.astype(np.float64)
    r = (dy * dy).astype(F).astype(np.float64)          # fp32 product of exact fp32 inputs, rounded once
    return (dx64 * dx64 + r).astype(F)                  # exact product + rounded addend, rounded once (53 bits suffice)


def render_backward(grad_depth_image, grad_weight_image, proj_points, pix, vis, weight_points, color_patch_width):
    """Rasterize_backward_kernel1/2 (point_render_cuda.cu:189-333): per point, in the reference thread's program
    order (j outer, i inner), fp32 adds of separately rounded products; grad_xy goes through double (`2.0`).  -> [B,N,3]"""
    p = np.asarray(proj_points, F)
    gd = np.asarray(grad_depth_image, F).reshape(p.shape[0], -1)
    gw = np.asarray(grad_weight_image, F).reshape(p.shape[0], -1)
    B, N, _ = p.shape
    H = W = None
    H, W = np.asarray(grad_depth_image).shape[1:3]
    cpw = color_patch_width
    ox, oy = _patch_offsets(cpw)
    divisor = F(((cpw + 1) // 2) ** 2)
    out = np.zeros((B, N, 3), F)
    for b in range(B):
        cx = pix[b, :, 0:1] + ox[None]
        cy = pix[b, :, 1:2] + oy[None]
        ok = (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H) & (vis[b, :, None] == 1)
        idx = np.where(ok, cy * W + cx, 0)
        gdp = gd[b][idx]
        # kernel1 :262-264: grad_depth*z + grad_weight, contracted by nvcc into ONE fma (verified bit for bit against
        # the reference's output, tests/golden); the atomicAdd into a zero-filled slot leaves it unchanged
        gwp = (gdp.astype(np.float64) * p[b, :, 2:3].astype(np.float64) + gw[b][idx].astype(np.float64)).astype(F)
        wpt = np.asarray(weight_points, F)[b]
        t = (gwp * wpt).astype(F)
        gxy = (-(t.astype(np.float64) * 2.0 / np.float64(divisor))).astype(F)    # :318, through double
        ctr_x = (2 * cx + 1).astype(F) / F(2)
        ctr_y = (2 * cy + 1).astype(F) / F(2)
        tx = (gxy * (p[b, :, 0:1] - ctr_x).astype(F)).astype(F)
        ty = (gxy * (p[b, :, 1:2] - ctr_y).astype(F)).astype(F)
        tz = (gdp * wpt).astype(F)
        acc = np.zeros((N, 3), F)
        for k in range(cpw * cpw):                                                # sequential fp32 adds, program order
            m = ok[:, k]
            acc[m, 0] = (acc[m, 0] + tx[m, k]).astype(F)
            acc[m, 1] = (acc[m, 1] + ty[m, k]).astype(F)
            acc[m, 2] = (acc[m, 2] + tz[m, k]).astype(F)
        out[b] = acc
    return out


def masker_forward(proj_points, height, width, mask_patch_width):
    """mask_kernel (point_masker_cuda.cu:4-37) on zeros (point_masker.cpp:55): -1 inside every point's patch."""
    p = np.asarray(proj_points, F)
    B = p.shape[0]
    pix = pixel_index(p, height, width)
    mask = np.zeros((B, height, width), F)
    ox, oy = _patch_offsets(mask_patch_width)
    for b in range(B):
        cx = pix[b, :, 0:1] + ox[None]
        cy = pix[b, :, 1:2] + oy[None]
        ok = (cx >= 0) & (cx < width) & (cy >= 0) & (cy < height)
        mask[b][cy[ok], cx[ok]] = -1.0
    return mask


def masker_backward(grad_mask_image, proj_points, epsilon, threshold, gamma=1000.0):
    """mask_backward_kernel (point_masker_cuda.cu:39-88): for every pixel with |g| > threshold,
    grad_z[i] += g * e_i / (epsilon + sum_i e_i), e_i = expf(-((px-cx)^2 + (py-cy)^2) / 1000).  Only z gets gradient.
    Terms in fp32 as written, accumulation over pixels in float64 (atomicAdd order is undefined in the reference)."""
    p = np.asarray(proj_points, F)
    g = np.asarray(grad_mask_image, F)
    B, H, W = g.shape[:3]
    g = g.reshape(B, H, W)
    out = np.zeros(p.shape, np.float64)
    for b in range(B):
        ys, xs = np.nonzero(np.abs(g[b]) > F(threshold))
        if ys.size == 0:
            continue
        ctr_x = (2 * xs + 1).astype(F) / F(2)
        ctr_y = (2 * ys + 1).astype(F) / F(2)
        for s in range(0, ys.size, 2048):                                     # bounded temporaries
            cxs, cys, gs = ctr_x[s:s + 2048], ctr_y[s:s + 2048], g[b, ys[s:s + 2048], xs[s:s + 2048]]
            dx = p[b, None, :, 0] - cxs[:, None]
            dy = p[b, None, :, 1] - cys[:, None]
            e = np.exp((-_fma_dist(dx, dy) / F(gamma)).astype(F)).astype(F)   # [pixels, N]
            ssum = np.full(e.shape[0], F(epsilon), F)
            for i in range(e.shape[1]):                                       # sequential fp32 sum over points :70-78
                ssum = (ssum + e[:, i]).astype(F)
            term = ((gs[:, None] * e).astype(F) / ssum[:, None]).astype(F)    # g * e / sum, left to right  :85
            out[b, :, 2] += term.astype(np.float64).sum(0)
    return out
