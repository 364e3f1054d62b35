"""CPU check of the rounding-error guard behind the filter + exact-resolve chamfer search (csrc/chamfer_filter.cu header).

The filter value f = fma(x0,q0, fma(x1,q1, fma(x2,q2, w))) with q = -2y, w = fl(|y|^2) is emulated in numpy (an fp32 fma
of fp32 inputs is float32(float64 product + float64 addend): the product is exact in 53 bits), compared with the exact
value in float64/longdouble, and the claimed bounds are asserted over millions of random pairs at several scales,
offsets and near-coincident configurations:

    |f - (|y|^2 - 2 x.y)|  <=  6.1 eps (b^2 + a b)                                   e(a,b)
    |g - |x-y|^2|          <=  e(a,b) + 3.01 eps a^2 + 1.01 eps D      (g = fl(f + fl(|x|^2)))
    guard_row(u, m)        >=  2 e(a, beta) + 10.1 eps D,   beta = a + sqrt(D)      what the proof needs for a row
    guard_col(w, m)        >=  2 (e(alpha, b) + 3.01 eps alpha^2 + 1.01 eps D) + 10.1 eps D + 2^-17 D
and, end to end, every exact minimiser of a row is a candidate (f_j <= min f + guard).
guard_row / guard_col are restated from the CUDA source, constant for constant.
"""
import numpy as np
import pytest

F = np.float32
EPS = 2.0 ** -24


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def norm2_rn(p):                       # csrc: fma(z, z, fma(y, y, x*x))
    return fma32(p[..., 2], p[..., 2], fma32(p[..., 1], p[..., 1], (p[..., 0] * p[..., 0]).astype(F)))


def filter_f(x, y):                    # x [..,3], y [..,3] broadcastable, fp32
    q = (F(-2.0) * y).astype(F)
    w = norm2_rn(y)
    return fma32(x[..., 0], q[..., 0], fma32(x[..., 1], q[..., 1], fma32(x[..., 2], q[..., 2], w)))


SQRT_LO = F(1.0 - 3e-7)                # the device uses sqrt.approx (<= 2 ulp): model its worst case from below


def guard_row(u, mt):                  # chamfer_filter.cu guard_row, fp32 arithmetic
    u, mt = u.astype(F), mt.astype(F)
    D0 = np.maximum(mt + u, F(0))
    Dub = F(1.001) * D0 + F(1e-5) * u
    a = np.sqrt(u) * SQRT_LO * F(1.0001)
    beta = a + np.sqrt(Dub) * SQRT_LO * F(1.0001)
    return (F(12.3) * beta * (beta + a) + F(10.3) * Dub) * F(EPS * 1.01) + F(1e-30)


def guard_col(w, mt):                  # chamfer_filter.cu guard_col
    w, mt = w.astype(F), mt.astype(F)
    D0 = np.maximum(mt, F(0))
    Dub = F(1.001) * D0 + F(1e-5) * w
    bn = np.sqrt(w) * SQRT_LO * F(1.0001)
    alpha = bn + np.sqrt(Dub) * SQRT_LO * F(1.0001)
    return (F(12.3) * bn * (bn + alpha) + F(6.3) * alpha * alpha + F(12.5) * Dub) * F(EPS * 1.01) + Dub * F(7.8e-6) + F(1e-30)


def _pairs(rng, n, kind, scale):
    if kind == "cube":
        x = rng.uniform(-1, 1, (n, 3)); y = rng.uniform(-1, 1, (n, 3))
    elif kind == "near":                # y within 1e-3..1e-7 of x: heavy cancellation
        x = rng.uniform(-1, 1, (n, 3)); y = x + rng.normal(size=(n, 3)) * 10.0 ** rng.uniform(-7, -3, (n, 1))
    elif kind == "offset":              # both far from the origin
        c = rng.uniform(-300, 300, (n, 3)); x = c + rng.uniform(-1, 1, (n, 3)); y = c + rng.uniform(-1, 1, (n, 3))
    else:                               # very different norms
        x = rng.uniform(-1, 1, (n, 3)) * 10.0 ** rng.uniform(-3, 3, (n, 1)); y = rng.uniform(-1, 1, (n, 3))
    return (x * scale).astype(F), (y * scale).astype(F)


@pytest.mark.parametrize("kind", ["cube", "near", "offset", "skewed"])
@pytest.mark.parametrize("scale", [1e-4, 1.0, 1e4])
def test_filter_error_bounds(kind, scale):
    rng = np.random.default_rng(hash((kind, scale)) % (2 ** 32))
    x, y = _pairs(rng, 400_000, kind, scale)
    xl, yl = x.astype(np.longdouble), y.astype(np.longdouble)
    a = np.sqrt((xl ** 2).sum(-1)); b = np.sqrt((yl ** 2).sum(-1))
    Fex = (yl ** 2).sum(-1) - 2 * (xl * yl).sum(-1)
    D = ((xl - yl) ** 2).sum(-1)
    f = filter_f(x, y)
    e = 6.1 * EPS * (b * b + a * b)
    assert np.all(np.abs(f.astype(np.longdouble) - Fex) <= e), "filter error exceeds e(a,b)"
    u = norm2_rn(x)
    g = (f + u).astype(F)
    eg = e + 3.01 * EPS * a * a + 1.01 * EPS * D
    assert np.all(np.abs(g.astype(np.longdouble) - D) <= eg), "column filter error exceeds its bound"
    # the guards dominate what the proof needs, evaluated at the true quantities of these pairs
    beta = a + np.sqrt(D)
    need_row = 2 * 6.1 * EPS * (beta * beta + a * beta) + 10.1 * EPS * D
    assert np.all(guard_row(u, f).astype(np.longdouble) >= need_row)
    alpha = b + np.sqrt(D)
    need_col = 2 * (6.1 * EPS * (b * b + alpha * b) + 3.01 * EPS * alpha * alpha + 1.01 * EPS * D) + 10.1 * EPS * D \
        + 2.0 ** -17 * D
    gt = (g.view(np.uint32) & np.uint32(0xFFFFFFE0)).view(F)          # the 5 packed low bits are cleared / overwritten
    assert np.all(guard_col(norm2_rn(y), np.minimum(g, gt)).astype(np.longdouble) >= need_col)


def test_exact_minimisers_are_always_candidates():
    """End to end on whole rows: with jf = argmin f, every j minimising the exact fp32 difference-form distance
    satisfies f_j <= f_jf + guard_row — the property the resolve stage relies on."""
    rng = np.random.default_rng(7)
    for kind, scale in (("cube", 1.0), ("offset", 1.0), ("near", 1e3), ("cube", 1e-5)):
        for _ in range(40):
            n, m = 64, 3000
            x, _ = _pairs(rng, n, kind, scale)
            if kind == "near":
                y = (x[rng.integers(0, n, m)].astype(np.float64) + rng.normal(size=(m, 3)) * 1e-6 * scale).astype(F)
            else:
                _, y = _pairs(rng, m, kind, scale)
            f = filter_f(x[:, None, :], y[None, :, :])                       # [n, m]
            dx = (x[:, None, :] - y[None, :, :]).astype(F)                   # exact path: fl differences ...
            d = fma32(dx[..., 2], dx[..., 2], fma32(dx[..., 1], dx[..., 1], (dx[..., 0] * dx[..., 0]).astype(F)))
            thr = f.min(1) + guard_row(norm2_rn(x), f.min(1))
            exact_min = d == d.min(1, keepdims=True)
            assert np.all(f[exact_min] <= np.broadcast_to(thr[:, None], f.shape)[exact_min])


def row_threshold(u, M1):              # chamfer_filter.cu row_threshold, fp32 arithmetic
    u, M1 = u.astype(F), M1.astype(F)
    return (M1 + (guard_row(u, M1) * F(1.0001) + F(1.2e-6) * np.abs(M1))).astype(F)


def test_tagged_group_minima_keep_every_exact_minimiser():
    """Row direction as the search records it: the minimum of f over each 16-column group carries the group's position
    in its 64-column tile in the two low mantissa bits, M1 = the smallest TAGGED group minimum, thr = row_threshold(u, M1).
    Every exact minimiser j* must (a) sit in a group whose tagged minimum is <= thr (so the resolve's candidate test on
    the recorded values finds the group) and (b) pass the per-column re-filter f_j* <= thr."""
    rng = np.random.default_rng(11)
    for kind, scale in (("cube", 1.0), ("offset", 1.0), ("near", 1e3), ("cube", 1e-5), ("skewed", 1.0), ("cube", 1e4)):
        for _ in range(25):
            n, m = 48, 2048
            x, _ = _pairs(rng, n, kind, scale)
            if kind == "near":
                y = (x[rng.integers(0, n, m)].astype(np.float64) + rng.normal(size=(m, 3)) * 1e-6 * scale).astype(F)
            else:
                _, y = _pairs(rng, m, kind, scale)
            f = filter_f(x[:, None, :], y[None, :, :])                       # [n, m]
            gm = f.reshape(n, m // 16, 16).min(2)                            # group minima
            gid = (np.arange(m // 16) % 4).astype(np.uint32)
            tagged = ((gm.view(np.uint32) & np.uint32(0xFFFFFFFC)) | gid[None, :]).view(F)
            M1 = tagged.min(1)
            thr = row_threshold(norm2_rn(x), M1)
            dx = (x[:, None, :] - y[None, :, :]).astype(F)
            d = fma32(dx[..., 2], dx[..., 2], fma32(dx[..., 1], dx[..., 1], (dx[..., 0] * dx[..., 0]).astype(F)))
            exact_min = d == d.min(1, keepdims=True)
            thr_b = np.broadcast_to(thr[:, None], f.shape)
            assert np.all(f[exact_min] <= thr_b[exact_min])
            tagged_of_col = np.repeat(tagged, 16, axis=1)
            assert np.all(tagged_of_col[exact_min] <= thr_b[exact_min])
            # the tag moves a value by less than 4 ulp
            assert np.all(np.abs(tagged.astype(np.float64) - gm.astype(np.float64)) <= 2.0 ** -21 * np.abs(gm.astype(np.float64)))
