// se(3) exponential map, rigid transform and the fused recurrent-loop warp for B200 (sm_100a).
//
// Replaces LieAlgebra.se3.exp / ExpMap / transform (reference model/LieAlgebra/se3.py:51-74, 102-112, 120-154,
// so3.py:14-24, sinc.py:5-17, 91-103, 120-132) and the warp call sites model/network.py:567-570, 652-655, 840-843.
// These are HBM/launch-bound elementwise kernels: 40 B/point for the fused warp.
#include "common.cuh"
#include "../../include/rma_b200.h"

namespace rma {

// sinc1(t) = sin t / t, sinc2(t) = (1 - cos t)/t^2, sinc3(t) = (t - sin t)/t^3.
// Same functions and the same |t| < 0.01 Taylor branch as sinc.py:5-17, 91-103, 120-132.  Away from the Taylor
// branch the reference's fp32 formulas cancel catastrophically for small t (1 - cos t, t - sin t); we evaluate the
// same functions in cancellation-free form so that fp32 results track the fp64 oracle to ~1e-6:
//   sinc2(t) = 0.5 * sinc1(t/2)^2 ;  sinc3(t) for t < 1 by its alternating series in t^2.
__device__ __forceinline__ void sinc123(float t, float& s1, float& s2, float& s3) {
  const float t2 = t * t;
  if (fabsf(t) < 0.01f) {
    s1 = 1.f - t2 / 6.f * (1.f - t2 / 20.f * (1.f - t2 / 42.f));
    s2 = 0.5f * (1.f - t2 / 12.f * (1.f - t2 / 30.f * (1.f - t2 / 56.f)));
    s3 = (1.f / 6.f) * (1.f - t2 / 20.f * (1.f - t2 / 42.f * (1.f - t2 / 72.f)));
    return;
  }
  float sn, cs;
  sincosf(t, &sn, &cs);
  s1 = sn / t;
  const float h = 0.5f * t;
  const float sh = sinf(h) / h;
  s2 = 0.5f * sh * sh;
  if (t < 1.f) {
    // 1/6 - t^2/120 + t^4/5040 - t^6/362880 + t^8/39916800 - t^10/6227020800 + t^12/1307674368000
    float p = 1.f / 1307674368000.f;
    p = fmaf(p, t2, -1.f / 6227020800.f);
    p = fmaf(p, t2, 1.f / 39916800.f);
    p = fmaf(p, t2, -1.f / 362880.f);
    p = fmaf(p, t2, 1.f / 5040.f);
    p = fmaf(p, t2, -1.f / 120.f);
    s3 = fmaf(p, t2, 1.f / 6.f);
  } else {
    s3 = (t - sn) / (t2 * t);
  }
}

// g = exp(x): rows 0..2 of the 4x4 (row-major, 12 floats: R | p).  se3.py:51-74.
__device__ __forceinline__ void se3_exp(const float* __restrict__ x, float (&g)[12]) {
  const float w1 = x[0], w2 = x[1], w3 = x[2], v1 = x[3], v2 = x[4], v3 = x[5];
  const float t = sqrtf(w1 * w1 + w2 * w2 + w3 * w3);
  float s1, s2, s3;
  sinc123(t, s1, s2, s3);
  // W = [w]x ; S = W W = w w^T - t^2 I
  const float S00 = -(w2 * w2 + w3 * w3), S11 = -(w1 * w1 + w3 * w3), S22 = -(w1 * w1 + w2 * w2);
  const float S01 = w1 * w2, S02 = w1 * w3, S12 = w2 * w3;
  // R = I + s1 W + s2 S
  g[0] = 1.f + s2 * S00;        g[1] = -s1 * w3 + s2 * S01;   g[2] = s1 * w2 + s2 * S02;
  g[4] = s1 * w3 + s2 * S01;    g[5] = 1.f + s2 * S11;        g[6] = -s1 * w1 + s2 * S12;
  g[8] = -s1 * w2 + s2 * S02;   g[9] = s1 * w1 + s2 * S12;    g[10] = 1.f + s2 * S22;
  // V = I + s2 W + s3 S ; p = V v
  const float V00 = 1.f + s3 * S00, V01 = -s2 * w3 + s3 * S01, V02 = s2 * w2 + s3 * S02;
  const float V10 = s2 * w3 + s3 * S01, V11 = 1.f + s3 * S11, V12 = -s2 * w1 + s3 * S12;
  const float V20 = -s2 * w2 + s3 * S02, V21 = s2 * w1 + s3 * S12, V22 = 1.f + s3 * S22;
  g[3] = V00 * v1 + V01 * v2 + V02 * v3;
  g[7] = V10 * v1 + V11 * v2 + V12 * v3;
  g[11] = V20 * v1 + V21 * v2 + V22 * v3;
}

// ExpMap.backward closed form (se3.py:133-152): go = rows 0..2 of grad_g (row-major 12 values).  The three cross
// products cancel (grad_w is a difference of sums over N points), so they are formed in double: the result then
// carries only the rounding of its inputs (gradients within 1e-5 of the fp64 oracle at any N).
__device__ __forceinline__ void se3_exp_bwd(const float (&g)[12], const double (&go)[12], float* __restrict__ gx) {
  double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    const double g0 = g[col], g1 = g[4 + col], g2 = g[8 + col];
    const double o0 = go[col], o1 = go[4 + col], o2 = go[8 + col];
    a += g1 * o2 - g2 * o1;
    b += g2 * o0 - g0 * o2;
    c += g0 * o1 - g1 * o0;
  }
  gx[0] = (float)a; gx[1] = (float)b; gx[2] = (float)c;
  gx[3] = (float)go[3]; gx[4] = (float)go[7]; gx[5] = (float)go[11];
}

__global__ void __launch_bounds__(256) se3_exp_fwd_kernel(const float* __restrict__ x, long long n,
                                                          float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float g[12];
  se3_exp(x + t * 6, g);
  float4* o = reinterpret_cast<float4*>(out + t * 16);
  o[0] = make_float4(g[0], g[1], g[2], g[3]);
  o[1] = make_float4(g[4], g[5], g[6], g[7]);
  o[2] = make_float4(g[8], g[9], g[10], g[11]);
  o[3] = make_float4(0.f, 0.f, 0.f, 1.f);
}

__global__ void __launch_bounds__(256) se3_exp_bwd_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ grad_g, long long n,
                                                          float* __restrict__ grad_x) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float g[12];
  double go[12];
  se3_exp(x + t * 6, g);
  const float4* gi = reinterpret_cast<const float4*>(grad_g + t * 16);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float4 v = gi[r];
    go[4 * r] = v.x; go[4 * r + 1] = v.y; go[4 * r + 2] = v.z; go[4 * r + 3] = v.w;
  }
  se3_exp_bwd(g, go, grad_x + t * 6);
}

// ---- transform ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se3_transform_fwd_kernel(
    const float* __restrict__ g, const float* __restrict__ a, long long a_st, long long a_sk, long long a_sc,
    long long G, long long P, float* __restrict__ out, long long o_st, long long o_sk, long long o_sc) {
  const long long total = G * P;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / P, k = idx - t * P;
    const float4* gm = reinterpret_cast<const float4*>(g + t * 16);
    const float4 r0 = gm[0], r1 = gm[1], r2 = gm[2];
    const float* p = a + t * a_st + k * a_sk;
    const float x = p[0], y = p[a_sc], z = p[2 * a_sc];
    float* o = out + t * o_st + k * o_sk;
    o[0] = r0.x * x + r0.y * y + r0.z * z + r0.w;
    o[o_sc] = r1.x * x + r1.y * y + r1.z * z + r1.w;
    o[2 * o_sc] = r2.x * x + r2.y * y + r2.z * z + r2.w;
  }
}

// Sum `v[12]` over the block and add the total into dst[0..11] with (double) atomics.  Everything beyond a thread's
// own few terms is accumulated in double: the sums feed cross products that cancel (se3_exp_bwd), and fp32 partial
// sums over N >= 5000 points cost 3e-5 of the gradient's norm (round-1 tests); the order of the double atomics is
// unspecified but moves the result by ~1e-16 relative.
__device__ __forceinline__ void block_sum12_atomic(double (&v)[12], double* dst, double* smem /* [warps][12] */) {
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) smem[warp * 12 + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double s = 0.0;
    for (unsigned w = 0; w < nwarps; ++w) s += smem[w * 12 + threadIdx.x];
    atomicAdd(dst + threadIdx.x, s);
  }
}

// One point's contribution to the 12 sums of grad_g = sum_k go_k [a_k^T 1]: fp32 products, double accumulation.
__device__ __forceinline__ void acc12(double (&acc)[12], float ox, float oy, float oz, float x, float y, float z) {
  acc[0] += (double)(ox * x); acc[1] += (double)(ox * y); acc[2] += (double)(ox * z); acc[3] += (double)ox;
  acc[4] += (double)(oy * x); acc[5] += (double)(oy * y); acc[6] += (double)(oy * z); acc[7] += (double)oy;
  acc[8] += (double)(oz * x); acc[9] += (double)(oz * y); acc[10] += (double)(oz * z); acc[11] += (double)oz;
}

// Backward, many points per transform: grid = (chunks, G); block-reduce the 12 sums in double, double atomics into
// gsum [G,12] (zeroed by the launcher); sum12_to_grad_g_kernel rounds them once into grad_g.
__global__ void __launch_bounds__(256) se3_transform_bwd_kernel(
    const float* __restrict__ g, const float* __restrict__ a, long long a_st, long long a_sk, long long a_sc,
    const float* __restrict__ go, long long go_st, long long go_sk, long long go_sc, long long P,
    double* __restrict__ gsum, float* __restrict__ grad_a, long long ga_st, long long ga_sk, long long ga_sc) {
  __shared__ double red[8 * 12];
  const long long t = blockIdx.y;
  const float4* gm = reinterpret_cast<const float4*>(g + t * 16);
  const float4 r0 = gm[0], r1 = gm[1], r2 = gm[2];
  double acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P;
       k += (long long)gridDim.x * blockDim.x) {
    const float* q = go + t * go_st + k * go_sk;
    const float ox = q[0], oy = q[go_sc], oz = q[2 * go_sc];
    if (grad_a) {
      float* d = grad_a + t * ga_st + k * ga_sk;
      d[0] = r0.x * ox + r1.x * oy + r2.x * oz;
      d[ga_sc] = r0.y * ox + r1.y * oy + r2.y * oz;
      d[2 * ga_sc] = r0.z * ox + r1.z * oy + r2.z * oz;
    }
    if (gsum) {
      const float* p = a + t * a_st + k * a_sk;
      acc12(acc, ox, oy, oz, p[0], p[a_sc], p[2 * a_sc]);
    }
  }
  if (gsum) block_sum12_atomic(acc, gsum + t * 12, red);
}

__global__ void __launch_bounds__(256) sum12_to_grad_g_kernel(const double* __restrict__ gsum, long long G,
                                                              float* __restrict__ grad_g) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G) return;
  const double* v = gsum + t * 12;
  float4* o = reinterpret_cast<float4*>(grad_g + t * 16);
  o[0] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
  o[1] = make_float4((float)v[4], (float)v[5], (float)v[6], (float)v[7]);
  o[2] = make_float4((float)v[8], (float)v[9], (float)v[10], (float)v[11]);
  o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Backward, few points per transform (per-point transforms, P small): one thread per transform, no atomics.
__global__ void __launch_bounds__(256) se3_transform_bwd_small_kernel(
    const float* __restrict__ g, const float* __restrict__ a, long long a_st, long long a_sk, long long a_sc,
    const float* __restrict__ go, long long go_st, long long go_sk, long long go_sc, long long G, long long P,
    float* __restrict__ grad_g, float* __restrict__ grad_a, long long ga_st, long long ga_sk, long long ga_sc) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G) return;
  const float4* gm = reinterpret_cast<const float4*>(g + t * 16);
  const float4 r0 = gm[0], r1 = gm[1], r2 = gm[2];
  double acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.0;
  for (long long k = 0; k < P; ++k) {
    const float* q = go + t * go_st + k * go_sk;
    const float ox = q[0], oy = q[go_sc], oz = q[2 * go_sc];
    if (grad_a) {
      float* d = grad_a + t * ga_st + k * ga_sk;
      d[0] = r0.x * ox + r1.x * oy + r2.x * oz;
      d[ga_sc] = r0.y * ox + r1.y * oy + r2.y * oz;
      d[2 * ga_sc] = r0.z * ox + r1.z * oy + r2.z * oz;
    }
    const float* p = a + t * a_st + k * a_sk;
    acc12(acc, ox, oy, oz, p[0], p[a_sc], p[2 * a_sc]);
  }
  if (grad_g) {
    float4* o = reinterpret_cast<float4*>(grad_g + t * 16);
    o[0] = make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]);
    o[1] = make_float4((float)acc[4], (float)acc[5], (float)acc[6], (float)acc[7]);
    o[2] = make_float4((float)acc[8], (float)acc[9], (float)acc[10], (float)acc[11]);
    o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- fused warp -----------------------------------------------------------------------------------------------
// grid = (chunks of N, B).  Thread 0 of each block evaluates exp(phi[b]) once into shared memory.
__global__ void __launch_bounds__(256) warp_fwd_kernel(
    const float* __restrict__ phi, const float* __restrict__ src, long long s_sb, long long s_sn, long long s_sc,
    const float* __restrict__ w, const float* __restrict__ prev, long long p_sb, long long p_sn, long long p_sc,
    int N, float* __restrict__ out, long long o_sb, long long o_sn, long long o_sc, float* __restrict__ g_out,
    float* __restrict__ rigid_out) {
  __shared__ float sg[12];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    float g[12];
    se3_exp(phi + (long long)b * 6, g);
#pragma unroll
    for (int i = 0; i < 12; ++i) sg[i] = g[i];
    if (g_out && blockIdx.x == 0) {
      float4* o = reinterpret_cast<float4*>(g_out + (long long)b * 16);
      o[0] = make_float4(g[0], g[1], g[2], g[3]);
      o[1] = make_float4(g[4], g[5], g[6], g[7]);
      o[2] = make_float4(g[8], g[9], g[10], g[11]);
      o[3] = make_float4(0.f, 0.f, 0.f, 1.f);
    }
  }
  __syncthreads();
  float g[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) g[i] = sg[i];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
    const float* p = src + b * s_sb + k * s_sn;
    const float x = p[0], y = p[s_sc], z = p[2 * s_sc];
    float rx = g[0] * x + g[1] * y + g[2] * z + g[3];
    float ry = g[4] * x + g[5] * y + g[6] * z + g[7];
    float rz = g[8] * x + g[9] * y + g[10] * z + g[11];
    if (rigid_out) {
      float* ro = rigid_out + b * o_sb + k * o_sn;
      ro[0] = rx; ro[o_sc] = ry; ro[2 * o_sc] = rz;
    }
    if (w) {
      const float ww = w[(long long)b * N + k];
      const float* q = prev + b * p_sb + k * p_sn;
      // torch.mul(w, rigid) + torch.mul(1 - w, prev)   (network.py:655)
      const float om = 1.f - ww;
      rx = ww * rx + om * q[0];
      ry = ww * ry + om * q[p_sc];
      rz = ww * rz + om * q[2 * p_sc];
    }
    float* o = out + b * o_sb + k * o_sn;
    o[0] = rx; o[o_sc] = ry; o[2 * o_sc] = rz;
  }
}

// `gr` (optional, strided like go) is the gradient that arrives through the returned rigid points
// (deform_rigid_points_list, network.py:654): the rigid points receive w*go + gr.
__global__ void __launch_bounds__(256) warp_bwd_kernel(
    const float* __restrict__ phi, const float* __restrict__ src, long long s_sb, long long s_sn, long long s_sc,
    const float* __restrict__ w, const float* __restrict__ prev, long long p_sb, long long p_sn, long long p_sc,
    const float* __restrict__ go, const float* __restrict__ gr, long long g_sb, long long g_sn, long long g_sc, int N,
    float* __restrict__ grad_w, float* __restrict__ grad_src, long long gs_sb, long long gs_sn, long long gs_sc,
    double* __restrict__ scratch) {
  __shared__ float sg[12];
  __shared__ double red[8 * 12];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    float g[12];
    se3_exp(phi + (long long)b * 6, g);
#pragma unroll
    for (int i = 0; i < 12; ++i) sg[i] = g[i];
  }
  __syncthreads();
  float g[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) g[i] = sg[i];
  double acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
    const float* p = src + b * s_sb + k * s_sn;
    const float x = p[0], y = p[s_sc], z = p[2 * s_sc];
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (go) {
      const float* q = go + b * g_sb + k * g_sn;
      ox = q[0]; oy = q[g_sc]; oz = q[2 * g_sc];
    }
    if (w) {
      const float ww = w[(long long)b * N + k];
      if (grad_w) {
        const float rx = g[0] * x + g[1] * y + g[2] * z + g[3];
        const float ry = g[4] * x + g[5] * y + g[6] * z + g[7];
        const float rz = g[8] * x + g[9] * y + g[10] * z + g[11];
        const float* pv = prev + b * p_sb + k * p_sn;
        grad_w[(long long)b * N + k] = ox * (rx - pv[0]) + oy * (ry - pv[p_sc]) + oz * (rz - pv[2 * p_sc]);
      }
      ox *= ww; oy *= ww; oz *= ww;
    }
    if (gr) {
      const float* q = gr + b * g_sb + k * g_sn;
      ox += q[0]; oy += q[g_sc]; oz += q[2 * g_sc];
    }
    if (grad_src) {
      float* d = grad_src + b * gs_sb + k * gs_sn;
      d[0] = g[0] * ox + g[4] * oy + g[8] * oz;
      d[gs_sc] = g[1] * ox + g[5] * oy + g[9] * oz;
      d[2 * gs_sc] = g[2] * ox + g[6] * oy + g[10] * oz;
    }
    acc12(acc, ox, oy, oz, x, y, z);
  }
  block_sum12_atomic(acc, scratch + (long long)b * 12, red);
}

// ---- contiguous fast paths: 4 points = 48 bytes = 3 x LDG.128 / STG.128 per tensor (the strided kernels above issue
// 3 scalar accesses per point and reach ~70 % / 48 % of the HBM copy bandwidth; these are for [B,N,3] contiguous,
// 16-byte aligned tensors with N % 4 == 0).  Per-point arithmetic is the same expression, so results are identical.
struct Pts4 { float x[4], y[4], z[4]; };
__device__ __forceinline__ Pts4 load_pts4(const float* base) {
  const float4* v = reinterpret_cast<const float4*>(base);
  const float4 a = v[0], b = v[1], c = v[2];
  Pts4 r;
  r.x[0] = a.x; r.y[0] = a.y; r.z[0] = a.z; r.x[1] = a.w; r.y[1] = b.x; r.z[1] = b.y;
  r.x[2] = b.z; r.y[2] = b.w; r.z[2] = c.x; r.x[3] = c.y; r.y[3] = c.z; r.z[3] = c.w;
  return r;
}
__device__ __forceinline__ void store_pts4(float* base, const Pts4& r) {
  float4* v = reinterpret_cast<float4*>(base);
  v[0] = make_float4(r.x[0], r.y[0], r.z[0], r.x[1]);
  v[1] = make_float4(r.y[1], r.z[1], r.x[2], r.y[2]);
  v[2] = make_float4(r.z[2], r.x[3], r.y[3], r.z[3]);
}

__global__ void __launch_bounds__(256) warp_fwd_vec_kernel(
    const float* __restrict__ phi, const float* __restrict__ src, const float* __restrict__ w,
    const float* __restrict__ prev, int N, float* __restrict__ out, float* __restrict__ g_out,
    float* __restrict__ rigid_out) {
  __shared__ float sg[12];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    float g[12];
    se3_exp(phi + (long long)b * 6, g);
#pragma unroll
    for (int i = 0; i < 12; ++i) sg[i] = g[i];
    if (g_out && blockIdx.x == 0) {
      float4* o = reinterpret_cast<float4*>(g_out + (long long)b * 16);
      o[0] = make_float4(g[0], g[1], g[2], g[3]);
      o[1] = make_float4(g[4], g[5], g[6], g[7]);
      o[2] = make_float4(g[8], g[9], g[10], g[11]);
      o[3] = make_float4(0.f, 0.f, 0.f, 1.f);
    }
  }
  __syncthreads();
  float g[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) g[i] = sg[i];
  const long long base = (long long)b * N;
  for (int k4 = blockIdx.x * blockDim.x + threadIdx.x; k4 < N / 4; k4 += gridDim.x * blockDim.x) {
    const long long o3 = (base + 4LL * k4) * 3;
    const Pts4 p = load_pts4(src + o3);
    Pts4 r;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      r.x[e] = g[0] * p.x[e] + g[1] * p.y[e] + g[2] * p.z[e] + g[3];
      r.y[e] = g[4] * p.x[e] + g[5] * p.y[e] + g[6] * p.z[e] + g[7];
      r.z[e] = g[8] * p.x[e] + g[9] * p.y[e] + g[10] * p.z[e] + g[11];
    }
    if (rigid_out) store_pts4(rigid_out + o3, r);
    if (w) {
      const float4 wv = *reinterpret_cast<const float4*>(w + base + 4LL * k4);
      const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
      const Pts4 q = load_pts4(prev + o3);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float om = 1.f - ww[e];
        r.x[e] = ww[e] * r.x[e] + om * q.x[e];
        r.y[e] = ww[e] * r.y[e] + om * q.y[e];
        r.z[e] = ww[e] * r.z[e] + om * q.z[e];
      }
    }
    store_pts4(out + o3, r);
  }
}

__global__ void __launch_bounds__(256) warp_bwd_vec_kernel(
    const float* __restrict__ phi, const float* __restrict__ src, const float* __restrict__ w,
    const float* __restrict__ prev, const float* __restrict__ go, const float* __restrict__ gr, int N,
    float* __restrict__ grad_w, float* __restrict__ grad_src, double* __restrict__ scratch) {
  __shared__ float sg[12];
  __shared__ double red[8 * 12];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    float g[12];
    se3_exp(phi + (long long)b * 6, g);
#pragma unroll
    for (int i = 0; i < 12; ++i) sg[i] = g[i];
  }
  __syncthreads();
  float g[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) g[i] = sg[i];
  double acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.0;
  const long long base = (long long)b * N;
  for (int k4 = blockIdx.x * blockDim.x + threadIdx.x; k4 < N / 4; k4 += gridDim.x * blockDim.x) {
    const long long o3 = (base + 4LL * k4) * 3;
    const Pts4 p = load_pts4(src + o3);
    Pts4 o;
    if (go) {
      o = load_pts4(go + o3);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) { o.x[e] = 0.f; o.y[e] = 0.f; o.z[e] = 0.f; }
    }
    if (w) {
      const float4 wv = *reinterpret_cast<const float4*>(w + base + 4LL * k4);
      const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
      if (grad_w) {
        const Pts4 pv = load_pts4(prev + o3);
        float gw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float rx = g[0] * p.x[e] + g[1] * p.y[e] + g[2] * p.z[e] + g[3];
          const float ry = g[4] * p.x[e] + g[5] * p.y[e] + g[6] * p.z[e] + g[7];
          const float rz = g[8] * p.x[e] + g[9] * p.y[e] + g[10] * p.z[e] + g[11];
          gw[e] = o.x[e] * (rx - pv.x[e]) + o.y[e] * (ry - pv.y[e]) + o.z[e] * (rz - pv.z[e]);
        }
        *reinterpret_cast<float4*>(grad_w + base + 4LL * k4) = make_float4(gw[0], gw[1], gw[2], gw[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { o.x[e] *= ww[e]; o.y[e] *= ww[e]; o.z[e] *= ww[e]; }
    }
    if (gr) {
      const Pts4 r = load_pts4(gr + o3);
#pragma unroll
      for (int e = 0; e < 4; ++e) { o.x[e] += r.x[e]; o.y[e] += r.y[e]; o.z[e] += r.z[e]; }
    }
    if (grad_src) {
      Pts4 d;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        d.x[e] = g[0] * o.x[e] + g[4] * o.y[e] + g[8] * o.z[e];
        d.y[e] = g[1] * o.x[e] + g[5] * o.y[e] + g[9] * o.z[e];
        d.z[e] = g[2] * o.x[e] + g[6] * o.y[e] + g[10] * o.z[e];
      }
      store_pts4(grad_src + o3, d);
    }
    // the thread's 4 points in fp32 (4 terms per sum), everything beyond in double
    float s[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[0] += o.x[e] * p.x[e]; s[1] += o.x[e] * p.y[e]; s[2] += o.x[e] * p.z[e]; s[3] += o.x[e];
      s[4] += o.y[e] * p.x[e]; s[5] += o.y[e] * p.y[e]; s[6] += o.y[e] * p.z[e]; s[7] += o.y[e];
      s[8] += o.z[e] * p.x[e]; s[9] += o.z[e] * p.y[e]; s[10] += o.z[e] * p.z[e]; s[11] += o.z[e];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] += (double)s[i];
  }
  block_sum12_atomic(acc, scratch + (long long)b * 12, red);
}

// true if t is [B,N,3] contiguous from a 16-byte aligned base (or null)
static bool contig_aos(const void* t, long long sb, long long sn, long long sc, int N) {
  return !t || (sc == 1 && sn == 3 && sb == 3LL * N && (((uintptr_t)t) & 15u) == 0);
}

// grad_g rows (scratch [B,12], double) -> grad_phi through ExpMap.backward.  grad_g_in (optional, [B,4,4]) is the
// gradient that arrives through the returned rigid matrix (rigid_matrix_list -> loss_on_tran, loss.py:224-230).
__global__ void warp_bwd_phi_kernel(const float* __restrict__ phi, const double* __restrict__ scratch,
                                    const float* __restrict__ grad_g_in, int B, float* __restrict__ grad_phi) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float g[12];
  double go[12];
  se3_exp(phi + (long long)b * 6, g);
#pragma unroll
  for (int i = 0; i < 12; ++i) go[i] = scratch[(long long)b * 12 + i];
  if (grad_g_in) {
#pragma unroll
    for (int i = 0; i < 12; ++i) go[i] += (double)grad_g_in[(long long)b * 16 + i];
  }
  se3_exp_bwd(g, go, grad_phi + (long long)b * 6);
}

static int grid1d(long long n, int block, long long cap) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace rma

using namespace rma;

extern "C" {

int rma_se3_exp_forward(const float* x, int64_t n, float* g, void* stream) {
  if (!x || !g || n <= 0) return RMA_E_BADARG;
  se3_exp_fwd_kernel<<<grid1d(n, 256, 0x7fffffffLL), 256, 0, (cudaStream_t)stream>>>(x, n, g);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_se3_exp_backward(const float* x, const float* grad_g, int64_t n, float* grad_x, void* stream) {
  if (!x || !grad_g || !grad_x || n <= 0) return RMA_E_BADARG;
  se3_exp_bwd_kernel<<<grid1d(n, 256, 0x7fffffffLL), 256, 0, (cudaStream_t)stream>>>(x, grad_g, n, grad_x);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_se3_transform_forward(const float* g, const float* a, int64_t a_st, int64_t a_sk, int64_t a_sc, int64_t G,
                              int64_t P, float* out, int64_t o_st, int64_t o_sk, int64_t o_sc, void* stream) {
  if (!g || !a || !out || G <= 0 || P <= 0) return RMA_E_BADARG;
  se3_transform_fwd_kernel<<<grid1d(G * P, 256, 148LL * 64), 256, 0, (cudaStream_t)stream>>>(
      g, a, a_st, a_sk, a_sc, G, P, out, o_st, o_sk, o_sc);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_se3_transform_backward(const float* g, const float* a, int64_t a_st, int64_t a_sk, int64_t a_sc,
                               const float* go, int64_t go_st, int64_t go_sk, int64_t go_sc, int64_t G, int64_t P,
                               float* grad_g, float* grad_a, int64_t ga_st, int64_t ga_sk, int64_t ga_sc,
                               double* scratch, void* stream_) {
  if (!g || !a || !go || G <= 0 || P <= 0) return RMA_E_BADARG;
  if (!grad_g && !grad_a) return RMA_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 64 || G > 65535) {
    se3_transform_bwd_small_kernel<<<grid1d(G, 256, 0x7fffffffLL), 256, 0, stream>>>(
        g, a, a_st, a_sk, a_sc, go, go_st, go_sk, go_sc, G, P, grad_g, grad_a, ga_st, ga_sk, ga_sc);
  } else {
    if (grad_g && (!scratch || (((uintptr_t)scratch) & 7u))) return RMA_E_BADARG;
    if (grad_g) RMA_CUDA_TRY(cudaMemsetAsync(scratch, 0, (size_t)G * 12 * sizeof(double), stream));
    long long chunks = (P + 256 * 4 - 1) / (256 * 4);
    if (chunks > 1024) chunks = 1024;
    dim3 grid((unsigned)chunks, (unsigned)G);
    se3_transform_bwd_kernel<<<grid, 256, 0, stream>>>(g, a, a_st, a_sk, a_sc, go, go_st, go_sk, go_sc, P,
                                                       grad_g ? scratch : nullptr, grad_a, ga_st, ga_sk, ga_sc);
    if (grad_g) {
      RMA_LAUNCH_CHECK();
      sum12_to_grad_g_kernel<<<grid1d(G, 256, 0x7fffffffLL), 256, 0, stream>>>(scratch, G, grad_g);
    }
  }
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_warp_forward(const float* phi, const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc, const float* w,
                     const float* prev, int64_t p_sb, int64_t p_sn, int64_t p_sc, int B, int N, float* out,
                     int64_t o_sb, int64_t o_sn, int64_t o_sc, float* g_out, float* rigid_out, void* stream) {
  if (!phi || !src || !out || B <= 0 || N <= 0 || B > 65535) return RMA_E_BADARG;
  if (w && !prev) return RMA_E_BADARG;
  if (N % 4 == 0 && contig_aos(src, s_sb, s_sn, s_sc, N) && contig_aos(out, o_sb, o_sn, o_sc, N) &&
      (!w || (contig_aos(prev, p_sb, p_sn, p_sc, N) && (((uintptr_t)w) & 15u) == 0)) &&
      (((uintptr_t)rigid_out) & 15u) == 0) {
    int vchunks = (N / 4 + 256 * 2 - 1) / (256 * 2);
    if (vchunks > 4096) vchunks = 4096;
    warp_fwd_vec_kernel<<<dim3((unsigned)vchunks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        phi, src, w, prev, N, out, g_out, rigid_out);
    RMA_LAUNCH_CHECK();
    return RMA_OK;
  }
  int chunks = (N + 256 * 2 - 1) / (256 * 2);
  if (chunks > 4096) chunks = 4096;
  dim3 grid((unsigned)chunks, (unsigned)B);
  warp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(phi, src, s_sb, s_sn, s_sc, w, prev, p_sb, p_sn, p_sc, N,
                                                         out, o_sb, o_sn, o_sc, g_out, rigid_out);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_warp_backward(const float* phi, const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc, const float* w,
                      const float* prev, int64_t p_sb, int64_t p_sn, int64_t p_sc, const float* go,
                      const float* go_rigid, int64_t g_sb, int64_t g_sn, int64_t g_sc, const float* grad_g_in, int B,
                      int N, float* grad_phi, float* grad_w, float* grad_src, int64_t gs_sb, int64_t gs_sn,
                      int64_t gs_sc, double* scratch, void* stream_) {
  if (!phi || !src || !grad_phi || !scratch || B <= 0 || N <= 0 || B > 65535) return RMA_E_BADARG;
  if ((((uintptr_t)scratch) & 7u) != 0) return RMA_E_BADARG;
  if (w && !prev) return RMA_E_BADARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  RMA_CUDA_TRY(cudaMemsetAsync(scratch, 0, (size_t)B * 12 * sizeof(double), stream));
  if (N % 4 == 0 && contig_aos(src, s_sb, s_sn, s_sc, N) && contig_aos(go, g_sb, g_sn, g_sc, N) &&
      contig_aos(go_rigid, g_sb, g_sn, g_sc, N) &&
      contig_aos(grad_src, gs_sb, gs_sn, gs_sc, N) &&
      (!w || (contig_aos(prev, p_sb, p_sn, p_sc, N) && (((uintptr_t)w) & 15u) == 0 && (((uintptr_t)grad_w) & 15u) == 0))) {
    int vchunks = (N / 4 + 256 * 2 - 1) / (256 * 2);
    if (vchunks > 2048) vchunks = 2048;
    warp_bwd_vec_kernel<<<dim3((unsigned)vchunks, (unsigned)B), 256, 0, stream>>>(phi, src, w, prev, go, go_rigid, N,
                                                                              grad_w, grad_src, scratch);
  } else {
    int chunks = (N + 256 * 4 - 1) / (256 * 4);
    if (chunks > 1024) chunks = 1024;
    dim3 grid((unsigned)chunks, (unsigned)B);
    warp_bwd_kernel<<<grid, 256, 0, stream>>>(phi, src, s_sb, s_sn, s_sc, w, prev, p_sb, p_sn, p_sc, go, go_rigid,
                                              g_sb, g_sn, g_sc, N, grad_w, grad_src, gs_sb, gs_sn, gs_sc, scratch);
  }
  RMA_LAUNCH_CHECK();
  warp_bwd_phi_kernel<<<(B + 127) / 128, 128, 0, stream>>>(phi, scratch, grad_g_in, B, grad_phi);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

}  // extern "C"
