// As-rigid-as-possible edge-length term of the reference's Loss.forward for B200 (sm_100a) — SURVEY.md §8f rank 3.
//
// Replaces, for all T stages of a training step at once, the torch composition of
//   Loss.loss_on_arap(neighbour_indexes, deformation_p, source_points)      model/loss.py:233-243
//   and its stage loop with the 10000 / 1000 / 100 stage factors             model/loss.py:334-345
// which per stage launches two advanced-indexing gathers to [B,V,n,3], ~12 elementwise / reduction kernels and their
// ~20 autograd kernels.  Here: ONE launch computes
//     loss = coef * sum_g weight[g] * sum_{v,k} ( sqrt(|d_j - d_v|^2 + 1e-5) - sqrt(|s_j - s_v|^2 + 1e-5) )^2,
//     j = idx[b,v,k],  g = stage * B + b
// and, fused, its gradient w.r.t. every deformation point (own part accumulated in registers, neighbour part
// scattered with red.global.add).  HBM-bound gather work: per (g, v) it reads 12 B own + 12 n B neighbours of the
// deformation, the same of the source, 4 n B of indices, and updates 12 (n + 1) B of gradient; the neighbour reads
// hit L2 (the graph is local).  fp32 throughout, sums in the reference's order ((x^2 + y^2) + z^2) + 1e-5.
#include "common.cuh"
#include "../../include/rma_b200.h"

namespace rma {

constexpr int kArapThreads = 256;

struct ArapArgs {
  const float* d; long long dsb, dsn, dsc;     // deformation [G,V,3], element strides
  const float* s; long long ssb, ssn, ssc;     // source      [B,V,3], element strides
  const int* idx;                              // [B,V,n] contiguous; entries outside [0,V) are skipped
  const float* weight;                         // null or [G]
  float coef;
  int G, B, V, n;
  float* loss;                                 // [1], zeroed by the caller; += coef * weight[g] * (block sum)
  float* grad;                                 // null or [G,V,3] contiguous, zeroed by the caller
};

__device__ __forceinline__ float edge_len(float ax, float ay, float az, float bx, float by, float bz,
                                          float& ex, float& ey, float& ez) {
  ex = ax - bx; ey = ay - by; ez = az - bz;     // neighbour - centre, as the reference subtracts (loss.py:237-238)
  return sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)), 0.00001f));
}

__global__ void __launch_bounds__(kArapThreads) arap_kernel(ArapArgs a) {
  __shared__ float swarp[kArapThreads / 32];
  pdl_wait();
  const int g = blockIdx.y;
  const int b = g % a.B;
  const int v = blockIdx.x * kArapThreads + threadIdx.x;
  const float wg = a.coef * (a.weight ? a.weight[g] : 1.f);
  float part = 0.f;
  if (v < a.V) {
    const float* dp = a.d + g * a.dsb;
    const float* sp = a.s + b * a.ssb;
    const float* pd = dp + v * a.dsn;
    const float* ps = sp + v * a.ssn;
    const float dvx = pd[0], dvy = pd[a.dsc], dvz = pd[2 * a.dsc];
    const float svx = ps[0], svy = ps[a.ssc], svz = ps[2 * a.ssc];
    const int* ip = a.idx + ((long long)b * a.V + v) * a.n;
    float* gp = a.grad ? a.grad + (long long)g * a.V * 3 : nullptr;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    for (int k = 0; k < a.n; ++k) {
      const int j = ip[k];
      if ((unsigned)j >= (unsigned)a.V) continue;
      const float* qd = dp + j * a.dsn;
      const float* qs = sp + j * a.ssn;
      float ex, ey, ez, fx, fy, fz;
      const float ld = edge_len(qd[0], qd[a.dsc], qd[2 * a.dsc], dvx, dvy, dvz, ex, ey, ez);
      const float ls = edge_len(qs[0], qs[a.ssc], qs[2 * a.ssc], svx, svy, svz, fx, fy, fz);
      const float diff = ld - ls;
      part = fmaf(diff, diff, part);
      if (gp) {
        // d/d(d_j) of wg * diff^2 = 2 wg diff (d_j - d_v) / ld; the centre gets the opposite
        const float f = 2.f * wg * diff / ld;
        const float gx = f * ex, gy = f * ey, gz = f * ez;
        ox -= gx; oy -= gy; oz -= gz;
        float* t = gp + (long long)j * 3;
        atomicAdd(t, gx); atomicAdd(t + 1, gy); atomicAdd(t + 2, gz);
      }
    }
    if (gp) {
      float* t = gp + (long long)v * 3;
      atomicAdd(t, ox); atomicAdd(t + 1, oy); atomicAdd(t + 2, oz);
    }
  }
  pdl_launch_dependents();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < kArapThreads / 32; ++w) tot += (double)swarp[w];
    atomicAdd(a.loss, (float)(tot * (double)wg));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-view depth comparison of loss_on_depth (model/loss.py:169-181) as one pass over the four rendered images:
//   ori = ori_depth / ori_weight, tar = tar_depth / tar_weight, dis = |ori - tar|,
//   mask = sign(ori * tar) * [dis <= 0.05]  (the reference builds the second factor as
//          (sign(sign(-1.0 * (dis - 0.05)) + 0.5) + 1.0) * 0.5, which is 1 at dis == 0.05 too),
//   loss += coef * sum mask * dis^2,   and the gradients w.r.t. ori_depth / ori_weight (mask is detached there).
// Replaces ~14 elementwise torch kernels forward and as many backward over [views*B, H, W] images: 24 B read + 8 B
// written per pixel instead of ~200 B.  HBM-bound streaming kernel, 4 pixels per thread (LDG.128 / STG.128).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDepthThreads = 256;

__device__ __forceinline__ float sign_of(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float depth_pixel(float od, float ow, float td, float tw, float thr, float coef,
                                             float& g_od, float& g_ow) {
  const float ori = __fdiv_rn(od, ow), tar = __fdiv_rn(td, tw);
  const float m1 = sign_of(__fmul_rn(ori, tar));
  const float diff = ori - tar;
  const float dis = fabsf(diff);
  const float thres = -1.0f * (dis - thr);
  const float m2 = (sign_of(sign_of(thres) + 0.5f) + 1.0f) * 0.5f;
  const float mask = m2 * m1;
  const float g_ori = coef * mask * 2.f * dis * sign_of(diff);
  // d ori / d od = 1 / ow, d ori / d ow = -od / ow^2 = -ori / ow: one reciprocal instead of two more IEEE divisions
  // (the kernel is bound by its divisions, profiles/r01m_exfunc_ncu_summary.md); 2 ulp, compared at 1e-5
  const float inv_ow = __frcp_rn(ow);
  g_od = g_ori * inv_ow;
  g_ow = -(g_ori * ori) * inv_ow;
  return mask * (dis * dis);
}

__global__ void __launch_bounds__(kDepthThreads) depth_view_term_kernel(
    const float* __restrict__ od, const float* __restrict__ ow, const float* __restrict__ td,
    const float* __restrict__ tw, long long n, float thr, float coef, float* __restrict__ loss,
    float* __restrict__ g_od, float* __restrict__ g_ow) {
  __shared__ float swarp[kDepthThreads / 32];
  pdl_wait();
  float part = 0.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * kDepthThreads;
  for (long long t = (long long)blockIdx.x * kDepthThreads + threadIdx.x; t < n4; t += stride) {
    const float4 a = reinterpret_cast<const float4*>(od)[t], b = reinterpret_cast<const float4*>(ow)[t];
    const float4 c = reinterpret_cast<const float4*>(td)[t], d = reinterpret_cast<const float4*>(tw)[t];
    float4 ga, gb;
    part += depth_pixel(a.x, b.x, c.x, d.x, thr, coef, ga.x, gb.x);
    part += depth_pixel(a.y, b.y, c.y, d.y, thr, coef, ga.y, gb.y);
    part += depth_pixel(a.z, b.z, c.z, d.z, thr, coef, ga.z, gb.z);
    part += depth_pixel(a.w, b.w, c.w, d.w, thr, coef, ga.w, gb.w);
    if (g_od) { reinterpret_cast<float4*>(g_od)[t] = ga; reinterpret_cast<float4*>(g_ow)[t] = gb; }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {      // tail pixels
    const long long t = (n4 << 2) + threadIdx.x;
    float ga, gb;
    part += depth_pixel(od[t], ow[t], td[t], tw[t], thr, coef, ga, gb);
    if (g_od) { g_od[t] = ga; g_ow[t] = gb; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < kDepthThreads / 32; ++w) tot += (double)swarp[w];
    atomicAdd(loss, (float)(tot * (double)coef));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-view projection of loss_on_depth / loss_on_mask (model/loss.py:161-166, 187-191): every point is rotated by
// each of the V view matrices and mapped to pixel coordinates,
//     out[v, b, n, r] = (sum_c R[v][r][c] * p[b, n, c] + add[r]) * mul[r],
// with the views folded into the batch dimension ([V*B, N, 3], ready for the renderer).  The reference does this with
// one bmm against the stacked rotations, a transpose, V slices, and 3-5 elementwise kernels per view (add, mul, div,
// cat, contiguous) plus their autograd mirror; here one launch forward and one backward
//     grad_p[b, n, c] = sum_v sum_r R[v][r][c] * mul[r] * g[v, b, n, r].
// `(x + 1) * (res - 1) / 2` of the reference is three roundings, the last one exact: (x + 1) * ((res - 1) / 2) gives
// the same bits.  HBM-bound: 12 B read + 12 V B written per point (and the reverse).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kProjThreads = 256;

__global__ void __launch_bounds__(kProjThreads) view_project_kernel(
    const float* __restrict__ pts, long long sb, long long sn, long long sc, const float* __restrict__ rot, int V, int B,
    int N, float ax, float ay, float az, float mx, float my, float mz, float* __restrict__ out) {
  extern __shared__ float srot[];
  pdl_wait();
  for (int t = threadIdx.x; t < V * 9; t += kProjThreads) srot[t] = rot[t];
  __syncthreads();
  const long long total = (long long)B * N;
  for (long long t = (long long)blockIdx.x * kProjThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kProjThreads) {
    const int b = (int)(t / N), n = (int)(t - (long long)b * N);
    const float* p = pts + b * sb + n * sn;
    const float x = p[0], y = p[sc], z = p[2 * sc];
    for (int v = 0; v < V; ++v) {
      const float* r = srot + v * 9;
      float* o = out + ((long long)v * total + t) * 3;
      o[0] = (fmaf(r[2], z, fmaf(r[1], y, r[0] * x)) + ax) * mx;
      o[1] = (fmaf(r[5], z, fmaf(r[4], y, r[3] * x)) + ay) * my;
      o[2] = (fmaf(r[8], z, fmaf(r[7], y, r[6] * x)) + az) * mz;
    }
  }
}

__global__ void __launch_bounds__(kProjThreads) view_project_backward_kernel(
    const float* __restrict__ gout, const float* __restrict__ rot, int V, long long total, float mx, float my, float mz,
    float* __restrict__ gpts) {
  extern __shared__ float srot[];
  pdl_wait();
  for (int t = threadIdx.x; t < V * 9; t += kProjThreads) {
    const int r = (t % 9) / 3;
    srot[t] = rot[t] * (r == 0 ? mx : (r == 1 ? my : mz));
  }
  __syncthreads();
  for (long long t = (long long)blockIdx.x * kProjThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kProjThreads) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int v = 0; v < V; ++v) {
      const float* r = srot + v * 9;
      const float* g = gout + ((long long)v * total + t) * 3;
      const float g0 = g[0], g1 = g[1], g2 = g[2];
      gx = fmaf(r[0], g0, fmaf(r[3], g1, fmaf(r[6], g2, gx)));
      gy = fmaf(r[1], g0, fmaf(r[4], g1, fmaf(r[7], g2, gy)));
      gz = fmaf(r[2], g0, fmaf(r[5], g1, fmaf(r[8], g2, gz)));
    }
    gpts[t * 3] = gx; gpts[t * 3 + 1] = gy; gpts[t * 3 + 2] = gz;
  }
}

}  // namespace rma

using namespace rma;

extern "C" {

int rma_arap_forward(const float* deformation, int64_t dsb, int64_t dsn, int64_t dsc,
                     const float* source, int64_t ssb, int64_t ssn, int64_t ssc,
                     const int32_t* neighbour_idx, int G, int B, int V, int n,
                     const float* group_weight, float coef, float* loss, float* grad_deformation, void* stream_) {
  if (!deformation || !source || !neighbour_idx || !loss) return RMA_E_BADARG;
  if (G <= 0 || B <= 0 || V <= 0 || n <= 0 || G % B != 0 || G > 65535) return RMA_E_BADARG;
  if ((long long)V > (1LL << 30) || (long long)G * V * 3 >= (1LL << 40)) return RMA_E_BADARG;
  ArapArgs a;
  a.d = deformation; a.dsb = dsb; a.dsn = dsn; a.dsc = dsc;
  a.s = source; a.ssb = ssb; a.ssn = ssn; a.ssc = ssc;
  a.idx = (const int*)neighbour_idx; a.weight = group_weight; a.coef = coef;
  a.G = G; a.B = B; a.V = V; a.n = n; a.loss = loss; a.grad = grad_deformation;
  const dim3 grid((unsigned)((V + kArapThreads - 1) / kArapThreads), (unsigned)G);
  arap_kernel<<<grid, kArapThreads, 0, (cudaStream_t)stream_>>>(a);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_depth_view_term(const float* ori_depth, const float* ori_weight, const float* tar_depth,
                        const float* tar_weight, int64_t n, float dis_threshold, float coef, float* loss,
                        float* grad_ori_depth, float* grad_ori_weight, void* stream_) {
  if (!ori_depth || !ori_weight || !tar_depth || !tar_weight || !loss || n <= 0) return RMA_E_BADARG;
  if ((grad_ori_depth == nullptr) != (grad_ori_weight == nullptr)) return RMA_E_BADARG;
  const uintptr_t al = (uintptr_t)ori_depth | (uintptr_t)ori_weight | (uintptr_t)tar_depth | (uintptr_t)tar_weight |
                       (uintptr_t)grad_ori_depth | (uintptr_t)grad_ori_weight;
  if (al & 15u) return RMA_E_BADARG;                        // LDG.128 / STG.128
  long long blocks = ((n >> 2) + kDepthThreads - 1) / kDepthThreads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  depth_view_term_kernel<<<(unsigned)blocks, kDepthThreads, 0, (cudaStream_t)stream_>>>(
      ori_depth, ori_weight, tar_depth, tar_weight, (long long)n, dis_threshold, coef, loss, grad_ori_depth,
      grad_ori_weight);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

static unsigned proj_grid(long long total) {
  long long g = (total + kProjThreads - 1) / kProjThreads;
  return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

int rma_view_project_forward(const float* points, int64_t sb, int64_t sn, int64_t sc, const float* rotations, int V,
                             int B, int N, const float* add3, const float* mul3, float* out, void* stream_) {
  if (!points || !rotations || !add3 || !mul3 || !out || V <= 0 || B <= 0 || N <= 0 || V > 4096) return RMA_E_BADARG;
  view_project_kernel<<<proj_grid((long long)B * N), kProjThreads, (size_t)V * 9 * sizeof(float),
                        (cudaStream_t)stream_>>>(points, (long long)sb, (long long)sn, (long long)sc, rotations, V, B, N,
                                                 add3[0], add3[1], add3[2], mul3[0], mul3[1], mul3[2], out);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_view_project_backward(const float* grad_out, const float* rotations, int V, int B, int N, const float* mul3,
                              float* grad_points, void* stream_) {
  if (!grad_out || !rotations || !mul3 || !grad_points || V <= 0 || B <= 0 || N <= 0 || V > 4096) return RMA_E_BADARG;
  view_project_backward_kernel<<<proj_grid((long long)B * N), kProjThreads, (size_t)V * 9 * sizeof(float),
                                 (cudaStream_t)stream_>>>(grad_out, rotations, V, (long long)B * N, mul3[0], mul3[1],
                                                          mul3[2], grad_points);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

}  // extern "C"
