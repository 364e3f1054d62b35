// point_render / point_masker for B200 (sm_100a): the reference's `model/exfunc` extension,
// re-designed.  Semantics follow model/exfunc/expackages/point_render/point_render_cuda.cu and
// point_masker/point_masker_cuda.cu (cited per kernel); structure does not:
//
//   reference                                                  here
//   ---------------------------------------------------------  --------------------------------------------------------
//   6 torch fills + 4 launches + 3 cudaDeviceSynchronize        3 launches, asynchronous on the caller's stream
//   one thread per point, 49 / 81 serial atomics each           one thread per (point, patch pixel): lanes hit
//                                                               consecutive pixels, so the L2 reductions coalesce
//   float min/max by atomicCAS loops                            native integer atomicMin/atomicMax on the float bits
//   weight_points [B,N,81] written, saved, re-read (42 MB at    never materialised unless asked for: the backward
//   B*N = 131072) + grad_weight_points [B,N,81] zero/fill/read  recomputes the Gaussian (same expf, same bits)
//   backward: 2 launches, per-point atomics to its own slot     1 launch, register accumulation in the same order
//   mask backward: every pixel thread loops over all N points   active pixels compacted in pixel order, points tiled
//   twice and atomically adds into all N slots                  through shared memory, no atomics, deterministic
#include "common.cuh"
#include "../../include/rma_b200.h"

namespace rma {

// ---- float atomic min / max through the integer units (any sign, no NaN; -0 is canonicalised to +0) ------------
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  v += 0.f;
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  v += 0.f;
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// point_render_cuda.cu:46-47 / point_masker_cuda.cu:16-17: id = max(min(int(p), size - 1), 0)
__device__ __forceinline__ int pixel_of(float p, int size) { return max(min((int)p, size - 1), 0); }

// ---------------------------------------------------------------------------------------------------------------
// point_render forward
// ---------------------------------------------------------------------------------------------------------------
// point_render.cpp:84-88: front = 0, back = -2, depth = 0, weight = epsilon
__global__ void __launch_bounds__(256) render_init_kernel(float* __restrict__ front, float* __restrict__ back,
                                                          float* __restrict__ depth, float* __restrict__ weight,
                                                          long long n, float epsilon) {
  pdl_wait();
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    front[t] = 0.f; back[t] = -2.f; depth[t] = 0.f; weight[t] = epsilon;
  }
}

// The three splatting kernels below run one thread per (point, patch pixel): lanes hit consecutive pixels, so the L2
// reductions of a patch row share sectors.  Index arithmetic is 32-bit inside a batch element (blockIdx.y walks the
// batch) and the patch width is a compile-time constant for the widths the reference uses (5, 7, 9; PW = 0 = any):
// a 64-bit `t / patch_area` is an emulated ~100-instruction division, which alone made these kernels twice as slow as
// their reduction traffic (tools/red_v2_microbench.cu).
template <int PW>
__device__ __forceinline__ void patch_split(unsigned t, int pw_runtime, unsigned& pt, int& j, int& i) {
  const unsigned pw = PW ? (unsigned)PW : (unsigned)pw_runtime;
  pt = t / (pw * pw);
  const unsigned k = t - pt * (pw * pw);
  j = (int)(k / pw);
  i = (int)(k - (unsigned)j * pw);
}

// visibility_kernel1 (point_render_cuda.cu:30-84): nearest / farthest z over every point whose visible patch covers
// the pixel.
template <int PW>
__global__ void __launch_bounds__(256) render_visibility_kernel(const float* __restrict__ proj, float* __restrict__ front,
                                                                float* __restrict__ back, int B, int N, int H, int W,
                                                                int vpw_runtime) {
  pdl_wait();
  const int vpw = PW ? PW : vpw_runtime;
  const unsigned total = (unsigned)N * (unsigned)(vpw * vpw);
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const float* pb = proj + (size_t)b * N * 3;
    float* fb = front + (size_t)b * H * W;
    float* bb = back + (size_t)b * H * W;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      unsigned pt; int j, i;
      patch_split<PW>(t, vpw, pt, j, i);
      const float* p = pb + pt * 3u;
      const int x = pixel_of(p[0], W) - (vpw - 1) / 2 + i, y = pixel_of(p[1], H) - (vpw - 1) / 2 + j;
      if (x >= 0 && x < W && y >= 0 && y < H) {
        const int o = y * W + x;
        const float z = p[2];
        atomic_min_float(fb + o, z);
        atomic_max_float(bb + o, z);
      }
    }
  }
}

// visibility_kernel2 + Rasterize_kernel1 + Rasterize_kernel2 (point_render_cuda.cu:86-227) in one pass.
// The point's visibility test is evaluated by each of its threads (two cached loads) instead of in a separate launch.
template <int PW>
__global__ void __launch_bounds__(256) render_raster_kernel(const float* __restrict__ proj, const float* __restrict__ front,
                                                            const float* __restrict__ back, int* __restrict__ pixel_index,
                                                            int* __restrict__ is_visible, float* __restrict__ weight_points,
                                                            float* __restrict__ depth, float* __restrict__ weight,
                                                            int B, int N, int H, int W, int cpw_runtime,
                                                            float threshold) {
  pdl_wait();
  const int cpw = PW ? PW : cpw_runtime;
  const unsigned cp2 = (unsigned)(cpw * cpw);
  const unsigned total = (unsigned)N * cp2;
  const float distance_divisor = float(((cpw + 1) / 2) * ((cpw + 1) / 2));
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const size_t img = (size_t)b * H * W, pts = (size_t)b * N;
    const float* pb = proj + pts * 3;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      unsigned pt; int j, i;
      patch_split<PW>(t, cpw, pt, j, i);
      const float* p = pb + pt * 3u;
      const float project_x = p[0], project_y = p[1], z = p[2];
      const int id_x = pixel_of(project_x, W), id_y = pixel_of(project_y, H);
      const int c = id_y * W + id_x;
      const float f = front[img + c], bk = back[img + c];
      // :110-111 — hidden if behind the middle of the local depth range and more than `threshold` behind the front
      // (the reference's `* 0.5` promotes to double; halving is exact in fp32 too, so the comparison is the same)
      const bool visible = !(z > ((f + bk) * 0.5f) && z > (f + threshold));
      if (i == 0 && j == 0) {
        pixel_index[(pts + pt) * 2] = id_x; pixel_index[(pts + pt) * 2 + 1] = id_y;
        is_visible[pts + pt] = visible ? 1 : 0;
      }
      const int current_x = id_x - (cpw - 1) / 2 + i, current_y = id_y - (cpw - 1) / 2 + j;
      float g = 0.f;
      if (visible && current_x >= 0 && current_x < W && current_y >= 0 && current_y < H) {
        const float center_x = float(2 * current_x + 1) * 0.5f, center_y = float(2 * current_y + 1) * 0.5f;   // :159-160 (exact)
        const float distance = (project_x - center_x) * (project_x - center_x) +
                               (project_y - center_y) * (project_y - center_y);                            // :161
        g = expf(-distance / distance_divisor);                                                             // :162
        const size_t o = img + (size_t)(current_y * W + current_x);
        atomicAdd(weight + o, g);                                                                           // :164
        atomicAdd(depth + o, g * z);                                                                        // :222
      }
      if (weight_points) weight_points[pts * cp2 + t] = g;   // zero outside the image / for hidden points (point_render.cpp:86)
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// point_render backward: Rasterize_backward_kernel1 + 2 (point_render_cuda.cu:229-333) fused.
// One thread per point, patch pixels in the reference thread's order (j outer, i inner) with separately rounded
// products and adds, so the result is the reference's bit for bit; the Gaussian is recomputed (identical expf).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) render_backward_kernel(const float* __restrict__ grad_depth,
                                                              const float* __restrict__ grad_weight,
                                                              const float* __restrict__ proj,
                                                              const int* __restrict__ pixel_index,
                                                              const int* __restrict__ is_visible,
                                                              float* __restrict__ grad_proj, long long npoints,
                                                              int N, int H, int W, int cpw) {
  pdl_wait();
  const float distance_divisor = float(((cpw + 1) / 2) * ((cpw + 1) / 2));
  for (long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x; pt < npoints;
       pt += (long long)gridDim.x * blockDim.x) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (is_visible[pt] != 0) {
      const float* p = proj + pt * 3;
      const float project_x = p[0], project_y = p[1], z = p[2];
      const int id_x = pixel_index[pt * 2], id_y = pixel_index[pt * 2 + 1];
      const long long img = (pt / N) * (long long)H * W;
      const int init_x = id_x - (cpw - 1) / 2, init_y = id_y - (cpw - 1) / 2;
      for (int j = 0; j < cpw; ++j) {
        const int current_y = init_y + j;
        if (current_y < 0 || current_y >= H) continue;
        const float center_y = float(2 * current_y + 1) * 0.5f;
        for (int i = 0; i < cpw; ++i) {
          const int current_x = init_x + i;
          if (current_x < 0 || current_x >= W) continue;
          const long long o = img + (long long)current_y * W + current_x;
          const float center_x = float(2 * current_x + 1) * 0.5f;
          const float distance = (project_x - center_x) * (project_x - center_x) +
                                 (project_y - center_y) * (project_y - center_y);
          const float wgt = expf(-distance / distance_divisor);            // == forward's weight_points entry
          const float gd = grad_depth[o];
          const float grad_depth_z = gd * z;                               // :262
          const float gwp = grad_depth_z + grad_weight[o];                 // :264 (added into a zero-filled slot)
          // :318 is -(gwp * wgt * 2.0 / divisor) through DOUBLE (the 2.0 literal).  With t = fl(gwp*wgt) and an integer
          // divisor d < 2^15, q = 2t/d is never within 2^-40 relative of a float rounding midpoint unless it sits on
          // it, so RN24(RN53(q)) == RN24(q): one IEEE float division gives the same bits as the reference's 39
          // FP64 instructions per patch pixel (checked bit for bit against its output, tests/golden/exfunc_*.npz)
          const float grad_xy = -__fdiv_rn(__fmul_rn(__fmul_rn(gwp, wgt), 2.f), distance_divisor);
          gx = __fadd_rn(gx, __fmul_rn(grad_xy, project_x - center_x));    // :319, :322 (atomicAdd: separate rounding)
          gy = __fadd_rn(gy, __fmul_rn(grad_xy, project_y - center_y));
          gz = __fadd_rn(gz, __fmul_rn(gd, wgt));                          // :321, :324
        }
      }
    }
    grad_proj[pt * 3] = gx; grad_proj[pt * 3 + 1] = gy; grad_proj[pt * 3 + 2] = gz;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// point_masker
// ---------------------------------------------------------------------------------------------------------------
// mask_kernel (point_masker_cuda.cu:4-37): -1 inside every point's patch (the image is zeroed by the caller).
template <int PW>
__global__ void __launch_bounds__(256) mask_forward_kernel(const float* __restrict__ proj, float* __restrict__ mask,
                                                           int B, int N, int H, int W, int mpw_runtime) {
  pdl_wait();
  const int mpw = PW ? PW : mpw_runtime;
  const unsigned total = (unsigned)N * (unsigned)(mpw * mpw);
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const float* pb = proj + (size_t)b * N * 3;
    float* mb = mask + (size_t)b * H * W;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      unsigned pt; int j, i;
      patch_split<PW>(t, mpw, pt, j, i);
      const float* p = pb + pt * 3u;
      const int x = pixel_of(p[0], W) - (mpw - 1) / 2 + i, y = pixel_of(p[1], H) - (mpw - 1) / 2 + j;
      if (x >= 0 && x < W && y >= 0 && y < H) mb[y * W + x] = -1.0f;
    }
  }
}

// mask_backward_kernel (point_masker_cuda.cu:39-88) in three atomic-free stages.
constexpr int kMaskThreads = 256;

// exp(-d / 1000) for the O(pixels x points) loops of the mask backward.  Both kernels are bound by the MUFU unit
// (16 ex2 / clk / SM), so everything around the MUFU.EX2 is kept below its 8 issue cycles per warp: coordinates are
// pre-scaled by k = sqrt(log2(e) / 1000), so that exp(-d / 1000) = ex2(-(dx'^2 + dy'^2)) needs no multiply, and
// ex2.approx.ftz is issued directly (exp2f adds a range test and two multiplies for results below 2^-126, which are
// irrelevant next to a sum >= epsilon = 1e-5).  Relative error of a term: 2 ulp (ex2.approx) + 2^-22 |x'| |dx'|
// (rounding of the scaled coordinates; |x'| <= 9.8 for a 256-pixel image): below 3e-6 for every term larger than
// e^-4 of the largest one.  The sums are compared at 1e-5 and the reference accumulates them with atomics in
// arbitrary order.  (The render kernels keep the exact expf: their outputs are compared bit for bit.)
constexpr float kMaskScale = 0.03798282566f;      // sqrt(1.4426950408889634 / 1000)

__device__ __forceinline__ float ex2_neg(float d) {         // 2^(-d)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-d));
  return r;
}

struct MaskEntry { float cx, cy, g, pad; };   // pixel centre * kMaskScale, upstream gradient

// Stage 1: compaction of the pixels with |g| > threshold, in pixel order (deterministic), in two passes over
// kMaskSegs segments per batch element: a count pass, then every segment writes behind the sum of the counts before it.
constexpr int kMaskSegs = 16;

__global__ void __launch_bounds__(256) mask_count_kernel(const float* __restrict__ grad_mask, int npix,
                                                         float threshold, int* __restrict__ segcount) {
  __shared__ int wsum[8];
  pdl_wait();
  const int b = blockIdx.y, seg = blockIdx.x;
  const int per = (npix + kMaskSegs - 1) / kMaskSegs;
  const int begin = min(seg * per, npix), end = min(begin + per, npix);
  int c = 0;
  for (int pix = begin + threadIdx.x; pix < end; pix += 256)
    c += fabsf(grad_mask[(long long)b * npix + pix]) > threshold ? 1 : 0;                       // :46
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += wsum[k];
    segcount[b * kMaskSegs + seg] = t;
  }
}

__global__ void __launch_bounds__(256) mask_compact_kernel(const float* __restrict__ grad_mask, int H, int W,
                                                           float threshold, const int* __restrict__ segcount,
                                                           MaskEntry* __restrict__ entries, int* __restrict__ counts) {
  __shared__ int warp_tot[8];
  __shared__ int base;
  const int b = blockIdx.y, seg = blockIdx.x, npix = H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (npix + kMaskSegs - 1) / kMaskSegs;
  const int begin = min(seg * per, npix), end = min(begin + per, npix);
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < seg; ++k) t += segcount[b * kMaskSegs + k];
    base = t;
    if (seg == kMaskSegs - 1) counts[b] = t + segcount[b * kMaskSegs + seg];
  }
  __syncthreads();
  MaskEntry* out = entries + (long long)b * npix;
  for (int start = begin; start < end; start += 256) {
    const int pix = start + threadIdx.x;
    float g = 0.f;
    bool on = false;
    if (pix < end) { g = grad_mask[(long long)b * npix + pix]; on = fabsf(g) > threshold; }
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = base;
    for (int k = 0; k < warp; ++k) before += warp_tot[k];
    if (on) {
      const int x = pix % W, y = pix / W;
      MaskEntry e;
      e.cx = float(2 * x + 1) * 0.5f * kMaskScale; e.cy = float(2 * y + 1) * 0.5f * kMaskScale;   // :59-60
      e.g = g; e.pad = 0.f;
      out[before + __popc(bal & ((1u << lane) - 1u))] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int k = 0; k < 8; ++k) tot += warp_tot[k]; base += tot; }
    __syncthreads();
  }
}

// Stage 2: thread = active pixel, split = a contiguous chunk of the points: partial sums of e over the chunk, points
// staged through shared memory; spart[pixel][split].  (Split over the points because a batch element has only a few
// thousand active pixels: pixels alone do not fill the machine evenly.)  Each CTA walks the pixel tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... of its batch element.
constexpr int kMaskMaxSplits = 8;
constexpr int kMaskTileCtas = 8;

__global__ void __launch_bounds__(kMaskThreads) mask_norm_kernel(const float* __restrict__ proj, int N, int npix,
                                                                 int splits, const MaskEntry* __restrict__ entries,
                                                                 const int* __restrict__ counts,
                                                                 float* __restrict__ spart) {
  __shared__ __align__(16) float2 sp[kMaskThreads];   // read back as float4 (two points per LDS.128)
  const int b = blockIdx.y, split = blockIdx.z;
  const int count = counts[b];
  const int per = (N + splits - 1) / splits;
  const int ibeg = min(split * per, N), iend = min(ibeg + per, N);
  const float* p = proj + (long long)b * N * 3;
  for (int tile = blockIdx.x * kMaskThreads; tile < count; tile += gridDim.x * kMaskThreads) {      // block-uniform
    const int e = tile + threadIdx.x;
    const bool on = e < count;
    const MaskEntry* ent = entries + (long long)b * npix + e;
    const float cx = on ? ent->cx : 0.f, cy = on ? ent->cy : 0.f;
    float s0 = 0.f, s1 = 0.f;
    for (int i0 = ibeg; i0 < iend; i0 += kMaskThreads) {
      __syncthreads();
      if (i0 + (int)threadIdx.x < iend)
        sp[threadIdx.x] = make_float2(p[(i0 + threadIdx.x) * 3] * kMaskScale, p[(i0 + threadIdx.x) * 3 + 1] * kMaskScale);
      __syncthreads();
      const int n = min(kMaskThreads, iend - i0);
      int i = 0;
#pragma unroll 4
      for (; i + 1 < n; i += 2) {
        const float4 q = *reinterpret_cast<const float4*>(&sp[i]);
        const float ax = q.x - cx, ay = q.y - cy, bx = q.z - cx, by = q.w - cy;
        s0 += ex2_neg(fmaf(ax, ax, ay * ay));
        s1 += ex2_neg(fmaf(bx, bx, by * by));
      }
      if (i < n) { const float ax = sp[i].x - cx, ay = sp[i].y - cy; s0 += ex2_neg(fmaf(ax, ax, ay * ay)); }
    }
    if (on) spart[((long long)b * npix + e) * kMaskMaxSplits + split] = s0 + s1;
  }
}

// Stage 3: gather.  Thread (point i, slice k) accumulates  sum g/s * e  over the k-th slice of the batch element's
// active pixels, in pixel order; the slices' partial sums are then added in slice order (stage 4), so the result
// is deterministic.  s = epsilon + the split sums of stage 2 (in split order) and g / s are formed once per staged
// pixel, which replaces the reference's per-pair division `g * e / s` (:85) by one multiply-add; the two differ by an
// ulp per term.
constexpr int kMaskSlices = 16;

__global__ void __launch_bounds__(kMaskThreads) mask_gather_kernel(const float* __restrict__ proj, int N, int npix,
                                                                   int splits, int slices, float epsilon,
                                                                   const MaskEntry* __restrict__ entries,
                                                                   const int* __restrict__ counts,
                                                                   const float* __restrict__ spart,
                                                                   float* __restrict__ partial) {
  __shared__ float4 se[kMaskThreads];
  const int b = blockIdx.y, slice = blockIdx.z;
  const int i = blockIdx.x * kMaskThreads + threadIdx.x;
  const bool on = i < N;
  const float* p = proj + ((long long)b * N + (on ? i : 0)) * 3;
  const float proj_x = p[0] * kMaskScale, proj_y = p[1] * kMaskScale;
  const int count = counts[b];
  const int per = (count + slices - 1) / slices;
  const int begin = min(slice * per, count), end = min(begin + per, count);
  const MaskEntry* ent = entries + (long long)b * npix;
  float acc0 = 0.f, acc1 = 0.f;
  for (int e0 = begin; e0 < end; e0 += kMaskThreads) {
    __syncthreads();
    if (e0 + (int)threadIdx.x < end) {
      const MaskEntry m = ent[e0 + threadIdx.x];
      const float* sq = spart + ((long long)b * npix + e0 + threadIdx.x) * kMaskMaxSplits;
      float sum = epsilon;
      for (int k = 0; k < splits; ++k) sum += sq[k];
      se[threadIdx.x] = make_float4(m.cx, m.cy, m.g / sum, 0.f);
    }
    __syncthreads();
    const int n = min(kMaskThreads, end - e0);
    int k = 0;
#pragma unroll 4
    for (; k + 1 < n; k += 2) {
      const float4 m0 = se[k], m1 = se[k + 1];
      const float ax = proj_x - m0.x, ay = proj_y - m0.y, bx = proj_x - m1.x, by = proj_y - m1.y;
      acc0 = fmaf(m0.z, ex2_neg(fmaf(ax, ax, ay * ay)), acc0);
      acc1 = fmaf(m1.z, ex2_neg(fmaf(bx, bx, by * by)), acc1);
    }
    if (k < n) {
      const float4 m0 = se[k];
      const float ax = proj_x - m0.x, ay = proj_y - m0.y;
      acc0 = fmaf(m0.z, ex2_neg(fmaf(ax, ax, ay * ay)), acc0);
    }
  }
  if (on) partial[((long long)b * slices + slice) * N + i] = acc0 + acc1;
}

// Stage 4: grad_z[i] = sum over slices, in slice order; x and y get no gradient (:85 touches only 3*i + 2).
__global__ void __launch_bounds__(256) mask_reduce_kernel(const float* __restrict__ partial, int N, long long npoints,
                                                          int slices, float* __restrict__ grad_proj) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npoints) return;
  const long long b = t / N;
  const int i = (int)(t - b * N);
  float acc = 0.f;
  for (int k = 0; k < slices; ++k) acc += partial[(b * slices + k) * N + i];
  grad_proj[t * 3] = 0.f; grad_proj[t * 3 + 1] = 0.f; grad_proj[t * 3 + 2] = acc;
}

static unsigned capped_grid(long long n, int block, int cap) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

// grid of a splatting kernel: x covers the N * pw^2 (point, patch pixel) threads of one batch element, y the batch,
// about 32 CTAs per SM in total
static dim3 splat_grid(int B, int N, int pw) {
  const long long total = (long long)N * pw * pw;
  const int by = B < 65535 ? B : 65535;
  long long gx = (total + 255) / 256, cap = (148LL * 32 + by - 1) / by;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)by, 1);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_grid(void (*kernel)(KArgs...), dim3 grid, unsigned block, cudaStream_t stream,
                                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// PW-specialised launch of a splatting kernel template (5, 7, 9 are the reference's patch widths, loss.py:34-35)
#define RMA_SPLAT_LAUNCH(KERNEL, pw, grid, st, ...)                                                       \
  ((pw) == 5   ? launch_pdl_grid(KERNEL<5>, grid, 256u, st, __VA_ARGS__)                                  \
   : (pw) == 7 ? launch_pdl_grid(KERNEL<7>, grid, 256u, st, __VA_ARGS__)                                  \
   : (pw) == 9 ? launch_pdl_grid(KERNEL<9>, grid, 256u, st, __VA_ARGS__)                                  \
               : launch_pdl_grid(KERNEL<0>, grid, 256u, st, __VA_ARGS__))

static bool render_sizes_ok(int B, int N, int H, int W, int pw) {
  if (B <= 0 || N <= 0 || H <= 0 || W <= 0 || pw <= 0 || pw > 255) return false;
  if ((long long)N * pw * pw >= (1LL << 31) - (1LL << 24)) return false;   // 32-bit thread index inside a batch element
  if ((long long)B * N >= (1LL << 31) / 3) return false;       // the reference indexes index*3 with int (cuda.cu:38)
  if ((long long)B * H * W >= (1LL << 31)) return false;
  return true;
}

}  // namespace rma

using namespace rma;

extern "C" {

int rma_point_render_forward(const float* proj_points, int B, int N, int H, int W, int visible_patch_width,
                             int color_patch_width, float epsilon, float threshold, int32_t* pixel_index,
                             int32_t* is_visible, float* weight_points, float* depth_image, float* weight_image,
                             float* front_image, float* back_image, void* stream_) {
  if (!proj_points || !pixel_index || !is_visible || !depth_image || !weight_image || !front_image || !back_image)
    return RMA_E_BADARG;
  if (!render_sizes_ok(B, N, H, W, visible_patch_width) || !render_sizes_ok(B, N, H, W, color_patch_width))
    return RMA_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  const long long npix = (long long)B * H * W;
  RMA_CUDA_TRY(launch_pdl(render_init_kernel, capped_grid(npix, 256, 148 * 8), 256u, 0, st, front_image, back_image,
                          depth_image, weight_image, npix, epsilon));
  RMA_CUDA_TRY(RMA_SPLAT_LAUNCH(render_visibility_kernel, visible_patch_width, splat_grid(B, N, visible_patch_width), st,
                                proj_points, front_image, back_image, B, N, H, W, visible_patch_width));
  RMA_CUDA_TRY(RMA_SPLAT_LAUNCH(render_raster_kernel, color_patch_width, splat_grid(B, N, color_patch_width), st,
                                proj_points, (const float*)front_image, (const float*)back_image, (int*)pixel_index,
                                (int*)is_visible, weight_points, depth_image, weight_image, B, N, H, W,
                                color_patch_width, threshold));
  return RMA_OK;
}

int rma_point_render_backward(const float* grad_depth_image, const float* grad_weight_image, const float* proj_points,
                              const int32_t* pixel_index, const int32_t* is_visible, int B, int N, int H, int W,
                              int color_patch_width, float* grad_proj_points, void* stream_) {
  if (!grad_depth_image || !grad_weight_image || !proj_points || !pixel_index || !is_visible || !grad_proj_points)
    return RMA_E_BADARG;
  if (!render_sizes_ok(B, N, H, W, color_patch_width)) return RMA_E_BADARG;
  const long long npts = (long long)B * N;
  RMA_CUDA_TRY(launch_pdl(render_backward_kernel, capped_grid(npts, 128, 148 * 64), 128u, 0, (cudaStream_t)stream_,
                          grad_depth_image, grad_weight_image, proj_points, (const int*)pixel_index,
                          (const int*)is_visible, grad_proj_points, npts, N, H, W, color_patch_width));
  return RMA_OK;
}

int rma_point_masker_forward(const float* proj_points, int B, int N, int H, int W, int mask_patch_width,
                             float* mask_image, void* stream_) {
  if (!proj_points || !mask_image || !render_sizes_ok(B, N, H, W, mask_patch_width)) return RMA_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  RMA_CUDA_TRY(cudaMemsetAsync(mask_image, 0, (size_t)B * H * W * sizeof(float), st));
  RMA_CUDA_TRY(RMA_SPLAT_LAUNCH(mask_forward_kernel, mask_patch_width, splat_grid(B, N, mask_patch_width), st,
                                proj_points, mask_image, B, N, H, W, mask_patch_width));
  return RMA_OK;
}

static size_t up256(size_t v) { return (v + 255) / 256 * 256; }

size_t rma_point_masker_backward_scratch_bytes(int B, int N, int H, int W) {
  if (B <= 0 || N <= 0 || H <= 0 || W <= 0) return 0;
  return up256((size_t)B * H * W * sizeof(MaskEntry)) + up256((size_t)B * (1 + kMaskSegs) * sizeof(int)) +
         up256((size_t)B * kMaskSlices * N * sizeof(float)) + up256((size_t)B * H * W * kMaskMaxSplits * sizeof(float));
}

int rma_point_masker_backward(const float* grad_mask_image, const float* proj_points, int B, int N, int H, int W,
                              float epsilon, float threshold, float* grad_proj_points, void* scratch,
                              size_t scratch_bytes, void* stream_) {
  if (!grad_mask_image || !proj_points || !grad_proj_points || !scratch) return RMA_E_BADARG;
  if (!render_sizes_ok(B, N, H, W, 1) || B > 65535) return RMA_E_BADARG;
  if (scratch_bytes < rma_point_masker_backward_scratch_bytes(B, N, H, W) || ((uintptr_t)scratch & 255u))
    return RMA_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream_;
  const int npix = H * W;
  MaskEntry* entries = (MaskEntry*)scratch;
  int* counts = (int*)((char*)scratch + up256((size_t)B * npix * sizeof(MaskEntry)));
  int* segcount = counts + B;
  float* partial = (float*)((char*)counts + up256((size_t)B * (1 + kMaskSegs) * sizeof(int)));
  float* spart = (float*)((char*)partial + up256((size_t)B * kMaskSlices * N * sizeof(float)));
  // The number of active pixels is only known on the device.  Size the grids for ~4 waves of 8 CTAs per SM assuming
  // about four 256-pixel tiles per batch element (a silhouette boundary in a 256x256 image); more pixels only add
  // iterations to the per-CTA loops.
  const int target = 4 * 148 * 8;
  int splits = (target + B * 4 - 1) / (B * 4);
  splits = splits < 1 ? 1 : (splits > kMaskMaxSplits ? kMaskMaxSplits : splits);
  if (splits > (N + kMaskThreads - 1) / kMaskThreads) splits = (N + kMaskThreads - 1) / kMaskThreads;
  const int point_ctas = (N + kMaskThreads - 1) / kMaskThreads;
  int slices = (target + B * point_ctas - 1) / (B * point_ctas);
  slices = slices < 1 ? 1 : (slices > kMaskSlices ? kMaskSlices : slices);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kMaskSegs, (unsigned)B, 1); cfg.blockDim = dim3(256, 1, 1); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    RMA_CUDA_TRY(cudaLaunchKernelEx(&cfg, mask_count_kernel, grad_mask_image, npix, threshold, segcount));
  }
  mask_compact_kernel<<<dim3(kMaskSegs, B), 256, 0, st>>>(grad_mask_image, H, W, threshold, segcount, entries, counts);
  RMA_LAUNCH_CHECK();
  const int tiles = (npix + kMaskThreads - 1) / kMaskThreads;
  mask_norm_kernel<<<dim3(tiles < kMaskTileCtas ? tiles : kMaskTileCtas, B, splits), kMaskThreads, 0, st>>>(
      proj_points, N, npix, splits, entries, counts, spart);
  RMA_LAUNCH_CHECK();
  mask_gather_kernel<<<dim3(point_ctas, B, slices), kMaskThreads, 0, st>>>(proj_points, N, npix, splits, slices,
                                                                            epsilon, entries, counts, spart, partial);
  RMA_LAUNCH_CHECK();
  mask_reduce_kernel<<<(unsigned)(((long long)B * N + 255) / 256), 256, 0, st>>>(partial, N, (long long)B * N, slices,
                                                                                 grad_proj_points);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

}  // extern "C"
