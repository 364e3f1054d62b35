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

// visibility_kernel1 (point_render_cuda.cu:30-84): nearest / farthest z over every point whose visible patch covers
// the pixel.  One thread per (point, patch pixel).
__global__ void __launch_bounds__(256) render_visibility_kernel(const float* __restrict__ proj, float* __restrict__ front,
                                                                float* __restrict__ back, long long npoints,
                                                                int N, int H, int W, int vpw) {
  pdl_wait();
  const int vp2 = vpw * vpw;
  const long long total = npoints * vp2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pt = t / vp2;
    const int k = (int)(t - pt * vp2), j = k / vpw, i = k - j * vpw;
    const float* p = proj + pt * 3;
    const int x = pixel_of(p[0], W) - (vpw - 1) / 2 + i, y = pixel_of(p[1], H) - (vpw - 1) / 2 + j;
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const long long o = (pt / N) * (long long)H * W + (long long)y * W + x;
      const float z = p[2];
      atomic_min_float(front + o, z);
      atomic_max_float(back + o, z);
    }
  }
}

// visibility_kernel2 + Rasterize_kernel1 + Rasterize_kernel2 (point_render_cuda.cu:86-227) in one pass.
// One thread per (point, colour patch pixel); the point's visibility test is evaluated by each of its threads
// (two cached loads) instead of in a separate launch.
__global__ void __launch_bounds__(256) render_raster_kernel(const float* __restrict__ proj, const float* __restrict__ front,
                                                            const float* __restrict__ back, int* __restrict__ pixel_index,
                                                            int* __restrict__ is_visible, float* __restrict__ weight_points,
                                                            float* __restrict__ depth, float* __restrict__ weight,
                                                            long long npoints, int N, int H, int W, int cpw,
                                                            float threshold) {
  pdl_wait();
  const int cp2 = cpw * cpw;
  const long long total = npoints * cp2;
  const float distance_divisor = float(((cpw + 1) / 2) * ((cpw + 1) / 2));
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pt = t / cp2;
    const int k = (int)(t - pt * cp2), j = k / cpw, i = k - j * cpw;
    const float* p = proj + pt * 3;
    const float project_x = p[0], project_y = p[1], z = p[2];
    const int id_x = pixel_of(project_x, W), id_y = pixel_of(project_y, H);
    const long long img = (pt / N) * (long long)H * W;
    const long long c = img + (long long)id_y * W + id_x;
    const float f = front[c], bk = back[c];
    // :110-111 — hidden if behind the middle of the local depth range and more than `threshold` behind the front
    // (the reference's `* 0.5` promotes to double; halving is exact in fp32 too, so the comparison is the same)
    const bool visible = !(z > ((f + bk) * 0.5f) && z > (f + threshold));
    if (k == 0) {
      pixel_index[pt * 2] = id_x; pixel_index[pt * 2 + 1] = id_y;
      is_visible[pt] = visible ? 1 : 0;
    }
    const int current_x = id_x - (cpw - 1) / 2 + i, current_y = id_y - (cpw - 1) / 2 + j;
    float g = 0.f;
    if (visible && current_x >= 0 && current_x < W && current_y >= 0 && current_y < H) {
      const float center_x = float(2 * current_x + 1) * 0.5f, center_y = float(2 * current_y + 1) * 0.5f;   // :159-160 (exact)
      const float distance = (project_x - center_x) * (project_x - center_x) +
                             (project_y - center_y) * (project_y - center_y);                            // :161
      g = expf(-distance / distance_divisor);                                                             // :162
      const long long o = img + (long long)current_y * W + current_x;
      atomicAdd(weight + o, g);                                                                           // :164
      atomicAdd(depth + o, g * z);                                                                        // :222
    }
    if (weight_points) weight_points[t] = g;     // zero outside the image / for hidden points (point_render.cpp:86)
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
__global__ void __launch_bounds__(256) mask_forward_kernel(const float* __restrict__ proj, float* __restrict__ mask,
                                                           long long npoints, int N, int H, int W, int mpw) {
  pdl_wait();
  const int mp2 = mpw * mpw;
  const long long total = npoints * mp2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pt = t / mp2;
    const int k = (int)(t - pt * mp2), j = k / mpw, i = k - j * mpw;
    const float* p = proj + pt * 3;
    const int x = pixel_of(p[0], W) - (mpw - 1) / 2 + i, y = pixel_of(p[1], H) - (mpw - 1) / 2 + j;
    if (x >= 0 && x < W && y >= 0 && y < H) mask[(pt / N) * (long long)H * W + (long long)y * W + x] = -1.0f;
  }
}

// mask_backward_kernel (point_masker_cuda.cu:39-88) in three atomic-free stages.
constexpr int kMaskThreads = 256;

// exp(-d / 1000) for the O(pixels x points) loops of the mask backward: one FMUL + MUFU.EX2 instead of an IEEE divide
// and the ~10-instruction expf.  Relative error <= |d/1000| * log2(e) * 2^-23 + 2 ulp: below 1e-6 for every term
// larger than 1e-4 of the largest one - the sums are compared at 1e-5 and the reference accumulates them with atomics
// in arbitrary order.  (The render kernels keep the exact expf: their outputs are compared bit for bit.)
__device__ __forceinline__ float mask_gauss(float distance) {
  return exp2f(distance * (-1.4426950408889634f / 1000.f));   // exp2f lowers to MUFU.EX2 (+ denormal scaling)
}

struct MaskEntry { float cx, cy, g, s; };   // pixel centre, upstream gradient (g / s after stage 2), epsilon + sum_i e_i

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
      e.cx = float(2 * x + 1) * 0.5f; e.cy = float(2 * y + 1) * 0.5f; e.g = g; e.s = 0.f;     // :59-60
      out[before + __popc(bal & ((1u << lane) - 1u))] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int k = 0; k < 8; ++k) tot += warp_tot[k]; base += tot; }
    __syncthreads();
  }
}

// Stage 2: one thread per active pixel: s = epsilon + sum_i expf(-d_i / 1000), points in order (:70-78), staged
// through shared memory.
__global__ void __launch_bounds__(kMaskThreads) mask_norm_kernel(const float* __restrict__ proj, int N, int npix,
                                                                 float epsilon, MaskEntry* __restrict__ entries,
                                                                 const int* __restrict__ counts) {
  __shared__ float2 sp[kMaskThreads];
  const int b = blockIdx.y;
  const int count = counts[b];
  if ((int)(blockIdx.x * kMaskThreads) >= count) return;      // block-uniform
  const int e = blockIdx.x * kMaskThreads + threadIdx.x;
  MaskEntry* ent = entries + (long long)b * npix + e;
  const bool on = e < count;
  const float cx = on ? ent->cx : 0.f, cy = on ? ent->cy : 0.f;
  float s = epsilon;
  const float* p = proj + (long long)b * N * 3;
  for (int i0 = 0; i0 < N; i0 += kMaskThreads) {
    __syncthreads();
    if (i0 + (int)threadIdx.x < N) sp[threadIdx.x] = make_float2(p[(i0 + threadIdx.x) * 3], p[(i0 + threadIdx.x) * 3 + 1]);
    __syncthreads();
    const int n = min(kMaskThreads, N - i0);
    for (int i = 0; i < n; ++i) {
      const float proj_x = sp[i].x, proj_y = sp[i].y;
      const float distance = (proj_x - cx) * (proj_x - cx) + (proj_y - cy) * (proj_y - cy);
      s += mask_gauss(distance);
    }
  }
  if (on) { ent->s = s; ent->g = ent->g / s; }     // g / s once per pixel, for the gather
}

// Stage 3: gather.  Thread (point i, slice k) accumulates  sum g/s * e  over the k-th slice of the batch element's
// active pixels, in pixel order; the slices' partial sums are then added in slice order (stage 4), so the result
// is deterministic.  g/s is formed once per pixel (stage 2), which replaces the reference's per-pair division
// `g * e / s` (:85) by one multiply-add; the two differ by an ulp per term.
constexpr int kMaskSlices = 16;

__global__ void __launch_bounds__(kMaskThreads) mask_gather_kernel(const float* __restrict__ proj, int N, int npix,
                                                                   const MaskEntry* __restrict__ entries,
                                                                   const int* __restrict__ counts,
                                                                   float* __restrict__ partial) {
  __shared__ MaskEntry se[kMaskThreads];
  const int b = blockIdx.y, slice = blockIdx.z;
  const int i = blockIdx.x * kMaskThreads + threadIdx.x;
  const bool on = i < N;
  const float* p = proj + ((long long)b * N + (on ? i : 0)) * 3;
  const float proj_x = p[0], proj_y = p[1];
  const int count = counts[b];
  const int per = (count + kMaskSlices - 1) / kMaskSlices;
  const int begin = min(slice * per, count), end = min(begin + per, count);
  const MaskEntry* ent = entries + (long long)b * npix;
  float acc = 0.f;
  for (int e0 = begin; e0 < end; e0 += kMaskThreads) {
    __syncthreads();
    if (e0 + (int)threadIdx.x < end) se[threadIdx.x] = ent[e0 + threadIdx.x];
    __syncthreads();
    const int n = min(kMaskThreads, end - e0);
#pragma unroll 4
    for (int k = 0; k < n; ++k) {
      const MaskEntry m = se[k];
      const float distance = (proj_x - m.cx) * (proj_x - m.cx) + (proj_y - m.cy) * (proj_y - m.cy);
      acc = fmaf(m.g, mask_gauss(distance), acc);        // m.g holds g / s here
    }
  }
  if (on) partial[((long long)b * kMaskSlices + slice) * N + i] = acc;
}

// Stage 4: grad_z[i] = sum over slices, in slice order; x and y get no gradient (:85 touches only 3*i + 2).
__global__ void __launch_bounds__(256) mask_reduce_kernel(const float* __restrict__ partial, int N, long long npoints,
                                                          float* __restrict__ grad_proj) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npoints) return;
  const long long b = t / N;
  const int i = (int)(t - b * N);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kMaskSlices; ++k) acc += partial[(b * kMaskSlices + k) * N + i];
  grad_proj[t * 3] = 0.f; grad_proj[t * 3 + 1] = 0.f; grad_proj[t * 3 + 2] = acc;
}

static unsigned capped_grid(long long n, int block, int cap) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

static bool render_sizes_ok(int B, int N, int H, int W, int pw) {
  if (B <= 0 || N <= 0 || H <= 0 || W <= 0 || pw <= 0 || pw > 255) return false;
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
  const long long npix = (long long)B * H * W, npts = (long long)B * N;
  RMA_CUDA_TRY(launch_pdl(render_init_kernel, capped_grid(npix, 256, 148 * 8), 256u, 0, st, front_image, back_image,
                          depth_image, weight_image, npix, epsilon));
  RMA_CUDA_TRY(launch_pdl(render_visibility_kernel,
                          capped_grid(npts * visible_patch_width * visible_patch_width, 256, 148 * 32), 256u, 0, st,
                          proj_points, front_image, back_image, npts, N, H, W, visible_patch_width));
  RMA_CUDA_TRY(launch_pdl(render_raster_kernel,
                          capped_grid(npts * color_patch_width * color_patch_width, 256, 148 * 32), 256u, 0, st,
                          proj_points, (const float*)front_image, (const float*)back_image, (int*)pixel_index,
                          (int*)is_visible, weight_points, depth_image, weight_image, npts, N, H, W, color_patch_width,
                          threshold));
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
  const long long npts = (long long)B * N;
  RMA_CUDA_TRY(launch_pdl(mask_forward_kernel, capped_grid(npts * mask_patch_width * mask_patch_width, 256, 148 * 32),
                          256u, 0, st, proj_points, mask_image, npts, N, H, W, mask_patch_width));
  return RMA_OK;
}

static size_t up256(size_t v) { return (v + 255) / 256 * 256; }

size_t rma_point_masker_backward_scratch_bytes(int B, int N, int H, int W) {
  if (B <= 0 || N <= 0 || H <= 0 || W <= 0) return 0;
  return up256((size_t)B * H * W * sizeof(MaskEntry)) + up256((size_t)B * (1 + kMaskSegs) * sizeof(int)) +
         up256((size_t)B * kMaskSlices * N * sizeof(float));
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
  // the grids cover the worst case (every pixel active); blocks beyond the batch element's count exit at once
  mask_norm_kernel<<<dim3((npix + kMaskThreads - 1) / kMaskThreads, B), kMaskThreads, 0, st>>>(
      proj_points, N, npix, epsilon, entries, counts);
  RMA_LAUNCH_CHECK();
  mask_gather_kernel<<<dim3((N + kMaskThreads - 1) / kMaskThreads, B, kMaskSlices), kMaskThreads, 0, st>>>(
      proj_points, N, npix, entries, counts, partial);
  RMA_LAUNCH_CHECK();
  mask_reduce_kernel<<<(unsigned)(((long long)B * N + 255) / 256), 256, 0, st>>>(partial, N, (long long)B * N,
                                                                                 grad_proj_points);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

}  // extern "C"
