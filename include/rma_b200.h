/* rma_b200.h — C ABI of the B200-native RMA-Net hot path (chamfer NN search + backward, se(3) exp-map warp).
 *
 * This is the drop-in boundary.  The reference has no native entry points for this path: chamfer is
 * pure torch (model/loss.py:45-70) and the warp is pure torch + one autograd.Function (model/LieAlgebra/se3.py:51-154).
 * The convention copied here is the one of the reference's only native extension, model/exfunc: a Python
 * torch.autograd.Function (functions/point_render_func.py:19-32) calls `ext.forward / ext.backward`, which unwrap
 * tensors to raw `float* / int*` + `int` sizes (expackages/point_render/point_render.cpp:29-38, 40-99, 101-133).
 * Every function below is the raw-pointer layer of that pattern for one reference symbol, cited per function.
 *
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device unless the name ends in `_host`;
 *   - fp32 data, int32 indices, element (not byte) strides, 64-bit;
 *   - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream); all calls are asynchronous on it,
 *     never synchronise the device, never allocate, and keep no state between calls (re-entrant, graph-capturable)
 *     — unlike the reference wrapper, which brackets every launch group with cudaDeviceSynchronize
 *     (point_render.cpp:54-55, 78-79, 95-96);
 *   - return value: 0 on success, a `cudaError_t` value (> 0) for CUDA failures, or a negative RMA_E_* code for
 *     argument errors.  Nothing calls exit() (the reference does: point_render.cpp:10-18).
 */
#ifndef RMA_B200_H_
#define RMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMA_B200_ABI_VERSION 3

#define RMA_OK 0
#define RMA_E_BADARG (-1)     /* null pointer / non-positive size / size overflow */
#define RMA_E_WORKSPACE (-2)  /* workspace too small or misaligned (256 B) */
#define RMA_E_ARCH (-3)       /* device is not sm_100 */

int rma_b200_abi_version(void);
const char* rma_b200_error_string(int code);
/* SM count and max SM clock (kHz) of the current device; used for the FP32 roofline denominator. */
int rma_b200_device_info(int* sm_count, int* clock_khz, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------------------------
 * Chamfer nearest-neighbour search, both directions.   Replaces compute_sqrdis_map + chamfer_dist
 * (model/loss.py:45-58, 60-70; callers loss.py:203, train_view.py:285-286, 379-380).
 *
 *   dist1[b,i] = min_j |x[b,i]-y[b,j]|^2     idx1[b,i] = first j attaining it      (i < N, j < M)
 *   dist2[b,j] = min_i |x[b,i]-y[b,j]|^2     idx2[b,j] = first i attaining it
 *
 * x is [B,N,3] with element strides (xsb, xsn, xsc); y is [B,M,3] with (ysb, ysn, ysc) — any strides, so the
 * batch-strided slices (train_view.py:369-370) and transposed stage outputs (loss.py:264) are taken as they are.
 * dist1/idx1 are contiguous [B,N]; dist2/idx2 contiguous [B,M].  idx1/idx2 may be NULL (forward-only metric calls).
 * Distances are evaluated in fp32 DIFFERENCE form (dx*dx, fma(dy,dy,.), fma(dz,dz,.)), never the expanded form;
 * the [B,N,M] matrix is never materialised.  workspace: rma_chamfer_workspace_bytes(B,N,M,opts) bytes, 256-B aligned.
 *
 * Per-call options (`opts`, may be NULL = all defaults; the SAME opts must be passed to workspace_bytes and to every
 * stage that works on that workspace).  Two search implementations return identical results: the filter + exact
 * resolve path (default) and the exact difference-form path, which takes over when the filter's group records
 * (B*N*M/64 bytes) would exceed record_cap_bytes (default 16 GiB = one 2^20 x 2^20 pair, whose filter workspace is 40 GiB; or
 * RMA_B200_FILTER_RECORD_CAP_MB read once at load time).  Everything is decided from the call's own arguments: the library keeps no mutable state between calls.
 */
#define RMA_CHAMFER_PATH_AUTO 0
#define RMA_CHAMFER_PATH_EXACT 1
#define RMA_CHAMFER_PATH_FILTER 2
typedef struct rma_chamfer_opts {
  int32_t path;               /* RMA_CHAMFER_PATH_*: force one search implementation (parity tests, A/B timing) */
  int32_t unit_run;           /* > 0: 32-column work units per search CTA (filter path) / column splits (exact path) */
  int64_t record_cap_bytes;   /* > 0: switch-over size of the automatic choice */
  int32_t y_batch_period;     /* > 0 (ABI 3; rma_chamfer_cd_forward* and rma_chamfer_prep only): y holds y_batch_period
                                 batch elements and batch element b of the call uses y[b % y_batch_period] - the stage
                                 loop's repeated target (loss.py:260-267) without materialising T copies of it.  B must be
                                 a multiple; idx / dist / gradient outputs stay per batch element of the call */
  int32_t reserved_;
} rma_chamfer_opts;
size_t rma_chamfer_workspace_bytes(int B, int N, int M, const rma_chamfer_opts* opts);
/* RMA_CHAMFER_PATH_EXACT or RMA_CHAMFER_PATH_FILTER: what (B,N,M,opts) runs (< 0: bad sizes). */
int rma_chamfer_path(int B, int N, int M, const rma_chamfer_opts* opts);
/* Guard statistics of the LAST resolve on this workspace (filter path; zeros on the exact path): rows on the slow
 * path, columns on the slow path, 512-row block rescans, unused.  Synchronises `stream`. */
int rma_chamfer_guard_stats(int B, int N, int M, void* workspace, const rma_chamfer_opts* opts,
                            uint32_t* stats4_host, void* stream);
int rma_chamfer_forward(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                        const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                        int B, int N, int M,
                        float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                        void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream);

/* The three stages of rma_chamfer_forward as separate calls on the same workspace (prep -> search -> finalize).
 * rma_chamfer_forward is exactly their composition; they exist so that a harness can bracket the dominant kernel
 * (search) with its own events (bench.py roofline leg) or capture the stages in a graph individually. */
int rma_chamfer_prep(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                     const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                     int B, int N, int M, void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts,
                     void* stream);
int rma_chamfer_search(int B, int N, int M, void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts,
                       void* stream);
int rma_chamfer_finalize(int B, int N, int M, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                         void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream);
/* Launch shape the search kernel will use for (B,N,M,opts) on the current device (reporting only).  On the filter path
 * the figures describe the VIRTUAL CTAs (groups of 1, 2 or 4 warps with their own shared memory and barrier; `threads`
 * threads each, `ctas_per_sm` of them inside one physical 256-thread CTA per SM) and `splits` is the longest run of
 * 32-column work units dealt to one of them. */
int rma_chamfer_search_config(int B, int N, int M, const rma_chamfer_opts* opts, int* ctas, int* threads, int* splits,
                              int* ctas_per_sm);

/* Host-only diagnostic (no device needed): checks the filter path's work decomposition for (B,N,M,opts) as it would be
 * derived on a device with `sms` SMs - every work unit covered by exactly one run, segment indices and record counts
 * inside what the workspace is carved for, search and resolve agreeing on every row block's segments.  Returns the
 * number of inconsistencies (0 = consistent) or -1 for unsupported sizes; *mode = 0 contiguous / 1 row-block aligned,
 * load_min_max[0..1] = fewest / most work units on one SM. */
int rma_chamfer_dealing_selfcheck(int B, int N, int M, int sms, const rma_chamfer_opts* opts, int* mode,
                                  int64_t* load_min_max);

/* Fused loss_on_cd (model/loss.py:200-205 on top of chamfer_dist): in the same three launches,
 *   *loss = 0.5 * (sum dist1 + sum dist2) / B                       (device scalar, overwritten)
 * and its gradient, split so that only the scattered half needs atomics:
 *   d loss / d x[b,i] = own_x[b,i,:] + scat_x[b,i,0:3]     own_x [B,N,3], scat_x [B,N,4] (16-byte aligned)
 *   d loss / d y[b,j] = own_y[b,j,:] + scat_y[b,j,0:3]     own_y [B,M,3], scat_y [B,M,4]
 * (all overwritten; pass own_x = scat_x = NULL and/or own_y = scat_y = NULL when that cloud needs no gradient).
 * The caller's backward is then rma_chamfer_cd_backward: one launch instead of a second scatter pass plus ~10
 * small torch kernels.  dist1/dist2/idx1/idx2 are optional outputs (NULL to skip). */
int rma_chamfer_cd_forward(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                           const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                           int B, int N, int M, float* loss, float* dist1, float* dist2, int32_t* idx1,
                           int32_t* idx2, float* own_x, float* scat_x, float* own_y, float* scat_y,
                           void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream);
/* The same with per-batch-element weights (device array [B] or NULL = 1) and an explicit coefficient:
 *   *loss = coef * sum_b batch_weight[b] * (sum dist1[b] + sum dist2[b]),   gradients scaled accordingly.
 * This is the stage loop of Loss.forward (loss.py:260-267) as ONE call: stack the T stage outputs as a [T*B,N,3] batch,
 * repeat the target, batch_weight[t*B + b] = weight_cd * gamma^(T-t-1), coef = 0.5 / B.
 * coef and every batch_weight[b] must be >= 0 (stage weights are): coef < 0 is RMA_E_BADARG, a negative weight makes
 * *loss NaN (the value is published as the largest rounded prefix of a sum of non-negative terms - repeated calls on
 * the same inputs then agree bit for bit without a completion ticket; csrc/chamfer_common.cuh, loss_commit). */
int rma_chamfer_cd_forward_weighted(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                                    const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                                    int B, int N, int M, const float* batch_weight, float coef, float* loss,
                                    float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* own_x,
                                    float* scat_x, float* own_y, float* scat_y, void* workspace,
                                    size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream);
/* out_x[p,c] = *scalar_dev * (own_x[p,c] + scat_x[p,c]) for p < nx points (out_x [nx,3]); same for y.  The upstream
 * scalar is read on the device (no host sync).  Either half may be empty (n = 0). */
int rma_chamfer_cd_backward(const float* own_x, const float* scat_x, int64_t nx, const float* own_y,
                            const float* scat_y, int64_t ny, const float* scalar_dev, float* out_x, float* out_y,
                            void* stream);

/* Backward of the above.  Replaces the implicit torch autograd of loss.py:45-70 (MinBackward x2, BmmBackward,
 * MulBackward — SURVEY.md §8a row a4).  g1 [B,N], g2 [B,M] contiguous upstream gradients (either may be NULL = 0).
 *   grad_x[b,i] = 2 g1[b,i] (x_i - y_idx1[b,i]) + sum_{j: idx2[b,j]=i} 2 g2[b,j] (x_i - y_j)
 *   grad_y[b,j] = 2 g2[b,j] (y_j - x_idx2[b,j]) + sum_{i: idx1[b,i]=j} 2 g1[b,i] (y_j - x_i)
 * grad_x [B,N,3], grad_y [B,M,3] contiguous; they are OVERWRITTEN (zeroed inside the call, then accumulated with
 * warp-aggregated atomics).  Either may be NULL when that cloud needs no gradient. */
int rma_chamfer_backward(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                         const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                         int B, int N, int M,
                         const float* g1, const float* g2, const int32_t* idx1, const int32_t* idx2,
                         float* grad_x, float* grad_y, void* stream);

/* Full squared-distance map, difference form: out[b,i,j] = |x[b,i]-y[b,j]|^2, contiguous [B,N,M].
 * Kept only so that `compute_sqrdis_map` (loss.py:45-58) stays callable; the hot path never uses it. */
int rma_sqrdis_map(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                   const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                   int B, int N, int M, float* out, void* stream);

/* Packed (distance,index) keys for the source-sharded huge-pair mode: key = float_bits(d) << 32 |
 * (idx + idx_offset).  d >= 0, so unsigned (and signed int64) order == (distance, then index) order and an
 * all-reduce-MIN over ranks merges the target->source direction with torch's lowest-index tie break. */
int rma_pack_min_keys(const float* dist, const int32_t* idx, int64_t n, int32_t idx_offset, uint64_t* keys,
                      void* stream);
int rma_unpack_min_keys(const uint64_t* keys, int64_t n, float* dist, int32_t* idx, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * se(3) exponential map.  Replaces LieAlgebra.se3.exp / ExpMap.forward (model/LieAlgebra/se3.py:51-74, 123-131)
 * with sinc1/2/3 and their |t| < 0.01 Taylor branches (sinc.py:5-17, 91-103, 120-132).
 * x [n,6] contiguous (w1,w2,w3,v1,v2,v3) -> g [n,4,4] contiguous, bottom row (0,0,0,1). */
int rma_se3_exp_forward(const float* x, int64_t n, float* g, void* stream);

/* ExpMap.backward (se3.py:133-152): grad_x[k] = sum_ij grad_g[i,j] (gen_k g)[i,j] — the left-perturbation derivative,
 * NOT the Jacobian of exp.  Closed form: grad_w = sum_c g[:3,c] x grad_g[:3,c],
 * grad_v = grad_g[:3,3].  x [n,6], grad_g [n,4,4] contiguous -> grad_x [n,6]. */
int rma_se3_exp_backward(const float* x, const float* grad_g, int64_t n, float* grad_x, void* stream);

/* se3.transform (se3.py:102-112): out = R a + p with R = g[:3,:3], p = g[:3,3].
 * g [G,4,4] contiguous; point (t, k) for t < G, k < P uses transform t.  a / out / grads are addressed as
 * base + t*s_t + k*s_k + c*s_c (c = coordinate), which covers both reference branches:
 *   g.dim()==a.dim():  a [*,3,N]  -> s_t = 3N, s_k = 1, s_c = N        (se3.py:108-109)
 *   else:              a [*,3]    -> g broadcast over P points: s_k = 3, s_c = 1 (hot call: g [B,1,4,4], a [B,N,3],
 *                                    network.py:569, 654, 843) or per-point transforms with P = 1. */
int rma_se3_transform_forward(const float* g, const float* a, int64_t a_st, int64_t a_sk, int64_t a_sc,
                              int64_t G, int64_t P, float* out, int64_t o_st, int64_t o_sk, int64_t o_sc,
                              void* stream);
/* Backward of transform (torch matmul autograd in the reference): grad_a = R^T go; grad_g[:3,:3] = sum_k go a^T;
 * grad_g[:3,3] = sum_k go; grad_g[3,:] = 0.  grad_g [G,4,4] contiguous (overwritten), grad_a strided like a.
 * Either output may be NULL.  The sums over the P points are accumulated in double (they feed the cancelling cross
 * products of ExpMap.backward) and rounded once; scratch: G*12 doubles (8-byte aligned), needed when grad_g is wanted
 * and P >= 64, zeroed inside the call. */
int rma_se3_transform_backward(const float* g, const float* a, int64_t a_st, int64_t a_sk, int64_t a_sc,
                               const float* go, int64_t go_st, int64_t go_sk, int64_t go_sc,
                               int64_t G, int64_t P, float* grad_g,
                               float* grad_a, int64_t ga_st, int64_t ga_sk, int64_t ga_sc, double* scratch,
                               void* stream);

/* Fused recurrent-loop warp (model/network.py:652-655; stage 0: 567-570; rigid net: 840-843):
 *   g = exp(phi[b]);  rigid = R src[b,k] + p;  out[b,k] = w ? w[b,k]*rigid + (1-w[b,k])*prev[b,k] : rigid
 * phi [B,6] contiguous; src/prev/out [B,N,3] with element strides (sb, sn, sc); w [B,N] contiguous or NULL
 * (then prev is ignored: stage 0 / rigid net).  g_out [B,4,4] (optional) receives exp(phi) (rigid_matrix_list);
 * rigid_out (optional, same strides as out) receives the un-blended rigid points (deform_rigid_points_list). */
int rma_warp_forward(const float* phi, const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                     const float* w, const float* prev, int64_t p_sb, int64_t p_sn, int64_t p_sc,
                     int B, int N, float* out, int64_t o_sb, int64_t o_sn, int64_t o_sc,
                     float* g_out, float* rigid_out, void* stream);
/* Backward of the fused warp w.r.t. phi and w (prev is detached in the reference, network.py:606; src is data).
 * Three gradients can arrive, one per output of the forward (any may be NULL, not all of go / go_rigid / grad_g_in
 * need be given): go = d/d out, go_rigid = d/d rigid_out (same strides as go; deform_rigid_points_list), grad_g_in
 * [B,4,4] = d/d g_out (rigid_matrix_list -> loss_on_tran, loss.py:224-230, 311-319).
 *   grad_w[b,k] = sum_c go[b,k,c] (rigid_c - prev_c);  grad_rigid = w*go (or go) + go_rigid;  then transform backward
 *   (+ grad_g_in) and ExpMap.backward chained:  grad_phi [B,6] (overwritten).  grad_src (optional, strided like src)
 *   = R^T grad_rigid.
 * The 12 per-batch sums and the cross products of ExpMap.backward are accumulated in double.
 * scratch: B*12 doubles (8-byte aligned), zeroed inside the call. */
int rma_warp_backward(const float* phi, const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                      const float* w, const float* prev, int64_t p_sb, int64_t p_sn, int64_t p_sc,
                      const float* go, const float* go_rigid, int64_t g_sb, int64_t g_sn, int64_t g_sc,
                      const float* grad_g_in, int B, int N, float* grad_phi, float* grad_w,
                      float* grad_src, int64_t gs_sb, int64_t gs_sn, int64_t gs_sc,
                      double* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * point_render / point_masker — the reference's `model/exfunc` CUDA extension.
 * These four entry points replace, one for one, the raw-pointer launchers the reference's pybind wrappers call:
 *   point_render_cuda_forward   point_render.cpp:29-33  / point_render_cuda.cu:337-352
 *   point_render_cuda_backward  point_render.cpp:35-38  / point_render_cuda.cu:354-367
 *   point_masker_cuda_forward   point_masker.cpp:29-30  / point_masker_cuda.cu:90-94
 *   point_masker_cuda_backward  point_masker.cpp:32-33  / point_masker_cuda.cu:96-100
 * with two differences in the contract: image / scratch buffers need NOT be pre-filled by the caller (the reference's
 * wrapper fills front = 0, back = -2, depth = 0, weight = epsilon, mask = 0, grads = 0 with torch before every call,
 * point_render.cpp:80-88, 119-120; here the kernels do it), and `weight_points` is optional (NULL): the backward
 * recomputes the Gaussian weights instead of reading a saved [B,N,cpw^2] tensor.
 *
 * proj_points [B,N,3] contiguous: (x, y) in pixels, z <= 0 by convention (smaller = closer; point_render_func.py:15).
 * Images are contiguous [B,H,W] (the reference's [B,H,W,1]).
 *
 * Forward: pixel_index[b,n] = (clamp(int(x),0,W-1), clamp(int(y),0,H-1));  front/back = min/max z over every point
 * whose visible_patch_width^2 patch covers the pixel;  is_visible = !(z > (front+back)/2 && z > front+threshold) at the
 * point's pixel;  for visible points g = expf(-|xy - pixel centre|^2 / ((cpw+1)/2)^2) over the colour patch:
 * weight_image = epsilon + sum g, depth_image = sum g*z  (fp32 atomics, summation order unspecified as in the
 * reference). */
int rma_point_render_forward(const float* proj_points, int B, int N, int H, int W, int visible_patch_width,
                             int color_patch_width, float epsilon, float threshold,
                             int32_t* pixel_index /* [B,N,2] */, int32_t* is_visible /* [B,N] */,
                             float* weight_points /* [B,N,cpw*cpw] or NULL */,
                             float* depth_image, float* weight_image,
                             float* front_image /* scratch [B,H,W] */, float* back_image /* scratch [B,H,W] */,
                             void* stream);
/* grad_proj_points [B,N,3] (overwritten): per point, in the reference thread's accumulation order, bit for bit:
 *   gwp = grad_depth*z + grad_weight;  grad_xy = -(gwp*g*2/divisor);  grad x,y += grad_xy*(xy - centre);
 *   grad z += grad_depth*g.   Hidden points get 0. */
int rma_point_render_backward(const float* grad_depth_image, const float* grad_weight_image, const float* proj_points,
                              const int32_t* pixel_index, const int32_t* is_visible, int B, int N, int H, int W,
                              int color_patch_width, float* grad_proj_points, void* stream);
/* mask_image [B,H,W] (overwritten): -1 inside every point's mask_patch_width^2 patch, 0 elsewhere. */
int rma_point_masker_forward(const float* proj_points, int B, int N, int H, int W, int mask_patch_width,
                             float* mask_image, void* stream);
/* grad_proj_points [B,N,3] (overwritten; only z is non-zero): for every pixel with |g| > threshold,
 *   grad_z[i] += g * e_i / (epsilon + sum_i e_i),  e_i = expf(-|xy_i - pixel centre|^2 / 1000).
 * Deterministic (no atomics): active pixels are compacted in pixel order.  scratch: see ..._scratch_bytes, 256-B aligned. */
size_t rma_point_masker_backward_scratch_bytes(int B, int N, int H, int W);
int rma_point_masker_backward(const float* grad_mask_image, const float* proj_points, int B, int N, int H, int W,
                              float epsilon, float threshold, float* grad_proj_points, void* scratch,
                              size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Brute-force k nearest neighbours.  Replaces the [B,N,N] matrix + torch.topk of
 *   knn(x, k)                        model/network.py:52-58  (k = 5, x [B,3,N]: pass strides (3N, 1, N))
 *   get_neighbor_index(vertices, n)  model/loss.py:99-110    (k = n + 1, skip_first = 1)
 * idx[b,i,:] = indices of the k candidates nearest to query i, ascending by (exact fp32 difference-form squared
 * distance, then index); the first `skip_first` of them are dropped, so idx (and dist, optional) are contiguous
 * [B,N,k-skip_first].  queries [B,N,3] / candidates [B,M,3] with element strides; pass the same tensor twice for
 * the reference's self-kNN.  k <= 32; slots beyond M candidates are -1. */
int rma_knn(const float* queries, int64_t qsb, int64_t qsn, int64_t qsc,
            const float* candidates, int64_t csb, int64_t csn, int64_t csc,
            int B, int N, int M, int k, int skip_first, int32_t* idx, float* dist, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * As-rigid-as-possible edge-length loss, all stages of a step in one launch.
 * Replaces Loss.loss_on_arap (model/loss.py:233-243) and its stage loop (model/loss.py:334-345):
 *   loss += coef * sum_g group_weight[g] * sum_{v,k} (sqrt(|d_j-d_v|^2 + 1e-5) - sqrt(|s_j-s_v|^2 + 1e-5))^2,
 *   j = neighbour_idx[b,v,k], b = g % B.
 * deformation [G,V,3] (G = stages * B) and source [B,V,3] with element strides; neighbour_idx [B,V,n] contiguous
 * (entries outside [0,V) are skipped); group_weight NULL (all 1) or [G]; loss [1] and grad_deformation (NULL or
 * [G,V,3] contiguous) must be ZEROED by the caller: the kernel accumulates d loss / d deformation into it. */
int rma_arap_forward(const float* deformation, int64_t dsb, int64_t dsn, int64_t dsc,
                     const float* source, int64_t ssb, int64_t ssn, int64_t ssc,
                     const int32_t* neighbour_idx, int G, int B, int V, int n,
                     const float* group_weight, float coef, float* loss, float* grad_deformation, void* stream);

/* Per-view depth comparison of loss_on_depth (model/loss.py:169-181) over n pixels (all views and batch elements):
 *   ori = ori_depth / ori_weight, tar = tar_depth / tar_weight, dis = |ori - tar|,
 *   mask = sign(ori * tar) * [dis <= dis_threshold],   *loss += coef * sum mask * dis^2,
 * and (when non-NULL) grad_ori_depth / grad_ori_weight [n] = d loss / d ori_depth, d loss / d ori_weight.
 * All arrays contiguous and 16-byte aligned; loss [1] must be zeroed by the caller. */
int rma_depth_view_term(const float* ori_depth, const float* ori_weight, const float* tar_depth,
                        const float* tar_weight, int64_t n, float dis_threshold, float coef, float* loss,
                        float* grad_ori_depth, float* grad_ori_weight, void* stream);

/* Multi-view projection of loss_on_depth / loss_on_mask (model/loss.py:161-166, 187-191), views folded into the batch:
 *   out[v,b,n,r] = (sum_c rotations[v][r][c] * points[b,n,c] + add3[r]) * mul3[r]        out [V*B, N, 3] contiguous.
 * points [B,N,3] with element strides; rotations [V,3,3] contiguous on the device (the reference's stacked
 * `all_rot_matrix_for_lfd` [3V,3]); add3 / mul3 are HOST arrays of 3 floats (depth: add (1,1,-1), mul ((res-1)/2,
 * (res-1)/2, 1); mask: add (1,1,-1), mul res/2).  backward: grad_points [B,N,3] contiguous (overwritten)
 *   = sum_v rotations[v]^T (mul3 * grad_out[v,b,n,:]). */
int rma_view_project_forward(const float* points, int64_t sb, int64_t sn, int64_t sc, const float* rotations, int V,
                             int B, int N, const float* add3, const float* mul3, float* out, void* stream);
int rma_view_project_backward(const float* grad_out, const float* rotations, int V, int B, int N, const float* mul3,
                              float* grad_points, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Host-buffer convenience entry (what a non-torch FFI binding would call): pageable or pinned HOST pointers,
 * contiguous [B,N,3] / [B,M,3]; allocates device memory, copies in, runs forward (+ backward of
 * loss_on_cd = 0.5 (sum dist1 + sum dist2) / B, loss.py:200-205, when grad pointers are non-NULL), copies out and
 * synchronises.  loss_host receives the scalar. */
int rma_chamfer_loss_host(const float* x_host, const float* y_host, int B, int N, int M,
                          float* loss_host, float* dist1_host, float* dist2_host,
                          int32_t* idx1_host, int32_t* idx2_host, float* grad_x_host, float* grad_y_host);

#ifdef __cplusplus
}
#endif
#endif /* RMA_B200_H_ */
