// Shared device helpers for the B200 (sm_100a) chamfer / se(3) kernels.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <utility>
#include <cuda_runtime.h>

typedef unsigned long long u64;

#define RMA_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t e__ = (expr);                      \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

// Launch-error check that does not synchronise (keeps calls asynchronous and graph-capturable).
#define RMA_LAUNCH_CHECK()                         \
  do {                                             \
    cudaError_t e__ = cudaPeekAtLastError();       \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

namespace rma {

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// Every chamfer kernel starts with pdl_wait() (griddepcontrol.wait: the previous kernel in the stream has completed
// and its writes are visible) and is launched with the programmatic-stream-serialization attribute, so its launch
// latency and CTA ramp overlap the tail of its predecessor instead of following it (a few microseconds per boundary
// on a ~90 us step).  Works in eager streams and under stream capture.  RMA_B200_NO_PDL=1 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline bool pdl_enabled() {
  static const int on = [] { const char* e = std::getenv("RMA_B200_NO_PDL"); return (e && e[0] == '1') ? 0 : 1; }();
  return on != 0;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_dep(bool pdl, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                     cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                     cudaStream_t stream, Args&&... args) {
  return launch_dep(true, kernel, grid, block, smem, stream, std::forward<Args>(args)...);
}

// Experiment hook: RMA_B200_PDL_SITES is a bit mask of the launch sites that keep the attribute
// (1 prep, 2 search, 4 resolve, 8 cd_backward; default all).
static inline bool pdl_site(int bit) {
  static const int mask = [] { const char* e = std::getenv("RMA_B200_PDL_SITES"); return e ? std::atoi(e) : 15; }();
  return (mask & bit) != 0;
}

// ---- packed fp32x2 arithmetic (sm_100+: FADD2 / FMUL2 / FFMA2). ------------------------------------------------
// One instruction performs two IEEE-754 round-to-nearest fp32 operations (no ftz), bit-identical to the scalar
// __fadd_rn / __fmul_rn / __fmaf_rn on each half; measured on B200 at the same FLOP rate as scalar FFMA with half
// the issue slots (profiles/r01_fp32_microbench.jsonl), which is what leaves room for the min bookkeeping.
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 3-input minimum (sm_100+: FMNMX3).
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Scalar squared distance with the SAME operation order as the packed search loop, so that re-evaluating a
// candidate reproduces the search kernel's value bit for bit:  dx*dx -> fma(dy,dy,.) -> fma(dz,dz,.).
// `nyx` etc. are the NEGATED target coordinates (x + (-y) == x - y exactly in IEEE arithmetic).
__device__ __forceinline__ float sqdist_neg(float x, float y, float z, float nyx, float nyy, float nyz) {
  float dx = __fadd_rn(x, nyx), dy = __fadd_rn(y, nyy), dz = __fadd_rn(z, nyz);
  float d = __fmul_rn(dx, dx);
  d = __fmaf_rn(dy, dy, d);
  d = __fmaf_rn(dz, dz, d);
  return d;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// ---- TMA 1-D bulk copy global -> shared (cp.async.bulk, SASS UBLKCP), completion counted on an mbarrier --------
// The search kernels stage their target ("column") chunks with it: the chunk is a contiguous run of float4 in the
// prep kernel's pack, so one elected thread issues one bulk copy per chunk and the copy of chunk k+1 runs behind the
// FP32 work on chunk k (double buffer).  Source and destination 16-byte aligned, size a multiple of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the TMA unit) before the first copy is issued
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, u64* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace rma
