// Pieces shared by the exact search path (chamfer.cu) and the filter + exact-resolve path (chamfer_filter.cu).
#pragma once
#include "common.cuh"

namespace rma {

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 0x7fffffffLL) g = 0x7fffffffLL;
  return (int)g;
}

// Sum (vx,vy,vz) over the lanes that share a key; the result is valid on the lowest lane of each peer group.
__device__ __forceinline__ void reduce_peers(unsigned peers, float& vx, float& vy, float& vz) {
  const unsigned lane = lane_id();
  unsigned rel = __popc(peers & ((1u << lane) - 1u));
  unsigned higher = peers & ~((2u << lane) - 1u);
  while (__any_sync(0xffffffffu, higher != 0)) {
    const int next = __ffs(higher);
    const int src = next ? next - 1 : (int)lane;
    const float tx = __shfl_sync(0xffffffffu, vx, src);
    const float ty = __shfl_sync(0xffffffffu, vy, src);
    const float tz = __shfl_sync(0xffffffffu, vz, src);
    if (next) { vx += tx; vy += ty; vz += tz; }
    const unsigned done = __ballot_sync(0xffffffffu, rel & 1u);
    higher &= ~done;
    rel >>= 1;
  }
}

// Warp-aggregated scatter: lanes with equal `key` (>= 0) are summed and the group's lowest lane issues one atomic
// per coordinate.  Every lane of the warp must call this; lanes with live == false contribute nothing.
__device__ __forceinline__ void scatter_add3(float* dst, long long key, bool live, float vx, float vy, float vz) {
  if (!live) key = -1 - (long long)lane_id();
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (__any_sync(0xffffffffu, peers != (1u << lane_id()))) reduce_peers(peers, vx, vy, vz);
  if (live && (unsigned)(__ffs(peers) - 1) == lane_id()) {
    atomicAdd(dst + 0, vx); atomicAdd(dst + 1, vy); atomicAdd(dst + 2, vz);
  }
}

// Same into a dense [.,3] float buffer whose base is 8-byte aligned: point p starts at byte 12 p, so either its (x, y) or
// its (y, z) pair is 8-byte aligned and goes out as one vector atomic - two L2 transactions per target instead of three
// (the generic backward is bound by L2 transactions, profiles/r01p_hbm_kernels_dram.md).
__device__ __forceinline__ void scatter_add3_v2(float* base, long long point, long long key, bool live, float vx,
                                                float vy, float vz) {
  if (!live) key = -1 - (long long)lane_id();
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (__any_sync(0xffffffffu, peers != (1u << lane_id()))) reduce_peers(peers, vx, vy, vz);
  if (live && (unsigned)(__ffs(peers) - 1) == lane_id()) {
    float* dst = base + point * 3;
    if ((point & 1) == 0) {
      atomicAdd(reinterpret_cast<float2*>(dst), make_float2(vx, vy));
      atomicAdd(dst + 2, vz);
    } else {
      atomicAdd(dst, vx);
      atomicAdd(reinterpret_cast<float2*>(dst + 1), make_float2(vy, vz));
    }
  }
}

// Warp-aggregated scatter into a float4-strided buffer: one 16-byte vector atomic (RED.128) per distinct target.
__device__ __forceinline__ void scatter_add4(float4* dst, long long key, bool live, float vx, float vy, float vz) {
  if (!live) key = -1 - (long long)lane_id();
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (__any_sync(0xffffffffu, peers != (1u << lane_id()))) reduce_peers(peers, vx, vy, vz);
  if (live && (unsigned)(__ffs(peers) - 1) == lane_id()) atomicAdd(dst, make_float4(vx, vy, vz, 0.f));
}

struct FinalizeArgs {
  float* dist1; float* dist2; int* idx1; int* idx2;   // any may be null
  float* loss;       // null, or [1] accumulator (zeroed by prep): += coef_loss * (sum dist1 + sum dist2)
  // Fused gradient of coef_loss * (sum dist1 + sum dist2), split so that only the scattered half needs atomics:
  //   d/dx[b,i] = own_x[b,i] + scat_x[b,i].xyz      own_*: plain stores, one writer per point
  //   d/dy[b,j] = own_y[b,j] + scat_y[b,j].xyz      scat_*: float4-strided, zeroed by prep, vector atomics
  float* own_x; float* own_y;      // null or [B,N,3] / [B,M,3]
  float4* scat_x; float4* scat_y;  // null or [B,N] / [B,M]
  float coef_loss;   // 0.5 / B for loss_on_cd
  const float* batch_weight;   // null, or [B]: batch element b counts with weight batch_weight[b] (stage weights)
};

constexpr int kFinWarps = 8;

// Fused loss value.  Every CTA adds its share - summed in DOUBLE - to a double accumulator in the workspace (zeroed by
// prep) with ONE value-returning atomic and publishes fl32(previous + own) with an integer atomicMax on the caller's
// float: the terms are non-negative (squared distances, weights >= 0), so the prefix sums only grow, their float
// roundings are monotone, and the largest one is the rounding of the total.  The order of the double additions moves the
// total by ~1e-16 relative, i.e. repeated runs agree to the last bit in practice, and no CTA has to know that it is the
// last one: a "last CTA rounds the sum" ticket (atomic, fence, atomic, and fence + load + store on the last CTA) cost
// 2.3 us at the tail of a 77 us step, float atomics are free but differ by ~1e-6 between runs.
// A negative weight breaks the monotonicity, so it poisons the sum instead: NaN, which as an integer is larger than every
// finite value and sticks (NaN / inf distances propagate the same way, as they do in the reference's torch.mean).
template <int kWarps>
__device__ __forceinline__ double loss_partial(bool valid, float d, float wb, float coef, double* lsum) {
  double sd = valid ? (double)d * (double)wb : 0.0;
  if (valid && wb < 0.f) sd = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
  if (lane_id() == 0) lsum[threadIdx.x >> 5] = sd;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kWarps; ++k) tot += lsum[k];
    tot *= (double)coef;
  }
  return tot;   // valid on thread 0
}

__device__ __forceinline__ void loss_commit(double* acc, double tot, float* loss) {
  const double prev = atomicAdd(acc, tot);
  const float r = (float)(prev + tot);
  atomicMax(reinterpret_cast<int*>(loss), r == r ? __float_as_int(r) : 0x7fc00000);
}


}  // namespace rma
