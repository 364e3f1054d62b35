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


}  // namespace rma
