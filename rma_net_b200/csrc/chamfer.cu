// Batched chamfer nearest-neighbour search (both directions) + backward scatter for B200 (sm_100a).
//
// Replaces compute_sqrdis_map / chamfer_dist (reference model/loss.py:45-58, 60-70) and their implicit autograd
// (SURVEY.md §8a rows a1-a4).  The [B,N,M] distance matrix is never materialised.  See DESIGN.md §3-§4.
//
// Pipeline of one forward call (3 launches on the caller's stream):
//   1. chamfer_prep_kernel     strided AoS inputs -> register/TMA-friendly packs in the workspace, best-key init
//   2. chamfer_search_kernel   every pair distance evaluated ONCE (packed fp32x2 difference form) and used for
//                              both directions: running row minima stay in registers, column minima go through a
//                              per-warp shared-memory transpose; winners are merged with 64-bit atomicMin on
//                              (distance bits << 32 | coarse index) keys
//   3. chamfer_finalize_kernel resolves the coarse index (a 32-column tile / a 16-row group) to the exact first
//                              minimum by re-evaluating that tile with the identical operation order
// Backward: memset + one scatter kernel with warp-aggregated atomics.
#include "chamfer_common.cuh"
#include "../../include/rma_b200.h"

#include <cstdio>
#include <cstdlib>

namespace rma {

constexpr int kRowsPerThread = 16;                      // R: source points held in registers per thread
constexpr int kRowsPerWarp = 32 * kRowsPerThread;       // 512: one "row block"
constexpr int kSearchWarps = 4;                         // warps per CTA (one per SM sub-partition)
constexpr int kSearchThreads = 32 * kSearchWarps;
constexpr int kTile = 32;                               // columns per tile (= lanes, for the column transpose)
constexpr int kChunk = 1024;                            // columns staged in shared memory at a time
constexpr int kCbStride = 36;                           // row stride (floats) of the per-warp column-min transpose
                                                        // buffer: 16-B aligned rows, conflict-free LDS.128 / STS.32
// dynamic shared memory of the search: column double buffer, per-warp transpose buffers, two mbarriers
constexpr size_t kSearchSmem = (size_t)2 * (kChunk + 4) * sizeof(float4) +
                               (size_t)kSearchWarps * 2 * kTile * kCbStride * sizeof(float) + 16;
constexpr float kRowSentinel = 1e18f;                   // padded source rows:  d ~ 3e36 against real targets
constexpr float kColSentinelNeg = 3e18f;                // padded target columns (stored negated): d ~ 2.7e37

struct Geometry {
  int B, N, M;
  int nRB;    // row blocks (512 rows) per batch element
  int Npad;   // nRB * 512
  int nRBC;   // CTA row blocks (2048 rows) per batch element
  int Mp;     // padded column count: roundup(M, 32) + 32
  int yper;   // > 0: batch element b reads y[b % yper] (rma_chamfer_opts.y_batch_period)
  __host__ __device__ size_t rowpack_floats() const { return (size_t)B * Npad * 3; }
  __host__ __device__ size_t colpack_float4() const { return (size_t)B * Mp; }
};

static Geometry make_geometry(int B, int N, int M) {
  Geometry g;
  g.B = B; g.N = N; g.M = M;
  g.nRB = (N + kRowsPerWarp - 1) / kRowsPerWarp;
  g.Npad = g.nRB * kRowsPerWarp;
  g.nRBC = (g.nRB + kSearchWarps - 1) / kSearchWarps;
  g.Mp = (M + kTile - 1) / kTile * kTile + kTile;
  g.yper = 0;
  return g;
}

struct Workspace {
  float* rowpack;   // [B][nRB][12][32] float4 : per-thread 16 rows as x[16], y[16], z[16], lane-coalesced
  float4* colpack;  // [B][Mp] float4          : (-x, -y, -z, 0)
  u64* rowbest;     // [B][Npad]               : (dist bits << 32) | tile start column
  u64* colbest;     // [B][Mp]                 : (dist bits << 32) | 16-row group id
  double* lossacc;  // [1]                     : fused loss summed in double (loss_commit, chamfer_common.cuh)
  size_t bytes;
};

static Workspace carve(const Geometry& g, void* base) {
  Workspace w;
  char* p = (char*)base;
  size_t off = 0;
  w.rowpack = (float*)(p + off);  off += align256(g.rowpack_floats() * sizeof(float));
  w.colpack = (float4*)(p + off); off += align256(g.colpack_float4() * sizeof(float4));
  w.rowbest = (u64*)(p + off);    off += align256((size_t)g.B * g.Npad * sizeof(u64));
  w.colbest = (u64*)(p + off);    off += align256((size_t)g.B * g.Mp * sizeof(u64));
  w.lossacc = (double*)(p + off); off += align256(sizeof(double));
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// 1. prep
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chamfer_prep_kernel(
    const float* __restrict__ x, long long xsb, long long xsn, long long xsc,
    const float* __restrict__ y, long long ysb, long long ysn, long long ysc,
    Geometry g, Workspace w, float* __restrict__ zero_loss, float4* __restrict__ zero_gx,
    float4* __restrict__ zero_gy) {
  const long long nrow = (long long)g.B * g.Npad;
  const long long ncol = (long long)g.B * g.Mp;
  pdl_wait();
  if (zero_loss && blockIdx.x == 0 && threadIdx.x == 0) {
    *zero_loss = 0.f;
    w.lossacc[0] = 0.0;
  }
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nrow + ncol;
       t += (long long)gridDim.x * blockDim.x) {
    if (t < nrow) {
      const int b = (int)(t / g.Npad), i = (int)(t % g.Npad);
      float px = kRowSentinel, py = kRowSentinel, pz = kRowSentinel;
      if (i < g.N) {
        const float* s = x + b * xsb + i * xsn;
        px = s[0]; py = s[xsc]; pz = s[2 * xsc];
        if (zero_gx) zero_gx[(long long)b * g.N + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int rb = i / kRowsPerWarp, lane = (i % kRowsPerWarp) / kRowsPerThread, r = i % kRowsPerThread;
      // float4 slot q = c*4 + r/4 of this (row block, lane); element r%4
      const size_t base = ((size_t)(b * g.nRB + rb) * 12) * 32 + lane;
      const int q = r >> 2, e = r & 3;
      w.rowpack[(base + (size_t)(q)*32) * 4 + e] = px;
      w.rowpack[(base + (size_t)(4 + q) * 32) * 4 + e] = py;
      w.rowpack[(base + (size_t)(8 + q) * 32) * 4 + e] = pz;
      w.rowbest[t] = ~0ull;
    } else {
      const long long u = t - nrow;
      const int b = (int)(u / g.Mp), j = (int)(u % g.Mp);
      float nx = kColSentinelNeg, ny = kColSentinelNeg, nz = kColSentinelNeg;
      if (j < g.M) {
        const float* s = y + (g.yper > 0 ? b % g.yper : b) * ysb + j * ysn;
        nx = -s[0]; ny = -s[ysc]; nz = -s[2 * ysc];
        if (zero_gy) zero_gy[(long long)b * g.M + j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      w.colpack[u] = make_float4(nx, ny, nz, 0.f);
      w.colbest[u] = ~0ull;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. search
// ---------------------------------------------------------------------------------------------------------------
// One column against the thread's 16 rows (8 packed row pairs): 3 FADD2 + FMUL2 + 2 FFMA2 per row pair.
// Returns the minimum over the 16 rows (7 FMNMX3 + 1 FMNMX).
// One column against the thread's 16 rows (8 packed row pairs): 3 FADD2 + FMUL2 + 2 FFMA2 per row pair.
// (ncx, ncy, ncz) are the NEGATED column coordinates; pack2(c, c) lowers to the broadcast-scalar operand form
// (FADD2 R, R.F32x2, Rc.F32), so a column costs one LDS.128.  Returns the minimum over the 16 rows
// (7 FMNMX3 + 1 FMNMX).
__device__ __forceinline__ float column_step(const u64 (&px)[8], const u64 (&py)[8], const u64 (&pz)[8],
                                             float ncx, float ncy, float ncz, float (&d)[16]) {
  const u64 cx = pack2(ncx, ncx), cy = pack2(ncy, ncy), cz = pack2(ncz, ncz);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const u64 dx = add2(px[k], cx), dy = add2(py[k], cy), dz = add2(pz[k], cz);
    u64 t = mul2(dx, dx);
    t = fma2(dy, dy, t);
    t = fma2(dz, dz, t);
    unpack2(t, d[2 * k], d[2 * k + 1]);
  }
  const float t0 = min3(d[0], d[1], d[2]);
  const float t1 = min3(d[3], d[4], d[5]);
  const float t2 = min3(d[6], d[7], d[8]);
  const float t3 = min3(d[9], d[10], d[11]);
  const float t4 = min3(d[12], d[13], d[14]);
  const float u0 = min3(t0, t1, d[15]);
  const float u1 = min3(t2, t3, t4);
  return fminf(u0, u1);
}

struct SearchArgs {
  const float4* rowpack;
  const float4* colpack;
  u64* rowbest;
  u64* colbest;
  int N, M, nRB, Npad, nRBC, Mp, S;
};

__global__ void __launch_bounds__(kSearchThreads, 2) chamfer_search_kernel(SearchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* scol = reinterpret_cast<float4*>(smem_raw);                                           // [2][kChunk + 4]
  float* cball = reinterpret_cast<float*>(smem_raw + (size_t)2 * (kChunk + 4) * sizeof(float4));  // [warps][2][32][36]
  u64* mbar = reinterpret_cast<u64*>(smem_raw + kSearchSmem - 16);                              // [2]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
  __syncthreads();
  pdl_wait();
  const int s = blockIdx.x % a.S;
  const int rbc = (blockIdx.x / a.S) % a.nRBC;
  const int b = blockIdx.x / (a.S * a.nRBC);
  const int rb = rbc * kSearchWarps + warp;
  const bool has_rows = rb < a.nRB;   // warp-uniform
  const int c0 = (int)((long long)s * a.M / a.S);
  const int c1 = (int)((long long)(s + 1) * a.M / a.S);
  const int ncols = (c1 - c0 + kTile - 1) / kTile * kTile;   // tiles may overrun c1 (into real or sentinel columns)

  float* cb = cball + warp * (2 * kTile * kCbStride);
  u64* colbest = a.colbest + (size_t)b * a.Mp;

  // The thread's 16 source points, packed as row pairs (2k, 2k+1).
  u64 px[8], py[8], pz[8];
  float m[kRowsPerThread];
  int mt[kRowsPerThread];
  if (has_rows) {
    const float4* rp = a.rowpack + ((size_t)(b * a.nRB + rb) * 12) * 32 + lane;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 vx = rp[(size_t)q * 32], vy = rp[(size_t)(4 + q) * 32], vz = rp[(size_t)(8 + q) * 32];
      px[2 * q] = pack2(vx.x, vx.y); px[2 * q + 1] = pack2(vx.z, vx.w);
      py[2 * q] = pack2(vy.x, vy.y); py[2 * q + 1] = pack2(vy.z, vy.w);
      pz[2 * q] = pack2(vz.x, vz.y); pz[2 * q + 1] = pack2(vz.z, vz.w);
    }
  }
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r) { m[r] = __int_as_float(0x7f800000); mt[r] = 0; }

  const float4* cp = a.colpack + ((size_t)b * a.Mp + c0);
  // Column chunks arrive by TMA bulk copy into a double buffer: chunk 0 is in flight while the rows above are
  // loaded, chunk k+1 while chunk k is computed.
  if (tid == 0 && ncols > 0) bulk_load(scol, cp, (unsigned)min(kChunk, ncols) * 16u, &mbar[0]);
  int cur = 0;
  unsigned phases = 0;   // bit s: parity the next wait on buffer s expects

  // Column direction, software-pipelined: while tile t is being computed, lane l reduces column l of tile t-1
  // (32 per-lane partial minima in the other half of the double buffer), four values per loop body, so the
  // dependent compare/select chain hides behind the FP32 work instead of stalling at every tile boundary.
  int pend_start = -1;   // first column of the tile whose column reduce is pending (warp-uniform)
  int parity = 0;

  for (int cs = 0; cs < ncols; cs += kChunk) {
    const int nc = min(kChunk, ncols - cs);
    // the other buffer was last read in the previous iteration, which ended with a __syncthreads
    if (tid == 0 && cs + kChunk < ncols)
      bulk_load(scol + (cur ^ 1) * (kChunk + 4), cp + cs + kChunk, (unsigned)min(kChunk, ncols - cs - kChunk) * 16u,
                &mbar[cur ^ 1]);
    mbar_wait(&mbar[cur], (phases >> cur) & 1u);
    phases ^= 1u << cur;
    const float4* schunk = scol + cur * (kChunk + 4);
    cur ^= 1;

#pragma unroll 1
    for (int t0 = 0; has_rows && t0 < nc; t0 += kTile) {
      float tm[kRowsPerThread];
#pragma unroll
      for (int r = 0; r < kRowsPerThread; ++r) tm[r] = __int_as_float(0x7f800000);
      const float4* sp = schunk + t0;
      float* cbw = cb + parity * (kTile * kCbStride) + lane;                          // this tile: [col][lane]
      const float4* cbr = reinterpret_cast<const float4*>(cb + (parity ^ 1) * (kTile * kCbStride) +
                                                          lane * kCbStride);          // previous tile: my column
      float pm = __int_as_float(0x7f800000);
      int pid = 0;

      // operands of the first four columns; every body prefetches the next four (the last prefetch of a chunk
      // reads the 4-column pad behind the staged data and is never used)
      float4 n0 = sp[0], n1 = sp[1], n2 = sp[2], n3 = sp[3];

#pragma unroll 1
      for (int jj = 0; jj < kTile; jj += 4) {
        const float4 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
        n0 = sp[jj + 4]; n1 = sp[jj + 5]; n2 = sp[jj + 6]; n3 = sp[jj + 7];
        const float4 pv = cbr[jj >> 2];          // 4 partial minima of the pending tile (garbage for the first)
        {
          float dA[16], dB[16];
          const float cmA = column_step(px, py, pz, q0.x, q0.y, q0.z, dA);
          const float cmB = column_step(px, py, pz, q1.x, q1.y, q1.z, dB);
#pragma unroll
          for (int r = 0; r < kRowsPerThread; ++r) tm[r] = min3(tm[r], dA[r], dB[r]);
          cbw[jj * kCbStride] = cmA;
          cbw[(jj + 1) * kCbStride] = cmB;
        }
        {
          float dA[16], dB[16];
          const float cmA = column_step(px, py, pz, q2.x, q2.y, q2.z, dA);
          const float cmB = column_step(px, py, pz, q3.x, q3.y, q3.z, dB);
#pragma unroll
          for (int r = 0; r < kRowsPerThread; ++r) tm[r] = min3(tm[r], dA[r], dB[r]);
          cbw[(jj + 2) * kCbStride] = cmA;
          cbw[(jj + 3) * kCbStride] = cmB;
        }
        // pending column reduce: ascending lane order + strict '<' = lowest lane (lowest rows) wins ties
        bool l = pv.x < pm; pm = l ? pv.x : pm; pid = l ? jj : pid;
        l = pv.y < pm; pm = l ? pv.y : pm; pid = l ? jj + 1 : pid;
        l = pv.z < pm; pm = l ? pv.z : pm; pid = l ? jj + 2 : pid;
        l = pv.w < pm; pm = l ? pv.w : pm; pid = l ? jj + 3 : pid;
      }

      // Row direction: strict '<' keeps the earliest tile on ties (lowest index, torch.min semantics).
      const int tile_start = c0 + cs + t0;
#pragma unroll
      for (int r = 0; r < kRowsPerThread; ++r) {
        const bool better = tm[r] < m[r];
        m[r] = better ? tm[r] : m[r];
        mt[r] = better ? tile_start : mt[r];
      }

      if (pend_start >= 0 && pend_start + lane < c1) {
        const u64 key = ((u64)__float_as_uint(pm) << 32) | (unsigned)(rb * 32 + pid);
        atomicMin(colbest + pend_start + lane, key);
      }
      pend_start = tile_start;
      parity ^= 1;
      __syncwarp();   // this tile's partial minima visible to all lanes; previous buffer free for reuse
    }
    __syncthreads();   // chunk fully consumed: the next iteration's prefetch refills this buffer
  }

  if (has_rows) {
    if (pend_start >= 0) {   // drain: column reduce of the last tile
      const float* col = cb + (parity ^ 1) * (kTile * kCbStride) + lane * kCbStride;
      float pm = __int_as_float(0x7f800000);
      int pid = 0;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float v = col[k];
        const bool l = v < pm;
        pm = l ? v : pm;
        pid = l ? k : pid;
      }
      if (pend_start + lane < c1) {
        const u64 key = ((u64)__float_as_uint(pm) << 32) | (unsigned)(rb * 32 + pid);
        atomicMin(colbest + pend_start + lane, key);
      }
    }
    const int i0 = rb * kRowsPerWarp + lane * kRowsPerThread;
    u64* rbest = a.rowbest + (size_t)b * a.Npad + i0;
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
      if (i0 + r < a.N) atomicMin(rbest + r, ((u64)__float_as_uint(m[r]) << 32) | (unsigned)mt[r]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. finalize: coarse winner -> exact (distance, first index); optionally fused loss_on_cd value + gradients
// ---------------------------------------------------------------------------------------------------------------
// One warp per block of 32 items.  Row item: the 32 columns of the winning tile are evaluated by the 32 lanes
// (one coalesced 1 KB read); the winning VALUE is already known from the key, so a ballot on equality gives the
// first minimum.  Column item: the 16 rows of the winning row group are evaluated by a half warp.
// Distances use sqdist_neg = the search kernel's operation order, so equality is exact.
__global__ void __launch_bounds__(kFinWarps * 32) chamfer_finalize_kernel(Geometry g, Workspace w, FinalizeArgs f) {
  __shared__ float4 stage[kFinWarps][32];
  __shared__ float stagev[kFinWarps][32];
  __shared__ unsigned hits[kFinWarps][32];
  __shared__ double lsum[kFinWarps];
  const long long nrow = (long long)g.B * g.N, ncol = (long long)g.B * g.M;
  const long long row_warps = (nrow + 31) / 32;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const long long wi = (long long)blockIdx.x * kFinWarps + warp;
  const float coef_grad = 2.f * f.coef_loss;
  float myd = 0.f, wb = 1.f;
  bool valid = false;
  pdl_wait();
  pdl_launch_dependents();

  if (wi < row_warps) {
    const long long t = wi * 32 + lane;
    valid = t < nrow;
    int b = 0, i = 0, jt = 0;
    float x = 0.f, y = 0.f, z = 0.f;
    if (valid) {
      b = (int)(t / g.N); i = (int)(t % g.N);
      const u64 key = w.rowbest[(size_t)b * g.Npad + i];
      jt = (int)(unsigned)(key & 0xffffffffu);
      myd = __uint_as_float((unsigned)(key >> 32));
      const int rb = i / kRowsPerWarp, ln = (i % kRowsPerWarp) / kRowsPerThread, r = i % kRowsPerThread;
      const size_t base = ((size_t)(b * g.nRB + rb) * 12) * 32 + ln;
      const int q = r >> 2, e = r & 3;
      x = w.rowpack[(base + (size_t)q * 32) * 4 + e];
      y = w.rowpack[(base + (size_t)(4 + q) * 32) * 4 + e];
      z = w.rowpack[(base + (size_t)(8 + q) * 32) * 4 + e];
    }
    stage[warp][lane] = make_float4(__int_as_float(b * g.Mp + jt), x, y, z);
    stagev[warp][lane] = myd;
    __syncwarp();
    const float4* cbase = w.colpack + lane;
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 16) {
      float4 c[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) c[k] = cbase[__float_as_int(stage[warp][r0 + k].x)];   // 16 loads in flight
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 sv = stage[warp][r0 + k];
        const float d = sqdist_neg(sv.y, sv.z, sv.w, c[k].x, c[k].y, c[k].z);
        const unsigned hit = __ballot_sync(0xffffffffu, d == stagev[warp][r0 + k]);
        hits[warp][r0 + k] = hit;     // same value from every lane
      }
    }
    __syncwarp();
    const unsigned myhit = hits[warp][lane];
    const int j = jt + (myhit ? __ffs(myhit) - 1 : 0);
    // no candidate reproduces the winning value: only with NaN coordinates (every comparison fails).  NaN distance,
    // index 0 - the same contract as the filter path's resolve.
    if (valid && !myhit) myd = __int_as_float(0x7fc00000);
    wb = (valid && f.batch_weight) ? f.batch_weight[b] : 1.f;
    if (valid) {
      if (f.dist1) f.dist1[t] = myd;
      if (f.idx1) f.idx1[t] = j;
    }
    if (f.own_x || f.scat_y) {
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (valid) {
        const float4 c = w.colpack[(size_t)b * g.Mp + j];
        const float cg = coef_grad * wb;
        vx = cg * (x + c.x); vy = cg * (y + c.y); vz = cg * (z + c.z);
        if (f.own_x) { float* o = f.own_x + t * 3; o[0] = vx; o[1] = vy; o[2] = vz; }
      }
      if (f.scat_y) scatter_add4(f.scat_y + ((long long)b * g.M + j), (long long)b * g.M + j, valid, -vx, -vy, -vz);
    }
  } else {
    const long long t = (wi - row_warps) * 32 + lane;
    valid = t < ncol;
    int b = 0, j = 0, grp = 0, fbase = 0;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (valid) {
      b = (int)(t / g.M); j = (int)(t % g.M);
      const u64 key = w.colbest[(size_t)b * g.Mp + j];
      grp = (int)(unsigned)(key & 0xffffffffu);
      myd = __uint_as_float((unsigned)(key >> 32));
      const float4 c = w.colpack[(size_t)b * g.Mp + j];
      nx = c.x; ny = c.y; nz = c.z;
      fbase = (((b * g.nRB + (grp >> 5)) * 12) * 32 + (grp & 31)) * 4;   // float index of (group, q = 0, e = 0)
    }
    stage[warp][lane] = make_float4(__int_as_float(fbase), nx, ny, nz);
    stagev[warp][lane] = myd;
    __syncwarp();
    const int half = lane >> 4, rr = lane & 15;
    const float* rbase = w.rowpack + ((rr >> 2) * 32) * 4 + (rr & 3);      // row rr inside a group
    float rx[16], ry[16], rz[16];
#pragma unroll
    for (int p = 0; p < 16; ++p) {                                          // 48 loads in flight
      const float* rp = rbase + __float_as_int(stage[warp][half * 16 + p].x);
      rx[p] = rp[0]; ry[p] = rp[4 * 32 * 4]; rz[p] = rp[8 * 32 * 4];
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const int r = half * 16 + p;                                          // item handled by this half warp
      const float4 sv = stage[warp][r];
      const float d = sqdist_neg(rx[p], ry[p], rz[p], sv.y, sv.z, sv.w);
      const unsigned hit = __ballot_sync(0xffffffffu, d == stagev[warp][r]);
      hits[warp][r] = (hit >> (half * 16)) & 0xffffu;   // same value from every lane of the half warp
    }
    __syncwarp();
    const unsigned myhit = hits[warp][lane];
    const int myi = myhit ? __ffs(myhit) - 1 : 0;
    if (valid && !myhit) myd = __int_as_float(0x7fc00000);   // NaN coordinates, as above
    const int i = grp * kRowsPerThread + myi;
    wb = (valid && f.batch_weight) ? f.batch_weight[b] : 1.f;
    if (valid) {
      if (f.dist2) f.dist2[t] = myd;
      if (f.idx2) f.idx2[t] = i;
    }
    if (f.own_y || f.scat_x) {
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (valid) {
        const float* rp = w.rowpack + fbase + ((myi >> 2) * 32) * 4 + (myi & 3);
        const float x = rp[0], y = rp[4 * 32 * 4], z = rp[8 * 32 * 4];
        // d/dy_j |y_j - x_i|^2 = 2 (y_j - x_i) = -2 (x_i + (-y_j))
        const float cg = coef_grad * wb;
        vx = -cg * (x + nx); vy = -cg * (y + ny); vz = -cg * (z + nz);
        if (f.own_y) { float* o = f.own_y + t * 3; o[0] = vx; o[1] = vy; o[2] = vz; }
      }
      if (f.scat_x) scatter_add4(f.scat_x + ((long long)b * g.N + i), (long long)b * g.N + i, valid, -vx, -vy, -vz);
    }
  }

  if (f.loss) {
    const double tot = loss_partial<kFinWarps>(valid, myd, wb, f.coef_loss, lsum);
    if (threadIdx.x == 0) loss_commit(w.lossacc, tot, f.loss);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: gradient scatter with warp-aggregated atomics (generic upstream gradients g1, g2)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chamfer_backward_kernel(
    const float* __restrict__ x, long long xsb, long long xsn, long long xsc,
    const float* __restrict__ y, long long ysb, long long ysn, long long ysc,
    int B, int N, int M, const float* __restrict__ g1, const float* __restrict__ g2,
    const int* __restrict__ idx1, const int* __restrict__ idx2, float* __restrict__ grad_x,
    float* __restrict__ grad_y, int vec_ok) {
  const long long nrow = (long long)B * N, ncol = (long long)B * M;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // grid covers whole warps
  // Every lane of the warp takes part in the match/shuffle; out-of-range or zero-gradient lanes carry a unique
  // negative key and contribute nothing.
  float vx = 0.f, vy = 0.f, vz = 0.f;
  long long key = -1 - (long long)lane_id();
  float* own = nullptr;
  float* other = nullptr;
  long long self = 0, partner = 0;
  bool live = false;
  if (t < nrow) {
    if (g1) {
      const int b = (int)(t / N), i = (int)(t % N);
      const int j = idx1[t];
      const float gg = 2.f * g1[t];
      const float* p = x + b * xsb + i * xsn;
      const float* q = y + b * ysb + j * ysn;
      vx = gg * (p[0] - q[0]); vy = gg * (p[xsc] - q[ysc]); vz = gg * (p[2 * xsc] - q[2 * ysc]);
      own = grad_x ? grad_x + t * 3 : nullptr;
      other = grad_y;
      self = t; partner = (long long)b * M + j;
      key = partner;
      live = true;
    }
  } else if (t < nrow + ncol) {
    if (g2) {
      const long long u = t - nrow;
      const int b = (int)(u / M), j = (int)(u % M);
      const int i = idx2[u];
      const float gg = 2.f * g2[u];
      const float* p = y + b * ysb + j * ysn;
      const float* q = x + b * xsb + i * xsn;
      vx = gg * (p[0] - q[0]); vy = gg * (p[ysc] - q[xsc]); vz = gg * (p[2 * ysc] - q[2 * xsc]);
      own = grad_y ? grad_y + u * 3 : nullptr;
      other = grad_x;
      self = u; partner = (long long)b * N + i;
      key = ncol + partner;   // x->y keys live in [0, ncol): disjoint (a warp can straddle both halves)
      live = true;
    }
  }
  // The kernel is bound by L2 transactions (profiles/r01p_hbm_kernels_dram.md): with 8-byte aligned outputs a point's
  // (x, y) or (y, z) pair goes out as one vector atomic, two transactions per point instead of three.
  if (live && own) {   // own-point term: one writer per address from this half, the other half scatters into it
    if (vec_ok) {
      if ((self & 1) == 0) { atomicAdd(reinterpret_cast<float2*>(own), make_float2(vx, vy)); atomicAdd(own + 2, vz); }
      else { atomicAdd(own, vx); atomicAdd(reinterpret_cast<float2*>(own + 1), make_float2(vy, vz)); }
    } else {
      atomicAdd(own + 0, vx); atomicAdd(own + 1, vy); atomicAdd(own + 2, vz);
    }
  }
  if (vec_ok) scatter_add3_v2(other, partner, key, live && other != nullptr, -vx, -vy, -vz);
  else scatter_add3(other ? other + partner * 3 : nullptr, key, live && other != nullptr, -vx, -vy, -vz);
}

__global__ void __launch_bounds__(256) sqrdis_map_kernel(
    const float* __restrict__ x, long long xsb, long long xsn, long long xsc,
    const float* __restrict__ y, long long ysb, long long ysn, long long ysc,
    int B, int N, int M, float* __restrict__ out) {
  const long long total = (long long)B * N * M;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % M);
    const long long bi = t / M;
    const int i = (int)(bi % N), b = (int)(bi / N);
    const float* p = x + b * xsb + i * xsn;
    const float* q = y + b * ysb + j * ysn;
    out[t] = sqdist_neg(p[0], p[xsc], p[2 * xsc], -q[0], -q[ysc], -q[2 * ysc]);
  }
}

__global__ void pack_keys_kernel(const float* __restrict__ dist, const int* __restrict__ idx, long long n,
                                 int off, u64* __restrict__ keys) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) keys[t] = ((u64)__float_as_uint(dist[t]) << 32) | (unsigned)(idx[t] + off);
}
__global__ void unpack_keys_kernel(const u64* __restrict__ keys, long long n, float* __restrict__ dist,
                                   int* __restrict__ idx) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const u64 k = keys[t];
    dist[t] = __uint_as_float((unsigned)(k >> 32));
    idx[t] = (int)(unsigned)(k & 0xffffffffu);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------

struct DeviceCaps { int sms; int search_occ; };

static int get_caps(DeviceCaps* out) {
  // Re-queried per call (a few microseconds of host time, no device work): keeps the library free of global
  // state so that it is safe across devices and threads.
  int dev = 0;
  RMA_CUDA_TRY(cudaGetDevice(&dev));
  RMA_CUDA_TRY(cudaDeviceGetAttribute(&out->sms, cudaDevAttrMultiProcessorCount, dev));
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSearchSmem));
  RMA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&out->search_occ, chamfer_search_kernel,
                                                             kSearchThreads, kSearchSmem));
  if (out->search_occ < 1) out->search_occ = 1;
  return 0;
}

// Number of column splits S: CTAs = B * nRBC * S, each sweeping ~M/S columns for its 2048 rows.  Co-resident CTAs
// share the SM's FP32 pipe, so time ~ ceil(CTAs / SMs) * (columns per CTA + fixed per-CTA overhead).
static int choose_splits(const Geometry& g, const DeviceCaps& caps) {
  const long long base = (long long)g.B * g.nRBC;
  const int smax = g.M / 64 > 1 ? (g.M / 64 < 4096 ? g.M / 64 : 4096) : 1;
  const double overhead_cols = 40.0;
  double best_cost = 1e300;
  int best = 1;
  for (int s = 1; s <= smax; ++s) {
    const long long ctas = base * s;
    const long long per_sm = (ctas + caps.sms - 1) / caps.sms;
    double cost = (double)per_sm * ((double)((g.M + s - 1) / s) + overhead_cols);
    if (per_sm == 1) cost *= 1.05;   // a single warp per sub-partition hides less latency
    if (cost < best_cost * 0.999) { best_cost = cost; best = s; }
    if (ctas > 64LL * caps.sms) break;
  }
  return best;
}

}  // namespace rma

using namespace rma;

// Filter + exact-resolve path (chamfer_filter.cu).
namespace rma {
bool f_sizes_ok(int B, int N, int M);
size_t f_record_bytes(int B, int N, int M);
size_t f_workspace_bytes(int B, int N, int M, int unit_run);
int f_launch_prep(const float* x, int64_t xsb, int64_t xsn, int64_t xsc, const float* y, int64_t ysb, int64_t ysn,
                  int64_t ysc, int B, int N, int M, int unit_run, int y_period, void* workspace, size_t workspace_bytes,
                  float* zero_loss, float4* zero_gx, float4* zero_gy, void* stream_);
int f_search_config(int B, int N, int M, int unit_run, int* ctas, int* threads, int* splits, int* ctas_per_sm);
int f_launch_search(int B, int N, int M, int unit_run, void* workspace, size_t workspace_bytes, void* stream_);
int f_launch_resolve(int B, int N, int M, int unit_run, const FinalizeArgs& f, void* workspace,
                     size_t workspace_bytes, void* stream_);
int f_read_stats(void* workspace, int B, int N, int M, int unit_run, unsigned* out4, void* stream_);
int f_dealing_selfcheck(int B, int N, int M, int sms, int unit_run, int* mode_out, long long* load2_out);
}  // namespace rma

// Which search runs: the filter path whenever its group records (B*N*M/64 bytes) stay below the cap, else the
// exact path (huge single pairs, SURVEY.md config C5 on few GPUs).  Both return identical results.  Every choice is
// a function of the call's own arguments (sizes + rma_chamfer_opts): the library keeps no mutable state.
// Default record cap: 16 GiB - C5 (one pair of 2^20 x 2^20 points) on ONE GPU is exactly at it and runs the filter path
// with a 40 GiB workspace (records 16 GiB + row entries 25 GiB) out of 180 GB of HBM: 204 ms of search at 87 % of the
// FP32 FMA peak against 238 ms on the exact path.  It was 4 GiB until r02n (the filter path only from 4 GPUs upward).
// The environment variable RMA_B200_FILTER_RECORD_CAP_MB, read once when the library is loaded, changes the default;
// callers short of memory pass record_cap_bytes or RMA_CHAMFER_PATH_EXACT per call (workspace 2.8 MiB per 2^16 points).
static const size_t g_default_record_cap = [] {
  const char* e = std::getenv("RMA_B200_FILTER_RECORD_CAP_MB");
  const long long mb = e ? std::atoll(e) : 0;
  return mb > 0 ? (size_t)mb << 20 : (size_t)16 << 30;
}();

struct ChamferOpts {
  int path;        // RMA_CHAMFER_PATH_*
  int unit_run;    // 0 = automatic
  size_t cap;      // filter record cap in bytes
  int y_period;    // 0 = every batch element has its own y
};

static ChamferOpts make_opts(const rma_chamfer_opts* o) {
  ChamferOpts r = {RMA_CHAMFER_PATH_AUTO, 0, g_default_record_cap, 0};
  if (o) {
    if (o->path == RMA_CHAMFER_PATH_EXACT || o->path == RMA_CHAMFER_PATH_FILTER) r.path = o->path;
    if (o->unit_run > 0) r.unit_run = o->unit_run;
    if (o->record_cap_bytes > 0) r.cap = (size_t)o->record_cap_bytes;
    if (o->y_batch_period > 0) r.y_period = o->y_batch_period;
  }
  return r;
}

static bool use_filter(int B, int N, int M, const ChamferOpts& o) {
  if (o.path == RMA_CHAMFER_PATH_EXACT || !f_sizes_ok(B, N, M)) return false;
  if (o.path == RMA_CHAMFER_PATH_FILTER) return true;
  return f_record_bytes(B, N, M) <= o.cap;
}

extern "C" {

int rma_b200_abi_version(void) { return RMA_B200_ABI_VERSION; }

const char* rma_b200_error_string(int code) {
  if (code == RMA_OK) return "ok";
  if (code == RMA_E_BADARG) return "rma_b200: bad argument (null pointer, non-positive size or size overflow)";
  if (code == RMA_E_WORKSPACE) return "rma_b200: workspace too small or not 256-byte aligned";
  if (code == RMA_E_ARCH) return "rma_b200: device is not sm_100 (B200)";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "rma_b200: unknown error";
}

int rma_b200_device_info(int* sm_count, int* clock_khz, int* cc_major, int* cc_minor) {
  int dev = 0;
  RMA_CUDA_TRY(cudaGetDevice(&dev));
  if (sm_count) RMA_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (clock_khz) RMA_CUDA_TRY(cudaDeviceGetAttribute(clock_khz, cudaDevAttrClockRate, dev));
  if (cc_major) RMA_CUDA_TRY(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) RMA_CUDA_TRY(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return RMA_OK;
}

static bool sizes_ok(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return false;
  // int32 point indices inside the kernels: padded per-batch sizes must fit, total slots use 64-bit.
  if ((long long)N > (1LL << 30) || (long long)M > (1LL << 30)) return false;
  if ((long long)B * ((long long)N + kRowsPerWarp) >= (1LL << 31) / 12) return false;
  if ((long long)B * ((long long)M + 2 * kTile) >= (1LL << 31) / 2) return false;
  return true;
}

size_t rma_chamfer_workspace_bytes(int B, int N, int M, const rma_chamfer_opts* opts) {
  if (!sizes_ok(B, N, M)) return 0;
  const ChamferOpts o = make_opts(opts);
  const Geometry g = make_geometry(B, N, M);
  size_t bytes = carve(g, nullptr).bytes;
  if (use_filter(B, N, M, o)) { const size_t fb = f_workspace_bytes(B, N, M, o.unit_run); if (fb > bytes) bytes = fb; }
  return bytes;
}

int rma_chamfer_dealing_selfcheck(int B, int N, int M, int sms, const rma_chamfer_opts* opts, int* mode,
                                  int64_t* load_min_max) {
  const ChamferOpts o = make_opts(opts);
  long long l2[2] = {0, 0};
  const int r = f_dealing_selfcheck(B, N, M, sms, o.unit_run, mode, l2);
  if (load_min_max) { load_min_max[0] = l2[0]; load_min_max[1] = l2[1]; }
  return r;
}

int rma_chamfer_path(int B, int N, int M, const rma_chamfer_opts* opts) {
  if (!sizes_ok(B, N, M)) return RMA_E_BADARG;
  return use_filter(B, N, M, make_opts(opts)) ? RMA_CHAMFER_PATH_FILTER : RMA_CHAMFER_PATH_EXACT;
}

int rma_chamfer_guard_stats(int B, int N, int M, void* workspace, const rma_chamfer_opts* opts,
                            uint32_t* stats4_host, void* stream) {
  if (!sizes_ok(B, N, M) || !workspace || !stats4_host) return RMA_E_BADARG;
  const ChamferOpts o = make_opts(opts);
  if (!use_filter(B, N, M, o)) { stats4_host[0] = stats4_host[1] = stats4_host[2] = stats4_host[3] = 0u; return RMA_OK; }
  return f_read_stats(workspace, B, N, M, o.unit_run, stats4_host, stream);
}

static int forward_setup(const float* x, const float* y, int B, int N, int M, void* workspace,
                         size_t workspace_bytes, Geometry* g, Workspace* w) {
  if (!x || !y || !workspace || !sizes_ok(B, N, M)) return RMA_E_BADARG;
  *g = make_geometry(B, N, M);
  if (((uintptr_t)workspace & 255u) != 0) return RMA_E_WORKSPACE;
  *w = carve(*g, workspace);
  if (workspace_bytes < w->bytes) return RMA_E_WORKSPACE;
  return RMA_OK;
}

static int launch_prep(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                       const float* y, int64_t ysb, int64_t ysn, int64_t ysc, int B, int N, int M,
                       const ChamferOpts& o, void* workspace, size_t workspace_bytes, float* zero_loss,
                       float4* zero_gx, float4* zero_gy, void* stream_) {
  if (o.y_period > 0 && (B <= 0 || B % o.y_period != 0)) return RMA_E_BADARG;
  if (x && y && sizes_ok(B, N, M) && use_filter(B, N, M, o))
    return f_launch_prep(x, xsb, xsn, xsc, y, ysb, ysn, ysc, B, N, M, o.unit_run, o.y_period, workspace, workspace_bytes,
                         zero_loss, zero_gx, zero_gy, stream_);
  Geometry g; Workspace w;
  if (int e = forward_setup(x, y, B, N, M, workspace, workspace_bytes, &g, &w)) return e;
  g.yper = o.y_period;
  const long long slots = (long long)B * (g.Npad + g.Mp);
  RMA_CUDA_TRY(launch_pdl(chamfer_prep_kernel, (unsigned)grid_for(slots, 256), 256u, 0, (cudaStream_t)stream_,
                          x, (long long)xsb, (long long)xsn, (long long)xsc, y, (long long)ysb, (long long)ysn,
                          (long long)ysc, g, w, zero_loss, zero_gx, zero_gy));
  return RMA_OK;
}

int rma_chamfer_prep(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                     const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                     int B, int N, int M, void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts,
                     void* stream_) {
  return launch_prep(x, xsb, xsn, xsc, y, ysb, ysn, ysc, B, N, M, make_opts(opts), workspace, workspace_bytes,
                     nullptr, nullptr, nullptr, stream_);
}

static int launch_search(int B, int N, int M, const ChamferOpts& o, void* workspace, size_t workspace_bytes,
                         void* stream_) {
  if (sizes_ok(B, N, M) && use_filter(B, N, M, o))
    return f_launch_search(B, N, M, o.unit_run, workspace, workspace_bytes, stream_);
  Geometry g; Workspace w;
  if (int e = forward_setup((const float*)workspace, (const float*)workspace, B, N, M, workspace, workspace_bytes,
                            &g, &w)) return e;
  DeviceCaps caps;
  if (int e = get_caps(&caps)) return e;
  SearchArgs a;
  a.rowpack = reinterpret_cast<const float4*>(w.rowpack);
  a.colpack = w.colpack;
  a.rowbest = w.rowbest;
  a.colbest = w.colbest;
  a.N = N; a.M = M; a.nRB = g.nRB; a.Npad = g.Npad; a.nRBC = g.nRBC; a.Mp = g.Mp;
  a.S = o.unit_run > 0 ? o.unit_run : choose_splits(g, caps);   // exact path: unit_run = column splits per row block
  const long long ctas = (long long)B * g.nRBC * a.S;
  if (ctas > 0x7fffffffLL) return RMA_E_BADARG;
  RMA_CUDA_TRY(launch_pdl(chamfer_search_kernel, (unsigned)ctas, (unsigned)kSearchThreads, kSearchSmem,
                          (cudaStream_t)stream_, a));
  return RMA_OK;
}

int rma_chamfer_search(int B, int N, int M, void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts,
                       void* stream_) {
  return launch_search(B, N, M, make_opts(opts), workspace, workspace_bytes, stream_);
}

static int launch_finalize(int B, int N, int M, const ChamferOpts& o, const FinalizeArgs& f, void* workspace,
                           size_t workspace_bytes, void* stream_) {
  if (sizes_ok(B, N, M) && use_filter(B, N, M, o))
    return f_launch_resolve(B, N, M, o.unit_run, f, workspace, workspace_bytes, stream_);
  Geometry g; Workspace w;
  if (int e = forward_setup((const float*)workspace, (const float*)workspace, B, N, M, workspace, workspace_bytes,
                            &g, &w)) return e;
  const long long warps = ((long long)B * N + 31) / 32 + ((long long)B * M + 31) / 32;
  RMA_CUDA_TRY(launch_pdl(chamfer_finalize_kernel, (unsigned)grid_for(warps, kFinWarps), (unsigned)(kFinWarps * 32), 0,
                          (cudaStream_t)stream_, g, w, f));
  return RMA_OK;
}

int rma_chamfer_finalize(int B, int N, int M, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                         void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream_) {
  if (!dist1 || !dist2) return RMA_E_BADARG;
  FinalizeArgs f = {dist1, dist2, idx1, idx2, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr};
  return launch_finalize(B, N, M, make_opts(opts), f, workspace, workspace_bytes, stream_);
}

int rma_chamfer_search_config(int B, int N, int M, const rma_chamfer_opts* opts, int* ctas, int* threads, int* splits,
                              int* ctas_per_sm) {
  if (!sizes_ok(B, N, M)) return RMA_E_BADARG;
  const ChamferOpts o = make_opts(opts);
  if (use_filter(B, N, M, o)) return f_search_config(B, N, M, o.unit_run, ctas, threads, splits, ctas_per_sm);
  const Geometry g = make_geometry(B, N, M);
  DeviceCaps caps;
  if (int e = get_caps(&caps)) return e;
  const int S = o.unit_run > 0 ? o.unit_run : choose_splits(g, caps);
  if (ctas) *ctas = B * g.nRBC * S;
  if (threads) *threads = kSearchThreads;
  if (splits) *splits = S;
  if (ctas_per_sm) *ctas_per_sm = caps.search_occ;
  return RMA_OK;
}

int rma_chamfer_forward(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                        const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                        int B, int N, int M, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                        void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream) {
  if (!dist1 || !dist2) return RMA_E_BADARG;
  if (int e = rma_chamfer_prep(x, xsb, xsn, xsc, y, ysb, ysn, ysc, B, N, M, workspace, workspace_bytes, opts, stream))
    return e;
  if (int e = rma_chamfer_search(B, N, M, workspace, workspace_bytes, opts, stream)) return e;
  return rma_chamfer_finalize(B, N, M, dist1, dist2, idx1, idx2, workspace, workspace_bytes, opts, stream);
}

int rma_chamfer_cd_forward_weighted(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                                    const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                                    int B, int N, int M, const float* batch_weight, float coef, float* loss,
                                    float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, float* own_x,
                                    float* scat_x, float* own_y, float* scat_y, void* workspace,
                                    size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream) {
  if (!loss || !(coef >= 0.f)) return RMA_E_BADARG;
  const ChamferOpts o = make_opts(opts);
  if ((own_x == nullptr) != (scat_x == nullptr) || (own_y == nullptr) != (scat_y == nullptr)) return RMA_E_BADARG;
  if ((((uintptr_t)scat_x) | ((uintptr_t)scat_y)) & 15u) return RMA_E_BADARG;
  if (int e = launch_prep(x, xsb, xsn, xsc, y, ysb, ysn, ysc, B, N, M, o, workspace, workspace_bytes, loss,
                          (float4*)scat_x, (float4*)scat_y, stream))
    return e;
  if (int e = launch_search(B, N, M, o, workspace, workspace_bytes, stream)) return e;
  FinalizeArgs f = {dist1, dist2, idx1, idx2, loss, own_x, own_y, (float4*)scat_x, (float4*)scat_y, coef,
                    batch_weight};
  return launch_finalize(B, N, M, o, f, workspace, workspace_bytes, stream);
}

int rma_chamfer_cd_forward(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                           const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                           int B, int N, int M, float* loss, float* dist1, float* dist2, int32_t* idx1,
                           int32_t* idx2, float* own_x, float* scat_x, float* own_y, float* scat_y,
                           void* workspace, size_t workspace_bytes, const rma_chamfer_opts* opts, void* stream) {
  if (B <= 0) return RMA_E_BADARG;
  return rma_chamfer_cd_forward_weighted(x, xsb, xsn, xsc, y, ysb, ysn, ysc, B, N, M, nullptr, 0.5f / (float)B, loss,
                                         dist1, dist2, idx1, idx2, own_x, scat_x, own_y, scat_y, workspace,
                                         workspace_bytes, opts, stream);
}

// out[p, c] = scalar * (own[p, c] + scat[p].c): combines the two gradient halves of the fused forward and applies
// the upstream scalar, both clouds in one launch.
__global__ void __launch_bounds__(256) cd_backward_kernel(const float* __restrict__ own_x,
                                                          const float4* __restrict__ scat_x, long long nx,
                                                          const float* __restrict__ own_y,
                                                          const float4* __restrict__ scat_y, long long ny,
                                                          const float* __restrict__ scalar,
                                                          float* __restrict__ out_x, float* __restrict__ out_y) {
  pdl_wait();
  const float sc = *scalar;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nx + ny;
       t += (long long)gridDim.x * blockDim.x) {
    const bool isx = t < nx;
    const long long p = isx ? t : t - nx;
    const float* own = (isx ? own_x : own_y) + p * 3;
    const float4 s4 = (isx ? scat_x : scat_y)[p];
    float* o = (isx ? out_x : out_y) + p * 3;
    o[0] = sc * (own[0] + s4.x); o[1] = sc * (own[1] + s4.y); o[2] = sc * (own[2] + s4.z);
  }
}

int rma_chamfer_cd_backward(const float* own_x, const float* scat_x, int64_t nx, const float* own_y,
                            const float* scat_y, int64_t ny, const float* scalar_dev, float* out_x, float* out_y,
                            void* stream_) {
  if (!scalar_dev || nx < 0 || ny < 0) return RMA_E_BADARG;
  if (nx > 0 && (!own_x || !scat_x || !out_x)) return RMA_E_BADARG;
  if (ny > 0 && (!own_y || !scat_y || !out_y)) return RMA_E_BADARG;
  if (nx + ny == 0) return RMA_OK;
  long long grid = (nx + ny + 255) / 256;
  if (grid > 148LL * 16) grid = 148LL * 16;
  RMA_CUDA_TRY(launch_dep(pdl_site(8), cd_backward_kernel, (unsigned)grid, 256u, 0, (cudaStream_t)stream_, own_x,
                          (const float4*)scat_x, (long long)nx, own_y, (const float4*)scat_y, (long long)ny, scalar_dev,
                          out_x, out_y));
  return RMA_OK;
}

int rma_chamfer_backward(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                         const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                         int B, int N, int M, const float* g1, const float* g2, const int32_t* idx1,
                         const int32_t* idx2, float* grad_x, float* grad_y, void* stream_) {
  if (!x || !y || !sizes_ok(B, N, M)) return RMA_E_BADARG;
  if ((g1 && !idx1) || (g2 && !idx2)) return RMA_E_BADARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (grad_x) RMA_CUDA_TRY(cudaMemsetAsync(grad_x, 0, (size_t)B * N * 3 * sizeof(float), stream));
  if (grad_y) RMA_CUDA_TRY(cudaMemsetAsync(grad_y, 0, (size_t)B * M * 3 * sizeof(float), stream));
  if ((!grad_x && !grad_y) || (!g1 && !g2)) return RMA_OK;
  const long long n = (long long)B * ((long long)N + M);
  // vector atomics need 8-byte aligned bases (any cudaMalloc / torch allocation; an odd float offset falls back)
  const int vec_ok = (((uintptr_t)grad_x | (uintptr_t)grad_y) & 7u) == 0;
  chamfer_backward_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, xsb, xsn, xsc, y, ysb, ysn, ysc, B, N, M, g1,
                                                                g2, idx1, idx2, grad_x, grad_y, vec_ok);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_sqrdis_map(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                   const float* y, int64_t ysb, int64_t ysn, int64_t ysc,
                   int B, int N, int M, float* out, void* stream_) {
  if (!x || !y || !out || B <= 0 || N <= 0 || M <= 0) return RMA_E_BADARG;
  const long long total = (long long)B * N * M;
  long long grid = (total + 255) / 256;
  if (grid > 148LL * 64) grid = 148LL * 64;
  sqrdis_map_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream_>>>(x, xsb, xsn, xsc, y, ysb, ysn, ysc, B, N, M,
                                                                  out);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_pack_min_keys(const float* dist, const int32_t* idx, int64_t n, int32_t idx_offset, uint64_t* keys,
                      void* stream_) {
  if (!dist || !idx || !keys || n <= 0) return RMA_E_BADARG;
  pack_keys_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(dist, idx, n, idx_offset, (u64*)keys);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

int rma_unpack_min_keys(const uint64_t* keys, int64_t n, float* dist, int32_t* idx, void* stream_) {
  if (!dist || !idx || !keys || n <= 0) return RMA_E_BADARG;
  unpack_keys_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>((const u64*)keys, n, dist, idx);
  RMA_LAUNCH_CHECK();
  return RMA_OK;
}

// loss_on_cd reduction for the host-buffer entry: 0.5 * (sum d1 + sum d2) / B in double.
__global__ void __launch_bounds__(256) loss_sum_kernel(const float* __restrict__ d1, long long n1,
                                                       const float* __restrict__ d2, long long n2,
                                                       double* __restrict__ acc) {
  double s = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n1 + n2;
       t += (long long)gridDim.x * blockDim.x)
    s += (double)(t < n1 ? d1[t] : d2[t - n1]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane_id() == 0) atomicAdd(acc, s);
}

__global__ void fill_kernel(float* p, long long n, float v) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}

int rma_chamfer_loss_host(const float* x_host, const float* y_host, int B, int N, int M, float* loss_host,
                          float* dist1_host, float* dist2_host, int32_t* idx1_host, int32_t* idx2_host,
                          float* grad_x_host, float* grad_y_host) {
  if (!x_host || !y_host || !sizes_ok(B, N, M)) return RMA_E_BADARG;
  const size_t nx = (size_t)B * N, ny = (size_t)B * M;
  const size_t wsb = rma_chamfer_workspace_bytes(B, N, M, nullptr);
  const bool want_grad = grad_x_host || grad_y_host;
  // one device allocation, carved
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  const size_t o_x = take(nx * 12), o_y = take(ny * 12), o_d1 = take(nx * 4), o_d2 = take(ny * 4),
               o_i1 = take(nx * 4), o_i2 = take(ny * 4), o_ws = take(wsb), o_acc = take(8),
               o_g1 = take(want_grad ? nx * 4 : 0), o_g2 = take(want_grad ? ny * 4 : 0),
               o_gx = take(want_grad ? nx * 12 : 0), o_gy = take(want_grad ? ny * 12 : 0);
  char* d = nullptr;
  RMA_CUDA_TRY(cudaMalloc((void**)&d, off));
  cudaStream_t st = 0;
  int rc = RMA_OK;
  auto fail = [&](int e) { cudaFree(d); return e; };
#define TRY_(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail((int)e__); } while (0)
  TRY_(cudaMemcpyAsync(d + o_x, x_host, nx * 12, cudaMemcpyHostToDevice, st));
  TRY_(cudaMemcpyAsync(d + o_y, y_host, ny * 12, cudaMemcpyHostToDevice, st));
  rc = rma_chamfer_forward((float*)(d + o_x), (int64_t)N * 3, 3, 1, (float*)(d + o_y), (int64_t)M * 3, 3, 1, B, N, M,
                           (float*)(d + o_d1), (float*)(d + o_d2), (int32_t*)(d + o_i1), (int32_t*)(d + o_i2),
                           d + o_ws, wsb, nullptr, st);
  if (rc) return fail(rc);
  TRY_(cudaMemsetAsync(d + o_acc, 0, 8, st));
  loss_sum_kernel<<<148, 256, 0, st>>>((float*)(d + o_d1), (long long)nx, (float*)(d + o_d2), (long long)ny,
                                       (double*)(d + o_acc));
  if (want_grad) {
    const float gval = 0.5f / (float)B;
    fill_kernel<<<grid_for((long long)nx, 256), 256, 0, st>>>((float*)(d + o_g1), (long long)nx, gval);
    fill_kernel<<<grid_for((long long)ny, 256), 256, 0, st>>>((float*)(d + o_g2), (long long)ny, gval);
    rc = rma_chamfer_backward((float*)(d + o_x), (int64_t)N * 3, 3, 1, (float*)(d + o_y), (int64_t)M * 3, 3, 1, B, N,
                              M, (float*)(d + o_g1), (float*)(d + o_g2), (int32_t*)(d + o_i1),
                              (int32_t*)(d + o_i2), (float*)(d + o_gx), (float*)(d + o_gy), st);
    if (rc) return fail(rc);
  }
  double acc = 0.0;
  TRY_(cudaMemcpyAsync(&acc, d + o_acc, 8, cudaMemcpyDeviceToHost, st));
  if (dist1_host) TRY_(cudaMemcpyAsync(dist1_host, d + o_d1, nx * 4, cudaMemcpyDeviceToHost, st));
  if (dist2_host) TRY_(cudaMemcpyAsync(dist2_host, d + o_d2, ny * 4, cudaMemcpyDeviceToHost, st));
  if (idx1_host) TRY_(cudaMemcpyAsync(idx1_host, d + o_i1, nx * 4, cudaMemcpyDeviceToHost, st));
  if (idx2_host) TRY_(cudaMemcpyAsync(idx2_host, d + o_i2, ny * 4, cudaMemcpyDeviceToHost, st));
  if (grad_x_host) TRY_(cudaMemcpyAsync(grad_x_host, d + o_gx, nx * 12, cudaMemcpyDeviceToHost, st));
  if (grad_y_host) TRY_(cudaMemcpyAsync(grad_y_host, d + o_gy, ny * 12, cudaMemcpyDeviceToHost, st));
  TRY_(cudaStreamSynchronize(st));
#undef TRY_
  if (loss_host) *loss_host = (float)(0.5 * acc / (double)B);
  cudaFree(d);
  return RMA_OK;
}

}  // extern "C"
