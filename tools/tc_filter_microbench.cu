// Evidence for the "no tensor cores" choice (BASELINE.json north_star): what would a tcgen05 filter
// tile cost?  NOT a product path - the product search is FP32 FMA by contract (csrc/chamfer_filter.cu).
//
// One CTA per SM.  A = 128 rows, B = 256 columns per MMA, K = 16 TF32 slots that hold a 3xTF32 split of the whole
// squared distance, so the accumulator IS the filter value and the epilogue has no FADD at all:
//     A row   = (xh0,xh1,xh2,  xl0,xl1,xl2,  xh0,xh1,xh2,  uh, ul, 1, 1, 0, 0, 0)          x = xh + xl,  u = |x|^2 = uh + ul
//     B col   = (qh0,qh1,qh2,  qh0,qh1,qh2,  ql0,ql1,ql2,  1,  1,  wh, wl, 0, 0, 0)        q = -2y = qh + ql, w = |y|^2
//     A . B   = xh.qh + xl.qh + xh.ql + u + w  ~  |x - y|^2       (tcgen05.mma kind::tf32, 2 instructions of K = 8, FP32 accumulate in TMEM)
// The epilogue (8 warps, thread = one row x one column half) reads the 128 x 256 accumulator with tcgen05.ld.32x32b.x32
// and keeps per row the running (smallest tagged 16-column group minimum, its 64-column tile, second smallest) exactly as
// the FP32 search kernel's row direction does (FMNMX3 chain, LOP3 tag, 3 FMNMX top-2).  The column direction of a real
// kernel would be the same tile with the operand roles swapped (a second MMA + the same epilogue), so one "pair
// evaluation for both directions" costs two of the tile passes timed here.
// Accumulators are double-buffered in TMEM (2 x 256 columns): the MMA of tile t+1 runs under the epilogue of tile t.
// B tiles are either resident in shared memory (mode 0) or streamed from an L2-resident array through an 8-stage ring of
// 16 KB bulk copies (mode 1: what a real kernel pays to feed N = 256 columns x 64 bytes per MMA).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/tc_filter_microbench tools/tc_filter_microbench.cu
//   tools/bin/tc_filter_microbench          -> JSON lines: cycles per pair and direction, error against fp64
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

typedef unsigned long long u64;
constexpr int TM = 128, TN = 256, TK = 16, NSTAGE = 8;
constexpr int A_BYTES = TM * TK * 4, B_BYTES = TN * TK * 4;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_wait(u64* b, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// K-major, no swizzle ("interleave"): 16-byte K chunks; chunk kc of row r at kc * LBO + (r / 8) * SBO + (r % 8) * 16
__device__ __forceinline__ u64 make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  u64 d = 0;
  d |= (u64)((saddr & 0x3ffffu) >> 4);
  d |= (u64)(lbo_bytes >> 4) << 16;
  d |= (u64)(sbo_bytes >> 4) << 32;
  d |= (u64)1 << 46;            // descriptor version: Blackwell
  return d;                      // layout_type 0 = no swizzle, base offset 0
}
// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 256
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TN >> 3) << 17) | ((unsigned)(TM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(unsigned tmem_d, u64 da, u64 db, unsigned accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(u64* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Args {
  const float* a_tiles;     // [n_a][TM*TK] pre-arranged canonical layout
  const float* b_tiles;     // [n_b][TN*TK]
  int n_b;                  // B tiles per pass over the columns
  int tiles;                // MMA tiles per CTA
  int stream_b;             // 0: B resident (NSTAGE tiles, loaded once), 1: every tile's B comes through the ring
  int epilogue;             // 0: none (MMA only), 1: tcgen05.ld only, 2: ld + row minima (the filter's row direction)
  float* out;               // [ctas][TM][2 column halves][3]
  long long* cycles;        // [ctas]
};

__global__ void __launch_bounds__(288, 1) tc_filter_kernel(Args g) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = reinterpret_cast<float*>(smem + A_BYTES);                              // [NSTAGE][B_BYTES]
  u64* bars = reinterpret_cast<u64*>(smem + A_BYTES + NSTAGE * B_BYTES);
  u64* bfull = bars;                  // [NSTAGE] bulk copy landed
  u64* bempty = bars + NSTAGE;        // [NSTAGE] the MMAs that read the stage have completed
  u64* afull = bars + 2 * NSTAGE;     // [2] accumulator ready
  u64* aempty = bars + 2 * NSTAGE + 2;  // [2] accumulator drained by the 4 epilogue warps
  u64* abar = bars + 2 * NSTAGE + 4;  // A tile landed
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 8); }
    mbar_init(abar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = tmem_base_s;
  const long long t_start = clock64();

  if (warp == 8) {
    if (lane == 0) {
      // ---------------- producer + MMA issuer (one thread) ----------------
      mbar_expect_tx(abar, A_BYTES);
      bulk_load(sA, g.a_tiles + (size_t)(blockIdx.x % 4) * (TM * TK), A_BYTES, abar);
      const int pre = g.tiles < NSTAGE ? g.tiles : NSTAGE;
      for (int s = 0; s < pre; ++s) {
        mbar_expect_tx(&bfull[s], B_BYTES);
        bulk_load(sB + (size_t)s * (TN * TK), g.b_tiles + (size_t)((blockIdx.x + s) % g.n_b) * (TN * TK), B_BYTES, &bfull[s]);
      }
      mbar_wait(abar, 0);
      const unsigned a_addr = smem_u32(sA);
      for (int t = 0; t < g.tiles; ++t) {
        const int s = t % NSTAGE, acc = t & 1;
        if (g.stream_b || t < NSTAGE) mbar_wait(&bfull[s], (unsigned)((t / NSTAGE) & 1));
        if (t >= 2) mbar_wait(&aempty[acc], (unsigned)(((t >> 1) - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned b_addr = smem_u32(sB + (size_t)s * (TN * TK));
        const unsigned d = tmem_base + (unsigned)(acc * TN);
#pragma unroll
        for (int k = 0; k < TK / 8; ++k)   // each instruction: K = 8 tf32 = two 16-byte chunks
          mma_tf32(d, make_desc(a_addr + k * 2 * (TM * 16), TM * 16, 128), make_desc(b_addr + k * 2 * (TN * 16), TN * 16, 128), k);
        mma_commit(&afull[acc]);
        if (g.stream_b) {
          mma_commit(&bempty[s]);
          // refill the stage that tile t - 1 used (its MMAs are older than this commit's, so they finish first)
          if (t >= 1 && t - 1 + NSTAGE < g.tiles) {
            const int ps = (t - 1) % NSTAGE;
            mbar_wait(&bempty[ps], (unsigned)(((t - 1) / NSTAGE) & 1));
            mbar_expect_tx(&bfull[ps], B_BYTES);
            bulk_load(sB + (size_t)ps * (TN * TK), g.b_tiles + (size_t)((blockIdx.x + t - 1 + NSTAGE) % g.n_b) * (TN * TK),
                      B_BYTES, &bfull[ps]);
          }
        }
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; warp w reads TMEM lanes 32 (w % 4) .. + 31 (its hardware quarter) and the
    // column half w / 4 of the accumulator: thread <-> (row, column half).  Loads are software-pipelined: the
    // tcgen05.ld of the next 32 columns is in flight while the minima of the current 32 are computed, and the second
    // warp of the sub-partition computes while this one waits.
    const float inf = __int_as_float(0x7f800000);
    const int quarter = warp & 3, chalf = warp >> 2;
    float m1 = inf, m2 = inf;
    int t1 = 0;
    float sink = 0.f;
    auto tmem_ld32 = [](unsigned (&v)[32], unsigned addr) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
          "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(addr));
    };
    for (int t = 0; t < g.tiles; ++t) {
      const int acc = t & 1;
      mbar_wait(&afull[acc], (unsigned)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (g.epilogue >= 1) {
        const unsigned taddr = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(acc * TN + chalf * (TN / 2));
        unsigned va[32], vb[32];
        float tm = inf, ts = inf;
        auto fold32 = [&](const unsigned (&v)[32], int c32) {      // c32: which 32-column chunk of this half (0..3)
          if (g.epilogue >= 2) {
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {        // two 16-column groups
              float gm = fminf(__uint_as_float(v[16 * gq]), __uint_as_float(v[16 * gq + 1]));
#pragma unroll
              for (int k = 2; k < 16; k += 2) gm = min3(gm, __uint_as_float(v[16 * gq + k]), __uint_as_float(v[16 * gq + k + 1]));
              const float tg = __uint_as_float((__float_as_uint(gm) & ~3u) | (unsigned)((c32 & 1) * 2 + gq));
              ts = fminf(ts, fmaxf(tm, tg));
              tm = fminf(tm, tg);
            }
            if (c32 & 1) {                          // end of a 64-column tile of the row direction
              const int tile = t * (TN / 64) + chalf * 2 + (c32 >> 1);
              m2 = min3(m2, ts, fmaxf(m1, tm));
              t1 = tm < m1 ? tile : t1;
              m1 = fminf(m1, tm);
              tm = inf; ts = inf;
            }
          } else {
            sink += __uint_as_float(v[0] ^ v[31]);
          }
        };
        tmem_ld32(va, taddr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tmem_ld32(vb, taddr + 32);
        fold32(va, 0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tmem_ld32(va, taddr + 64);
        fold32(vb, 1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tmem_ld32(vb, taddr + 96);
        fold32(va, 2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        fold32(vb, 3);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&aempty[acc]);
    }
    float* o = g.out + (((size_t)blockIdx.x * TM + quarter * 32 + lane) * 2 + chalf) * 3;
    o[0] = m1 + sink * 0.f; o[1] = __int_as_float(t1); o[2] = m2;
  }
  __syncthreads();
  if (tid == 0) g.cycles[blockIdx.x] = clock64() - t_start;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

static float tf32_hi(float v) { uint32_t b; memcpy(&b, &v, 4); b &= 0xffffe000u; float r; memcpy(&r, &b, 4); return r; }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const int n_a = 4, n_b = 64;          // 4 row tiles, 64 column tiles (16384 columns, 1 MB of B operands: L2 resident)
  std::vector<float> xs(n_a * TM * 3), ys((size_t)n_b * TN * 3);
  srand(1);
  for (auto& v : xs) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : ys) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  auto put = [](std::vector<float>& dst, size_t tile_off, int rows, int r, const float* kv) {
    for (int k = 0; k < TK; ++k) dst[tile_off + (size_t)(k / 4) * (rows * 4) + (size_t)(r / 8) * 32 + (r % 8) * 4 + (k % 4)] = kv[k];
  };
  std::vector<float> ha((size_t)n_a * TM * TK), hb((size_t)n_b * TN * TK);
  for (int t = 0; t < n_a; ++t)
    for (int r = 0; r < TM; ++r) {
      const float* p = &xs[(size_t)(t * TM + r) * 3];
      const float u = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
      float kv[TK] = {0};
      for (int c = 0; c < 3; ++c) { kv[c] = tf32_hi(p[c]); kv[3 + c] = tf32_hi(p[c] - kv[c]); kv[6 + c] = kv[c]; }
      kv[9] = tf32_hi(u); kv[10] = tf32_hi(u - kv[9]); kv[11] = 1.f; kv[12] = 1.f;
      put(ha, (size_t)t * TM * TK, TM, r, kv);
    }
  for (int t = 0; t < n_b; ++t)
    for (int r = 0; r < TN; ++r) {
      const float* p = &ys[(size_t)(t * TN + r) * 3];
      const float w = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
      float kv[TK] = {0};
      for (int c = 0; c < 3; ++c) { const float q = -2.f * p[c]; kv[c] = tf32_hi(q); kv[3 + c] = kv[c]; kv[6 + c] = tf32_hi(q - kv[c]); }
      kv[9] = 1.f; kv[10] = 1.f; kv[11] = tf32_hi(w); kv[12] = tf32_hi(w - kv[11]);
      put(hb, (size_t)t * TN * TK, TN, r, kv);
    }
  float *da, *db, *dout;
  long long* dcyc;
  CK(cudaMalloc(&da, ha.size() * 4)); CK(cudaMalloc(&db, hb.size() * 4));
  CK(cudaMalloc(&dout, (size_t)sms * TM * 6 * 4)); CK(cudaMalloc(&dcyc, sms * 8));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = A_BYTES + (size_t)NSTAGE * B_BYTES + 256;
  CK(cudaFuncSetAttribute(tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  // ---- correctness of the MMA + layout: 64 tiles (one pass over all columns) against fp64 on the host ----
  {
    Args g{da, db, n_b, n_b, 1, 2, dout, dcyc};
    tc_filter_kernel<<<1, 288, smem>>>(g);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(TM * 6);
    CK(cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_rel = 0;
    int bad_tile = 0;
    for (int r = 0; r < TM; ++r) {
      double best = 1e300; int bj = -1;
      for (int j = 0; j < n_b * TN; ++j) {
        double d = 0;
        for (int c = 0; c < 3; ++c) { const double e = (double)xs[r * 3 + c] - (double)ys[(size_t)j * 3 + c]; d += e * e; }
        if (d < best) { best = d; bj = j; }
      }
      const int hsel = o[r * 6] <= o[r * 6 + 3] ? 0 : 1;          // merge the two column halves
      const double err = fabs((double)o[r * 6 + hsel * 3] - best);
      const double scale = xs[r*3]*xs[r*3]+xs[r*3+1]*xs[r*3+1]+xs[r*3+2]*xs[r*3+2] + 3.0;
      if (err > max_err) max_err = err;
      if (err / scale > max_rel) max_rel = err / scale;
      int t1; memcpy(&t1, &o[r * 6 + hsel * 3 + 1], 4);
      if (t1 != bj / 64) ++bad_tile;     // CTA 0 starts at B tile 0: tile index = column / 64
    }
    printf("{\"test\": \"correctness: min_j D(row, j) over 16384 columns, 3xTF32 K=16 tcgen05.mma vs fp64\", \"max_abs_err\": %.3e, "
           "\"max_err_rel_to_u_plus_w\": %.3e, \"rows_with_wrong_tile\": %d}\n", max_err, max_rel, bad_tile);
  }
  // ---- timing ----
  const int tiles = 2048;
  const char* epi_name[3] = {"MMA only (no epilogue)", "MMA + tcgen05.ld of the whole accumulator", "MMA + tcgen05.ld + row minima (filter row direction)"};
  for (int stream_b = 0; stream_b < 2; ++stream_b)
    for (int epi = 0; epi < 3; ++epi) {
      Args g{da, db, n_b, tiles, stream_b, epi, dout, dcyc};
      tc_filter_kernel<<<sms, 288, smem>>>(g);
      CK(cudaDeviceSynchronize());
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      tc_filter_kernel<<<sms, 288, smem>>>(g);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<long long> cyc(sms);
      CK(cudaMemcpy(cyc.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (auto c : cyc) avg += (double)c / sms;
      const double pairs_per_sm = (double)tiles * TM * TN;
      // one direction = one tile pass; cycles per pair and direction per SM sub-partition lane, comparable with the FP32
      // kernel's cycles per pair (which serve both directions): x 128 lanes / pairs
      const double cyc_per_pair_lane = avg * 128.0 / pairs_per_sm;
      printf("{\"test\": \"%s, B %s\", \"tiles_per_cta\": %d, \"ms\": %.4f, \"cycles_per_tile\": %.1f, "
             "\"cycles_per_pair_one_direction\": %.3f, \"cycles_per_pair_both_directions\": %.3f, \"G_pairs_per_s_both_directions\": %.0f}\n",
             epi_name[epi], stream_b ? "streamed (8 x 16 KB bulk-copy ring from L2)" : "resident in shared memory", tiles, ms,
             avg / tiles, cyc_per_pair_lane, 2 * cyc_per_pair_lane, (double)sms * pairs_per_sm / 2 / (ms * 1e-3) / 1e9);
    }
  return 0;
}
