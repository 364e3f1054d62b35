// Chamfer nearest-neighbour search, "filter + exact resolve" path (B200, sm_100a).  Same results as the exact path in
// chamfer.cu (fp32 difference-form distances, first-index ties = torch.min semantics of model/loss.py:68-69), at
// 6 instead of 8 issue cycles per point pair.  DESIGN.md §4.
//
//   search   every pair is evaluated ONCE in the EXPANDED form the reference itself uses (loss.py:54-57), as packed
//            fp32x2 over two columns with the row value as broadcast scalar (FFMA2 acc, Qpair, x_row.F32, acc):
//              f_ij = |y_j|^2 - 2 x_i.y_j   (3 FFMA, chain seeded with |y_j|^2)         -> row direction  min_j f_ij
//              g_ij = f_ij + |x_i|^2        (1 FADD)                                    -> column direction min_i g_ij
//            The expanded form cancels, so f/g are only a FILTER.  Per row the minimum of f over each 64-column tile
//            is folded into a running (smallest tile minimum, its tile, second smallest) per search segment; per
//            column and 512-row block the two smallest 16-row-group minima are written (group id packed into the 5
//            low mantissa bits).  No atomics, no index bookkeeping in the inner loop.
//   resolve  per row: every tile whose filter minimum lies within a rigorous rounding-error guard G of the smallest
//            one is a candidate; its 64 columns are re-evaluated in the exact difference form and the (distance,
//            index)-lexicographic minimum is kept (more than one candidate tile: every segment whose smallest entry
//            is within G is re-evaluated whole).  Same per column with 16-row groups.  The candidate set provably
//            contains every exact minimiser (guard derivation below), so the result is the exact path's bit for bit.
//
// Guard.  eps = 2^-24, a = |x_i|, b = |y_j|, D = |x_i - y_j|^2, all on the fp32 inputs.
//   w_j = fl(|y_j|^2), u_i = fl(|x_i|^2): three roundings each, relative error <= 3.01 eps.
//   f   = fma(x0,q0, fma(x1,q1, fma(x2,q2, w))), q = -2y exact: every partial sum is bounded by b^2 + 2ab, so
//         |f - (|y|^2 - 2x.y)| <= 3.01 eps b^2 + 3 eps (b^2 + 2ab) <= 6.1 eps (b^2 + ab) =: e(a,b).
//   g   = fl(f + u): |g - D| <= e(a,b) + 3.01 eps a^2 + 1.01 eps D, and the 5 packed bits cost 2^-18 |g| more.
//   The exact path returns d = fl-difference-form with |d - D| <= 5.01 eps D.  If jf minimises f and j* minimises d,
//   then D_j* <= D_jf (1 + 10.1 eps), hence f_j* <= f_jf + 2 e(a, beta) + 10.1 eps D_jf with beta >= b_jf, b_j*;
//   beta = a + sqrt(1.0001 D_jf) by the triangle inequality, and D_jf <= 1.001 (f_jf + u_i)^+ + 1e-5 u_i.
//   guard_row / guard_col below evaluate these bounds with >= 1% slack plus an absolute 1e-30 for underflow.
#include <atomic>

#include "chamfer_common.cuh"
#include "../../include/rma_b200.h"

namespace rma {

constexpr int kFRowsPerThread = 16;
constexpr int kFRowsPerWarp = 32 * kFRowsPerThread;     // 512: one row block
constexpr int kFWarps = 4;                              // warps per search CTA for clouds of more than 1024 rows: a CTA row
                                                        // block is 2048 rows.  Smaller clouds run 2- or 1-warp CTAs
                                                        // (template parameter kW; 1024 / 512 rows per CTA row block), so
                                                        // that no warp of a CTA is left without rows (64 x 1024 ran at
                                                        // 33 % of peak with two of four warps idle)
constexpr int kFThreads = 32 * kFWarps;                 // threads of the largest search CTA (reporting)
static inline int f_warps_for(int nRB) { return nRB >= 3 ? 4 : nRB; }   // 512-row blocks per batch element -> kW
constexpr int kFColTile = 32;                           // columns per column-direction tile (= lanes)
constexpr int kFColTop = 2;                             // packed group minima recorded per column and row block
constexpr int kFRowTile = 64;                           // columns per work unit (CTA row block x 64 columns)
// Columns per row-direction tile minimum ("sub-tile", template parameter kSub of the search and the resolve): what the
// resolve re-evaluates per row.  32 instead of 64 cuts the resolve by a quarter (C2: 13.5 -> 10.3 us, C4: 160 -> 129 us,
// 16 x 2048: 10.1 -> 7.6 us) for twice the 16-row epilogues in the search (+2-3 %: C2 +1.4 us, C4 +10 us): a net gain
// while the resolve (linear in N + M) weighs enough against the search (N * M), i.e. for small clouds.
constexpr int kFSubSmall = 32, kFSubLarge = 64;
#ifndef RMA_F_SUBSWITCH
#define RMA_F_SUBSWITCH 4096
#endif
constexpr int kFSubSwitch = RMA_F_SUBSWITCH;            // min(N, M) <= this: 32-column sub-tiles (tools/ build A/B variants)
static inline int f_sub_for(int N, int M) { return (N < M ? N : M) <= kFSubSwitch ? kFSubSmall : kFSubLarge; }
constexpr int kFSubTiles = 16;                          // 64-column tiles per row-direction entry (sub-segment): bounds
                                                        // the slow path's re-evaluation to 1024 columns per candidate
#ifndef RMA_F_CHUNK
#define RMA_F_CHUNK 1024
#endif
#ifndef RMA_F_MINCTAS
#define RMA_F_MINCTAS 2
#endif
constexpr int kFChunk = RMA_F_CHUNK;                     // columns staged in shared memory at a time by a 4-warp CTA
                                                         // (kW warps: kFChunk * kW / 4, the same bytes per SM)
constexpr int kFCbStride = 36;
constexpr float kFBig = 1e30f;                          // |x|^2 of padded rows / |y|^2 of padded columns
constexpr float kEps = 5.9604644775390625e-8f;          // 2^-24

struct FGeometry {
  int B, N, M;
  int nRB;    // 512-row blocks per batch element
  int Npad;   // nRB * 512
  int W;      // warps per search CTA (4, or 2 / 1 for clouds of <= 1024 / 512 rows)
  int RC;     // rows per CTA row block: 512 * W
  int nRBC;   // CTA row blocks (RC rows) per batch element
  int nT;     // 64-column tiles per batch element
  int Mp;     // nT * 64
  int U;      // work units (CTA row block x 64-column tile) per search CTA
  int Smax;   // max segments (CTAs) that share one CTA row block
  int P;      // row-direction entries per segment: ceil(U / kFSubTiles)
  long long units() const { return (long long)B * nRBC * nT; }
};

// Work decomposition of the search: the B*nRBC*nT units are dealt out in contiguous runs of U to the CTAs, so every
// CTA has the same amount of work (+-1 unit) whatever B, N, M are; a run may cross into the next row block / batch
// element, in which case the CTA reloads its rows ("segment").  U is chosen for a whole number of balanced waves.
static int f_choose_unit_run(long long units, int nT, int slots) {
  // cost of w waves ~ w * (U * 64 columns + fixed per-CTA overhead of ~40 column-times), U = ceil(units/(slots*w))
  double best_cost = 1e300;
  int best = 1;
  for (int w = 1; w <= 64; ++w) {
    long long U = (units + (long long)slots * w - 1) / ((long long)slots * w);
    if (U < 1) U = 1;
    const double cost = (double)w * ((double)U * kFRowTile + 40.0);
    if (cost < best_cost * 0.995) { best_cost = cost; best = (int)(U > 0x3fffffff ? 0x3fffffff : U); }
    if (U <= 2) break;
  }
  (void)nT;
  return best;
}

static FGeometry f_make_geometry(int B, int N, int M, int slots, int force_run) {
  FGeometry g;
  g.B = B; g.N = N; g.M = M;
  g.nRB = (N + kFRowsPerWarp - 1) / kFRowsPerWarp;
  g.Npad = g.nRB * kFRowsPerWarp;
  g.W = f_warps_for(g.nRB);
  g.RC = g.W * kFRowsPerWarp;
  g.nRBC = (g.nRB + g.W - 1) / g.W;
  g.nT = (M + kFRowTile - 1) / kFRowTile;
  g.Mp = g.nT * kFRowTile;
  g.U = force_run > 0 ? force_run : f_choose_unit_run(g.units(), g.nT, slots);
  g.Smax = (g.nT + g.U - 1) / g.U + 1;
  g.P = ((g.U < g.nT ? g.U : g.nT) + kFSubTiles - 1) / kFSubTiles;
  return g;
}

struct FWorkspace {
  float4* rowpack;   // [B][Npad]               (x, y, z, |x|^2); padded rows (0,0,0,1e30)           (resolve)
  float4* rowsoa;    // [B][nRB][16][32]        per search thread its 16 rows as x[16], y[16], z[16], u[16],
                     //                         lane-coalesced: 16 LDG.128 whose halves are the packed row pairs
  float4* colpack;   // [B][Mp]                 (-2x, -2y, -2z, |y|^2); padded columns (0,0,0,1e30)            (resolve)
  float4* colpair;   // [B][Mp/2][2]            the same, two columns (c, c+1) interleaved as (q0c, q0c', q1c, q1c')
                     //                         (q2c, q2c', wc, wc'): the search's fp32x2 operands as they are read from
                     //                         shared memory, so a chunk is one contiguous TMA bulk copy
  float* rowtop;     // [B*nRBC][Smax*P][3][RC] per segment and run of <= 16 tiles in it ("entry"): smallest sub-tile
                     //                         minimum, its sub-tile, second smallest; unused entries of a segment (inf, 0, inf)
  float* colrec;     // [B][nRB][2][Mp]         two smallest packed group minima of g over the row block
  unsigned* stats;   // [4] rows on the slow path, columns on the slow path, row-block rescans, unused
  size_t bytes;
};

static FWorkspace f_carve(const FGeometry& g, void* base) {
  FWorkspace w;
  char* p = (char*)base;
  size_t off = 0;
  w.rowpack = (float4*)(p + off); off += align256((size_t)g.B * g.Npad * sizeof(float4));
  w.rowsoa = (float4*)(p + off);  off += align256((size_t)g.B * g.Npad * sizeof(float4));
  w.colpack = (float4*)(p + off); off += align256((size_t)g.B * g.Mp * sizeof(float4));
  w.colpair = (float4*)(p + off); off += align256((size_t)g.B * g.Mp * sizeof(float4));
  w.rowtop = (float*)(p + off);   off += align256((size_t)g.B * g.nRBC * g.Smax * g.P * 3 * g.RC * sizeof(float));
  w.colrec = (float*)(p + off);   off += align256((size_t)g.B * g.nRB * kFColTop * g.Mp * sizeof(float));
  w.stats = (unsigned*)(p + off); off += 256;
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float norm2_rn(float x, float y, float z) {
  return __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
}

__global__ void __launch_bounds__(256) chamfer_fprep_kernel(
    const float* __restrict__ x, long long xsb, long long xsn, long long xsc,
    const float* __restrict__ y, long long ysb, long long ysn, long long ysc,
    FGeometry g, FWorkspace w, float* __restrict__ zero_loss, float4* __restrict__ zero_gx,
    float4* __restrict__ zero_gy) {
  const long long nrow = (long long)g.B * g.Npad;
  const long long ncol = (long long)g.B * g.Mp;
  pdl_wait();
  pdl_launch_dependents();   // let the search grid's launch overlap this kernel; its CTAs wait for our completion
  if (blockIdx.x == 0 && threadIdx.x < 4) w.stats[threadIdx.x] = 0u;
  if (zero_loss && blockIdx.x == 0 && threadIdx.x == 0) *zero_loss = 0.f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nrow + ncol;
       t += (long long)gridDim.x * blockDim.x) {
    if (t < nrow) {
      const int b = (int)(t / g.Npad), i = (int)(t % g.Npad);
      float4 v = make_float4(0.f, 0.f, 0.f, kFBig);
      if (i < g.N) {
        const float* s = x + b * xsb + i * xsn;
        v.x = s[0]; v.y = s[xsc]; v.z = s[2 * xsc];
        v.w = norm2_rn(v.x, v.y, v.z);
        if (zero_gx) zero_gx[(long long)b * g.N + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      w.rowpack[t] = v;
      const int rb = i / kFRowsPerWarp, ln = (i % kFRowsPerWarp) / kFRowsPerThread, r = i % kFRowsPerThread;
      float* soa = reinterpret_cast<float*>(w.rowsoa + ((size_t)(b * g.nRB + rb) * 16) * 32 + ln);
      // float4 slot c*4 + r/4 (c = x,y,z,u), element r%4; slots are 32 float4 apart
      soa[(size_t)(r >> 2) * 128 + (r & 3)] = v.x;
      soa[(size_t)(4 + (r >> 2)) * 128 + (r & 3)] = v.y;
      soa[(size_t)(8 + (r >> 2)) * 128 + (r & 3)] = v.z;
      soa[(size_t)(12 + (r >> 2)) * 128 + (r & 3)] = v.w;
    } else {
      const long long u = t - nrow;
      const int b = (int)(u / g.Mp), j = (int)(u % g.Mp);
      float4 v = make_float4(0.f, 0.f, 0.f, kFBig);
      if (j < g.M) {
        const float* s = y + b * ysb + j * ysn;
        const float px = s[0], py = s[ysc], pz = s[2 * ysc];
        v = make_float4(-2.f * px, -2.f * py, -2.f * pz, norm2_rn(px, py, pz));
        if (zero_gy) zero_gy[(long long)b * g.M + j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      w.colpack[u] = v;
      // pair pack: columns (c, c+1) sit in neighbouring lanes (nrow, Mp and the grid stride are even, so the pair
      // never straddles a warp, the row/column boundary or the loop bound); the even lane stores (q0c, q0c', q1c, q1c'),
      // the odd one (q2c, q2c', wc, wc') - one coalesced 16-byte store each
      const bool odd = (u & 1) != 0;
      const float s0 = odd ? v.x : v.z, s1 = odd ? v.y : v.w;            // what the partner lane needs from me
      const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
      w.colpair[u] = odd ? make_float4(r0, v.z, r1, v.w) : make_float4(v.x, r0, v.y, r1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------------------------
// Insert v into the ascending pair (p1 <= p2): 3 FMNMX, no predicates.  (Two entries per row block suffice: a second
// group within the guard - ~5e-5 of the columns at C2 - makes the resolve re-evaluate the whole 512-row block.)
__device__ __forceinline__ void top2_insert(float& p1, float& p2, float v) {
  p2 = fminf(p2, fmaxf(p1, v));
  p1 = fminf(p1, v);
}

// dynamic shared memory: column double buffer, per-warp transpose buffers, per-row running top entries, two mbarriers
__host__ __device__ constexpr size_t f_smem_base(int kW) {
  return (size_t)2 * (kFChunk * kW / 4 + 4) * sizeof(float4) + (size_t)kW * 2 * kFColTile * kFCbStride * sizeof(float);
}
__host__ __device__ constexpr size_t f_smem_top(int kW) { return f_smem_base(kW) + (size_t)3 * 4 * (32 * kW) * sizeof(float4); }
__host__ __device__ constexpr size_t f_smem(int kW) { return f_smem_top(kW) + 16; }

struct FSearchArgs {
  const float4* rowsoa;
  const float4* colpair;
  float* rowtop;
  float* colrec;
  int N, M, nRB, Npad, nRBC, nT, Mp, U, Smax, P;
  long long units;
};

template <int kSub, int kW>
__global__ void __launch_bounds__(32 * kW, RMA_F_MINCTAS * 4 / kW) chamfer_fsearch_kernel(FSearchArgs a) {
  static_assert(kSub == kFRowTile || kSub == kFColTile, "row sub-tile is one or two 32-column passes");
  static_assert(kW == 1 || kW == 2 || kW == 4, "1, 2 or 4 warps per search CTA");
  constexpr int kThreads = 32 * kW, kRowsPerCta = kFRowsPerWarp * kW, kChunk = kFChunk * kW / 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* scol = reinterpret_cast<float4*>(smem_raw);                                              // [2][kChunk + 4]
  float* cball = reinterpret_cast<float*>(smem_raw + (size_t)2 * (kChunk + 4) * sizeof(float4));    // [warps][2][32][36]
  u64* mbar = reinterpret_cast<u64*>(smem_raw + f_smem_top(kW));                                   // [2]
  // per-row running (smallest tile minimum, its tile, second smallest): touched once per 64-column tile, so it lives
  // in shared memory ([plane][q][thread] float4, conflict-free) instead of 48 registers
  float4* stop = reinterpret_cast<float4*>(smem_raw + f_smem_base(kW)) + threadIdx.x;             // [3][4][kThreads]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* cb = cball + warp * (2 * kFColTile * kFCbStride);
  const float inf = __int_as_float(0x7f800000);
  const unsigned lane_tag = (unsigned)lane;
  unsigned tag_mask = ~31u;                                    // kept in a register: (bits & mask) | tag is then ONE LOP3
  asm volatile("" : "+r"(tag_mask));

  long long unit = (long long)blockIdx.x * a.U;
  const long long unit_end = min(unit + a.U, a.units);
  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
  __syncthreads();
  int cur = 0;           // column buffer the next chunk lands in
  unsigned phases = 0;   // bit s: parity the next wait on buffer s expects
  pdl_wait();
  pdl_launch_dependents();   // the resolve grid may be launched (it blocks in its own pdl_wait until we are done)

  while (unit < unit_end) {
    // ---- one segment: a run of `ntile` 64-column tiles of one CTA row block ----
    const int brb = (int)(unit / a.nT);
    const int tile0 = (int)(unit - (long long)brb * a.nT);
    const int ntile = (int)min((long long)(a.nT - tile0), unit_end - unit);
    unit += ntile;
    const int b = brb / a.nRBC, rbc = brb - b * a.nRBC;
    const int rb = rbc * kW + warp;
    const bool has_rows = rb < a.nRB;   // warp-uniform
    const int seg = blockIdx.x - (int)(((long long)brb * a.nT) / a.U);   // which of the row block's segments
    const int c0 = tile0 * kFRowTile;
    const int ncols = ntile * kFRowTile;

    // Column chunks arrive by TMA bulk copy (one contiguous run of the interleaved pair pack) into a double buffer:
    // the segment's first chunk is in flight while the rows are loaded, chunk k+1 while chunk k is computed.  Both
    // buffers are free here: every chunk iteration ends with a __syncthreads.
    const float4* cp = a.colpair + ((size_t)b * a.Mp + c0);
    if (tid == 0) bulk_load(scol + cur * (kChunk + 4), cp, (unsigned)min(kChunk, ncols) * 16u, &mbar[cur]);

    u64 px[8], py[8], pz[8], pu[8];
    if (has_rows) {
      const float4* rp = a.rowsoa + ((size_t)(b * a.nRB + rb) * 16) * 32 + lane;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 vx = rp[q * 32], vy = rp[(4 + q) * 32], vz = rp[(8 + q) * 32], vu = rp[(12 + q) * 32];
        px[2 * q] = pack2(vx.x, vx.y); px[2 * q + 1] = pack2(vx.z, vx.w);
        py[2 * q] = pack2(vy.x, vy.y); py[2 * q + 1] = pack2(vy.z, vy.w);
        pz[2 * q] = pack2(vz.x, vz.y); pz[2 * q + 1] = pack2(vz.z, vz.w);
        pu[2 * q] = pack2(vu.x, vu.y); pu[2 * q + 1] = pack2(vu.z, vu.w);
      }
    }
    // The running (smallest, its tile, second smallest) goes to the segment's next entry every kFSubTiles tiles and at
    // the segment's end, and starts again empty; entries the segment does not fill are written empty.
    float* top = a.rowtop + ((size_t)(brb * a.Smax + seg) * a.P * 3) * kRowsPerCta + warp * kFRowsPerWarp +
                 lane * kFRowsPerThread;
    int sub_tiles = 0, sub = 0;
    auto reset_top = [&]() {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        stop[q * kThreads] = make_float4(inf, inf, inf, inf);
        stop[(4 + q) * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
        stop[(8 + q) * kThreads] = make_float4(inf, inf, inf, inf);
      }
    };
    auto flush_top = [&]() {
      float* o = top + (size_t)sub * (3 * kRowsPerCta);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        reinterpret_cast<float4*>(o)[q] = stop[q * kThreads];
        reinterpret_cast<float4*>(o + kRowsPerCta)[q] = stop[(4 + q) * kThreads];          // tile ids (int bits)
        reinterpret_cast<float4*>(o + 2 * kRowsPerCta)[q] = stop[(8 + q) * kThreads];
      }
      ++sub;
      sub_tiles = 0;
      reset_top();
    };
    reset_top();

    float* crec = a.colrec + ((size_t)(b * a.nRB + rb) * kFColTop) * a.Mp + c0 + lane;

    // Column direction, software-pipelined: while tile t is computed, lane l folds column l of tile t-1 (32 per-lane
    // packed partial minima in the other half of the double buffer) into its running top-3, four values per body.
    int pend_col = -1;   // column offset (relative to c0) of the tile whose reduce is pending; warp-uniform
    int parity = 0;

    for (int cs = 0; cs < ncols; cs += kChunk) {
      const int nc = min(kChunk, ncols - cs);
      // two columns (c, c+1) lie interleaved as (q0c, q0c', q1c, q1c') (q2c, q2c', wc, wc'): an LDS.128 delivers two
      // ready-made fp32x2 operands of the column pair
      if (tid == 0 && cs + kChunk < ncols)
        bulk_load(scol + (cur ^ 1) * (kChunk + 4), cp + cs + kChunk,
                  (unsigned)min(kChunk, ncols - cs - kChunk) * 16u, &mbar[cur ^ 1]);
      mbar_wait(&mbar[cur], (phases >> cur) & 1u);
      phases ^= 1u << cur;
      const float4* schunk = scol + cur * (kChunk + 4);
      cur ^= 1;

#pragma unroll 1
      for (int t0 = 0; has_rows && t0 < nc; t0 += kFRowTile) {
        float tm[kFRowsPerThread];
#pragma unroll 1
        for (int h = 0; h < kFRowTile; h += kFColTile) {
          if (kSub == kFColTile || h == 0) {
#pragma unroll
            for (int r = 0; r < kFRowsPerThread; ++r) tm[r] = inf;
          }
          const float4* sp = schunk + t0 + h;
          float* cbw = cb + parity * (kFColTile * kFCbStride) + lane;
          const float4* cbr = reinterpret_cast<const float4*>(cb + (parity ^ 1) * (kFColTile * kFCbStride) +
                                                              lane * kFCbStride);
          float p1 = inf, p2 = inf;

          // Four columns against the 16 rows, plus four pending values of the previous tile's column reduce.
          auto body = [&](const float4& q0, const float4& q1, const float4& q2, const float4& q3, int jj) {
            const float4 pv = cbr[jj >> 2];          // 4 packed partial minima of the pending tile
            // fp32x2 lanes = two COLUMNS, the row value is the broadcast scalar.  The operand that stays fixed over
            // the 16-row inner loop is then the 64-bit column pair, which the operand-reuse cache keeps (2 register
            // words saved per FFMA2 instead of 1): FFMA2 acc, Qpair.reuse, x_row.F32, acc.
            float cm[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const float4 A = h2 ? q2 : q0, Bq = h2 ? q3 : q1;
              const u64 Q0 = pack2(A.x, A.y), Q1 = pack2(A.z, A.w), Q2 = pack2(Bq.x, Bq.y), WW = pack2(Bq.z, Bq.w);
              // chain step by step over the 16 rows (ptxas re-schedules this freely; forcing the order with
              // basic-block fences so that the 64-bit column operand is reused 16 times in a row was measured: no gain)
              u64 acc[16];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float za, zb;
                unpack2(pz[k], za, zb);
                acc[2 * k] = fma2(Q2, pack2(za, za), WW); acc[2 * k + 1] = fma2(Q2, pack2(zb, zb), WW);
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float ya, yb;
                unpack2(py[k], ya, yb);
                acc[2 * k] = fma2(Q1, pack2(ya, ya), acc[2 * k]); acc[2 * k + 1] = fma2(Q1, pack2(yb, yb), acc[2 * k + 1]);
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float xa, xb;
                unpack2(px[k], xa, xb);
                acc[2 * k] = fma2(Q0, pack2(xa, xa), acc[2 * k]); acc[2 * k + 1] = fma2(Q0, pack2(xb, xb), acc[2 * k + 1]);
              }
              float cl = inf, ch = inf;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float ua, ub, al, ah, bl, bh;
                unpack2(pu[k], ua, ub);
                unpack2(acc[2 * k], al, ah); unpack2(acc[2 * k + 1], bl, bh);
                tm[2 * k] = min3(tm[2 * k], al, ah);
                tm[2 * k + 1] = min3(tm[2 * k + 1], bl, bh);
                unpack2(add2(acc[2 * k], pack2(ua, ua)), al, ah); unpack2(add2(acc[2 * k + 1], pack2(ub, ub)), bl, bh);
                cl = min3(cl, al, bl);
                ch = min3(ch, ah, bh);
              }
              cm[2 * h2] = cl; cm[2 * h2 + 1] = ch;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              cbw[(jj + c) * kFCbStride] = __uint_as_float((__float_as_uint(cm[c]) & tag_mask) | lane_tag);
            top2_insert(p1, p2, pv.x);
            top2_insert(p1, p2, pv.y);
            top2_insert(p1, p2, pv.z);
            top2_insert(p1, p2, pv.w);
          };

          // Operands are prefetched one body ahead into the OTHER register set (ping-pong, unrolled by two) so that no
          // register copies are needed: a MOV/IMAD.MOV here queues behind the saturated FMA pipe and stalls the warp.
          // The last prefetch of a chunk reads the 4-column pad behind the staged data and is never used.
          float4 a0 = sp[0], a1 = sp[1], a2 = sp[2], a3 = sp[3];
#pragma unroll 1
          for (int jj = 0; jj < kFColTile; jj += 8) {
            const float4 b0 = sp[jj + 4], b1 = sp[jj + 5], b2 = sp[jj + 6], b3 = sp[jj + 7];
            body(a0, a1, a2, a3, jj);
            a0 = sp[jj + 8]; a1 = sp[jj + 9]; a2 = sp[jj + 10]; a3 = sp[jj + 11];
            body(b0, b1, b2, b3, jj + 4);
          }

          if (pend_col >= 0) {
            float* o = crec + pend_col;
            o[0] = p1; o[a.Mp] = p2;
          }
          pend_col = cs + t0 + h;
          parity ^= 1;
          __syncwarp();   // this tile's partial minima visible to all lanes; previous buffer free for reuse

          // Row direction: the sub-tile's 16 filter minima are folded into the running (smallest, its sub-tile, second
          // smallest) of the segment.  (A per-tile record in global memory - B*N*M/16 bytes, read only when a guard
          // fails - cost 2.8 us of the 83 us C2 step in stores that evict dirty L2 lines; the slow path re-evaluates
          // the candidate SEGMENTS instead.)
          if (kSub == kFColTile || h + kFColTile == kFRowTile) {
            const int tile = tile0 * (kFRowTile / kSub) + (cs + t0 + h) / kSub;
            const float tilef = __int_as_float(tile);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 m1 = stop[q * kThreads], t1 = stop[(4 + q) * kThreads], m2 = stop[(8 + q) * kThreads];
              m2.x = fminf(m2.x, fmaxf(m1.x, tm[4 * q]));     t1.x = tm[4 * q] < m1.x ? tilef : t1.x;
              m2.y = fminf(m2.y, fmaxf(m1.y, tm[4 * q + 1])); t1.y = tm[4 * q + 1] < m1.y ? tilef : t1.y;
              m2.z = fminf(m2.z, fmaxf(m1.z, tm[4 * q + 2])); t1.z = tm[4 * q + 2] < m1.z ? tilef : t1.z;
              m2.w = fminf(m2.w, fmaxf(m1.w, tm[4 * q + 3])); t1.w = tm[4 * q + 3] < m1.w ? tilef : t1.w;
              m1.x = fminf(m1.x, tm[4 * q]); m1.y = fminf(m1.y, tm[4 * q + 1]);
              m1.z = fminf(m1.z, tm[4 * q + 2]); m1.w = fminf(m1.w, tm[4 * q + 3]);
              stop[q * kThreads] = m1; stop[(4 + q) * kThreads] = t1; stop[(8 + q) * kThreads] = m2;
            }
          }
        }
        if (++sub_tiles == kFSubTiles) flush_top();
      }
      __syncthreads();   // chunk fully consumed: a later bulk copy may refill this buffer
    }

    if (has_rows) {
      if (pend_col >= 0) {   // drain: column reduce of the segment's last tile
        const float* col = cb + (parity ^ 1) * (kFColTile * kFCbStride) + lane * kFCbStride;
        float p1 = inf, p2 = inf;
#pragma unroll
        for (int k = 0; k < 32; ++k) top2_insert(p1, p2, col[k]);
        float* o = crec + pend_col;
        o[0] = p1; o[a.Mp] = p2;
      }
      if (sub_tiles > 0) flush_top();
      while (sub < a.P) flush_top();   // empty entries
      __syncwarp();   // the next segment reuses the transpose buffers
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// resolve
// ---------------------------------------------------------------------------------------------------------------
// Rigorous guards (derivation in the file header).  u = fl(|x_i|^2), mt = smallest filter value f of the row.
__device__ __forceinline__ float guard_row(float u, float mt) {
  const float D0 = fmaxf(mt + u, 0.f);
  const float Dub = 1.001f * D0 + 1e-5f * u;
  const float a = sqrtf(u) * 1.0001f;
  const float beta = a + sqrtf(Dub) * 1.0001f;
  return (12.3f * beta * (beta + a) + 10.3f * Dub) * (kEps * 1.01f) + 1e-30f;
}
// wj = fl(|y_j|^2), mt = smallest (packed, truncated) value of g for the column, i.e. an estimate of the distance.
__device__ __forceinline__ float guard_col(float wj, float mt) {
  const float D0 = fmaxf(mt, 0.f);
  const float Dub = 1.001f * D0 + 1e-5f * wj;
  const float bn = sqrtf(wj) * 1.0001f;
  const float alpha = bn + sqrtf(Dub) * 1.0001f;
  return (12.3f * bn * (bn + alpha) + 6.3f * alpha * alpha + 12.5f * Dub) * (kEps * 1.01f) +
         Dub * 7.8e-6f /* 2^-17 * 1.02 */ + 1e-30f;
}

// Exact difference-form distance between (x,y,z) and the column stored as q = -2*(yx,yy,yz): x + 0.5 q == x - y
// exactly, so this reproduces sqdist_neg / the oracle's operation order bit for bit.
__device__ __forceinline__ float sqdist_q(float x, float y, float z, const float4 q) {
  const float dx = __fmaf_rn(0.5f, q.x, x), dy = __fmaf_rn(0.5f, q.y, y), dz = __fmaf_rn(0.5f, q.z, z);
  float d = __fmul_rn(dx, dx);
  d = __fmaf_rn(dy, dy, d);
  d = __fmaf_rn(dz, dz, d);
  return d;
}

// (distance bits, index) lexicographic minimum over the kLanes lanes of an aligned lane group.
template <int kLanes>
__device__ __forceinline__ void group_min(unsigned& d, unsigned& idx) {
#pragma unroll
  for (int o = 1; o < kLanes; o <<= 1) {
    const unsigned od = __shfl_xor_sync(0xffffffffu, d, o);
    const unsigned oi = __shfl_xor_sync(0xffffffffu, idx, o);
    const bool take = od < d || (od == d && oi < idx);
    d = take ? od : d;
    idx = take ? oi : idx;
  }
}

// Result of one fast-path round for owner lane (round * kItems + k): the (distance, index) minimum of lane group k
// (kLanes lanes per item, kItems = 32 / kLanes items per round).  Usually exactly one lane of a group holds a
// survivor, then a ballot finds it and two SHFL fetch it; otherwise the butterfly runs.  Warp-uniform control flow.
template <int kLanes>
__device__ __forceinline__ void round_result(unsigned d, unsigned idx, unsigned& rd, unsigned& ri) {
  constexpr unsigned kItems = 32u / kLanes, gmask = (1u << kLanes) - 1u;
  const unsigned full = 0xffffffffu, lane = lane_id();
  const unsigned hit = __ballot_sync(full, d != 0xffffffffu);
  const unsigned gb = (hit >> (lane & ~(unsigned)(kLanes - 1))) & gmask;   // my own group's survivors
  if (__all_sync(full, (gb & (gb - 1u)) == 0u)) {                            // every group has at most one
    const unsigned k = lane % kItems, gk = (hit >> (k * kLanes)) & gmask;
    const int src = (int)(k * kLanes) + (gk ? __ffs(gk) - 1 : 0);
    rd = __shfl_sync(full, d, src); ri = __shfl_sync(full, idx, src);
  } else {
    group_min<kLanes>(d, idx);
    rd = __shfl_sync(full, d, (lane % kItems) * kLanes); ri = __shfl_sync(full, idx, (lane % kItems) * kLanes);
  }
}

// Work distribution: the ceil(B*N/32) + ceil(B*M/32) warp items are dealt out evenly over the grid (CTA c takes items
// [c*total/grid, (c+1)*total/grid), at most one per warp).  For a single-wave problem the grid is exactly
// SMs x 4 resident CTAs, so every SM carries the same load wherever the block scheduler (or an early programmatic
// launch behind the search's tail) places the CTAs: with one CTA per 8 items, 512 CTAs on 148 SMs leave some SMs with
// 4 x 8 items against an average of 27.7.
//
// Lanes per item in the fast path: kSub / 8 lanes re-evaluate one row's candidate sub-tile (8 columns each) or one
// column's candidate group of 16 rows; a round handles kItems = 32 / kLanes items.
template <int kSub>
__global__ void __launch_bounds__(kFinWarps * 32, 4) chamfer_fresolve_kernel(FGeometry g, FWorkspace w, FinalizeArgs f) {
  constexpr int kLanes = kSub / 8, kItems = 32 / kLanes, kRounds = 32 / kItems;
  constexpr int kGroupRows = kFRowsPerThread / kLanes;      // rows of a 16-row group per lane
  __shared__ double lsum[kFinWarps];
  // per-warp staging of the 32 items' broadcast data: a round reads its 4 items with two LDS.128 (one address per
  // 8-lane group, conflict-free) instead of seven SHFL
  __shared__ float4 sA[kFinWarps][32];
  __shared__ int4 sB[kFinWarps][32];
  // 32-bit index math throughout: B*Npad, B*Mp and the unit count are < 2^31 (f_sizes_ok); a 64-bit division
  // costs ~80 emulated instructions per lane and there were five of them
  const unsigned nrow = (unsigned)g.B * (unsigned)g.N, ncol = (unsigned)g.B * (unsigned)g.M;
  const unsigned row_warps = (nrow + 31u) / 32u;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int grp = lane / kLanes, sub = lane % kLanes;
  const unsigned total_warps = row_warps + (ncol + 31u) / 32u;
  const unsigned w_begin = (unsigned)(((u64)blockIdx.x * total_warps) / gridDim.x);
  const unsigned w_end = (unsigned)(((u64)(blockIdx.x + 1) * total_warps) / gridDim.x);
  // warps beyond the CTA's share take the out-of-range item (valid == false on every lane)
  const unsigned wi = w_begin + (unsigned)warp < w_end ? w_begin + (unsigned)warp : total_warps;
  const float coef_grad = 2.f * f.coef_loss;
  const unsigned full = 0xffffffffu;
  const float inf = __int_as_float(0x7f800000);
  float myd = 0.f, wb = 1.f;
  bool valid = false;
  pdl_wait();
  pdl_launch_dependents();   // the backward combine may queue behind us (it waits for our completion itself)

  if (wi < row_warps) {
    // ---------------------------------------------------------------- rows: lane <-> row, candidates = 64-column tiles
    const unsigned t = wi * 32u + lane;
    valid = t < nrow;
    int b = 0, i = 0;
    float4 rp = make_float4(0.f, 0.f, 0.f, 0.f);
    float M1 = inf, M2 = inf;
    int T1 = 0;
    if (valid) {
      b = (int)(t / (unsigned)g.N); i = (int)(t - (unsigned)b * (unsigned)g.N);
      rp = w.rowpack[(size_t)b * g.Npad + i];
      const int rbc = i / g.RC, brb = b * g.nRBC + rbc;
      const int kfirst = (int)(((unsigned)brb * (unsigned)g.nT) / (unsigned)g.U);
      const int klast = (int)((((unsigned)brb + 1u) * (unsigned)g.nT - 1u) / (unsigned)g.U);
      const float* top = w.rowtop + ((size_t)brb * g.Smax * g.P * 3) * g.RC + (i - rbc * g.RC);
      const int nseg = (klast - kfirst + 1) * g.P;  // entries: P per segment, consecutive
      constexpr int kSegBatch = 12;                 // up to 12 entries in flight: one latency for C2 (10)
      for (int s0 = 0; s0 < nseg; s0 += kSegBatch) {
        float a1[kSegBatch], a2[kSegBatch];
        int at[kSegBatch];
#pragma unroll
        for (int k = 0; k < kSegBatch; ++k) {
          const bool on = s0 + k < nseg;
          const float* tp = top + (size_t)(on ? s0 + k : s0) * (3 * g.RC);
          a1[k] = on ? tp[0] : inf;
          at[k] = __float_as_int(tp[g.RC]);
          a2[k] = on ? tp[2 * g.RC] : inf;
        }
#pragma unroll
        for (int k = 0; k < kSegBatch; ++k) {
          M2 = fminf(fminf(M2, a2[k]), fmaxf(M1, a1[k]));
          if (a1[k] < M1) { M1 = a1[k]; T1 = at[k]; }
        }
      }
    }
    const float thr = M1 + guard_row(rp.w, M1);
#if defined(RMA_B200_EXPERIMENTS) && (defined(RMA_EXP_NO_SLOW) || defined(RMA_EXP_NO_SLOW_ROWS))   // timing experiment only (tools/resolve_variants.py): wrong results for the slow-path items
    const bool fast = valid;
#else
    const bool fast = valid && M2 > thr;
#endif
    unsigned best_d = 0xffffffffu, best_j = 0;

    // Fast path (one candidate sub-tile of kSub columns): kLanes lanes per row, 8 columns per lane, kItems rows per
    // round.  Inside the sub-tile the same guard applies per COLUMN: every exact minimiser has f_j <= thr, so each
    // column is first re-filtered with the search's own 3-FFMA chain (bit-identical f; padded columns carry
    // |y|^2 = 1e30 and never pass) and only the survivors - typically one per row - are evaluated in the exact
    // difference form; the pass bits are collected branch-free and the exact evaluations sit under one branch.
    sA[warp][lane] = make_float4(rp.x, rp.y, rp.z, thr);
    sB[warp][lane] = make_int4(b, T1, fast ? 1 : 0, 0);
    __syncwarp();
#pragma unroll 1
    for (int round = 0; round < kRounds; ++round) {
      const int src = round * kItems + grp;
      const float4 sa = sA[warp][src];
      const int4 sb4 = sB[warp][src];
      const float x = sa.x, y = sa.y, z = sa.z, sthr = sa.w;
      unsigned d = 0xffffffffu, dj = 0xffffffffu;
      if (sb4.z) {
        const int j0 = sb4.y * kSub + sub;
        const float4* cq = w.colpack + (size_t)sb4.x * g.Mp + j0;
        float4 q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = cq[kLanes * c];
        unsigned pass = 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float fj = __fmaf_rn(x, q[c].x, __fmaf_rn(y, q[c].y, __fmaf_rn(z, q[c].z, q[c].w)));
          pass |= fj <= sthr ? 1u << c : 0u;
        }
        while (pass) {                                     // ascending columns: strict '<' keeps the first
          const int c = __ffs(pass) - 1;
          pass &= pass - 1u;
          const unsigned dc = __float_as_uint(sqdist_q(x, y, z, cq[kLanes * c]));
          if (dc < d) { d = dc; dj = (unsigned)(j0 + kLanes * c); }
        }
      }
      unsigned rd, rj;
      round_result<kLanes>(d, dj, rd, rj);
      if (lane / kItems == round) { best_d = rd; best_j = rj; }
    }

    // Slow path (guard failed: several tiles within the rounding error of the smallest).  Every candidate tile lies in
    // an entry (a run of up to 16 tiles of one search segment) whose smallest tile minimum is within the guard, so the
    // warp re-evaluates those entries' tiles; the lowest column among the lanes that hold the minimum = torch.min's
    // first index.
    unsigned slow = __ballot_sync(full, valid && !fast);
    if (slow && lane == 0) atomicAdd(w.stats + 0, (unsigned)__popc(slow));
    while (slow) {
      const int src = __ffs(slow) - 1; slow &= slow - 1;
      const int sb = __shfl_sync(full, b, src), si = __shfl_sync(full, i, src);
      const float sthr = __shfl_sync(full, thr, src);
      const float x = __shfl_sync(full, rp.x, src), y = __shfl_sync(full, rp.y, src), z = __shfl_sync(full, rp.z, src);
      const int srbc = si / g.RC;
      const unsigned sbrb = (unsigned)(sb * g.nRBC + srbc);
      const unsigned u0 = sbrb * (unsigned)g.nT, u1 = u0 + (unsigned)g.nT;       // the row block's units [u0, u1)
      const unsigned kfirst = u0 / (unsigned)g.U, klast = (u1 - 1u) / (unsigned)g.U;
      const float* top = w.rowtop + ((size_t)sbrb * g.Smax * g.P * 3) * g.RC + (si - srbc * g.RC);
      const float4* cq = w.colpack + (size_t)sb * g.Mp;
      unsigned bd = 0xffffffffu, bj = 0xffffffffu;
      // One sub-tile for this lane's group: the fast path's re-filter (f_j <= thr, bit-identical to the search's
      // chain; padded columns never pass) in front of the exact evaluation, 8 loads in flight per lane.
      auto eval_tile = [&](unsigned tile) {
        const int j0 = (int)tile * kSub + sub;
        float4 q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = cq[j0 + kLanes * c];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float fj = __fmaf_rn(x, q[c].x, __fmaf_rn(y, q[c].y, __fmaf_rn(z, q[c].z, q[c].w)));
          if (fj <= sthr) {
            const unsigned d = __float_as_uint(sqdist_q(x, y, z, q[c])), jc = (unsigned)(j0 + kLanes * c);
            if (d < bd || (d == bd && jc < bj)) { bd = d; bj = jc; }
          }
        }
      };
      const unsigned nent = (klast - kfirst + 1u) * (unsigned)g.P;
      for (unsigned base = 0; base < nent; base += 32u) {
        // one entry per lane, all in flight; empty entries hold inf
        const unsigned e = base + (unsigned)lane;
        const float* tp = top + (size_t)(e < nent ? e : base) * (3 * g.RC);
        const float a1 = e < nent ? tp[0] : inf;
        const unsigned t1 = (unsigned)__float_as_int(tp[g.RC]);
        const float a2 = e < nent ? tp[2 * g.RC] : inf;
        const bool hit = a1 <= sthr;
        // An entry whose SECOND smallest tile minimum is above the threshold has one candidate tile, the recorded one:
        // the usual case (two near-tied tiles in two entries) then costs a single round.
        const unsigned singles = __ballot_sync(full, hit && !(a2 <= sthr));
        unsigned wholes = __ballot_sync(full, hit && a2 <= sthr);
        const int nsingle = __popc(singles);
        for (int r = 0; r < nsingle; r += kItems) {       // kItems sub-tiles per round, one per lane group
          const int idx = r + grp;
          const unsigned from = idx < nsingle ? __fns(singles, 0, idx + 1) : 0u;
          const unsigned tile = __shfl_sync(full, t1, (int)from);
          if (idx < nsingle) eval_tile(tile);
        }
        while (wholes) {                                  // every sub-tile of the entry
          const unsigned he = base + (unsigned)__ffs(wholes) - 1u; wholes &= wholes - 1u;
          const unsigned k = kfirst + he / (unsigned)g.P, pe = he % (unsigned)g.P;
          const unsigned ua = max(k * (unsigned)g.U, u0), ub = min((k + 1u) * (unsigned)g.U, u1);   // the segment's units
          const unsigned ta = ua - u0 + pe * kFSubTiles;              // first 64-column tile of the entry, within the row block
          const unsigned tb = min(ta + kFSubTiles, ub - u0);          // the entry's tiles [ta, tb)
          constexpr unsigned kPer = kFRowTile / kSub;                  // sub-tiles per tile
          for (unsigned tr = ta * kPer; tr < tb * kPer; tr += (unsigned)kItems)
            if (tr + (unsigned)grp < tb * kPer) eval_tile(tr + (unsigned)grp);
        }
      }
      const unsigned mn = __reduce_min_sync(full, bd);
      const unsigned jm = __reduce_min_sync(full, bd == mn ? bj : 0xffffffffu);
      if (lane == src) { best_d = mn; best_j = jm; }
    }

    myd = __uint_as_float(best_d);
    wb = (valid && f.batch_weight) ? f.batch_weight[b] : 1.f;
    // no candidate at all happens only for NaN coordinates (every comparison fails): NaN distance, index 0
    const int j = best_d == 0xffffffffu ? 0 : (int)best_j;
    if (valid) {
      if (f.dist1) f.dist1[t] = myd;
      if (f.idx1) f.idx1[t] = j;
    }
    if (f.own_x || f.scat_y) {
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (valid) {
        const float4 q = w.colpack[(size_t)b * g.Mp + j];
        const float cg = coef_grad * wb;
        vx = cg * __fmaf_rn(0.5f, q.x, rp.x);
        vy = cg * __fmaf_rn(0.5f, q.y, rp.y);
        vz = cg * __fmaf_rn(0.5f, q.z, rp.z);
        if (f.own_x) { float* o = f.own_x + (size_t)t * 3; o[0] = vx; o[1] = vy; o[2] = vz; }
      }
      if (f.scat_y) scatter_add4(f.scat_y + ((long long)b * g.M + j), (long long)b * g.M + j, valid, -vx, -vy, -vz);
    }
  } else {
    // ------------------------------------------------------------- columns: lane <-> column, candidates = 16-row groups
    const unsigned t = (wi - row_warps) * 32u + lane;
    valid = t < ncol;
    int b = 0, j = 0;
    float4 cq = make_float4(0.f, 0.f, 0.f, 0.f);
    float M1 = inf, M2 = inf;
    int R1 = 0;
    if (valid) {
      b = (int)(t / (unsigned)g.M); j = (int)(t - (unsigned)b * (unsigned)g.M);
      cq = w.colpack[(size_t)b * g.Mp + j];
      const float* rec = w.colrec + (size_t)b * g.nRB * kFColTop * g.Mp + j;
      constexpr int kRbBatch = 8;                   // records of 8 row blocks in flight (C2: all of them)
      for (int r0 = 0; r0 < g.nRB; r0 += kRbBatch) {
        float a1[kRbBatch], a2[kRbBatch];
#pragma unroll
        for (int k = 0; k < kRbBatch; ++k) {
          const bool on = r0 + k < g.nRB;
          const float* rp_ = rec + (size_t)(on ? r0 + k : r0) * kFColTop * g.Mp;
          a1[k] = on ? rp_[0] : inf;
          a2[k] = on ? rp_[g.Mp] : inf;
        }
#pragma unroll
        for (int k = 0; k < kRbBatch; ++k) {
          M2 = fminf(fminf(M2, a2[k]), fmaxf(M1, a1[k]));
          if (a1[k] < M1) { M1 = a1[k]; R1 = r0 + k; }
        }
      }
    }
    const float thr = M1 + guard_col(cq.w, M1);
#if defined(RMA_B200_EXPERIMENTS) && (defined(RMA_EXP_NO_SLOW) || defined(RMA_EXP_NO_SLOW_COLS))   // timing experiment only (tools/resolve_variants.py): wrong results for the slow-path items
    const bool fast = valid;
#else
    const bool fast = valid && M2 > thr;
#endif
    const int G1 = (R1 * 32 + (int)(__float_as_uint(M1) & 31u)) * kFRowsPerThread;   // first row of the candidate group
    unsigned best_d = 0xffffffffu, best_i = 0;

    // Fast path (one candidate group of 16 rows): kLanes lanes per column, kGroupRows rows per lane, kItems columns per
    // round.  As for the rows: inside the candidate group the guard applies per ROW (g_i = fl(f + |x_i|^2) <= thr for
    // every exact minimiser; padded rows carry |x|^2 = 1e30), so the 16 rows are re-filtered with the search's chain
    // and only the survivors are evaluated exactly.
    sA[warp][lane] = make_float4(cq.x, cq.y, cq.z, cq.w);
    sB[warp][lane] = make_int4(b, G1, fast ? 1 : 0, __float_as_int(thr));
    __syncwarp();
#pragma unroll 1
    for (int round = 0; round < kRounds; ++round) {
      const int src = round * kItems + grp;
      const float4 q = sA[warp][src];
      const int4 sb4 = sB[warp][src];
      const float sthr = __int_as_float(sb4.w);
      unsigned d = 0xffffffffu, di = 0xffffffffu;
      if (sb4.z) {
        const int r0 = sb4.y + sub;
        const float4* rq = w.rowpack + (size_t)sb4.x * g.Npad + r0;
        float4 pr[kGroupRows];
#pragma unroll
        for (int k = 0; k < kGroupRows; ++k) pr[k] = rq[kLanes * k];
        unsigned pass = 0u;
#pragma unroll
        for (int k = 0; k < kGroupRows; ++k) {
          const float gk = __fadd_rn(__fmaf_rn(pr[k].x, q.x, __fmaf_rn(pr[k].y, q.y, __fmaf_rn(pr[k].z, q.z, q.w))), pr[k].w);
          pass |= gk <= sthr ? 1u << k : 0u;
        }
        while (pass) {                                     // ascending rows: strict '<' keeps the first
          const int k = __ffs(pass) - 1;
          pass &= pass - 1u;
          const float4 pk = rq[kLanes * k];
          const unsigned dk = __float_as_uint(sqdist_q(pk.x, pk.y, pk.z, q));
          if (dk < d) { d = dk; di = (unsigned)(r0 + kLanes * k); }
        }
      }
      unsigned rd, ri;
      round_result<kLanes>(d, di, rd, ri);
      if (lane / kItems == round) { best_d = rd; best_i = ri; }
    }

    // Slow path: every row block's smallest entry within the guard is a candidate group; a second-smallest entry
    // within the guard means the row block may hold more than the recorded two, so all of its 512 rows are re-evaluated.
    unsigned slow = __ballot_sync(full, valid && !fast);
    if (slow && lane == 0) atomicAdd(w.stats + 1, (unsigned)__popc(slow));
    while (slow) {
      const int src = __ffs(slow) - 1; slow &= slow - 1;
      const int sb = __shfl_sync(full, b, src), sj = __shfl_sync(full, j, src);
      const float sthr = __shfl_sync(full, thr, src);
      const float4 q = make_float4(__shfl_sync(full, cq.x, src), __shfl_sync(full, cq.y, src),
                                   __shfl_sync(full, cq.z, src), 0.f);
      const float* rec = w.colrec + (size_t)sb * g.nRB * kFColTop * g.Mp + sj;
      unsigned bd = 0xffffffffu, bi = 0xffffffffu;
      const int nE = g.nRB * kFColTop;
      for (int base = 0; base < nE; base += 32) {
        const int e = base + lane;
        const float p = e < nE ? rec[(size_t)e * g.Mp] : inf;
        unsigned hits = __ballot_sync(full, p <= sthr);
        while (hits) {
          const int hl = __ffs(hits) - 1; hits &= hits - 1;
          const int he = base + hl, hrb = he / kFColTop, hk = he - hrb * kFColTop;
          const unsigned pbits = __float_as_uint(__shfl_sync(full, p, hl));
          int first, n;
          if (hk < kFColTop - 1) { first = (hrb * 32 + (int)(pbits & 31u)) * kFRowsPerThread; n = kFRowsPerThread; }
          else { first = hrb * kFRowsPerWarp; n = kFRowsPerWarp; if (lane == 0) atomicAdd(w.stats + 2, 1u); }
          unsigned dm = 0xffffffffu, im = 0xffffffffu;
          for (int o = lane; o < n; o += 32) {
            const int row = first + o;
            if (row < g.N) {
              const float4 pr = w.rowpack[(size_t)sb * g.Npad + row];
              const unsigned d = __float_as_uint(sqdist_q(pr.x, pr.y, pr.z, q));
              if (d < dm) { dm = d; im = (unsigned)row; }
            }
          }
          const unsigned mn = __reduce_min_sync(full, dm);
          const unsigned ii = __reduce_min_sync(full, dm == mn ? im : 0xffffffffu);
          if (mn < bd || (mn == bd && ii < bi)) { bd = mn; bi = ii; }
        }
      }
      if (lane == src) { best_d = bd; best_i = bi; }
    }

    myd = __uint_as_float(best_d);
    wb = (valid && f.batch_weight) ? f.batch_weight[b] : 1.f;
    const int i = best_d == 0xffffffffu ? 0 : (int)best_i;
    if (valid) {
      if (f.dist2) f.dist2[t] = myd;
      if (f.idx2) f.idx2[t] = i;
    }
    if (f.own_y || f.scat_x) {
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (valid) {
        const float4 p = w.rowpack[(size_t)b * g.Npad + i];
        // d/dy_j |y_j - x_i|^2 = 2 (y_j - x_i) = -2 (x_i - y_j)
        const float cg = coef_grad * wb;
        vx = -cg * __fmaf_rn(0.5f, cq.x, p.x);
        vy = -cg * __fmaf_rn(0.5f, cq.y, p.y);
        vz = -cg * __fmaf_rn(0.5f, cq.z, p.z);
        if (f.own_y) { float* o = f.own_y + (size_t)t * 3; o[0] = vx; o[1] = vy; o[2] = vz; }
      }
      if (f.scat_x) scatter_add4(f.scat_x + ((long long)b * g.N + i), (long long)b * g.N + i, valid, -vx, -vy, -vz);
    }
  }

  if (f.loss) {
    double sd = valid ? (double)myd * (double)wb : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
    if (lane == 0) lsum[warp] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int k = 0; k < kFinWarps; ++k) tot += lsum[k];
      atomicAdd(f.loss, (float)(tot * (double)f.coef_loss));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side (called from the dispatcher in chamfer.cu)
// ---------------------------------------------------------------------------------------------------------------
bool f_sizes_ok(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return false;
  if ((long long)N > (1LL << 30) || (long long)M > (1LL << 30)) return false;
  const FGeometry g = f_make_geometry(B, N, M, 296, 1);
  if ((long long)B * g.Npad >= (1LL << 31) || (long long)B * g.Mp >= (1LL << 31)) return false;
  if ((long long)B * g.nRBC >= (1LL << 24)) return false;
  if ((long long)B * g.nRBC * g.nT >= (1LL << 31)) return false;   // unit indices are 32-bit in the resolve
  return true;
}

// Bytes of the filter's group records alone (the dispatcher's size criterion; device independent): B*N*M/64.
size_t f_record_bytes(int B, int N, int M) {
  const FGeometry g = f_make_geometry(B, N, M, 296, 1);
  return ((size_t)g.B * g.nRB * kFColTop * g.Mp) * sizeof(float);
}

// SM count and resident search CTAs per SM of the current device.  Immutable device properties, memoised per device
// (the occupancy query costs ~10 us of host time, which matters for eager callers at ~90 us per step).  The memo is
// a pure cache of constants: one packed atomic word per device (release / acquire), so concurrent first calls from
// several threads at worst query twice and store the same value.
static int f_caps(int* sms, int* occ) {
  static std::atomic<unsigned> memo[64];   // 0 = empty, else (sms << 8) | occ
  int dev = 0;
  RMA_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64) {
    const unsigned m = memo[dev].load(std::memory_order_acquire);
    if (m) { *sms = (int)(m >> 8); *occ = (int)(m & 0xffu); return 0; }
  }
  RMA_CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_fsearch_kernel<kFSubLarge, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)f_smem(4)));
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_fsearch_kernel<kFSubSmall, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)f_smem(4)));
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_fsearch_kernel<kFSubSmall, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)f_smem(2)));
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_fsearch_kernel<kFSubSmall, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)f_smem(1)));
  // occupancy of the 4-warp CTA; the 2- and 1-warp CTAs use half / a quarter of its registers and shared memory and
  // are launch-bounded to 2x / 4x as many resident CTAs (the same warps per SM)
  RMA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, chamfer_fsearch_kernel<kFSubLarge, 4>, kFThreads,
                                                             f_smem(4)));
  if (*occ < 1) *occ = 1;
  if (*occ > 255) *occ = 255;
  if (dev >= 0 && dev < 64) memo[dev].store(((unsigned)*sms << 8) | (unsigned)*occ, std::memory_order_release);
  return 0;
}

// Geometry for the current device (the unit run depends on SM count x resident CTAs).  Re-derived per call: the
// library keeps no state, and every stage of one forward derives the identical geometry.
static int f_geometry(int B, int N, int M, int unit_run, FGeometry* g, int* occ_out) {
  int sms = 0, occ = 0;
  if (int e = f_caps(&sms, &occ)) return e;
  const int W = f_warps_for((N + kFRowsPerWarp - 1) / kFRowsPerWarp);
  occ *= kFWarps / W;                       // resident CTAs per SM for this CTA shape
  *g = f_make_geometry(B, N, M, sms * occ, unit_run);
  if (occ_out) *occ_out = occ;
  return 0;
}

size_t f_workspace_bytes(int B, int N, int M, int unit_run) {
  FGeometry g;
  if (f_geometry(B, N, M, unit_run, &g, nullptr)) return 0;
  return f_carve(g, nullptr).bytes;
}

static int f_setup(int B, int N, int M, int unit_run, void* workspace, size_t workspace_bytes, FGeometry* g,
                   FWorkspace* w) {
  if (!workspace || !f_sizes_ok(B, N, M)) return RMA_E_BADARG;
  if (int e = f_geometry(B, N, M, unit_run, g, nullptr)) return e;
  if (((uintptr_t)workspace & 255u) != 0) return RMA_E_WORKSPACE;
  *w = f_carve(*g, workspace);
  if (workspace_bytes < w->bytes) return RMA_E_WORKSPACE;
  return RMA_OK;
}

int f_launch_prep(const float* x, int64_t xsb, int64_t xsn, int64_t xsc,
                  const float* y, int64_t ysb, int64_t ysn, int64_t ysc, int B, int N, int M, int unit_run,
                  void* workspace, size_t workspace_bytes, float* zero_loss, float4* zero_gx, float4* zero_gy,
                  void* stream_) {
  if (!x || !y) return RMA_E_BADARG;
  FGeometry g; FWorkspace w;
  if (int e = f_setup(B, N, M, unit_run, workspace, workspace_bytes, &g, &w)) return e;
  const long long slots = (long long)B * (g.Npad + g.Mp);
  RMA_CUDA_TRY(launch_dep(pdl_site(1), chamfer_fprep_kernel, (unsigned)grid_for(slots, 256), 256u, 0, (cudaStream_t)stream_,
                          x, (long long)xsb, (long long)xsn, (long long)xsc, y, (long long)ysb, (long long)ysn,
                          (long long)ysc, g, w, zero_loss, zero_gx, zero_gy));
  return RMA_OK;
}

int f_search_config(int B, int N, int M, int unit_run, int* ctas, int* threads, int* splits, int* ctas_per_sm) {
  if (!f_sizes_ok(B, N, M)) return RMA_E_BADARG;
  FGeometry g;
  int occ = 0;
  if (int e = f_geometry(B, N, M, unit_run, &g, &occ)) return e;
  if (ctas) *ctas = (int)((g.units() + g.U - 1) / g.U);
  if (threads) *threads = 32 * g.W;
  if (splits) *splits = g.U;          // for this path: 64-column tiles per CTA
  if (ctas_per_sm) *ctas_per_sm = occ;
  return RMA_OK;
}

int f_launch_search(int B, int N, int M, int unit_run, void* workspace, size_t workspace_bytes, void* stream_) {
  FGeometry g; FWorkspace w;
  if (int e = f_setup(B, N, M, unit_run, workspace, workspace_bytes, &g, &w)) return e;
  FSearchArgs a;
  a.rowsoa = w.rowsoa; a.colpair = w.colpair; a.rowtop = w.rowtop; a.colrec = w.colrec;
  a.N = N; a.M = M; a.nRB = g.nRB; a.Npad = g.Npad; a.nRBC = g.nRBC; a.nT = g.nT; a.Mp = g.Mp;
  a.U = g.U; a.Smax = g.Smax; a.P = g.P; a.units = g.units();
  const long long ctas = (a.units + g.U - 1) / g.U;
  if (ctas > 0x7fffffffLL) return RMA_E_BADARG;
  const bool small = f_sub_for(N, M) == kFSubSmall;
  cudaStream_t st = (cudaStream_t)stream_;
  // clouds of <= 1024 rows always take 32-column sub-tiles (f_sub_for: min(N, M) <= 4096)
  if (g.W == 4 && !small)
    RMA_CUDA_TRY(launch_dep(pdl_site(2), chamfer_fsearch_kernel<kFSubLarge, 4>, (unsigned)ctas, 128u, f_smem(4), st, a));
  else if (g.W == 4)
    RMA_CUDA_TRY(launch_dep(pdl_site(2), chamfer_fsearch_kernel<kFSubSmall, 4>, (unsigned)ctas, 128u, f_smem(4), st, a));
  else if (g.W == 2)
    RMA_CUDA_TRY(launch_dep(pdl_site(2), chamfer_fsearch_kernel<kFSubSmall, 2>, (unsigned)ctas, 64u, f_smem(2), st, a));
  else
    RMA_CUDA_TRY(launch_dep(pdl_site(2), chamfer_fsearch_kernel<kFSubSmall, 1>, (unsigned)ctas, 32u, f_smem(1), st, a));
  return RMA_OK;
}

int f_launch_resolve(int B, int N, int M, int unit_run, const FinalizeArgs& f, void* workspace,
                     size_t workspace_bytes, void* stream_) {
  FGeometry g; FWorkspace w;
  if (int e = f_setup(B, N, M, unit_run, workspace, workspace_bytes, &g, &w)) return e;
  const long long warps = ((long long)B * N + 31) / 32 + ((long long)B * M + 31) / 32;
  int sms = 0, occ = 0;
  if (int e = f_caps(&sms, &occ)) return e;
  long long grid = grid_for(warps, kFinWarps);            // one CTA per 8 items when that is more than one wave
  const long long wave = (long long)sms * 4;              // resident resolve CTAs (launch bounds: 4 per SM)
  if (grid <= wave) grid = warps < wave ? warps : wave;   // single wave: the same number of CTAs on every SM
  if (f_sub_for(N, M) == kFSubSmall)
    RMA_CUDA_TRY(launch_dep(pdl_site(4), chamfer_fresolve_kernel<kFSubSmall>, (unsigned)grid,
                            (unsigned)(kFinWarps * 32), 0, (cudaStream_t)stream_, g, w, f));
  else
    RMA_CUDA_TRY(launch_dep(pdl_site(4), chamfer_fresolve_kernel<kFSubLarge>, (unsigned)grid,
                            (unsigned)(kFinWarps * 32), 0, (cudaStream_t)stream_, g, w, f));
  return RMA_OK;
}

int f_read_stats(void* workspace, int B, int N, int M, int unit_run, unsigned* out4_host, void* stream_) {
  FGeometry g;
  if (int e = f_geometry(B, N, M, unit_run, &g, nullptr)) return e;
  FWorkspace w = f_carve(g, workspace);
  RMA_CUDA_TRY(cudaMemcpyAsync(out4_host, w.stats, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream_));
  RMA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream_));
  return RMA_OK;
}

}  // namespace rma
