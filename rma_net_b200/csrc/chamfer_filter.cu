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
#include <type_traits>
#include <vector>

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
constexpr int kFThreads = 32 * kFWarps;                 // threads of the largest virtual search CTA (reporting)
constexpr int kFSearchThreads = 256;                    // threads of a physical search CTA: 8 / kW virtual CTAs, one physical CTA per SM
static inline int f_warps_for(int nRB) { return nRB >= 3 ? 4 : nRB; }   // 512-row blocks per batch element -> kW
constexpr int kFColTile = 32;                           // columns per column-direction tile (= lanes)
constexpr int kFColTop = 2;                             // packed group minima recorded per column and row block
constexpr int kFRowTile = 64;                           // columns per row-direction tile (its four 16-column groups carry 2 tag bits)
constexpr int kFUnit = 32;                              // columns per work unit (CTA row block x 32 columns = one column-direction pass):
                                                        // half a tile, so that small problems split evenly over the SMs
// Row direction: per row the minimum of f over each GROUP of 16 columns carries the group's position inside its
// 64-column tile in the two low mantissa bits; per tile the two smallest tagged group minima are kept (registers), per
// search segment the running (smallest, its tile, second smallest) over the tiles (shared memory).  The resolve then
// re-evaluates the 16 columns of ONE group per row (256 bytes) instead of a whole tile.
constexpr int kFGroup = 16;
constexpr int kFGroupsPerTile = kFRowTile / kFGroup;    // 4: two tag bits
constexpr unsigned kFRowTagMask = (unsigned)kFGroupsPerTile - 1u;
constexpr int kFSubTiles = 16;                          // 64-column tiles per row-direction entry (sub-segment): bounds
                                                        // the slow path's re-evaluation to 1024 columns per candidate
// Scheduling fence between the FFMA2 runs and the minima of a loop body: an instruction
// ptxas does not move arithmetic across.  griddepcontrol.launch_dependents (SASS PREEXIT) is one - a repeated call is a
// no-op for the launch logic - and the cheapest of those tried (pmevent also fences but stalls longer; __syncwarp,
// %clock reads, nanosleep and fence.proxy.async do not stop the scheduler).  -DRMA_F_NO_FENCE builds the unfenced loop
// (tools/ A/B).
#ifdef RMA_F_NO_FENCE
#define RMA_F_FENCE() ((void)0)
#else
#define RMA_F_FENCE() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#endif
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

// Timeline instrumentation for tools/ builds (-DRMA_B200_TIMELINE): per kernel phase the earliest and latest
// %globaltimer over all CTAs, kept behind the guard statistics.  Not compiled into the product library.
#ifdef RMA_B200_TIMELINE
__device__ __forceinline__ void tl_mark(unsigned* stats, int slot) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    unsigned long long* p = reinterpret_cast<unsigned long long*>(stats + 8) + 2 * slot;
    atomicMin(p, t);
    atomicMax(p + 1, t);
  }
}
#define RMA_TL(stats, slot) tl_mark(stats, slot)
#else
#define RMA_TL(stats, slot) ((void)0)
#endif

// t / d for t < 2^31 with m = min(ceil(2^32 / d), 2^32 - 1): umulhi(t, m) is off by at most one either way (the
// estimate t m / 2^32 lies in [t/d - t/2^32, t/d + t/2^32)), so one multiply-back and two selects make it exact - six
// instructions against ~20 for the emulated 32-bit division (the resolve had five of them per lane).
__host__ __device__ __forceinline__ unsigned fdiv(unsigned t, unsigned d, unsigned m) {
#ifdef __CUDA_ARCH__
  unsigned q = __umulhi(t, m);
#else
  unsigned q = (unsigned)(((unsigned long long)t * m) >> 32);
#endif
  const int r = (int)(t - q * d);
  q = r < 0 ? q - 1u : q;
  q = r >= (int)d ? q + 1u : q;
  return q;
}

// How the B*nRBC*nT work units of the search are dealt out to its CTAs (one wave that fills every resident slot; a
// CTA's share of one row block is a "segment", for which it loads the rows once and writes one set of row records).
//   mode 0  contiguous runs over the whole unit range: the first nL CTAs take U units, the others Us = U - 1; a run may
//           cross into the next row block / batch element (the CTA then starts a second segment: row reload, record
//           flush, a drained pipeline - about one unit's time).  Used when there are more row blocks than CTAs, where a
//           run spans row blocks anyway, or when forced (rma_chamfer_opts.unit_run, tests).
//   mode 1  row-block aligned: the first Rhi row blocks are shared by clo + 1 CTAs each, the others by clo; inside a row
//           block of c CTAs the first nT % c runs are nT / c + 1 units long, the others nT / c.  No CTA crosses a row
//           block, so every CTA has exactly one segment.  CTAs are dispatched breadth-first in index order (block k and
//           block k + #SMs share an SM: tools/cta_placement_probe.cu), so the search maps its block index through a
//           serpentine over the runs SORTED by length (cls_*: up to four classes of equal length) and every SM gets
//           long runs together with short ones.
// The host picks the mode with the smaller per-SM maximum (f_make_geometry).
struct FDeal {
  int mode;
  // mode 0
  int U, Us, nL;
  unsigned uL;         // nL * U: first unit of the short CTAs
  unsigned magicU, magicUs;
  // mode 1
  int Rhi, clo;
  int ncls;
  int cls_base[5];     // first sorted index of class i; cls_base[ncls] = nC
  int cls_per[4];      // runs of this class per row block
  int cls_rb0[4];      // first row block of the class
  int cls_pos0[4];     // first position (segment index inside the row block) of the class
};

__host__ __device__ __forceinline__ long long f_cta_start(int k, int U, int Us, int nL, unsigned uL) {
  return k <= nL ? (long long)k * U : (long long)uL + (long long)(k - nL) * Us;
}
// mode 1: units [ua, ub) (relative to the row block) of segment `pos` of a row block shared by c CTAs
__host__ __device__ __forceinline__ void f_aligned_run(int nT, int c, int pos, unsigned& ua, unsigned& ub) {
  const int q = nT / c, r = nT - q * c;
  ua = (unsigned)(pos * q + (pos < r ? pos : r));
  ub = ua + (unsigned)(q + (pos < r ? 1 : 0));
}

struct FGeometry {
  int B, N, M;
  int nRB;    // 512-row blocks per batch element
  int Npad;   // nRB * 512
  int W;      // warps per search CTA (4, or 2 / 1 for clouds of <= 1024 / 512 rows)
  int RC;     // rows per CTA row block: 512 * W
  int nRBC;   // CTA row blocks (RC rows) per batch element
  int nT;     // 32-column work units per batch element (even: Mp is a multiple of the 64-column tile)
  int Mp;     // nT * 32
  FDeal d;    // how the units are dealt out to the search CTAs
  int nC;     // virtual search CTAs (runs)
  int G;      // physical search CTAs (grid): virtual CTA v = block + layer * G, layer = warp group of the block
  int maxrun; // units of the longest run
  int yper;   // > 0: batch element b reads y[b % yper] (rma_chamfer_opts.y_batch_period); prep only
  int Smax;   // max segments (CTAs) that share one CTA row block
  int P;      // row-direction entries per segment: a run of n units touches <= n / 2 + 1 tiles, one entry per 16 tiles
  unsigned magicN, magicM;   // ceil(2^32 / d) (clipped to 2^32 - 1): fast exact division in the resolve (fdiv)
  int rc_shift;                      // log2(RC)
  long long units() const { return (long long)B * nRBC * nT; }
};

static FGeometry f_make_geometry(int B, int N, int M, int sms, int occ, int force_run) {
  FGeometry g;
  g.B = B; g.N = N; g.M = M;
  g.nRB = (N + kFRowsPerWarp - 1) / kFRowsPerWarp;
  g.Npad = g.nRB * kFRowsPerWarp;
  g.W = f_warps_for(g.nRB);
  g.RC = g.W * kFRowsPerWarp;
  g.nRBC = (g.nRB + g.W - 1) / g.W;
  g.nT = ((M + kFRowTile - 1) / kFRowTile) * (kFRowTile / kFUnit);
  g.Mp = g.nT * kFUnit;
  const long long units = g.units();
  const long long slots = (long long)sms * occ;
  FDeal& d = g.d;
  d = FDeal{};
  int maxrun;
  if (force_run > 0) {
    d.U = force_run;
    const long long nc = (units + force_run - 1) / force_run;
    g.nC = (int)(nc > 0x7fffffffLL ? 0x7fffffffLL : nc);
    d.nL = g.nC;
#ifdef RMA_F_UNIFORM_RUNS   // tools/ A/B only: the round-1 dealing, ceil(units / slots) units for every CTA
  } else if (true) {
    const long long U = (units + slots - 1) / slots;
    d.U = (int)(U < 1 ? 1 : U);
    g.nC = (int)((units + d.U - 1) / d.U);
    d.nL = g.nC;
#endif
  } else {
    g.nC = (int)(units < slots ? units : slots);
    if (g.nC < 1) g.nC = 1;
    const long long U = (units + g.nC - 1) / g.nC;
    d.U = (int)(U > 0x3fffffff ? 0x3fffffff : (U < 1 ? 1 : U));
    d.nL = (int)(units - (long long)g.nC * (d.U - 1));       // in [1, nC]
  }
  {
    const long long G = (long long)g.nC <= slots ? ((long long)g.nC < sms ? g.nC : sms) : ((long long)g.nC + occ - 1) / occ;
    g.G = (int)(G > 0x7fffffffLL ? 0x7fffffffLL : G);
  }
  d.Us = d.U > 1 ? d.U - 1 : 1;
  d.uL = (unsigned)((long long)d.nL * d.U > 0x7fffffffLL ? 0x7fffffffLL : (long long)d.nL * d.U);
  const int shortest = d.nL < g.nC ? d.Us : d.U;
  g.Smax = (g.nT + shortest - 1) / shortest + 1;
  maxrun = d.U;
  auto magic = [](unsigned dv) { const unsigned long long m = ((1ull << 32) + dv - 1) / dv; return (unsigned)(m > 0xffffffffull ? 0xffffffffull : m); };
  d.magicU = magic((unsigned)d.U);
  d.magicUs = magic((unsigned)d.Us);

#ifndef RMA_F_UNIFORM_RUNS
  // Row-block aligned dealing (mode 1) when every row block can have its own CTAs and the sorted serpentine loads the
  // SMs no worse than the contiguous runs do (whose row-block crossings cost about one unit's time on the SM that has one).
  const long long R = (long long)B * g.nRBC;
  if (force_run <= 0 && (long long)g.nC >= R && g.nC > 1) {
    const int clo = (int)(g.nC / R), Rhi = (int)(g.nC - (long long)clo * R);
    struct Cls { int len, per, rb0, pos0; long long count; } c[4];
    int n = 0;
    auto add = [&](int rb0, long long nrb, int cc) {
      if (nrb <= 0 || cc <= 0) return;
      const int q = g.nT / cc, r = g.nT - q * cc;
      if (r > 0) c[n++] = {q + 1, r, rb0, 0, nrb * r};
      if (cc - r > 0 && q > 0) c[n++] = {q, cc - r, rb0, r, nrb * (cc - r)};
    };
    add(0, Rhi, clo + 1);
    add(Rhi, R - Rhi, clo);
    for (int i = 1; i < n; ++i)                     // insertion sort, longest runs first
      for (int j = i; j > 0 && c[j].len > c[j - 1].len; --j) { const Cls t = c[j]; c[j] = c[j - 1]; c[j - 1] = t; }
    long long base[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) base[i + 1] = base[i] + c[i].count;
    if (base[n] == (long long)g.nC) {               // every CTA has a run (c <= nT in both row-block classes)
      auto len_sorted = [&](long long sidx) { int i = 0; while (i + 1 < n && sidx >= base[i + 1]) ++i; return c[i].len; };
      auto serp = [&](int j, int t) -> long long {   // sorted index of the CTA that SM j runs in layer t, or -1
        const long long lo = (long long)t * g.G, sz = (g.nC - lo) < g.G ? (g.nC - lo) : g.G;
        if (j >= sz) return -1;
        return lo + ((t & 1) ? sz - 1 - j : j);
      };
      auto sm_load = [&](int j) { long long tot = 0; for (int t = 0; (long long)t * g.G < g.nC; ++t) { const long long x = serp(j, t); if (x >= 0) tot += len_sorted(x); } return tot; };
      // the load is piecewise constant in j: evaluate next to every class boundary of every layer
      long long cost1 = sm_load(0);
      for (int t = 0; (long long)t * g.G < g.nC; ++t)
        for (int i = 1; i < n; ++i) {
          const long long bidx = base[i] - (long long)t * g.G;
          const long long sz = (g.nC - (long long)t * g.G) < g.G ? (g.nC - (long long)t * g.G) : g.G;
          if (bidx <= 0 || bidx >= sz) continue;
          const long long jb = (t & 1) ? sz - 1 - bidx : bidx;
          for (long long jj = jb - 1; jj <= jb + 1; ++jj)
            if (jj >= 0 && jj < g.G) { const long long v = sm_load((int)jj); if (v > cost1) cost1 = v; }
        }
      long long cost0 = 0;
      for (int t = 0; (long long)t * g.G < g.nC; ++t) cost0 += ((long long)t * g.G < d.nL) ? d.U : d.Us;
      cost0 += (R > 1) ? 1 : 0;                     // a crossing
      if (cost1 <= cost0) {
        d.mode = 1;
        d.Rhi = Rhi; d.clo = clo; d.ncls = n;
        for (int i = 0; i < n; ++i) { d.cls_base[i] = (int)base[i]; d.cls_per[i] = c[i].per; d.cls_rb0[i] = c[i].rb0; d.cls_pos0[i] = c[i].pos0; }
        d.cls_base[n] = g.nC;
        g.Smax = Rhi > 0 ? clo + 1 : clo;
        maxrun = c[0].len;
      }
    }
  }
#endif
  g.maxrun = maxrun;
  g.yper = 0;
  g.P = ((maxrun < g.nT ? maxrun : g.nT) / 2 + 1 + kFSubTiles - 1) / kFSubTiles;
  g.magicN = magic((unsigned)N); g.magicM = magic((unsigned)M);
  g.rc_shift = g.W == 4 ? 11 : g.W == 2 ? 10 : 9;
  return g;
}

// Segments of row block brb: how many CTAs share it (nseg) and, for mode 0, the first of them (kfirst).
__host__ __device__ __forceinline__ void f_rowblock_segments(const FDeal& d, unsigned brb, int nT, int& kfirst, int& nseg) {
  if (d.mode == 1) {
    kfirst = 0;
    nseg = (int)brb < d.Rhi ? d.clo + 1 : d.clo;
    return;
  }
  const unsigned ufirst = brb * (unsigned)nT, ulast = ufirst + (unsigned)nT - 1u;
  auto cta_of = [&](unsigned u) {
    return u < d.uL ? fdiv(u, (unsigned)d.U, d.magicU) : (unsigned)d.nL + fdiv(u - d.uL, (unsigned)d.Us, d.magicUs);
  };
  kfirst = (int)cta_of(ufirst);
  nseg = (int)cta_of(ulast) - kfirst + 1;
}
// Units [ua, ub), relative to the row block, of its segment number srel.
__host__ __device__ __forceinline__ void f_segment_units(const FDeal& d, unsigned brb, int nT, int kfirst, int nseg,
                                                         unsigned srel, unsigned& ua, unsigned& ub) {
  if (d.mode == 1) { f_aligned_run(nT, nseg, (int)srel, ua, ub); return; }
  const unsigned u0 = brb * (unsigned)nT, u1 = u0 + (unsigned)nT;
  const unsigned ks = (unsigned)f_cta_start(kfirst + (int)srel, d.U, d.Us, d.nL, d.uL);
  const unsigned ke = (unsigned)f_cta_start(kfirst + (int)srel + 1, d.U, d.Us, d.nL, d.uL);
  ua = (ks > u0 ? ks : u0) - u0; ub = (ke < u1 ? ke : u1) - u0;
}
// The run of virtual CTA (block, layer) of a grid of G physical CTAs: units [unit, unit_end) and, in aligned mode, its
// segment index inside its row block.  Shared by the search kernel and the host-side self-check.
__host__ __device__ __forceinline__ void f_virtual_run(const FDeal& d, int nT, int nC, long long units, int G, int block,
                                                       int layer, long long& unit, long long& unit_end, int& seg_aligned) {
  seg_aligned = 0;
  if (d.mode == 1) {
    // serpentine over the runs sorted by length: layer `layer` of the SM that runs physical CTA `block`
    const int lo = layer * G, rest = nC - lo, sz = G < rest ? G : rest;
    const int sidx = lo + ((layer & 1) ? sz - 1 - block : block);
    int ci = 0;
    while (ci + 1 < d.ncls && sidx >= d.cls_base[ci + 1]) ++ci;
    const int o = sidx - d.cls_base[ci];
    const int rbk = d.cls_rb0[ci] + o / d.cls_per[ci];
    seg_aligned = d.cls_pos0[ci] + o % d.cls_per[ci];
    unsigned ua, ub;
    f_aligned_run(nT, rbk < d.Rhi ? d.clo + 1 : d.clo, seg_aligned, ua, ub);
    unit = (long long)rbk * nT + ua;
    unit_end = (long long)rbk * nT + ub;
  } else {
    const int v = block + layer * G;
    unit = f_cta_start(v, d.U, d.Us, d.nL, d.uL);
    const long long e = f_cta_start(v + 1, d.U, d.Us, d.nL, d.uL);
    unit_end = e < units ? e : units;
  }
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
  float* colrec;     // [B][nRB][Mp][2]         two smallest packed group minima of g over the row block, interleaved
                     //                         per column: one 8-byte store in the search, one 8-byte load in the resolve
  unsigned* stats;   // [64]: [0..3] rows on the slow path, columns on the slow path, row-block rescans, unused;
                     //       [8..39] timeline marks (tools/ builds only); [48..49] the loss accumulator (double; loss_commit)
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
  RMA_TL(w.stats, 0);
  if (blockIdx.x == 0 && threadIdx.x < 4) w.stats[threadIdx.x] = 0u;
  if (blockIdx.x == 0 && threadIdx.x >= 48 && threadIdx.x < 50) w.stats[threadIdx.x] = 0u;   // loss accumulator (double)
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
        const float* s = y + (g.yper > 0 ? b % g.yper : b) * ysb + j * ysn;
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
  RMA_TL(w.stats, 1);
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
  unsigned* stats;
  int N, M, nRB, Npad, nRBC, nT, Mp, nC, Smax, P;
  long long units;
  FDeal d;
};

// One physical CTA of 256 threads per SM holds 8 / kW "virtual CTAs" (groups of kW warps: what used to be a CTA of its
// own, with its own shared memory, mbarriers and named barrier).  Which runs share an SM is then decided by the dealing
// (FDeal) and not by where the block scheduler happens to put the second CTA of an SM: 296 separate CTAs land
// breadth-first only approximately (the first 148 blocks cover 142 SMs, tools/cta_placement_probe.cu), and six SMs with
// two long runs are enough to lose what the balanced dealing gains.
template <int kW>
__global__ void __launch_bounds__(kFSearchThreads, 1) chamfer_fsearch_kernel(FSearchArgs a) {
  static_assert(kW == 1 || kW == 2 || kW == 4, "1, 2 or 4 warps per virtual search CTA");
  constexpr int kThreads = 32 * kW, kRowsPerCta = kFRowsPerWarp * kW, kChunk = kFChunk * kW / 4;
  extern __shared__ __align__(16) unsigned char smem_all[];
  const int group = (int)threadIdx.x / kThreads;                       // virtual CTA slot on this SM (layer of the dealing)
  const int vblock = (int)blockIdx.x + group * (int)gridDim.x;         // virtual CTA index
  if (vblock >= a.nC) return;                                           // the whole group leaves: its barrier is its own
  unsigned char* smem_raw = smem_all + (size_t)group * f_smem(kW);
  auto group_sync = [&]() {
    if (kW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kThreads) : "memory");
  };
  const int tid = (int)threadIdx.x - group * kThreads, lane = tid & 31, warp = tid >> 5;
  float4* scol = reinterpret_cast<float4*>(smem_raw);                                              // [2][kChunk + 4]
  float* cball = reinterpret_cast<float*>(smem_raw + (size_t)2 * (kChunk + 4) * sizeof(float4));    // [warps][2][32][36]
  u64* mbar = reinterpret_cast<u64*>(smem_raw + f_smem_top(kW));                                   // [2]
  // per-row running (smallest tile minimum, its tile, second smallest): touched once per 64-column tile, so it lives
  // in shared memory ([plane][q][thread] float4, conflict-free) instead of 48 registers
  float4* stop = reinterpret_cast<float4*>(smem_raw + f_smem_base(kW)) + tid;                     // [3][4][kThreads]

  float* cb = cball + warp * (2 * kFColTile * kFCbStride);
  const float inf = __int_as_float(0x7f800000);
  const unsigned lane_tag = (unsigned)lane;
  unsigned tag_mask = ~31u;                                    // kept in a register: (bits & mask) | tag is then ONE LOP3
  asm volatile("" : "+r"(tag_mask));
  unsigned rtag_mask = ~kFRowTagMask;
  asm volatile("" : "+r"(rtag_mask));

  long long unit, unit_end;
  int seg_aligned;
  f_virtual_run(a.d, a.nT, a.nC, a.units, (int)gridDim.x, (int)blockIdx.x, group, unit, unit_end, seg_aligned);
  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
  group_sync();
  int cur = 0;           // column buffer the next chunk lands in
  unsigned phases = 0;   // bit s: parity the next wait on buffer s expects
  pdl_wait();
  pdl_launch_dependents();   // the resolve grid may be launched (it blocks in its own pdl_wait until we are done)
  RMA_TL(a.stats, 2);

  while (unit < unit_end) {
    // ---- one segment: a run of `nunit` 32-column units of one CTA row block ----
    const int brb = (int)(unit / a.nT);
    const int unit0 = (int)(unit - (long long)brb * a.nT);
    const int nunit = (int)min((long long)(a.nT - unit0), unit_end - unit);
    unit += nunit;
    const int b = brb / a.nRBC, rbc = brb - b * a.nRBC;
    const int rb = rbc * kW + warp;
    const bool has_rows = rb < a.nRB;   // warp-uniform
    const long long ub0 = (long long)brb * a.nT;                         // the row block's first unit: which CTA holds it?
    const int seg = a.d.mode == 1 ? seg_aligned                          // which of the row block's segments
                                  : vblock - (int)(ub0 < (long long)a.d.uL ? ub0 / a.d.U : a.d.nL + (ub0 - a.d.uL) / a.d.Us);
    const int c0 = unit0 * kFUnit;
    const int ncols = nunit * kFUnit;

    // Column chunks arrive by TMA bulk copy (one contiguous run of the interleaved pair pack) into a double buffer:
    // the segment's first chunk is in flight while the rows are loaded, chunk k+1 while chunk k is computed.  Both
    // buffers are free here: every chunk iteration ends with a group barrier.
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

    float2* crec = reinterpret_cast<float2*>(a.colrec) + (size_t)(b * a.nRB + rb) * a.Mp + c0 + lane;

    // Column direction, software-pipelined: while tile t is computed, lane l folds column l of tile t-1 (32 per-lane
    // packed partial minima in the other half of the double buffer) into its running top-2, four values per body.
    int pend_col = -1;   // column offset (relative to c0) of the tile whose reduce is pending; warp-uniform
    int parity = 0;
    // tile-level (smallest, second smallest) tagged group minimum per row; a tile can straddle two chunks
    float tm[kFRowsPerThread], ts[kFRowsPerThread];
#pragma unroll
    for (int r = 0; r < kFRowsPerThread; ++r) tm[r] = ts[r] = inf;

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
      for (int t0 = 0; has_rows && t0 < nc; t0 += kFUnit) {
        const int hh = unit0 + (cs + t0) / kFUnit;     // this pass: half hh & 1 of the 64-column tile hh >> 1
        const float4* sp = schunk + t0;
        float* cbw = cb + parity * (kFColTile * kFCbStride) + lane;
        const float4* cbr = reinterpret_cast<const float4*>(cb + (parity ^ 1) * (kFColTile * kFCbStride) +
                                                            lane * kFCbStride);
        float p1 = inf, p2 = inf;
        float gm[kFRowsPerThread];       // running minimum of f over the current 16-column group

        // Four columns against the 16 rows, plus four pending values of the previous tile's column reduce.
        // kFirst: the group's first body starts the group minima (no reset instructions, a 2-input minimum).
        auto body = [&](const float4& q0, const float4& q1, const float4& q2, const float4& q3, int jj, auto first_tag) {
          constexpr bool kFirst = decltype(first_tag)::value;
          const float4 pv = cbr[jj >> 2];          // 4 packed partial minima of the pending tile
          // fp32x2 lanes = two COLUMNS, the row value is the broadcast scalar.  The operand that stays fixed over
          // the 16-row inner loop is then the 64-bit column pair, which the operand-reuse cache keeps (2 register
          // words saved per FFMA2 instead of 1): FFMA2 acc, Qpair.reuse, x_row.F32, acc.
          float cm[4];
          // The 96 FFMA2 of the body (two column pairs x three chain steps x 16 rows) are issued as six runs of 16
          // that share their 64-bit column operand (operand-reuse latch: an FFMA2 whose column pair comes from the
          // latch reads 3 register words = 2 issue cycles, otherwise 5 words = 3 cycles), then the minima.  Left to
          // itself ptxas alternates FFMA2 and FMNMX3 one to one, which loses the latch on every FFMA2; RMA_F_FENCE()
          // keeps the two groups apart.
          u64 acc[2][16];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 A = h2 ? q2 : q0, Bq = h2 ? q3 : q1;
            const u64 Q0 = pack2(A.x, A.y), Q1 = pack2(A.z, A.w), Q2 = pack2(Bq.x, Bq.y), WW = pack2(Bq.z, Bq.w);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float za, zb;
              unpack2(pz[k], za, zb);
              acc[h2][2 * k] = fma2(Q2, pack2(za, za), WW); acc[h2][2 * k + 1] = fma2(Q2, pack2(zb, zb), WW);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float ya, yb;
              unpack2(py[k], ya, yb);
              acc[h2][2 * k] = fma2(Q1, pack2(ya, ya), acc[h2][2 * k]);
              acc[h2][2 * k + 1] = fma2(Q1, pack2(yb, yb), acc[h2][2 * k + 1]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float xa, xb;
              unpack2(px[k], xa, xb);
              acc[h2][2 * k] = fma2(Q0, pack2(xa, xa), acc[h2][2 * k]);
              acc[h2][2 * k + 1] = fma2(Q0, pack2(xb, xb), acc[h2][2 * k + 1]);
            }
          }
          RMA_F_FENCE();
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float cl = inf, ch = inf;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float ua, ub, al, ah, bl, bh;
              unpack2(pu[k], ua, ub);
              unpack2(acc[h2][2 * k], al, ah); unpack2(acc[h2][2 * k + 1], bl, bh);
              if (kFirst && h2 == 0) {
                gm[2 * k] = fminf(al, ah);
                gm[2 * k + 1] = fminf(bl, bh);
              } else {
                gm[2 * k] = min3(gm[2 * k], al, ah);
                gm[2 * k + 1] = min3(gm[2 * k + 1], bl, bh);
              }
              unpack2(add2(acc[h2][2 * k], pack2(ua, ua)), al, ah);
              unpack2(add2(acc[h2][2 * k + 1], pack2(ub, ub)), bl, bh);
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
          RMA_F_FENCE();
        };

        // One 16-column group per trip (four bodies).  Operands are prefetched one body ahead into the OTHER register
        // set (ping-pong) so that no register copies are needed: a MOV/IMAD.MOV here queues behind the saturated FMA
        // pipe and stalls the warp.  The last prefetch of a chunk reads the 4-column pad behind the staged data and is
        // never used.
        float4 a0 = sp[0], a1 = sp[1], a2 = sp[2], a3 = sp[3];
        unsigned gid = (unsigned)((hh & 1) * (kFUnit / kFGroup));                   // group position inside the 64-column tile
#pragma unroll 1
        for (int jj = 0; jj < kFColTile; jj += kFGroup, ++gid) {
          float4 b0 = sp[jj + 4], b1 = sp[jj + 5], b2 = sp[jj + 6], b3 = sp[jj + 7];
          body(a0, a1, a2, a3, jj, std::true_type{});
          a0 = sp[jj + 8]; a1 = sp[jj + 9]; a2 = sp[jj + 10]; a3 = sp[jj + 11];
          body(b0, b1, b2, b3, jj + 4, std::false_type{});
          b0 = sp[jj + 12]; b1 = sp[jj + 13]; b2 = sp[jj + 14]; b3 = sp[jj + 15];
          body(a0, a1, a2, a3, jj + 8, std::false_type{});
          a0 = sp[jj + 16]; a1 = sp[jj + 17]; a2 = sp[jj + 18]; a3 = sp[jj + 19];
          body(b0, b1, b2, b3, jj + 12, std::false_type{});
          // fold the group's tagged minima into the tile's two smallest (3 FMNMX + 1 LOP3 per row and group)
#pragma unroll
          for (int r = 0; r < kFRowsPerThread; ++r) {
            const float v = __uint_as_float((__float_as_uint(gm[r]) & rtag_mask) | gid);
            ts[r] = fminf(ts[r], fmaxf(tm[r], v));
            tm[r] = fminf(tm[r], v);
          }
        }

        if (pend_col >= 0) {
          crec[pend_col] = make_float2(p1, p2);
        }
        pend_col = cs + t0;
        parity ^= 1;
        __syncwarp();   // this tile's partial minima visible to all lanes; previous buffer free for reuse

        if ((hh & 1) || cs + t0 + kFUnit == ncols) {     // the tile (or the segment's share of it) is complete
          // Row direction: the tile's two smallest tagged group minima are folded into the running (smallest, its tile,
          // second smallest) of the segment.  (A per-tile record in global memory - B*N*M/16 bytes, read only when a
          // guard fails - cost 2.8 us of the 83 us C2 step in stores that evict dirty L2 lines; the slow path
          // re-evaluates the candidate SEGMENTS instead.)
          {
            const int tile = hh >> 1;
            const float tilef = __int_as_float(tile);
  #pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 m1 = stop[q * kThreads], t1 = stop[(4 + q) * kThreads], m2 = stop[(8 + q) * kThreads];
              m2.x = min3(m2.x, ts[4 * q], fmaxf(m1.x, tm[4 * q]));         t1.x = tm[4 * q] < m1.x ? tilef : t1.x;
              m2.y = min3(m2.y, ts[4 * q + 1], fmaxf(m1.y, tm[4 * q + 1])); t1.y = tm[4 * q + 1] < m1.y ? tilef : t1.y;
              m2.z = min3(m2.z, ts[4 * q + 2], fmaxf(m1.z, tm[4 * q + 2])); t1.z = tm[4 * q + 2] < m1.z ? tilef : t1.z;
              m2.w = min3(m2.w, ts[4 * q + 3], fmaxf(m1.w, tm[4 * q + 3])); t1.w = tm[4 * q + 3] < m1.w ? tilef : t1.w;
              m1.x = fminf(m1.x, tm[4 * q]); m1.y = fminf(m1.y, tm[4 * q + 1]);
              m1.z = fminf(m1.z, tm[4 * q + 2]); m1.w = fminf(m1.w, tm[4 * q + 3]);
              stop[q * kThreads] = m1; stop[(4 + q) * kThreads] = t1; stop[(8 + q) * kThreads] = m2;
            }
          }
#pragma unroll
          for (int r = 0; r < kFRowsPerThread; ++r) tm[r] = ts[r] = inf;
          if (++sub_tiles == kFSubTiles) flush_top();
        }
      }
      group_sync();   // chunk fully consumed: a later bulk copy may refill this buffer
    }

    if (has_rows) {
      if (pend_col >= 0) {   // drain: column reduce of the segment's last tile
        const float* col = cb + (parity ^ 1) * (kFColTile * kFCbStride) + lane * kFCbStride;
        float p1 = inf, p2 = inf;
#pragma unroll
        for (int k = 0; k < 32; ++k) top2_insert(p1, p2, col[k]);
        crec[pend_col] = make_float2(p1, p2);
      }
      if (sub_tiles > 0) flush_top();
      while (sub < a.P) flush_top();   // empty entries
      __syncwarp();   // the next segment reuses the transpose buffers
    }
  }
  RMA_TL(a.stats, 3);
}

// ---------------------------------------------------------------------------------------------------------------
// resolve
// ---------------------------------------------------------------------------------------------------------------
// sqrt.approx (one MUFU instruction, <= 2 ulp; 0 -> 0, inf -> inf) instead of the IEEE sqrtf (~12 instructions and a
// divergent slow path for denormals): the guards multiply every square root by 1.0001, which covers 2 ulp = 2.4e-7.
__device__ __forceinline__ float sqrt_fast(float v) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// Rigorous guards (derivation in the file header).  u = fl(|x_i|^2), mt = smallest filter value f of the row.
__device__ __forceinline__ float guard_row(float u, float mt) {
  const float D0 = fmaxf(mt + u, 0.f);
  const float Dub = 1.001f * D0 + 1e-5f * u;
  const float a = sqrt_fast(u) * 1.0001f;
  const float beta = a + sqrt_fast(Dub) * 1.0001f;
  return (12.3f * beta * (beta + a) + 10.3f * Dub) * (kEps * 1.01f) + 1e-30f;
}
// wj = fl(|y_j|^2), mt = smallest (packed, truncated) value of g for the column, i.e. an estimate of the distance.
__device__ __forceinline__ float guard_col(float wj, float mt) {
  const float D0 = fmaxf(mt, 0.f);
  const float Dub = 1.001f * D0 + 1e-5f * wj;
  const float bn = sqrt_fast(wj) * 1.0001f;
  const float alpha = bn + sqrt_fast(Dub) * 1.0001f;
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

// Tagged row-direction values: a group minimum v is stored as t(v) = (bits(v) & ~3) | group, |t(v) - v| < 4 ulp(v)
// <= 2^-21 |v|.  With M1 the smallest TAGGED group minimum of a row, every group that holds an exact minimiser has
// t(gm) <= M1 + G (1 + 2^-21) + 2.02 * 2^-21 |M1| = M1 + ... + 0.97e-6 |M1| (the 1.2e-6 below also covers the rounding of the final sum; G = guard_row at the true filter minimum; guard_row's own slack in
// its distance estimate, 1e-5 u + 1e-3 D, covers the 2^-21 (u + D) by which M1 may differ from it).
__device__ __forceinline__ float row_threshold(float u, float M1) {
  return M1 + (guard_row(u, M1) * 1.0001f + 1.2e-6f * fabsf(M1));
}

// Fast-path evaluation of the 32 items of a warp, 16 candidates each (a row's candidate group of 16 columns / a
// column's candidate group of 16 rows: 256 contiguous bytes of colpack / rowpack).
//   phase 1  two items per round, one candidate per lane (a half-warp reads its item's 256 bytes as one coalesced
//            LDG.128): the candidate is re-filtered with the search's own chain (bit-identical f / g; padding carries
//            1e30 and never passes) against the item's threshold; a ballot hands every item's 16 pass bits to its owner
//            lane.  All loads of eight rounds are in flight together.
//   phase 2  lane <-> item: the owner evaluates its survivors - usually one - in the exact difference form, ascending,
//            strict '<' (torch.min's first index).
// sA[item] = the item's broadcast data, sB[item] = (first candidate's index in `cand`, fast flag, threshold bits, -).
template <bool kColumns>
__device__ __forceinline__ void fast_items(const float4* __restrict__ cand, const float4* sA, const int2* sB, float4* sQ,
                                           int lane, unsigned& best_d, unsigned& best_k, float4& best_c) {
  const int half = lane >> 4;
  const unsigned c = (unsigned)lane & 15u, sh = ((unsigned)lane & 1u) * 16u;
  unsigned mymask = 0u;
  // Rounds of two items.  Loads run two batches of four rounds ahead of the arithmetic (32 registers in flight), so the
  // sixteen loads cost little more than one L2 round trip.  A candidate that passes is also dropped into its item's
  // shared-memory slot: with exactly one survivor - the usual case - the owner lane evaluates it from there without a
  // second dependent trip to L2.
  auto load4 = [&](int r0, float4 (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k)          // unconditional: an item off the fast path has base 0 and threshold -inf
      v[k] = cand[(unsigned)sB[2 * (r0 + k) + half].x + c];
  };
  auto eval4 = [&](int r0, const float4 (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int item = 2 * (r0 + k) + half;
      const float4 it = sA[item];
      const float thr = __int_as_float(sB[item].y);
      float fv;
      if (kColumns)   // it = the column (q0, q1, q2, w), v = a row (x, y, z, u):  g = fl(f + u)
        fv = __fadd_rn(__fmaf_rn(v[k].x, it.x, __fmaf_rn(v[k].y, it.y, __fmaf_rn(v[k].z, it.z, it.w))), v[k].w);
      else            // it = the row (x, y, z, .), v = a column (q0, q1, q2, w)
        fv = __fmaf_rn(it.x, v[k].x, __fmaf_rn(it.y, v[k].y, __fmaf_rn(it.z, v[k].z, v[k].w)));
      const bool pass = fv <= thr;
      if (pass) sQ[item] = v[k];
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      if ((lane >> 1) == r0 + k) mymask = (bal >> sh) & 0xffffu;
    }
  };
  float4 va[4], vb[4];
  load4(0, va);
  load4(4, vb);
  eval4(0, va);
  load4(8, va);
  eval4(4, vb);
  load4(12, vb);
  eval4(8, va);
  eval4(12, vb);
  __syncwarp();
  const float4 me = sA[lane];
  const unsigned base = (unsigned)sB[lane].x;
  if (mymask != 0u && (mymask & (mymask - 1u)) == 0u) {     // one survivor: its data is in the item's slot
    const float4 cv = sQ[lane];
    best_d = __float_as_uint(kColumns ? sqdist_q(cv.x, cv.y, cv.z, me) : sqdist_q(me.x, me.y, me.z, cv));
    best_k = base + (unsigned)(__ffs(mymask) - 1);
    best_c = cv;
    mymask = 0u;
  }
  while (mymask) {                                          // several survivors: ascending, strict '<' keeps the first
    const int k = __ffs(mymask) - 1;
    mymask &= mymask - 1u;
    const float4 cv = cand[base + (unsigned)k];
    const unsigned d = __float_as_uint(kColumns ? sqdist_q(cv.x, cv.y, cv.z, me) : sqdist_q(me.x, me.y, me.z, cv));
    if (d < best_d) { best_d = d; best_k = base + (unsigned)k; best_c = cv; }
  }
}

// Row-direction records of one row: the smallest tagged group minimum over the row block's entries, its tile, and the
// second smallest.  kRC = rows per CTA row block (compile-time: the three planes of an entry then sit at immediate
// offsets from one running pointer).
template <int kRC>
__device__ __forceinline__ void scan_row_entries(const float* top, int nseg, float& M1, int& T1, float& M2) {
  const float inf = __int_as_float(0x7f800000);
  constexpr int kSegBatch = 12;                 // up to 12 entries in flight: one latency for C2 (10)
  for (int s0 = 0; s0 < nseg; s0 += kSegBatch) {
    float a1[kSegBatch], a2[kSegBatch];
    int at[kSegBatch];
    const float* tp = top + (size_t)s0 * (3 * kRC);
#pragma unroll
    for (int k = 0; k < kSegBatch; ++k) {
      const bool on = s0 + k < nseg;
      const float* q = on ? tp + k * (3 * kRC) : tp;
      a1[k] = q[0]; at[k] = __float_as_int(q[kRC]); a2[k] = q[2 * kRC];
      a1[k] = on ? a1[k] : inf;
      a2[k] = on ? a2[k] : inf;
    }
#pragma unroll
    for (int k = 0; k < kSegBatch; ++k) {
      M2 = fminf(fminf(M2, a2[k]), fmaxf(M1, a1[k]));
      if (a1[k] < M1) { M1 = a1[k]; T1 = at[k]; }
    }
  }
}

// Slow path of ONE row (out of line: its registers must not weigh on the fast path, which runs at the kernel's 64-register
// cap).  Every candidate group lies in an entry (a run of up to 16 tiles of one search segment) whose smallest group
// minimum is within the guard: an entry whose second smallest is not contributes its one recorded group, otherwise all
// of its columns.  The lowest column among the lanes that hold the minimum = torch.min's first index.
__device__ __noinline__ void row_slow_item(const float* __restrict__ rowtop, const float4* __restrict__ colpack, int RC,
                                           int nRBC, int nT, const FDeal* __restrict__ dp, int Smax, int P, int Mp,
                                           int sb, int si, float sthr,
                                           float x, float y, float z, unsigned& out_d, unsigned& out_j, float4& out_q) {
  const unsigned full = 0xffffffffu;
  const float inf = __int_as_float(0x7f800000);
  const int lane = lane_id(), half = lane >> 4, hl = lane & 15;
  const int srbc = si / RC;
  const unsigned sbrb = (unsigned)(sb * nRBC + srbc);
  const FDeal& d = *dp;
  int kfirst, nsegs;
  f_rowblock_segments(d, sbrb, nT, kfirst, nsegs);
  const float* top = rowtop + ((size_t)sbrb * Smax * P * 3) * RC + (si - srbc * RC);
  const float4* scq = colpack + (size_t)sb * Mp;
  unsigned bd = 0xffffffffu, bj = 0xffffffffu;
  float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
  // Latency matters more than work here: a warp with a slow item is the last to finish, and the kernel is as long
  // as its slowest warp.  So the candidates of a slow item are loaded in batches (every load of a batch in flight
  // together) and only then evaluated: one L2 round trip per batch of 8 groups / 128 columns.
  auto eval_col = [&](const float4 q, unsigned jc) {
    const float fj = __fmaf_rn(x, q.x, __fmaf_rn(y, q.y, __fmaf_rn(z, q.z, q.w)));
    if (fj <= sthr) {
      const unsigned d = __float_as_uint(sqdist_q(x, y, z, q));
      if (d < bd || (d == bd && jc < bj)) { bd = d; bj = jc; bq = q; }
    }
  };
  const unsigned nent = (unsigned)nsegs * (unsigned)P;
  for (unsigned base = 0; base < nent; base += 32u) {
    // one entry per lane, all in flight; empty entries hold inf
    const unsigned e = base + (unsigned)lane;
    const float* tp = top + (size_t)(e < nent ? e : base) * (3 * RC);
    const float a1 = e < nent ? tp[0] : inf;
    const unsigned g1 = (unsigned)__float_as_int(tp[RC]) * kFGroupsPerTile + (__float_as_uint(a1) & kFRowTagMask);
    const float a2 = e < nent ? tp[2 * RC] : inf;
    const bool hit = a1 <= sthr;
    const unsigned singles = __ballot_sync(full, hit && !(a2 <= sthr));
    unsigned wholes = __ballot_sync(full, hit && a2 <= sthr);
    const int nsingle = __popc(singles);
    for (int r = 0; r < nsingle; r += 8) {            // eight groups per batch: half-warp <-> group, lane <-> column
      float4 q[4];
      unsigned jc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = min(r + 2 * k + half, nsingle - 1);       // clamped: re-evaluating a candidate is harmless
        const unsigned from = __fns(singles, 0, idx + 1);
        jc[k] = __shfl_sync(full, g1, (int)from) * kFGroup + (unsigned)hl;
        q[k] = scq[jc[k]];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) eval_col(q[k], jc[k]);
    }
    while (wholes) {                                  // every column of the entry: 128 per batch, lane <-> 4 columns
      const unsigned he = base + (unsigned)__ffs(wholes) - 1u; wholes &= wholes - 1u;
      const unsigned srel = he / (unsigned)P, pe = he % (unsigned)P;
      unsigned ua, ub;                                              // the segment's 32-column units within the row block
      f_segment_units(d, sbrb, nT, kfirst, nsegs, srel, ua, ub);
      const unsigned ta = (ua >> 1) + pe * kFSubTiles;            // first 64-column tile of the entry
      const unsigned tb = min(ta + kFSubTiles, ((ub - 1u) >> 1) + 1u);   // the entry's tiles [ta, tb)
      // the segment's first and last tile may be halves
      const unsigned c_lo = max(ta * kFRowTile, ua * kFUnit), c_hi = min(tb * kFRowTile, ub * kFUnit);
#pragma unroll 1
      for (unsigned cb0 = c_lo; cb0 < c_hi; cb0 += 128u) {
        float4 q[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) q[kk] = scq[min(cb0 + (unsigned)lane + 32u * kk, c_hi - 1u)];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) eval_col(q[kk], min(cb0 + (unsigned)lane + 32u * kk, c_hi - 1u));
      }
    }
  }
  const unsigned mn = __reduce_min_sync(full, bd);
  const unsigned jm = __reduce_min_sync(full, bd == mn ? bj : 0xffffffffu);
  const int owner = __ffs(__ballot_sync(full, bd == mn && bj == jm)) - 1;
  const float qx = __shfl_sync(full, bq.x, owner), qy = __shfl_sync(full, bq.y, owner);
  const float qz = __shfl_sync(full, bq.z, owner);
  out_d = mn; out_j = jm; out_q = make_float4(qx, qy, qz, 0.f);
}

// Slow path of ONE column (out of line, see row_slow_item): every row block's first record within the guard is a
// candidate group of 16 rows; a SECOND record within the guard means the row block may hold more than the recorded two,
// so all of its 512 rows are re-evaluated.
__device__ __noinline__ void col_slow_item(const float* __restrict__ colrec, const float4* __restrict__ rowpack,
                                           unsigned* stats, int nRB, int Mp, int Npad, int N, int sb, int sj, float sthr,
                                           float qx, float qy, float qz, unsigned& out_d, unsigned& out_i) {
  const unsigned full = 0xffffffffu;
  const float inf = __int_as_float(0x7f800000);
  const int lane = lane_id(), half = lane >> 4, hl = lane & 15;
  const float4 q = make_float4(qx, qy, qz, 0.f);
  const float* rec = colrec + ((size_t)sb * nRB * Mp + sj) * kFColTop;   // record e of this column: rec[(e / 2) * 2 Mp + e % 2]
  unsigned bd = 0xffffffffu, bi = 0xffffffffu;
  const float4* srp = rowpack + (size_t)sb * Npad;
  auto eval_row = [&](const float4 pr, unsigned row) {            // padded rows sit at the origin: excluded by index
    if (row < (unsigned)N) {
      const unsigned d = __float_as_uint(sqdist_q(pr.x, pr.y, pr.z, q));
      if (d < bd || (d == bd && row < bi)) { bd = d; bi = row; }
    }
  };
  const int nE = nRB * kFColTop;
  for (int base = 0; base < nE; base += 32) {
    const int e = base + lane;                                     // entry e = row block e / 2, record e % 2
    const float p = e < nE ? rec[(size_t)(e >> 1) * (kFColTop * Mp) + (e & 1)] : inf;
    const unsigned hits = __ballot_sync(full, p <= sthr);
    // a row block whose SECOND record is within the guard may hold more candidates than the two recorded: rescan it
    unsigned blocks = hits & 0xaaaaaaaau;
    unsigned groups = hits & 0x55555555u & ~(blocks >> 1);        // first records of the other row blocks
    const unsigned first_row = ((unsigned)(e >> 1) * 32u + (__float_as_uint(p) & 31u)) * kFRowsPerThread;
    const int ngroup = __popc(groups);
    for (int r = 0; r < ngroup; r += 8) {            // eight 16-row groups per batch: half-warp <-> group, lane <-> row
      float4 pr[4];
      unsigned row[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = min(r + 2 * k + half, ngroup - 1);
        const unsigned from = __fns(groups, 0, idx + 1);
        row[k] = __shfl_sync(full, first_row, (int)from) + (unsigned)hl;
        pr[k] = srp[row[k]];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) eval_row(pr[k], row[k]);
    }
    while (blocks) {                                  // all 512 rows of the block: 128 per batch
      const unsigned hrb = (unsigned)(base + __ffs(blocks) - 1) >> 1; blocks &= blocks - 1u;
      if (lane == 0) atomicAdd(stats + 2, 1u);
#pragma unroll 1
      for (unsigned r0 = hrb * kFRowsPerWarp; r0 < (hrb + 1u) * kFRowsPerWarp; r0 += 128u) {
        float4 pr[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) pr[kk] = srp[r0 + (unsigned)lane + 32u * kk];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) eval_row(pr[kk], r0 + (unsigned)lane + 32u * kk);
      }
    }
  }
  {
    const unsigned mn = __reduce_min_sync(full, bd);
    const unsigned ii = __reduce_min_sync(full, bd == mn ? bi : 0xffffffffu);
    bd = mn; bi = ii;
  }
  out_d = bd; out_i = bi;
}

// Work distribution: the ceil(B*N/32) + ceil(B*M/32) warp items are dealt out evenly over the grid (CTA c takes
// total/grid items, the first total%grid CTAs one more; at most one per warp).  For a single-wave problem the grid is exactly
// SMs x 4 resident CTAs, so every SM carries the same load wherever the block scheduler (or an early programmatic
// launch behind the search's tail) places the CTAs: with one CTA per 8 items, 512 CTAs on 148 SMs leave some SMs with
// 4 x 8 items against an average of 27.7.
__global__ void __launch_bounds__(kFinWarps * 32, 4) chamfer_fresolve_kernel(const __grid_constant__ FGeometry g, FWorkspace w, FinalizeArgs f,
                                                                             unsigned wbase, unsigned wrem) {
  __shared__ double lsum[kFinWarps];
  // per-warp staging of the 32 items' broadcast data
  __shared__ float4 sA[kFinWarps][32];
  __shared__ int2 sB[kFinWarps][32];
  __shared__ float4 sQ[kFinWarps][32];      // per item: the candidate that passed the re-filter (fast_items)
  // 32-bit index math throughout: B*Npad, B*Mp and the unit count are < 2^31 (f_sizes_ok); a 64-bit division
  // costs ~80 emulated instructions per lane and there were five of them
  const unsigned nrow = (unsigned)g.B * (unsigned)g.N, ncol = (unsigned)g.B * (unsigned)g.M;
  const unsigned row_warps = (nrow + 31u) / 32u;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned total_warps = row_warps + (ncol + 31u) / 32u;
  // CTA c takes wbase = total / grid items, the first wrem = total % grid CTAs one more (computed on the host: a 64-bit
  // division here costs ~100 instructions per thread)
  const unsigned w_begin = blockIdx.x * wbase + min(blockIdx.x, wrem);
  const unsigned w_end = w_begin + wbase + (blockIdx.x < wrem ? 1u : 0u);
  // warps beyond the CTA's share take the out-of-range item (valid == false on every lane)
  const unsigned wi = w_begin + (unsigned)warp < w_end ? w_begin + (unsigned)warp : total_warps;
  const float coef_grad = 2.f * f.coef_loss;
  const unsigned full = 0xffffffffu;
  const float inf = __int_as_float(0x7f800000);
  float myd = 0.f, wb = 1.f;
  bool valid = false;
  RMA_TL(w.stats, 4);
  // The dependency wait sits in the two branches below, behind everything that needs only the prep kernel's output
  // (index arithmetic, the item's own pack entry, the segment list).  This grid cannot be launched before every search
  // CTA has passed its own dependency wait, i.e. before prep has completed and flushed, so prep's output is in L2 by
  // then; the two loads in front of the wait read it there (ld.global.cg: no L1 line of an earlier step can answer).
  // A CTA that becomes resident while the last search CTAs still run has that round trip behind it when they are
  // done: C2 step -0.4 us, C4 -3.2 us.
  pdl_launch_dependents();   // the backward combine may queue behind us (it waits for our completion itself)

  // Row items and column items alternate over the warp index (row warps read ~10 records per item, column warps
  // 2 per row block: interleaved, every CTA - and so every SM - carries the same mix and they finish together).
  const unsigned col_warps = total_warps - row_warps, both = 2u * min(row_warps, col_warps);
  const bool is_row = wi < both ? (wi & 1u) == 0u : row_warps > col_warps;
  const unsigned widx = wi < both ? wi >> 1 : wi - (both >> 1);
  if (is_row) {
    // ------------------------------------------------------------ rows: lane <-> row, candidates = 16-column groups
    const unsigned t = widx * 32u + lane;
    valid = t < nrow;
    int b = 0, i = 0;
    float4 rp = make_float4(0.f, 0.f, 0.f, 0.f);
    float M1 = inf, M2 = inf;
    int T1 = 0;
    const float* top = w.rowtop;
    int nseg = 0;
    if (valid) {
      b = (int)fdiv(t, (unsigned)g.N, g.magicN); i = (int)(t - (unsigned)b * (unsigned)g.N);
      rp = __ldcg(w.rowpack + (size_t)b * g.Npad + i);   // before the dependency wait: from L2, where prep's flush put it
      const int rbc = i >> g.rc_shift, brb = b * g.nRBC + rbc;
      int kfirst, nsegs;
      f_rowblock_segments(g.d, (unsigned)brb, g.nT, kfirst, nsegs);
      top = w.rowtop + ((size_t)brb * g.Smax * g.P * 3) * g.RC + (i - rbc * g.RC);
      nseg = nsegs * g.P;  // entries: P per segment, consecutive
    }
    pdl_wait();              // from here on: what the search wrote
    RMA_TL(w.stats, 5);
    if (valid) {
      if (g.W == 4) scan_row_entries<4 * kFRowsPerWarp>(top, nseg, M1, T1, M2);
      else if (g.W == 2) scan_row_entries<2 * kFRowsPerWarp>(top, nseg, M1, T1, M2);
      else scan_row_entries<kFRowsPerWarp>(top, nseg, M1, T1, M2);
    }
    const float thr = row_threshold(rp.w, M1);
#if defined(RMA_B200_EXPERIMENTS) && (defined(RMA_EXP_NO_SLOW) || defined(RMA_EXP_NO_SLOW_ROWS))   // timing experiment only (tools/resolve_variants.py): wrong results for the slow-path items
    const bool fast = valid;
#else
    const bool fast = valid && M2 > thr;
#endif
    unsigned best_d = 0xffffffffu, best_j = 0;
    float4 best_q = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* cq = w.colpack + (size_t)b * g.Mp;

    // Fast path: the one candidate group (tile T1, position in the two tag bits of M1)
    sA[warp][lane] = make_float4(rp.x, rp.y, rp.z, 0.f);
    sB[warp][lane] = fast ? make_int2(b * g.Mp + T1 * kFRowTile + (int)(__float_as_uint(M1) & kFRowTagMask) * kFGroup,
                                      __float_as_int(thr))
                          : make_int2(0, (int)0xff800000u);
    __syncwarp();
    fast_items<false>(w.colpack, sA[warp], sB[warp], sQ[warp], lane, best_d, best_j, best_q);
    best_j -= (unsigned)(b * g.Mp);

    // Slow path (guard failed: several groups within the rounding error of the smallest).  Every candidate group lies
    // in an entry (a run of up to 16 tiles of one search segment) whose smallest group minimum is within the guard: an
    // entry whose second smallest is not contributes its one recorded group, otherwise all of its groups; a half-warp
    // re-filters and evaluates one group per round.  The lowest column among the lanes that hold the minimum =
    // torch.min's first index.
    unsigned slow = __ballot_sync(full, valid && !fast);
    if (slow && lane == 0) atomicAdd(w.stats + 0, (unsigned)__popc(slow));
    while (slow) {
      const int src = __ffs(slow) - 1; slow &= slow - 1;
      unsigned sd, sj;
      float4 sq;
      row_slow_item(w.rowtop, w.colpack, g.RC, g.nRBC, g.nT, &g.d, g.Smax, g.P, g.Mp, __shfl_sync(full, b, src),
                    __shfl_sync(full, i, src), __shfl_sync(full, thr, src), __shfl_sync(full, rp.x, src),
                    __shfl_sync(full, rp.y, src), __shfl_sync(full, rp.z, src), sd, sj, sq);
      if (lane == src) { best_d = sd; best_j = sj; best_q = sq; }
    }

    myd = __uint_as_float(best_d);
    wb = (valid && f.batch_weight) ? f.batch_weight[b] : 1.f;
    // no candidate at all happens only for NaN coordinates (every comparison fails): NaN distance, index 0
    const bool none = best_d == 0xffffffffu;
    const int j = none ? 0 : (int)best_j;
    if (valid) {
      if (f.dist1) f.dist1[t] = myd;
      if (f.idx1) f.idx1[t] = j;
    }
    if (f.own_x || f.scat_y) {
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (valid) {
        const float4 q = none ? cq[0] : best_q;
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
    const unsigned t = widx * 32u + lane;
    valid = t < ncol;
    int b = 0, j = 0;
    float4 cq = make_float4(0.f, 0.f, 0.f, 0.f);
    float M1 = inf, M2 = inf;
    int R1 = 0;
    if (valid) {
      b = (int)fdiv(t, (unsigned)g.M, g.magicM); j = (int)(t - (unsigned)b * (unsigned)g.M);
      cq = __ldcg(w.colpack + (size_t)b * g.Mp + j);     // before the dependency wait: from L2
    }
    pdl_wait();              // from here on: what the search wrote
    if (valid) {
      const float2* rec = reinterpret_cast<const float2*>(w.colrec) + (size_t)b * g.nRB * g.Mp + j;
      constexpr int kRbBatch = 8;                   // records of 8 row blocks in flight (C2: all of them)
      for (int r0 = 0; r0 < g.nRB; r0 += kRbBatch) {
        float2 a[kRbBatch];
        const float2* rp_ = rec + (size_t)r0 * g.Mp;
#pragma unroll
        for (int k = 0; k < kRbBatch; ++k) {
          const bool on = r0 + k < g.nRB;
          a[k] = *(on ? rp_ : rec);
          a[k] = on ? a[k] : make_float2(inf, inf);
          rp_ += g.Mp;
        }
#pragma unroll
        for (int k = 0; k < kRbBatch; ++k) {
          M2 = fminf(fminf(M2, a[k].y), fmaxf(M1, a[k].x));
          if (a[k].x < M1) { M1 = a[k].x; R1 = r0 + k; }
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
    float4 best_p = make_float4(0.f, 0.f, 0.f, 0.f);

    // Fast path (one candidate group of 16 rows).  As for the rows: inside the candidate group the guard applies per
    // ROW (g_i = fl(f + |x_i|^2) <= thr for every exact minimiser; padded rows carry |x|^2 = 1e30).
    sA[warp][lane] = cq;
    sB[warp][lane] = fast ? make_int2(b * g.Npad + G1, __float_as_int(thr)) : make_int2(0, (int)0xff800000u);
    __syncwarp();
    fast_items<true>(w.rowpack, sA[warp], sB[warp], sQ[warp], lane, best_d, best_i, best_p);
    best_i -= (unsigned)(b * g.Npad);

    // Slow path: every row block's smallest entry within the guard is a candidate group; a second-smallest entry
    // within the guard means the row block may hold more than the recorded two, so all of its 512 rows are re-evaluated.
    unsigned slow = __ballot_sync(full, valid && !fast);
    if (slow && lane == 0) atomicAdd(w.stats + 1, (unsigned)__popc(slow));
    while (slow) {
      const int src = __ffs(slow) - 1; slow &= slow - 1;
      unsigned sd, si_;
      col_slow_item(w.colrec, w.rowpack, w.stats, g.nRB, g.Mp, g.Npad, g.N, __shfl_sync(full, b, src),
                    __shfl_sync(full, j, src), __shfl_sync(full, thr, src), __shfl_sync(full, cq.x, src),
                    __shfl_sync(full, cq.y, src), __shfl_sync(full, cq.z, src), sd, si_);
      if (lane == src) {
        best_d = sd; best_i = si_;
        if (sd != 0xffffffffu) best_p = w.rowpack[(size_t)b * g.Npad + si_];
      }
    }

    myd = __uint_as_float(best_d);
    wb = (valid && f.batch_weight) ? f.batch_weight[b] : 1.f;
    const bool none = best_d == 0xffffffffu;
    const int i = none ? 0 : (int)best_i;
    if (valid) {
      if (f.dist2) f.dist2[t] = myd;
      if (f.idx2) f.idx2[t] = i;
    }
    if (f.own_y || f.scat_x) {
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (valid) {
        const float4 p = none ? w.rowpack[(size_t)b * g.Npad] : best_p;
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
    const double tot = loss_partial<kFinWarps>(valid, myd, wb, f.coef_loss, lsum);
    if (threadIdx.x == 0) {
#ifdef RMA_LOSS_FLOAT_ATOMIC   // tools/ A/B only: the round-1 float atomics (run-to-run differences of ~1e-6 relative)
      atomicAdd(f.loss, (float)tot);
#else
      loss_commit(reinterpret_cast<double*>(w.stats + 48), tot, f.loss);
#endif
    }
  }
  RMA_TL(w.stats, 6);
}

// ---------------------------------------------------------------------------------------------------------------
// host side (called from the dispatcher in chamfer.cu)
// ---------------------------------------------------------------------------------------------------------------
bool f_sizes_ok(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return false;
  if ((long long)N > (1LL << 30) || (long long)M > (1LL << 30)) return false;
  const FGeometry g = f_make_geometry(B, N, M, 148, 2, 1);
  if ((long long)B * g.Npad >= (1LL << 31) || (long long)B * g.Mp >= (1LL << 31)) return false;
  if ((long long)B * g.nRBC >= (1LL << 24)) return false;
  if ((long long)B * g.nRBC * g.nT >= (1LL << 31)) return false;   // unit indices are 32-bit in the resolve
  return true;
}

// Bytes of the filter's group records alone (the dispatcher's size criterion; device independent): B*N*M/64.
size_t f_record_bytes(int B, int N, int M) {
  const FGeometry g = f_make_geometry(B, N, M, 148, 2, 1);
  return ((size_t)g.B * g.nRB * kFColTop * g.Mp) * sizeof(float);
}

// SM count of the current device (and the one-time shared-memory opt-in of the search kernels).  Immutable device
// properties, memoised per device (the queries cost ~10 us of host time, which matters for eager callers at ~90 us per
// step).  The memo is a pure cache of constants: one atomic word per device (release / acquire), so concurrent first
// calls from several threads at worst query twice and store the same value.
static int f_caps(int* sms) {
  static std::atomic<unsigned> memo[64];   // 0 = empty, else the SM count
  int dev = 0;
  RMA_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64) {
    const unsigned m = memo[dev].load(std::memory_order_acquire);
    if (m) { *sms = (int)m; return 0; }
  }
  RMA_CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  // a physical search CTA carries 8 / kW virtual CTAs, each with its own f_smem(kW) bytes: ~189 KB for every kW
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_fsearch_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * f_smem(4))));
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_fsearch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * f_smem(2))));
  RMA_CUDA_TRY(cudaFuncSetAttribute(chamfer_fsearch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * f_smem(1))));
  if (*sms < 1) *sms = 1;
  if (dev >= 0 && dev < 64) memo[dev].store((unsigned)*sms, std::memory_order_release);
  return 0;
}

// Geometry for the current device (the dealing depends on SM count x virtual CTAs per SM).  Re-derived per call: the
// library keeps no state, and every stage of one forward derives the identical geometry.
static int f_geometry(int B, int N, int M, int unit_run, FGeometry* g, int* occ_out) {
  int sms = 0;
  if (int e = f_caps(&sms)) return e;
  const int W = f_warps_for((N + kFRowsPerWarp - 1) / kFRowsPerWarp);
  const int occ = (kFSearchThreads / 32) / W;   // virtual CTAs per physical CTA = per SM
  *g = f_make_geometry(B, N, M, sms, occ, unit_run);
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
                  int y_period, void* workspace, size_t workspace_bytes, float* zero_loss, float4* zero_gx,
                  float4* zero_gy, void* stream_) {
  if (!x || !y) return RMA_E_BADARG;
  FGeometry g; FWorkspace w;
  if (int e = f_setup(B, N, M, unit_run, workspace, workspace_bytes, &g, &w)) return e;
  g.yper = y_period;
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
  if (ctas) *ctas = g.nC;
  if (threads) *threads = 32 * g.W;
  if (splits) *splits = g.maxrun;     // for this path: 32-column units of the longest run
  if (ctas_per_sm) *ctas_per_sm = occ;
  return RMA_OK;
}

int f_launch_search(int B, int N, int M, int unit_run, void* workspace, size_t workspace_bytes, void* stream_) {
  FGeometry g; FWorkspace w;
  if (int e = f_setup(B, N, M, unit_run, workspace, workspace_bytes, &g, &w)) return e;
  FSearchArgs a;
  a.rowsoa = w.rowsoa; a.colpair = w.colpair; a.rowtop = w.rowtop; a.colrec = w.colrec; a.stats = w.stats;
  a.N = N; a.M = M; a.nRB = g.nRB; a.Npad = g.Npad; a.nRBC = g.nRBC; a.nT = g.nT; a.Mp = g.Mp;
  a.nC = g.nC; a.d = g.d; a.Smax = g.Smax; a.P = g.P; a.units = g.units();
  if ((a.units + g.d.U - 1) / g.d.U > 0x7fffffffLL) return RMA_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  const unsigned grid = (unsigned)g.G;
  if (g.W == 4)
    RMA_CUDA_TRY(launch_dep(pdl_site(2), chamfer_fsearch_kernel<4>, grid, (unsigned)kFSearchThreads, 2 * f_smem(4), st, a));
  else if (g.W == 2)
    RMA_CUDA_TRY(launch_dep(pdl_site(2), chamfer_fsearch_kernel<2>, grid, (unsigned)kFSearchThreads, 4 * f_smem(2), st, a));
  else
    RMA_CUDA_TRY(launch_dep(pdl_site(2), chamfer_fsearch_kernel<1>, grid, (unsigned)kFSearchThreads, 8 * f_smem(1), st, a));
  return RMA_OK;
}

int f_launch_resolve(int B, int N, int M, int unit_run, const FinalizeArgs& f, void* workspace,
                     size_t workspace_bytes, void* stream_) {
  FGeometry g; FWorkspace w;
  if (int e = f_setup(B, N, M, unit_run, workspace, workspace_bytes, &g, &w)) return e;
  const long long warps = ((long long)B * N + 31) / 32 + ((long long)B * M + 31) / 32;
  int sms = 0;
  if (int e = f_caps(&sms)) return e;
  long long grid = grid_for(warps, kFinWarps);            // one CTA per 8 items when that is more than one wave
  const long long wave = (long long)sms * 4;              // resident resolve CTAs (launch bounds: 4 per SM)
  if (grid <= wave) grid = warps < wave ? warps : wave;   // single wave: the same number of CTAs on every SM
  RMA_CUDA_TRY(launch_dep(pdl_site(4), chamfer_fresolve_kernel, (unsigned)grid, (unsigned)(kFinWarps * 32), 0,
                          (cudaStream_t)stream_, g, w, f, (unsigned)(warps / grid), (unsigned)(warps % grid)));
  return RMA_OK;
}

// Host-only self-check of the dealing for a device with `sms` SMs: every unit is covered by exactly one run, no run of
// the aligned mode crosses a row block, segment indices and entry counts stay inside what the workspace is carved for,
// and the resolve's view of a row block's segments (f_rowblock_segments / f_segment_units) is the search's.  Returns
// the number of inconsistencies (0 = consistent), -1 for unsupported sizes; load[0..1] = min / max units per SM.
int f_dealing_selfcheck(int B, int N, int M, int sms, int unit_run, int* mode_out, long long* load2_out) {
  if (sms < 1 || !f_sizes_ok(B, N, M)) return -1;
  const int W = f_warps_for((N + kFRowsPerWarp - 1) / kFRowsPerWarp);
  const int occ = (kFSearchThreads / 32) / W;
  const FGeometry g = f_make_geometry(B, N, M, sms, occ, unit_run);
  const long long units = g.units();
  if (units > (1LL << 26)) return -1;
  std::vector<unsigned char> cover((size_t)units, 0);
  std::vector<long long> load((size_t)g.G, 0);
  int bad = 0;
  for (int layer = 0; layer < occ || (long long)layer * g.G < g.nC; ++layer) {
    if ((long long)layer * g.G >= g.nC) break;
    for (int block = 0; block < g.G; ++block) {
      if (block + (long long)layer * g.G >= g.nC) continue;
      if (layer >= occ) { ++bad; continue; }          // more layers than a physical CTA has warp groups
      long long unit, unit_end; int seg_aligned;
      f_virtual_run(g.d, g.nT, g.nC, units, g.G, block, layer, unit, unit_end, seg_aligned);
      if (unit < 0 || unit_end > units || unit_end < unit || unit_end - unit > g.maxrun) { ++bad; continue; }
      if (unit_end == unit && unit_run <= 0) ++bad;   // only a forced run length may leave the last CTA empty
      load[(size_t)block] += unit_end - unit;
      for (long long u = unit; u < unit_end; ++u) if (++cover[(size_t)u] != 1) ++bad;
      int nsegments = 0;
      while (unit < unit_end) {                       // the kernel's walk over the run's segments
        const int brb = (int)(unit / g.nT), unit0 = (int)(unit - (long long)brb * g.nT);
        const long long left = unit_end - unit;
        const int nunit = (int)((long long)(g.nT - unit0) < left ? (long long)(g.nT - unit0) : left);
        unit += nunit;
        ++nsegments;
        int kfirst, nsegs;
        f_rowblock_segments(g.d, (unsigned)brb, g.nT, kfirst, nsegs);
        const int v = block + layer * g.G;
        const int seg = g.d.mode == 1 ? seg_aligned : v - kfirst;
        if (seg < 0 || seg >= nsegs || nsegs > g.Smax) { ++bad; continue; }
        unsigned ua, ub;
        f_segment_units(g.d, (unsigned)brb, g.nT, kfirst, nsegs, (unsigned)seg, ua, ub);
        if ((int)ua != unit0 || (int)ub != unit0 + nunit) ++bad;
        const int tiles = ((unit0 + nunit - 1) >> 1) - (unit0 >> 1) + 1;
        if (tiles > g.P * kFSubTiles) ++bad;
      }
      if (g.d.mode == 1 && nsegments != 1) ++bad;
    }
  }
  for (long long u = 0; u < units; ++u) if (cover[(size_t)u] != 1) ++bad;
  if (mode_out) *mode_out = g.d.mode;
  if (load2_out) {
    long long lo = load[0], hi = load[0];
    for (long long v : load) { lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
    load2_out[0] = lo; load2_out[1] = hi;
  }
  return bad;
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

#ifdef RMA_B200_TIMELINE
// tools/ only: copy the 8 (min, max) timestamp pairs to the host and reset them
int f_timeline_read(void* workspace, int B, int N, int M, int unit_run, unsigned long long* out16_host, void* stream_) {
  FGeometry g;
  if (int e = f_geometry(B, N, M, unit_run, &g, nullptr)) return e;
  FWorkspace w = f_carve(g, workspace);
  cudaStream_t st = (cudaStream_t)stream_;
  RMA_CUDA_TRY(cudaMemcpyAsync(out16_host, w.stats + 8, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  RMA_CUDA_TRY(cudaStreamSynchronize(st));
  unsigned long long init[16];
  for (int k = 0; k < 8; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
  RMA_CUDA_TRY(cudaMemcpyAsync(w.stats + 8, init, sizeof(init), cudaMemcpyHostToDevice, st));
  RMA_CUDA_TRY(cudaStreamSynchronize(st));
  return RMA_OK;
}
#endif

}  // namespace rma

#ifdef RMA_B200_TIMELINE
extern "C" int rma_debug_timeline_read(void* workspace, int B, int N, int M, unsigned long long* out16_host, void* stream) {
  return rma::f_timeline_read(workspace, B, N, M, 0, out16_host, stream);
}
#endif
