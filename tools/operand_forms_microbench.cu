// Operand forms of the packed fp32x2 instructions in the shapes the filter search uses (B200, sm_100a): what does one
// FFMA2 / FADD2 cost when the repeated operand is a 64-bit pair or a .F32 broadcast scalar, in runs that share that
// operand (operand-reuse latch) or interleaved with FMNMX3?  Every test runs 16 independent accumulators; the
// shared operands are reloaded from shared memory every trip (as the search kernel's column operands are), so nothing
// is hoisted.  Cycles per instruction per SM sub-partition from the kernel time at 1 and 2 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/operand_forms_microbench tools/operand_forms_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

enum { T_COLPAIR_ROWSCALAR, T_ROWPAIR_COLDUP, T_ROWPAIR_COLSCALAR, T_ADD_SCALAR, T_ADD_PAIR, T_COLPAIR_ROWSCALAR_MIN,
       T_ROWPAIR_COLDUP_MIN, T_MIN3_ONLY, T_MIN2_ONLY, T_CHAIN_COLPAIR, T_CHAIN_ROWPAIR, T_COUNT };
static const char* names[T_COUNT] = {
    "ffma2 run16: Qpair(shared) x row.F32 + acc   [search kernel's form]",
    "ffma2 run16: Xpair(rows) x Qdup pair(shared) + acc   [all F32x2]",
    "ffma2 run16: Xpair(rows) x q.F32(shared) + acc",
    "fadd2 run16: acc + u.F32 (distinct scalars)",
    "fadd2 run16: acc + Upair (distinct pairs)",
    "ffma2 Qpair x row.F32 + acc interleaved 1:1 with fmnmx3",
    "ffma2 Xpair x Qdup + acc interleaved 1:1 with fmnmx3",
    "fmnmx3 only (m = min3(m, lo, hi) of a pair)",
    "fmnmx x2 only (m = min(min(m, lo), hi))",
    "chain of 3 ffma2, Qpair x row.F32, runs of 16 (steps: seed WW shared)",
    "chain of 3 ffma2, Xpair x Qdup pair, runs of 16 (seed Wdup shared)"};
static const int instr[T_COUNT] = {16, 16, 16, 16, 16, 32, 32, 16, 32, 48, 48};

template <int T>
__global__ void __launch_bounds__(128) kern(float* out, int iters, float seed, const float4* __restrict__ tab) {
  __shared__ float4 sq[64];
  if (threadIdx.x < 64) sq[threadIdx.x] = tab[threadIdx.x];
  __syncthreads();
  u64 acc[16], X[16], U[16];
  float xs[16], us[16], m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float tv = threadIdx.x * 1e-7f;
    acc[i] = pack2(tv + i, -tv - i);
    X[i] = pack2(1.0f + (seed + i) * 1e-8f + tv, 1.0f - (seed + i) * 1e-8f + tv);
    U[i] = pack2(seed * 1e-3f + i * 1e-6f + tv, seed * 2e-3f + i * 1e-6f + tv);
    xs[i] = 1.0f + i * 1e-7f + tv;
    us[i] = i * 1e-9f + tv;
    m[i] = 1e30f + i + tv;
  }
  for (int it = 0; it < iters; ++it) {
    const float4 qa = sq[(it * 2) & 63], qb = sq[(it * 2 + 1) & 63];   // uniform address: broadcast LDS.128
    const u64 Q0 = pack2(qa.x, qa.y), Q1 = pack2(qa.z, qa.w), Q2 = pack2(qb.x, qb.y), WW = pack2(qb.z, qb.w);
    const u64 D0 = pack2(qa.x, qa.x), D1 = pack2(qa.y, qa.y), D2 = pack2(qa.z, qa.z), DW = pack2(qa.w, qa.w);
    if (T == T_COLPAIR_ROWSCALAR) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma2(Q0, pack2(xs[i], xs[i]), acc[i]);
    }
    if (T == T_ROWPAIR_COLDUP) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma2(X[i], Q0, acc[i]);
    }
    if (T == T_ROWPAIR_COLSCALAR) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma2(X[i], D0, acc[i]);
    }
    if (T == T_ADD_SCALAR) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = add2(acc[i], pack2(us[i] * qa.x, us[i] * qa.x));
    }
    if (T == T_ADD_PAIR) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = add2(acc[i], U[i]);
    }
    if (T == T_COLPAIR_ROWSCALAR_MIN) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float lo, hi; unpack2(U[i], lo, hi);
        m[i] = fmin3(m[i], lo, hi);
        acc[i] = fma2(Q0, pack2(xs[i], xs[i]), acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) U[i] = acc[(i + 5) & 15];
    }
    if (T == T_ROWPAIR_COLDUP_MIN) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float lo, hi; unpack2(U[i], lo, hi);
        m[i] = fmin3(m[i], lo, hi);
        acc[i] = fma2(X[i], Q0, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) U[i] = acc[(i + 5) & 15];
    }
    if (T == T_MIN3_ONLY || T == T_MIN2_ONLY) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float lo, hi; unpack2(acc[i], lo, hi);
        if (T == T_MIN3_ONLY) m[i] = fmin3(m[i], lo + qa.x, hi);
        else { m[i] = fminf(m[i], lo + qa.x); m[i] = fminf(m[i], hi); }
      }
    }
    if (T == T_CHAIN_COLPAIR) {
      u64 t[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = fma2(Q2, pack2(xs[i], xs[i]), WW);
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = fma2(Q1, pack2(us[i], us[i]), t[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = fma2(Q0, pack2(m[i], m[i]), t[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = t[i] ^ acc[i];   // consume (LOP3 x2 per pair, not counted)
    }
    if (T == T_CHAIN_ROWPAIR) {
      u64 t[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = fma2(X[i], Q2, WW);
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = fma2(U[i], Q1, t[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = fma2(acc[i], Q0, t[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = t[i];
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { float l2, h2; unpack2(acc[i], l2, h2); s += l2 + h2 + m[i]; unpack2(U[i], l2, h2); s += l2 + h2; }
  if (s == 12345.f) out[threadIdx.x] = s;
}

template <int T>
static void run(int sms, float* out, const float4* tab) {
  const int iters = 20000;
  for (int wps = 1; wps <= 2; ++wps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<T><<<sms * wps, 128>>>(out, 10, 1.0f, tab);
    cudaEventRecord(e0);
    kern<T><<<sms * wps, 128>>>(out, iters, 1.0f, tab);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * wps);       // SMSP cycles per trip and warp
    printf("{\"test\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_trip\": %.2f, \"counted_instr_per_trip\": %d, \"cycles_per_instr\": %.3f}\n",
           names[T], wps, cyc, instr[T], cyc / instr[T]);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; float4* tab;
  cudaMalloc(&out, 4096); cudaMalloc(&tab, 64 * 16);
  float4 h[64];
  for (int i = 0; i < 64; ++i) h[i] = make_float4(1.f + i * 1e-6f, 1.f - i * 1e-6f, 1e-3f * i, 1e-4f);
  cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice);
  int n = p.multiProcessorCount;
  run<T_COLPAIR_ROWSCALAR>(n, out, tab); run<T_ROWPAIR_COLDUP>(n, out, tab); run<T_ROWPAIR_COLSCALAR>(n, out, tab);
  run<T_ADD_SCALAR>(n, out, tab); run<T_ADD_PAIR>(n, out, tab); run<T_COLPAIR_ROWSCALAR_MIN>(n, out, tab);
  run<T_ROWPAIR_COLDUP_MIN>(n, out, tab); run<T_MIN3_ONLY>(n, out, tab); run<T_MIN2_ONLY>(n, out, tab);
  run<T_CHAIN_COLPAIR>(n, out, tab); run<T_CHAIN_ROWPAIR>(n, out, tab);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
