// Dispatch / register-file microbenchmarks for B200 (sm_100a): how many cycles per warp-instruction do packed
// fp32x2 ops and FMNMX3 cost when their operands are distinct registers (no operand-reuse cache hits), and do
// ALU-pipe minima overlap with FMA-pipe packed ops?  Results drive DESIGN.md §4.3.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dispatch_microbench dispatch_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

enum { T_FADD2_DISTINCT, T_FADD2_SHARED, T_FFMA2_SQ, T_FMNMX3_DISTINCT, T_FMNMX_2IN, T_MIX_FADD2_MIN3, T_MIX_FFMA2_MIN3,
       T_SCALAR_FADD_DISTINCT, T_SCALAR_FFMA_SQ, T_MIX_FADD2_3_MIN3_1, T_MIX_SCALAR6_MIN1, T_FADD2_BCAST, T_COUNT };
static const char* names[T_COUNT] = {"fadd2_distinct", "fadd2_shared_b", "ffma2_a*a+d", "fmnmx3_distinct", "fmnmx_2in",
  "mix_1fadd2_1min3", "mix_1ffma2_1min3", "scalar_fadd_distinct", "scalar_ffma_a*a+d", "mix_3fadd2_1min3",
  "mix_6scalar_1min", "fadd2_bcast_scalar_b"};
// warp-instructions per inner iteration (16 chains)
static const int instr_per_iter[T_COUNT] = {16, 16, 16, 16, 16, 32, 32, 16, 16, 64, 112, 16};

template <int T>
__global__ void __launch_bounds__(128) kern(float* out, int iters, float seed, long long* cyc) {
  u64 a[16], b[16];
  float m[16], p[16], q[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float tv = threadIdx.x * 1e-7f;   // per-thread values: keeps operands in vector registers (not UR)
    a[i] = pack2(seed + i + tv, seed - i + tv);
    b[i] = pack2(seed * 1e-3f + i * 1e-6f + tv, seed * 2e-3f + i * 1e-6f + tv);
    m[i] = 1e30f + i + tv; p[i] = 1e29f + seed * i + tv * 1e29f; q[i] = 1e28f - seed * i + tv * 1e28f;
  }
  const u64 bs = pack2(seed * 1e-3f + threadIdx.x * 1e-7f, seed * 1e-3f + threadIdx.x * 1e-7f);
  float sb = seed * 1e-3f + threadIdx.x * 1e-7f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (T == T_FADD2_DISTINCT) a[i] = add2(a[i], b[i]);
        if (T == T_FADD2_SHARED) a[i] = add2(a[i], bs);
        if (T == T_FADD2_BCAST) a[i] = add2(a[i], pack2(sb, sb));
        if (T == T_FFMA2_SQ) b[i] = fma2(a[i], a[i], b[i]);
        if (T == T_FMNMX3_DISTINCT) m[i] = min3(m[i], p[i], q[i]);
        if (T == T_FMNMX_2IN) m[i] = fminf(m[i], p[i]);
        if (T == T_MIX_FADD2_MIN3) { a[i] = add2(a[i], b[i]); m[i] = min3(m[i], p[i], q[i]); }
        if (T == T_MIX_FFMA2_MIN3) { b[i] = fma2(a[i], a[i], b[i]); m[i] = min3(m[i], p[i], q[i]); }
        if (T == T_SCALAR_FADD_DISTINCT) m[i] = __fadd_rn(m[i], p[i]);
        if (T == T_SCALAR_FFMA_SQ) m[i] = __fmaf_rn(p[i], p[i], m[i]);
        if (T == T_MIX_FADD2_3_MIN3_1) {
          // the search kernel's ratio: 3 packed ops : 1 min3, min3 consumes the packed result halves
          a[i] = add2(a[i], b[i]);
          u64 t = mul2(a[i], a[i]);
          b[i] = fma2(a[i], a[i], t);
          float lo, hi; unpack2(b[i], lo, hi);
          m[i] = min3(m[i], lo, hi);
        }
        if (T == T_MIX_SCALAR6_MIN1) {
          float dx = __fadd_rn(p[i], q[i]), dy = __fadd_rn(p[i], m[i]), dz = __fadd_rn(q[i], sb);
          float d = __fmul_rn(dx, dx);
          d = __fmaf_rn(dy, dy, d);
          d = __fmaf_rn(dz, dz, d);
          m[i] = fminf(m[i], d);
        }
      }
    }
    sb += 1e-9f;
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { float lo, hi; unpack2(a[i], lo, hi); float l2, h2; unpack2(b[i], l2, h2); s += lo + hi + l2 + h2 + m[i]; }
  if (s == 12345.f) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int T>
static void run(int sms, float* out, long long* cyc) {
  const int iters = 4000;
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    kern<T><<<sms * warps_per_smsp, 128>>>(out, iters, 1.5f, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<T><<<sms * warps_per_smsp, 128>>>(out, iters, 1.5f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double n = (double)iters * 4 * instr_per_iter[T];
    double per = (double)c / n;
    // steady-state SMSP throughput from wall time at the max SM clock (1.965 GHz): cycles per warp-instruction
    // per sub-partition with `warps_per_smsp` warps sharing it
    double per_time = ms * 1e-3 * 1.965e9 / (n * warps_per_smsp);
    printf("{\"test\": \"%s\", \"warps_per_smsp\": %d, \"cta0_cycles_per_warp_instr\": %.3f, "
           "\"time_cycles_per_smsp_instr\": %.3f, \"ms\": %.4f}\n",
           names[T], warps_per_smsp, per, per_time, ms);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  run<T_FADD2_DISTINCT>(p.multiProcessorCount, out, cyc);
  run<T_FADD2_SHARED>(p.multiProcessorCount, out, cyc);
  run<T_FADD2_BCAST>(p.multiProcessorCount, out, cyc);
  run<T_FFMA2_SQ>(p.multiProcessorCount, out, cyc);
  run<T_FMNMX3_DISTINCT>(p.multiProcessorCount, out, cyc);
  run<T_FMNMX_2IN>(p.multiProcessorCount, out, cyc);
  run<T_MIX_FADD2_MIN3>(p.multiProcessorCount, out, cyc);
  run<T_MIX_FFMA2_MIN3>(p.multiProcessorCount, out, cyc);
  run<T_SCALAR_FADD_DISTINCT>(p.multiProcessorCount, out, cyc);
  run<T_SCALAR_FFMA_SQ>(p.multiProcessorCount, out, cyc);
  run<T_MIX_FADD2_3_MIN3_1>(p.multiProcessorCount, out, cyc);
  run<T_MIX_SCALAR6_MIN1>(p.multiProcessorCount, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
