// FFMA2 operand forms on B200: how many issue cycles does the packed fp32x2 FMA take when the B and/or C operands
// are broadcast scalars (SASS: FFMA2 R, Ra.F32x2, Rb.F32, Rc.F32)?  16 independent chains per thread so latency is
// hidden; cycles per warp instruction per SM sub-partition from clock64 on CTA 0 and from the kernel time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_forms_microbench tools/ffma2_forms_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

enum { T_FULL, T_BCAST_B, T_BCAST_C, T_BCAST_BC, T_CHAIN3_BC, T_CHAIN3_FULL, T_SCALAR3, T_ADD2_FULL, T_ADD2_BCAST, T_COUNT };
static const char* names[T_COUNT] = {"ffma2 a*b+c all F32x2", "ffma2 b=bcast", "ffma2 c=bcast", "ffma2 b,c=bcast",
                                     "filter chain (3 ffma2, bcast b / first c)", "filter chain (3 ffma2, all F32x2)",
                                     "filter chain scalar (6 ffma)", "fadd2 all F32x2", "fadd2 b=bcast"};
static const int instr_per_chain[T_COUNT] = {1, 1, 1, 1, 3, 3, 6, 1, 1};

template <int T>
__global__ void __launch_bounds__(128) kern(float* out, int iters, float seed, long long* cyc) {
  u64 a[16], b[16], acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float tv = threadIdx.x * 1e-7f;
    a[i] = pack2(1.0f + (seed + i) * 1e-8f + tv, 1.0f - (seed + i) * 1e-8f + tv);
    b[i] = pack2(seed * 1e-3f + i * 1e-6f + tv, seed * 2e-3f + i * 1e-6f + tv);
    acc[i] = pack2(tv, -tv);
  }
  float s0 = seed * 1e-3f + threadIdx.x * 1e-7f, s1 = seed * 3e-3f - threadIdx.x * 1e-7f, s2 = seed * 1e-4f, s3 = seed * 7e-4f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // the scalars change every step so that nothing can be hoisted
      s0 += 1e-9f; s1 -= 1e-9f;
      const u64 bs0 = pack2(s0, s0), bs1 = pack2(s1, s1), bs2 = pack2(s2, s2), bs3 = pack2(s3, s3);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (T == T_FULL) acc[i] = fma2(a[i], b[i], acc[i]);
        if (T == T_BCAST_B) acc[i] = fma2(a[i], bs0, acc[i]);
        if (T == T_BCAST_C) acc[i] = fma2(acc[i], b[i], bs0);
        if (T == T_BCAST_BC) acc[i] = fma2(acc[i], bs0, bs1);
        if (T == T_CHAIN3_BC) { u64 t = fma2(a[i], bs0, bs3); t = fma2(b[i], bs1, t); acc[i] = fma2(acc[i], bs2, t); }
        if (T == T_CHAIN3_FULL) { u64 t = fma2(a[i], b[(i + 1) & 15], acc[(i + 3) & 15]); t = fma2(b[i], a[(i + 5) & 15], t); acc[i] = fma2(acc[i], b[(i + 7) & 15], t); }
        if (T == T_SCALAR3) {
          float al, ah, bl, bh, cl, ch; unpack2(a[i], al, ah); unpack2(b[i], bl, bh); unpack2(acc[i], cl, ch);
          float tl = __fmaf_rn(al, s0, s3), th = __fmaf_rn(ah, s0, s3);
          tl = __fmaf_rn(bl, s1, tl); th = __fmaf_rn(bh, s1, th);
          cl = __fmaf_rn(cl, s2, tl); ch = __fmaf_rn(ch, s2, th);
          acc[i] = pack2(cl, ch);
        }
        if (T == T_ADD2_FULL) acc[i] = add2(acc[i], b[i]);
        if (T == T_ADD2_BCAST) acc[i] = add2(acc[i], bs0);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { float l2, h2; unpack2(acc[i], l2, h2); s += l2 + h2; }
  if (s == 12345.f) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int T>
static void run(int sms, float* out, long long* cyc) {
  const int iters = 4000;
  for (int wps = 1; wps <= 2; ++wps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<T><<<sms * wps, 128>>>(out, 10, 1.0f, cyc);
    cudaEventRecord(e0);
    kern<T><<<sms * wps, 128>>>(out, iters, 1.0f, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double groups = (double)iters * 4 * 16;
    printf("{\"test\": \"%s\", \"warps_per_smsp\": %d, \"cta0_cycles_per_group\": %.3f, \"instr_per_group\": %d, \"cycles_per_instr\": %.3f, \"ms\": %.4f}\n",
           names[T], wps, (double)c / groups, instr_per_chain[T], (double)c / groups / instr_per_chain[T] / 1.0, ms);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  int n = p.multiProcessorCount;
  run<T_FULL>(n, out, cyc); run<T_BCAST_B>(n, out, cyc); run<T_BCAST_C>(n, out, cyc); run<T_BCAST_BC>(n, out, cyc);
  run<T_CHAIN3_BC>(n, out, cyc); run<T_CHAIN3_FULL>(n, out, cyc); run<T_SCALAR3>(n, out, cyc);
  run<T_ADD2_FULL>(n, out, cyc); run<T_ADD2_BCAST>(n, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
