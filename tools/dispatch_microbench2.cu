// Which min flavour overlaps with packed FP32 ops on B200?  (follow-up of dispatch_microbench.cu)
// Each test: 16 independent chains; per chain and iteration one FFMA2 (b = a*a + b) followed by min op(s) on the
// fresh result halves, so nothing can be hoisted.  Build as dispatch_microbench.cu.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

enum { T_NONE, T_FMIN3, T_FMIN2x2, T_FMIN2x1, T_UMIN2x2, T_UMIN2x1, T_UMIN3, T_SMIN2x2, T_FSETP_SEL, T_LOP, T_COUNT };
static const char* names[T_COUNT] = {"ffma2_only", "ffma2+fmin3", "ffma2+2xfmin", "ffma2+1xfmin", "ffma2+2xumin", "ffma2+1xumin",
                                     "ffma2+umin3(nested)", "ffma2+2xsmin", "ffma2+fsetp+sel", "ffma2+2xlop"};
static const int instr_per_chain[T_COUNT] = {1, 2, 3, 2, 3, 2, 2, 3, 3, 3};

template <int T>
__global__ void __launch_bounds__(128) kern(float* out, int iters, float seed, long long* cyc) {
  u64 a[16], b[16];
  float m[16];
  unsigned um[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float tv = threadIdx.x * 1e-7f;
    a[i] = pack2(1.0f + (seed + i) * 1e-8f + tv, 1.0f - (seed + i) * 1e-8f + tv);
    b[i] = pack2(seed * 1e-3f + i * 1e-6f + tv, seed * 2e-3f + i * 1e-6f + tv);
    m[i] = 1e30f + i + tv;
    um[i] = 0x7f000000u + i + threadIdx.x;
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        b[i] = fma2(a[i], a[i], b[i]);
        float lo, hi; unpack2(b[i], lo, hi);
        if (T == T_FMIN3) m[i] = fmin3(m[i], lo, hi);
        if (T == T_FMIN2x2) { m[i] = fminf(m[i], lo); m[i] = fminf(m[i], hi); }
        if (T == T_FMIN2x1) m[i] = fminf(m[i], lo);
        if (T == T_UMIN2x2) { um[i] = min(um[i], __float_as_uint(lo)); um[i] = min(um[i], __float_as_uint(hi)); }
        if (T == T_UMIN2x1) um[i] = min(um[i], __float_as_uint(lo));
        if (T == T_UMIN3) um[i] = min(um[i], min(__float_as_uint(lo), __float_as_uint(hi)));
        if (T == T_SMIN2x2) { um[i] = (unsigned)min((int)um[i], __float_as_int(lo)); um[i] = (unsigned)min((int)um[i], __float_as_int(hi)); }
        if (T == T_FSETP_SEL) { const bool l = lo < m[i]; m[i] = l ? lo : m[i]; um[i] = l ? (unsigned)it : um[i]; }
        if (T == T_LOP) { um[i] = (um[i] & __float_as_uint(lo)) ^ __float_as_uint(hi); um[i] |= __float_as_uint(lo) >> 3; }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { float l2, h2; unpack2(b[i], l2, h2); s += l2 + h2 + m[i] + (float)um[i]; }
  if (s == 12345.f) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int T>
static void run(int sms, float* out, long long* cyc) {
  const int iters = 8000;
  for (int w = 1; w <= 4; w *= 2) {
    kern<T><<<sms * w, 128>>>(out, iters, 1.5f, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<T><<<sms * w, 128>>>(out, iters, 1.5f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double chains = (double)iters * 4 * 16;
    double cyc_per_chain = ms * 1e-3 * 1.965e9 / (chains * w);   // SMSP cycles per (FFMA2 + min ops) group
    printf("{\"test\": \"%s\", \"warps_per_smsp\": %d, \"smsp_cycles_per_group\": %.3f, \"instr_per_group\": %d, \"ms\": %.4f}\n",
           names[T], w, cyc_per_chain, instr_per_chain[T], ms);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  int n = p.multiProcessorCount;
  run<T_NONE>(n, out, cyc); run<T_FMIN3>(n, out, cyc); run<T_FMIN2x2>(n, out, cyc); run<T_FMIN2x1>(n, out, cyc);
  run<T_UMIN2x2>(n, out, cyc); run<T_UMIN2x1>(n, out, cyc); run<T_UMIN3>(n, out, cyc); run<T_SMIN2x2>(n, out, cyc);
  run<T_FSETP_SEL>(n, out, cyc); run<T_LOP>(n, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
