// The exact arithmetic mix of one column pair of chamfer_fsearch_kernel (16 rows x 2 columns = 32 point pairs):
// 48 FFMA2 (pair x scalar + pair / scalar), 16 FADD2, 16 FMNMX3 (row minima), 16 FMNMX3 (column minima), on registers
// only - no shared memory, no records - with ablations.  Cycles per point pair per lane tell the floor of the mix
// itself, against which the kernel's measured 8.0 cycles/pair is to be read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/filter_mix_microbench tools/filter_mix_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

enum { V_FULL, V_NOROWMIN, V_NOCOLMIN, V_NOADD, V_FMAONLY, V_MIN2, V_DUP_FMAONLY, V_DUP_FULL, V_COUNT };
static const char* names[V_COUNT] = {"full mix", "no row minima", "no column minima (and no FADD2 use)", "no FADD2 (column minima on f)",
                                     "FFMA2 chain only", "full mix with 2-input FMNMX instead of FMNMX3",
                                     "row pairs x duplicated column pairs from smem (no .F32 broadcast): FFMA2 chain only",
                                     "row pairs x duplicated column pairs from smem: full mix"};

template <int V>
__global__ void __launch_bounds__(128) kern(float* out, int iters, float seed, long long* cyc) {
  float x[16], y[16], z[16], u[16], tm[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float tv = threadIdx.x * 1e-4f + r * 1e-2f;
    x[r] = seed + tv; y[r] = seed * 0.5f - tv; z[r] = seed * 0.25f + 2 * tv; u[r] = x[r] * x[r] + y[r] * y[r] + z[r] * z[r];
    tm[r] = 1e30f;
  }
  float cl = 1e30f, ch = 1e30f;
  float q0a = seed * -2.f, q0b = seed * -2.1f, q1a = seed * 1.3f, q1b = seed * 1.4f, q2a = seed * 0.7f, q2b = seed * 0.6f, wa = 3.f * seed, wb = 2.9f * seed;
  u64 keep = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // a new column pair every iteration (as from an LDS.128 pair in the kernel)
    q0a += 1e-3f; q0b -= 1e-3f; q1a += 2e-3f; q2b -= 1e-3f; wa += 1e-3f; wb += 2e-3f;
    const u64 Q0 = pack2(q0a, q0b), Q1 = pack2(q1a, q1b), Q2 = pack2(q2a, q2b), WW = pack2(wa, wb);
    u64 acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = fma2(Q2, pack2(z[r], z[r]), WW);
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = fma2(Q1, pack2(y[r], y[r]), acc[r]);
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = fma2(Q0, pack2(x[r], x[r]), acc[r]);
#pragma unroll
    for (int r = 0; r < 16; r += 2) {
      float al, ah, bl, bh;
      unpack2(acc[r], al, ah); unpack2(acc[r + 1], bl, bh);
      if (V == V_FULL || V == V_NOCOLMIN || V == V_NOADD) { tm[r] = min3(tm[r], al, ah); tm[r + 1] = min3(tm[r + 1], bl, bh); }
      if (V == V_MIN2) { tm[r] = fminf(fminf(tm[r], al), ah); tm[r + 1] = fminf(fminf(tm[r + 1], bl), bh); }
      if (V == V_FULL || V == V_NOROWMIN || V == V_MIN2) {
        unpack2(add2(acc[r], pack2(u[r], u[r])), al, ah); unpack2(add2(acc[r + 1], pack2(u[r + 1], u[r + 1])), bl, bh);
      }
      if (V == V_FULL || V == V_NOROWMIN || V == V_NOADD) { cl = min3(cl, al, bl); ch = min3(ch, ah, bh); }
      if (V == V_MIN2) { cl = fminf(fminf(cl, al), bl); ch = fminf(fminf(ch, ah), bh); }
      if (V == V_FMAONLY || V == V_NOROWMIN || V == V_NOCOLMIN) keep ^= acc[r] ^ acc[r + 1];   // 2 LOP3 per row pair keep the chain alive
    }
  }
  long long t1 = clock64();
  float s = cl + ch + (float)(keep & 0xff);
#pragma unroll
  for (int r = 0; r < 16; ++r) s += tm[r];
  if (s == 12345.f) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// Row pairs (x_r0, x_r1) in registers, column operands duplicated (q, q) and loaded as 64-bit pairs from shared memory:
// every FFMA2 operand is a true F32x2 pair.  One iteration = 2 columns x 16 rows = 32 point pairs, as above.
template <int V>
__global__ void __launch_bounds__(128) kern_dup(float* out, int iters, float seed, long long* cyc) {
  __shared__ u64 scol[64 * 4];
  for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) { const float v = seed * (1.f + 0.01f * i); scol[i] = pack2(v, v); }
  __syncthreads();
  u64 px[8], py[8], pz[8], pu[8];
  float tm[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float tv = threadIdx.x * 1e-4f + k * 1e-2f;
    px[k] = pack2(seed + tv, seed - tv); py[k] = pack2(seed * .5f - tv, seed * .5f + tv); pz[k] = pack2(seed * .25f + tv, seed * .3f - tv);
    pu[k] = pack2(3 * seed + tv, 3 * seed - tv); tm[2 * k] = tm[2 * k + 1] = 1e30f;
  }
  float cm[2] = {1e30f, 1e30f};
  u64 keep = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const volatile u64* sp = scol + ((it * 2 + c) & 63) * 4;
      const u64 Q0 = sp[0], Q1 = sp[1], Q2 = sp[2], WW = sp[3];
      u64 acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fma2(pz[k], Q2, WW);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fma2(py[k], Q1, acc[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fma2(px[k], Q0, acc[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (V == V_DUP_FULL) {
          float lo, hi;
          unpack2(acc[k], lo, hi);
          tm[2 * k] = fminf(tm[2 * k], lo); tm[2 * k + 1] = fminf(tm[2 * k + 1], hi);
          unpack2(add2(acc[k], pu[k]), lo, hi);
          cm[c] = fminf(fminf(cm[c], lo), hi);
        } else {
          keep ^= acc[k];
        }
      }
    }
  }
  long long t1 = clock64();
  float s = cm[0] + cm[1] + (float)(keep & 0xff);
#pragma unroll
  for (int r = 0; r < 16; ++r) s += tm[r];
  if (s == 12345.f) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int V>
static void run(int sms, float* out, long long* cyc) {
  const int iters = 20000;
  for (int wps = 1; wps <= 2; ++wps) {
    if (V >= V_DUP_FMAONLY) kern_dup<V><<<sms * wps, 128>>>(out, 100, 1.0f, cyc); else kern<V><<<sms * wps, 128>>>(out, 100, 1.0f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (V >= V_DUP_FMAONLY) kern_dup<V><<<sms * wps, 128>>>(out, iters, 1.0f, cyc); else kern<V><<<sms * wps, 128>>>(out, iters, 1.0f, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // each SM sub-partition runs `wps` warps; cycles per point pair per lane = kernel cycles / (iters * 32 pairs * wps)
    const double cyc_total = ms * 1e-3 * 1.965e9;
    printf("{\"variant\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_pair\": %.3f, \"ms\": %.4f}\n", names[V], wps,
           cyc_total / ((double)iters * 32 * wps), ms);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const int n = p.multiProcessorCount;
  run<V_FULL>(n, out, cyc); run<V_NOROWMIN>(n, out, cyc); run<V_NOCOLMIN>(n, out, cyc); run<V_NOADD>(n, out, cyc);
  run<V_FMAONLY>(n, out, cyc); run<V_MIN2>(n, out, cyc); run<V_DUP_FMAONLY>(n, out, cyc); run<V_DUP_FULL>(n, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
