// FP32 issue-rate microbenchmark for B200 (sm_100a).
//
// Answers the questions the chamfer search kernel design depends on:
//   * scalar FFMA rate per SM per clock (the "FP32 FMA peak" the roofline is quoted against),
//   * packed FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...) rate: same FLOPs with half the issue slots?
//   * whether FMNMX3 / CREDUX issued between packed ops hide in the freed issue slots.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_microbench fp32_microbench.cu
// Prints one JSON object per variant on stdout.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
__device__ __forceinline__ float credux_min(float v) {
  float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r;
}

enum { V_FFMA = 0, V_FFMA2 = 1, V_FADD2 = 2, V_FMNMX3 = 3, V_MIX_SCALAR = 4, V_MIX_PACKED = 5,
       V_MIX_PACKED_REDUX = 6, V_FFMA_FMNMX = 7, V_COUNT = 8 };

// Each variant executes `iters` iterations of a fixed body; OPS = FP32-pipe lane-ops per thread per
// iteration (an FFMA2 counts 2), ISSUE = warp instructions per iteration (approx., body only).
template <int V>
__global__ void __launch_bounds__(128) bench(float* out, int iters, float a0, float b0, long long* cyc) {
  float a = a0, b = b0;
  long long t0 = clock64();
  if (V == V_FFMA) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = a + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.f) out[threadIdx.x] = s;
  } else if (V == V_FFMA2 || V == V_FADD2) {
    u64 acc[16];
    u64 pa = pack2(a, a), pb = pack2(b, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = pack2(a + i, b + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = (V == V_FFMA2) ? fma2(acc[i], pa, pb) : add2(acc[i], pb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float lo, hi; unpack2(acc[i], lo, hi); s += lo + hi; }
    if (s == 12345.f) out[threadIdx.x] = s;
  } else if (V == V_FMNMX3) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = a + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = min3(acc[i], acc[(i + 1) & 15], b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.f) out[threadIdx.x] = s;
  } else if (V == V_FFMA_FMNMX) {
    // 3 FFMA : 1 FMNMX (the expanded-form per-direction inner loop)
    float x[8], m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a + i; m[i] = 1e30f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float d = __fmaf_rn(x[i], a, b);
          d = __fmaf_rn(x[i], b, d);
          d = __fmaf_rn(x[i], d, a);
          m[i] = fminf(m[i], d);
        }
      a += 1e-7f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += m[i];
    if (s == 12345.f) out[threadIdx.x] = s;
  } else if (V == V_MIX_SCALAR) {
    // scalar difference form, 16 rows x 1 column per body: 3 FADD + FMUL + 2 FFMA + 1 FMNMX per pair
    float x[16], y[16], z[16], m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = a + i; y[i] = b + i; z[i] = a - i; m[i] = 1e30f; }
    for (int it = 0; it < iters; ++it) {
      float cx = a, cy = b, cz = a + b;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float dx = x[i] + cx, dy = y[i] + cy, dz = z[i] + cz;
        float d = __fmul_rn(dx, dx);
        d = __fmaf_rn(dy, dy, d);
        d = __fmaf_rn(dz, dz, d);
        m[i] = fminf(m[i], d);
      }
      a += 1e-7f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += m[i];
    if (s == 12345.f) out[threadIdx.x] = s;
  } else if (V == V_MIX_PACKED || V == V_MIX_PACKED_REDUX) {
    // packed difference form, 16 rows x 2 columns per body:
    //   per column: 8 x (3 FADD2 + FMUL2 + 2 FFMA2) ; 8 FMNMX3 column tree ; (+ CREDUX)
    //   per 2 columns: 16 FMNMX3 row mins
    u64 px[8], py[8], pz[8];
    float m[16];
    float cacc = 1e30f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      px[i] = pack2(a + i, a - i); py[i] = pack2(b + i, b - i); pz[i] = pack2(a * i, b * i);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = 1e30f;
    for (int it = 0; it < iters; ++it) {
      float d[2][16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        u64 cx = pack2(a + c, a + c), cy = pack2(b + c, b + c), cz = pack2(a - c, a - c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          u64 dx = add2(px[i], cx), dy = add2(py[i], cy), dz = add2(pz[i], cz);
          u64 t = mul2(dx, dx);
          t = fma2(dy, dy, t);
          t = fma2(dz, dz, t);
          unpack2(t, d[c][2 * i], d[c][2 * i + 1]);
        }
        float c0 = min3(d[c][0], d[c][1], d[c][2]);
        float c1 = min3(d[c][3], d[c][4], d[c][5]);
        float c2 = min3(d[c][6], d[c][7], d[c][8]);
        float c3 = min3(d[c][9], d[c][10], d[c][11]);
        float c4 = min3(d[c][12], d[c][13], d[c][14]);
        float c5 = min3(c0, c1, d[c][15]);
        float c6 = min3(c2, c3, c4);
        float cm = fminf(c5, c6);
        if (V == V_MIX_PACKED_REDUX) {
          float v = credux_min(cm);
          unsigned mask = __ballot_sync(0xffffffffu, cm == v);
          if ((threadIdx.x & 31) == 0 && v < -1.f) out[it & 1023] = v + (float)mask;
        } else {
          cacc = fminf(cacc, cm);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) m[i] = min3(m[i], d[0][i], d[1][i]);
      a += 1e-7f;
    }
    float s = cacc;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += m[i];
    if (s == 12345.f) out[threadIdx.x] = s;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

struct Desc { const char* name; double fp32_ops; double issue; };
static const Desc descs[V_COUNT] = {
  {"ffma_scalar", 64, 64},       {"ffma2_packed", 128, 64},     {"fadd2_packed", 128, 64},
  {"fmnmx3", 0, 64},             {"diff_scalar_16x1", 96, 112}, {"diff_packed_16x2", 192, 128},
  {"diff_packed_16x2_redux", 192, 136}, {"ffma3_fmnmx1", 96, 128}};

template <int V>
static void run(int sms, int ctas_per_sm, int iters, float* out, long long* cyc, int clock_khz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = sms * ctas_per_sm;
  bench<V><<<grid, 128>>>(out, iters, 1.0f, 0.5f, cyc);  // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    bench<V><<<grid, 128>>>(out, iters, 1.0f, 0.5f, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  double threads = (double)grid * 128;
  double ops = descs[V].fp32_ops * iters * threads;
  double wi = descs[V].issue * iters * threads / 32.0;
  double sec = best * 1e-3;
  // per-SM per-clock figures use the on-device cycle count of CTA 0 (same work on every CTA)
  double ops_clk_sm = descs[V].fp32_ops * iters * 128.0 * ctas_per_sm / (double)c;
  double issue_clk_smsp = descs[V].issue * iters * ctas_per_sm / (double)c;
  printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"cycles_cta0\": %lld, "
         "\"fp32_lane_ops_per_s\": %.4e, \"warp_instr_per_s\": %.4e, \"fp32_ops_per_clk_per_sm\": %.2f, "
         "\"issue_per_clk_per_smsp\": %.3f, \"eff_mhz\": %.1f}\n",
         descs[V].name, ctas_per_sm, best, c, ops / sec, wi / sec, ops_clk_sm, issue_clk_smsp,
         (double)c / sec * 1e-6);
  (void)clock_khz;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, khz);
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long));
  const int iters = 20000;
  for (int cps = 1; cps <= 4; cps *= 2) {
    run<V_FFMA>(sms, cps, iters, out, cyc, khz);
    run<V_FFMA2>(sms, cps, iters, out, cyc, khz);
    run<V_FADD2>(sms, cps, iters, out, cyc, khz);
    run<V_FMNMX3>(sms, cps, iters, out, cyc, khz);
    run<V_FFMA_FMNMX>(sms, cps, iters, out, cyc, khz);
    run<V_MIX_SCALAR>(sms, cps, iters, out, cyc, khz);
    run<V_MIX_PACKED>(sms, cps, iters, out, cyc, khz);
    run<V_MIX_PACKED_REDUX>(sms, cps, iters, out, cyc, khz);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
