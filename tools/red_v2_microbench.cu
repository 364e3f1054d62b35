// Does one red.global.add.v2.f32 cost less than two scalar red.global.add.f32?  Address pattern of the point
// rasteriser: every point adds into a 9x9 pixel patch of a 256x256 image (one thread per (point, patch pixel)).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/red_v2_microbench tools/red_v2_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

template <int MODE>
__global__ void k(const int* __restrict__ px, const int* __restrict__ py, float* __restrict__ a, float* __restrict__ b,
                  long long total, int N) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pt = t / 81;
    const int kk = (int)(t - pt * 81), j = kk / 9, i = kk - j * 9;
    const int x = px[pt] - 4 + i, y = py[pt] - 4 + j;
    if (x < 0 || x >= 256 || y < 0 || y >= 256) continue;
    const long long o = (pt / N) * 65536 + y * 256 + x;
    const float g = 1.0f + i * 0.01f;
    if (MODE == 0) { atomicAdd(a + o, g); atomicAdd(b + o, g * 0.5f); }
    if (MODE == 1) { red_v2(a + 2 * o, g, g * 0.5f); }
    if (MODE == 2) { atomicAdd(a + o, g); }
  }
}

// the same update with 32-bit index arithmetic and a compile-time patch width (the 64-bit `t / 81` of the loop above is
// an emulated division of ~100 instructions)
template <int MODE>
__global__ void k32(const int* __restrict__ px, const int* __restrict__ py, float* __restrict__ a, float* __restrict__ b,
                    unsigned total, int N) {
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned pt = t / 81u;
    const int kk = (int)(t - pt * 81u), j = kk / 9, i = kk - j * 9;
    const int x = px[pt] - 4 + i, y = py[pt] - 4 + j;
    if (x < 0 || x >= 256 || y < 0 || y >= 256) continue;
    const unsigned o = (pt / (unsigned)N) * 65536u + y * 256 + x;
    const float g = 1.0f + i * 0.01f;
    if (MODE == 0) { atomicAdd(a + o, g); atomicAdd(b + o, g * 0.5f); }
    if (MODE == 1) { red_v2(a + 2 * (size_t)o, g, g * 0.5f); }
  }
}

int main() {
  const int B = 200, N = 8192;
  const long long npts = (long long)B * N, total = npts * 81;
  int *hx = new int[npts], *hy = new int[npts];
  unsigned s = 12345;
  for (long long i = 0; i < npts; ++i) { s = s * 1664525u + 1013904223u; hx[i] = 64 + (s >> 16) % 128; s = s * 1664525u + 1013904223u; hy[i] = 64 + (s >> 16) % 128; }
  int *px, *py; float *a, *b;
  cudaMalloc(&px, npts * 4); cudaMalloc(&py, npts * 4);
  cudaMalloc(&a, (size_t)B * 65536 * 8); cudaMalloc(&b, (size_t)B * 65536 * 4);
  cudaMemcpy(px, hx, npts * 4, cudaMemcpyHostToDevice); cudaMemcpy(py, hy, npts * 4, cudaMemcpyHostToDevice);
  cudaMemset(a, 0, (size_t)B * 65536 * 8); cudaMemset(b, 0, (size_t)B * 65536 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 32;
  for (int mode = 0; mode < 5; ++mode) {
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, 256>>>(px, py, a, b, total, N);
      if (mode == 1) k<1><<<grid, 256>>>(px, py, a, b, total, N);
      if (mode == 2) k<2><<<grid, 256>>>(px, py, a, b, total, N);
      if (mode == 3) k32<0><<<grid, 256>>>(px, py, a, b, (unsigned)total, N);
      if (mode == 4) k32<1><<<grid, 256>>>(px, py, a, b, (unsigned)total, N);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"G_lane_updates_per_s\": %.1f}\n",
           mode == 0 ? "2 x red.f32 (separate images)" : mode == 1 ? "1 x red.v2.f32 (interleaved)" : mode == 2 ? "1 x red.f32"
           : mode == 3 ? "2 x red.f32, 32-bit indices" : "1 x red.v2.f32, 32-bit indices",
           best, total / best / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
