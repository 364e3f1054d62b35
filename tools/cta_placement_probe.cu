// Which SM does block k of a single-wave grid land on?  (2 resident CTAs per SM, 296 blocks.)
// Prints for the first / second half of the block indices how many distinct SMs they cover, and the first 24 assignments.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/cta_placement_probe tools/cta_placement_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <set>
#include <vector>
__global__ void __launch_bounds__(128, 2) probe(unsigned* smid, unsigned long long* t0, int spin) {
  extern __shared__ float sm[];
  unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (threadIdx.x == 0) { smid[blockIdx.x] = s; t0[blockIdx.x] = t; }
  long long c0 = clock64();
  while (clock64() - c0 < spin) { }
  if (spin < 0) sm[0] = 1.f;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 100 * 1024;   // two CTAs per SM
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n : {2 * sms, 2 * sms - 3, sms + 40}) {
    unsigned* d; unsigned long long* dt;
    cudaMalloc(&d, n * 4); cudaMalloc(&dt, n * 8);
    probe<<<n, 128, smem>>>(d, dt, 200000);
    cudaDeviceSynchronize();
    std::vector<unsigned> h(n); cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
    std::set<unsigned> a(h.begin(), h.begin() + (n < sms ? n : sms)), b;
    if (n > sms) b = std::set<unsigned>(h.begin() + sms, h.end());
    int same_pair = 0;
    for (int k = 0; k + 1 < n; k += 2) same_pair += h[k] == h[k + 1];
    printf("blocks %d: first %d blocks cover %zu SMs, the rest %zu SMs; consecutive pairs (2k,2k+1) on the same SM: %d\n  first 24:", n, sms, a.size(), b.size(), same_pair);
    for (int k = 0; k < 24; ++k) printf(" %u", h[k]);
    printf("\n  blocks %d..: ", sms);
    for (int k = sms; k < sms + 12 && k < n; ++k) printf(" %u", h[k]);
    printf("\n");
    cudaFree(d); cudaFree(dt);
  }
  return 0;
}
