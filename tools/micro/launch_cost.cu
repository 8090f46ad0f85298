// Micro-benchmark: how long do grids of short-lived CTAs take on this GPU?  (calibrates the latency model used in DESIGN.md)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_empty() {}
__global__ void k_chain(const int *p, int *out, int hops) {
  int i = blockIdx.x * 131 + threadIdx.x;
  for (int h = 0; h < hops; ++h) i = p[i & 0xfffff];
  if (i == -12345) out[0] = i;
}
__global__ void k_smem(int *out, int bytes) {
  extern __shared__ int s[];
  s[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (s[(threadIdx.x + 1) & 127] == -1) out[0] = 1;
}
template <typename F>
float timeit(F f, int it = 50) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < it; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return 1e3f * ms / it;
}
int main() {
  int *p, *out;
  cudaMalloc(&p, 4 << 20); cudaMalloc(&out, 4);
  cudaMemset(p, 0, 4 << 20);
  for (int ctas : {148, 1312, 2624, 8192, 16384}) {
    for (int thr : {128, 256}) {
      float e = timeit([&] { k_empty<<<ctas, thr>>>(); });
      float c1 = timeit([&] { k_chain<<<ctas, thr>>>(p, out, 1); });
      float c4 = timeit([&] { k_chain<<<ctas, thr>>>(p, out, 4); });
      float c16 = timeit([&] { k_chain<<<ctas, thr>>>(p, out, 16); });
      cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 44 * 1024);
      float s = timeit([&] { k_smem<<<ctas, thr, 44 * 1024>>>(out, 0); });
      printf("ctas %6d thr %3d: empty %6.2f us  chain1 %6.2f  chain4 %6.2f  chain16 %6.2f  smem44K %6.2f\n", ctas, thr, e, c1, c4, c16, s);
    }
  }
  return 0;
}
