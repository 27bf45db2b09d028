// micro-benchmark: MUFU.EX2, FFMA and TMEM-free ALU issue rates per SM (cycles per warp-instruction per SMSP)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ex2(float* out, int iters, int mode) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (mode == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (mode == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] = fmaf(a[i], 1.0001f, 0.5f); a[i] = a[i] + 1.0f; a[i] = fmaf(a[i], 0.999f, 0.25f); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("mode %d warps/SM %d: %.2f cycles per warp-level op group (16 independent chains)\n", mode, blockDim.x / 32, (double)(t1 - t0) / (iters * 16.0));
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512}) { k_ex2<<<148, threads>>>(d, 2000, mode); cudaDeviceSynchronize(); }
  return 0;
}
