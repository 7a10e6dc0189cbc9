// Build on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o f32x2_issue f32x2_issue.cu
// Micro-benchmark: issue cost of packed fp32 (mul/add .f32x2, sm_100+) against scalar FMUL/FADD.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float c, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 1.0f + 0.001f * (threadIdx.x + i);
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c));
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c));
    }
  } else {
    unsigned long long p[4], cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(p[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  float h0[4], h1[4];
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, 1.0000001f, iters); else k<1><<<148 * 8, 256>>>(out, 1.0000001f, iters);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      // flops: 16 per thread per iteration
      printf("mode %d: %.3f ms, %.2f Tflop/s (mul+add counted once each)\n", mode, ms, 16.0 * iters * 148 * 8 * 256 / ms * 1e-9);
    }
    cudaMemcpy(mode ? h1 : h0, out, 16, cudaMemcpyDeviceToHost);
  }
  printf("bits equal: %d %d %d %d\n", h0[0] == h1[0], h0[1] == h1[1], h0[2] == h1[2], h0[3] == h1[3]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
