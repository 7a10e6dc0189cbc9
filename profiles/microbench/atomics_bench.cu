// Microbenchmark: throughput of the reduction primitives the backward pass can be built from
// (global RED scalar / v2 / v4 with spread and clustered addresses, shared-memory int atomics,
// warp shuffles).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench atomics_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int VEC>
__global__ void red_kernel(float* out, uint32_t n_rows, int iters, int cluster) {
  // each lane adds VEC floats to row r; `cluster` consecutive lanes share a row (contention)
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    uint32_t r = hash((tid / cluster) * 9781u + it * 7919u) % n_rows;
    float* p = out + (size_t)r * 4;
    if (VEC == 1) atomicAdd(p + (tid & 3), 1.0f);
    else if (VEC == 2) atomicAdd(reinterpret_cast<float2*>(p) + (tid & 1), make_float2(1.f, 1.f));
    else atomicAdd(reinterpret_cast<float4*>(p), make_float4(1.f, 1.f, 1.f, 1.f));
  }
}

__global__ void smem_int_atomic_kernel(int* out, int iters, int spread) {
  __shared__ int acc[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) acc[i] = 0;
  __syncthreads();
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    uint32_t r = spread ? hash(tid * 31u + it) & 2047u : (hash(tid / 8 + it) & 2047u);
    atomicAdd(&acc[r], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = acc[0];
}

__global__ void smem_cas_float_kernel(float* out, int iters, int spread) {
  __shared__ float acc[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) acc[i] = 0;
  __syncthreads();
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    uint32_t r = spread ? hash(tid * 31u + it) & 2047u : (hash(tid / 8 + it) & 2047u);
    atomicAdd(&acc[r], 1.0f);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = acc[0];
}

__global__ void shfl_kernel(float* out, int iters) {
  float v = threadIdx.x, a = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { a += __shfl_xor_sync(0xffffffffu, v, d); v += 1.0f; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  const uint32_t n_rows = 1u << 21;   // 32 MB of float4 rows (L2 resident)
  float* buf; cudaMalloc(&buf, (size_t)n_rows * 16); cudaMemset(buf, 0, (size_t)n_rows * 16);
  int* ibuf; cudaMalloc(&ibuf, 1 << 20);
  const int blocks = 148 * 16, threads = 256, iters = 64;
  const double ops = (double)blocks * threads * iters;
  for (int cluster : {1, 8, 32}) {
    float t1 = time_ms([&] { red_kernel<1><<<blocks, threads>>>(buf, n_rows, iters, cluster); });
    float t2 = time_ms([&] { red_kernel<2><<<blocks, threads>>>(buf, n_rows, iters, cluster); });
    float t4 = time_ms([&] { red_kernel<4><<<blocks, threads>>>(buf, n_rows, iters, cluster); });
    printf("REDG lanes-per-row=%2d: scalar %.1f Gop/s  v2 %.1f Gop/s  v4 %.1f Gop/s (ops = lane-atomics)\n", cluster,
           ops / t1 / 1e6, ops / t2 / 1e6, ops / t4 / 1e6);
  }
  for (int spread : {1, 0}) {
    float ti = time_ms([&] { smem_int_atomic_kernel<<<blocks, threads>>>(ibuf, iters * 4, spread); });
    float tf = time_ms([&] { smem_cas_float_kernel<<<blocks, threads>>>((float*)ibuf, iters * 4, spread); });
    printf("ATOMS %s: int add %.1f Gop/s   float add (CAS loop) %.1f Gop/s\n", spread ? "spread" : "8-lane clusters",
           ops * 4 / ti / 1e6, ops * 4 / tf / 1e6);
  }
  float ts = time_ms([&] { shfl_kernel<<<blocks, threads>>>(buf, 256); });
  printf("SHFL.BFLY: %.1f G lane-shuffles/s  (%.2f warp-shuffles/clk/SM at 1.9 GHz)\n",
         (double)blocks * threads * 256 * 5 / ts / 1e6, (double)blocks * threads * 256 * 5 / 32 / (ts * 1e-3) / 148 / 1.9e9);
  return 0;
}
