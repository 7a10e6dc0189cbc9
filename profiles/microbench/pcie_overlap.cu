// Does an H2D copy on one stream overlap a D2H copy on another (pinned buffers)?
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
int main() {
  const size_t n = 600u << 20;
  char *h_a, *h_b, *d_a, *d_b;
  cudaMallocHost(&h_a, n); cudaMallocHost(&h_b, n); cudaMalloc(&d_a, n); cudaMalloc(&d_b, n);
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int rep = 0; rep < 2; ++rep) {
    double t0 = now(); cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s1); cudaStreamSynchronize(s1); double t1 = now();
    cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2); double t2 = now();
    cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s2);
    cudaStreamSynchronize(s1); cudaStreamSynchronize(s2); double t3 = now();
    cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, 0); cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s2);
    cudaStreamSynchronize(0); cudaStreamSynchronize(s2); double t4 = now();
    printf("H2D %.1f GB/s  D2H %.1f GB/s  both(2 streams) %.2f ms vs sum %.2f ms  both(default+nonblocking) %.2f ms\n", n / (t1 - t0) / 1e9, n / (t2 - t1) / 1e9,
           (t3 - t2) * 1e3, (t2 - t0) * 1e3, (t4 - t3) * 1e3);
  }
  return 0;
}
