// vertex_stage.cu -- the step immediately in front of the hot path: world -> clip space
// (reference src/common/camera_utils.py:142-170 transform_homogeneous, (M V^T)^T with w = 1) and its
// backward.  The reference does it with torch.cat + matmul; for multi-view fitting, where one mesh
// [V,3] is seen by B views, that materialises [B,V,4] homogeneous copies and reduces a [B,V,3]
// gradient afterwards.  Here: one kernel forward, one backward that sums over the views on the fly and
// writes the shared [V,3] gradient that the single all-reduce then moves over NVLink.
#include "pmr_internal.cuh"

namespace pmr {

// clip[b][v] = M_b * (x, y, z, 1).  world is [V,3] (shared, batch stride 0) or [B,V,3].
__global__ void __launch_bounds__(256)
transform_forward_kernel(const float *__restrict__ matrices, const float *__restrict__ world, int V,
                         long long world_batch_stride, float4 *__restrict__ clip) {
  __shared__ float m[16];
  const int b = blockIdx.y;
  if (threadIdx.x < 16) m[threadIdx.x] = matrices[(size_t)b * 16 + threadIdx.x];
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float *p = world + (size_t)b * world_batch_stride + (size_t)v * 3;
  const float x = p[0], y = p[1], z = p[2];
  float4 o;
  o.x = m[0] * x + m[1] * y + m[2] * z + m[3];
  o.y = m[4] * x + m[5] * y + m[6] * z + m[7];
  o.z = m[8] * x + m[9] * y + m[10] * z + m[11];
  o.w = m[12] * x + m[13] * y + m[14] * z + m[15];
  clip[(size_t)b * V + v] = o;
}

// d_world[v][k] = sum_b sum_c M_b[c][k] * d_clip[b][v][c]  (shared) or per b (not shared).
__global__ void __launch_bounds__(256)
transform_backward_kernel(const float *__restrict__ matrices, const float4 *__restrict__ d_clip, int B, int V,
                          int shared, float *__restrict__ d_world) {
  extern __shared__ float ms[];     // [B or 1][12]: columns 0..2 of every row of M_b
  const int b0 = shared ? 0 : blockIdx.y;
  const int nb = shared ? B : 1;
  for (int i = threadIdx.x; i < nb * 12; i += blockDim.x) {
    const int bb = b0 + i / 12, r = (i % 12) / 3, k = i % 3;
    ms[i] = matrices[(size_t)bb * 16 + r * 4 + k];
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float gx = 0.0f, gy = 0.0f, gz = 0.0f;
  for (int i = 0; i < nb; ++i) {
    const float4 g = __ldg(d_clip + (size_t)(b0 + i) * V + v);
    const float *m = ms + i * 12;
    gx += m[0] * g.x + m[3] * g.y + m[6] * g.z + m[9] * g.w;
    gy += m[1] * g.x + m[4] * g.y + m[7] * g.z + m[10] * g.w;
    gz += m[2] * g.x + m[5] * g.y + m[8] * g.z + m[11] * g.w;
  }
  float *o = d_world + ((size_t)(shared ? 0 : b0) * V + v) * 3;
  o[0] = gx; o[1] = gy; o[2] = gz;
}

int transform_forward_impl(Context *ctx, const float *matrices, const float *world, int B, int V, int shared,
                           float *clip, cudaStream_t stream) {
  if (B == 0 || V == 0) return PMR_OK;
  transform_forward_kernel<<<dim3((V + 255) / 256, B), 256, 0, stream>>>(matrices, world, V,
                                                                        shared ? 0ll : (long long)V * 3,
                                                                        reinterpret_cast<float4 *>(clip));
  ctx->launches += 1;
  return check_launch(ctx, "transform_forward_kernel");
}

int transform_backward_impl(Context *ctx, const float *matrices, const float *d_clip, int B, int V, int shared,
                            float *d_world, cudaStream_t stream) {
  if (B == 0 || V == 0) return PMR_OK;
  const size_t smem = (size_t)(shared ? B : 1) * 12 * sizeof(float);
  if (smem > 48 * 1024) return set_error(ctx, PMR_ERR_SIZE, "too many views for one transform_backward launch");
  transform_backward_kernel<<<dim3((V + 255) / 256, shared ? 1 : B), 256, smem, stream>>>(
      matrices, reinterpret_cast<const float4 *>(d_clip), B, V, shared, d_world);
  ctx->launches += 1;
  return check_launch(ctx, "transform_backward_kernel");
}

}  // namespace pmr
