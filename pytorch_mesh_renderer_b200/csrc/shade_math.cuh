// shade_math.cuh -- per-pixel arithmetic of the diffuse + ambient Phong term (reference
// src/mesh_renderer/render.py:201-228 and phong_shader :231-325), shared by the standalone shading kernels
// (shade.cu) and by the render path that shades inside the rasterizer's resolve / backward kernels.
#pragma once
#include <cuda_runtime.h>

namespace pmr {

constexpr int kMaxLights = 16;
constexpr float kNormalizeEps = 1e-12f;     // torch.nn.functional.normalize default

struct Lights {
  float pos[kMaxLights][3];
  float intensity[kMaxLights][3];
  float ambient[3];
};

__device__ __forceinline__ void load_lights(Lights &sm, const float *__restrict__ light_positions,
                                            const float *__restrict__ light_intensities,
                                            const float *__restrict__ ambient, int b, int L) {
  for (int i = threadIdx.x; i < L * 3; i += blockDim.x) {
    sm.pos[i / 3][i % 3] = light_positions[(size_t)b * L * 3 + i];
    sm.intensity[i / 3][i % 3] = light_intensities[(size_t)b * L * 3 + i];
  }
  if (threadIdx.x < 3) sm.ambient[threadIdx.x] = ambient != nullptr ? ambient[(size_t)b * 3 + threadIdx.x] : 0.0f;
  __syncthreads();
}

// v / max(|v|, eps) (render.py:201 and :318-321) as one IEEE reciprocal and three multiplies (within
// 1.5 ulp of the three divisions torch performs).  Returns 1 / max(|v|, eps); `len` receives |v|.
__device__ __forceinline__ float normalize3(const float v[3], float out[3], float &len) {
  len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const float inv = 1.0f / fmaxf(len, kNormalizeEps);
  out[0] = v[0] * inv; out[1] = v[1] * inv; out[2] = v[2] * inv;
  return inv;
}

// Backward of normalize3: g = d(loss)/d(out) -> d(loss)/d(v).  torch: v / norm.clamp_min(eps); the
// clamp passes gradient to the norm when norm >= eps (then v / norm is `unit` itself), and not below.
__device__ __forceinline__ void normalize3_backward(const float unit[3], float len, float inv,
                                                    const float g[3], float dv[3]) {
  const float dot = g[0] * unit[0] + g[1] * unit[1] + g[2] * unit[2];      // = sum(g * v) / denom
  const float through_norm = len >= kNormalizeEps ? dot * inv : 0.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) dv[k] = g[k] * inv - through_norm * unit[k];
}

// RGBA of one pixel from its nine interpolated channels [normal, world position, diffuse colour].
__device__ __forceinline__ float4 shade_diffuse_pixel(const float n_raw[3], const float pos[3], const float kd[3],
                                                      const Lights &lights, int L, bool has_ambient) {
  // background pixels carry diffuse = -1 in every channel (render.py:197, :215)
  const float alpha = (kd[0] >= 0.0f || kd[1] >= 0.0f || kd[2] >= 0.0f) ? 1.0f : 0.0f;
  float n[3], len;
  normalize3(n_raw, n, len);
  float rgb[3] = {0.0f, 0.0f, 0.0f};
  for (int l = 0; l < L; ++l) {
    const float d[3] = {lights.pos[l][0] - pos[0], lights.pos[l][1] - pos[1], lights.pos[l][2] - pos[2]};
    float u[3];
    normalize3(d, u, len);
    const float ndl = fminf(fmaxf(n[0] * u[0] + n[1] * u[1] + n[2] * u[2], 0.0f), 1.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] += kd[c] * ndl * lights.intensity[l][c];
  }
  if (has_ambient) {
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = lights.ambient[c] * kd[c] + rgb[c];
  }
  const bool valid = alpha > 0.5f;
  return make_float4(valid ? rgb[0] : 0.0f, valid ? rgb[1] : 0.0f, valid ? rgb[2] : 0.0f, alpha);
}

// d(rgb) of a valid pixel -> d(normal), d(position), d(diffuse colour)  (zero for masked pixels).
__device__ __forceinline__ void shade_diffuse_pixel_backward(const float n_raw[3], const float pos[3], const float kd[3],
                                                             const float g[3], const Lights &lights, int L,
                                                             bool has_ambient, float d_n[3], float d_pos[3], float d_kd[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) d_n[k] = d_pos[k] = d_kd[k] = 0.0f;
  if (!(kd[0] >= 0.0f || kd[1] >= 0.0f || kd[2] >= 0.0f)) return;
  float n[3], n_len;
  const float n_inv = normalize3(n_raw, n, n_len);
  float d_unit_n[3] = {0.0f, 0.0f, 0.0f};
  for (int l = 0; l < L; ++l) {
    const float d[3] = {lights.pos[l][0] - pos[0], lights.pos[l][1] - pos[1], lights.pos[l][2] - pos[2]};
    float u[3], d_len;
    const float d_inv = normalize3(d, u, d_len);
    const float s = n[0] * u[0] + n[1] * u[1] + n[2] * u[2];
    const float ndl = fminf(fmaxf(s, 0.0f), 1.0f);
    float d_ndl = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d_kd[c] += g[c] * (ndl * lights.intensity[l][c]);
      d_ndl += g[c] * (kd[c] * lights.intensity[l][c]);
    }
    const float d_s = (s >= 0.0f && s <= 1.0f) ? d_ndl : 0.0f;        // torch.clamp passes on the closed interval
    const float d_u[3] = {d_s * n[0], d_s * n[1], d_s * n[2]};
    float d_d[3];
    normalize3_backward(u, d_len, d_inv, d_u, d_d);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d_unit_n[k] += d_s * u[k];
      d_pos[k] -= d_d[k];                                              // d = light - position
    }
  }
  if (has_ambient) {
#pragma unroll
    for (int c = 0; c < 3; ++c) d_kd[c] += g[c] * lights.ambient[c];
  }
  normalize3_backward(n, n_len, n_inv, d_unit_n, d_n);
}

}  // namespace pmr
