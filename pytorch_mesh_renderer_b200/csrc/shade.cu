// shade.cu -- the direct caller of the hot path: per-pixel Phong lighting of the interpolated
// attribute image, diffuse + ambient terms (reference src/mesh_renderer/render.py:201-228 and
// phong_shader :231-386 without the specular branch, i.e. `render` called without specular colors --
// the configuration of mesh_renderer_test.py:30-70 and BASELINE config c1).
//
// The reference runs ~20 torch ops that materialise [B,L,P,3] temporaries several times the size of
// the rasterizer's own output; here one kernel reads the 9 interpolated channels of a pixel
// ([normal, world position, diffuse colour], render.py:181) and writes RGBA, and one kernel maps
// d(RGBA) back to d(pixel channels).  Gradients with respect to the lights / ambient colour are not
// produced here (the Python layer keeps the torch-op path for callers that ask for them), and the
// specular branch, whose per-(image, light) L2 normalisation over all pixels (:347-353) needs extra
// passes, stays on the torch-op path as well.
//
// Arithmetic follows the torch ops of the reference term by term (F.normalize with eps = 1e-12,
// clamp to [0, 1] with gradient passed on the closed interval, lights summed in order), but libraries
// differ (torch CPU there), so parity is within the float tolerance stated in tests/test_gpu_shade.py,
// not bit-exact.
#include "pmr_internal.cuh"

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

// One thread per pixel.  pixels [B,H,W,A] (channels 0..8 used), rgba [B,H,W,4] with rows flipped
// (row 0 of the result is the top of the image, render.py:382-386).
__global__ void __launch_bounds__(256)
shade_diffuse_forward_kernel(const float *__restrict__ pixels, const float *__restrict__ light_positions,
                             const float *__restrict__ light_intensities, const float *__restrict__ ambient,
                             int L, int A, int W, int H, float4 *__restrict__ rgba) {
  __shared__ Lights sm;
  const int b = blockIdx.y;
  load_lights(sm, light_positions, light_intensities, ambient, b, L);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= W * H) return;
  const int y = p / W, x = p - y * W;
  const float *px = pixels + ((size_t)b * H * W + p) * A;
  const float n_raw[3] = {px[0], px[1], px[2]}, pos[3] = {px[3], px[4], px[5]}, kd[3] = {px[6], px[7], px[8]};
  // background pixels carry diffuse = -1 in every channel (render.py:197, :215)
  const float alpha = (kd[0] >= 0.0f || kd[1] >= 0.0f || kd[2] >= 0.0f) ? 1.0f : 0.0f;
  float n[3], len;
  normalize3(n_raw, n, len);
  float rgb[3] = {0.0f, 0.0f, 0.0f};
  for (int l = 0; l < L; ++l) {
    const float d[3] = {sm.pos[l][0] - pos[0], sm.pos[l][1] - pos[1], sm.pos[l][2] - pos[2]};
    float u[3];
    normalize3(d, u, len);
    const float ndl = fminf(fmaxf(n[0] * u[0] + n[1] * u[1] + n[2] * u[2], 0.0f), 1.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] += kd[c] * ndl * sm.intensity[l][c];
  }
  if (ambient != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = sm.ambient[c] * kd[c] + rgb[c];
  }
  float4 o;
  const bool valid = alpha > 0.5f;
  o.x = valid ? rgb[0] : 0.0f; o.y = valid ? rgb[1] : 0.0f; o.z = valid ? rgb[2] : 0.0f; o.w = alpha;
  rgba[((size_t)b * H + (H - 1 - y)) * W + x] = o;
}

// d(rgba) -> d(pixels) (channels 0..8; further channels, if any, are written as zero).
__global__ void __launch_bounds__(256)
shade_diffuse_backward_kernel(const float4 *__restrict__ grad_rgba, const float *__restrict__ pixels,
                              const float *__restrict__ light_positions, const float *__restrict__ light_intensities,
                              const float *__restrict__ ambient, int L, int A, int W, int H,
                              float *__restrict__ d_pixels) {
  __shared__ Lights sm;
  extern __shared__ float stage[];          // [256][A]: the CTA's pixels are contiguous in d_pixels
  const int b = blockIdx.y;
  load_lights(sm, light_positions, light_intensities, ambient, b, L);
  const int p0 = blockIdx.x * blockDim.x, p = p0 + threadIdx.x;
  const bool in_image = p < W * H;
  const int y = in_image ? p / W : 0, x = in_image ? p - y * W : 0;
  const float *px = pixels + ((size_t)b * H * W + (in_image ? p : 0)) * A;
  float *out = stage + threadIdx.x * A;
  const float n_raw[3] = {px[0], px[1], px[2]}, pos[3] = {px[3], px[4], px[5]}, kd[3] = {px[6], px[7], px[8]};
  float d_n[3] = {0.0f, 0.0f, 0.0f}, d_pos[3] = {0.0f, 0.0f, 0.0f}, d_kd[3] = {0.0f, 0.0f, 0.0f};
  const bool valid = in_image && (kd[0] >= 0.0f || kd[1] >= 0.0f || kd[2] >= 0.0f);
  if (valid) {
    const float4 g4 = grad_rgba[((size_t)b * H + (H - 1 - y)) * W + x];
    const float g[3] = {g4.x, g4.y, g4.z};                 // alpha comes from a comparison: no gradient
    float n[3], n_len;
    const float n_inv = normalize3(n_raw, n, n_len);
    float d_unit_n[3] = {0.0f, 0.0f, 0.0f};
    for (int l = 0; l < L; ++l) {
      const float d[3] = {sm.pos[l][0] - pos[0], sm.pos[l][1] - pos[1], sm.pos[l][2] - pos[2]};
      float u[3], d_len;
      const float d_inv = normalize3(d, u, d_len);
      const float s = n[0] * u[0] + n[1] * u[1] + n[2] * u[2];
      const float ndl = fminf(fmaxf(s, 0.0f), 1.0f);
      float d_ndl = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        d_kd[c] += g[c] * (ndl * sm.intensity[l][c]);
        d_ndl += g[c] * (kd[c] * sm.intensity[l][c]);
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
    if (ambient != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) d_kd[c] += g[c] * sm.ambient[c];
    }
    normalize3_backward(n, n_len, n_inv, d_unit_n, d_n);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { out[k] = d_n[k]; out[3 + k] = d_pos[k]; out[6 + k] = d_kd[k]; }
  for (int k = 9; k < A; ++k) out[k] = 0.0f;
  __syncthreads();
  // coalesced write-back of the CTA's contiguous run of pixels (A is odd or even: staging rows of A floats
  // are read linearly, so there are no bank conflicts here; the writes above conflict only for even A)
  const int n_pixels = min((int)blockDim.x, W * H - p0);
  float *dst = d_pixels + ((size_t)b * H * W + p0) * A;
  for (int i = threadIdx.x; i < n_pixels * A; i += blockDim.x) dst[i] = stage[i];
}

int shade_diffuse_forward_impl(Context *ctx, const float *pixels, const float *light_positions,
                               const float *light_intensities, const float *ambient, int B, int L, int A, int W,
                               int H, float *rgba, cudaStream_t stream) {
  if (B == 0) return PMR_OK;
  if (L > kMaxLights) return set_error(ctx, PMR_ERR_SIZE, "at most %d lights", kMaxLights);
  StageScope timed(ctx, PMR_STAGE_SHADE, stream);
  shade_diffuse_forward_kernel<<<dim3((unsigned)(((long long)W * H + 255) / 256), B), 256, 0, stream>>>(
      pixels, light_positions, light_intensities, ambient, L, A, W, H, reinterpret_cast<float4 *>(rgba));
  ctx->launches += 1;
  return check_launch(ctx, "shade_diffuse_forward_kernel");
}

int shade_diffuse_backward_impl(Context *ctx, const float *grad_rgba, const float *pixels,
                                const float *light_positions, const float *light_intensities, const float *ambient,
                                int B, int L, int A, int W, int H, float *d_pixels, cudaStream_t stream) {
  if (B == 0) return PMR_OK;
  if (L > kMaxLights) return set_error(ctx, PMR_ERR_SIZE, "at most %d lights", kMaxLights);
  StageScope timed(ctx, PMR_STAGE_SHADE, stream);
  const size_t smem = (size_t)256 * A * sizeof(float);
  if (smem > 40 * 1024) return set_error(ctx, PMR_ERR_SIZE, "at most 40 pixel channels");
  shade_diffuse_backward_kernel<<<dim3((unsigned)(((long long)W * H + 255) / 256), B), 256, smem, stream>>>(
      reinterpret_cast<const float4 *>(grad_rgba), pixels, light_positions, light_intensities, ambient, L, A, W, H,
      d_pixels);
  ctx->launches += 1;
  return check_launch(ctx, "shade_diffuse_backward_kernel");
}

}  // namespace pmr
