// shade.cu -- the direct caller of the hot path: per-pixel Phong lighting of the interpolated
// attribute image, diffuse + ambient terms (reference src/mesh_renderer/render.py:201-228 and
// phong_shader :231-386 without the specular branch, i.e. `render` called without specular colors --
// the configuration of mesh_renderer_test.py:30-70 and BASELINE config c1).
//
// The reference runs ~20 torch ops that materialise [B,L,P,3] temporaries several times the size of
// the rasterizer's own output; here one kernel reads the 9 interpolated channels of a pixel
// ([normal, world position, diffuse colour], render.py:181) and writes RGBA, and one kernel maps
// d(RGBA) back to d(pixel channels).  Gradients with respect to the lights / ambient colour / camera are
// not produced here (the Python layer keeps the torch-op path for callers that ask for them).
//
// The specular branch (render.py:326-372; 12 or 13 channels: + specular colour (+ per-vertex shininess))
// normalises the reflection . view products of each (image, light) by their L2 norm over ALL pixels of the
// image (:347-353), so it takes two passes each way: shade_phong_norm_kernel accumulates the squared
// products, shade_phong_forward_kernel shades; backward, shade_phong_sums_kernel accumulates
// sum(d t . x) per (image, light) -- the term through which every pixel's gradient reaches every other
// pixel -- and shade_phong_backward_kernel finishes.  The per-(image, light) sums are float atomics: their
// order, hence their last bits, varies from run to run.
//
// Arithmetic follows the torch ops of the reference term by term (F.normalize with eps = 1e-12,
// clamp to [0, 1] with gradient passed on the closed interval, lights summed in order), but libraries
// differ (torch CPU there), so parity is within the float tolerance stated in tests/test_gpu_shade.py,
// not bit-exact.
#include "pmr_internal.cuh"
#include "shade_math.cuh"

namespace pmr {

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
  const float4 o = shade_diffuse_pixel(n_raw, pos, kd, sm, L, ambient != nullptr);
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
  if (in_image) {
    const float4 g4 = grad_rgba[((size_t)b * H + (H - 1 - y)) * W + x];
    const float g[3] = {g4.x, g4.y, g4.z};                 // alpha comes from a comparison: no gradient
    shade_diffuse_pixel_backward(n_raw, pos, kd, g, sm, L, ambient != nullptr, d_n, d_pos, d_kd);
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

// ---------------------------------------------------------------------------------------------
// Diffuse + ambient + specular
// ---------------------------------------------------------------------------------------------

struct PhongScene {
  Lights lights;
  float camera[3];
  float shininess;          // per-image value (used when the pixel carries no shininess channel)
  float norm2[kMaxLights];  // sum over the image's pixels of (reflection . view)^2, per light
  float sum_gx[kMaxLights]; // backward: sum over pixels of d(loss)/d(t) * x, per light
};

__device__ __forceinline__ void load_scene(PhongScene &sm, const float *__restrict__ light_positions,
                                           const float *__restrict__ light_intensities, const float *__restrict__ ambient,
                                           const float *__restrict__ camera, const float *__restrict__ shininess,
                                           const float *__restrict__ norm2, const float *__restrict__ sum_gx, int b, int L) {
  if (threadIdx.x < 3) sm.camera[threadIdx.x] = camera[(size_t)b * 3 + threadIdx.x];
  if (threadIdx.x == 3) sm.shininess = shininess != nullptr ? shininess[b] : 0.0f;
  if ((int)threadIdx.x < L) {
    sm.norm2[threadIdx.x] = norm2 != nullptr ? norm2[(size_t)b * L + threadIdx.x] : 0.0f;
    sm.sum_gx[threadIdx.x] = sum_gx != nullptr ? sum_gx[(size_t)b * L + threadIdx.x] : 0.0f;
  }
  load_lights(sm.lights, light_positions, light_intensities, ambient, b, L);      // ends with __syncthreads()
}

// Everything the specular term of one (pixel, light) needs, forward and backward.
struct SpecularTerm {
  float u[3], u_len, u_inv;      // unit vector to the light
  float s, ndl;                  // n . u and its clamp
  float m[3], mhat[3], m_len, m_inv;   // mirror direction 2 ndl n - u and its normalisation
  float x;                       // mhat . chat  (render.py:343-345)
};

__device__ __forceinline__ void specular_term(const float n[3], const float pos[3], const float chat[3],
                                              const float light[3], SpecularTerm &t) {
  const float d[3] = {light[0] - pos[0], light[1] - pos[1], light[2] - pos[2]};
  t.u_inv = normalize3(d, t.u, t.u_len);
  t.s = n[0] * t.u[0] + n[1] * t.u[1] + n[2] * t.u[2];
  t.ndl = fminf(fmaxf(t.s, 0.0f), 1.0f);
#pragma unroll
  for (int k = 0; k < 3; ++k) t.m[k] = 2.0f * t.ndl * n[k] - t.u[k];
  t.m_inv = normalize3(t.m, t.mhat, t.m_len);
  t.x = t.mhat[0] * chat[0] + t.mhat[1] * chat[1] + t.mhat[2] * chat[2];
}

// Pixel index and validity shared by the four kernels.
struct PhongPixel {
  bool in_image, valid;
  int x, y;
  float n_raw[3], n[3], n_len, n_inv, pos[3], kd[3], ks[3], shin;
  float c[3], chat[3], c_len, c_inv;
};

__device__ __forceinline__ void load_phong_pixel(const float *__restrict__ pixels, int b, int p, int A, int W, int H,
                                                 const PhongScene &sm, PhongPixel &q) {
  q.in_image = p < W * H;
  q.y = q.in_image ? p / W : 0;
  q.x = q.in_image ? p - q.y * W : 0;
  const float *px = pixels + ((size_t)b * H * W + (q.in_image ? p : 0)) * A;
#pragma unroll
  for (int k = 0; k < 3; ++k) { q.n_raw[k] = px[k]; q.pos[k] = px[3 + k]; q.kd[k] = px[6 + k]; q.ks[k] = px[9 + k]; }
  q.shin = A > 12 ? px[12] : sm.shininess;
  q.valid = q.in_image && (q.kd[0] >= 0.0f || q.kd[1] >= 0.0f || q.kd[2] >= 0.0f);
  q.n_inv = normalize3(q.n_raw, q.n, q.n_len);
#pragma unroll
  for (int k = 0; k < 3; ++k) q.c[k] = sm.camera[k] - q.pos[k];
  q.c_inv = normalize3(q.c, q.chat, q.c_len);
}

// Block-wide sum of one value per light into global accumulators (one atomic per CTA and light).
__device__ __forceinline__ void accumulate_per_light(const float *mine, int L, float *scratch /*[8][kMaxLights]*/,
                                                     float *__restrict__ global_sums) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int l = 0; l < L; ++l) {
    float v = mine[l];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) scratch[warp * kMaxLights + l] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < L) {
    float v = 0.0f;
    for (int w = 0; w < 8; ++w) v += scratch[w * kMaxLights + threadIdx.x];
    atomicAdd(global_sums + threadIdx.x, v);
  }
}

// Pass 1 forward: norm2[b][l] += sum over pixels of x^2  (every pixel of the image counts, background too).
__global__ void __launch_bounds__(256)
shade_phong_norm_kernel(const float *__restrict__ pixels, const float *__restrict__ light_positions,
                        const float *__restrict__ light_intensities, const float *__restrict__ camera,
                        int L, int A, int W, int H, float *__restrict__ norm2) {
  __shared__ PhongScene sm;
  __shared__ float scratch[8 * kMaxLights];
  const int b = blockIdx.y;
  load_scene(sm, light_positions, light_intensities, nullptr, camera, nullptr, nullptr, nullptr, b, L);
  PhongPixel q;
  load_phong_pixel(pixels, b, blockIdx.x * blockDim.x + threadIdx.x, A, W, H, sm, q);
  float mine[kMaxLights];
  for (int l = 0; l < L; ++l) {
    SpecularTerm t;
    specular_term(q.n, q.pos, q.chat, sm.lights.pos[l], t);
    mine[l] = q.in_image ? t.x * t.x : 0.0f;
  }
  accumulate_per_light(mine, L, scratch, norm2 + (size_t)b * L);
}

// y = where(ndl != 0, clamp(x / denom, 0, 1), 0); specularity = y^shininess  (render.py:347-366)
__device__ __forceinline__ float specular_strength(const SpecularTerm &t, float denom, float shin, float &y) {
  const float tt = t.x / denom;
  y = t.ndl != 0.0f ? fminf(fmaxf(tt, 0.0f), 1.0f) : 0.0f;
  return powf(y, shin);
}

__global__ void __launch_bounds__(256)
shade_phong_forward_kernel(const float *__restrict__ pixels, const float *__restrict__ light_positions,
                           const float *__restrict__ light_intensities, const float *__restrict__ ambient,
                           const float *__restrict__ camera, const float *__restrict__ shininess,
                           const float *__restrict__ norm2, int L, int A, int W, int H, float4 *__restrict__ rgba) {
  __shared__ PhongScene sm;
  const int b = blockIdx.y;
  load_scene(sm, light_positions, light_intensities, ambient, camera, shininess, norm2, nullptr, b, L);
  PhongPixel q;
  load_phong_pixel(pixels, b, blockIdx.x * blockDim.x + threadIdx.x, A, W, H, sm, q);
  if (!q.in_image) return;
  float rgb[3] = {0.0f, 0.0f, 0.0f}, spec[3] = {0.0f, 0.0f, 0.0f};
  for (int l = 0; l < L; ++l) {
    SpecularTerm t;
    specular_term(q.n, q.pos, q.chat, sm.lights.pos[l], t);
    float y;
    const float sp = specular_strength(t, fmaxf(sqrtf(sm.norm2[l]), kNormalizeEps), q.shin, y);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rgb[c] += q.kd[c] * t.ndl * sm.lights.intensity[l][c];
      spec[c] += q.ks[c] * sp * sm.lights.intensity[l][c];
    }
  }
  if (ambient != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = sm.lights.ambient[c] * q.kd[c] + rgb[c];
  }
  float4 o;
  o.x = q.valid ? rgb[0] + spec[0] : 0.0f;
  o.y = q.valid ? rgb[1] + spec[1] : 0.0f;
  o.z = q.valid ? rgb[2] + spec[2] : 0.0f;
  o.w = q.valid ? 1.0f : 0.0f;
  rgba[((size_t)b * H + (H - 1 - q.y)) * W + q.x] = o;
}

// d(loss)/d(t) of one (pixel, light), t = x / denom before the clamp: through the specular sum, the power
// (torch.pow's gradient: exponent * base^(exponent-1), zero where the exponent is zero), the `where` on
// n.l != 0 and the clamp to [0, 1] (gradient on the closed interval).  Also the gradient with respect to a
// per-pixel shininess (result * log(base), zero where base == 0 and exponent >= 0).
__device__ __forceinline__ float specular_dt(const SpecularTerm &t, float denom, float shin, const float g[3],
                                             const float ks[3], const float intensity[3], bool valid,
                                             float &sp, float &d_shin) {
  float y;
  sp = specular_strength(t, denom, shin, y);
  d_shin = 0.0f;
  (void)valid;      // masked pixels arrive with g = 0 and run the same formulas, so that 0 * inf turns into NaN
                    // exactly where torch's autograd produces it (background pixels carry shininess -1)
  const float d_sp = g[0] * (ks[0] * intensity[0]) + g[1] * (ks[1] * intensity[1]) + g[2] * (ks[2] * intensity[2]);
  if (!(y == 0.0f && shin >= 0.0f)) d_shin = d_sp * (sp * logf(y));
  const float d_y = shin == 0.0f ? 0.0f : d_sp * (shin * powf(y, shin - 1.0f));
  const float tt = t.x / denom;
  return (t.ndl != 0.0f && tt >= 0.0f && tt <= 1.0f) ? d_y : 0.0f;
}

// Pass 1 backward: sum_gx[b][l] += sum over pixels of d(loss)/d(t) * x.
__global__ void __launch_bounds__(256)
shade_phong_sums_kernel(const float4 *__restrict__ grad_rgba, const float *__restrict__ pixels,
                        const float *__restrict__ light_positions, const float *__restrict__ light_intensities,
                        const float *__restrict__ camera, const float *__restrict__ shininess,
                        const float *__restrict__ norm2, int L, int A, int W, int H, float *__restrict__ sum_gx) {
  __shared__ PhongScene sm;
  __shared__ float scratch[8 * kMaxLights];
  const int b = blockIdx.y;
  load_scene(sm, light_positions, light_intensities, nullptr, camera, shininess, norm2, nullptr, b, L);
  PhongPixel q;
  load_phong_pixel(pixels, b, blockIdx.x * blockDim.x + threadIdx.x, A, W, H, sm, q);
  float g[3] = {0.0f, 0.0f, 0.0f};
  if (q.valid) {
    const float4 g4 = grad_rgba[((size_t)b * H + (H - 1 - q.y)) * W + q.x];
    g[0] = g4.x; g[1] = g4.y; g[2] = g4.z;
  }
  float mine[kMaxLights];
  for (int l = 0; l < L; ++l) {
    SpecularTerm t;
    specular_term(q.n, q.pos, q.chat, sm.lights.pos[l], t);
    float sp, d_shin;
    const float dt = specular_dt(t, fmaxf(sqrtf(sm.norm2[l]), kNormalizeEps), q.shin, g, q.ks, sm.lights.intensity[l],
                                 q.valid, sp, d_shin);
    mine[l] = dt * t.x;
  }
  accumulate_per_light(mine, L, scratch, sum_gx + (size_t)b * L);
}

__global__ void __launch_bounds__(256)
shade_phong_backward_kernel(const float4 *__restrict__ grad_rgba, const float *__restrict__ pixels,
                            const float *__restrict__ light_positions, const float *__restrict__ light_intensities,
                            const float *__restrict__ ambient, const float *__restrict__ camera,
                            const float *__restrict__ shininess, const float *__restrict__ norm2,
                            const float *__restrict__ sum_gx, int L, int A, int W, int H, float *__restrict__ d_pixels) {
  __shared__ PhongScene sm;
  extern __shared__ float stage[];          // [256][A]
  const int b = blockIdx.y;
  load_scene(sm, light_positions, light_intensities, ambient, camera, shininess, norm2, sum_gx, b, L);
  const int p0 = blockIdx.x * blockDim.x;
  PhongPixel q;
  load_phong_pixel(pixels, b, p0 + threadIdx.x, A, W, H, sm, q);
  float g[3] = {0.0f, 0.0f, 0.0f};
  if (q.valid) {
    const float4 g4 = grad_rgba[((size_t)b * H + (H - 1 - q.y)) * W + q.x];
    g[0] = g4.x; g[1] = g4.y; g[2] = g4.z;
  }
  float d_unit_n[3] = {0.0f, 0.0f, 0.0f}, d_pos[3] = {0.0f, 0.0f, 0.0f}, d_kd[3] = {0.0f, 0.0f, 0.0f};
  float d_ks[3] = {0.0f, 0.0f, 0.0f}, d_chat[3] = {0.0f, 0.0f, 0.0f}, d_shin_total = 0.0f;
  for (int l = 0; l < L; ++l) {
    SpecularTerm t;
    specular_term(q.n, q.pos, q.chat, sm.lights.pos[l], t);
    const float norm = sqrtf(sm.norm2[l]), denom = fmaxf(norm, kNormalizeEps);
    float sp, d_shin;
    const float dt = specular_dt(t, denom, q.shin, g, q.ks, sm.lights.intensity[l], q.valid, sp, d_shin);
    d_shin_total += d_shin;
    // through x / clamp_min(|x|_2, eps) over the image: d x = d t / denom - [norm >= eps] (S / denom^2) (x / norm)
    float d_x = dt / denom;
    if (norm >= kNormalizeEps && norm > 0.0f) d_x -= (sm.sum_gx[l] / (denom * denom)) * (t.x / norm);
    if (!q.in_image) d_x = 0.0f;
    // x = mhat . chat
    const float d_mhat[3] = {d_x * q.chat[0], d_x * q.chat[1], d_x * q.chat[2]};
    float d_m[3];
    normalize3_backward(t.mhat, t.m_len, t.m_inv, d_mhat, d_m);
    // m = 2 ndl n - u
    float d_ndl = 2.0f * (d_m[0] * q.n[0] + d_m[1] * q.n[1] + d_m[2] * q.n[2]);
    float d_u[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d_chat[k] += d_x * t.mhat[k];
      d_unit_n[k] += 2.0f * t.ndl * d_m[k];
      d_u[k] = -d_m[k];
    }
    // diffuse and specular colour terms (masked pixels have g = 0; 0 * inf stays NaN as in torch's autograd)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d_kd[c] += g[c] * (t.ndl * sm.lights.intensity[l][c]);
      d_ks[c] += g[c] * (sp * sm.lights.intensity[l][c]);
      d_ndl += g[c] * (q.kd[c] * sm.lights.intensity[l][c]);
    }
    const float d_s = (t.s >= 0.0f && t.s <= 1.0f) ? d_ndl : 0.0f;
    float d_d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { d_u[k] += d_s * q.n[k]; d_unit_n[k] += d_s * t.u[k]; }
    normalize3_backward(t.u, t.u_len, t.u_inv, d_u, d_d);
#pragma unroll
    for (int k = 0; k < 3; ++k) d_pos[k] -= d_d[k];                      // d = light - position
  }
  if (ambient != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) d_kd[c] += g[c] * sm.lights.ambient[c];
  }
  float d_n[3], d_c[3];
  normalize3_backward(q.n, q.n_len, q.n_inv, d_unit_n, d_n);
  normalize3_backward(q.chat, q.c_len, q.c_inv, d_chat, d_c);
  float *out = stage + threadIdx.x * A;
#pragma unroll
  for (int k = 0; k < 3; ++k) { out[k] = d_n[k]; out[3 + k] = d_pos[k] - d_c[k]; out[6 + k] = d_kd[k]; out[9 + k] = d_ks[k]; }
  if (A > 12) out[12] = d_shin_total;
  for (int k = 13; k < A; ++k) out[k] = 0.0f;
  __syncthreads();
  const int n_pixels = min((int)blockDim.x, W * H - p0);
  float *dst = d_pixels + ((size_t)b * H * W + p0) * A;
  for (int i = threadIdx.x; i < n_pixels * A; i += blockDim.x) dst[i] = stage[i];
}

int shade_phong_forward_impl(Context *ctx, const float *pixels, const float *light_positions,
                             const float *light_intensities, const float *ambient, const float *camera,
                             const float *shininess, int B, int L, int A, int W, int H, float *norm2, float *rgba,
                             cudaStream_t stream) {
  if (B == 0) return PMR_OK;
  if (L > kMaxLights) return set_error(ctx, PMR_ERR_SIZE, "at most %d lights", kMaxLights);
  StageScope timed(ctx, PMR_STAGE_SHADE, stream);
  const dim3 grid((unsigned)(((long long)W * H + 255) / 256), B);
  PMR_CUDA(ctx, cudaMemsetAsync(norm2, 0, (size_t)B * L * sizeof(float), stream));
  shade_phong_norm_kernel<<<grid, 256, 0, stream>>>(pixels, light_positions, light_intensities, camera, L, A, W, H, norm2);
  shade_phong_forward_kernel<<<grid, 256, 0, stream>>>(pixels, light_positions, light_intensities, ambient, camera,
                                                       shininess, norm2, L, A, W, H, reinterpret_cast<float4 *>(rgba));
  ctx->launches += 2;
  return check_launch(ctx, "shade_phong_forward_kernel");
}

int shade_phong_backward_impl(Context *ctx, const float *grad_rgba, const float *pixels, const float *light_positions,
                              const float *light_intensities, const float *ambient, const float *camera,
                              const float *shininess, const float *norm2, int B, int L, int A, int W, int H,
                              float *sum_gx, float *d_pixels, cudaStream_t stream) {
  if (B == 0) return PMR_OK;
  if (L > kMaxLights) return set_error(ctx, PMR_ERR_SIZE, "at most %d lights", kMaxLights);
  const size_t smem = (size_t)256 * A * sizeof(float);
  if (smem > 40 * 1024) return set_error(ctx, PMR_ERR_SIZE, "at most 40 pixel channels");
  StageScope timed(ctx, PMR_STAGE_SHADE, stream);
  const dim3 grid((unsigned)(((long long)W * H + 255) / 256), B);
  PMR_CUDA(ctx, cudaMemsetAsync(sum_gx, 0, (size_t)B * L * sizeof(float), stream));
  shade_phong_sums_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(grad_rgba), pixels, light_positions,
                                                    light_intensities, camera, shininess, norm2, L, A, W, H, sum_gx);
  shade_phong_backward_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<const float4 *>(grad_rgba), pixels,
                                                           light_positions, light_intensities, ambient, camera, shininess,
                                                           norm2, sum_gx, L, A, W, H, d_pixels);
  ctx->launches += 2;
  return check_launch(ctx, "shade_phong_backward_kernel");
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
