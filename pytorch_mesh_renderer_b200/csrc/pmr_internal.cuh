// pmr_internal.cuh -- context, workspace and error plumbing shared by the translation units
// behind include/pmr_b200.h.  No torch types anywhere: the library is plain CUDA runtime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/pmr_b200.h"

namespace pmr {

// Screen tile of the large-triangle pass (one CTA at a time): 16x16 pixels = 8 warps x (8x4) pixels.
constexpr int kTileW = 16, kTileH = 16, kTileShiftX = 4, kTileShiftY = 4;

struct Context;
int set_error(Context *ctx, int code, const char *fmt, ...);

// Grow-only device buffer owned by a context (stream-ordered use: one context serves one stream
// at a time, like the single-threaded reference entry points it replaces).
struct Buffer {
  void *ptr = nullptr;
  size_t bytes = 0;
  int reserve(Context *ctx, size_t need);
  void release();
};

struct StageInterval {
  cudaEvent_t begin, end;
  int stage;
};

struct Context {
  int device = 0;
  int sm_count = 148;
  long long peer_wait_cycles = 19000000000ll;   // reduce_partials_kernel gives up after this many SM cycles (~10 s)
  int no_tma = 0;                      // PMR_NO_TMA: per-pixel loads instead of tensor-map boxes (tests of that path)
  int strip_blocks_override = 0;       // PMR_STRIP_BLOCKS: 8x4 blocks per warp of the backward (and, up to 4, resolve) kernel (0 = automatic)
  int small_mesh_threshold = 64;      // T at or below this: the tile kernel alone, every tile walks all triangles
  Buffer bins, scratch, keys, centers, staging;   // staging: device copies of the host entry point's buffers
  int centers_w = -1, centers_h = -1;  // image size the pixel-centre table was built for
  // host entry point: uploads and downloads run on their own streams beside the caller's (kernels)
  cudaStream_t copy_stream = nullptr, down_stream = nullptr;
  cudaEvent_t call_begin = nullptr;
  std::vector<cudaEvent_t> host_events;     // per slice: inputs up, gradient up, forward done, backward done
  const int *last_large_count = nullptr;    // device: large triangles per image of the last pipeline forward
  int last_large_images = 0;
  long long launches = 0;             // kernels launched through this context (bench gpu_launches)
  char error[512] = {0};
  // stage timing (pmr_enable_stage_timing)
  bool timing = false;
  std::vector<StageInterval> intervals;     // recorded, not yet read
  std::vector<cudaEvent_t> spare_events;
  double stage_ms[PMR_STAGE_COUNT] = {};
  long long stage_n[PMR_STAGE_COUNT] = {};
};

// RAII bracket around one stage: records an event pair on `stream` when timing is enabled.
struct StageScope {
  Context *ctx;
  cudaStream_t stream;
  StageInterval iv;
  bool active;
  StageScope(Context *c, int stage, cudaStream_t s);
  ~StageScope();
};

int check_launch(Context *ctx, const char *what);

#define PMR_CUDA(ctx, call)                                                                   \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ::pmr::set_error((ctx), PMR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// The render path: lights for shading inside the resolve / backward kernels (9 attribute channels).
struct ShadeArgs {
  const float *light_positions, *light_intensities, *ambient;   // [B,L,3], [B,L,3], [B,3] or nullptr
  int L;
  float *rgba;               // forward: [B,H,W,4] out
  const float *grad_rgba;    // backward: [B,H,W,4] in
  const float *background;   // backward: [9]
};

// raster_forward.cu
int forward_impl(Context *ctx, const float *verts, const int32_t *tris, int B, int V, int T, int W, int H,
                 int32_t *ids, float *bary, float *z, const float *attrs, const float *bg, int A,
                 float *image, cudaStream_t stream, const ShadeArgs *shade = nullptr);
int interpolate_impl(Context *ctx, const float *attrs, const int32_t *tris, const int32_t *ids,
                     const float *bary, const float *bg, int B, int V, int A, int W, int H, float *out,
                     cudaStream_t stream);
// raster_backward.cu
int backward_impl(Context *ctx, const float *df_dbary, const float *grad_image, const float *verts,
                  const float *attrs, const int32_t *tris, const int32_t *ids, const float *bary,
                  int B, int V, int T, int A, int W, int H, float *d_verts, float *d_attrs, int mode,
                  cudaStream_t stream, const ShadeArgs *shade = nullptr);

// vertex_stage.cu
int transform_forward_impl(Context *ctx, const float *matrices, const float *world, int B, int V, int shared,
                           float *clip, cudaStream_t stream);
int transform_backward_impl(Context *ctx, const float *matrices, const float *d_clip, int B, int V, int shared,
                            float *d_world, cudaStream_t stream);

// mesh_normals.cu
int vertex_incidence_impl(Context *ctx, const int32_t *tris, int T, int V, int32_t *offsets, int32_t *incidence,
                          cudaStream_t stream);
int vertex_normals_forward_impl(Context *ctx, const float *verts, const int32_t *tris, const int32_t *offsets,
                                const int32_t *incidence, int B, int V, float *raw, float *normals,
                                cudaStream_t stream);
int vertex_normals_backward_impl(Context *ctx, const float *grad_normals, const float *raw, const float *verts,
                                 const int32_t *tris, const int32_t *offsets, const int32_t *incidence, int B, int V,
                                 float *grad_raw, float *d_verts, cudaStream_t stream);

// peer_exchange.cu
size_t peer_exchange_bytes(long long n_floats, int world);
int transform_backward_exchange_impl(Context *ctx, const float *matrices, const float *d_clip, int B, int V,
                                     void *const *peers, int rank, int world, long long epoch, float *d_world,
                                     cudaStream_t stream);

// shade.cu
int shade_diffuse_forward_impl(Context *ctx, const float *pixels, const float *light_positions,
                               const float *light_intensities, const float *ambient, int B, int L, int A, int W,
                               int H, float *rgba, cudaStream_t stream);
int shade_diffuse_backward_impl(Context *ctx, const float *grad_rgba, const float *pixels,
                                const float *light_positions, const float *light_intensities, const float *ambient,
                                int B, int L, int A, int W, int H, float *d_pixels, cudaStream_t stream);

int shade_phong_forward_impl(Context *ctx, const float *pixels, const float *light_positions,
                             const float *light_intensities, const float *ambient, const float *camera,
                             const float *shininess, int B, int L, int A, int W, int H, float *norm2, float *rgba,
                             cudaStream_t stream);
int shade_phong_backward_impl(Context *ctx, const float *grad_rgba, const float *pixels, const float *light_positions,
                              const float *light_intensities, const float *ambient, const float *camera,
                              const float *shininess, const float *norm2, int B, int L, int A, int W, int H,
                              float *sum_gx, float *d_pixels, cudaStream_t stream);

}  // namespace pmr
