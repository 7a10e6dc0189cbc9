// c_api.cu -- extern "C" surface of libpmr_b200.so (include/pmr_b200.h): context management,
// argument validation and the host-buffer round trip.  Kernels live in raster_forward.cu and
// raster_backward.cu.
#include <stdarg.h>

#include <new>

#include <stdlib.h>

#include "pmr_internal.cuh"

struct pmr_context : public pmr::Context {};

namespace pmr {

int set_error(Context *ctx, int code, const char *fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->error, sizeof(ctx->error), fmt, ap);
    va_end(ap);
  }
  return code;
}

int check_launch(Context *ctx, const char *what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(ctx, PMR_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return PMR_OK;
}

int Buffer::reserve(Context *ctx, size_t need) {
  if (need <= bytes) return PMR_OK;
  // Grow geometrically; the old block may still be in use by enqueued work of this context's
  // stream, and cudaFree synchronises the device before releasing it.
  size_t want = bytes + bytes / 2;
  if (want < need) want = need;
  want = (want + 255) & ~(size_t)255;
  if (ptr) {
    cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  const cudaError_t e = cudaMalloc(&ptr, want);
  if (e != cudaSuccess) {
    ptr = nullptr;
    return set_error(ctx, PMR_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
  }
  bytes = want;
  return PMR_OK;
}

void Buffer::release() {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  bytes = 0;
}

static cudaEvent_t take_event(Context *ctx) {
  cudaEvent_t e = nullptr;
  if (!ctx->spare_events.empty()) {
    e = ctx->spare_events.back();
    ctx->spare_events.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}

StageScope::StageScope(Context *c, int stage, cudaStream_t s) : ctx(c), stream(s), active(c && c->timing) {
  if (!active) return;
  iv.stage = stage;
  iv.begin = take_event(ctx);
  iv.end = take_event(ctx);
  cudaEventRecord(iv.begin, stream);
}

StageScope::~StageScope() {
  if (!active) return;
  cudaEventRecord(iv.end, stream);
  ctx->intervals.push_back(iv);
}

static int validate_common(Context *ctx, int B, int V, int T, int W, int H) {
  if (!ctx) return PMR_ERR_INVALID;
  if (B < 0 || V < 0 || T < 0) return set_error(ctx, PMR_ERR_INVALID, "negative batch/vertex/triangle count");
  // rasterize.py:98-101 raises ValueError for non-positive sizes.
  if (W <= 0) return set_error(ctx, PMR_ERR_INVALID, "Image width must be > 0.");
  if (H <= 0) return set_error(ctx, PMR_ERR_INVALID, "Image height must be > 0.");
  if (W > 32768 || H > 32768) return set_error(ctx, PMR_ERR_SIZE, "image larger than 32768 pixels on a side");
  if (B > 65535) return set_error(ctx, PMR_ERR_SIZE, "more than 65535 images per call (the image index is a grid dimension)");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  return PMR_OK;
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace pmr

using pmr::Context;
using pmr::set_error;

extern "C" {

int pmr_version(void) { return 100; }

int pmr_create(int device, pmr_context **out) {
  if (!out) return PMR_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return PMR_ERR_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) return PMR_ERR_CUDA;
  pmr_context *ctx = new (std::nothrow) pmr_context();
  if (!ctx) return PMR_ERR_INVALID;
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (const char *v = getenv("PMR_STRIP_BLOCKS")) ctx->strip_blocks_override = atoi(v);    // tuning experiments only
  if (const char *v = getenv("PMR_NO_TMA")) ctx->no_tma = atoi(v);
  {
    // how long a rank waits for its peers' partial sums before it declares the exchange broken (seconds)
    double seconds = 10.0;
    if (const char *v = getenv("PMR_PEER_WAIT_SECONDS")) seconds = atof(v) > 0.0 ? atof(v) : seconds;
    int khz = 1900000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    ctx->peer_wait_cycles = (long long)(seconds * 1e3 * (double)khz);
  }
  *out = ctx;
  return PMR_OK;
}

void pmr_destroy(pmr_context *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  ctx->bins.release();
  ctx->scratch.release();
  ctx->keys.release();
  ctx->centers.release();
  ctx->staging.release();
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->down_stream) cudaStreamDestroy(ctx->down_stream);
  for (cudaEvent_t e : ctx->host_events) cudaEventDestroy(e);
  if (ctx->call_begin) cudaEventDestroy(ctx->call_begin);
  for (const pmr::StageInterval &iv : ctx->intervals) { cudaEventDestroy(iv.begin); cudaEventDestroy(iv.end); }
  for (cudaEvent_t e : ctx->spare_events) cudaEventDestroy(e);
  delete ctx;
}

const char *pmr_last_error(const pmr_context *ctx) { return ctx ? ctx->error : "null context"; }
long long pmr_launch_count(const pmr_context *ctx) { return ctx ? ctx->launches : 0; }
long long pmr_last_large_triangles(pmr_context *ctx) {
  // Diagnostics only: blocking copy of the per-image counters the last pipeline forward left on the device.
  if (!ctx) return PMR_ERR_INVALID;
  if (ctx->last_large_count == nullptr || ctx->last_large_images <= 0) return 0;
  std::vector<int> counts((size_t)ctx->last_large_images);
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(counts.data(), ctx->last_large_count, counts.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return set_error(ctx, PMR_ERR_CUDA, "reading the large-triangle counters failed: %s", cudaGetErrorString(cudaGetLastError()));
  long long total = 0;
  for (int c : counts) total += c;
  return total;
}

int pmr_enable_stage_timing(pmr_context *ctx, int enable) {
  if (!ctx) return PMR_ERR_INVALID;
  ctx->timing = enable != 0;
  return PMR_OK;
}

int pmr_read_stage_timing(pmr_context *ctx, double *ms, long long *counts, int reset) {
  if (!ctx) return PMR_ERR_INVALID;
  for (const pmr::StageInterval &iv : ctx->intervals) {
    float t = 0.0f;
    PMR_CUDA(ctx, cudaEventSynchronize(iv.end));
    PMR_CUDA(ctx, cudaEventElapsedTime(&t, iv.begin, iv.end));
    ctx->stage_ms[iv.stage] += t;
    ctx->stage_n[iv.stage] += 1;
    ctx->spare_events.push_back(iv.begin);
    ctx->spare_events.push_back(iv.end);
  }
  ctx->intervals.clear();
  for (int k = 0; k < PMR_STAGE_COUNT; ++k) {
    if (ms) ms[k] = ctx->stage_ms[k];
    if (counts) counts[k] = ctx->stage_n[k];
    if (reset) { ctx->stage_ms[k] = 0.0; ctx->stage_n[k] = 0; }
  }
  return PMR_OK;
}

int pmr_set_small_mesh_threshold(pmr_context *ctx, int triangles) {
  if (!ctx || triangles < 0) return PMR_ERR_INVALID;
  ctx->small_mesh_threshold = triangles;
  return PMR_OK;
}

int pmr_rasterize_forward(pmr_context *ctx, const float *vertices, const int32_t *triangles, int B, int V,
                          int T, int W, int H, int32_t *ids, float *bary, float *z, void *stream) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (B == 0) return PMR_OK;
  if (!ids || !bary || !z || (T > 0 && (!vertices || !triangles)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(vertices)) return set_error(ctx, PMR_ERR_INVALID, "vertices must be 16-byte aligned");
  return pmr::forward_impl(ctx, vertices, triangles, B, V, T, W, H, ids, bary, z, nullptr, nullptr, 0, nullptr,
                           (cudaStream_t)stream);
}

int pmr_rasterize_backward(pmr_context *ctx, const float *df_dbary, const float *vertices,
                           const int32_t *triangles, const int32_t *ids, const float *bary, int B, int V, int T,
                           int W, int H, float *df_dvertices, int mode, void *stream) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (B == 0 || V == 0) return PMR_OK;
  if (!df_dbary || !vertices || !ids || !bary || !df_dvertices || (T > 0 && !triangles))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(vertices)) return set_error(ctx, PMR_ERR_INVALID, "vertices must be 16-byte aligned");
  return pmr::backward_impl(ctx, df_dbary, nullptr, vertices, nullptr, triangles, ids, bary, B, V, T, 0, W, H,
                            df_dvertices, nullptr, mode, (cudaStream_t)stream);
}

int pmr_interpolate_forward(pmr_context *ctx, const float *attributes, const int32_t *triangles,
                            const int32_t *ids, const float *bary, const float *background, int B, int V, int T,
                            int A, int W, int H, float *image, void *stream) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (A < 0) return set_error(ctx, PMR_ERR_INVALID, "negative attribute count");
  if (B == 0 || A == 0) return PMR_OK;
  if (!attributes || !triangles || !ids || !bary || !background || !image)
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (T == 0) return set_error(ctx, PMR_ERR_INVALID, "interpolation needs at least one triangle");
  return pmr::interpolate_impl(ctx, attributes, triangles, ids, bary, background, B, V, A, W, H, image,
                               (cudaStream_t)stream);
}

int pmr_rasterize_interpolate_forward(pmr_context *ctx, const float *vertices, const float *attributes,
                                      const int32_t *triangles, const float *background, int B, int V, int T,
                                      int A, int W, int H, int32_t *ids, float *bary, float *z, float *image,
                                      void *stream) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (A <= 0) return set_error(ctx, PMR_ERR_INVALID, "attribute count must be > 0");
  if (B == 0) return PMR_OK;
  if (!ids || !bary || !z || !image || !background || (T > 0 && (!vertices || !triangles || !attributes)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(vertices)) return set_error(ctx, PMR_ERR_INVALID, "vertices must be 16-byte aligned");
  return pmr::forward_impl(ctx, vertices, triangles, B, V, T, W, H, ids, bary, z, attributes, background, A,
                           image, (cudaStream_t)stream);
}

int pmr_rasterize_interpolate_backward(pmr_context *ctx, const float *grad_image, const float *vertices,
                                       const float *attributes, const int32_t *triangles, const int32_t *ids,
                                       const float *bary, int B, int V, int T, int A, int W, int H,
                                       float *d_vertices, float *d_attributes, int mode, void *stream) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (A <= 0) return set_error(ctx, PMR_ERR_INVALID, "attribute count must be > 0");
  if (B == 0 || V == 0) return PMR_OK;
  if (!grad_image || !vertices || !attributes || !ids || !bary || (T > 0 && !triangles))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(vertices)) return set_error(ctx, PMR_ERR_INVALID, "vertices must be 16-byte aligned");
  return pmr::backward_impl(ctx, nullptr, grad_image, vertices, attributes, triangles, ids, bary, B, V, T, A, W,
                            H, d_vertices, d_attributes, mode, (cudaStream_t)stream);
}

int pmr_rasterize_clip_space_host(pmr_context *ctx, const float *vertices, const float *attributes,
                                  const int32_t *triangles, const float *background, const float *grad_image,
                                  int B, int V, int T, int A, int W, int H, float *image, float *d_vertices,
                                  float *d_attributes, int32_t *ids, float *bary, float *z, int mode,
                                  void *stream_) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (A <= 0) return set_error(ctx, PMR_ERR_INVALID, "attribute count must be > 0");
  if (B == 0) return PMR_OK;
  if (!vertices || !attributes || !background || !image || (T > 0 && !triangles))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t P = (size_t)B * H * W;
  const size_t n_v = (size_t)B * V * 4 * sizeof(float), n_a = (size_t)B * V * A * sizeof(float);
  const size_t n_t = (size_t)T * 3 * sizeof(int32_t), n_bg = (size_t)A * sizeof(float);
  const size_t n_img = P * A * sizeof(float), n_ids = P * sizeof(int32_t), n_bary = P * 3 * sizeof(float);
  auto up = [](size_t n) { return (n + 255) & ~(size_t)255; };
  const bool bwd = grad_image != nullptr;
  const size_t need = up(n_v) + up(n_a) + up(n_t) + up(n_bg) + up(n_img) * (bwd ? 2 : 1) + up(n_ids) +
                      up(n_bary) + up(n_ids) + (bwd ? up(n_v) + up(n_a) : 0) + 256;
  rc = ctx->staging.reserve(ctx, need);      // device staging area of the host entry point, owned by the context
  if (rc) return rc;
  char *cur = (char *)ctx->staging.ptr;
  auto take = [&](size_t n) { char *p = cur; cur += up(n); return p; };
  float *d_v = (float *)take(n_v), *d_a = (float *)take(n_a);
  int32_t *d_t = (int32_t *)take(n_t);
  float *d_bg = (float *)take(n_bg), *d_img = (float *)take(n_img);
  int32_t *d_ids = (int32_t *)take(n_ids);
  float *d_bary = (float *)take(n_bary), *d_z = (float *)take(n_ids);
  float *d_g = bwd ? (float *)take(n_img) : nullptr;
  float *d_dv = bwd ? (float *)take(n_v) : nullptr, *d_da = bwd ? (float *)take(n_a) : nullptr;

  // The batch is cut into slices that flow through three streams: uploads (copy_stream), kernels
  // (the caller's stream) and downloads (down_stream).  While slice k is rasterized, slice k+1's inputs
  // and image gradient are on their way up and slice k-1's image is on its way down: PCIe is full
  // duplex and both copy engines stay busy for the whole call, so the call costs about
  // max(bytes up, bytes down) / link bandwidth instead of the sum of its phases.
  if (!ctx->copy_stream) {
    PMR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    PMR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    PMR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->call_begin, cudaEventDisableTiming));
  }
  constexpr int kMaxSlices = 8;
  const int n_slices = B < kMaxSlices ? B : kMaxSlices;
  while ((int)ctx->host_events.size() < 4 * kMaxSlices) {
    cudaEvent_t e;
    PMR_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->host_events.push_back(e);
  }
  cudaStream_t upload = ctx->copy_stream, download = ctx->down_stream;
  PMR_CUDA(ctx, cudaEventRecord(ctx->call_begin, stream));                 // staging area free from here on
  PMR_CUDA(ctx, cudaStreamWaitEvent(upload, ctx->call_begin, 0));
  if (n_t) PMR_CUDA(ctx, cudaMemcpyAsync(d_t, triangles, n_t, cudaMemcpyHostToDevice, upload));
  PMR_CUDA(ctx, cudaMemcpyAsync(d_bg, background, n_bg, cudaMemcpyHostToDevice, upload));
  const size_t px = (size_t)H * W;                                          // per image
  // uploads of all slices in the order the kernels need them (the copy engine is a FIFO): a slice's
  // mesh data, then its image gradient.  (Measured on the B200 box, c2, 688 MB each way: this order
  // 14.9 ms per call; all mesh data first, so that the downloads start at once and both directions
  // are saturated for the whole call, 15.9 ms; one slice 15.9 ms; 2..16 slices within 0.3 ms.)
  for (int k = 0; k < n_slices; ++k) {
    const int b0 = (int)((long long)B * k / n_slices), b1 = (int)((long long)B * (k + 1) / n_slices);
    const size_t nb = (size_t)(b1 - b0);
    PMR_CUDA(ctx, cudaMemcpyAsync(d_v + (size_t)b0 * V * 4, vertices + (size_t)b0 * V * 4, nb * V * 4 * sizeof(float),
                                  cudaMemcpyHostToDevice, upload));
    PMR_CUDA(ctx, cudaMemcpyAsync(d_a + (size_t)b0 * V * A, attributes + (size_t)b0 * V * A, nb * V * A * sizeof(float),
                                  cudaMemcpyHostToDevice, upload));
    PMR_CUDA(ctx, cudaEventRecord(ctx->host_events[4 * k + 0], upload));
    if (bwd) {
      PMR_CUDA(ctx, cudaMemcpyAsync(d_g + (size_t)b0 * px * A, grad_image + (size_t)b0 * px * A,
                                    nb * px * A * sizeof(float), cudaMemcpyHostToDevice, upload));
      PMR_CUDA(ctx, cudaEventRecord(ctx->host_events[4 * k + 1], upload));
    }
  }
  for (int k = 0; k < n_slices; ++k) {
    const int b0 = (int)((long long)B * k / n_slices), b1 = (int)((long long)B * (k + 1) / n_slices);
    const int nb = b1 - b0;
    cudaEvent_t in_ready = ctx->host_events[4 * k + 0], grad_ready = ctx->host_events[4 * k + 1];
    cudaEvent_t fwd_done = ctx->host_events[4 * k + 2], bwd_done = ctx->host_events[4 * k + 3];
    const size_t p0 = (size_t)b0 * px, np = (size_t)nb * px;
    PMR_CUDA(ctx, cudaStreamWaitEvent(stream, in_ready, 0));
    rc = pmr::forward_impl(ctx, d_v + (size_t)b0 * V * 4, d_t, nb, V, T, W, H, d_ids + p0, d_bary + 3 * p0, d_z + p0,
                           d_a + (size_t)b0 * V * A, d_bg, A, d_img + p0 * A, stream);
    if (rc) return rc;
    PMR_CUDA(ctx, cudaEventRecord(fwd_done, stream));
    PMR_CUDA(ctx, cudaStreamWaitEvent(download, fwd_done, 0));
    PMR_CUDA(ctx, cudaMemcpyAsync(image + p0 * A, d_img + p0 * A, np * A * sizeof(float), cudaMemcpyDeviceToHost, download));
    if (ids) PMR_CUDA(ctx, cudaMemcpyAsync(ids + p0, d_ids + p0, np * sizeof(int32_t), cudaMemcpyDeviceToHost, download));
    if (bary) PMR_CUDA(ctx, cudaMemcpyAsync(bary + 3 * p0, d_bary + 3 * p0, np * 3 * sizeof(float), cudaMemcpyDeviceToHost, download));
    if (z) PMR_CUDA(ctx, cudaMemcpyAsync(z + p0, d_z + p0, np * sizeof(float), cudaMemcpyDeviceToHost, download));
    if (bwd) {
      PMR_CUDA(ctx, cudaStreamWaitEvent(stream, grad_ready, 0));
      float *dv_k = d_vertices ? d_dv + (size_t)b0 * V * 4 : nullptr, *da_k = d_attributes ? d_da + (size_t)b0 * V * A : nullptr;
      rc = pmr::backward_impl(ctx, nullptr, d_g + p0 * A, d_v + (size_t)b0 * V * 4, d_a + (size_t)b0 * V * A, d_t,
                              d_ids + p0, d_bary + 3 * p0, nb, V, T, A, W, H, dv_k, da_k, mode, stream);
      if (rc) return rc;
      PMR_CUDA(ctx, cudaEventRecord(bwd_done, stream));
      PMR_CUDA(ctx, cudaStreamWaitEvent(download, bwd_done, 0));
      if (d_vertices)
        PMR_CUDA(ctx, cudaMemcpyAsync(d_vertices + (size_t)b0 * V * 4, dv_k, (size_t)nb * V * 4 * sizeof(float),
                                      cudaMemcpyDeviceToHost, download));
      if (d_attributes)
        PMR_CUDA(ctx, cudaMemcpyAsync(d_attributes + (size_t)b0 * V * A, da_k, (size_t)nb * V * A * sizeof(float),
                                      cudaMemcpyDeviceToHost, download));
    }
  }
  PMR_CUDA(ctx, cudaStreamSynchronize(stream));
  PMR_CUDA(ctx, cudaStreamSynchronize(download));
  return PMR_OK;
}

int pmr_transform_forward(pmr_context *ctx, const float *matrices, const float *world_vertices, int B, int V,
                          int shared, float *clip_vertices, void *stream) {
  if (!ctx) return PMR_ERR_INVALID;
  if (B < 0 || V < 0) return set_error(ctx, PMR_ERR_INVALID, "negative batch/vertex count");
  if (B == 0 || V == 0) return PMR_OK;
  if (!matrices || !world_vertices || !clip_vertices) return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(clip_vertices)) return set_error(ctx, PMR_ERR_INVALID, "clip_vertices must be 16-byte aligned");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  return pmr::transform_forward_impl(ctx, matrices, world_vertices, B, V, shared, clip_vertices, (cudaStream_t)stream);
}

int pmr_transform_backward(pmr_context *ctx, const float *matrices, const float *d_clip_vertices, int B, int V,
                           int shared, float *d_world_vertices, void *stream) {
  if (!ctx) return PMR_ERR_INVALID;
  if (B < 0 || V < 0) return set_error(ctx, PMR_ERR_INVALID, "negative batch/vertex count");
  if (B == 0 || V == 0) return PMR_OK;
  if (!matrices || !d_clip_vertices || !d_world_vertices) return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(d_clip_vertices)) return set_error(ctx, PMR_ERR_INVALID, "d_clip_vertices must be 16-byte aligned");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  return pmr::transform_backward_impl(ctx, matrices, d_clip_vertices, B, V, shared, d_world_vertices,
                                      (cudaStream_t)stream);
}

size_t pmr_peer_exchange_bytes(long long n_floats, int world) {
  if (n_floats < 0 || world < 1 || world > PMR_MAX_PEERS) return 0;
  return pmr::peer_exchange_bytes(n_floats, world);
}

int pmr_peer_alloc(pmr_context *ctx, size_t bytes, void **ptr, void *handle) {
  if (!ctx) return PMR_ERR_INVALID;
  if (!ptr || !handle || bytes == 0) return set_error(ctx, PMR_ERR_INVALID, "null pointer or zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == PMR_PEER_HANDLE_BYTES, "handle size");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  void *p = nullptr;
  PMR_CUDA(ctx, cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle), p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_error(ctx, PMR_ERR_CUDA, "peer buffer setup failed: %s", cudaGetErrorString(e));
  }
  *ptr = p;
  return PMR_OK;
}

int pmr_peer_open(pmr_context *ctx, const void *handle, void **ptr) {
  if (!ctx) return PMR_ERR_INVALID;
  if (!ptr || !handle) return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  PMR_CUDA(ctx, cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PMR_OK;
}

int pmr_peer_close(pmr_context *ctx, void *ptr) {
  if (!ctx) return PMR_ERR_INVALID;
  if (!ptr) return PMR_OK;
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  PMR_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
  return PMR_OK;
}

int pmr_peer_free(pmr_context *ctx, void *ptr) {
  if (!ctx) return PMR_ERR_INVALID;
  if (!ptr) return PMR_OK;
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  PMR_CUDA(ctx, cudaFree(ptr));
  return PMR_OK;
}

int pmr_peer_status(pmr_context *ctx, const void *own_buffer, int *status) {
  if (!ctx) return PMR_ERR_INVALID;
  if (!own_buffer || !status) return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  PMR_CUDA(ctx, cudaMemcpy(status, static_cast<const char *>(own_buffer) + 128, sizeof(int), cudaMemcpyDeviceToHost));
  return PMR_OK;
}

int pmr_transform_backward_exchange(pmr_context *ctx, const float *matrices, const float *d_clip_vertices, int B,
                                    int V, void *const *peer_buffers, int rank, int world, long long epoch,
                                    float *d_world_vertices, void *stream) {
  if (!ctx) return PMR_ERR_INVALID;
  if (B < 0 || V < 0) return set_error(ctx, PMR_ERR_INVALID, "negative batch/vertex count");
  if (world < 1 || world > PMR_MAX_PEERS || rank < 0 || rank >= world)
    return set_error(ctx, PMR_ERR_INVALID, "rank / world outside 0 <= rank < world <= PMR_MAX_PEERS");
  if (epoch < 0) return set_error(ctx, PMR_ERR_INVALID, "epochs count from 1 (0: counted on the device)");
  if (epoch >= (1ll << 30))   // the flags carry the epoch as a 30-bit stamp compared with <
    return set_error(ctx, PMR_ERR_SIZE, "2^30 exchange steps on one set of buffers: allocate a new set");
  if (!matrices || !d_clip_vertices || !d_world_vertices || !peer_buffers)
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  for (int r = 0; r < world; ++r)
    if (!peer_buffers[r]) return set_error(ctx, PMR_ERR_INVALID, "null peer buffer");
  if (!pmr::aligned16(d_clip_vertices)) return set_error(ctx, PMR_ERR_INVALID, "d_clip_vertices must be 16-byte aligned");
  if (V == 0) return PMR_OK;
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  return pmr::transform_backward_exchange_impl(ctx, matrices, d_clip_vertices, B, V, peer_buffers, rank, world, epoch,
                                               d_world_vertices, (cudaStream_t)stream);
}

int pmr_vertex_incidence(pmr_context *ctx, const int32_t *triangles, int T, int V, int32_t *offsets,
                         int32_t *incidence, void *stream) {
  if (!ctx) return PMR_ERR_INVALID;
  if (T < 0 || V < 0) return set_error(ctx, PMR_ERR_INVALID, "negative vertex/triangle count");
  if (T >= (1 << 30)) return set_error(ctx, PMR_ERR_SIZE, "more than 2^30 triangles");
  if (!offsets || (T > 0 && (!triangles || !incidence))) return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  return pmr::vertex_incidence_impl(ctx, triangles, T, V, offsets, incidence, (cudaStream_t)stream);
}

int pmr_vertex_normals_forward(pmr_context *ctx, const float *vertices, const int32_t *triangles,
                               const int32_t *offsets, const int32_t *incidence, int B, int V, int T, float *raw,
                               float *normals, void *stream) {
  if (!ctx) return PMR_ERR_INVALID;
  if (B < 0 || V < 0 || T < 0) return set_error(ctx, PMR_ERR_INVALID, "negative batch/vertex/triangle count");
  if (B > 65535) return set_error(ctx, PMR_ERR_SIZE, "more than 65535 meshes per call (the mesh index is a grid dimension)");
  if (B == 0 || V == 0) return PMR_OK;
  if (!vertices || !offsets || !normals || (T > 0 && (!triangles || !incidence)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  return pmr::vertex_normals_forward_impl(ctx, vertices, triangles, offsets, incidence, B, V, raw, normals,
                                          (cudaStream_t)stream);
}

int pmr_vertex_normals_backward(pmr_context *ctx, const float *grad_normals, const float *raw, const float *vertices,
                                const int32_t *triangles, const int32_t *offsets, const int32_t *incidence, int B,
                                int V, int T, float *grad_raw, float *d_vertices, void *stream) {
  if (!ctx) return PMR_ERR_INVALID;
  if (B < 0 || V < 0 || T < 0) return set_error(ctx, PMR_ERR_INVALID, "negative batch/vertex/triangle count");
  if (B > 65535) return set_error(ctx, PMR_ERR_SIZE, "more than 65535 meshes per call (the mesh index is a grid dimension)");
  if (B == 0 || V == 0) return PMR_OK;
  if (!grad_normals || !raw || !vertices || !offsets || !grad_raw || !d_vertices || (T > 0 && (!triangles || !incidence)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  PMR_CUDA(ctx, cudaSetDevice(ctx->device));
  return pmr::vertex_normals_backward_impl(ctx, grad_normals, raw, vertices, triangles, offsets, incidence, B, V,
                                           grad_raw, d_vertices, (cudaStream_t)stream);
}

int pmr_shade_diffuse_forward(pmr_context *ctx, const float *pixels, const float *light_positions,
                              const float *light_intensities, const float *ambient, int B, int L, int A, int W,
                              int H, float *rgba, void *stream) {
  int rc = pmr::validate_common(ctx, B, 0, 0, W, H);
  if (rc) return rc;
  if (A < 9 || L < 0) return set_error(ctx, PMR_ERR_INVALID, "shading needs at least 9 pixel channels and L >= 0");
  if (B == 0) return PMR_OK;
  if (!pixels || !rgba || (L > 0 && (!light_positions || !light_intensities)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(rgba)) return set_error(ctx, PMR_ERR_INVALID, "rgba must be 16-byte aligned");
  return pmr::shade_diffuse_forward_impl(ctx, pixels, light_positions, light_intensities, ambient, B, L, A, W, H, rgba,
                                         (cudaStream_t)stream);
}

int pmr_shade_diffuse_backward(pmr_context *ctx, const float *grad_rgba, const float *pixels,
                               const float *light_positions, const float *light_intensities, const float *ambient,
                               int B, int L, int A, int W, int H, float *d_pixels, void *stream) {
  int rc = pmr::validate_common(ctx, B, 0, 0, W, H);
  if (rc) return rc;
  if (A < 9 || L < 0) return set_error(ctx, PMR_ERR_INVALID, "shading needs at least 9 pixel channels and L >= 0");
  if (B == 0) return PMR_OK;
  if (!grad_rgba || !pixels || !d_pixels || (L > 0 && (!light_positions || !light_intensities)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(grad_rgba)) return set_error(ctx, PMR_ERR_INVALID, "grad_rgba must be 16-byte aligned");
  return pmr::shade_diffuse_backward_impl(ctx, grad_rgba, pixels, light_positions, light_intensities, ambient, B, L, A,
                                          W, H, d_pixels, (cudaStream_t)stream);
}

int pmr_shade_phong_forward(pmr_context *ctx, const float *pixels, const float *light_positions,
                            const float *light_intensities, const float *ambient, const float *camera_position,
                            const float *shininess, int B, int L, int A, int W, int H, float *norm2, float *rgba,
                            void *stream) {
  int rc = pmr::validate_common(ctx, B, 0, 0, W, H);
  if (rc) return rc;
  if (A < 12 || L < 0) return set_error(ctx, PMR_ERR_INVALID, "specular shading needs at least 12 pixel channels and L >= 0");
  if (A == 12 && !shininess) return set_error(ctx, PMR_ERR_INVALID, "12 channels carry no shininess: pass it per image");
  if (B == 0) return PMR_OK;
  if (!pixels || !rgba || !norm2 || !camera_position || (L > 0 && (!light_positions || !light_intensities)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(rgba)) return set_error(ctx, PMR_ERR_INVALID, "rgba must be 16-byte aligned");
  return pmr::shade_phong_forward_impl(ctx, pixels, light_positions, light_intensities, ambient, camera_position,
                                       A > 12 ? nullptr : shininess, B, L, A, W, H, norm2, rgba, (cudaStream_t)stream);
}

int pmr_shade_phong_backward(pmr_context *ctx, const float *grad_rgba, const float *pixels,
                             const float *light_positions, const float *light_intensities, const float *ambient,
                             const float *camera_position, const float *shininess, const float *norm2, int B, int L,
                             int A, int W, int H, float *sum_gx, float *d_pixels, void *stream) {
  int rc = pmr::validate_common(ctx, B, 0, 0, W, H);
  if (rc) return rc;
  if (A < 12 || L < 0) return set_error(ctx, PMR_ERR_INVALID, "specular shading needs at least 12 pixel channels and L >= 0");
  if (A == 12 && !shininess) return set_error(ctx, PMR_ERR_INVALID, "12 channels carry no shininess: pass it per image");
  if (B == 0) return PMR_OK;
  if (!grad_rgba || !pixels || !d_pixels || !norm2 || !sum_gx || !camera_position ||
      (L > 0 && (!light_positions || !light_intensities)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(grad_rgba)) return set_error(ctx, PMR_ERR_INVALID, "grad_rgba must be 16-byte aligned");
  return pmr::shade_phong_backward_impl(ctx, grad_rgba, pixels, light_positions, light_intensities, ambient,
                                        camera_position, A > 12 ? nullptr : shininess, norm2, B, L, A, W, H, sum_gx,
                                        d_pixels, (cudaStream_t)stream);
}

int pmr_render_diffuse_forward(pmr_context *ctx, const float *vertices, const float *attributes,
                               const int32_t *triangles, const float *background, const float *light_positions,
                               const float *light_intensities, const float *ambient, int B, int V, int T, int L,
                               int W, int H, int32_t *ids, float *bary, float *z, float *rgba, void *stream) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (L < 0 || L > 16) return set_error(ctx, PMR_ERR_SIZE, "0 to 16 lights");
  if (B == 0) return PMR_OK;
  if (!ids || !bary || !z || !rgba || !background || (T > 0 && (!vertices || !triangles || !attributes)) ||
      (L > 0 && (!light_positions || !light_intensities)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(vertices) || !pmr::aligned16(rgba))
    return set_error(ctx, PMR_ERR_INVALID, "vertices and rgba must be 16-byte aligned");
  pmr::ShadeArgs shade = {light_positions, light_intensities, ambient, L, rgba, nullptr, background};
  return pmr::forward_impl(ctx, vertices, triangles, B, V, T, W, H, ids, bary, z, attributes, background, 9, nullptr,
                           (cudaStream_t)stream, &shade);
}

int pmr_render_diffuse_backward(pmr_context *ctx, const float *grad_rgba, const float *vertices,
                                const float *attributes, const int32_t *triangles, const float *background,
                                const float *light_positions, const float *light_intensities, const float *ambient,
                                const int32_t *ids, const float *bary, int B, int V, int T, int L, int W, int H,
                                float *d_vertices, float *d_attributes, void *stream) {
  int rc = pmr::validate_common(ctx, B, V, T, W, H);
  if (rc) return rc;
  if (L < 0 || L > 16) return set_error(ctx, PMR_ERR_SIZE, "0 to 16 lights");
  if (B == 0 || V == 0) return PMR_OK;
  if (!grad_rgba || !vertices || !attributes || !ids || !bary || !background || (T > 0 && !triangles) ||
      (L > 0 && (!light_positions || !light_intensities)))
    return set_error(ctx, PMR_ERR_INVALID, "null pointer argument");
  if (!pmr::aligned16(vertices) || !pmr::aligned16(grad_rgba))
    return set_error(ctx, PMR_ERR_INVALID, "vertices and grad_rgba must be 16-byte aligned");
  pmr::ShadeArgs shade = {light_positions, light_intensities, ambient, L, nullptr, grad_rgba, background};
  return pmr::backward_impl(ctx, nullptr, nullptr, vertices, attributes, triangles, ids, bary, B, V, T, 9, W, H,
                            d_vertices, d_attributes, PMR_BACKWARD_ATOMIC, (cudaStream_t)stream, &shade);
}

}  // extern "C"
