/*
 * pmr_b200.h -- C ABI of libpmr_b200.so, the sm_100a implementation of pytorch_mesh_renderer's
 * barycentric rasterization hot path.
 *
 * This is the drop-in boundary.  The reference crosses exactly one language boundary on this
 * path: Python -> `rasterize_triangles_cpp.forward / .backward`
 *   (/root/reference/src/mesh_renderer/kernels/rasterize_triangles.cpp:302-307, :131-137,
 *    exported at :421-424; called from rasterize_triangles_ext.py:39 and :56),
 * and then interpolates attributes with torch ops in rasterize.py:118-150.  The entry points
 * below are what a binding for that path binds: plain pointers and sizes, no torch types.
 * They are batched over images (the reference loops `for b in range(batch_size)` in Python,
 * rasterize.py:112); B = 1 is the reference's unbatched call.
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers on the context's device unless the function name
 *     ends in `_host`, in which case they are host pointers and the call copies in and out.
 *   - Layouts are the reference's, row-major and densely packed:
 *       vertices   float32 [B, V, 4]   clip-space x y z w          (K.cpp:279-284)
 *       triangles  int32   [T, 3]      shared by all images        (K.cpp:285-287)
 *       ids        int32   [B, H, W]   0 also means "no triangle"  (K.cpp:290-294)
 *       bary       float32 [B, H, W, 3]  (0,0,0) where empty       (K.cpp:295-298)
 *       z          float32 [B, H, W]     1.0 where empty           (K.cpp:299-301)
 *       attributes float32 [B, V, A];  image float32 [B, H, W, A]  (rasterize.py:70-95)
 *     Row iy = 0 is NDC y = -1.  Argument order is (width, height) like the reference's.
 *   - `vertices` must be 16-byte aligned.  Vertex indices are not range-checked (neither does the
 *     reference, K.cpp:331-337).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Kernels are
 *     enqueued on it and nothing on the device-pointer entry points waits for the device (the
 *     passes are capturable in a CUDA graph once the workspace has grown to size).  A context
 *     must not be used from two streams at once.
 *   - Every entry point that takes a context makes the context's device the calling thread's current CUDA
 *     device (cudaSetDevice) and leaves it so; callers that juggle several devices restore their own.
 *   - Every function returns PMR_OK (0) or a negative PMR_ERR_* code; pmr_last_error() gives the
 *     message.  There is no CPU fallback: without a CUDA device pmr_create fails.
 *   - Outputs are fully written by the callee (no pre-initialisation required).
 */
#ifndef PMR_B200_H_
#define PMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PMR_API __attribute__((visibility("default")))
#else
#define PMR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define PMR_OK 0
#define PMR_ERR_INVALID (-1) /* bad argument: NULL pointer, non-positive size, misalignment */
#define PMR_ERR_CUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define PMR_ERR_SIZE (-3)    /* problem too large for 32-bit tile bookkeeping */
#define PMR_ERR_NO_DEVICE (-4)

/* Backward accumulation modes. */
#define PMR_BACKWARD_ATOMIC 0  /* throughput mode: warp-aggregated atomics, fp32 sums in arbitrary order */
#define PMR_BACKWARD_ORDERED 1 /* parity mode: every vertex component is summed sequentially in ascending
                                  pixel order, corner 0..2 within a pixel -- the reference's order
                                  (K.cpp:156-157, 232-269), bit-reproducible */

typedef struct pmr_context pmr_context;

PMR_API int pmr_version(void);

/* Creates a context (grow-only device workspace) on CUDA device `device`. */
PMR_API int pmr_create(int device, pmr_context **out);
PMR_API void pmr_destroy(pmr_context *ctx);
PMR_API const char *pmr_last_error(const pmr_context *ctx);

/* Number of kernels launched through this context so far (bench.py's gpu_launches). */
PMR_API long long pmr_launch_count(const pmr_context *ctx);
/* Triangles whose pixel box exceeded 16x16 in the last pipeline forward pass, summed over its images
 * (the ones raster_tile_kernel drew).  Diagnostics: synchronises the device. */
PMR_API long long pmr_last_large_triangles(pmr_context *ctx);
/* Meshes with at most this many triangles (default 64, effective maximum 1024) run the tile kernel alone. */
PMR_API int pmr_set_small_mesh_threshold(pmr_context *ctx, int triangles);

/*
 * Per-stage device timing (CUDA events recorded on the launching stream around each stage while
 * enabled).  bench.py uses it to attribute step time to kernels for the roofline report.
 * pmr_read_stage_timing waits for the recorded events, ADDS the elapsed milliseconds and the
 * number of timed intervals per stage into ms[PMR_STAGE_COUNT] / counts[PMR_STAGE_COUNT]'s
 * running totals held by the context, copies the totals out and, if `reset` is non-zero,
 * clears them.
 */
#define PMR_STAGE_BIN 0      /* clearing the depth keys and the large-triangle counters */
#define PMR_STAGE_RASTER 1   /* raster_tile_kernel (big triangles; whole pass for tiny meshes) */
#define PMR_STAGE_BACKWARD 2 /* backward kernels (atomic: one kernel; ordered: radix sort by vertex + in-order fold) */
#define PMR_STAGE_INTERP 3   /* standalone interpolate_kernel */
#define PMR_STAGE_SCATTER 4  /* scatter_small_kernel (small triangles -> depth keys; lists the large ones) */
#define PMR_STAGE_RESOLVE 5  /* resolve_kernel (depth keys -> ids / bary / z / interpolated image) */
#define PMR_STAGE_SHADE 6    /* shade_diffuse kernels (the Phong caller of the path) */
#define PMR_STAGE_COUNT 7
PMR_API int pmr_enable_stage_timing(pmr_context *ctx, int enable);
PMR_API int pmr_read_stage_timing(pmr_context *ctx, double *ms, long long *counts, int reset);

/*
 * rasterize_triangles forward.  Replaces rasterize_triangles_cpp.forward
 * (rasterize_triangles.cpp:302-419) for B images at once.
 */
PMR_API int pmr_rasterize_forward(pmr_context *ctx, const float *vertices, const int32_t *triangles,
                          int B, int V, int T, int image_width, int image_height,
                          int32_t *ids, float *bary, float *z, void *stream);

/*
 * rasterize_triangles backward.  Replaces rasterize_triangles_cpp.backward
 * (rasterize_triangles.cpp:131-273): df_dbary [B,H,W,3] -> df_dvertices [B,V,4]
 * (columns x, y, w; the z column is written as zero).
 */
PMR_API int pmr_rasterize_backward(pmr_context *ctx, const float *df_dbary, const float *vertices,
                           const int32_t *triangles, const int32_t *ids, const float *bary,
                           int B, int V, int T, int image_width, int image_height,
                           float *df_dvertices, int mode, void *stream);

/*
 * Attribute interpolation from existing raster buffers.  Replaces the torch-op chain of
 * rasterize_clip_space (rasterize.py:118-150): gather corner attributes, weight by barycentrics,
 * alpha = clamp(2*sum(b), 0, 1), blend with `background` [A].
 */
PMR_API int pmr_interpolate_forward(pmr_context *ctx, const float *attributes, const int32_t *triangles,
                            const int32_t *ids, const float *bary, const float *background,
                            int B, int V, int T, int A, int image_width, int image_height,
                            float *image, void *stream);

/*
 * Fused rasterize + interpolate: rasterize_clip_space (rasterize.py:66-152) in one pass.  Also
 * returns the raster buffers the backward pass needs.
 */
PMR_API int pmr_rasterize_interpolate_forward(pmr_context *ctx, const float *vertices, const float *attributes,
                                      const int32_t *triangles, const float *background,
                                      int B, int V, int T, int A, int image_width, int image_height,
                                      int32_t *ids, float *bary, float *z, float *image, void *stream);

/*
 * Backward of rasterize_clip_space: grad_image [B,H,W,A] -> d_vertices [B,V,4] and
 * d_attributes [B,V,A] in one pass (the reference runs autograd through rasterize.py:130-150 and
 * then rasterize_triangles.cpp:131-273).  Either output may be NULL to skip it.
 */
PMR_API int pmr_rasterize_interpolate_backward(pmr_context *ctx, const float *grad_image, const float *vertices,
                                       const float *attributes, const int32_t *triangles,
                                       const int32_t *ids, const float *bary,
                                       int B, int V, int T, int A, int image_width, int image_height,
                                       float *d_vertices, float *d_attributes, int mode, void *stream);

/*
 * Host-buffer round trip of rasterize_clip_space forward + backward: copies vertices, attributes,
 * triangles, background and grad_image to the device, runs the fused forward and backward, copies
 * image, d_vertices and d_attributes (and, when non-NULL, ids / bary / z) back.  This is the call
 * bench.py times for `e2e`.  All pointers are HOST pointers (pinned memory makes the copies
 * asynchronous); the call returns after the results are in host memory.
 */
PMR_API int pmr_rasterize_clip_space_host(pmr_context *ctx, const float *vertices, const float *attributes,
                                  const int32_t *triangles, const float *background,
                                  const float *grad_image,
                                  int B, int V, int T, int A, int image_width, int image_height,
                                  float *image, float *d_vertices, float *d_attributes,
                                  int32_t *ids, float *bary, float *z, int mode, void *stream);

/*
 * Vertex stage in front of the path: world -> clip space, clip[b][v] = M_b * (x, y, z, 1)
 * (reference src/common/camera_utils.py:142-170 transform_homogeneous), and its backward with respect to
 * the vertices.  matrices float32 [B,4,4] row-major; `shared` != 0: world_vertices is ONE mesh [V,3] seen
 * by all B views and d_world [V,3] receives the gradient summed over the views (the buffer that the
 * multi-GPU path all-reduces); `shared` == 0: world_vertices / d_world are [B,V,3].
 */
PMR_API int pmr_transform_forward(pmr_context *ctx, const float *matrices, const float *world_vertices,
                                  int B, int V, int shared, float *clip_vertices, void *stream);
PMR_API int pmr_transform_backward(pmr_context *ctx, const float *matrices, const float *d_clip_vertices,
                                   int B, int V, int shared, float *d_world_vertices, void *stream);

/*
 * Multi-GPU exchange of the shared-mesh gradient over peer memory (one process per GPU of
 * one NVLink / NVSwitch box).  pmr_transform_backward_exchange is pmr_transform_backward with shared = 1 FUSED with
 * the sum over the ranks: the kernel that reduces this rank's views stores its partial [V,3] into a slot of every
 * peer's exchange buffer (plain stores over NVLink) and raises a flag there; a second kernel waits for the world's
 * flags and adds the slots in rank order into d_world_vertices.  Every rank ends with the bit-identical sum over
 * all ranks' views; there is no collective-library call and no host synchronisation in the step.
 *
 * Setup, once: every rank allocates an exchange buffer of pmr_peer_exchange_bytes(3 * V, world) bytes with
 * pmr_peer_alloc (zero-filled; `handle` receives PMR_PEER_HANDLE_BYTES bytes to send to the peers by any means),
 * opens every peer's handle with pmr_peer_open, and the ranks meet at a host barrier before the first step.
 * peer_buffers[r] is rank r's buffer as mapped in THIS process (own allocation at [rank]).  `epoch` counts the
 * calls on this set of buffers from 1 (below 2^30: PMR_ERR_SIZE after that, allocate a new set) and must be the
 * same number on every rank for the same step; 0 lets the device count (a step counter in the rank's own
 * buffer): the call then has no per-step argument and can be captured in a CUDA graph and replayed.  Use one
 * convention per set of buffers.  A rank whose peer does not deliver within PMR_PEER_WAIT_SECONDS (default 10 s)
 * gives up waiting instead of hanging the device: pmr_peer_status reports 1 from then on and this and every
 * later call on the buffers fills d_world_vertices with NaN -- check the status before using the gradient.
 * Teardown: host barrier, pmr_peer_close on the opened pointers, pmr_peer_free on the own one.  world <= PMR_MAX_PEERS.
 */
#define PMR_MAX_PEERS 16
#define PMR_PEER_HANDLE_BYTES 64
PMR_API size_t pmr_peer_exchange_bytes(long long n_floats, int world);
PMR_API int pmr_peer_alloc(pmr_context *ctx, size_t bytes, void **ptr, void *handle);
PMR_API int pmr_peer_open(pmr_context *ctx, const void *handle, void **ptr);
PMR_API int pmr_peer_close(pmr_context *ctx, void *ptr);
PMR_API int pmr_peer_free(pmr_context *ctx, void *ptr);
PMR_API int pmr_peer_status(pmr_context *ctx, const void *own_buffer, int *status);
PMR_API int pmr_transform_backward_exchange(pmr_context *ctx, const float *matrices, const float *d_clip_vertices,
                                            int B, int V, void *const *peer_buffers, int rank, int world,
                                            long long epoch, float *d_world_vertices, void *stream);

/*
 * Vertex normals of a triangle mesh (reference src/common/meshes.py:3-35 compute_vertex_normals; SURVEY.md
 * section 8f row 4) for loops that move the geometry every step.
 *
 * pmr_vertex_incidence turns the topology, ONCE, into the table both directions gather through:
 * offsets int32 [V+1] and incidence int32 [3T], row v = the (corner, triangle) pairs with
 * triangles[triangle][corner] == v, coded corner << 30 | triangle and sorted ascending -- the order in which
 * the reference's three index_add_ passes (meshes.py:23-33) reach the vertex.  T < 2^30; vertex ids outside
 * [0, V) are skipped (the reference raises an index error).  The rows are ordered by counting inside each row, so the
 * build costs the SUM OF SQUARED VALENCES in comparisons (the pole of a UV sphere with 700 longitudes: nothing; one
 * vertex shared by 10^6 triangles: minutes) -- a one-time cost per topology, sized for meshes, not for star graphs.
 *
 * pmr_vertex_normals_forward: vertices float32 [B,V,3] -> normals float32 [B,V,3] (normalised with
 * eps = 1e-6, meshes.py:34) and, when raw != NULL, the summed un-normalised normals [B,V,3] that the backward
 * needs.  Sums run in the reference's order with torch's CPU arithmetic, so the result is bit-reproducible.
 * pmr_vertex_normals_backward: grad_normals [B,V,3] -> d_vertices [B,V,3]; grad_raw [B,V,3] is caller-provided
 * scratch (the gradient at the un-normalised normals).  No atomics in either direction.
 */
PMR_API int pmr_vertex_incidence(pmr_context *ctx, const int32_t *triangles, int T, int V, int32_t *offsets,
                                 int32_t *incidence, void *stream);
PMR_API int pmr_vertex_normals_forward(pmr_context *ctx, const float *vertices, const int32_t *triangles,
                                       const int32_t *offsets, const int32_t *incidence, int B, int V, int T,
                                       float *raw, float *normals, void *stream);
PMR_API int pmr_vertex_normals_backward(pmr_context *ctx, const float *grad_normals, const float *raw,
                                        const float *vertices, const int32_t *triangles, const int32_t *offsets,
                                        const int32_t *incidence, int B, int V, int T, float *grad_raw,
                                        float *d_vertices, void *stream);

/*
 * Direct caller of the path: per-pixel Phong lighting, diffuse + ambient terms, of the interpolated
 * attribute image (reference src/mesh_renderer/render.py:201-228 + phong_shader :231-386 for a
 * `render` call without specular colours).  pixels float32 [B,H,W,A], A >= 9, channels
 * [normal xyz, world position xyz, diffuse rgb] (render.py:181); light_positions / light_intensities
 * float32 [B,L,3], L <= 16; ambient float32 [B,3] or NULL.  rgba float32 [B,H,W,4] (16-byte aligned),
 * rows flipped (row 0 = top of the image, render.py:382-386), alpha = 1 where the diffuse colour is not
 * the -1 background.  The backward maps grad_rgba to d_pixels [B,H,W,A] (channels beyond 8 get zero);
 * gradients with respect to the lights are not produced.
 */
PMR_API int pmr_shade_diffuse_forward(pmr_context *ctx, const float *pixels, const float *light_positions,
                                      const float *light_intensities, const float *ambient,
                                      int B, int L, int A, int image_width, int image_height,
                                      float *rgba, void *stream);
PMR_API int pmr_shade_diffuse_backward(pmr_context *ctx, const float *grad_rgba, const float *pixels,
                                       const float *light_positions, const float *light_intensities,
                                       const float *ambient, int B, int L, int A, int image_width,
                                       int image_height, float *d_pixels, void *stream);

/*
 * The same with the specular term (render.py:326-372): pixels float32 [B,H,W,A], A = 12: [normal, position,
 * diffuse, specular rgb] with `shininess` float32 [B] (one exponent per image), or A = 13 with the exponent in
 * channel 12 (`shininess` ignored, may be NULL); camera_position float32 [B,3].  The reference divides the
 * reflection . view products of each (image, light) by their L2 norm over all pixels of the image before the
 * power: the forward pass returns those sums of squares in norm2 [B,L] (written by the call) and the backward
 * pass takes them back, together with a scratch buffer sum_gx [B,L] it overwrites.  Gradients reach the
 * pixel channels only (all 12 / 13 of them).
 */
PMR_API int pmr_shade_phong_forward(pmr_context *ctx, const float *pixels, const float *light_positions,
                                    const float *light_intensities, const float *ambient,
                                    const float *camera_position, const float *shininess,
                                    int B, int L, int A, int image_width, int image_height,
                                    float *norm2, float *rgba, void *stream);
PMR_API int pmr_shade_phong_backward(pmr_context *ctx, const float *grad_rgba, const float *pixels,
                                     const float *light_positions, const float *light_intensities,
                                     const float *ambient, const float *camera_position, const float *shininess,
                                     const float *norm2, int B, int L, int A, int image_width, int image_height,
                                     float *sum_gx, float *d_pixels, void *stream);

/*
 * The render path in one piece (render.py:183-228 for a call without specular colours): rasterize, interpolate
 * the nine channels [normal, world position, diffuse colour] and light them, without ever storing the
 * [B,H,W,9] attribute image: the forward writes ids / barycentrics / z and RGBA [B,H,W,4] (rows flipped), the
 * backward goes from grad_rgba straight to d_vertices [B,V,4] and d_attributes [B,V,9] (atomic accumulation).
 * attributes float32 [B,V,9]; background float32 [9] (render.py:197 passes -1); lights as in
 * pmr_shade_diffuse_forward.  Results equal pmr_rasterize_interpolate_forward followed by
 * pmr_shade_diffuse_forward bit for bit.
 */
PMR_API int pmr_render_diffuse_forward(pmr_context *ctx, const float *vertices, const float *attributes,
                                       const int32_t *triangles, const float *background,
                                       const float *light_positions, const float *light_intensities,
                                       const float *ambient, int B, int V, int T, int L,
                                       int image_width, int image_height,
                                       int32_t *ids, float *bary, float *z, float *rgba, void *stream);
PMR_API int pmr_render_diffuse_backward(pmr_context *ctx, const float *grad_rgba, const float *vertices,
                                        const float *attributes, const int32_t *triangles, const float *background,
                                        const float *light_positions, const float *light_intensities,
                                        const float *ambient, const int32_t *ids, const float *bary,
                                        int B, int V, int T, int L, int image_width, int image_height,
                                        float *d_vertices, float *d_attributes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PMR_B200_H_ */
