// raster_forward.cu -- screen-tile binning and the per-tile raster kernel (forward pass of
// rasterize_triangles, K.cpp:302-419, batched over images) with the attribute interpolation
// of rasterize_clip_space (rast.py:118-150) fused into its epilogue.
//
// Pipeline per call (all on one stream):
//   bin_count_kernel   one thread per (image, triangle): pixel box -> tile range, per-tile counts
//   bin_offsets_kernel warp-aggregated allocation of one contiguous list range per tile
//   bin_fill_kernel    writes triangle ids into the tile lists (order inside a list is arbitrary)
//   raster_tile_kernel one CTA per 16x16-pixel tile: stages triangle setup records into shared
//                      memory, every warp culls them against its own 8x4 pixel block with a
//                      ballot, each lane owns one pixel and keeps its depth-test winner in
//                      registers (no atomics), then writes ids / barycentrics / z (+ attributes).
// Meshes with few triangles skip binning: every tile walks the whole triangle array.
//
// The depth rule is order independent (min z, then max id -- SURVEY.md F1), so list order
// does not matter and the result is deterministic.
#include "pmr_internal.cuh"
#include "raster_math.cuh"

namespace pmr {

// ---------------------------------------------------------------------------------------------
// Binning
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void load_triangle(const float *__restrict__ verts_b,
                                              const int32_t *__restrict__ tris, int t,
                                              float4 &a, float4 &b, float4 &c) {
  const int i0 = __ldg(tris + 3 * (size_t)t + 0);
  const int i1 = __ldg(tris + 3 * (size_t)t + 1);
  const int i2 = __ldg(tris + 3 * (size_t)t + 2);
  const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
  a = __ldg(v4 + i0);
  b = __ldg(v4 + i1);
  c = __ldg(v4 + i2);
}

// Tile range of a pixel box, packed as four uint16 (tx0, tx1, ty0, ty1; exclusive upper ends).
__device__ __forceinline__ uint2 tile_range_of(const PixelBox &box) {
  uint2 r;
  if (box.left >= box.right || box.bottom >= box.top) { r.x = 0u; r.y = 0u; return r; }
  const unsigned tx0 = (unsigned)box.left >> kTileShiftX, tx1 = ((unsigned)box.right + kTileW - 1) >> kTileShiftX;
  const unsigned ty0 = (unsigned)box.bottom >> kTileShiftY, ty1 = ((unsigned)box.top + kTileH - 1) >> kTileShiftY;
  r.x = tx0 | (tx1 << 16);
  r.y = ty0 | (ty1 << 16);
  return r;
}

// Visits every tile of the ranges held by the lanes of a warp.  Ranges of up to
// kSerialTiles tiles are walked by their own lane; larger ones are walked by the whole warp
// so that one screen-filling triangle does not serialise 16k atomics on a single thread.
template <typename Visit>
__device__ __forceinline__ void for_each_tile(uint2 range, int tiles_x, Visit visit) {
  const int tx0 = range.x & 0xffff, tx1 = range.x >> 16;
  const int ty0 = range.y & 0xffff, ty1 = range.y >> 16;
  const int nx = tx1 - tx0, n = nx * (ty1 - ty0);
  constexpr int kSerialTiles = 8;
  if (n > 0 && n <= kSerialTiles) {
    for (int ty = ty0; ty < ty1; ++ty)
      for (int tx = tx0; tx < tx1; ++tx) visit(ty * tiles_x + tx, /*owner_lane=*/-1);
  }
  unsigned big = __ballot_sync(0xffffffffu, n > kSerialTiles);
  const int lane = threadIdx.x & 31;
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int sx0 = __shfl_sync(0xffffffffu, tx0, src), snx = __shfl_sync(0xffffffffu, nx, src);
    const int sy0 = __shfl_sync(0xffffffffu, ty0, src), sn = __shfl_sync(0xffffffffu, n, src);
    for (int k = lane; k < sn; k += 32) {
      const int ty = sy0 + k / snx, tx = sx0 + k % snx;
      visit(ty * tiles_x + tx, src);
    }
  }
}

__global__ void __launch_bounds__(256)
bin_count_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris,
                 int V, int T, int W, int H, float half_w, float half_h, int tiles_x, int tiles_per_image,
                 uint2 *__restrict__ tri_ranges, int *__restrict__ tile_counts) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  uint2 range = make_uint2(0u, 0u);
  if (t < T) {
    float4 p0, p1, p2;
    load_triangle(verts + (size_t)b * V * 4, tris, t, p0, p1, p2);
    range = tile_range_of(triangle_box(p0, p1, p2, half_w, half_h, W, H));
    tri_ranges[(size_t)b * T + t] = range;
  }
  int *counts = tile_counts + (size_t)b * tiles_per_image;
  for_each_tile(range, tiles_x, [&](int tile, int) { atomicAdd(counts + tile, 1); });
}

// One contiguous range per tile; ranges are handed out warp by warp from a global cursor, so
// their order in the list buffer is arbitrary (nothing depends on it).
__global__ void __launch_bounds__(256)
bin_offsets_kernel(const int *__restrict__ tile_counts, int n_tiles, int *__restrict__ tile_offsets,
                   unsigned long long *__restrict__ total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int c = i < n_tiles ? tile_counts[i] : 0;
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (lane == 31 && warp_total > 0) base = atomicAdd(total, (unsigned long long)warp_total);
  base = __shfl_sync(0xffffffffu, base, 31);
  // Offsets beyond 2^31 entries are refused on the host before the fill kernel runs.
  if (i < n_tiles) tile_offsets[i] = (int)(base + (unsigned long long)(incl - c));
}

__global__ void __launch_bounds__(256)
bin_fill_kernel(const uint2 *__restrict__ tri_ranges, int T, int tiles_x, int tiles_per_image,
                const int *__restrict__ tile_offsets, int *__restrict__ tile_cursors,
                int32_t *__restrict__ tile_lists) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint2 range = t < T ? tri_ranges[(size_t)b * T + t] : make_uint2(0u, 0u);
  const int *offsets = tile_offsets + (size_t)b * tiles_per_image;
  int *cursors = tile_cursors + (size_t)b * tiles_per_image;
  const int lane = threadIdx.x & 31;
  for_each_tile(range, tiles_x, [&](int tile, int owner) {
    const int tri = owner < 0 ? t : (t - lane + owner);
    const int slot = atomicAdd(cursors + tile, 1);
    tile_lists[(size_t)offsets[tile] + slot] = tri;
  });
}

// ---------------------------------------------------------------------------------------------
// Per-tile raster kernel
// ---------------------------------------------------------------------------------------------

constexpr int kChunk = 256;   // triangles staged per round == threads per CTA

struct TileSmem {
  // Setup record of one staged triangle, split into float4 planes so that staging stores are
  // conflict free and the warp-uniform reads in the pixel loop are broadcasts.
  float4 r0[kChunk];   // m0 m1 m2 | id
  float4 r1[kChunk];   // m3 m4 m5 | z0
  float4 r2[kChunk];   // m6 m7 m8 | z1
  float4 r3[kChunk];   // z2 | w0 w1 w2
  int4 box[kChunk];    // left right bottom top (pixels)
};

template <int A_STATIC>
__global__ void __launch_bounds__(kChunk)
raster_tile_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris,
                   int V, int T, int W, int H, float half_w, float half_h, int tiles_x, int tiles_per_image,
                   const int *__restrict__ tile_counts, const int *__restrict__ tile_offsets,
                   const int32_t *__restrict__ tile_lists,
                   int32_t *__restrict__ out_ids, float *__restrict__ out_bary, float *__restrict__ out_z,
                   const float *__restrict__ attrs, const float *__restrict__ background, int A_dyn,
                   float *__restrict__ out_image) {
  __shared__ TileSmem sm;
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  const int b = blockIdx.y;
  const int tile = blockIdx.x;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 8 warps, each an 8x4 pixel block; blocks are laid out 2 across, 4 down inside the 16x16 tile.
  const int blk_x0 = tile_x * kTileW + (warp & 1) * 8;
  const int blk_y0 = tile_y * kTileH + (warp >> 1) * 4;
  const int ix = blk_x0 + (lane & 7);
  const int iy = blk_y0 + (lane >> 3);
  const float px = pixel_center(ix, half_w);
  const float py = pixel_center(iy, half_h);
  const float *verts_b = verts + (size_t)b * V * 4;

  int n_list;
  const int32_t *list = nullptr;
  if (tile_lists != nullptr) {
    const size_t g = (size_t)b * tiles_per_image + tile;
    n_list = tile_counts[g];
    list = tile_lists + tile_offsets[g];
  } else {
    n_list = T;            // small mesh: every tile walks all triangles
  }

  Fragment best;
  fragment_clear(best);

  for (int base = 0; base < n_list; base += kChunk) {
    const int n_here = min(kChunk, n_list - base);
    __syncthreads();       // previous chunk fully consumed
    if (threadIdx.x < n_here) {
      const int t = list ? list[base + threadIdx.x] : base + threadIdx.x;
      float4 p0, p1, p2;
      load_triangle(verts_b, tris, t, p0, p1, p2);
      const PixelBox bx = triangle_box(p0, p1, p2, half_w, half_h, W, H);
      float m[9];
      adjugate_signed(p0.x, p1.x, p2.x, p0.y, p1.y, p2.y, p0.w, p1.w, p2.w, m);
      sm.r0[threadIdx.x] = make_float4(m[0], m[1], m[2], __int_as_float(t));
      sm.r1[threadIdx.x] = make_float4(m[3], m[4], m[5], p0.z);
      sm.r2[threadIdx.x] = make_float4(m[6], m[7], m[8], p1.z);
      sm.r3[threadIdx.x] = make_float4(p2.z, p0.w, p1.w, p2.w);
      sm.box[threadIdx.x] = make_int4(bx.left, bx.right, bx.bottom, bx.top);
    }
    __syncthreads();

    for (int g0 = 0; g0 < n_here; g0 += 32) {
      // Coarse: lane j holds triangle g0+j; does its pixel box touch this warp's 8x4 block?
      bool touches = false;
      if (g0 + lane < n_here) {
        const int4 bx = sm.box[g0 + lane];
        touches = bx.x < blk_x0 + 8 && bx.y > blk_x0 && bx.z < blk_y0 + 4 && bx.w > blk_y0;
      }
      unsigned todo = __ballot_sync(0xffffffffu, touches);
      while (todo) {
        const int j = g0 + __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 bx = sm.box[j];
        // The reference only visits pixels inside the triangle's own box (K.cpp:374-375).
        if (ix >= bx.x && ix < bx.y && iy >= bx.z && iy < bx.w) {
          const float4 q0 = sm.r0[j], q1 = sm.r1[j], q2 = sm.r2[j], q3 = sm.r3[j];
          const float m[9] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
          const float zc[3] = {q1.w, q2.w, q3.x};
          const float wc[3] = {q3.y, q3.z, q3.w};
          fragment_test(m, zc, wc, px, py, __float_as_int(q0.w), best);
        }
      }
    }
  }

  if (ix >= W || iy >= H) return;
  const size_t p = ((size_t)b * H + iy) * W + ix;
  const bool covered = best.id >= 0;
  const int id = covered ? best.id : 0;
  out_ids[p] = id;
  out_z[p] = best.z;
  out_bary[3 * p + 0] = best.b0;
  out_bary[3 * p + 1] = best.b1;
  out_bary[3 * p + 2] = best.b2;

  if (out_image != nullptr) {
    // rast.py:118-150: corner attributes weighted by barycentrics, alpha, background blend.
    float *o = out_image + p * A;
    if (!covered) {
      for (int a = 0; a < A; ++a) o[a] = __ldg(background + a);
    } else {
      const float *at = attrs + (size_t)b * V * A;
      const float *c0 = at + (size_t)__ldg(tris + 3 * (size_t)id + 0) * A;
      const float *c1 = at + (size_t)__ldg(tris + 3 * (size_t)id + 1) * A;
      const float *c2 = at + (size_t)__ldg(tris + 3 * (size_t)id + 2) * A;
      const float alpha = coverage_alpha(best.b0, best.b1, best.b2);
      const float one_minus = 1.0f - alpha;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const float img = __ldg(c0 + a) * best.b0 + __ldg(c1 + a) * best.b1 + __ldg(c2 + a) * best.b2;
        o[a] = alpha * img + one_minus * __ldg(background + a);
      }
    }
  }
}

// Standalone interpolation (rast.py:118-150) from existing id / barycentric buffers.
__global__ void __launch_bounds__(256)
interpolate_kernel(const float *__restrict__ attrs, const int32_t *__restrict__ tris,
                   const int32_t *__restrict__ ids, const float *__restrict__ bary,
                   const float *__restrict__ background, int V, int A, long long pixels_per_image,
                   long long total_pixels, float *__restrict__ out) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_pixels) return;
  const int b = (int)(p / pixels_per_image);
  const float b0 = bary[3 * p], b1 = bary[3 * p + 1], b2 = bary[3 * p + 2];
  const int id = ids[p];
  const float *at = attrs + (size_t)b * V * A;
  const float *c0 = at + (size_t)__ldg(tris + 3 * (size_t)id + 0) * A;
  const float *c1 = at + (size_t)__ldg(tris + 3 * (size_t)id + 1) * A;
  const float *c2 = at + (size_t)__ldg(tris + 3 * (size_t)id + 2) * A;
  const float alpha = coverage_alpha(b0, b1, b2);
  const float one_minus = 1.0f - alpha;
  float *o = out + p * A;
  for (int a = 0; a < A; ++a) {
    const float img = __ldg(c0 + a) * b0 + __ldg(c1 + a) * b1 + __ldg(c2 + a) * b2;
    o[a] = alpha * img + one_minus * __ldg(background + a);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------

static int launch_raster(Context *ctx, const float *verts, const int32_t *tris, int B, int V, int T,
                         int W, int H, const int *counts, const int *offsets, const int32_t *lists,
                         int32_t *ids, float *bary, float *z, const float *attrs, const float *bg, int A,
                         float *image, cudaStream_t stream) {
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const int tiles = tiles_x * tiles_y;
  const float half_w = (float)(0.5 * W), half_h = (float)(0.5 * H);   // K.cpp:309-310
  dim3 grid(tiles, B);
  StageScope timed(ctx, PMR_STAGE_RASTER, stream);
#define PMR_LAUNCH(AS)                                                                              \
  raster_tile_kernel<AS><<<grid, kChunk, 0, stream>>>(verts, tris, V, T, W, H, half_w, half_h,      \
                                                     tiles_x, tiles, counts, offsets, lists, ids,   \
                                                     bary, z, attrs, bg, A, image)
  if (image == nullptr) PMR_LAUNCH(0);
  else if (A == 4) PMR_LAUNCH(4);
  else if (A == 9) PMR_LAUNCH(9);
  else if (A == 12) PMR_LAUNCH(12);
  else if (A == 13) PMR_LAUNCH(13);
  else PMR_LAUNCH(0);
#undef PMR_LAUNCH
  ctx->launches += 1;
  return check_launch(ctx, "raster_tile_kernel");
}

int forward_impl(Context *ctx, const float *verts, const int32_t *tris, int B, int V, int T, int W, int H,
                 int32_t *ids, float *bary, float *z, const float *attrs, const float *bg, int A,
                 float *image, cudaStream_t stream) {
  if (B == 0 || W == 0 || H == 0) return PMR_OK;
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const int tiles = tiles_x * tiles_y;
  const float half_w = (float)(0.5 * W), half_h = (float)(0.5 * H);

  if (T <= ctx->small_mesh_threshold) {
    return launch_raster(ctx, verts, tris, B, V, T, W, H, nullptr, nullptr, nullptr, ids, bary, z, attrs,
                         bg, A, image, stream);
  }

  const size_t n_tiles = (size_t)B * tiles;
  int32_t *lists = nullptr;
  int *counts = nullptr, *offsets = nullptr;
  {
  StageScope timed(ctx, PMR_STAGE_BIN, stream);
  if (n_tiles > (size_t)INT_MAX) return set_error(ctx, PMR_ERR_SIZE, "too many screen tiles");
  // workspace: [total u64 | counts | cursors | offsets | ranges]
  int rc = ctx->bins.reserve(ctx, 16 + n_tiles * 3 * sizeof(int) + (size_t)B * T * sizeof(uint2) + 64);
  if (rc) return rc;
  char *base = (char *)ctx->bins.ptr;
  unsigned long long *total = (unsigned long long *)base;
  counts = (int *)(base + 16);
  int *cursors = counts + n_tiles;
  offsets = cursors + n_tiles;
  uint2 *ranges = (uint2 *)(((uintptr_t)(offsets + n_tiles) + 15) & ~(uintptr_t)15);

  PMR_CUDA(ctx, cudaMemsetAsync(base, 0, 16 + n_tiles * 2 * sizeof(int), stream));
  dim3 tgrid((T + 255) / 256, B);
  bin_count_kernel<<<tgrid, 256, 0, stream>>>(verts, tris, V, T, W, H, half_w, half_h, tiles_x, tiles,
                                             ranges, counts);
  bin_offsets_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, stream>>>(counts, (int)n_tiles, offsets,
                                                                           total);
  ctx->launches += 2;
  rc = check_launch(ctx, "bin_count/bin_offsets");
  if (rc) return rc;

  // The list length is data dependent: read it back (8 bytes, pinned) to size the buffer.
  PMR_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                stream));
  PMR_CUDA(ctx, cudaStreamSynchronize(stream));
  const unsigned long long n_entries = *ctx->mailbox;
  ctx->last_bin_entries = n_entries;
  if (n_entries >= (1ull << 31)) return set_error(ctx, PMR_ERR_SIZE, "tile lists exceed 2^31 entries");
  rc = ctx->lists.reserve(ctx, (size_t)(n_entries + 1) * sizeof(int32_t));
  if (rc) return rc;
  lists = (int32_t *)ctx->lists.ptr;

  bin_fill_kernel<<<tgrid, 256, 0, stream>>>(ranges, T, tiles_x, tiles, offsets, cursors, lists);
  ctx->launches += 1;
  rc = check_launch(ctx, "bin_fill_kernel");
  if (rc) return rc;
  }
  return launch_raster(ctx, verts, tris, B, V, T, W, H, counts, offsets, lists, ids, bary, z, attrs, bg, A,
                       image, stream);
}

int interpolate_impl(Context *ctx, const float *attrs, const int32_t *tris, const int32_t *ids,
                     const float *bary, const float *bg, int B, int V, int A, int W, int H, float *out,
                     cudaStream_t stream) {
  const long long ppi = (long long)W * H, total = ppi * B;
  if (total == 0 || A == 0) return PMR_OK;
  StageScope timed(ctx, PMR_STAGE_INTERP, stream);
  interpolate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(attrs, tris, ids, bary, bg, V, A,
                                                                         ppi, total, out);
  ctx->launches += 1;
  return check_launch(ctx, "interpolate_kernel");
}

}  // namespace pmr
