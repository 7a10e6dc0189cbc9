// raster_forward.cu -- forward pass of rasterize_triangles (K.cpp:302-419, batched over images) with
// the attribute interpolation of rasterize_clip_space (rast.py:118-150) fused into its last kernel.
//
// Depth is resolved with ONE packed 64-bit key per pixel, (orderable depth bits << 32) | ~id, whose
// minimum is the reference's winner (min z, then max id -- SURVEY.md F1).  Pipeline (one stream):
//   scatter_small_kernel  one warp per 32 triangles of an image.  Triangles whose pixel box is at most
//                         16x16 are rasterized right here: their boxes are cut into row segments of
//                         <= 4 pixels, the segments of the warp's triangles are laid out back to back
//                         in shared memory and dealt to the lanes (balanced whatever the sizes), the
//                         inside pixels are compacted and dealt to the lanes again for barycentrics /
//                         depth, and each one does atomicMin on the pixel's key in global memory (L2).
//                         No per-tile duplication of triangle setup, no binning for these triangles.
//                         Larger triangles are only counted into the 16x16 screen tiles they touch.
//   bin_offsets_kernel    warp-aggregated allocation of one contiguous list range per tile
//   bin_fill_kernel       writes the ids of the LARGE triangles into the tile lists
//   raster_tile_kernel    one CTA per tile for the large triangles: setup records staged in shared
//                         memory, a warp per 8x4 pixel block culls them with a ballot, one lane per
//                         pixel keeps its winner in registers (parts of large triangles that are small
//                         inside the tile take the segment path on a shared-memory key), result merged
//                         into the global keys.  Skipped when no triangle is large.
//   resolve_kernel        one warp per 8x4 pixel block: decodes the winner, re-evaluates its
//                         barycentrics / depth once (same arithmetic, same bits), interpolates the
//                         attributes and stores ids / z / barycentrics / image as 16-byte vectors
//                         through a shared-memory transpose.
// Meshes with few triangles (<= small_mesh_threshold) run raster_tile_kernel alone: every tile walks
// the whole triangle array and writes the outputs itself.
//
// Neither list order nor atomic order can change a result, so the pass is deterministic.
#include "pmr_internal.cuh"
#include "raster_math.cuh"

namespace pmr {

// ---------------------------------------------------------------------------------------------
// Binning
// ---------------------------------------------------------------------------------------------

// Pixel box packed as four uint16: x = left | right << 16, y = bottom | top << 16 (W, H <= 32768).
__device__ __forceinline__ uint2 pack_box(const PixelBox &box) {
  if (box.left >= box.right || box.bottom >= box.top) return make_uint2(0u, 0u);
  return make_uint2((unsigned)box.left | ((unsigned)box.right << 16),
                    (unsigned)box.bottom | ((unsigned)box.top << 16));
}

__device__ __forceinline__ int4 unpack_box(uint2 p) {
  return make_int4((int)(p.x & 0xffffu), (int)(p.x >> 16), (int)(p.y & 0xffffu), (int)(p.y >> 16));
}

// Visits every tile touched by the boxes held by the lanes of a warp.  Ranges of up to
// kSerialTiles tiles are walked by their own lane; larger ones are walked by the whole warp
// so that one screen-filling triangle does not serialise 16k atomics on a single thread.
template <typename Visit>
__device__ __forceinline__ void for_each_tile(uint2 packed, int tiles_x, Visit visit) {
  const int4 box = unpack_box(packed);
  const bool empty = box.x >= box.y || box.z >= box.w;
  const int tx0 = box.x >> kTileShiftX, tx1 = empty ? tx0 : (box.y + kTileW - 1) >> kTileShiftX;
  const int ty0 = box.z >> kTileShiftY, ty1 = empty ? ty0 : (box.w + kTileH - 1) >> kTileShiftY;
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
bin_fill_kernel(const uint2 *__restrict__ tri_boxes, int T, int tiles_x, int tiles_per_image,
                const int *__restrict__ tile_offsets, int *__restrict__ tile_cursors,
                int32_t *__restrict__ tile_lists) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint2 packed = t < T ? tri_boxes[(size_t)b * T + t] : make_uint2(0u, 0u);
  const int *offsets = tile_offsets + (size_t)b * tiles_per_image;
  int *cursors = tile_cursors + (size_t)b * tiles_per_image;
  const int lane = threadIdx.x & 31;
  for_each_tile(packed, tiles_x, [&](int tile, int owner) {
    const int tri = owner < 0 ? t : (t - lane + owner);
    const int slot = atomicAdd(cursors + tile, 1);
    tile_lists[(size_t)offsets[tile] + slot] = tri;
  });
}

// ---------------------------------------------------------------------------------------------
// Per-tile raster kernel
// ---------------------------------------------------------------------------------------------

constexpr int kChunk = 256;                 // triangles staged per round == threads per CTA
constexpr int kTilePixels = kTileW * kTileH;
constexpr int kWarps = kChunk / 32;
constexpr int kWarpSegCap = 256;            // row segments of small triangles a warp holds per round

struct TileSmem {
  // Setup record of one staged triangle, split into float4 planes so that staging stores are
  // conflict free and warp-uniform reads are broadcasts.
  float4 r0[kChunk];   // m0 m1 m2 | id
  float4 r1[kChunk];   // m3 m4 m5 | z0
  float4 r2[kChunk];   // m6 m7 m8 | z1
  float4 r3[kChunk];   // z2 | w0 w1 w2
  int4 box[kChunk];    // the triangle's pixel box: left right bottom top
  float zlo[kChunk];   // conservative lower bound of the triangle's depth (-inf unless all w > 0)
  unsigned long long key[kTilePixels];   // packed (depth, ~id) minimum per pixel (small-triangle path)
  unsigned segs[kWarps][kWarpSegCap];    // per warp: slot | row << 8 | first column << 12 | width << 16
  unsigned short hits[kWarps][128];      // per warp: inside pixels of 32 segments (slot << 8 | pixel)
  unsigned short big_list[kChunk];
  int depth_bucket[32], depth_cursor[32];   // front-to-back ordering of the big list (32 depth slices)
  float cx[kTileW], cy[kTileH];          // pixel-centre NDC coordinates of the tile's columns / rows
  int n_big;
};

// After the raster loop the record planes are dead; the epilogue reuses them to transpose each warp's
// 8x4 pixel block into row-contiguous runs so that global stores are full 16-byte vectors.
static_assert(sizeof(float4) * kChunk * 4 >= sizeof(float) * kWarps * 32 * 16, "epilogue staging must fit");

// Appends the slots of the threads with `flag` set to `list` (order irrelevant).
__device__ __forceinline__ void append_slots(bool flag, unsigned short *list, int *count) {
  const unsigned votes = __ballot_sync(0xffffffffu, flag);
  if (votes == 0u) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(count, __popc(votes));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (flag) list[base + __popc(votes & ((1u << lane) - 1u))] = (unsigned short)threadIdx.x;
}

// Writes N floats per pixel of the warp's 8x4 block from the shared staging area (`stage`: 4 rows x
// 8*N contiguous floats) to `dst_row0` (+ row * row_stride floats).  The vector path needs 16-byte
// aligned rows and a full-width block; otherwise scalars.  N is a compile-time constant so that the
// row/column split is a multiply-shift, not a division.
template <int N>
__device__ __forceinline__ void store_block_rows(const float *stage, float *dst_row0, int row_stride,
                                                 int cols, int rows, bool vec_ok) {
  const int lane = threadIdx.x & 31;
  constexpr int run = 8 * N;                           // floats per full block row
  if (vec_ok && cols == 8) {
    constexpr int v4_per_row = 2 * N, total = 4 * v4_per_row;
    const float4 *stage4 = reinterpret_cast<const float4 *>(stage);     // rows are back to back
#pragma unroll
    for (int i = 0; i < (total + 31) / 32; ++i) {
      const int k = lane + 32 * i;
      const int r = k / v4_per_row, c = k - r * v4_per_row;
      if (k < total && r < rows)
        *reinterpret_cast<float4 *>(dst_row0 + (r * row_stride + 4 * c)) = stage4[k];
    }
  } else {
    const int live = cols * N;
    for (int k = lane; k < rows * run; k += 32) {
      const int r = k / run, c = k - r * run;
      if (c < live) dst_row0[r * row_stride + c] = stage[k];
    }
  }
}

// Barycentrics / depth of a known winner (triangle t covers the pixel and passed the depth range).
__device__ __forceinline__ void evaluate_winner(const float4 &p0, const float4 &p1, const float4 &p2,
                                                float px, float py, int t, Fragment &out) {
  float m[9], e[3], esum, bc[3], z;
  adjugate_signed(p0.x, p1.x, p2.x, p0.y, p1.y, p2.y, p0.w, p1.w, p2.w, m);
  edge_values(m, px, py, e);
  edges_inside(e, esum);
  const float zc[3] = {p0.z, p1.z, p2.z}, wc[3] = {p0.w, p1.w, p2.w};
  fragment_depth(e, esum, zc, wc, bc, z);
  out.z = z; out.id = t; out.b0 = bc[0]; out.b1 = bc[1]; out.b2 = bc[2];
}

// One row of a block (bytes contiguous in shared and in global memory, both 16-byte aligned, size a
// multiple of 16) handed to the bulk-copy engine: cp.async.bulk shared::cta -> global (SASS UBLKCP).
// The issuing lane commits the group and must wait for the reads (bulk_store_wait) before the
// shared memory is reused or the CTA exits.
__device__ __forceinline__ void bulk_store_row(float *dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Output of one warp's 8x4 pixel block (lane = pixel): ids / z / barycentrics and, when requested, the
// interpolated attributes of rast.py:118-150.  ids and z rows are full 32-byte sectors as they are.
// Barycentrics (12 B/pixel) and attributes (4A B/pixel) are transposed through `stage` (32*16 floats of
// shared memory owned by the warp) into row-contiguous runs; full, aligned blocks are then written by
// the bulk-copy engine (one elected lane issues one cp.async.bulk per block row: 96 B and 32*A B), which
// takes the copy loops off the instruction-issue-bound SMs; edge / unaligned blocks use vector or
// scalar stores.
template <int A_STATIC>
__device__ __forceinline__ void block_epilogue(float *stage, int b, int blk_x0, int blk_y0, int W, int H, int V,
                                               const Fragment &best, const int32_t *__restrict__ tris,
                                               const float *__restrict__ attrs, const float *__restrict__ background,
                                               int A, int32_t *__restrict__ out_ids, float *__restrict__ out_bary,
                                               float *__restrict__ out_z, float *__restrict__ out_image,
                                               const float *corners = nullptr) {
  // `corners`: the winner's 3*A corner attributes already in registers ([corner][attribute]), or
  // nullptr to gather them here.
  const int lane = threadIdx.x & 31;
  const bool covered = best.id >= 0;
  const int id = covered ? best.id : 0;
  const int cols = min(8, W - blk_x0), rows = min(4, H - blk_y0);     // <= 0: block outside the image
  if (cols <= 0 || rows <= 0) return;
  const bool vec_ok = (W & 3) == 0 &&
      (((uintptr_t)out_ids | (uintptr_t)out_z | (uintptr_t)out_bary | (uintptr_t)out_image) & 15) == 0;
  const size_t p0 = ((size_t)b * H + blk_y0) * W + blk_x0;            // first pixel of the block
  const bool in_image = (lane & 7) < cols && (lane >> 3) < rows;
  if (in_image) {
    const size_t p = p0 + (size_t)(lane >> 3) * W + (lane & 7);
    out_ids[p] = id;
    out_z[p] = best.z;
  }
  // staging layout: barycentrics in floats [0, 96), attributes in [96, 96 + 32*A)
  constexpr int kImageAt = 96;
  constexpr bool staged = A_STATIC > 0 && kImageAt + 32 * A_STATIC <= 32 * 16;
  const bool bulk = vec_ok && cols == 8 && rows == 4 && (out_image == nullptr || staged);
  stage[3 * lane + 0] = best.b0;
  stage[3 * lane + 1] = best.b1;
  stage[3 * lane + 2] = best.b2;

  if (out_image != nullptr) {
    // rast.py:118-150: corner attributes weighted by barycentrics, alpha, background blend.
    float *dst = staged ? stage + kImageAt + lane * A : out_image + (p0 + (size_t)(lane >> 3) * W + (lane & 7)) * A;
    if (staged || in_image) {
      if (!covered) {
        for (int a = 0; a < A; ++a) dst[a] = __ldg(background + a);
      } else {
        const float alpha = coverage_alpha(best.b0, best.b1, best.b2);
        const float one_minus = 1.0f - alpha;
        if (corners != nullptr) {
#pragma unroll
          for (int a = 0; a < A; ++a) {
            const float img = corners[a] * best.b0 + corners[A + a] * best.b1 + corners[2 * A + a] * best.b2;
            dst[a] = alpha * img + one_minus * __ldg(background + a);
          }
        } else {
          const float *at = attrs + (size_t)b * V * A;
          const float *c0 = at + (size_t)__ldg(tris + 3 * (size_t)id + 0) * A;
          const float *c1 = at + (size_t)__ldg(tris + 3 * (size_t)id + 1) * A;
          const float *c2 = at + (size_t)__ldg(tris + 3 * (size_t)id + 2) * A;
#pragma unroll
          for (int a = 0; a < A; ++a) {
            const float img = __ldg(c0 + a) * best.b0 + __ldg(c1 + a) * best.b1 + __ldg(c2 + a) * best.b2;
            dst[a] = alpha * img + one_minus * __ldg(background + a);
          }
        }
      }
    }
  }
  __syncwarp();
  if (bulk) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy reads
    __syncwarp();
    if (lane == 0) {
      float *bary_row = out_bary + 3 * p0, *image_row = out_image != nullptr ? out_image + p0 * A : nullptr;
      const unsigned stage_at = (unsigned)__cvta_generic_to_shared(stage);
      const int bary_pitch = 3 * W, image_pitch = W * A;                // floats per image row
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        bulk_store_row(bary_row, stage_at + r * 96, 96u);
        bary_row += bary_pitch;
        if (out_image != nullptr) {
          bulk_store_row(image_row, stage_at + 4 * (kImageAt + r * 8 * A), 32u * (unsigned)A);
          image_row += image_pitch;
        }
      }
      bulk_store_commit_and_wait();
    }
    return;
  }
  store_block_rows<3>(stage, out_bary + 3 * p0, W * 3, cols, rows, vec_ok);
  if (out_image != nullptr && staged)
    store_block_rows<(A_STATIC > 0 ? A_STATIC : 1)>(stage + kImageAt, out_image + p0 * A, W * A, cols, rows, vec_ok);
}

template <int A_STATIC>
__global__ void __launch_bounds__(kChunk, 5)
raster_tile_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris,
                   int V, int T, int W, int H, float half_w, float half_h, int tiles_per_image,
                   const int *__restrict__ tile_counts, const int *__restrict__ tile_offsets,
                   const int32_t *__restrict__ tile_lists, const uint2 *__restrict__ tri_boxes,
                   int32_t *__restrict__ out_ids, float *__restrict__ out_bary, float *__restrict__ out_z,
                   const float *__restrict__ attrs, const float *__restrict__ background, int A_dyn,
                   float *__restrict__ out_image, unsigned long long *__restrict__ keys_out) {
  __shared__ __align__(16) TileSmem sm;
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  const int b = blockIdx.z;
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int tile_x0 = blockIdx.x * kTileW, tile_y0 = blockIdx.y * kTileH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 8 warps, each an 8x4 pixel block; blocks are laid out 2 across, 4 down inside the 16x16 tile.
  const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
  const int blk_x0 = tile_x0 + (warp & 1) * 8, blk_y0 = tile_y0 + (warp >> 1) * 4;
  const int ix = tile_x0 + lx, iy = tile_y0 + ly;
  const float *verts_b = verts + (size_t)b * V * 4;

  sm.key[threadIdx.x] = kEmptyKey;
  if (threadIdx.x < kTileW) sm.cx[threadIdx.x] = pixel_center(tile_x0 + threadIdx.x, half_w);
  else if (threadIdx.x < kTileW + kTileH) sm.cy[threadIdx.x - kTileW] = pixel_center(tile_y0 + threadIdx.x - kTileW, half_h);
  if (threadIdx.x == 0) sm.n_big = 0;
  if (threadIdx.x < 32) { sm.depth_bucket[threadIdx.x] = 0; sm.depth_cursor[threadIdx.x] = 0; }

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
  __syncthreads();         // publishes key / cx / cy / n_big
  const float px = sm.cx[lx], py = sm.cy[ly];
  // pixel-centre range of this warp's 8x4 block (for the conservative edge test of the big path)
  const float blk_px0 = sm.cx[(warp & 1) * 8], blk_px1 = sm.cx[(warp & 1) * 8 + 7];
  const float blk_py0 = sm.cy[(warp >> 1) * 4], blk_py1 = sm.cy[(warp >> 1) * 4 + 3];
  const float blk_pxabs = fmaxf(fabsf(blk_px0), fabsf(blk_px1)), blk_pyabs = fmaxf(fabsf(blk_py0), fabsf(blk_py1));
  if (keys_out != nullptr && ix < W && iy < H) {
    // binned pipeline: start from what the small triangles already drew (depth and id are all the
    // depth rule needs; barycentrics are not produced in this mode)
    const unsigned long long seen = keys_out[((size_t)b * H + iy) * W + ix];
    if (seen != kEmptyKey) { best.z = ordered_to_float((unsigned)(seen >> 32)); best.id = depth_key_id(seen); }
  }

  for (int base = 0; base < n_list; base += kChunk) {
    const int n_here = min(kChunk, n_list - base);
    if (base > 0) {
      __syncthreads();     // previous chunk's records fully consumed
      if (threadIdx.x == 0) sm.n_big = 0;
      if (threadIdx.x < 32) { sm.depth_bucket[threadIdx.x] = 0; sm.depth_cursor[threadIdx.x] = 0; }
      __syncthreads();
    }

    // ---- stage: the chunk's triangles are dealt round-robin to the warps (entry lane*8 + warp goes
    // to this thread) so that every warp owns a similar share of the small-triangle work.
    const int entry = lane * kWarps + warp;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0, n_seg = 0;
    bool overlaps = false;
    if (entry < n_here) {
      const int t = list ? list[base + entry] : base + entry;
      float4 p0, p1, p2;
      load_triangle(verts_b, tris, t, p0, p1, p2);
      int4 bx;
      if (tri_boxes != nullptr) {
        bx = unpack_box(__ldg(tri_boxes + (size_t)b * T + t));
      } else {
        const PixelBox pb = triangle_box(p0, p1, p2, half_w, half_h, W, H);
        bx = make_int4(pb.left, pb.right, pb.bottom, pb.top);
      }
      float m[9];
      adjugate_signed(p0.x, p1.x, p2.x, p0.y, p1.y, p2.y, p0.w, p1.w, p2.w, m);
      sm.r0[threadIdx.x] = make_float4(m[0], m[1], m[2], __int_as_float(t));
      sm.r1[threadIdx.x] = make_float4(m[3], m[4], m[5], p0.z);
      sm.r2[threadIdx.x] = make_float4(m[6], m[7], m[8], p1.z);
      sm.r3[threadIdx.x] = make_float4(p2.z, p0.w, p1.w, p2.w);
      sm.box[threadIdx.x] = bx;
      // Depth bound for the hierarchical z test of the big path: with all w > 0 the pixel depth
      // (b.z)/(b.w) is a positive-weight average of z_i/w_i, so it is >= min_i z_i/w_i; the computed
      // depth differs from the exact one by < 10 ulp of max|z_i/w_i| (three rounded barycentrics, two
      // 3-term dot products, one division), covered by the 2^-19 relative slack.
      float zlo = -INFINITY;
      if (p0.w > 0.0f && p1.w > 0.0f && p2.w > 0.0f) {
        const float d0 = p0.z / p0.w, d1 = p1.z / p1.w, d2 = p2.z / p2.w;
        const float dabs = fmaxf(fmaxf(fabsf(d0), fabsf(d1)), fabsf(d2));
        zlo = fminf(fminf(d0, d1), d2) - dabs * 1.9073486e-6f - 1e-30f;
      }
      sm.zlo[threadIdx.x] = zlo;
      // the box inside this tile, in tile-local pixel coordinates
      x0 = max(bx.x, tile_x0) - tile_x0; x1 = min(bx.y, tile_x0 + kTileW) - tile_x0;
      y0 = max(bx.z, tile_y0) - tile_y0; y1 = min(bx.w, tile_y0 + kTileH) - tile_y0;
      overlaps = x1 > x0 && y1 > y0;
      if (overlaps && (x1 - x0) * (y1 - y0) <= 64) n_seg = (y1 - y0) * ((x1 - x0 + 3) >> 2);
    }
    // ---- the warp lays the row segments of ITS small triangles out back to back (warp scan).
    // Triangles whose segments do not fit the warp's buffer, and all large ones, take the big path.
    const int seg_end = warp_inclusive_scan(n_seg);
    const bool small = n_seg > 0 && seg_end <= kWarpSegCap;
    unsigned *segs = sm.segs[warp];
    if (small) {
      int k = seg_end - n_seg;
      const int per_row = (x1 - x0 + 3) >> 2;          // 1..4 segments per row
      for (int yy = y0; yy < y1; ++yy) {
        const unsigned head = threadIdx.x | (yy << 8);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < per_row) segs[k + q] = head | ((x0 + 4 * q) << 12) | (min(4, x1 - x0 - 4 * q) << 16);
        k += per_row;
      }
    }
    int total_segs = small ? seg_end : 0;              // the fitting segments form a prefix
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total_segs = max(total_segs, __shfl_xor_sync(0xffffffffu, total_segs, d));
    // Large triangles are walked front to back (32 slices of their depth bound), which lets the
    // hierarchical z test of the big path reject most of what lies behind the first few layers.
    const bool is_big = overlaps && !small;
    int slice = 0;
    if (is_big) {
      const float zl = sm.zlo[threadIdx.x];
      slice = zl > -1.0f ? min(31, (int)((zl + 1.0f) * 16.0f)) : 0;
      atomicAdd(&sm.depth_bucket[slice], 1);
      atomicAdd(&sm.n_big, 1);
    }
    __syncwarp();

    // ---- small triangles (warp-local: no block barrier): 32 row segments per round
    unsigned short *hits = sm.hits[warp];
    for (int s0 = 0; s0 < total_segs; s0 += 32) {
      // pass 1: inside test on the (up to) four pixels of this lane's segment
      int j = 0, yy = 0, xs = 0;
      unsigned inside = 0u;
      if (s0 + lane < total_segs) {
        const unsigned seg = segs[s0 + lane];
        j = seg & 0xffu; yy = (seg >> 8) & 0xfu; xs = (seg >> 12) & 0xfu;
        const int width = seg >> 16;
        const float4 q0 = sm.r0[j], q1 = sm.r1[j], q2 = sm.r2[j];
        const float m[9] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
        const float cyv = sm.cy[yy];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float e[3], esum;
          edge_values(m, sm.cx[(xs + k) & (kTileW - 1)], cyv, e);
          if (edges_inside(e, esum) && k < width) inside |= 1u << k;
        }
      }
      // compact the inside pixels of the warp and deal them to the lanes again
      const int mine = __popc(inside);
      const int upto = warp_inclusive_scan(mine);
      const int n_hits = __shfl_sync(0xffffffffu, upto, 31);
      int at = upto - mine;
      while (inside) {
        const int k = __ffs(inside) - 1;
        inside &= inside - 1;
        hits[at++] = (unsigned short)((j << 8) | (yy * kTileW + xs + k));
      }
      __syncwarp();
      // pass 2: barycentrics / depth for exactly those pixels, depth resolve by packed atomicMin
      for (int h = lane; h < n_hits; h += 32) {
        const unsigned hit = hits[h];
        const int jj = hit >> 8, pix = hit & 0xffu;
        const float4 q0 = sm.r0[jj], q1 = sm.r1[jj], q2 = sm.r2[jj], q3 = sm.r3[jj];
        const float m[9] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
        const float zc[3] = {q1.w, q2.w, q3.x};
        const float wc[3] = {q3.y, q3.z, q3.w};
        float e[3], esum, bc[3], z;
        edge_values(m, sm.cx[pix & (kTileW - 1)], sm.cy[pix >> kTileShiftX], e);
        edges_inside(e, esum);
        if (fragment_depth(e, esum, zc, wc, bc, z))
          atomicMin(&sm.key[pix], depth_key(z, __float_as_int(q0.w)));
      }
      __syncwarp();
    }

    // ---- big triangles: warp per 8x4 block, ballot cull, lane per pixel (needs everyone's records)
    __syncthreads();
    const int n_big = sm.n_big;
    if (n_big > 0) {
      // exclusive prefix of the slice counts (every warp computes it for itself), then placement
      const int count = sm.depth_bucket[lane];
      const int start = warp_inclusive_scan(count) - count;
      const int my_start = __shfl_sync(0xffffffffu, start, slice);
      if (is_big) sm.big_list[my_start + atomicAdd(&sm.depth_cursor[slice], 1)] = (unsigned short)threadIdx.x;
      __syncthreads();
    }
    for (int g0 = 0; g0 < n_big; g0 += 32) {
      // Farthest depth any pixel of this warp's block currently holds (1.0 while a pixel is empty):
      // a triangle whose depth bound lies beyond it cannot change the block (K.cpp:401 rejects z > zbuf).
      const float block_zmax = ordered_to_float(__reduce_max_sync(0xffffffffu, float_to_ordered(best.z)));
      bool touches = false;
      if (g0 + lane < n_big) {
        const int jj = sm.big_list[g0 + lane];
        const int4 bx = sm.box[jj];
        touches = bx.x < blk_x0 + 8 && bx.y > blk_x0 && bx.z < blk_y0 + 4 && bx.w > blk_y0 &&
                  sm.zlo[jj] <= block_zmax;
        if (touches) {
          // Conservative edge test: an edge function is linear, so its largest exact value over the
          // block's pixel centres sits at a corner; the fp32 evaluation at any pixel is within
          // 3 ulp-sums of it, the corner evaluation too.  If even that bound is negative for one edge,
          // no pixel of the block can pass the inside test (K.cpp:93-98).
          const float4 q0 = sm.r0[jj], q1 = sm.r1[jj], q2 = sm.r2[jj];
          const float ea[3] = {q0.x, q1.x, q2.x}, eb[3] = {q0.y, q1.y, q2.y}, ec[3] = {q0.z, q1.z, q2.z};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float hi = ea[i] * (ea[i] >= 0.0f ? blk_px1 : blk_px0) + eb[i] * (eb[i] >= 0.0f ? blk_py1 : blk_py0) + ec[i];
            const float mag = fabsf(ea[i]) * blk_pxabs + fabsf(eb[i]) * blk_pyabs + fabsf(ec[i]);
            if (hi < -9.5367432e-7f * mag) touches = false;             // 2^-20 = 16 ulp of the magnitude sum
          }
        }
      }
      unsigned todo = __ballot_sync(0xffffffffu, touches);
      while (todo) {
        const int j = sm.big_list[g0 + __ffs(todo) - 1];
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
  __syncthreads();           // all keys final; record planes dead from here on

  // ---- resolve: minimum of the two paths
  const unsigned long long key_small = sm.key[ly * kTileW + lx];
  const unsigned long long key_big = best.id >= 0 ? depth_key(best.z, best.id) : kEmptyKey;
  if (keys_out != nullptr) {
    // large-triangle pass of the binned pipeline: merge into the global keys (this CTA is the only
    // writer of its pixels now; the scatter kernel has finished), resolve_kernel does the rest.
    if (ix < W && iy < H) {
      const size_t p = ((size_t)b * H + iy) * W + ix;
      const unsigned long long mine = min(key_small, key_big);
      if (mine < keys_out[p]) keys_out[p] = mine;
    }
    return;
  }
  if (key_small < key_big) {     // the winner came through the key buffer: re-evaluate it once
    const int t = depth_key_id(key_small);
    float4 p0, p1, p2;
    load_triangle(verts_b, tris, t, p0, p1, p2);
    evaluate_winner(p0, p1, p2, px, py, t, best);
  }
  float *stage = reinterpret_cast<float *>(sm.r0) + warp * (32 * 16);
  block_epilogue<A_STATIC>(stage, b, blk_x0, blk_y0, W, H, V, best, tris, attrs, background, A,
                           out_ids, out_bary, out_z, out_image);
}

// ---------------------------------------------------------------------------------------------
// Small triangles: global scatter
// ---------------------------------------------------------------------------------------------

// Pixel-centre NDC coordinates of every column and row (K.cpp:376-377), built once per image size.
__global__ void pixel_centers_kernel(float *__restrict__ centers, int W, int H, float half_w, float half_h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) centers[i] = pixel_center(i, half_w);
  else if (i < W + H) centers[i] = pixel_center(i - W, half_h);
}

constexpr int kScatterWarps = 8;
constexpr int kScatterSegCap = 512;          // row segments a warp holds per round (one triangle has <= 64)
constexpr int kSmallBox = 16;                // largest box side the scatter path takes

struct ScatterWarpSmem {
  float4 r0[32], r1[32], r2[32], r3[32];     // setup records of the warp's 32 triangles (planes as in TileSmem)
  int2 origin[32];                           // left, bottom of each triangle's pixel box
  unsigned short segs[kScatterSegCap];       // slot | dy << 5 | dx << 9 | width << 13
  unsigned short hits[160];                  // queue of inside pixels: slot | dy << 5 | x offset << 9
};

__global__ void __launch_bounds__(kScatterWarps * 32)
scatter_small_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris, int V, int T, int W, int H,
                     float half_w, float half_h, int tiles_x, int tiles_per_image,
                     const float *__restrict__ centers, uint2 *__restrict__ tri_boxes,
                     int *__restrict__ tile_counts, unsigned long long *__restrict__ keys) {
  __shared__ __align__(16) ScatterWarpSmem sm_all[kScatterWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScatterWarpSmem &sm = sm_all[warp];
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const float *cx = centers, *cy = centers + W;
  unsigned long long *keys_b = keys + (size_t)b * H * W;

  // ---- setup: one triangle per lane
  int n_seg = 0, bw = 0, bh = 0;
  uint2 big_box = make_uint2(0u, 0u);
  if (t < T) {
    float4 p0, p1, p2;
    load_triangle(verts + (size_t)b * V * 4, tris, t, p0, p1, p2);
    const PixelBox box = triangle_box(p0, p1, p2, half_w, half_h, W, H);
    bw = box.right - box.left; bh = box.top - box.bottom;
    if (bw > 0 && bh > 0) {
      if (bw <= kSmallBox && bh <= kSmallBox) {
        float m[9];
        adjugate_signed(p0.x, p1.x, p2.x, p0.y, p1.y, p2.y, p0.w, p1.w, p2.w, m);
        sm.r0[lane] = make_float4(m[0], m[1], m[2], __int_as_float(t));
        sm.r1[lane] = make_float4(m[3], m[4], m[5], p0.z);
        sm.r2[lane] = make_float4(m[6], m[7], m[8], p1.z);
        sm.r3[lane] = make_float4(p2.z, p0.w, p1.w, p2.w);
        sm.origin[lane] = make_int2(box.left, box.bottom);
        n_seg = bh * ((bw + 3) >> 2);
      } else {
        big_box = pack_box(box);
      }
    }
    tri_boxes[(size_t)b * T + t] = big_box;            // empty for small triangles: bin_fill skips them
  }
  // large triangles are only counted into the tiles they touch (raster_tile_kernel draws them)
  int *counts = tile_counts + (size_t)b * tiles_per_image;
  for_each_tile(big_box, tiles_x, [&](int tile, int) { atomicAdd(counts + tile, 1); });

  // pass 2 of the scatter: barycentrics / depth of queued inside pixels, depth resolve by packed
  // atomicMin in global memory (L2).  Lane h of the warp takes queue entry first + h.
  auto shade_hits = [&](int first, int count) {
    if (lane < count) {
      const unsigned hit = sm.hits[first + lane];
      const int j = hit & 31u, dy = (hit >> 5) & 15u, xo = hit >> 9;
      const float4 q0 = sm.r0[j], q1 = sm.r1[j], q2 = sm.r2[j], q3 = sm.r3[j];
      const float m[9] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
      const float zc[3] = {q1.w, q2.w, q3.x};
      const float wc[3] = {q3.y, q3.z, q3.w};
      const int2 org = sm.origin[j];
      const int x = org.x + xo, y = org.y + dy;
      float e[3], esum, bc[3], z;
      edge_values(m, __ldg(cx + x), __ldg(cy + y), e);
      edges_inside(e, esum);
      if (fragment_depth(e, esum, zc, wc, bc, z))
        atomicMin(keys_b + (size_t)y * W + x, depth_key(z, __float_as_int(q0.w)));
    }
  };

  // ---- rounds: the longest prefix (in lane order) of the remaining small triangles whose segments fit
  int queued = 0;                                       // inside pixels waiting in sm.hits (warp-uniform)
  unsigned remaining = __ballot_sync(0xffffffffu, n_seg > 0);
  while (remaining) {
    const int mine = (remaining >> lane) & 1u ? n_seg : 0;
    const int seg_end = warp_inclusive_scan(mine);
    const bool fits = mine > 0 && seg_end <= kScatterSegCap;
    if (fits) {
      int k = seg_end - mine;
      const int per_row = (bw + 3) >> 2;               // 1..4 segments per row
      for (int dy = 0; dy < bh; ++dy) {
        const unsigned head = lane | (dy << 5);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < per_row) sm.segs[k + q] = (unsigned short)(head | ((4 * q) << 9) | (min(4, bw - 4 * q) << 13));
        k += per_row;
      }
    }
    int total_segs = fits ? seg_end : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total_segs = max(total_segs, __shfl_xor_sync(0xffffffffu, total_segs, d));
    remaining &= ~__ballot_sync(0xffffffffu, fits);
    __syncwarp();

    for (int s0 = 0; s0 < total_segs; s0 += 32) {
      // pass 1: inside test on the (up to) four pixels of this lane's segment
      unsigned seg = 0u, inside = 0u;
      if (s0 + lane < total_segs) {
        seg = sm.segs[s0 + lane];
        const int j = seg & 31u, dy = (seg >> 5) & 15u, dx = (seg >> 9) & 15u, width = seg >> 13;
        const float4 q0 = sm.r0[j], q1 = sm.r1[j], q2 = sm.r2[j];
        const float m[9] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
        const int2 org = sm.origin[j];
        const float cyv = __ldg(cy + org.y + dy);
        const float *cxp = cx + org.x + dx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < width) {
            float e[3], esum;
            edge_values(m, __ldg(cxp + k), cyv, e);
            if (edges_inside(e, esum)) inside |= 1u << k;
          }
        }
      }
      // append the warp's inside pixels to the queue (at most 4 per lane, < 32 were waiting)
      const int mine_hits = __popc(inside);
      const int upto = warp_inclusive_scan(mine_hits);
      int at = queued + upto - mine_hits;
      const unsigned base_bits = seg & 0x1ffu, dx0 = (seg >> 9) & 15u;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (inside & (1u << k)) sm.hits[at++] = (unsigned short)(base_bits | ((dx0 + k) << 9));
      queued += __shfl_sync(0xffffffffu, upto, 31);
      __syncwarp();
      // shade full warps of queued pixels; keep the remainder (< 32) for the next round
      int first = 0;
      for (; queued - first >= 32; first += 32) shade_hits(first, 32);
      const int left = queued - first;
      unsigned short carry = 0;
      if (first > 0 && lane < left) carry = sm.hits[first + lane];
      __syncwarp();
      if (first > 0 && lane < left) sm.hits[lane] = carry;
      queued = left;
      __syncwarp();
    }
  }
  shade_hits(0, queued);                                // final partial warp of pixels
}

// ---------------------------------------------------------------------------------------------
// Resolve: depth keys -> outputs
// ---------------------------------------------------------------------------------------------

constexpr int kResolveWarps = 8;

template <int A_STATIC>
__global__ void __launch_bounds__(kResolveWarps * 32)
resolve_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris, int V, int W, int H,
               const float *__restrict__ centers,
               const unsigned long long *__restrict__ keys,
               int32_t *__restrict__ out_ids, float *__restrict__ out_bary, float *__restrict__ out_z,
               const float *__restrict__ attrs, const float *__restrict__ background, int A_dyn,
               float *__restrict__ out_image) {
  __shared__ __align__(16) float stage_all[kResolveWarps][32 * 16];
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the CTA covers a 16x16 pixel tile: 2 blocks across, 4 down
  const int b = blockIdx.z;
  const int blk_x0 = (blockIdx.x * 2 + (warp & 1)) * 8, blk_y0 = (blockIdx.y * (kResolveWarps / 2) + (warp >> 1)) * 4;
  if (blk_x0 >= W || blk_y0 >= H) return;
  const int ix = blk_x0 + (lane & 7), iy = blk_y0 + (lane >> 3);
  Fragment best;
  fragment_clear(best);
  if (ix < W && iy < H) {
    const unsigned long long key = keys[((size_t)b * H + iy) * W + ix];
    if (key != kEmptyKey) {
      const int t = depth_key_id(key);
      float4 p0, p1, p2;
      load_triangle(verts + (size_t)b * V * 4, tris, t, p0, p1, p2);
      evaluate_winner(p0, p1, p2, __ldg(centers + ix), __ldg(centers + W + iy), t, best);
    }
  }
  // (Gathering the 3*A corner attributes here, together with the vertices, was measured: 60 registers
  // instead of 38 cost more occupancy than the shorter dependency chain gained: 0.355 -> 0.411 ms.)
  block_epilogue<A_STATIC>(stage_all[warp], b, blk_x0, blk_y0, W, H, V, best, tris, attrs, background, A,
                           out_ids, out_bary, out_z, out_image);
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
                         const uint2 *boxes, int32_t *ids, float *bary, float *z, const float *attrs, const float *bg, int A,
                         float *image, unsigned long long *keys_out, cudaStream_t stream) {
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const int tiles = tiles_x * tiles_y;
  const float half_w = (float)(0.5 * W), half_h = (float)(0.5 * H);   // K.cpp:309-310
  dim3 grid(tiles_x, tiles_y, B);
  StageScope timed(ctx, PMR_STAGE_RASTER, stream);
#define PMR_LAUNCH(AS)                                                                              \
  raster_tile_kernel<AS><<<grid, kChunk, 0, stream>>>(verts, tris, V, T, W, H, half_w, half_h,      \
                                                     tiles, counts, offsets, lists, boxes, ids,     \
                                                     bary, z, attrs, bg, A, image, keys_out)
  if (image == nullptr || keys_out != nullptr) PMR_LAUNCH(0);
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
    // tiny mesh: one kernel, every tile walks all triangles and writes the outputs itself
    return launch_raster(ctx, verts, tris, B, V, T, W, H, nullptr, nullptr, nullptr, nullptr, ids, bary, z, attrs,
                         bg, A, image, nullptr, stream);
  }

  const size_t n_tiles = (size_t)B * tiles, n_pixels = (size_t)B * H * W;
  if (n_tiles > (size_t)INT_MAX) return set_error(ctx, PMR_ERR_SIZE, "too many screen tiles");
  int rc;
  // pixel-centre table of this image size (rebuilt only when the size changes)
  if (ctx->centers_w != W || ctx->centers_h != H) {
    rc = ctx->centers.reserve(ctx, (size_t)(W + H) * sizeof(float));
    if (rc) return rc;
    pixel_centers_kernel<<<(W + H + 255) / 256, 256, 0, stream>>>((float *)ctx->centers.ptr, W, H, half_w, half_h);
    ctx->launches += 1;
    ctx->centers_w = W; ctx->centers_h = H;
  }
  const float *centers = (const float *)ctx->centers.ptr;
  // workspace: [total u64 | counts | cursors | offsets | boxes], keys
  rc = ctx->bins.reserve(ctx, 16 + n_tiles * 3 * sizeof(int) + (size_t)B * T * sizeof(uint2) + 64);
  if (rc) return rc;
  rc = ctx->keys.reserve(ctx, n_pixels * sizeof(unsigned long long));
  if (rc) return rc;
  char *base = (char *)ctx->bins.ptr;
  unsigned long long *total = (unsigned long long *)base;
  int *counts = (int *)(base + 16);
  int *cursors = counts + n_tiles;
  int *offsets = cursors + n_tiles;
  uint2 *boxes = (uint2 *)(((uintptr_t)(offsets + n_tiles) + 15) & ~(uintptr_t)15);
  unsigned long long *keys = (unsigned long long *)ctx->keys.ptr;

  {
    StageScope timed(ctx, PMR_STAGE_BIN, stream);
    PMR_CUDA(ctx, cudaMemsetAsync(base, 0, 16 + n_tiles * 2 * sizeof(int), stream));
    PMR_CUDA(ctx, cudaMemsetAsync(keys, 0xff, n_pixels * sizeof(unsigned long long), stream));   // kEmptyKey
  }
  {
    StageScope timed(ctx, PMR_STAGE_SCATTER, stream);
    scatter_small_kernel<<<dim3((T + kScatterWarps * 32 - 1) / (kScatterWarps * 32), B), kScatterWarps * 32, 0, stream>>>(
        verts, tris, V, T, W, H, half_w, half_h, tiles_x, tiles, centers, boxes, counts, keys);
    ctx->launches += 1;
    rc = check_launch(ctx, "scatter_small_kernel");
    if (rc) return rc;
  }
  unsigned long long n_entries = 0;
  {
    StageScope timed(ctx, PMR_STAGE_BIN, stream);
    bin_offsets_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, stream>>>(counts, (int)n_tiles, offsets, total);
    ctx->launches += 1;
    rc = check_launch(ctx, "bin_offsets_kernel");
    if (rc) return rc;
    // The list length of the large triangles is data dependent: read it back (8 bytes, pinned).
    PMR_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    PMR_CUDA(ctx, cudaStreamSynchronize(stream));
    n_entries = *ctx->mailbox;
    ctx->last_bin_entries = n_entries;
    if (n_entries >= (1ull << 31)) return set_error(ctx, PMR_ERR_SIZE, "tile lists exceed 2^31 entries");
    if (n_entries > 0) {
      rc = ctx->lists.reserve(ctx, (size_t)(n_entries + 1) * sizeof(int32_t));
      if (rc) return rc;
      bin_fill_kernel<<<dim3((T + 255) / 256, B), 256, 0, stream>>>(boxes, T, tiles_x, tiles, offsets, cursors,
                                                                    (int32_t *)ctx->lists.ptr);
      ctx->launches += 1;
      rc = check_launch(ctx, "bin_fill_kernel");
      if (rc) return rc;
    }
  }
  if (n_entries > 0) {
    rc = launch_raster(ctx, verts, tris, B, V, T, W, H, counts, offsets, (const int32_t *)ctx->lists.ptr, boxes,
                       ids, bary, z, attrs, bg, A, image, keys, stream);
    if (rc) return rc;
  }
  {
    StageScope timed(ctx, PMR_STAGE_RESOLVE, stream);
    dim3 grid((W + 15) / 16, (H + 2 * kResolveWarps - 1) / (2 * kResolveWarps), B);
#define PMR_RESOLVE(AS)                                                                                          \
  resolve_kernel<AS><<<grid, kResolveWarps * 32, 0, stream>>>(verts, tris, V, W, H, centers, keys, ids,          \
                                                              bary, z, attrs, bg, A, image)
    if (image == nullptr) PMR_RESOLVE(0);
    else if (A == 4) PMR_RESOLVE(4);
    else if (A == 9) PMR_RESOLVE(9);
    else if (A == 12) PMR_RESOLVE(12);
    else if (A == 13) PMR_RESOLVE(13);
    else PMR_RESOLVE(0);
#undef PMR_RESOLVE
    ctx->launches += 1;
    rc = check_launch(ctx, "resolve_kernel");
  }
  return rc;
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
