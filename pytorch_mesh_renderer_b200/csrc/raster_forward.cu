// raster_forward.cu -- screen-tile binning and the per-tile raster kernel (forward pass of
// rasterize_triangles, K.cpp:302-419, batched over images) with the attribute interpolation
// of rasterize_clip_space (rast.py:118-150) fused into its epilogue.
//
// Pipeline per call (all on one stream):
//   bin_count_kernel   one thread per (image, triangle): pixel box (kept, 8 bytes) -> per-tile counts
//   bin_offsets_kernel warp-aggregated allocation of one contiguous list range per tile
//   bin_fill_kernel    writes triangle ids into the tile lists (order inside a list is arbitrary)
//   raster_tile_kernel one CTA per 16x16-pixel tile.  Triangle setup records (edge equations, z, w,
//                      pixel box) are staged into shared memory 256 at a time and split by the size
//                      of their box inside the tile:
//                        small (<= 64 pixels): the boxes are cut into row segments of up to four
//                          pixels, the segments of all small triangles are laid out back to back and
//                          dealt to the 256 threads (balanced work whatever the triangle sizes);
//                          a thread runs the inside test on its four pixels, the inside pixels of a
//                          warp are compacted and dealt to its lanes again for barycentrics / depth,
//                          and depth is resolved with a packed 64-bit atomicMin(depth bits, ~id) on a
//                          shared-memory key per pixel;
//                        big: one WARP per 8x4 pixel block culls the records with a ballot and each
//                          lane tests its own pixel, keeping its winner in registers.
//                      The per-pixel minimum of both paths is the winner; a winner that came through
//                      the key buffer is re-evaluated once (same arithmetic, same bits) and ids /
//                      barycentrics / z (+ interpolated attributes) are written.
// Meshes with few triangles skip binning: every tile walks the whole triangle array.
//
// The depth rule is order independent (min z, then max id -- SURVEY.md F1), so neither list order
// nor atomic order matters and the result is deterministic.
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

__global__ void __launch_bounds__(256)
bin_count_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris,
                 int V, int T, int W, int H, float half_w, float half_h, int tiles_x, int tiles_per_image,
                 uint2 *__restrict__ tri_boxes, int *__restrict__ tile_counts) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  uint2 packed = make_uint2(0u, 0u);
  if (t < T) {
    float4 p0, p1, p2;
    load_triangle(verts + (size_t)b * V * 4, tris, t, p0, p1, p2);
    packed = pack_box(triangle_box(p0, p1, p2, half_w, half_h, W, H));
    tri_boxes[(size_t)b * T + t] = packed;
  }
  int *counts = tile_counts + (size_t)b * tiles_per_image;
  for_each_tile(packed, tiles_x, [&](int tile, int) { atomicAdd(counts + tile, 1); });
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
  unsigned long long key[kTilePixels];   // packed (depth, ~id) minimum per pixel (small-triangle path)
  unsigned segs[kWarps][kWarpSegCap];    // per warp: slot | row << 8 | first column << 12 | width << 16
  unsigned short hits[kWarps][128];      // per warp: inside pixels of 32 segments (slot << 8 | pixel)
  unsigned short big_list[kChunk];
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

__device__ __forceinline__ int warp_inclusive_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += up;
  }
  return v;
}

// Writes N floats per pixel of the warp's 8x4 block from the shared staging area (`stage`: 4 rows x
// 8*N contiguous floats) to `dst_row0` (+ row * row_stride floats).  The vector path needs 16-byte
// aligned rows and a full-width block; otherwise scalars.  N is a compile-time constant so that the
// row/column split is a multiply-shift, not a division.
template <int N>
__device__ __forceinline__ void store_block_rows(const float *stage, float *dst_row0, size_t row_stride,
                                                 int cols, int rows, bool vec_ok) {
  const int lane = threadIdx.x & 31;
  constexpr int run = 8 * N;                           // floats per full block row
  if (vec_ok && cols == 8) {
    constexpr int v4_per_row = 2 * N;
#pragma unroll
    for (int k0 = 0; k0 < 4 * v4_per_row; k0 += 32) {
      const int k = k0 + lane;
      const int r = k / v4_per_row, c = k % v4_per_row;
      if (k < 4 * v4_per_row && r < rows)
        reinterpret_cast<float4 *>(dst_row0 + r * row_stride)[c] = reinterpret_cast<const float4 *>(stage + r * run)[c];
    }
  } else {
    const int live = cols * N;
    for (int k = lane; k < rows * run; k += 32) {
      const int r = k / run, c = k % run;
      if (c < live) dst_row0[r * row_stride + c] = stage[r * run + c];
    }
  }
}

template <int A_STATIC>
__global__ void __launch_bounds__(kChunk, 5)
raster_tile_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris,
                   int V, int T, int W, int H, float half_w, float half_h, int tiles_per_image,
                   const int *__restrict__ tile_counts, const int *__restrict__ tile_offsets,
                   const int32_t *__restrict__ tile_lists, const uint2 *__restrict__ tri_boxes,
                   int32_t *__restrict__ out_ids, float *__restrict__ out_bary, float *__restrict__ out_z,
                   const float *__restrict__ attrs, const float *__restrict__ background, int A_dyn,
                   float *__restrict__ out_image) {
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

  for (int base = 0; base < n_list; base += kChunk) {
    const int n_here = min(kChunk, n_list - base);
    if (base > 0) {
      __syncthreads();     // previous chunk's records fully consumed
      if (threadIdx.x == 0) sm.n_big = 0;
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
    append_slots(overlaps && !small, sm.big_list, &sm.n_big);
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
    for (int g0 = 0; g0 < n_big; g0 += 32) {
      bool touches = false;
      if (g0 + lane < n_big) {
        const int4 bx = sm.box[sm.big_list[g0 + lane]];
        touches = bx.x < blk_x0 + 8 && bx.y > blk_x0 && bx.z < blk_y0 + 4 && bx.w > blk_y0;
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

  // ---- resolve: minimum of the two paths; re-evaluate the winner if it came from the key buffer
  const unsigned long long key_small = sm.key[ly * kTileW + lx];
  const unsigned long long key_big = best.id >= 0 ? depth_key(best.z, best.id) : kEmptyKey;
  if (key_small < key_big) {
    const int t = depth_key_id(key_small);
    float4 p0, p1, p2;
    load_triangle(verts_b, tris, t, p0, p1, p2);
    float m[9], e[3], esum, bc[3], z;
    adjugate_signed(p0.x, p1.x, p2.x, p0.y, p1.y, p2.y, p0.w, p1.w, p2.w, m);
    edge_values(m, px, py, e);
    edges_inside(e, esum);
    const float zc[3] = {p0.z, p1.z, p2.z}, wc[3] = {p0.w, p1.w, p2.w};
    fragment_depth(e, esum, zc, wc, bc, z);
    best.z = z; best.id = t; best.b0 = bc[0]; best.b1 = bc[1]; best.b2 = bc[2];
  }
  const bool covered = best.id >= 0;
  const int id = covered ? best.id : 0;

  // ---- epilogue: transpose the warp's block through shared memory, store whole 16-byte vectors
  float *stage = reinterpret_cast<float *>(sm.r0) + warp * (32 * 16);
  const int cols = min(8, W - blk_x0), rows = min(4, H - blk_y0);     // <= 0: block outside the image
  if (cols <= 0 || rows <= 0) return;
  const bool vec_ok = (W & 3) == 0 &&
      (((uintptr_t)out_ids | (uintptr_t)out_z | (uintptr_t)out_bary | (uintptr_t)out_image) & 15) == 0;
  const size_t p0 = ((size_t)b * H + blk_y0) * W + blk_x0;            // first pixel of the block
  // ids (as raw bits), z, barycentrics
  stage[lane] = __int_as_float(id);
  stage[32 + lane] = best.z;
  stage[64 + 3 * lane + 0] = best.b0;
  stage[64 + 3 * lane + 1] = best.b1;
  stage[64 + 3 * lane + 2] = best.b2;
  __syncwarp();
  store_block_rows<1>(stage, reinterpret_cast<float *>(out_ids) + p0, (size_t)W, cols, rows, vec_ok);
  store_block_rows<1>(stage + 32, out_z + p0, (size_t)W, cols, rows, vec_ok);
  store_block_rows<3>(stage + 64, out_bary + 3 * p0, (size_t)W * 3, cols, rows, vec_ok);
  __syncwarp();

  if (out_image != nullptr) {
    // rast.py:118-150: corner attributes weighted by barycentrics, alpha, background blend.
    float *o = stage + lane * A;                      // compiled-in attribute counts go through the staging area
    const bool staged = A_STATIC > 0 && A_STATIC <= 16;
    float *direct = out_image + (p0 + (size_t)(lane >> 3) * W + (lane & 7)) * A;
    float *dst = staged ? o : direct;
    const bool in_image = (lane & 7) < cols && (lane >> 3) < rows;
    if (staged || in_image) {
      if (!covered) {
        for (int a = 0; a < A; ++a) dst[a] = __ldg(background + a);
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
          dst[a] = alpha * img + one_minus * __ldg(background + a);
        }
      }
    }
    if (staged) {
      __syncwarp();
      store_block_rows<(A_STATIC > 0 ? A_STATIC : 1)>(stage, out_image + p0 * A, (size_t)W * A, cols, rows, vec_ok);
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
                         const uint2 *boxes, int32_t *ids, float *bary, float *z, const float *attrs, const float *bg, int A,
                         float *image, cudaStream_t stream) {
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const int tiles = tiles_x * tiles_y;
  const float half_w = (float)(0.5 * W), half_h = (float)(0.5 * H);   // K.cpp:309-310
  dim3 grid(tiles_x, tiles_y, B);
  StageScope timed(ctx, PMR_STAGE_RASTER, stream);
#define PMR_LAUNCH(AS)                                                                              \
  raster_tile_kernel<AS><<<grid, kChunk, 0, stream>>>(verts, tris, V, T, W, H, half_w, half_h,      \
                                                     tiles, counts, offsets, lists, boxes, ids,     \
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
    return launch_raster(ctx, verts, tris, B, V, T, W, H, nullptr, nullptr, nullptr, nullptr, ids, bary, z, attrs,
                         bg, A, image, stream);
  }

  const size_t n_tiles = (size_t)B * tiles;
  int32_t *lists = nullptr;
  int *counts = nullptr, *offsets = nullptr;
  uint2 *ranges = nullptr;      // packed pixel box per (image, triangle)
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
  ranges = (uint2 *)(((uintptr_t)(offsets + n_tiles) + 15) & ~(uintptr_t)15);

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
  return launch_raster(ctx, verts, tris, B, V, T, W, H, counts, offsets, lists, ranges, ids, bary, z, attrs, bg, A,
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
