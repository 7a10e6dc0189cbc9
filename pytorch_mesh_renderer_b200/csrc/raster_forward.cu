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
//                         Larger triangles are appended (id + pixel box) to the image's LARGE list.
//   raster_tile_kernel    persistent CTAs take (image, 64x64 macro tile) items of the images that have
//                         large triangles from a work counter: the image's large list is scanned once
//                         per macro tile with a box test into shared memory, each 16x16 screen tile
//                         filters that short list, the survivors are staged 256 at a time as setup
//                         records and walked front to back by a tile-level depth bound; a warp per 8x4
//                         pixel block culls them (box, hierarchical z, conservative edge test) with a
//                         ballot, one lane per pixel keeps its winner in registers (in a mesh of small
//                         triangles the ones that are small inside the tile take the segment path on a
//                         shared-memory key), result merged into the global keys.  Images without large
//                         triangles cost one counter load per CTA; the whole pass needs no host round
//                         trip (no list sizes to read back).
//   resolve_kernel        one warp per 8x4 pixel block: decodes the winner, re-evaluates its
//                         barycentrics / depth once (same arithmetic, same bits), interpolates the
//                         attributes and stores ids / z as full sectors, barycentrics / image through a
//                         shared-memory transpose and the bulk-copy engine.  With SHADE (render path)
//                         the nine interpolated channels are lit in registers and only RGBA is written.
// Meshes with few triangles (<= small_mesh_threshold) run raster_tile_kernel alone: every tile walks
// the whole triangle array and writes the outputs itself.
//
// Neither list order nor atomic order can change a result, so the pass is deterministic.
#include "pmr_internal.cuh"
#include "raster_math.cuh"
#include "shade_math.cuh"
#include "tensor_maps.cuh"

namespace pmr {

// ---------------------------------------------------------------------------------------------
// Pixel boxes
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

// ---------------------------------------------------------------------------------------------
// Per-tile raster kernel
// ---------------------------------------------------------------------------------------------

constexpr int kChunk = 256;                 // triangles staged per round == threads per CTA
constexpr int kTilePixels = kTileW * kTileH;
constexpr int kWarps = kChunk / 32;
constexpr int kWarpSegCap = 256;            // row segments of small triangles a warp holds per round

constexpr int kMacroCapacity = 1024;        // see kMacroCap below

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
  unsigned short cand[2 * kChunk];       // macro-list entries that passed the box-vs-tile test
  uint2 mbox[kMacroCapacity];            // macro list: packed pixel boxes ...
  int mslot[kMacroCapacity];             // ... and positions in the image's candidate list
  unsigned short active[kChunk];         // images of the current slice that have candidates (B <= 65535)
  int warp_count[kWarps];
  int n_big, n_cand, n_macro, next_item;
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
// Tensor maps of the outputs that are staged: barycentrics [B*H][3W] box 24 x 4, image [B*H][A*W] box 8A x 4;
// and of the resolve kernel's input, the depth keys as [B*H][2W] 32-bit words, box 64 x 4 (a strip of 32 x 4 pixels).
struct OutputMaps {
  CUtensorMap bary, image, keys;
};

template <int A_STATIC>
__device__ __forceinline__ void block_epilogue(float *stage, int b, int blk_x0, int blk_y0, int W, int H, int V,
                                               const Fragment &best, const int32_t *__restrict__ tris,
                                               const float *__restrict__ attrs, const float *__restrict__ background,
                                               int A, int32_t *__restrict__ out_ids, float *__restrict__ out_bary,
                                               float *__restrict__ out_z, float *__restrict__ out_image,
                                               const float *corners = nullptr, const OutputMaps *maps = nullptr,
                                               bool defer_wait = false) {
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
  // `maps` (stage 128-byte aligned): the two staged arrays leave as ONE tensor store each, issued by one
  // lane; columns beyond the image are clipped by the copy engine.  The eight per-row bulk copies below cost
  // ~130 warp instructions per block (every lane's copy is issued in turn through uniform registers), more
  // than a quarter of the resolve kernel.
  if (maps != nullptr && vec_ok && rows == 4 && (out_image == nullptr || staged)) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy reads
    __syncwarp();
    // defer_wait: the caller waits for the reads before the stage is written again or the warp exits -- on the
    // lane elect.sync picks, which is the same lane every time for the same (full) mask.  (elect.sync rather than
    // `lane == 0`: the copy instructions take their operands from uniform registers, and only behind an election
    // does the compiler know that one lane issues them; otherwise it loops over the "active lanes" -- 11
    // instructions per copy.)
    if (elect_one()) {
      const unsigned stage_at = (unsigned)__cvta_generic_to_shared(stage);
      tma_store_2d(&maps->bary, 3 * blk_x0, b * H + blk_y0, stage_at);
      if (out_image != nullptr) tma_store_2d(&maps->image, A * blk_x0, b * H + blk_y0, stage_at + 4u * (unsigned)kImageAt);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (!defer_wait) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  if (bulk) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy reads
    __syncwarp();
    // Eight lanes issue one row each in the same instruction (lanes 0..3: barycentric rows, 4..7: image rows)
    // and wait for their own bulk group; issuing all eight from one lane cost a quarter of the kernel's
    // instructions (ncu, profiles/r01/lines_r1c_*).
    if (lane < 8 && (lane < 4 || out_image != nullptr)) {
      const int r = lane & 3;
      const bool is_image = lane >= 4;
      const size_t row_pixel = p0 + (size_t)r * W;
      float *dst = is_image ? out_image + row_pixel * A : out_bary + 3 * row_pixel;
      const unsigned stage_at = (unsigned)__cvta_generic_to_shared(stage);
      const unsigned src = stage_at + (is_image ? 4u * (unsigned)(kImageAt + r * 8 * A) : 96u * (unsigned)r);
      bulk_store_row(dst, src, is_image ? 32u * (unsigned)A : 96u);
      bulk_store_commit_and_wait();
    }
    return;
  }
  store_block_rows<3>(stage, out_bary + 3 * p0, W * 3, cols, rows, vec_ok);
  if (out_image != nullptr && staged)
    store_block_rows<(A_STATIC > 0 ? A_STATIC : 1)>(stage + kImageAt, out_image + p0 * A, W * A, cols, rows, vec_ok);
}

// Conservative lower bound of the depth any pixel centre in the rectangle [px0,px1] x [py0,py1] can get
// from one triangle (all w > 0), or -inf when no bound can be given.  Hierarchical z at tile granularity:
// the triangle-level bound (min of the vertex depths) is useless for a large slanted triangle that is far
// behind the surface in THIS tile.  Used when a triangle is staged for a tile: it becomes the key of the
// front-to-back order and the value the per-block hierarchical z test compares.
//
// Why it is conservative.  The reference's depth at a pixel (K.cpp:384-397) is, up to roundings,
//   Z(e) = sum_i e_i z_i / sum_i e_i w_i  with the COMPUTED edge values e_i >= 0 as weights.
// With E_i(p) = a_i px + b_i py + c_i the exact edge functions of the fp32 coefficients, N(p) = sum z_i E_i(p)
// and D(p) = sum w_i E_i(p) are affine in p, so R = N / D is quasi-linear wherever D > 0: over the rectangle
// its minimum sits at a corner.  Error terms, with eps = 2^-24, mag_i >= |a_i px| + |b_i py| + |c_i| on the
// rectangle, Ed = sum w_i mag_i, dabs = max |z_i / w_i|, q = 16 eps Ed / Dmin:
//   |e_i - E_i(p)| <= 4 eps mag_i  ==>  Z(e) >= R(p) - (q/3) S,  S = spread of {z_i / w_i} and R   (weights lemma:
//                                       Z(e) - R(p) = sum_i (e_i - E_i) w_i (z_i/w_i - R) / sum_i e_i w_i)
//   corner values computed from the affine coefficients in fp32:  |R~(c) - R(c)| <= q (dabs + |R~|)
//   roundings of K.cpp:384-397 (three divisions, two dot products, one division): <= 10 eps dabs
//   the approximate divisions used here: 2^-21 relative.
// Everything is charged twice below, and the bound is only used when q <= 1/64 and D > 0 at all four corners
// (then D > 0 on the whole rectangle even with its own evaluation error).
__device__ __forceinline__ float block_depth_bound(const float ea[3], const float eb[3], const float ec[3],
                                                   const float mag[3], const float zc[3], const float wc[3],
                                                   float px0, float px1, float py0, float py1) {
  const float an = __fmaf_rn(zc[2], ea[2], __fmaf_rn(zc[1], ea[1], zc[0] * ea[0]));
  const float bn = __fmaf_rn(zc[2], eb[2], __fmaf_rn(zc[1], eb[1], zc[0] * eb[0]));
  const float cn = __fmaf_rn(zc[2], ec[2], __fmaf_rn(zc[1], ec[1], zc[0] * ec[0]));
  const float ad = __fmaf_rn(wc[2], ea[2], __fmaf_rn(wc[1], ea[1], wc[0] * ea[0]));
  const float bd = __fmaf_rn(wc[2], eb[2], __fmaf_rn(wc[1], eb[1], wc[0] * eb[0]));
  const float cd = __fmaf_rn(wc[2], ec[2], __fmaf_rn(wc[1], ec[1], wc[0] * ec[0]));
  const float ny0 = __fmaf_rn(bn, py0, cn), ny1 = __fmaf_rn(bn, py1, cn);
  const float dy0 = __fmaf_rn(bd, py0, cd), dy1 = __fmaf_rn(bd, py1, cd);
  const float n00 = __fmaf_rn(an, px0, ny0), n10 = __fmaf_rn(an, px1, ny0), n01 = __fmaf_rn(an, px0, ny1), n11 = __fmaf_rn(an, px1, ny1);
  const float d00 = __fmaf_rn(ad, px0, dy0), d10 = __fmaf_rn(ad, px1, dy0), d01 = __fmaf_rn(ad, px0, dy1), d11 = __fmaf_rn(ad, px1, dy1);
  const float dmin = fminf(fminf(d00, d10), fminf(d01, d11));
  const float ed = __fmaf_rn(wc[2], mag[2], __fmaf_rn(wc[1], mag[1], wc[0] * mag[0]));
  // q = 16 eps Ed / Dmin <= 1/64  <=>  Ed * 2^-14 <= Dmin   (also rejects Dmin <= 0 and NaN)
  if (!(ed * 6.1035156e-5f <= dmin)) return -INFINITY;
  const float q = __fdividef(ed * 9.5367432e-7f, dmin);
  const float r00 = __fdividef(n00, d00), r10 = __fdividef(n10, d10), r01 = __fdividef(n01, d01), r11 = __fdividef(n11, d11);
  const float rlow = fminf(fminf(r00, r10), fminf(r01, r11));
  const float rabs = fmaxf(fmaxf(fabsf(r00), fabsf(r10)), fmaxf(fabsf(r01), fabsf(r11)));
  const float dabs = fmaxf(fmaxf(fabsf(__fdividef(zc[0], wc[0])), fabsf(__fdividef(zc[1], wc[1]))),
                           fabsf(__fdividef(zc[2], wc[2])));
  const float scale = (dabs + rabs) * 1.0001f;
  // S <= 2 (dabs + rabs):  2 * [ (q/3) S + q (dabs + rabs) ] <= 4 q scale;  roundings: 2^-19 scale
  return rlow - scale * (4.0f * q + 1.9073486e-6f) - 1e-30f;
}

// Packed pixel box vs the pixel rectangle [x0, x0 + w) x [y0, y0 + h).
__device__ __forceinline__ bool box_touches_rect(uint2 packed, int x0, int y0, int w, int h) {
  const int4 bx = unpack_box(packed);
  return bx.x < x0 + w && bx.y > x0 && bx.z < y0 + h && bx.w > y0;
}

constexpr int kMacroTiles = 4;               // a macro tile is up to 4 x 4 screen tiles = 64 x 64 pixels
constexpr int kMacroCap = kMacroCapacity;    // candidates a macro tile holds per pass
constexpr int kScanRound = 2 * kChunk;       // candidates examined per scan round (two per thread)
constexpr int kTinyMeshMax = kMacroCap;      // largest mesh the tiny-mesh mode accepts

// Persistent kernel.  Work items are (image, macro tile) pairs of the images that have candidates; CTA i
// takes items i, i + gridDim.x, ...  Per item the image's candidate list is scanned ONCE against the
// macro tile (64x64 pixels) into shared memory; each of its 16 screen tiles then filters that short
// list against its own rectangle, stages the survivors 256 at a time and rasterizes them.  Nothing here
// needs list sizes on the host: memory is bounded by the per-image list (T entries) and, when a macro
// tile has more than kMacroCap candidates, the item is done in several passes that merge through the
// global depth keys.
//   large_count == nullptr   tiny-mesh mode (T <= kTinyMeshMax): every tile walks all T triangles, boxes
//                            are computed here and the tile writes ids / barycentrics / z / image itself
//                            (keys_out == nullptr);
//   large_count != nullptr   pipeline mode: image b has large_count[b] large triangles, ids and packed
//                            boxes in large_ids / large_boxes[b * T ...]; the result is merged into keys_out.
template <int A_STATIC, bool PIPELINE>
__global__ void __launch_bounds__(kChunk, 4)
raster_tile_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris,
                   int B, int V, int T, int W, int H, float half_w, float half_h, int tiles_x, int tiles_y,
                   int macro_shift, int *__restrict__ work_cursor, const int *__restrict__ large_count, const int32_t *__restrict__ large_ids,
                   const uint2 *__restrict__ large_boxes,
                   int32_t *__restrict__ out_ids, float *__restrict__ out_bary, float *__restrict__ out_z,
                   const float *__restrict__ attrs, const float *__restrict__ background, int A_dyn,
                   float *__restrict__ out_image, unsigned long long *__restrict__ keys_out) {
  __shared__ __align__(16) TileSmem sm;
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lanes_below = (1u << lane) - 1u;
  const int macro_tiles = 1 << macro_shift;               // 1, 2 or 4 screen tiles on a side
  const int macros_x = (tiles_x + macro_tiles - 1) >> macro_shift, macros_y = (tiles_y + macro_tiles - 1) >> macro_shift;
  const int macros_per_image = macros_x * macros_y;
  // Work items are handed out through a global counter (item costs differ by an order of magnitude
  // between busy and empty screen regions); `grabbed` only ever grows, the slices consume it in order.
  auto grab_item = [&]() {
    __syncthreads();
    if (threadIdx.x == 0) sm.next_item = atomicAdd(work_cursor, 1);
    __syncthreads();
    return sm.next_item;
  };
  int grabbed = grab_item();
  int slice_begin = 0;
  constexpr bool pipeline = PIPELINE;                     // compile-time: the tiny-mesh epilogue and the
                                                          // barycentrics of the running winner drop out
  // 8 warps, each an 8x4 pixel block; blocks are laid out 2 across, 4 down inside the 16x16 tile.
  const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);

  for (int b0 = 0; b0 < B; b0 += kChunk) {
    // ---- images of this slice that have candidates, in image order (every CTA builds the same list)
    __syncthreads();
    const bool has_work = b0 + (int)threadIdx.x < B && (!pipeline || large_count[b0 + threadIdx.x] > 0);
    const unsigned work_votes = __ballot_sync(0xffffffffu, has_work);
    if (lane == 0) sm.warp_count[warp] = __popc(work_votes);
    __syncthreads();
    int n_active = 0, my_at = 0;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) { if (k == warp) my_at = n_active; n_active += sm.warp_count[k]; }
    if (has_work) sm.active[my_at + __popc(work_votes & lanes_below)] = (unsigned short)(b0 + threadIdx.x);
    __syncthreads();

  const int slice_items = n_active * macros_per_image;
  for (; grabbed < slice_begin + slice_items; grabbed = grab_item()) {
    const int item = grabbed - slice_begin;
    const int ai = item / macros_per_image, macro = item - ai * macros_per_image;
    const int b = sm.active[ai];
    const int macro_ty = macro / macros_x, macro_tx = macro - macro_ty * macros_x;
    const int macro_x0 = macro_tx * macro_tiles * kTileW, macro_y0 = macro_ty * macro_tiles * kTileH;
    const int n_cand = pipeline ? large_count[b] : T;
    const float *verts_b = verts + (size_t)b * V * 4;
    const int32_t *cand_ids = pipeline ? large_ids + (size_t)b * T : nullptr;
    const uint2 *cand_boxes = pipeline ? large_boxes + (size_t)b * T : nullptr;

  int next = 0;              // candidates of the image scanned so far (uniform)
  do {
    // ---- macro list: candidates whose box touches the macro tile
    __syncthreads();         // previous pass / item has finished with the macro list
    if (threadIdx.x == 0) sm.n_macro = 0;
    __syncthreads();
    int n_macro = 0;
    while (next < n_cand && n_macro + kScanRound <= kMacroCap) {
      const int c0 = next + threadIdx.x, c1 = c0 + kChunk;
      uint2 box0 = make_uint2(0u, 0u), box1 = box0;
      bool keep0 = c0 < n_cand, keep1 = c1 < n_cand;
      if (pipeline) {
        if (keep0) box0 = __ldg(cand_boxes + c0);
        if (keep1) box1 = __ldg(cand_boxes + c1);
        keep0 = keep0 && box_touches_rect(box0, macro_x0, macro_y0, macro_tiles * kTileW, macro_tiles * kTileH);
        keep1 = keep1 && box_touches_rect(box1, macro_x0, macro_y0, macro_tiles * kTileW, macro_tiles * kTileH);
      }
      const unsigned v0 = __ballot_sync(0xffffffffu, keep0), v1 = __ballot_sync(0xffffffffu, keep1);
      int at = 0;
      if (lane == 0 && (v0 | v1)) at = atomicAdd(&sm.n_macro, __popc(v0) + __popc(v1));
      at = __shfl_sync(0xffffffffu, at, 0);
      if (keep0) { const int q = at + __popc(v0 & lanes_below); sm.mslot[q] = c0; sm.mbox[q] = box0; }
      if (keep1) { const int q = at + __popc(v0) + __popc(v1 & lanes_below); sm.mslot[q] = c1; sm.mbox[q] = box1; }
      next += kScanRound;
      n_macro += __syncthreads_count(keep0);               // barriers; the same totals in every thread
      n_macro += __syncthreads_count(keep1);
    }
    if (n_macro == 0 && pipeline) continue;                // nothing of this range reaches the macro tile

  for (int sub = 0; sub < macro_tiles * macro_tiles; ++sub) {
    const int tile_tx = macro_tx * macro_tiles + (sub & (macro_tiles - 1)), tile_ty = macro_ty * macro_tiles + (sub >> macro_shift);
    if (tile_tx >= tiles_x || tile_ty >= tiles_y) continue;
    const int tile_x0 = tile_tx * kTileW, tile_y0 = tile_ty * kTileH;
    const int blk_x0 = tile_x0 + (warp & 1) * 8, blk_y0 = tile_y0 + (warp >> 1) * 4;
    const int ix = tile_x0 + lx, iy = tile_y0 + ly;

    __syncthreads();         // the previous tile's epilogue has finished with the shared arrays
    sm.key[threadIdx.x] = kEmptyKey;
    if (threadIdx.x < kTileW) sm.cx[threadIdx.x] = pixel_center(tile_x0 + threadIdx.x, half_w);
    else if (threadIdx.x < kTileW + kTileH) sm.cy[threadIdx.x - kTileW] = pixel_center(tile_y0 + threadIdx.x - kTileW, half_h);
    if (threadIdx.x == 0) { sm.n_big = 0; sm.n_cand = 0; }
    if (threadIdx.x < 32) { sm.depth_bucket[threadIdx.x] = 0; sm.depth_cursor[threadIdx.x] = 0; }

    Fragment best;
    fragment_clear(best);
    __syncthreads();         // publishes key / cx / cy / counters
    const float px = sm.cx[lx], py = sm.cy[ly];
    // pixel-centre range of this warp's 8x4 block (for the conservative edge test of the big path)
    const float blk_px0 = sm.cx[(warp & 1) * 8], blk_px1 = sm.cx[(warp & 1) * 8 + 7];
    const float blk_py0 = sm.cy[(warp >> 1) * 4], blk_py1 = sm.cy[(warp >> 1) * 4 + 3];
    const float blk_pxabs = fmaxf(fabsf(blk_px0), fabsf(blk_px1)), blk_pyabs = fmaxf(fabsf(blk_py0), fabsf(blk_py1));
    const float tile_pxabs = fmaxf(fabsf(sm.cx[0]), fabsf(sm.cx[kTileW - 1])), tile_pyabs = fmaxf(fabsf(sm.cy[0]), fabsf(sm.cy[kTileH - 1]));
    if (PIPELINE && ix < W && iy < H) {
      // pipeline mode: start from what is already drawn (small triangles, earlier passes); depth and id
      // are all the depth rule needs, barycentrics are not produced in this mode
      const unsigned long long seen = keys_out[((size_t)b * H + iy) * W + ix];
      if (seen != kEmptyKey) { best.z = ordered_to_float((unsigned)(seen >> 32)); best.id = depth_key_id(seen); }
    }

    int mnext = 0;           // macro-list entries examined so far (uniform)
    int pending = 0;         // survivors waiting in sm.cand (uniform)
    bool first_chunk = true;
    while (mnext < n_macro || pending > 0) {
      // ---- filter: box-vs-tile test on 256 macro entries per round until a chunk is full (sm.cand holds
      // up to 2 * kChunk entries: a round adds at most kChunk to fewer than kChunk)
      while (pending < kChunk && mnext < n_macro) {
        const int c = mnext + threadIdx.x;
        bool keep = c < n_macro;
        if (keep && pipeline) keep = box_touches_rect(sm.mbox[c], tile_x0, tile_y0, kTileW, kTileH);
        const unsigned votes = __ballot_sync(0xffffffffu, keep);
        int at = 0;
        if (lane == 0 && votes) at = atomicAdd(&sm.n_cand, __popc(votes));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (keep) sm.cand[at + __popc(votes & lanes_below)] = (unsigned short)c;
        mnext += kChunk;
        pending += __syncthreads_count(keep);             // barrier; the same total in every thread
      }

      const int n_here = min(kChunk, pending);
      if (n_here == 0) break;
      if (!first_chunk) {
        if (threadIdx.x == 0) sm.n_big = 0;
        if (threadIdx.x < 32) { sm.depth_bucket[threadIdx.x] = 0; sm.depth_cursor[threadIdx.x] = 0; }
        __syncthreads();
      }
      first_chunk = false;

      // ---- stage: the chunk's triangles are dealt round-robin to the warps (entry lane*8 + warp goes
      // to this thread) so that every warp owns a similar share of the small-triangle work.
      const int entry = lane * kWarps + warp;
      int x0 = 0, x1 = 0, y0 = 0, y1 = 0, n_seg = 0;
      bool overlaps = false;
      if (entry < n_here) {
        const int mc = sm.cand[entry];
        const int t = pipeline ? __ldg(cand_ids + sm.mslot[mc]) : sm.mslot[mc];
        float4 p0, p1, p2;
        load_triangle(verts_b, tris, t, p0, p1, p2);
        int4 bx;
        if (pipeline) {
          bx = unpack_box(sm.mbox[mc]);
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
        // the box inside this tile, in tile-local pixel coordinates
        x0 = max(bx.x, tile_x0) - tile_x0; x1 = min(bx.y, tile_x0 + kTileW) - tile_x0;
        y0 = max(bx.z, tile_y0) - tile_y0; y1 = min(bx.w, tile_y0 + kTileH) - tile_y0;
        overlaps = x1 > x0 && y1 > y0;
        if (overlaps && zlo > -INFINITY) {
          // Tighten the bound to THIS tile (block_depth_bound over the tile's pixel centres): a large slanted
          // triangle is then sorted, and culled, by where it is here, not by its nearest vertex.
          const float ea[3] = {m[0], m[3], m[6]}, eb[3] = {m[1], m[4], m[7]}, ec[3] = {m[2], m[5], m[8]};
          const float zc[3] = {p0.z, p1.z, p2.z}, wc[3] = {p0.w, p1.w, p2.w};
          float mag[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) mag[i] = fabsf(ea[i]) * tile_pxabs + fabsf(eb[i]) * tile_pyabs + fabsf(ec[i]);
          zlo = fmaxf(zlo, block_depth_bound(ea, eb, ec, mag, zc, wc, sm.cx[0], sm.cx[kTileW - 1], sm.cy[0], sm.cy[kTileH - 1]));
        }
        sm.zlo[threadIdx.x] = zlo;
        if (overlaps && (x1 - x0) * (y1 - y0) <= 64) n_seg = (y1 - y0) * ((x1 - x0 + 3) >> 2);
      }
      // ---- the warp lays the row segments of ITS small triangles out back to back (warp scan).
      // Triangles whose segments do not fit the warp's buffer, and all large ones, take the big path.
      // The segment path pays when a good part of the warp has such triangles (a mesh of small triangles);
      // a few stragglers (corners of large triangles clipped by the tile) would run its loops with one
      // or two lanes active, so they take the big path, which draws anything.
      if (__popc(__ballot_sync(0xffffffffu, n_seg > 0)) < 8) n_seg = 0;
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
            // (Evaluating block_depth_bound once more here, for the warp's own 8x4 block, was measured on c5:
            // with the tile-level bound already in zlo it culls too little to pay -- 9.65 ms without, 10.0 with.)
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
      __syncthreads();         // the chunk's records and candidate entries are consumed
      // survivors beyond this chunk move to the front of the filter list
      const int left = pending - n_here;
      unsigned short moved = 0;
      if ((int)threadIdx.x < left) moved = sm.cand[kChunk + threadIdx.x];
      __syncthreads();
      if ((int)threadIdx.x < left) sm.cand[threadIdx.x] = moved;
      if (threadIdx.x == 0) sm.n_cand = left;
      pending = left;
      __syncthreads();
    }

    // ---- resolve: minimum of the two paths (all keys final; record planes dead from here on)
    const unsigned long long key_small = sm.key[ly * kTileW + lx];
    const unsigned long long key_big = best.id >= 0 ? depth_key(best.z, best.id) : kEmptyKey;
    if constexpr (PIPELINE) {
      // large-triangle pass of the pipeline: merge into the global keys (this CTA is the only
      // writer of its pixels now; the scatter kernel has finished), resolve_kernel does the rest.
      if (ix < W && iy < H) {
        const size_t p = ((size_t)b * H + iy) * W + ix;
        const unsigned long long mine = min(key_small, key_big);
        if (mine < keys_out[p]) keys_out[p] = mine;
      }
    } else {
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
  }                          // tiles of the macro tile
  } while (next < n_cand);   // passes over the image's candidates
  }                          // work items
  slice_begin += slice_items;
  }                          // image slices
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

constexpr int kScatterWarps = 2;     // warps are independent; small CTAs refill SM slots sooner (8: 0.397, 4: 0.387, 2: 0.380 ms on c2)
constexpr int kScatterSegCap = 512;          // row segments a warp holds per round (one triangle has <= 64)
constexpr int kSmallBox = 16;                // largest box side the scatter path takes

struct ScatterWarpSmem {
  float4 r0[32], r1[32], r2[32], r3[32];     // setup records of the warp's 32 triangles (planes as in TileSmem)
  int2 origin[32];                           // left, bottom of each triangle's pixel box
  unsigned short segs[kScatterSegCap];       // slot | dy << 5 | dx << 9 | width << 13
  unsigned char spans[kSmallBox][32];        // per box row and triangle: first candidate column | (columns - 1) << 4, 0xff: none
  unsigned short hits[160];                  // queue of inside pixels: slot | dy << 5 | x offset << 9
};

// (48 registers, 5 CTAs per SM; forcing 6 CTAs -- 40 registers, small spills -- measured 2 % slower.)
__global__ void __launch_bounds__(kScatterWarps * 32)
scatter_small_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris, int V, int T, int W, int H,
                     float half_w, float half_h, const float *__restrict__ centers,
                     int *__restrict__ large_count, int32_t *__restrict__ large_ids,
                     uint2 *__restrict__ large_boxes, unsigned long long *__restrict__ keys) {
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
  bool is_large = false;
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
        // Candidate columns of every box row.  A pixel passes the inside test only if all three edge values,
        // evaluated in fp32 as K.cpp:93-98 does, are >= 0; the fp32 value of edge i differs from the exact
        // a*x + b*y + c by at most 2^-22 * (|a| + |b| + |c|) =: d (pixel centres lie in [-1, 1]).  Along a row
        // the exact value is linear in x, so the pixels that can pass lie on one side of the crossing
        // x = -(b*y + c) / a, shifted by d / |a|.  An edge bounds the row only when that shift (and the rounding
        // of this very computation, of the same size) is below 1/64 of a pixel -- |a| >= 2^-16 * W/2 * (|a| +
        // |b| + |c|) -- and the span is widened by a quarter pixel on both sides; nearly horizontal edges bound
        // nothing.  On c2 (boxes of 6.3 x 6.3 pixels around triangles of 7.7) this halves the row segments.
        // (Taking the rows of boxes of <= 4 columns whole -- one segment per row whatever the span -- was
        // measured: c3 0.731 -> 0.717 ms, but c2 0.326 -> 0.332 and c4 1.695 -> 1.741: empty rows matter more.)
        float cross[3];                        // -1 / a of the edges that bound rows, else 0
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float mag = fabsf(m[3 * i]) + fabsf(m[3 * i + 1]) + fabsf(m[3 * i + 2]);
          cross[i] = fabsf(m[3 * i]) >= 1.52587890625e-5f * half_w * mag ? -1.0f / m[3 * i] : 0.0f;
        }
        const float left_f = (float)box.left, last_f = (float)(bw - 1);
        for (int dy = 0; dy < bh; ++dy) {
          const float y = __ldg(cy + box.bottom + dy);
          float x_lo = -2.0f, x_hi = 2.0f;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float x = (m[3 * i + 1] * y + m[3 * i + 2]) * cross[i];      // NaN (0 * inf): ignored by fmaxf / fminf
            if (cross[i] < 0.0f) x_lo = fmaxf(x_lo, x);                        // a > 0: inside to the right
            if (cross[i] > 0.0f) x_hi = fminf(x_hi, x);
          }
          // pixel k has its centre at (k + 0.5) / half_w - 1
          const float k_lo = ceilf((x_lo + 1.0f) * half_w - 0.75f), k_hi = floorf((x_hi + 1.0f) * half_w - 0.25f);
          const int d_lo = (int)fminf(fmaxf(k_lo - left_f, 0.0f), 16.0f);
          const int d_hi = (int)fmaxf(fminf(k_hi - left_f, last_f), -1.0f);
          const int columns = d_hi - d_lo + 1;
          sm.spans[dy][lane] = columns > 0 ? (unsigned char)(d_lo | ((columns - 1) << 4)) : (unsigned char)0xff;
          if (columns > 0) n_seg += (columns + 3) >> 2;
        }
      } else {
        big_box = pack_box(box);
        is_large = true;
      }
    }
  }
  // large triangles go to the image's large list (raster_tile_kernel draws them); one atomic per warp
  const unsigned large_votes = __ballot_sync(0xffffffffu, is_large);
  if (large_votes) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(large_count + b, __popc(large_votes));
    slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(large_votes & ((1u << lane) - 1u));
    if (is_large) {
      large_ids[(size_t)b * T + slot] = t;
      large_boxes[(size_t)b * T + slot] = big_box;
    }
  }

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
        atomicMin(keys_b + (unsigned)(y * W + x), depth_key(z, __float_as_int(q0.w)));   // H * W <= 2^30
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
      for (int dy = 0; dy < bh; ++dy) {
        const unsigned span = sm.spans[dy][lane];
        if (span == 0xffu) continue;
        const unsigned head = lane | (dy << 5);
        const int first = span & 15u, columns = (span >> 4) + 1, per_row = (columns + 3) >> 2;   // 1..4 segments
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < per_row) sm.segs[k + q] = (unsigned short)(head | ((first + 4 * q) << 9) | (min(4, columns - 4 * q) << 13));
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
        const int yi = org.y + dy, xi = org.x + dx;      // 32-bit indices: one IMAD.WIDE per table
        const float cyv = __ldg(centers + (W + yi));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < width) {
            float e[3], esum;
            edge_values(m, __ldg(centers + (xi + k)), cyv, e);
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
constexpr int kResolveStrip = 4;             // most blocks of 8 x 4 pixels a warp resolves, left to right (fewer for small batches)

// One warp per strip of kResolveStrip blocks, 2 x 4 strips per CTA (64 x 16 pixels).  With tensor maps the strip's
// depth keys (32 x 4 pixels, 1 KB) arrive by ONE bulk tensor copy issued when the warp starts: a warp that
// resolves a single block spends a fifth of its life waiting for its first load (ncu, profiles/r02: 20 % of the
// stall samples on the key), here that wait is paid once per strip and the tensor stores of block i drain while
// block i + 1 is computed.  40 registers, 6 CTAs per SM: at 8 CTAs (32 registers) the loop state spills and the
// kernel takes 0.385 ms instead of 0.279 (c2); 5 / 4 CTAs: 0.296 / 0.312.  Against one block per warp (0.286 ms,
// 202 M instructions) the strips cost 35 M instructions and win 2 % on c2 / c4, lose 4 % on c5.
// SHADE: the render path (render.py:198-228 without specular colours).  The nine interpolated channels
// [normal, world position, diffuse colour] of a pixel never leave the registers: they are lit right here
// (shade_math.cuh) and only RGBA is written, rows flipped as phong_shader returns them (render.py:382-386).
template <int A_STATIC, bool SHADE, int MIN_CTAS = (SHADE ? 5 : 6)>
__global__ void __launch_bounds__(kResolveWarps * 32, MIN_CTAS)
resolve_kernel(const __grid_constant__ OutputMaps maps, int use_tma, int strip_blocks,
               const float *__restrict__ verts, const int32_t *__restrict__ tris, int V, int W, int H,
               const float *__restrict__ centers,
               const unsigned long long *__restrict__ keys,
               int32_t *__restrict__ out_ids, float *__restrict__ out_bary, float *__restrict__ out_z,
               const float *__restrict__ attrs, const float *__restrict__ background, int A_dyn,
               float *__restrict__ out_image, const float *__restrict__ light_positions,
               const float *__restrict__ light_intensities, const float *__restrict__ ambient, int L,
               float4 *__restrict__ out_rgba) {
  __shared__ __align__(128) float stage_all[kResolveWarps][32 * 16];
  __shared__ __align__(128) unsigned long long keys_all[kResolveWarps][4 * 8 * kResolveStrip];
  __shared__ unsigned long long ready_all[kResolveWarps];
  __shared__ Lights lights;
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  if (SHADE) load_lights(lights, light_positions, light_intensities, ambient, b, L);     // block barrier inside
  const int x0 = (blockIdx.x * 2 + (warp & 1)) * (8 * strip_blocks), y0 = (blockIdx.y * (kResolveWarps / 2) + (warp >> 1)) * 4;
  if (x0 >= W || y0 >= H) return;
  const int n_blocks = min(strip_blocks, (W - x0 + 7) >> 3);
  const int iy = y0 + (lane >> 3);
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&ready_all[warp]);
  if (use_tma) {
    if (elect_one()) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(256 * strip_blocks)) : "memory");
      tma_load_2d((unsigned)__cvta_generic_to_shared(keys_all[warp]), &maps.keys, 2 * x0, b * H + y0, bar);
    }
    __syncwarp();
  }
  const unsigned tid_kept = threadIdx.x;
  for (int it = 0; it < n_blocks; ++it) {
    // what depends on the thread only is re-derived from an opaque copy of its index: kept in registers across
    // the loop it pushes the per-pixel arithmetic into spills
    // (the CTA's coordinates too: 24 bytes of spills at 40 registers otherwise)
    unsigned tid = tid_kept, b_now, cta_x, cta_y;
    asm volatile("" : "+r"(tid));
    asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(b_now));
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta_x));
    asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(cta_y));
    const int b = (int)b_now;
    const int lane = tid & 31, warp = tid >> 5;
    const int x0 = (cta_x * 2 + (warp & 1)) * (8 * strip_blocks), y0 = (cta_y * (kResolveWarps / 2) + (warp >> 1)) * 4;
    const int iy = y0 + (lane >> 3);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&ready_all[warp]);
    const int blk_x0 = x0 + 8 * it, ix = blk_x0 + (lane & 7);
    const bool in_image = ix < W && iy < H;
    unsigned long long key = kEmptyKey;
    if (use_tma) {
      if (it == 0) {
        unsigned done = 0;
        while (!done) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
      }
      if (in_image) key = keys_all[warp][(lane >> 3) * (8 * strip_blocks) + it * 8 + (lane & 7)];   // (columns beyond W: zero fill)
    } else if (in_image) {
      key = keys[((size_t)b * H + iy) * W + ix];
    }
    Fragment best;
    fragment_clear(best);
    if (key != kEmptyKey) {
      const int t = depth_key_id(key);
      float4 p0, p1, p2;
      load_triangle(verts + (size_t)b * V * 4, tris, t, p0, p1, p2);
      evaluate_winner(p0, p1, p2, __ldg(centers + ix), __ldg(centers + W + iy), t, best);
    }
    // (Gathering the 3*A corner attributes here, together with the vertices, was measured: 60 registers
    // instead of 38 cost more occupancy than the shorter dependency chain gained: 0.355 -> 0.411 ms.)
    if (use_tma && it > 0) {               // the previous block's tensor stores have read the stage
      if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
    block_epilogue<SHADE ? 0 : A_STATIC>(stage_all[warp], b, blk_x0, y0, W, H, V, best, tris, attrs, background, A,
                                         out_ids, out_bary, out_z, SHADE ? nullptr : out_image, nullptr,
                                         use_tma ? &maps : nullptr, true);
    if (SHADE && in_image) {
      // the same interpolation as block_epilogue (rast.py:118-150), into registers
      float px[9];
      if (best.id < 0) {
#pragma unroll
        for (int a = 0; a < 9; ++a) px[a] = __ldg(background + a);
      } else {
        const float alpha = coverage_alpha(best.b0, best.b1, best.b2);
        const float one_minus = 1.0f - alpha;
        const float *at = attrs + (size_t)b * V * 9;
        const float *c0 = at + (size_t)__ldg(tris + 3 * (size_t)best.id + 0) * 9;
        const float *c1 = at + (size_t)__ldg(tris + 3 * (size_t)best.id + 1) * 9;
        const float *c2 = at + (size_t)__ldg(tris + 3 * (size_t)best.id + 2) * 9;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const float img = __ldg(c0 + a) * best.b0 + __ldg(c1 + a) * best.b1 + __ldg(c2 + a) * best.b2;
          px[a] = alpha * img + one_minus * __ldg(background + a);
        }
      }
      out_rgba[((size_t)b * H + (H - 1 - iy)) * W + ix] = shade_diffuse_pixel(px, px + 3, px + 6, lights, L, ambient != nullptr);
    }
    if (!use_tma) __syncwarp();            // the stage is rewritten by the next block
  }
  if (use_tma && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
                         int W, int H, int *work_cursor, const int *large_count, const int32_t *large_ids, const uint2 *large_boxes,
                         int32_t *ids, float *bary, float *z, const float *attrs, const float *bg, int A,
                         float *image, unsigned long long *keys_out, cudaStream_t stream) {
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const float half_w = (float)(0.5 * W), half_h = (float)(0.5 * H);   // K.cpp:309-310
  // persistent CTAs: 4 per SM (launch bounds), fewer when there is less work than that.  Macro tiles of
  // 4x4 screen tiles amortise the candidate scan 16-fold; small jobs use smaller ones to keep every SM busy.
  const long long resident = (long long)ctx->sm_count * 4;
  static_assert(kMacroTiles == 4, "macro_shift below starts at log2(kMacroTiles)");
  int macro_shift = 2;
  long long n_items = 0;
  for (;; --macro_shift) {
    const int macro_tiles = 1 << macro_shift;
    n_items = (long long)B * ((tiles_x + macro_tiles - 1) / macro_tiles) * ((tiles_y + macro_tiles - 1) / macro_tiles);
    if (macro_shift == 0 || n_items >= 4 * resident) break;
  }
  if (n_items > (long long)INT_MAX / 2) return set_error(ctx, PMR_ERR_SIZE, "too many screen tiles");
  const int grid = (int)(n_items < resident ? n_items : resident);
  StageScope timed(ctx, PMR_STAGE_RASTER, stream);
#define PMR_LAUNCH(AS, PIPE)                                                                        \
  raster_tile_kernel<AS, PIPE><<<grid, kChunk, 0, stream>>>(verts, tris, B, V, T, W, H, half_w, half_h, \
                                                     tiles_x, tiles_y, macro_shift, work_cursor,    \
                                                     large_count, large_ids,                        \
                                                     large_boxes, ids, bary, z, attrs, bg, A, image, keys_out)
  if (keys_out != nullptr) PMR_LAUNCH(0, true);
  else if (image == nullptr) PMR_LAUNCH(0, false);
  else if (A == 4) PMR_LAUNCH(4, false);
  else if (A == 9) PMR_LAUNCH(9, false);
  else if (A == 12) PMR_LAUNCH(12, false);
  else if (A == 13) PMR_LAUNCH(13, false);
  else PMR_LAUNCH(0, false);
#undef PMR_LAUNCH
  ctx->launches += 1;
  return check_launch(ctx, "raster_tile_kernel");
}

int forward_impl(Context *ctx, const float *verts, const int32_t *tris, int B, int V, int T, int W, int H,
                 int32_t *ids, float *bary, float *z, const float *attrs, const float *bg, int A,
                 float *image, cudaStream_t stream, const ShadeArgs *shade) {
  if (B == 0 || W == 0 || H == 0) return PMR_OK;
  const float half_w = (float)(0.5 * W), half_h = (float)(0.5 * H);

  if (shade == nullptr && T <= ctx->small_mesh_threshold && T <= kTinyMeshMax) {
    // tiny mesh: one kernel, every tile walks all triangles and writes the outputs itself
    int rc0 = ctx->bins.reserve(ctx, 16);
    if (rc0) return rc0;
    PMR_CUDA(ctx, cudaMemsetAsync(ctx->bins.ptr, 0, 16, stream));       // the work counter
    return launch_raster(ctx, verts, tris, B, V, T, W, H, (int *)ctx->bins.ptr, nullptr, nullptr, nullptr, ids, bary, z,
                         attrs, bg, A, image, nullptr, stream);
  }

  const size_t n_pixels = (size_t)B * H * W;
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
  // workspace: [work cursor | large_count[B] | pad to 16 | large_boxes[B*T] (8 B) | large_ids[B*T]], keys
  const size_t count_bytes = (((size_t)(B + 1) * sizeof(int)) + 15) & ~(size_t)15;
  rc = ctx->bins.reserve(ctx, count_bytes + (size_t)B * T * (sizeof(uint2) + sizeof(int32_t)));
  if (rc) return rc;
  rc = ctx->keys.reserve(ctx, n_pixels * sizeof(unsigned long long));
  if (rc) return rc;
  char *base = (char *)ctx->bins.ptr;
  int *work_cursor = (int *)base;
  int *large_count = work_cursor + 1;
  uint2 *large_boxes = (uint2 *)(base + count_bytes);
  int32_t *large_ids = (int32_t *)(large_boxes + (size_t)B * T);
  unsigned long long *keys = (unsigned long long *)ctx->keys.ptr;
  ctx->last_large_count = large_count;
  ctx->last_large_images = B;

  {
    StageScope timed(ctx, PMR_STAGE_BIN, stream);
    PMR_CUDA(ctx, cudaMemsetAsync(work_cursor, 0, (size_t)(B + 1) * sizeof(int), stream));
    // (Letting the resolve kernel clear each key it reads, to save this memset, made that kernel 5x slower:
    // 0.31 -> 1.47 ms on c2, every load of the chain behind the store stalled.  Storing an empty box back over
    // each strip by one tensor store per warp, the buffer kept clean between calls: memset 0.025 -> 0.004 ms
    // but resolve 0.279 -> 0.293: 0.6 % of the step, not worth a buffer whose state outlives the call.)
    PMR_CUDA(ctx, cudaMemsetAsync(keys, 0xff, n_pixels * sizeof(unsigned long long), stream));   // kEmptyKey
  }
  {
    StageScope timed(ctx, PMR_STAGE_SCATTER, stream);
    scatter_small_kernel<<<dim3((T + kScatterWarps * 32 - 1) / (kScatterWarps * 32), B), kScatterWarps * 32, 0, stream>>>(
        verts, tris, V, T, W, H, half_w, half_h, centers, large_count, large_ids, large_boxes, keys);
    ctx->launches += 1;
    rc = check_launch(ctx, "scatter_small_kernel");
    if (rc) return rc;
  }
  // Large triangles: always launched (how many there are is only known on the device); CTAs of images
  // without large triangles fall through their tiles in a few hundred nanoseconds each.
  rc = launch_raster(ctx, verts, tris, B, V, T, W, H, work_cursor, large_count, large_ids, large_boxes,
                     ids, bary, z, attrs, bg, A, image, keys, stream);
  if (rc) return rc;
  {
    StageScope timed(ctx, PMR_STAGE_RESOLVE, stream);
    // strips of four blocks unless that leaves fewer than a dozen waves of CTAs (c4 at 32 views per GPU: 9): the
    // last, partly filled wave would show
    int strip = kResolveStrip;
    const long long cta_rows = (long long)((H + 2 * kResolveWarps - 1) / (2 * kResolveWarps)) * B;
    while (strip > 1 && cta_rows * ((W + 16 * strip - 1) / (16 * strip)) < 12LL * 6 * ctx->sm_count) strip >>= 1;
    if (ctx->strip_blocks_override > 0) strip = min(kResolveStrip, ctx->strip_blocks_override);   // tests of the strip loop
    dim3 grid((W + 16 * strip - 1) / (16 * strip), (H + 2 * kResolveWarps - 1) / (2 * kResolveWarps), B);
    // the staged outputs as tensor maps (PMR_NO_TMA=1 or an array the copy engine cannot describe: per-row copies)
    OutputMaps maps;
    const bool image_staged = shade == nullptr && image != nullptr && (A == 4 || A == 9 || A == 12 || A == 13);
    const int use_tma = !ctx->no_tma && (H & 3) == 0 &&
                        make_block_map(&maps.keys, CU_TENSOR_MAP_DATA_TYPE_UINT32, keys, 2LL * W, (long long)B * H, 16 * strip) &&
                        make_block_map(&maps.bary, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, bary, 3LL * W, (long long)B * H, 24) &&
                        (!image_staged || make_block_map(&maps.image, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, image,
                                                         (long long)A * W, (long long)B * H, 8 * A));
#define PMR_RESOLVE(AS)                                                                                          \
  resolve_kernel<AS, false><<<grid, kResolveWarps * 32, 0, stream>>>(maps, use_tma, strip, verts, tris, V, W, H, centers, \
                                                              keys, ids, bary, z, attrs, bg, A, image, nullptr,  \
                                                              nullptr, nullptr, 0, nullptr)
    if (shade != nullptr)
      resolve_kernel<9, true><<<grid, kResolveWarps * 32, 0, stream>>>(
          maps, use_tma, strip, verts, tris, V, W, H, centers, keys, ids, bary, z, attrs, bg, 9, nullptr, shade->light_positions,
          shade->light_intensities, shade->ambient, shade->L, reinterpret_cast<float4 *>(shade->rgba));
    else if (image == nullptr) PMR_RESOLVE(0);
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
