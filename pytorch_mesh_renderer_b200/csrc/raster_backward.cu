// raster_backward.cu -- backward of rasterize_triangles (K.cpp:131-273) batched over images,
// optionally fused with the backward of the attribute interpolation (autograd of
// rast.py:118-150), in two accumulation modes:
//
//   PMR_BACKWARD_ATOMIC   backward_blocks_kernel: one warp per 8x4 pixel block; the per-pixel sums are
//                         parked in shared memory, rows sorted by triangle, and lane c adds column c
//                         over each triangle's rows: one fire-and-forget atomic per (block, triangle,
//                         column).  fp32 sums in arbitrary order.  (Attribute counts without a
//                         specialised instance fall back to backward_atomic_kernel, one thread per
//                         pixel.)  With SHADE it is also the backward of the fused render path.
//   PMR_BACKWARD_ORDERED  one warp per (image, vertex): walks the union of the pixel boxes of the
//                         vertex's triangles in ascending pixel order, lanes evaluate the
//                         per-pixel terms in parallel and the sums are then folded strictly in
//                         pixel order, corner 0..2 within a pixel -- the reference's summation
//                         order (K.cpp:156-157, :232-269; index_put_ order of rast.py:130-132),
//                         so the result is bit-reproducible and equals the reference's.
//
// Per-pixel arithmetic is shared with the forward pass (raster_math.cuh) and follows the
// reference op for op; compile with -fmad=false.
#include <cuda.h>
#include <stddef.h>

#include "pmr_internal.cuh"
#include "raster_math.cuh"
#include "shade_math.cuh"

namespace pmr {

// Everything the backward pass needs to know about one covered pixel.
struct PixelGrad {
  int vid[3];        // vertex ids of the pixel's triangle
  float terms[9];    // vertex_terms(): [3*corner + component]
  float b[3];
  float alpha;
};

// Loads the pixel's triangle, derives d(loss)/d(bary) (given directly, or from the image gradient
// through the interpolation), and evaluates the nine vertex terms.
//   fused:  d_img_a = g_a * alpha;  d_b_k = sum_a d_img_a * corner_k[a]  (torch order).
template <bool FUSED>
__device__ __forceinline__ void pixel_grad_loaded(const int vid[3], const float4 &p0, const float4 &p1, const float4 &p2,
                                                  const float *__restrict__ attrs_b, const float *bary_p,
                                                  const float *g_p, int A, PixelGrad &out) {
#pragma unroll
  for (int j = 0; j < 3; ++j) out.vid[j] = vid[j];
  out.b[0] = bary_p[0]; out.b[1] = bary_p[1]; out.b[2] = bary_p[2];
  float g[3];
  if (FUSED) {
    const float s = 2.0f * out.b[0] + 2.0f * out.b[1] + 2.0f * out.b[2];
    out.alpha = fminf(fmaxf(s, 0.0f), 1.0f);
    const float alpha = out.alpha;
    const float *c[3] = {attrs_b + (size_t)out.vid[0] * A, attrs_b + (size_t)out.vid[1] * A,
                         attrs_b + (size_t)out.vid[2] * A};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float *ck = c[k];
      g[k] = torch_inner_sum(A, [&](int a) { return (g_p[a] * alpha) * __ldg(ck + a); });
    }
    // Covered pixels have s ~ 2: the clamp of rast.py:145-146 is saturated and passes no gradient,
    // so the 2*d_alpha term of the reference's autograd is exactly zero here.
  } else {
    out.alpha = 1.0f;
    g[0] = g_p[0]; g[1] = g_p[1]; g[2] = g_p[2];
  }
  float m[9];
  const float det = adjugate_signed(p0.x, p1.x, p2.x, p0.y, p1.y, p2.y, p0.w, p1.w, p2.w, m);
  vertex_terms(m, fabsf(det), out.b, g, out.terms);
}

template <bool FUSED>
__device__ __forceinline__ void pixel_grad(const float *__restrict__ verts_b, const float *__restrict__ attrs_b,
                                           const int32_t *__restrict__ tris, int id, const float *bary_p,
                                           const float *g_p, int A, PixelGrad &out) {
  int vid[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) vid[j] = __ldg(tris + 3 * (size_t)id + j);
  const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
  const float4 p0 = __ldg(v4 + vid[0]), p1 = __ldg(v4 + vid[1]), p2 = __ldg(v4 + vid[2]);
  pixel_grad_loaded<FUSED>(vid, p0, p1, p2, attrs_b, bary_p, g_p, A, out);
}

__device__ __forceinline__ bool pixel_is_covered(int id, const float *bary_p) {
  // K.cpp:162
  return !(id == 0 && bary_p[0] + bary_p[1] + bary_p[2] < kDegenerateBarySum);
}

// x, y, w gradients land in columns 0, 1, 3 of [V,4] (K.cpp:232-269).
__device__ __forceinline__ int column_of(int c) { return c == 2 ? 3 : c; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Atomic mode
// ---------------------------------------------------------------------------------------------

template <bool FUSED, int A_STATIC>
__global__ void __launch_bounds__(256)
backward_atomic_kernel(const float *__restrict__ grad, const float *__restrict__ verts,
                       const float *__restrict__ attrs, const int32_t *__restrict__ tris,
                       const int32_t *__restrict__ ids, const float *__restrict__ bary,
                       int V, int A_dyn, long long pixels_per_image, long long total_pixels,
                       float *__restrict__ d_verts, float *__restrict__ d_attrs) {
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool in_range = p < total_pixels;
  int id = -1;
  float bp[3] = {0.0f, 0.0f, 0.0f};
  if (in_range) {
    id = ids[p];
    bp[0] = bary[3 * p]; bp[1] = bary[3 * p + 1]; bp[2] = bary[3 * p + 2];
    if (!pixel_is_covered(id, bp)) id = -1;
  }
  const unsigned covered = __ballot_sync(0xffffffffu, id >= 0);
  if (covered == 0u) return;
  // All 32 pixels of a warp lie in one image only if the row stride allows; key on (image, id).
  const int b = in_range ? (int)(p / pixels_per_image) : 0;
  const long long key = id >= 0 ? ((long long)b << 32) | (unsigned)id : -1;
  const int leader = __ffs(covered) - 1;
  const long long leader_key = __shfl_sync(0xffffffffu, key, leader);
  const bool uniform = __all_sync(0xffffffffu, key == leader_key || key < 0);

  PixelGrad pg;
  const float *g_p = nullptr;
  if (id >= 0) {
    const float *verts_b = verts + (size_t)b * V * 4;
    const float *attrs_b = FUSED ? attrs + (size_t)b * V * A : nullptr;
    g_p = grad + (size_t)p * (FUSED ? A : 3);
    pixel_grad<FUSED>(verts_b, attrs_b, tris, id, bp, g_p, A, pg);
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) pg.terms[k] = 0.0f;
    pg.b[0] = pg.b[1] = pg.b[2] = 0.0f; pg.alpha = 0.0f;
    pg.vid[0] = pg.vid[1] = pg.vid[2] = 0;
  }

  if (uniform) {
    // One triangle for the whole warp: reduce, then one lane per value issues the atomic.
    const int v0 = __shfl_sync(0xffffffffu, pg.vid[0], leader);
    const int v1 = __shfl_sync(0xffffffffu, pg.vid[1], leader);
    const int v2 = __shfl_sync(0xffffffffu, pg.vid[2], leader);
    const int lb = __shfl_sync(0xffffffffu, b, leader);
    const int vv[3] = {v0, v1, v2};
    if (d_verts != nullptr) {
      float *dv = d_verts + (size_t)lb * V * 4;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float s = warp_sum(pg.terms[k]);
        if (lane == k) atomicAdd(dv + (size_t)vv[k / 3] * 4 + column_of(k % 3), s);
      }
    }
    if (FUSED && d_attrs != nullptr) {
      float *da = d_attrs + (size_t)lb * V * A;
      for (int a = 0; a < A; ++a) {
        const float d_img = id >= 0 ? g_p[a] * pg.alpha : 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float s = warp_sum(d_img * pg.b[k]);
          if (lane == ((3 * a + k) & 31)) atomicAdd(da + (size_t)vv[k] * A + a, s);
        }
      }
    }
    return;
  }
  if (id < 0) return;
  if (d_verts != nullptr) {
    float *dv = d_verts + (size_t)b * V * 4;
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(dv + (size_t)pg.vid[k / 3] * 4 + column_of(k % 3), pg.terms[k]);
  }
  if (FUSED && d_attrs != nullptr) {
    float *da = d_attrs + (size_t)b * V * A;
    for (int a = 0; a < A; ++a) {
      const float d_img = g_p[a] * pg.alpha;
#pragma unroll
      for (int k = 0; k < 3; ++k) atomicAdd(da + (size_t)pg.vid[k] * A + a, d_img * pg.b[k]);
    }
  }
}

// Predicated fire-and-forget float add (RED): lanes without a destination skip it without a branch.
__device__ __forceinline__ void red_add_if(bool ok, float *addr, float v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\t@p red.global.add.f32 [%1], %2;\n\t}"
               ::"r"((int)ok), "l"(addr), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Atomic mode, warp-aggregated: the throughput path
// ---------------------------------------------------------------------------------------------
//
// One warp owns a horizontal STRIP of 8x4 pixel blocks (the shape the forward kernel rasterizes) and
// walks it block by block, one lane per pixel.  The three per-pixel inputs of a block -- ids [4][8],
// barycentrics [4][8][3], gradient rows [4][8][A] -- are boxes of 2-D tensor maps over [B*H][W*c] and
// arrive in shared memory by TMA (cp.async.bulk.tensor.2d, one elected lane, completion on a per-warp
// mbarrier; SASS UTMALDG): the box of the NEXT block is requested as soon as the lanes have read the
// current one, so it lands while the current block is processed, no registers are held for it and the SM
// issues no address arithmetic, LDG or STS for it.  Everything that depends on the lane only (its role in
// the reduction, shared-memory addresses, barrier) is set up once per strip.
//
// Per block, every covered lane evaluates its per-pixel sums -- per corner k the three vertex terms and
// the A products (g_a*alpha)*b_k, E = 3 + A columns per corner -- and parks them as ONE row of float4
// slots in shared memory.  Lanes are grouped by triangle id with match.any and the rows of one triangle
// are made consecutive (warp scan over the group sizes); a group is cut into PIECES of at most
// kPieceRows rows.  The reduction runs SLOTS pieces side by side: U = ceil(E/4) lanes per corner,
// J = 3U lanes per piece, lane (corner, u) owning the four columns e = u, u+U, u+2U, u+3U of its corner,
// which the row layout keeps in one float4 -- one 128-bit shared load and four adds per (piece, row) for
// four columns, and per piece ONE fire-and-forget atomic per column; the i-th atomic instruction covers
// U consecutive words of each of the three vertices' gradient rows.
//
// The sums have no fixed order in this mode, so the two places where the reference's ORDER (not its
// per-pixel arithmetic) costs instructions are relaxed: d(out)/d(bary) is an FMA chain over the
// attributes instead of torch's blocked inner sum, and the nine quotients by |det| skip the
// per-numerator window test of SharedDivisor (identical bits inside the window).  The per-pixel terms
// of K.cpp:180-269 are evaluated op for op as everywhere else.

constexpr int kPieceRows = 8;
constexpr int kStripWarps = 4;          // warps (= block rows of 4 pixels) per CTA

template <bool FUSED, int A>
struct BlockRows {
  static constexpr int E = 3 + (FUSED ? A : 0);      // columns per corner: x, y, w terms, then the attributes
  static constexpr int U = (E + 3) / 4;              // lanes per corner (four columns each)
  static constexpr int J = 3 * U;                    // lanes per piece
  static constexpr int ROW4 = J | 1;                 // float4 slots per row; odd keeps 128-bit row stores apart
  static constexpr int SLOTS = 32 / J;               // pieces reduced side by side
  static_assert(J <= 32, "too many attribute channels for the block kernel");
};

// One row of a piece into the lane's four column sums; lanes whose piece is shorter skip it (predicated,
// no branch: the warp runs the longest piece's row count exactly once).
#define PMR_PIECE_ROW(k)                                                                                   \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 a, b, c, d;\n\tsetp.gt.s32 p, %5, %6;\n\t"                 \
               "@p ld.shared.v4.f32 {a, b, c, d}, [%4+%7];\n\t"                                            \
               "@p add.rn.f32 %0, %0, a;\n\t@p add.rn.f32 %1, %1, b;\n\t"                                  \
               "@p add.rn.f32 %2, %2, c;\n\t@p add.rn.f32 %3, %3, d;\n\t}"                                 \
               : "+f"(acc0), "+f"(acc1), "+f"(acc2), "+f"(acc3)                                            \
               : "r"(q), "r"(len), "n"(k), "n"((k) * ROW4 * 16) : "memory");

__device__ __forceinline__ void red_add(float *addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// The per-block inputs as tensor maps: ids [B*H][W] box 8x4, barycentrics [B*H][3W] box 24x4, gradient
// [B*H][GC*W] box 8GC x 4 (GC channels per pixel).
struct BlockMaps {
  CUtensorMap ids, bary, grad;
};

// SHADE (render path, A = 9): `grad` is d(RGBA) [B,H,W,4] with flipped rows; the pixel's nine interpolated
// channels are recomputed from the corner attributes (they were never stored) and the gradient passes through
// the diffuse + ambient lighting (shade_math.cuh) before it enters the interpolation backward.
template <bool FUSED, int A_STATIC, bool SHADE = false>
__global__ void __launch_bounds__(kStripWarps * 32, SHADE ? 4 : 6)
backward_blocks_kernel(const __grid_constant__ BlockMaps maps, int use_tma,
                       const float *__restrict__ grad, const float *__restrict__ verts,
                       const float *__restrict__ attrs, const int32_t *__restrict__ tris,
                       const int32_t *__restrict__ ids, const float *__restrict__ bary,
                       int V, int W, int H, int strip_blocks, float *__restrict__ d_verts,
                       float *__restrict__ d_attrs,
                       const float *__restrict__ light_positions = nullptr,
                       const float *__restrict__ light_intensities = nullptr,
                       const float *__restrict__ ambient = nullptr, int L = 0,
                       const float *__restrict__ background = nullptr) {
  constexpr int A = A_STATIC;
  using R = BlockRows<FUSED, A>;
  constexpr int E = R::E, U = R::U, J = R::J, ROW4 = R::ROW4, SLOTS = R::SLOTS;
  constexpr int GC = SHADE ? 4 : (FUSED ? A : 3);    // gradient channels per pixel
  struct __align__(128) Boxes {              // the TMA boxes of one block (each 128-byte aligned)
    int ids[32];
    float bary[96];
    float grad[32 * GC];
  };
  struct __align__(128) WarpArea {
    float4 rows[32 * ROW4];                  // 32 rows of ROW4 float4 slots
    int4 vids[32];                           // per row: the triangle's vertex ids
    Boxes boxes[2];                          // block `it` lives in boxes[it & 1] while block it + 1 is fetched
    unsigned long long ready;                // mbarrier of the fetches (one is in flight at a time)
    unsigned short pieces[32];               // per piece: first row | rows << 8
  };
  static_assert(offsetof(WarpArea, boxes) % 128 == 0 && sizeof(Boxes) % 128 == 0 && offsetof(Boxes, bary) % 128 == 0 &&
                offsetof(Boxes, grad) % 128 == 0, "TMA destinations must be 128-byte aligned");
  __shared__ WarpArea areas[kStripWarps];
  __shared__ Lights lights;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  if (SHADE) load_lights(lights, light_positions, light_intensities, ambient, b, L);     // block barrier inside
  const int y0 = (blockIdx.y * kStripWarps + warp) * 4, xs = blockIdx.x * strip_blocks * 8;
  if (y0 >= H || xs >= W) return;
  const int n_blocks = min(strip_blocks, (W - xs + 7) >> 3);
  const int iy = y0 + (lane >> 3);
  const bool row_ok = iy < H;
  const float *verts_b = verts + (size_t)b * V * 4;
  const float *attrs_b = FUSED ? attrs + (size_t)b * V * A : nullptr;
  float *dv_b = d_verts ? d_verts + (size_t)b * V * 4 : nullptr;
  float *da_b = (FUSED && d_attrs) ? d_attrs + (size_t)b * V * A : nullptr;

  WarpArea &area = areas[warp];
  const unsigned area_at = (unsigned)__cvta_generic_to_shared(&area);
  const unsigned bar = area_at + (unsigned)offsetof(WarpArea, ready);

  unsigned phase = 0;
  if (use_tma) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  auto fetch_block = [&](int it) {                   // all lanes call; one elected lane issues the copies
    if (elect_one()) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(128 * (4 + GC))) : "memory");
      const int x = xs + it * 8, y = b * H + y0;
      const unsigned to = area_at + (unsigned)offsetof(WarpArea, boxes) + (it & 1) * (unsigned)sizeof(Boxes);
      tma_load_2d(to + (unsigned)offsetof(Boxes, ids), &maps.ids, x, y, bar);
      tma_load_2d(to + (unsigned)offsetof(Boxes, bary), &maps.bary, 3 * x, y, bar);
      // the render path's gradient image is flipped: image row H-1-iy belongs to pixel row iy
      tma_load_2d(to + (unsigned)offsetof(Boxes, grad), &maps.grad, GC * x, SHADE ? b * H + (H - 4 - y0) : y, bar);
    }
  };
  if (use_tma) fetch_block(0);

  for (int it = 0; it < n_blocks; ++it) {
    // Values that depend on the lane only are re-derived per block from an opaque copy of the lane index:
    // hoisted out of the loop they would occupy registers through the whole per-pixel arithmetic.
    int lane_now = lane;
    asm volatile("" : "+r"(lane_now));
    const bool in_image = row_ok && xs + it * 8 + (lane & 7) < W;
    int id = -1;
    float bp[3] = {0.0f, 0.0f, 0.0f};
    const Boxes &box = area.boxes[it & 1];
    if (use_tma) {
      unsigned done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      phase ^= 1u;
      if (it + 1 < n_blocks) fetch_block(it + 1);                 // into the other boxes, while this block is processed
      if (in_image) {
        id = box.ids[lane];
        bp[0] = box.bary[3 * lane]; bp[1] = box.bary[3 * lane + 1]; bp[2] = box.bary[3 * lane + 2];
      }
    } else if (in_image) {
      const size_t px = ((size_t)b * H + iy) * W + xs + it * 8 + (lane & 7);
      id = ids[px];
      bp[0] = bary[3 * px]; bp[1] = bary[3 * px + 1]; bp[2] = bary[3 * px + 2];
    }
    if (id >= 0 && !pixel_is_covered(id, bp)) id = -1;
    if (__ballot_sync(0xffffffffu, id >= 0) == 0u) continue;

    // the dependent chain id -> triangle -> vertices
    int vid[3] = {0, 0, 0};
    float4 pv0 = make_float4(0.f, 0.f, 0.f, 0.f), pv1 = pv0, pv2 = pv0;
    if (id >= 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) vid[k] = __ldg(tris + 3 * (size_t)id + k);
      const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
      pv0 = __ldg(v4 + vid[0]); pv1 = __ldg(v4 + vid[1]); pv2 = __ldg(v4 + vid[2]);
    }
    float g_local[SHADE ? 9 : GC];
    if (id >= 0) {                                  // the pixel's gradient row
      if constexpr (SHADE) {
        const float4 g4 = use_tma ? reinterpret_cast<const float4 *>(box.grad)[(3 - (lane >> 3)) * 8 + (lane & 7)]
                                  : reinterpret_cast<const float4 *>(grad)[((size_t)b * H + (H - 1 - iy)) * W + xs + it * 8 + (lane & 7)];
        g_local[0] = g4.x; g_local[1] = g4.y; g_local[2] = g4.z;
      } else if (use_tma) {
#pragma unroll
        for (int a = 0; a < GC; ++a) g_local[a] = box.grad[lane * GC + a];
      } else {
        const float *g_p = grad + (((size_t)b * H + iy) * W + xs + it * 8 + (lane & 7)) * GC;
#pragma unroll
        for (int a = 0; a < GC; ++a) g_local[a] = __ldg(g_p + a);
      }
    }
    if constexpr (SHADE) {
      if (id >= 0) {
        // the pixel's interpolated channels, exactly as the forward pass computed them (rast.py:118-150)
        const float alpha = coverage_alpha(bp[0], bp[1], bp[2]);
        const float one_minus = 1.0f - alpha;
        const float *c0 = attrs_b + (size_t)vid[0] * A, *c1 = attrs_b + (size_t)vid[1] * A, *c2 = attrs_b + (size_t)vid[2] * A;
        float px[9];
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const float img = __ldg(c0 + a) * bp[0] + __ldg(c1 + a) * bp[1] + __ldg(c2 + a) * bp[2];
          px[a] = alpha * img + one_minus * __ldg(background + a);
        }
        const float g[3] = {g_local[0], g_local[1], g_local[2]};
        shade_diffuse_pixel_backward(px, px + 3, px + 6, g, lights, L, ambient != nullptr, g_local, g_local + 3, g_local + 6);
      }
    }

    // Group the covered lanes by triangle and give every covered lane a ROW: the rows of one triangle
    // are consecutive (groups ordered by their first lane).  Uncovered lanes get private keys, match
    // nobody and own no row.  Every kPieceRows-th lane of a group heads a piece.
    const unsigned peers = __match_any_sync(0xffffffffu, id >= 0 ? id : -1 - lane);
    const int leader = __ffs(peers) - 1;
    const int group_size = __popc(peers), rank = __popc(peers & ((1u << lane) - 1u));
    const int lead_size = (id >= 0 && lane == leader) ? group_size : 0;
    const int before = warp_inclusive_scan(lead_size) - lead_size;        // rows of the groups led by lower lanes
    const int pos = __shfl_sync(0xffffffffu, before, leader) + rank;
    const bool head = id >= 0 && (rank & (kPieceRows - 1)) == 0;
    const unsigned heads = __reduce_or_sync(0xffffffffu, head ? (1u << pos) : 0u);   // bit r: row r starts a piece
    const int n_pieces = __popc(heads);

    if (id >= 0) {
      if (head) area.pieces[__popc(heads & ((1u << pos) - 1u))] = (unsigned short)(pos | (min(kPieceRows, group_size - rank) << 8));
      // d(loss)/d(bary): given, or through the interpolation (d_b_k = sum_a (g_a*alpha) * corner_k[a])
      float gb[3];
      float *ga = g_local;                       // g_a * alpha, in place
      if (FUSED) {
        // Covered pixels have a barycentric sum ~ 1: the clamp of rast.py:145-146 is saturated and passes
        // no gradient, so the 2*d_alpha term of the reference's autograd is exactly zero here.
        const float alpha = coverage_alpha(bp[0], bp[1], bp[2]);
        const float *c[3] = {attrs_b + (size_t)vid[0] * A, attrs_b + (size_t)vid[1] * A, attrs_b + (size_t)vid[2] * A};
#pragma unroll
        for (int a = 0; a < A; ++a) ga[a] *= alpha;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float acc = ga[0] * __ldg(c[k]);
#pragma unroll
          for (int a = 1; a < A; ++a) acc = __fmaf_rn(ga[a], __ldg(c[k] + a), acc);
          gb[k] = acc;
        }
      } else {
        gb[0] = g_local[0]; gb[1] = g_local[1]; gb[2] = g_local[2];
      }
      float m[9], terms[9];
      const float det = adjugate_signed(pv0.x, pv1.x, pv2.x, pv0.y, pv1.y, pv2.y, pv0.w, pv1.w, pv2.w, m);
      vertex_terms<false>(m, fabsf(det), bp, gb, terms);
      float4 *row4 = area.rows + pos * ROW4;
#pragma unroll
      for (int jj = 0; jj < J; ++jj) {              // slot (corner k, uu): columns e = uu + i*U
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = jj / U, e = jj % U + i * U;   // constants after unrolling
          if (e >= E) v[i] = 0.0f;
          else if (e < 3) v[i] = terms[3 * k + e];
          else v[i] = ga[FUSED ? e - 3 : 0] * bp[k];
        }
        row4[jj] = make_float4(v[0], v[1], v[2], v[3]);
      }
      area.vids[pos] = make_int4(vid[0], vid[1], vid[2], 0);
    }
    __syncwarp();

    // Reduction: SLOTS pieces per round; a lane adds its four columns over the rows of its piece and issues
    // one atomic per column.  Column e of corner k goes to d_verts[vtx_k*4 + {0,1,3}[e]] for e < 3 and to
    // d_attrs[vtx_k*A + e-3] otherwise.
    // The lane's role: piece slot, corner, columns e = u + i*U of that corner.
    const int slot = lane_now / J, j = lane_now - slot * J;
    const int corner = min(j / U, 2), u = j - (j / U) * U;
    const bool reducer = slot < SLOTS;
    const unsigned col_at = area_at + (reducer ? j : 0) * 16;
    for (int t = 0; t < n_pieces; t += SLOTS) {
      const int piece = t + slot;
      const int packed = (reducer && piece < n_pieces) ? area.pieces[piece] : 0;
      const int first = packed & 0xff, len = packed >> 8;           // len == 0: nothing to do in this round
      const unsigned q = col_at + first * (ROW4 * 16);
      const int longest = __reduce_max_sync(0xffffffffu, len);      // warp-uniform trip count
      float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
      static_assert(kPieceRows == 8, "the jump table below is written for pieces of up to 8 rows");
      switch (longest) {                                  // straight-line code per piece length
        case 8: PMR_PIECE_ROW(7)
        case 7: PMR_PIECE_ROW(6)
        case 6: PMR_PIECE_ROW(5)
        case 5: PMR_PIECE_ROW(4)
        case 4: PMR_PIECE_ROW(3)
        case 3: PMR_PIECE_ROW(2)
        case 2: PMR_PIECE_ROW(1)
        default: PMR_PIECE_ROW(0)
      }
      if (len > 0) {
        const int vtx = reinterpret_cast<const int *>(area.vids + first)[corner];
        const float sums[4] = {acc0, acc1, acc2, acc3};
        float *to_vert = dv_b + (unsigned)(vtx * 4 + column_of(u < 3 ? u : 0));   // column u of the vertex row
        float *to_attr = da_b + (unsigned)(vtx * A + u) - 3;                        // column u of the attribute row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i * U >= E) continue;                             // beyond the last column for every lane
          const int e = u + i * U;                              // column of the corner
          const bool is_vert = (i * U >= 3) ? false : ((U <= 3 && i == 0) ? true : e < 3);
          const bool exists = ((i + 1) * U <= E) ? true : e < E;
          if (is_vert) {
            // U < 3: columns x, y of i = 0 and w (or an attribute) of i = 1 ...: the row is walked in steps of U
            if (dv_b != nullptr) red_add(i == 0 ? to_vert : dv_b + (unsigned)(vtx * 4 + column_of(e)), sums[i]);
          } else if (exists) {
            if (da_b != nullptr) red_add(to_attr + i * U, sums[i]);
          }
        }
      }
    }
    __syncwarp();                                   // rows, pieces and vertex ids are rewritten by the next block
  }
}
#undef PMR_PIECE_ROW

// ---------------------------------------------------------------------------------------------
// Ordered (parity) mode
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
vertex_box_init_kernel(int4 *__restrict__ vbox, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vbox[i] = make_int4(INT_MAX, INT_MIN, INT_MAX, INT_MIN);   // left right bottom top
}

// Union, per (image, vertex), of the pixel boxes of the triangles that use the vertex.  A pixel
// drawn from triangle t always lies inside t's box (K.cpp:374-375), so the union bounds every
// pixel that can contribute to the vertex.
__global__ void __launch_bounds__(256)
vertex_box_kernel(const float *__restrict__ verts, const int32_t *__restrict__ tris, int V, int T,
                  int W, int H, float half_w, float half_h, int4 *__restrict__ vbox) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int i0 = __ldg(tris + 3 * (size_t)t), i1 = __ldg(tris + 3 * (size_t)t + 1), i2 = __ldg(tris + 3 * (size_t)t + 2);
  const float4 *v4 = reinterpret_cast<const float4 *>(verts + (size_t)b * V * 4);
  const PixelBox bx = triangle_box(__ldg(v4 + i0), __ldg(v4 + i1), __ldg(v4 + i2), half_w, half_h, W, H);
  if (bx.left >= bx.right || bx.bottom >= bx.top) return;
  const int vid[3] = {i0, i1, i2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    int *q = reinterpret_cast<int *>(vbox + (size_t)b * V + vid[j]);
    atomicMin(q + 0, bx.left);
    atomicMax(q + 1, bx.right);
    atomicMin(q + 2, bx.bottom);
    atomicMax(q + 3, bx.top);
  }
}

constexpr int kOrderedWarps = 8;

template <bool FUSED>
__global__ void __launch_bounds__(kOrderedWarps * 32)
backward_ordered_kernel(const float *__restrict__ grad, const float *__restrict__ verts,
                        const float *__restrict__ attrs, const int32_t *__restrict__ tris,
                        const int32_t *__restrict__ ids, const float *__restrict__ bary,
                        const int4 *__restrict__ vbox, int V, int A, int W, int H, long long n_pairs,
                        float *__restrict__ d_verts, float *__restrict__ d_attrs) {
  extern __shared__ float rows_all[];   // [warp][lane][corner][3 + A]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * kOrderedWarps + warp;
  if (pair >= n_pairs) return;
  const int b = (int)(pair / V), v = (int)(pair % V);
  const int ncomp = 3 + (FUSED ? A : 0);
  float *rows = rows_all + (size_t)warp * 32 * 3 * ncomp;
  const int4 box = vbox[pair];
  const float *verts_b = verts + (size_t)b * V * 4;
  const float *attrs_b = FUSED ? attrs + (size_t)b * V * A : nullptr;

  constexpr int kMaxSlots = 4;          // components lane, lane+32, ... (A <= 125)
  float acc[kMaxSlots] = {0.0f, 0.0f, 0.0f, 0.0f};

  const int bw = box.y - box.x;
  const long long n = (box.x < box.y && box.z < box.w) ? (long long)bw * (box.w - box.z) : 0;
  for (long long k0 = 0; k0 < n; k0 += 32) {
    const long long k = k0 + lane;
    int id = -1, corners = 0;
    long long p = 0;
    float bp[3];
    if (k < n) {
      const int iy = box.z + (int)(k / bw), ix = box.x + (int)(k % bw);
      p = ((long long)b * H + iy) * W + ix;
      id = ids[p];
      bp[0] = bary[3 * p]; bp[1] = bary[3 * p + 1]; bp[2] = bary[3 * p + 2];
      if (pixel_is_covered(id, bp)) {
#pragma unroll
        for (int j = 0; j < 3; ++j) corners |= (__ldg(tris + 3 * (size_t)id + j) == v) << j;
      }
    }
    const unsigned hits = __ballot_sync(0xffffffffu, corners != 0);
    if (hits == 0u) continue;
    if (corners != 0) {
      PixelGrad pg;
      const float *g_p = grad + (size_t)p * (FUSED ? A : 3);
      pixel_grad<FUSED>(verts_b, attrs_b, tris, id, bp, g_p, A, pg);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (corners & (1 << j)) {
          float *row = rows + ((size_t)lane * 3 + j) * ncomp;
          row[0] = pg.terms[3 * j + 0]; row[1] = pg.terms[3 * j + 1]; row[2] = pg.terms[3 * j + 2];
          if (FUSED)
            for (int a = 0; a < A; ++a) row[3 + a] = (g_p[a] * pg.alpha) * pg.b[j];
        }
      }
    }
    __syncwarp();
    // Fold in pixel order (ascending lane), corner order within a pixel.
    unsigned todo = hits;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int cm = __shfl_sync(0xffffffffu, corners, src);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (cm & (1 << j)) {
          const float *row = rows + ((size_t)src * 3 + j) * ncomp;
#pragma unroll
          for (int s = 0; s < kMaxSlots; ++s) {
            const int c = lane + 32 * s;
            if (c < ncomp) acc[s] += row[c];
          }
        }
      }
    }
    __syncwarp();
  }

  if (d_verts != nullptr) {
    float *dv = d_verts + (size_t)pair * 4;
    if (lane < 3) dv[column_of(lane)] = acc[0];
    if (lane == 3) dv[2] = 0.0f;        // z column never receives gradient
  }
  if (FUSED && d_attrs != nullptr) {
    float *da = d_attrs + (size_t)pair * A;
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
      const int c = lane + 32 * s;
      if (c >= 3 && c < ncomp) da[c - 3] = acc[s];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------

// Blocks per strip: long strips amortise the per-warp setup, but the grid must still fill the GPU
// (about 32 warps on each of the SMs, several times over).
static int strip_blocks_for(const Context *ctx, int W, int H, int B) {
  if (ctx->strip_blocks_override > 0) return ctx->strip_blocks_override;
  const long long block_rows = (long long)((H + 3) / 4) * B;
  const long long wanted = 4LL * 32 * ctx->sm_count;
  for (int n = 8; n > 1; n >>= 1)
    if (block_rows * ((W + 8 * n - 1) / (8 * n)) >= wanted) return n;
  return 1;
}

static dim3 strip_grid(int W, int H, int B, int strip) {
  return dim3((W + 8 * strip - 1) / (8 * strip), (H + 4 * kStripWarps - 1) / (4 * kStripWarps), B);
}

// cuTensorMapEncodeTiled through the runtime (the library does not link libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult found;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &found) != cudaSuccess ||
        found != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Tensor map over a dense [rows][inner] array of 4-byte elements with boxes of 4 rows.  False when the
// array cannot be described (base or row pitch not 16-byte aligned, box too wide, no driver entry point).
static bool make_block_map(CUtensorMap *map, CUtensorMapDataType type, const void *base, long long inner,
                           long long rows, int box_inner) {
  EncodeTiledFn fn = encode_tiled();
  if (fn == nullptr || base == nullptr || ((uintptr_t)base & 15) != 0 || (inner * 4) % 16 != 0 || box_inner > 256 ||
      inner <= 0 || rows <= 0 || inner >= (1LL << 32) || rows >= (1LL << 32))
    return false;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t pitch[1] = {(cuuint64_t)inner * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, 4};
  const cuuint32_t step[2] = {1, 1};
  return fn(map, type, 2, const_cast<void *>(base), dims, pitch, box, step, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int make_block_maps(const Context *ctx, BlockMaps *maps, const int32_t *ids, const float *bary, const float *grad,
                           int grad_channels, int B, int W, int H) {
  if (ctx->no_tma) return 0;
  const long long rows = (long long)B * H;
  return make_block_map(&maps->ids, CU_TENSOR_MAP_DATA_TYPE_INT32, ids, W, rows, 8) &&
         make_block_map(&maps->bary, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, bary, 3LL * W, rows, 24) &&
         make_block_map(&maps->grad, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, grad, (long long)grad_channels * W, rows, 8 * grad_channels);
}

int backward_impl(Context *ctx, const float *df_dbary, const float *grad_image, const float *verts,
                  const float *attrs, const int32_t *tris, const int32_t *ids, const float *bary,
                  int B, int V, int T, int A, int W, int H, float *d_verts, float *d_attrs, int mode,
                  cudaStream_t stream, const ShadeArgs *shade) {
  if (shade != nullptr) {
    // render path: one kernel from d(RGBA) to the vertex / attribute gradients (atomic accumulation)
    const long long n = (long long)B * V;
    if (n == 0) return PMR_OK;
    if (A != 9 || mode != PMR_BACKWARD_ATOMIC)
      return set_error(ctx, PMR_ERR_INVALID, "the render backward needs 9 attribute channels and the atomic mode");
    StageScope timed(ctx, PMR_STAGE_BACKWARD, stream);
    if (d_verts) PMR_CUDA(ctx, cudaMemsetAsync(d_verts, 0, (size_t)n * 4 * sizeof(float), stream));
    if (d_attrs) PMR_CUDA(ctx, cudaMemsetAsync(d_attrs, 0, (size_t)n * 9 * sizeof(float), stream));
    if ((long long)W * H * B == 0 || T == 0) return PMR_OK;
    const int strip = strip_blocks_for(ctx, W, H, B);
    BlockMaps maps;
    const int use_tma = make_block_maps(ctx, &maps, ids, bary, shade->grad_rgba, 4, B, W, H);
    backward_blocks_kernel<true, 9, true><<<strip_grid(W, H, B, strip), kStripWarps * 32, 0, stream>>>(
        maps, use_tma, shade->grad_rgba, verts, attrs, tris, ids, bary, V, W, H, strip, d_verts, d_attrs,
        shade->light_positions, shade->light_intensities, shade->ambient, shade->L, shade->background);
    ctx->launches += 1;
    return check_launch(ctx, "backward_blocks_kernel (render)");
  }
  const bool fused = grad_image != nullptr;
  const float *grad = fused ? grad_image : df_dbary;
  const long long ppi = (long long)W * H, total = ppi * B;
  const long long n_pairs = (long long)B * V;
  if (n_pairs == 0) return PMR_OK;
  StageScope timed(ctx, PMR_STAGE_BACKWARD, stream);

  if (mode == PMR_BACKWARD_ATOMIC) {
    if (d_verts) PMR_CUDA(ctx, cudaMemsetAsync(d_verts, 0, (size_t)n_pairs * 4 * sizeof(float), stream));
    if (fused && d_attrs) PMR_CUDA(ctx, cudaMemsetAsync(d_attrs, 0, (size_t)n_pairs * A * sizeof(float), stream));
    if (total == 0 || T == 0) return PMR_OK;
    const int strip = strip_blocks_for(ctx, W, H, B);
    BlockMaps maps;
    const int use_tma = make_block_maps(ctx, &maps, ids, bary, grad, fused ? A : 3, B, W, H);
#define PMR_BLOCKS(F, AS)                                                                                 \
  backward_blocks_kernel<F, AS><<<strip_grid(W, H, B, strip), kStripWarps * 32, 0, stream>>>(            \
      maps, use_tma, grad, verts, attrs, tris, ids, bary, V, W, H, strip, d_verts, d_attrs)
    if (!fused) PMR_BLOCKS(false, 1);
    else if (A == 9) PMR_BLOCKS(true, 9);
    else if (A == 3) PMR_BLOCKS(true, 3);
    else if (A == 4) PMR_BLOCKS(true, 4);
    else if (A == 12) PMR_BLOCKS(true, 12);
    else if (A == 13) PMR_BLOCKS(true, 13);
    else {
      // other attribute counts: one thread per pixel, per-lane atomics
      const unsigned grid = (unsigned)((total + 255) / 256);
      backward_atomic_kernel<true, 0><<<grid, 256, 0, stream>>>(grad, verts, attrs, tris, ids, bary, V, A, ppi, total,
                                                                d_verts, d_attrs);
    }
#undef PMR_BLOCKS
    ctx->launches += 1;
    return check_launch(ctx, "backward_atomic_kernel");
  }

  if (mode != PMR_BACKWARD_ORDERED) return set_error(ctx, PMR_ERR_INVALID, "unknown backward mode %d", mode);
  // shared memory of the ordered kernel: kOrderedWarps * 32 * 3 * (3 + A) floats <= 227 KB
  if (fused && A > 72) return set_error(ctx, PMR_ERR_SIZE, "ordered backward supports at most 72 attributes");
  int rc = ctx->scratch.reserve(ctx, (size_t)n_pairs * sizeof(int4));
  if (rc) return rc;
  int4 *vbox = (int4 *)ctx->scratch.ptr;
  const float half_w = (float)(0.5 * W), half_h = (float)(0.5 * H);
  vertex_box_init_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, stream>>>(vbox, n_pairs);
  ctx->launches += 1;
  if (T > 0 && total > 0) {
    vertex_box_kernel<<<dim3((T + 255) / 256, B), 256, 0, stream>>>(verts, tris, V, T, W, H, half_w, half_h, vbox);
    ctx->launches += 1;
  }
  const int ncomp = 3 + (fused ? A : 0);
  const size_t smem = (size_t)kOrderedWarps * 32 * 3 * ncomp * sizeof(float);
  const unsigned grid = (unsigned)((n_pairs + kOrderedWarps - 1) / kOrderedWarps);
  if (fused) {
    PMR_CUDA(ctx, cudaFuncSetAttribute(backward_ordered_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    backward_ordered_kernel<true><<<grid, kOrderedWarps * 32, smem, stream>>>(grad, verts, attrs, tris, ids, bary,
                                                                            vbox, V, A, W, H, n_pairs, d_verts, d_attrs);
  } else {
    backward_ordered_kernel<false><<<grid, kOrderedWarps * 32, smem, stream>>>(grad, verts, attrs, tris, ids, bary,
                                                                             vbox, V, A, W, H, n_pairs, d_verts, d_attrs);
  }
  ctx->launches += 1;
  return check_launch(ctx, "backward_ordered_kernel");
}

}  // namespace pmr
