// raster_backward.cu -- backward of rasterize_triangles (K.cpp:131-273) batched over images,
// optionally fused with the backward of the attribute interpolation (autograd of
// rast.py:118-150), in two accumulation modes:
//
//   PMR_BACKWARD_ATOMIC   backward_blocks_kernel: one warp per strip of 8x4 pixel blocks, the per-pixel inputs
//                         fetched by TMA; the per-pixel sums are parked in shared memory as rows of float4
//                         slots, rows sorted by triangle, and a lane adds four columns over a triangle's rows:
//                         one fire-and-forget atomic per (block, triangle, column).  fp32 sums in arbitrary
//                         order over the reference's per-pixel terms.  (Attribute counts without a specialised
//                         instance fall back to backward_atomic_kernel, one thread per pixel.)  With SHADE it is
//                         also the backward of the fused render path.
//   PMR_BACKWARD_ORDERED  a stable radix sort of the (pixel, corner) entries by vertex id, then one warp per
//                         run of vertices folds the entries' terms strictly in entry order -- pixel ascending,
//                         corner 0..2 within a pixel: the reference's summation order (K.cpp:156-157, :232-269;
//                         index_put_ order of rast.py:130-132), so the result is bit-reproducible and equals the
//                         reference's.
//
// Per-pixel arithmetic is shared with the forward pass (raster_math.cuh) and follows the
// reference op for op; compile with -fmad=false.
#include <stddef.h>

#include "pmr_internal.cuh"
#include "raster_math.cuh"
#include "shade_math.cuh"
#include "tensor_maps.cuh"

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
// mbarrier; SASS UTMALDG): the boxes of the NEXT block are requested once the values of the current one
// have been used (after the row stores), so they land during the reduction, no registers are held for them
// and the SM issues no address arithmetic, LDG or STS for them.  What depends on the lane only (its role in
// the reduction, shared-memory addresses, barrier) is re-derived per block from an opaque copy of the
// thread index (kept in registers across the loop it would be spilled).
//
// Per block: coverage and triangle id of every pixel in raster order; the pixels are then SORTED by triangle
// inside the warp (match.any + one warp scan give every pixel its row; lane L takes over the pixel whose row
// is L) so that the rows of one triangle are consecutive and a lane's row is its lane index.  Every covered
// lane evaluates its per-pixel sums -- per corner k the three vertex terms and the A products (g_a*alpha)*b_k,
// E = 3 + A columns per corner -- and parks them as ONE row of float4 slots in shared memory (conflict-free:
// a quarter warp writes eight consecutive rows); a triangle's rows are cut into PIECES of at most
// kPieceRows rows.  The reduction runs SLOTS pieces side by side: U = ceil(E/4) lanes per corner,
// J = 3U lanes per piece, lane (corner, u) owning the four columns e = u, u+U, u+2U, u+3U of its corner,
// which the row layout keeps in one float4 -- one 128-bit shared load and four adds per (piece, row) for
// four columns, and per piece ONE fire-and-forget atomic per column; the i-th atomic instruction covers
// U consecutive words of each of the three vertices' gradient rows.
//
// The per-pixel terms are the reference's bits (K.cpp:180-269 op for op; d(out)/d(bary) folded over the
// attributes in torch's order): only the ORDER of the sums over pixels differs from the reference in this
// mode.  The nine quotients by |det| skip the per-numerator window test of SharedDivisor (identical bits
// inside the window, i.e. for every quotient between 1e-30 and 1e30 in magnitude).

constexpr int kPieceRows = 16;
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

__device__ __forceinline__ void red_add(float *addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// The per-block inputs as tensor maps: ids [B*H][W] box 8x4, barycentrics [B*H][3W] box 24x4, gradient
// [B*H][GC*W] box 8GC x 4 (GC channels per pixel).
struct BlockMaps {
  CUtensorMap ids, bary, grad;
};

// SHADE (render path, A = 9): `grad` is d(RGBA) [B,H,W,4] with flipped rows; the pixel's nine interpolated
// channels are recomputed from the corner attributes (they were never stored) and the gradient passes through
// the diffuse + ambient lighting (shade_math.cuh) before it enters the interpolation backward.
template <bool FUSED, int A_STATIC, bool SHADE = false, int MIN_CTAS = (SHADE ? 4 : 8)>
__global__ void __launch_bounds__(kStripWarps * 32, MIN_CTAS)
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
  // the reduction reads up to kPieceRows - 1 rows past a piece (values unused): they must stay inside the area
  constexpr int kTail = 32 * 16 + 32 * 4 + 96 * 4 + 32 * GC * 4 + 8 + 64 + 32;
  constexpr int kNeeded = (kPieceRows - 1) * ROW4 * 16;
  struct __align__(128) WarpArea {
    float4 rows[32 * ROW4];                  // 32 rows of ROW4 float4 slots
    int4 vids[32];                           // per row: the triangle's vertex ids
    int ids[32];                             // the TMA boxes of the block in flight (each 128-byte aligned)
    float bary[96];
    float grad[32 * GC];
    unsigned long long ready;                // their mbarrier (one fetch is in flight at a time)
    unsigned short pieces[32];               // per piece: first row | rows << 8 | continues the previous piece's group << 15
    unsigned char order[32];                 // row -> the pixel (lane of the block's raster order) that owns it
    char pad[kNeeded > kTail ? kNeeded - kTail : 1];
  };
  static_assert(offsetof(WarpArea, ids) % 128 == 0 && offsetof(WarpArea, bary) % 128 == 0 &&
                offsetof(WarpArea, grad) % 128 == 0, "TMA destinations must be 128-byte aligned");
  __shared__ WarpArea areas[kStripWarps];
  __shared__ Lights lights;

  const int b = blockIdx.z;
  if (SHADE) load_lights(lights, light_positions, light_intensities, ambient, b, L);     // block barrier inside
  const int xs = blockIdx.x * strip_blocks * 8;
  if ((int)(blockIdx.y * kStripWarps + (threadIdx.x >> 5)) * 4 >= H || xs >= W) return;
  const int n_blocks = min(strip_blocks, (W - xs + 7) >> 3);
  const float *verts_b = verts + (size_t)b * V * 4;
  const float *attrs_b = FUSED ? attrs + (size_t)b * V * A : nullptr;
  float *dv_b = d_verts ? d_verts + (size_t)b * V * 4 : nullptr;
  float *da_b = (FUSED && d_attrs) ? d_attrs + (size_t)b * V * A : nullptr;

  // Everything that depends on the thread only (lane, warp, the warp's shared-memory area, the block row) is
  // re-derived from an OPAQUE copy of the thread index wherever it is needed: hoisted out of the loop and kept
  // in registers it would be spilled around the per-pixel arithmetic, and a spilled loop variable is a
  // local-memory round trip at the top of every block.
  struct Where {
    int lane, warp, y0;
    unsigned area_at;
    WarpArea *area;
  };
  const unsigned tid_kept = threadIdx.x;
  auto where = [&]() {
    unsigned tid = tid_kept;
    asm volatile("" : "+r"(tid));            // opaque: what is derived from it is derived here, not before the loop
    Where w;
    w.lane = tid & 31;
    w.warp = tid >> 5;
    w.y0 = (blockIdx.y * kStripWarps + w.warp) * 4;
    w.area = &areas[w.warp];
    w.area_at = (unsigned)__cvta_generic_to_shared(w.area);
    return w;
  };
  auto fetch_block = [&](const Where &w, int it) {    // all lanes call; one elected lane issues the copies
    if (elect_one()) {
      const unsigned bar = w.area_at + (unsigned)offsetof(WarpArea, ready);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(128 * (4 + GC))) : "memory");
      const int x = xs + it * 8, y = b * H + w.y0;
      tma_load_2d(w.area_at + (unsigned)offsetof(WarpArea, ids), &maps.ids, x, y, bar);
      tma_load_2d(w.area_at + (unsigned)offsetof(WarpArea, bary), &maps.bary, 3 * x, y, bar);
      // the render path's gradient image is flipped: image row H-1-iy belongs to pixel row iy
      tma_load_2d(w.area_at + (unsigned)offsetof(WarpArea, grad), &maps.grad, GC * x, SHADE ? b * H + (H - 4 - w.y0) : y, bar);
    }
  };
  if (use_tma) {
    const Where w = where();
    if (w.lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(w.area_at + (unsigned)offsetof(WarpArea, ready)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    fetch_block(w, 0);
  }

  for (int it = 0; it < n_blocks; ++it) {
    // ---- the block's pixels in raster order: triangle id, coverage
    int id0 = -1;
    {
      const Where w = where();
      const int lane = w.lane, iy = w.y0 + (lane >> 3), ix = xs + it * 8 + (lane & 7);
      const bool in_image = iy < H && ix < W;
      if (use_tma) {
        // exactly one fetch per block: the barrier's phase is the block's parity
        const unsigned bar = w.area_at + (unsigned)offsetof(WarpArea, ready);
        unsigned done = 0;
        while (!done) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(bar), "r"((unsigned)(it & 1)) : "memory");
        }
        const WarpArea &area = *w.area;
        if (in_image) {
          id0 = area.ids[lane];
          if (id0 == 0) {           // K.cpp:162: id 0 with barycentrics summing below 0.9 draws nothing
            const float b3[3] = {area.bary[3 * lane], area.bary[3 * lane + 1], area.bary[3 * lane + 2]};
            if (!pixel_is_covered(0, b3)) id0 = -1;
          }
        }
      } else if (in_image) {
        const size_t px = ((size_t)b * H + iy) * W + ix;
        id0 = ids[px];
        if (id0 == 0) {
          const float b3[3] = {bary[3 * px], bary[3 * px + 1], bary[3 * px + 2]};
          if (!pixel_is_covered(0, b3)) id0 = -1;
        }
      }
    }
    const unsigned covered_lanes = __ballot_sync(0xffffffffu, id0 >= 0);
    if (covered_lanes == 0u) {
      if (use_tma && it + 1 < n_blocks) {
        __syncwarp();
        fetch_block(where(), it + 1);
      }
      continue;
    }

    // ---- sort the block's pixels by triangle: lane L takes over the pixel whose ROW is L.  The rows of one
    // triangle are consecutive (groups ordered by their first lane, uncovered pixels last), and because a
    // lane later writes row `lane`, the 128-bit row stores of a quarter warp go to eight consecutive rows --
    // conflict-free, where rows in raster order of the pixels cost 7.6 wavefronts per store instead of 4.
    int id, rank, group_size;
    float bp[3] = {0.0f, 0.0f, 0.0f};
    float g_local[SHADE ? 9 : GC];
    {
      const Where w = where();
      const int lane = w.lane;
      WarpArea &area = *w.area;
      const unsigned peers0 = __match_any_sync(0xffffffffu, id0 >= 0 ? id0 : -1 - lane);
      const int leader0 = __ffs(peers0) - 1;
      const int size0 = __popc(peers0), rank0 = __popc(peers0 & ((1u << lane) - 1u));
      const int lead_size = (id0 >= 0 && lane == leader0) ? size0 : 0;
      const int before = warp_inclusive_scan(lead_size) - lead_size;      // rows of the groups led by lower lanes
      const int group_row = __shfl_sync(0xffffffffu, before, leader0);    // by all lanes: not under the condition
      const int row0 = id0 >= 0 ? group_row + rank0
                                : __popc(covered_lanes) + __popc(~covered_lanes & ((1u << lane) - 1u));
      area.order[row0] = (unsigned char)lane;
      __syncwarp();
      const int src = area.order[lane];          // the pixel (lane of the raster order) this lane takes over
      id = __shfl_sync(0xffffffffu, id0, src);
      rank = __shfl_sync(0xffffffffu, rank0, src);
      group_size = __shfl_sync(0xffffffffu, size0, src);
      if (use_tma) {
        bp[0] = area.bary[3 * src]; bp[1] = area.bary[3 * src + 1]; bp[2] = area.bary[3 * src + 2];
        if constexpr (SHADE) {
          const float4 g4 = reinterpret_cast<const float4 *>(area.grad)[(3 - (src >> 3)) * 8 + (src & 7)];
          g_local[0] = g4.x; g_local[1] = g4.y; g_local[2] = g4.z;
        } else {
#pragma unroll
          for (int a = 0; a < GC; ++a) g_local[a] = area.grad[src * GC + a];
        }
        // (The boxes are refilled further down, once these values have been USED: a load that has merely been
        // issued can still be in the L1 queue when the copy engine's data arrives.)
      } else if (id >= 0) {
        const int iy = w.y0 + (src >> 3), ix = xs + it * 8 + (src & 7);
        const size_t px = ((size_t)b * H + iy) * W + ix;
        bp[0] = bary[3 * px]; bp[1] = bary[3 * px + 1]; bp[2] = bary[3 * px + 2];
        if constexpr (SHADE) {
          const float4 g4 = reinterpret_cast<const float4 *>(grad)[((size_t)b * H + (H - 1 - iy)) * W + ix];
          g_local[0] = g4.x; g_local[1] = g4.y; g_local[2] = g4.z;
        } else {
#pragma unroll
          for (int a = 0; a < GC; ++a) g_local[a] = __ldg(grad + px * GC + a);
        }
      }
    }

    // the dependent chain id -> triangle -> vertices.  (Base pointers pass through an opaque copy for the same
    // reason as the thread index above: hoisted out of the loop each would hold two registers throughout.)
    int vid[3] = {0, 0, 0};
    float4 pv0 = make_float4(0.f, 0.f, 0.f, 0.f), pv1 = pv0, pv2 = pv0;
    const int32_t *tris_now = tris;
    const float *verts_now = verts_b, *attrs_now = attrs_b;
    asm volatile("" : "+l"(tris_now), "+l"(verts_now), "+l"(attrs_now));
    if (id >= 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) vid[k] = __ldg(tris_now + 3 * (size_t)id + k);
      const float4 *v4 = reinterpret_cast<const float4 *>(verts_now);
      pv0 = __ldg(v4 + vid[0]); pv1 = __ldg(v4 + vid[1]); pv2 = __ldg(v4 + vid[2]);
    }
    if constexpr (SHADE) {
      if (id >= 0) {
        // the pixel's interpolated channels, exactly as the forward pass computed them (rast.py:118-150)
        const float alpha = coverage_alpha(bp[0], bp[1], bp[2]);
        const float one_minus = 1.0f - alpha;
        const float *c0 = attrs_now + (size_t)vid[0] * A, *c1 = attrs_now + (size_t)vid[1] * A, *c2 = attrs_now + (size_t)vid[2] * A;
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

    // Every covered lane's row is its lane index; every kPieceRows-th lane of a triangle heads a piece.
    const Where wg = where();
    const int lane = wg.lane;
    WarpArea &area = *wg.area;
    const int pos = lane;
    const bool head = id >= 0 && (rank & (kPieceRows - 1)) == 0;
    const unsigned heads = __ballot_sync(0xffffffffu, head);               // bit r: row r starts a piece
    const int n_pieces = __popc(heads);

    float gb[3] = {0.0f, 0.0f, 0.0f};
    if (id >= 0) {
      // d(loss)/d(bary): given, or through the interpolation (d_b_k = sum_a (g_a*alpha) * corner_k[a])
      float *ga = g_local;                       // g_a * alpha, in place
      if (FUSED) {
        // Covered pixels have a barycentric sum ~ 1: the clamp of rast.py:145-146 is saturated and passes
        // no gradient, so the 2*d_alpha term of the reference's autograd is exactly zero here.
        const float alpha = coverage_alpha(bp[0], bp[1], bp[2]);
#pragma unroll
        for (int a = 0; a < A; ++a) ga[a] *= alpha;
#pragma unroll
        for (int k = 0; k < 3; ++k) {              // torch's order: the per-pixel terms are the reference's bits
          const float *ck = attrs_now + (size_t)vid[k] * A;
          gb[k] = torch_inner_sum(A, [&](int a) { return ga[a] * __ldg(ck + a); });
        }
      } else {
        gb[0] = g_local[0]; gb[1] = g_local[1]; gb[2] = g_local[2];
      }
    }
    if (id >= 0) {
      if (head) area.pieces[__popc(heads & ((1u << pos) - 1u))] =
          (unsigned short)(pos | (min(kPieceRows, group_size - rank) << 8) | (rank > 0 ? 0x8000 : 0));
      float *ga = g_local;
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
    // The boxes of the next block: requested only now, when every value read from this block's boxes has gone
    // through the arithmetic above (the row stores could not issue before the loads had returned) -- a
    // shared-memory load and the copy engine's write are ordered by nothing but that; issued right after
    // the loads, a fraction of 1e-5 of the gradient entries came out wrong under load (run-to-run check,
    // profiles/tools/run_to_run.py).  The copy still lands during the reduction below.
    if (use_tma && it + 1 < n_blocks) fetch_block(wg, it + 1);

    // Reduction: SLOTS pieces per round; a lane adds its four columns over the rows of its piece and issues
    // one atomic per column.  Column e of corner k goes to d_verts[vtx_k*4 + {0,1,3}[e]] for e < 3 and to
    // d_attrs[vtx_k*A + e-3] otherwise.
    // The lane's role: piece slot, corner, columns e = u + i*U of that corner.
    const Where wr = where();
    float *dv_now = dv_b, *da_now = da_b;
    asm volatile("" : "+l"(dv_now), "+l"(da_now));
    // J == 9 (A = 9): the 27 reducers are laid out so that every quarter warp reads ONE contiguous 128 bytes of a
    // row (lanes 8s .. 8s+7: slots j = 0..7 of piece s; lanes 24, 25, 26: slot j = 8 of pieces 0, 1, 2): a 128-bit
    // shared load then takes its minimum of four wavefronts instead of six or seven.
    constexpr bool kQuarters = J == 9 && SLOTS == 3;
    const int slot = kQuarters ? (wr.lane < 24 ? wr.lane >> 3 : wr.lane - 24) : wr.lane / J;
    const int j = kQuarters ? (wr.lane < 24 ? wr.lane & 7 : 8) : wr.lane - slot * J;
    const int next_slot_lane = kQuarters ? (wr.lane < 24 ? wr.lane + 8 : wr.lane + 1) : wr.lane + J;   // same j, slot + 1
    const int corner = min(j / U, 2), u = j - (j / U) * U;
    const bool reducer = slot < SLOTS;
    const float4 *area_rows = wr.area->rows;
    for (int t = 0; t < n_pieces; t += SLOTS) {
      const int piece = t + slot;
      const int packed = (reducer && piece < n_pieces) ? area.pieces[piece] : 0;
      const int first = packed & 0xff, len = (packed >> 8) & 0x7f;  // len == 0: nothing to do in this round
      const bool continues = (packed & 0x8000) != 0;                // same triangle as the piece before
      const float4 *q = area_rows + first * ROW4 + (reducer ? j : 0);
      const int longest = __reduce_max_sync(0xffffffffu, len);      // warp-uniform trip count
      float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
      // Two rows per step, both loads issued before the adds.  Loads and adds are predicated on the piece's own
      // length, the entry step on the longest piece of the round (warp-uniform): the warp runs the longest
      // piece's steps exactly once and short pieces cost no shared-memory wavefronts for rows they do not have.
      // ONE block of PTX: as C++ (conditional loads, accumulators through "+f" constraints of one asm per row)
      // every step carried eight register moves and a convergence barrier around its loads, 26 instructions
      // where 12 do the work.
      static_assert(kPieceRows == 16, "the steps below are written for pieces of up to 16 rows");
      {
        const unsigned q_at = (unsigned)__cvta_generic_to_shared(q);
#define PMR_STEP(K, K1, OA, OB)                                                             \
  "S" #K ":\n\t"                                                                            \
  "setp.gt.s32 q0, %4, " #K ";\n\t"                                                         \
  "setp.gt.s32 q1, %4, " #K1 ";\n\t"                                                        \
  "@q0 ld.shared.v4.f32 {a0, a1, a2, a3}, [%6+" OA "];\n\t"                                 \
  "@q1 ld.shared.v4.f32 {b0, b1, b2, b3}, [%6+" OB "];\n\t"                                 \
  "@q0 add.rn.f32 %0, %0, a0;\n\t@q0 add.rn.f32 %1, %1, a1;\n\t"                            \
  "@q0 add.rn.f32 %2, %2, a2;\n\t@q0 add.rn.f32 %3, %3, a3;\n\t"                            \
  "@q1 add.rn.f32 %0, %0, b0;\n\t@q1 add.rn.f32 %1, %1, b1;\n\t"                            \
  "@q1 add.rn.f32 %2, %2, b2;\n\t@q1 add.rn.f32 %3, %3, b3;\n\t"
        asm volatile(
            "{\n\t.reg .pred p, q0, q1;\n\t.reg .f32 a0, a1, a2, a3, b0, b1, b2, b3;\n\t"
            "setp.lt.s32 p, %5, 3;\n\t@p bra.uni S0;\n\t"
            "setp.lt.s32 p, %5, 5;\n\t@p bra.uni S2;\n\t"
            "setp.lt.s32 p, %5, 7;\n\t@p bra.uni S4;\n\t"
            "setp.lt.s32 p, %5, 9;\n\t@p bra.uni S6;\n\t"
            "setp.lt.s32 p, %5, 11;\n\t@p bra.uni S8;\n\t"
            "setp.lt.s32 p, %5, 13;\n\t@p bra.uni S10;\n\t"
            "setp.lt.s32 p, %5, 15;\n\t@p bra.uni S12;\n\t"
            PMR_STEP(14, 15, "%21", "%22") PMR_STEP(12, 13, "%19", "%20") PMR_STEP(10, 11, "%17", "%18")
            PMR_STEP(8, 9, "%15", "%16") PMR_STEP(6, 7, "%13", "%14") PMR_STEP(4, 5, "%11", "%12")
            PMR_STEP(2, 3, "%9", "%10") PMR_STEP(0, 1, "%7", "%8")
            "}"
            : "+f"(acc0), "+f"(acc1), "+f"(acc2), "+f"(acc3)
            : "r"(len), "r"(longest), "r"(q_at),
              "n"(0 * ROW4 * 16), "n"(1 * ROW4 * 16), "n"(2 * ROW4 * 16), "n"(3 * ROW4 * 16), "n"(4 * ROW4 * 16),
              "n"(5 * ROW4 * 16), "n"(6 * ROW4 * 16), "n"(7 * ROW4 * 16), "n"(8 * ROW4 * 16), "n"(9 * ROW4 * 16),
              "n"(10 * ROW4 * 16), "n"(11 * ROW4 * 16), "n"(12 * ROW4 * 16), "n"(13 * ROW4 * 16),
              "n"(14 * ROW4 * 16), "n"(15 * ROW4 * 16)
            : "memory");
#undef PMR_STEP
      }
      // A triangle with more than kPieceRows pixels in the block was cut into several pieces: the sums of the
      // pieces that share this round are added up (highest slot first) so that the triangle costs one set of
      // atomics, not one per piece -- what matters when few vertices take all the traffic (large triangles).
      const unsigned merging = __ballot_sync(0xffffffffu, continues && slot > 0 && len > 0);
      if (merging) {
#pragma unroll
        for (int sl = SLOTS - 1; sl >= 1; --sl) {
          const float x0 = __shfl_sync(0xffffffffu, acc0, next_slot_lane), x1 = __shfl_sync(0xffffffffu, acc1, next_slot_lane);
          const float x2 = __shfl_sync(0xffffffffu, acc2, next_slot_lane), x3 = __shfl_sync(0xffffffffu, acc3, next_slot_lane);
          if (slot == sl - 1 && ((merging >> (kQuarters ? sl * 8 : sl * J)) & 1u)) { acc0 += x0; acc1 += x1; acc2 += x2; acc3 += x3; }
        }
      }
      if (len > 0 && !(continues && slot > 0)) {
        const int vtx = reinterpret_cast<const int *>(area.vids + first)[corner];
        const float sums[4] = {acc0, acc1, acc2, acc3};
        float *to_vert = dv_now + (unsigned)(vtx * 4 + column_of(u < 3 ? u : 0));   // column u of the vertex row
        float *to_attr = da_now + (unsigned)(vtx * A + u) - 3;                        // column u of the attribute row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i * U >= E) continue;                             // beyond the last column for every lane
          const int e = u + i * U;                              // column of the corner
          const bool is_vert = (i * U >= 3) ? false : ((U <= 3 && i == 0) ? true : e < 3);
          const bool exists = ((i + 1) * U <= E) ? true : e < E;
          if (is_vert) {
            // U < 3: columns x, y of i = 0 and w (or an attribute) of i = 1 ...: the row is walked in steps of U
            if (dv_now != nullptr) red_add(i == 0 ? to_vert : dv_now + (unsigned)(vtx * 4 + column_of(e)), sums[i]);
          } else if (exists) {
            if (da_now != nullptr) red_add(to_attr + i * U, sums[i]);
          }
        }
      }
    }
    __syncwarp();                                   // rows, pieces and vertex ids are rewritten by the next block
  }
}

// ---------------------------------------------------------------------------------------------
// Ordered (parity) mode: sort by vertex, fold in pixel order
// ---------------------------------------------------------------------------------------------
//
// The reference adds a pixel's contributions to its triangle's three vertices in row-major pixel order, corner
// 0..2 within a pixel, in fp32 (K.cpp:156-157, :232-269; the same order for d(attributes) through index_put_,
// rast.py:130-132 with one torch thread).  Per image that is a list of 3*H*W ENTRIES, entry e = 3*pixel + corner
// with key = the corner's vertex id (V for pixels that draw nothing), already in the order in which each
// vertex must receive its terms.  A STABLE sort of the entries by key therefore lays every vertex's terms out
// consecutively and in the reference's order:
//   1. least-significant-digit radix sort, 8 bits (or fewer) per pass, one warp per tile of kSortTile entries:
//      a counting kernel (per-tile digit histogram), an exclusive scan over (digit, tile) per image, and a
//      scatter kernel in which a warp walks its tile 32 entries at a time -- rank among equal digits of the
//      chunk by match.any, running per-digit offsets in shared memory -- so that equal keys keep their order.
//      The first pass reads its keys straight from the id / triangle buffers; nothing is staged for it.
//   2. segment bounds per (image, vertex) from the places where the sorted key changes;
//   3. one warp per (image, vertex) walks its segment 32 entries at a time: the lanes evaluate the entries'
//      terms in parallel (same per-pixel arithmetic as everywhere), park them in shared memory, and lane c folds
//      component c over the rows strictly in order.
// Cost is linear in the number of pixels for any mesh (the box-walking kernel this replaced was quadratic in
// the triangle size) and every load of the fold is a gather of a few neighbouring pixels.

constexpr int kSortTile = 2048;            // entries per warp and pass
constexpr int kSortWarps = 4;              // warps per CTA (independent)
constexpr int kSortBins = 256;

// key of entry e of image b, from the forward buffers (first pass) or from the previous pass
struct EntrySource {
  const uint2 *sorted;                     // {key, payload} of the previous pass, or nullptr: generate
  const int32_t *ids, *tris;
  const float *bary;
  int V;
};

__device__ __forceinline__ uint2 load_entry(const EntrySource &src, size_t image_entries, int b, unsigned e,
                                            size_t pixels_per_image) {
  if (src.sorted != nullptr) return src.sorted[(size_t)b * image_entries + e];
  const unsigned pixel = e / 3u, corner = e - 3u * pixel;
  const size_t p = (size_t)b * pixels_per_image + pixel;
  const int id = src.ids[p];
  bool covered = true;
  if (id == 0) {                           // K.cpp:162: id 0 with barycentrics summing below 0.9 draws nothing
    const float *bp = src.bary + 3 * p;
    covered = !(bp[0] + bp[1] + bp[2] < kDegenerateBarySum);
  }
  const unsigned key = covered ? (unsigned)__ldg(src.tris + 3 * (size_t)id + corner) : (unsigned)src.V;
  return make_uint2(key, e);
}

// counts[b][digit][tile]: how many entries of the tile have this digit
__global__ void __launch_bounds__(kSortWarps * 32)
sort_count_kernel(EntrySource src, unsigned entries, size_t pixels_per_image, int tiles, int shift, unsigned mask,
                  unsigned *__restrict__ counts) {
  __shared__ unsigned hist_all[kSortWarps][kSortBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * kSortWarps + warp, b = blockIdx.y;
  unsigned *hist = hist_all[warp];
  for (int i = lane; i < kSortBins; i += 32) hist[i] = 0u;
  __syncwarp();
  if (tile < tiles) {
    const unsigned begin = (unsigned)tile * kSortTile, end = min(begin + (unsigned)kSortTile, entries);
    for (unsigned e0 = begin; e0 < end; e0 += 32) {
      const unsigned e = e0 + lane;
      const bool live = e < end;
      const unsigned digit = live ? (load_entry(src, entries, b, e, pixels_per_image).x >> shift) & mask : 0xffffffffu - lane;
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      if (live && lane == __ffs(peers) - 1) hist[digit] += __popc(peers);
      __syncwarp();
    }
    for (int i = lane; i < kSortBins; i += 32) counts[((size_t)b * kSortBins + i) * tiles + tile] = hist[i];
  }
}

// exclusive prefix sum of counts[b][:] (digit-major, then tile): one CTA per image
__global__ void __launch_bounds__(1024)
sort_scan_kernel(unsigned *__restrict__ counts, int n) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry;
  unsigned *c = counts + (size_t)blockIdx.x * n;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024 * 4) {
    const int i0 = base + threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + k < n ? c[i0 + k] : 0u;
    const unsigned mine = v[0] + v[1] + v[2] + v[3];
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_sums[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += up;
      }
      warp_sums[lane] = wi - w;                       // exclusive over the warps
    }
    __syncthreads();
    unsigned at = carry + warp_sums[warp] + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) c[i0 + k] = at;
      at += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = at;             // the last thread holds the running total
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSortWarps * 32)
sort_scatter_kernel(EntrySource src, unsigned entries, size_t pixels_per_image, int tiles, int shift, unsigned mask,
                    const unsigned *__restrict__ offsets, uint2 *__restrict__ out) {
  __shared__ unsigned next_all[kSortWarps][kSortBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * kSortWarps + warp, b = blockIdx.y;
  if (tile >= tiles) return;
  unsigned *next = next_all[warp];
  for (int i = lane; i < kSortBins; i += 32) next[i] = offsets[((size_t)b * kSortBins + i) * tiles + tile];
  __syncwarp();
  uint2 *out_b = out + (size_t)b * entries;
  const unsigned begin = (unsigned)tile * kSortTile, end = min(begin + (unsigned)kSortTile, entries);
  for (unsigned e0 = begin; e0 < end; e0 += 32) {
    const unsigned e = e0 + lane;
    const bool live = e < end;
    uint2 entry = make_uint2(0u, 0u);
    if (live) entry = load_entry(src, entries, b, e, pixels_per_image);
    const unsigned digit = live ? (entry.x >> shift) & mask : 0xffffffffu - lane;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const unsigned before = __popc(peers & ((1u << lane) - 1u));      // equal digits in lower lanes: stable
    unsigned at = 0u;
    if (live) at = next[digit] + before;
    __syncwarp();
    if (live) {
      out_b[at] = entry;
      if (lane == __ffs(peers) - 1) next[digit] += __popc(peers);
    }
    __syncwarp();
  }
}

// bounds[b][v] = {first, last + 1} of vertex v's run in the sorted entries of image b (zeros: no entry)
__global__ void __launch_bounds__(256)
segment_bounds_kernel(const uint2 *__restrict__ sorted, unsigned entries, int V, uint2 *__restrict__ bounds) {
  const int b = blockIdx.y;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= entries) return;
  const uint2 *s = sorted + (size_t)b * entries;
  const unsigned key = s[i].x;
  const unsigned prev = i > 0 ? s[i - 1].x : 0xffffffffu;
  if (key != prev) {
    unsigned *bb = reinterpret_cast<unsigned *>(bounds + (size_t)b * (V + 1));
    bb[2 * key] = i;
    if (i > 0) bb[2 * prev + 1] = i;
  }
  if (i == entries - 1) reinterpret_cast<unsigned *>(bounds + (size_t)b * (V + 1))[2 * key + 1] = entries;
}

constexpr int kFoldWarps = 4;

// The terms ONE entry (pixel, corner j) adds to its vertex: the three vertex terms of corner j and, fused, the
// A products (g_a*alpha)*b_j -- the same expressions as vertex_terms() / the interpolation backward evaluate
// for all three corners, restricted to the corner that is asked for (same operations, same bits).
template <bool FUSED, int A_STATIC>
__device__ __forceinline__ void entry_terms(const float *__restrict__ verts_b, const float *__restrict__ attrs_b,
                                            const int32_t *__restrict__ tris, int id, const float bp[3],
                                            const float *__restrict__ g_p, int A_dyn, int j, float *row) {
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  int vid[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) vid[k] = __ldg(tris + 3 * (size_t)id + k);
  const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
  const float4 p0 = __ldg(v4 + vid[0]), p1 = __ldg(v4 + vid[1]), p2 = __ldg(v4 + vid[2]);
  const float bj = j == 0 ? bp[0] : (j == 1 ? bp[1] : bp[2]);
  float g[3];
  if (FUSED) {
    const float alpha = coverage_alpha(bp[0], bp[1], bp[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float *ck = attrs_b + (size_t)vid[k] * A;
      g[k] = torch_inner_sum(A, [&](int a) { return (__ldg(g_p + a) * alpha) * __ldg(ck + a); });
    }
    if (A_STATIC > 0) {
#pragma unroll
      for (int a = 0; a < (A_STATIC > 0 ? A_STATIC : 1); ++a) row[3 + a] = (__ldg(g_p + a) * alpha) * bj;
    } else {
      for (int a = 0; a < A; ++a) row[3 + a] = (__ldg(g_p + a) * alpha) * bj;
    }
  } else {
    g[0] = g_p[0]; g[1] = g_p[1]; g[2] = g_p[2];
  }
  float m[9];
  const float det = adjugate_signed(p0.x, p1.x, p2.x, p0.y, p1.y, p2.y, p0.w, p1.w, p2.w, m);
  const SharedDivisor by_det(fabsf(det));
#pragma unroll
  for (int c = 0; c < 3; ++c) {                      // K.cpp:187-269 for corner j
    const float sc = m[c] + m[3 + c] + m[6 + c];
    const float d0 = (-m[0 + c]) * bj + sc * bp[0] * bj;
    const float d1 = (-m[3 + c]) * bj + sc * bp[1] * bj;
    const float d2 = (-m[6 + c]) * bj + sc * bp[2] * bj;
    row[c] = by_det.divide(g[0] * d0 + g[1] * d1 + g[2] * d2);
  }
}

// One warp folds the entries of `vpw` consecutive vertices of one image (their runs are consecutive in the
// sorted list): 32 entries at a time, lanes evaluate the entries' terms into shared-memory rows, then lane c
// adds component c row by row IN ORDER, writing a vertex's sums out when the key changes.
template <bool FUSED, int A_STATIC>
__global__ void __launch_bounds__(kFoldWarps * 32)
backward_fold_kernel(const float *__restrict__ grad, const float *__restrict__ verts,
                     const float *__restrict__ attrs, const int32_t *__restrict__ tris,
                     const int32_t *__restrict__ ids, const float *__restrict__ bary,
                     const uint2 *__restrict__ sorted, const uint2 *__restrict__ bounds, unsigned entries,
                     int V, int A_dyn, size_t pixels_per_image, int vpw, int groups_per_image,
                     float *__restrict__ d_verts, float *__restrict__ d_attrs) {
  extern __shared__ float fold_smem[];   // [warp]: 32 rows of ncomp floats, then 32 keys
  const int A = A_STATIC > 0 ? A_STATIC : A_dyn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x * kFoldWarps + warp, b = blockIdx.y;
  if (group >= groups_per_image) return;
  const int ncomp = 3 + (FUSED ? A : 0);
  float *rows = fold_smem + (size_t)warp * (32 * ncomp + 32);
  unsigned *row_keys = reinterpret_cast<unsigned *>(rows + 32 * ncomp);
  const int v_first = group * vpw, v_count = min(vpw, V - v_first);
  const float *verts_b = verts + (size_t)b * V * 4;
  const float *attrs_b = FUSED ? attrs + (size_t)b * V * A : nullptr;
  const uint2 *sorted_b = sorted + (size_t)b * entries;
  float *dv_b = d_verts ? d_verts + (size_t)b * V * 4 : nullptr;
  float *da_b = (FUSED && d_attrs) ? d_attrs + (size_t)b * V * A : nullptr;

  // the stream of this warp: from the first present vertex's run to the last one's (vpw <= 32)
  uint2 mine = make_uint2(0u, 0u);
  if (lane < v_count) mine = bounds[(size_t)b * (V + 1) + v_first + lane];
  const bool present = mine.y > mine.x;
  const unsigned begin = __reduce_min_sync(0xffffffffu, present ? mine.x : 0xffffffffu);
  const unsigned end = __reduce_max_sync(0xffffffffu, present ? mine.y : 0u);
  // vertices without entries: zero gradient
  if (lane < v_count && !present) {
    const int v = v_first + lane;
    if (dv_b) *reinterpret_cast<float4 *>(dv_b + (size_t)v * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (da_b) for (int a = 0; a < A; ++a) da_b[(size_t)v * A + a] = 0.0f;
  }
  if (begin >= end) return;

  constexpr int kMaxSlots = 3;          // components lane, lane + 32, lane + 64 (A <= 93)
  float acc[kMaxSlots] = {0.0f, 0.0f, 0.0f};
  unsigned current = 0xffffffffu;       // vertex whose sums are in acc
  auto flush = [&](unsigned v) {
    if (dv_b != nullptr) {
      if (lane < 3) dv_b[(size_t)v * 4 + column_of(lane)] = acc[0];
      if (lane == 3) dv_b[(size_t)v * 4 + 2] = 0.0f;       // z column never receives gradient
    }
    if (da_b != nullptr) {
#pragma unroll
      for (int s2 = 0; s2 < kMaxSlots; ++s2) {
        const int c = lane + 32 * s2;
        if (c >= 3 && c < ncomp) da_b[(size_t)v * A + c - 3] = acc[s2];
      }
    }
#pragma unroll
    for (int s2 = 0; s2 < kMaxSlots; ++s2) acc[s2] = 0.0f;
  };
  // Software pipeline over the chunks of 32 entries: an entry leads to its pixel, the pixel to its triangle, the
  // triangle to its vertices -- four dependent loads.  The entry of chunk c+2 and the pixel data (id,
  // barycentrics) of chunk c+1 are in flight while chunk c is evaluated, so two links of the chain are hidden.
  auto load_entry_at = [&](unsigned k) { return k < end ? sorted_b[k] : make_uint2(0xffffffffu, 0u); };
  struct PixelHead { int id; float b0, b1, b2; };
  auto load_pixel_head = [&](const uint2 &entry) {
    PixelHead h = {0, 0.0f, 0.0f, 0.0f};
    if (entry.x != 0xffffffffu) {
      const size_t p = (size_t)b * pixels_per_image + entry.y / 3u;
      h.id = ids[p];
      h.b0 = bary[3 * p]; h.b1 = bary[3 * p + 1]; h.b2 = bary[3 * p + 2];
    }
    return h;
  };
  uint2 entry0 = load_entry_at(begin + lane), entry1 = load_entry_at(begin + 32 + lane);
  PixelHead head0 = load_pixel_head(entry0);
  for (unsigned k0 = begin; k0 < end; k0 += 32) {
    const int count = (int)min(32u, end - k0);
    const uint2 entry2 = load_entry_at(k0 + 64 + lane);
    const PixelHead head1 = load_pixel_head(entry1);
    if (entry0.x != 0xffffffffu) {
      const unsigned pixel = entry0.y / 3u;
      const int corner = (int)(entry0.y - 3u * pixel);
      const size_t p = (size_t)b * pixels_per_image + pixel;
      const float bp[3] = {head0.b0, head0.b1, head0.b2};
      entry_terms<FUSED, A_STATIC>(verts_b, attrs_b, tris, head0.id, bp, grad + p * (FUSED ? A : 3), A, corner,
                                   rows + (size_t)lane * ncomp);
      row_keys[lane] = entry0.x;
    }
    entry0 = entry1; entry1 = entry2; head0 = head1;
    __syncwarp();
    // fold strictly in entry order; the rows of one vertex within the chunk are fetched eight at a time so
    // that only the adds are serial
    int r = 0;
    while (r < count) {
      const unsigned key = row_keys[r];
      if (key != current) {
        if (current != 0xffffffffu) flush(current);
        current = key;
      }
      int run = 1;                       // rows r .. r+run-1 share the key (warp-uniform)
      while (r + run < count && row_keys[r + run] == key) ++run;
#pragma unroll
      for (int s2 = 0; s2 < kMaxSlots; ++s2) {
        const int c = lane + 32 * s2;
        if (c < ncomp) {
          const float *col = rows + (size_t)r * ncomp + c;
          int i = 0;
          for (; i + 8 <= run; i += 8) {
            float t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = col[(size_t)(i + u) * ncomp];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[s2] += t[u];
          }
          for (; i < run; ++i) acc[s2] += col[(size_t)i * ncomp];
        }
      }
      r += run;
    }
    __syncwarp();
  }
  if (current != 0xffffffffu) flush(current);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------

// Blocks per strip: long strips amortise the per-warp setup, but the grid must still fill the GPU
// (about 32 warps on each of the SMs, several times over).
static int strip_blocks_for(const Context *ctx, int W, int H, int B) {
  if (ctx->strip_blocks_override > 0) return ctx->strip_blocks_override;
  const long long block_rows = (long long)((H + 3) / 4) * B;
  // a dozen waves of 8 CTAs x 4 warps per SM: with fewer, the last, partly filled wave shows (c4 at 32 views
  // per GPU ran 6.9 waves with strips of 8)
  const long long wanted = 12LL * 32 * ctx->sm_count;
  for (int n = 8; n > 1; n >>= 1)
    if (block_rows * ((W + 8 * n - 1) / (8 * n)) >= wanted) return n;
  return 1;
}

static dim3 strip_grid(int W, int H, int B, int strip) {
  return dim3((W + 8 * strip - 1) / (8 * strip), (H + 4 * kStripWarps - 1) / (4 * kStripWarps), B);
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
    if (d_verts && fused && d_attrs == d_verts + (size_t)n_pairs * 4) {      // one allocation: one memset
      PMR_CUDA(ctx, cudaMemsetAsync(d_verts, 0, (size_t)n_pairs * (4 + A) * sizeof(float), stream));
    } else {
      if (d_verts) PMR_CUDA(ctx, cudaMemsetAsync(d_verts, 0, (size_t)n_pairs * 4 * sizeof(float), stream));
      if (fused && d_attrs) PMR_CUDA(ctx, cudaMemsetAsync(d_attrs, 0, (size_t)n_pairs * A * sizeof(float), stream));
    }
    if (total == 0 || T == 0) return PMR_OK;
    const int strip = strip_blocks_for(ctx, W, H, B);
    BlockMaps maps;
    const int use_tma = make_block_maps(ctx, &maps, ids, bary, grad, fused ? A : 3, B, W, H);
#define PMR_BLOCKS(F, AS)                                                                                 \
  backward_blocks_kernel<F, AS><<<strip_grid(W, H, B, strip), kStripWarps * 32, 0, stream>>>(            \
      maps, use_tma, grad, verts, attrs, tris, ids, bary, V, W, H, strip, d_verts, d_attrs);
    if (!fused) PMR_BLOCKS(false, 1)
    else if (A == 9) PMR_BLOCKS(true, 9)
    else if (A == 3) PMR_BLOCKS(true, 3)
    else if (A == 4) PMR_BLOCKS(true, 4)
    else if (A == 12) PMR_BLOCKS(true, 12)
    else if (A == 13) PMR_BLOCKS(true, 13)
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
  if (fused && A > 93) return set_error(ctx, PMR_ERR_SIZE, "ordered backward supports at most 93 attributes");
  if (total == 0 || T == 0) {
    if (d_verts) PMR_CUDA(ctx, cudaMemsetAsync(d_verts, 0, (size_t)n_pairs * 4 * sizeof(float), stream));
    if (fused && d_attrs) PMR_CUDA(ctx, cudaMemsetAsync(d_attrs, 0, (size_t)n_pairs * A * sizeof(float), stream));
    return PMR_OK;
  }
  if (3 * ppi >= (1LL << 32)) return set_error(ctx, PMR_ERR_SIZE, "ordered backward: image too large (3*H*W must fit 32 bits)");
  const unsigned entries = (unsigned)(3 * ppi);
  const int tiles = (int)((entries + kSortTile - 1) / kSortTile);
  int key_bits = 1;
  while ((1LL << key_bits) <= (long long)V) ++key_bits;           // keys are 0 .. V
  const int passes = (key_bits + 7) / 8;
  const int digit_bits = (key_bits + passes - 1) / passes;
  // workspace: two entry buffers, the per-tile digit counters, the segment bounds
  const size_t entry_bytes = (((size_t)B * entries * sizeof(uint2)) + 255) & ~(size_t)255;
  const size_t count_bytes = (((size_t)B * kSortBins * tiles * sizeof(unsigned)) + 255) & ~(size_t)255;
  const size_t bound_bytes = (size_t)B * (V + 1) * sizeof(uint2);
  int rc = ctx->scratch.reserve(ctx, 2 * entry_bytes + count_bytes + bound_bytes);
  if (rc) return rc;
  char *base = (char *)ctx->scratch.ptr;
  uint2 *buffers[2] = {(uint2 *)base, (uint2 *)(base + entry_bytes)};
  unsigned *counts = (unsigned *)(base + 2 * entry_bytes);
  uint2 *bounds = (uint2 *)(base + 2 * entry_bytes + count_bytes);
  EntrySource src;
  src.sorted = nullptr; src.ids = ids; src.tris = tris; src.bary = bary; src.V = V;
  const dim3 sort_grid((tiles + kSortWarps - 1) / kSortWarps, B);
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * digit_bits;
    const unsigned mask = (1u << digit_bits) - 1u;
    uint2 *out = buffers[pass & 1];
    sort_count_kernel<<<sort_grid, kSortWarps * 32, 0, stream>>>(src, entries, (size_t)ppi, tiles, shift, mask, counts);
    sort_scan_kernel<<<B, 1024, 0, stream>>>(counts, kSortBins * tiles);
    sort_scatter_kernel<<<sort_grid, kSortWarps * 32, 0, stream>>>(src, entries, (size_t)ppi, tiles, shift, mask, counts, out);
    ctx->launches += 3;
    src.sorted = out;
  }
  const uint2 *sorted = src.sorted;
  PMR_CUDA(ctx, cudaMemsetAsync(bounds, 0, bound_bytes, stream));
  segment_bounds_kernel<<<dim3((entries + 255) / 256, B), 256, 0, stream>>>(sorted, entries, V, bounds);
  ctx->launches += 1;
  const int ncomp = 3 + (fused ? A : 0);
  const size_t smem = (size_t)kFoldWarps * (32 * ncomp + 32) * sizeof(float);
  // vertices per warp: enough to keep the 32 lanes busy (about 512 entries per warp when the runs are short)
  long long per_vertex = 3 * ppi / (V > 0 ? V : 1);
  int vpw = (int)(512 / (per_vertex > 0 ? per_vertex : 1));
  vpw = vpw < 1 ? 1 : (vpw > 32 ? 32 : vpw);
  const int groups = (V + vpw - 1) / vpw;
  const dim3 fold_grid((groups + kFoldWarps - 1) / kFoldWarps, B);
#define PMR_FOLD(F, AS)                                                                                       \
  {                                                                                                           \
    PMR_CUDA(ctx, cudaFuncSetAttribute(backward_fold_kernel<F, AS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    backward_fold_kernel<F, AS><<<fold_grid, kFoldWarps * 32, smem, stream>>>(                                \
        grad, verts, attrs, tris, ids, bary, sorted, bounds, entries, V, A, (size_t)ppi, vpw, groups, d_verts, d_attrs); \
  }
  if (!fused) PMR_FOLD(false, 0)
  else if (A == 9) PMR_FOLD(true, 9)
  else if (A == 4) PMR_FOLD(true, 4)
  else PMR_FOLD(true, 0)
#undef PMR_FOLD
  ctx->launches += 1;
  return check_launch(ctx, "backward_fold_kernel");
}

}  // namespace pmr
