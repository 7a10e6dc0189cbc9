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
// One warp owns an 8x4 pixel block (the same shape the forward kernel rasterizes), one lane per
// pixel.  Every covered lane evaluates its NV = 9 (+3A) per-pixel sums (9 vertex terms, and
// g_a*alpha*b_k for every corner k and attribute a) and parks them in a shared-memory row.  Lanes
// are grouped by triangle id with match.any; the warp then walks the (group, column) pairs 32 at
// a time, each lane adding one column over the lanes of one group, and issues ONE atomic per pair:
// NV atomics per (block, triangle) instead of per pixel, and the lanes of one instruction hit
// consecutive words of the same vertex rows, which the memory system merges per 32-byte sector
// (profiles/microbench/atomics_bench.cu: 3.8x the lane rate of scattered atomics).

// SHADE (render path, A = 9): `grad` is d(RGBA) [B,H,W,4] with flipped rows; the pixel's nine interpolated
// channels are recomputed from the corner attributes (they were never stored) and the gradient passes through
// the diffuse + ambient lighting (shade_math.cuh) before it enters the interpolation backward.
template <bool FUSED, int A_STATIC, int kBlockWarps, bool SHADE = false>
__global__ void __launch_bounds__(kBlockWarps * 32, (kBlockWarps == 8 ? (SHADE ? 4 : 5) : (A_STATIC == 9 ? 10 : 8)))
backward_blocks_kernel(const float *__restrict__ grad, const float *__restrict__ verts,
                       const float *__restrict__ attrs, const int32_t *__restrict__ tris,
                       const int32_t *__restrict__ ids, const float *__restrict__ bary,
                       int V, int W, int H, float *__restrict__ d_verts, float *__restrict__ d_attrs,
                       const float *__restrict__ light_positions = nullptr,
                       const float *__restrict__ light_intensities = nullptr,
                       const float *__restrict__ ambient = nullptr, int L = 0,
                       const float *__restrict__ background = nullptr) {
  constexpr int A = A_STATIC;
  constexpr int NV = 9 + (FUSED ? 3 * A : 0);
  constexpr int STRIDE = (NV + 3) | 1;           // NV sums + 3 vertex ids, odd => conflict-free rows
  // per warp: 32 rows of NV sums + 3 vertex ids; the gradient staging area aliases the rows (it is
  // consumed into registers before the first row is written)
  __shared__ __align__(16) float rows_all[kBlockWarps][32 * STRIDE + 36];   // + padding read by column-less lanes
  __shared__ __align__(8) unsigned long long grad_ready[kBlockWarps];      // mbarriers of the bulk gradient loads
  static_assert(32 * STRIDE >= 32 * A, "gradient staging must fit the row area");

  __shared__ Lights lights;
  // The CTA covers 2 x (kBlockWarps/2) pixel blocks: 16 pixels wide, 2*kBlockWarps rows high.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  if (SHADE) load_lights(lights, light_positions, light_intensities, ambient, b, L);     // block barrier inside
  const int x0 = (blockIdx.x * 2 + (warp & 1)) * 8, y0 = (blockIdx.y * (kBlockWarps / 2) + (warp >> 1)) * 4;
  if (x0 >= W || y0 >= H) return;
  const int ix = x0 + (lane & 7), iy = y0 + (lane >> 3);
  const bool in_image = ix < W && iy < H;
  const long long p = ((long long)b * H + iy) * W + ix;

  int id = -1;
  float bp[3] = {0.0f, 0.0f, 0.0f};
  if (in_image) {
    id = ids[p];
    bp[0] = bary[3 * p]; bp[1] = bary[3 * p + 1]; bp[2] = bary[3 * p + 2];
    if (!pixel_is_covered(id, bp)) id = -1;
  }
  const unsigned covered = __ballot_sync(0xffffffffu, id >= 0);
  if (covered == 0u) return;

  float *rows = rows_all[warp];
  float g_local[FUSED ? A : 3];
  // Start the dependent chain ids -> triangle -> vertices first, so that its latency overlaps the
  // streaming loads of the gradient rows.
  const float *verts_b = verts + (size_t)b * V * 4;
  const float *attrs_b = FUSED ? attrs + (size_t)b * V * A : nullptr;
  int vid[3] = {0, 0, 0};
  if (id >= 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) vid[j] = __ldg(tris + 3 * (size_t)id + j);
  }
  float4 pv0 = make_float4(0.f, 0.f, 0.f, 0.f), pv1 = pv0, pv2 = pv0;
  if constexpr (SHADE) {
    if (id >= 0) {
      const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
      pv0 = __ldg(v4 + vid[0]); pv1 = __ldg(v4 + vid[1]); pv2 = __ldg(v4 + vid[2]);
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
      const float4 g4 = reinterpret_cast<const float4 *>(grad)[((size_t)b * H + (H - 1 - iy)) * W + ix];
      const float g[3] = {g4.x, g4.y, g4.z};
      shade_diffuse_pixel_backward(px, px + 3, px + 6, g, lights, L, ambient != nullptr, g_local, g_local + 3, g_local + 6);
    }
  } else if (FUSED) {
    // Stage the block's gradient rows through shared memory: a block row is 8*A contiguous
    // floats, read as float4 when the image rows keep them 16-byte aligned.
    float *stage = rows;
    if ((W & 7) == 0 && ((uintptr_t)grad & 15) == 0) {
      // The block's four gradient rows (8*A contiguous floats each, 16-byte aligned) are fetched by
      // the bulk-copy engine straight into shared memory (cp.async.bulk global -> shared::cta,
      // completion on a per-warp mbarrier; SASS UBLKCP): no registers held, no LDG/STS pairs issued
      // by the SM, and the copy is in flight while the vertices are gathered.
      const float *block_grad = grad + (((long long)b * H + y0) * W + x0) * A;   // first pixel of the block
      const int n_rows = min(4, H - y0);
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&grad_ready[warp]);
      if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(n_rows * 32 * A)) : "memory");
      }
      __syncwarp();                               // barrier initialised and armed before the copies are issued
      if (lane < n_rows) {                        // one row per lane, all issued by the same instruction
        const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + lane * 8 * A);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(block_grad + (size_t)lane * W * A), "r"((unsigned)(32 * A)), "r"(bar) : "memory");
      }
      if (id >= 0) {
        const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
        pv0 = __ldg(v4 + vid[0]); pv1 = __ldg(v4 + vid[1]); pv2 = __ldg(v4 + vid[2]);
      }
      __syncwarp();                               // the barrier is initialised before anyone polls it
      unsigned done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar) : "memory");
      }
#pragma unroll
      for (int a = 0; a < A; ++a) g_local[a] = stage[lane * A + a];
      __syncwarp();                               // staging consumed; the area becomes the rows
    } else {
      if (in_image) {
#pragma unroll
        for (int a = 0; a < A; ++a) g_local[a] = __ldg(grad + p * A + a);
      }
      if (id >= 0) {
        const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
        pv0 = __ldg(v4 + vid[0]); pv1 = __ldg(v4 + vid[1]); pv2 = __ldg(v4 + vid[2]);
      }
    }
  } else {
    if (in_image) { g_local[0] = grad[3 * p]; g_local[1] = grad[3 * p + 1]; g_local[2] = grad[3 * p + 2]; }
    if (id >= 0) {
      const float4 *v4 = reinterpret_cast<const float4 *>(verts_b);
      pv0 = __ldg(v4 + vid[0]); pv1 = __ldg(v4 + vid[1]); pv2 = __ldg(v4 + vid[2]);
    }
  }

  // Group the covered lanes by triangle and give every covered lane a ROW: the rows of one triangle
  // are consecutive (groups ordered by their first lane), so the reduction below is a linear walk.
  // Uncovered lanes get private keys, match nobody and own no row.
  const unsigned peers = __match_any_sync(0xffffffffu, id >= 0 ? id : -1 - lane);
  const int leader = __ffs(peers) - 1;
  const int group_size = __popc(peers), rank = __popc(peers & ((1u << lane) - 1u));
  const int lead_size = (id >= 0 && lane == leader) ? group_size : 0;
  const int before = warp_inclusive_scan(lead_size) - lead_size;        // rows of the groups led by lower lanes
  const int pos = __shfl_sync(0xffffffffu, before, leader) + rank;
  // bit r set: row r is the last row of its group
  unsigned ends = __reduce_or_sync(0xffffffffu, (id >= 0 && rank == group_size - 1) ? (1u << pos) : 0u);

  if (id >= 0) {
    PixelGrad pg;
    pixel_grad_loaded<FUSED>(vid, pv0, pv1, pv2, attrs_b, bp, g_local, A, pg);
    float *row = rows + pos * STRIDE;
#pragma unroll
    for (int k = 0; k < 9; ++k) row[k] = pg.terms[k];
    if (FUSED) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int a = 0; a < A; ++a) row[9 + k * A + a] = (g_local[a] * pg.alpha) * pg.b[k];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) row[NV + j] = __int_as_float(pg.vid[j]);
  }
  __syncwarp();

  // Lane c owns column c (and column c + 32 when NV > 32) of every group: it adds its column over the
  // group's rows and issues ONE atomic per (group, column).  Column -> (corner, destination) depends
  // on the lane only: columns 0..8 are the vertex terms [3*corner + {x,y,w}] -> d_verts[vtx*4 + {0,1,3}],
  // columns 9.. are [corner][attribute] -> d_attrs[vtx*A + attribute].  All lanes issue the same
  // instruction; consecutive lanes hit consecutive words of a vertex row.
  const int c0 = lane, c1 = lane + 32;
  const int corner0 = c0 < 9 ? c0 / 3 : (c0 - 9) / (A > 0 ? A : 1);
  const int corner1 = (c1 - 9) / (A > 0 ? A : 1);
  float *dv = d_verts ? d_verts + (size_t)b * V * 4 : nullptr;
  float *da = (FUSED && d_attrs) ? d_attrs + (size_t)b * V * A : nullptr;
  float *dst0 = c0 < 9 ? (dv ? dv + column_of(c0 % 3) : nullptr) : (da ? da + (c0 - 9) % (A > 0 ? A : 1) : nullptr);
  float *dst1 = da ? da + (c1 - 9) % (A > 0 ? A : 1) : nullptr;
  const int pitch0 = c0 < 9 ? 4 : A;
  if (c0 >= NV) dst0 = nullptr;
  if (c1 >= NV) dst1 = nullptr;
  // Lanes without a column read column 0 / a neighbouring row's words (the row area is padded) and
  // drop the sum, so that the row loads are unpredicated.
  const float *col = rows + (c0 < NV ? c0 : 0);
  const int *vtx0_at = reinterpret_cast<const int *>(rows) + NV + corner0;
  const int *vtx1_at = reinterpret_cast<const int *>(rows) + NV + (c1 < NV ? corner1 : 0);
  int r = 0;
  while (ends) {
    const int last = __ffs(ends) - 1;                   // rows r..last form one group (warp-uniform)
    ends &= ends - 1;
    const float *q = col + r * STRIDE;
    int n = last - r + 1;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
    for (; n > 8; --n, q += STRIDE) {                    // rare: more than 8 pixels of one triangle
      acc0 += q[0];
      if (NV > 32) acc1 += q[32];
    }
#define PMR_ROW(k) acc0 += q[(k) * STRIDE]; if (NV > 32) acc1 += q[(k) * STRIDE + 32];
    switch (n) {                                        // straight-line code per group size
      case 8: PMR_ROW(7)
      case 7: PMR_ROW(6)
      case 6: PMR_ROW(5)
      case 5: PMR_ROW(4)
      case 4: PMR_ROW(3)
      case 3: PMR_ROW(2)
      case 2: PMR_ROW(1)
      default: PMR_ROW(0)
    }
#undef PMR_ROW
    r = last + 1;
    red_add_if(dst0 != nullptr, dst0 + (unsigned)(vtx0_at[last * STRIDE] * pitch0), acc0);
    if (NV > 32) red_add_if(dst1 != nullptr, dst1 + (unsigned)(vtx1_at[last * STRIDE] * A), acc1);
  }
}

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
    backward_blocks_kernel<true, 9, 8, true><<<dim3((W + 15) / 16, (H + 15) / 16, B), 256, 0, stream>>>(
        shade->grad_rgba, verts, attrs, tris, ids, bary, V, W, H, d_verts, d_attrs, shade->light_positions,
        shade->light_intensities, shade->ambient, shade->L, shade->background);
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
#define PMR_BLOCKS(F, AS, WARPS)                                                                          \
  backward_blocks_kernel<F, AS, WARPS><<<dim3((W + 15) / 16, (H + 2 * WARPS - 1) / (2 * WARPS), B), WARPS * 32, 0, stream>>>(  \
      grad, verts, attrs, tris, ids, bary, V, W, H, d_verts, d_attrs)
    if (!fused) PMR_BLOCKS(false, 1, 8);
    else if (A == 9) PMR_BLOCKS(true, 9, 4);      // CTAs of 4 warps (16x8 pixels): 0.622 -> 0.601 ms on c2; 2 warps: 0.620
    else if (A == 4) PMR_BLOCKS(true, 4, 8);
    else if (A == 12) PMR_BLOCKS(true, 12, 4);
    else if (A == 13) PMR_BLOCKS(true, 13, 4);
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
  if (fused && A > 125) return set_error(ctx, PMR_ERR_SIZE, "ordered backward supports at most 125 attributes");
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
