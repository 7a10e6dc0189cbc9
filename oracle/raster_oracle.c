/*
 * raster_oracle.c -- CPU restatement of the reference rasterization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pytorch_mesh_renderer_b200/ may import, link
 * or execute this file.  It is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity pinning: this restatement is checked bit-for-bit (ids, barycentrics, z, vertex
 * gradients) against the reference's own kernel compiled from
 * /root/reference/src/mesh_renderer/kernels/rasterize_triangles.cpp (oracle/_ref, see
 * oracle/build_ref.py) by tests/test_oracle_pinning.py and against the committed vectors
 * the tests/golden npz vectors that were generated from that kernel and from the reference's
 * rasterize_clip_space (tests/golden/make_golden.py).
 *
 * Plain C99, one thread, no libm beyond floorf/ceilf/fabsf.  Must be compiled WITHOUT
 * floating-point contraction (-ffp-contract=off, no -march=native): the reference object
 * code has no FMA and every rounding below is significant.
 *
 * File:line citations are into /root/reference/src/mesh_renderer/ ("K.cpp" =
 * kernels/rasterize_triangles.cpp, "rast.py" = rasterize.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* K.cpp:13 -- a pixel with id 0 whose barycentrics sum below this is "no triangle". */
#define PMR_DEGENERATE_BARY_SUM 0.9f

/*
 * Adjugate of M = [[x0 x1 x2],[y0 y1 y2],[w0 w1 w2]] without the 1/det factor, sign
 * corrected so that a point inside the triangle sees non-negative edge values for either
 * winding.  K.cpp:61-87.  Each entry is (p*q) - (r*s) with both products rounded; the
 * determinant is the left-to-right sum (x0*adj[0] + x1*adj[3]) + x2*adj[6]; det == 0
 * leaves the sign alone.  Returns det (before the flip).
 */
static float adjugate_signed(const float x[3], const float y[3], const float w[3],
                             float adj[9])
{
    adj[0] = y[1] * w[2] - w[1] * y[2];
    adj[1] = x[2] * w[1] - w[2] * x[1];
    adj[2] = x[1] * y[2] - y[1] * x[2];
    adj[3] = y[2] * w[0] - w[2] * y[0];
    adj[4] = x[0] * w[2] - w[0] * x[2];
    adj[5] = x[2] * y[0] - y[2] * x[0];
    adj[6] = y[0] * w[1] - w[0] * y[1];
    adj[7] = x[1] * w[0] - w[1] * x[0];
    adj[8] = x[0] * y[1] - y[0] * x[1];
    const float det = x[0] * adj[0] + x[1] * adj[3] + x[2] * adj[6];
    if (det < 0.0f) {
        for (int k = 0; k < 9; ++k) adj[k] = -adj[k];
    }
    return det;
}

/* K.cpp:19-31: round outward, convert to int, clamp into [lo, hi]. */
static int clamp_floor_min3(float a, float b, float c, int lo, int hi)
{
    float m = a < b ? a : b;   /* std::min(std::min(a,b),c) */
    m = m < c ? m : c;
    int v = (int)floorf(m);
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

static int clamp_ceil_max3(float a, float b, float c, int lo, int hi)
{
    float m = a > b ? a : b;
    m = m > c ? m : c;
    int v = (int)ceilf(m);
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

/*
 * Forward pass for ONE image.  K.cpp:302-419.
 *   verts [V,4] clip-space xyzw, tris [T,3], outputs ids [H,W], bary [H,W,3], z [H,W].
 * Mixed precision exactly as the reference:
 *   - half extents are 0.5*W computed in double, stored as float (K.cpp:309-310);
 *   - bounding-box projection divides in float, then adds 1.0 and scales in double, then
 *     rounds to float (K.cpp:361-366);
 *   - pixel centres are ((i + 0.5) / half) - 1.0 in double, rounded to float (K.cpp:376-377).
 * Depth rule (K.cpp:401): reject z < -1, z > 1 or z > zbuf; an equal z therefore
 * overwrites, so among equal depths the largest triangle id wins.
 * If counters != NULL it receives {N_bbox, N_inside, N_zpass} (pixel visits, inside-test
 * passes, depth-test passes) for the flop accounting of BASELINE.md section 4.
 */
void pmr_oracle_forward(const float *verts, const int32_t *tris, int T, int W, int H,
                        int32_t *ids, float *bary, float *z, int64_t *counters)
{
    const float half_w = (float)(0.5 * W);
    const float half_h = (float)(0.5 * H);
    const size_t P = (size_t)W * (size_t)H;
    int64_t n_bbox = 0, n_inside = 0, n_zpass = 0;

    memset(ids, 0, P * sizeof(int32_t));            /* K.cpp:313-315 */
    memset(bary, 0, P * 3 * sizeof(float));         /* K.cpp:316-318 */
    for (size_t p = 0; p < P; ++p) z[p] = 1.0f;     /* K.cpp:319-321 */

    /* Pixel-centre tables: same value for every triangle, so hoisting is exact. */
    float *cx = (float *)malloc(sizeof(float) * (size_t)(W > 0 ? W : 1));
    float *cy = (float *)malloc(sizeof(float) * (size_t)(H > 0 ? H : 1));
    for (int ix = 0; ix < W; ++ix) cx[ix] = (float)(((ix + 0.5) / half_w) - 1.0);
    for (int iy = 0; iy < H; ++iy) cy[iy] = (float)(((iy + 0.5) / half_h) - 1.0);

    for (int t = 0; t < T; ++t) {                   /* K.cpp:330 ascending id */
        const float *p0 = verts + 4 * (size_t)tris[3 * t + 0];
        const float *p1 = verts + 4 * (size_t)tris[3 * t + 1];
        const float *p2 = verts + 4 * (size_t)tris[3 * t + 2];
        const float x[3] = {p0[0], p1[0], p2[0]};
        const float y[3] = {p0[1], p1[1], p2[1]};
        const float zc[3] = {p0[2], p1[2], p2[2]};
        const float w[3] = {p0[3], p1[3], p2[3]};

        if (w[0] < 0 && w[1] < 0 && w[2] < 0) continue;   /* K.cpp:339 */

        float adj[9];
        adjugate_signed(x, y, w, adj);                    /* K.cpp:350-353 */

        int left = 0, right = W, bottom = 0, top = H;     /* K.cpp:356 */
        if (w[0] > 0 && w[1] > 0 && w[2] > 0) {           /* K.cpp:360-371 */
            const float sx0 = (float)(((double)(x[0] / w[0]) + 1.0) * (double)half_w);
            const float sx1 = (float)(((double)(x[1] / w[1]) + 1.0) * (double)half_w);
            const float sx2 = (float)(((double)(x[2] / w[2]) + 1.0) * (double)half_w);
            const float sy0 = (float)(((double)(y[0] / w[0]) + 1.0) * (double)half_h);
            const float sy1 = (float)(((double)(y[1] / w[1]) + 1.0) * (double)half_h);
            const float sy2 = (float)(((double)(y[2] / w[2]) + 1.0) * (double)half_h);
            left = clamp_floor_min3(sx0, sx1, sx2, 0, W);
            right = clamp_ceil_max3(sx0, sx1, sx2, 0, W);
            bottom = clamp_floor_min3(sy0, sy1, sy2, 0, H);
            top = clamp_ceil_max3(sy0, sy1, sy2, 0, H);
        }

        for (int iy = bottom; iy < top; ++iy) {           /* K.cpp:374-375 */
            const float py = cy[iy];
            for (int ix = left; ix < right; ++ix) {
                const float px = cx[ix];
                ++n_bbox;
                /* K.cpp:39-48: e_i = ((a*px) + (b*py)) + c, four roundings. */
                const float e0 = adj[0] * px + adj[1] * py + adj[2];
                const float e1 = adj[3] * px + adj[4] * py + adj[5];
                const float e2 = adj[6] * px + adj[7] * py + adj[8];
                /* K.cpp:93-98: all >= 0 and not all zero. */
                if (!(e0 >= 0 && e1 >= 0 && e2 >= 0)) continue;
                if (!(e0 > 0 || e1 > 0 || e2 > 0)) continue;
                ++n_inside;

                const float esum = e0 + e1 + e2;          /* K.cpp:384 */
                const float b0 = e0 / esum;               /* K.cpp:385-387 */
                const float b1 = e1 / esum;
                const float b2 = e2 / esum;
                const float cz = b0 * zc[0] + b1 * zc[1] + b2 * zc[2];  /* K.cpp:395 */
                const float cw = b0 * w[0] + b1 * w[1] + b2 * w[2];     /* K.cpp:396 */
                const float depth = cz / cw;                            /* K.cpp:397 */

                const size_t p = (size_t)iy * (size_t)W + (size_t)ix;
                if (depth < -1.0f || depth > 1.0f || depth > z[p]) continue;  /* K.cpp:401 */
                ++n_zpass;
                ids[p] = t;                                /* K.cpp:405-409 */
                z[p] = depth;
                bary[3 * p + 0] = b0;
                bary[3 * p + 1] = b1;
                bary[3 * p + 2] = b2;
            }
        }
    }
    free(cx);
    free(cy);
    if (counters) {
        counters[0] = n_bbox;
        counters[1] = n_inside;
        counters[2] = n_zpass;
    }
}

/*
 * The nine per-pixel vertex-gradient terms of K.cpp:180-269 for one covered pixel.
 * out[3*j + c] is the contribution to vertex j of the triangle, component c in
 * {x, y, w}.  Operation order is the reference's:
 *   s_c      = (adj[c] + adj[3+c]) + adj[6+c]                       K.cpp:187-198
 *   d(i,c,j) = ((-adj[3i+c]) * b_j) + ((s_c * b_i) * b_j)           K.cpp:202-230
 *   out      = ((g0*d(0,c,j) + g1*d(1,c,j)) + g2*d(2,c,j)) / |det|  K.cpp:232-269
 */
static void pixel_vertex_terms(const float x[3], const float y[3], const float w[3],
                               const float b[3], const float g[3], float out[9])
{
    float adj[9];
    const float abs_det = fabsf(adjugate_signed(x, y, w, adj));
    for (int c = 0; c < 3; ++c) {
        const float s = adj[c] + adj[3 + c] + adj[6 + c];
        for (int j = 0; j < 3; ++j) {
            const float d0 = (-adj[0 + c]) * b[j] + s * b[0] * b[j];
            const float d1 = (-adj[3 + c]) * b[j] + s * b[1] * b[j];
            const float d2 = (-adj[6 + c]) * b[j] + s * b[2] * b[j];
            out[3 * j + c] = (g[0] * d0 + g[1] * d1 + g[2] * d2) / abs_det;
        }
    }
}

/*
 * Backward pass for ONE image.  K.cpp:131-273.  dverts [V,4] is zeroed here; only
 * columns 0 (x), 1 (y), 3 (w) receive gradient.  Pixels are visited row-major and each
 * term is added in fp32 in that order (K.cpp:156-157, 232-269) -- the summation order is
 * part of the reference's result.
 * dbary may be strided: element (p, k) is dbary[p*dbary_pixel_stride + k].
 */
void pmr_oracle_backward(const float *dbary, int64_t dbary_pixel_stride,
                         const float *verts, const int32_t *tris,
                         const int32_t *ids, const float *bary,
                         int V, int W, int H, float *dverts)
{
    static const int column_of[3] = {0, 1, 3};
    const size_t P = (size_t)W * (size_t)H;
    memset(dverts, 0, (size_t)V * 4 * sizeof(float));     /* K.cpp:144-146 */
    for (size_t p = 0; p < P; ++p) {
        const int32_t t = ids[p];
        const float b[3] = {bary[3 * p], bary[3 * p + 1], bary[3 * p + 2]};
        if (t == 0 && b[0] + b[1] + b[2] < PMR_DEGENERATE_BARY_SUM) continue;  /* K.cpp:162 */
        const int32_t vid[3] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
        float x[3], y[3], w[3];
        for (int j = 0; j < 3; ++j) {                     /* K.cpp:166-178 */
            x[j] = verts[4 * (size_t)vid[j] + 0];
            y[j] = verts[4 * (size_t)vid[j] + 1];
            w[j] = verts[4 * (size_t)vid[j] + 3];
        }
        const float *gp = dbary + (size_t)p * (size_t)dbary_pixel_stride;
        const float g[3] = {gp[0], gp[1], gp[2]};
        float terms[9];
        pixel_vertex_terms(x, y, w, b, g, terms);
        for (int j = 0; j < 3; ++j)
            for (int c = 0; c < 3; ++c)
                dverts[4 * (size_t)vid[j] + column_of[c]] += terms[3 * j + c];
    }
}

/*
 * Same per-pixel fp32 terms, accumulated in double.  Not the reference's result: it is
 * the yardstick for SURVEY.md F5 (how far any fp32 summation order, the reference's
 * included, sits from the exactly-summed value).
 */
void pmr_oracle_backward_f64acc(const float *dbary, int64_t dbary_pixel_stride,
                                const float *verts, const int32_t *tris,
                                const int32_t *ids, const float *bary,
                                int V, int W, int H, double *dverts)
{
    static const int column_of[3] = {0, 1, 3};
    const size_t P = (size_t)W * (size_t)H;
    memset(dverts, 0, (size_t)V * 4 * sizeof(double));
    for (size_t p = 0; p < P; ++p) {
        const int32_t t = ids[p];
        const float b[3] = {bary[3 * p], bary[3 * p + 1], bary[3 * p + 2]};
        if (t == 0 && b[0] + b[1] + b[2] < PMR_DEGENERATE_BARY_SUM) continue;
        const int32_t vid[3] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
        float x[3], y[3], w[3];
        for (int j = 0; j < 3; ++j) {
            x[j] = verts[4 * (size_t)vid[j] + 0];
            y[j] = verts[4 * (size_t)vid[j] + 1];
            w[j] = verts[4 * (size_t)vid[j] + 3];
        }
        const float *gp = dbary + (size_t)p * (size_t)dbary_pixel_stride;
        const float g[3] = {gp[0], gp[1], gp[2]};
        float terms[9];
        pixel_vertex_terms(x, y, w, b, g, terms);
        for (int j = 0; j < 3; ++j)
            for (int c = 0; c < 3; ++c)
                dverts[4 * (size_t)vid[j] + column_of[c]] += (double)terms[3 * j + c];
    }
}

/*
 * Attribute interpolation for ONE image, the torch-op chain of rast.py:118-150 written as
 * scalar loops:
 *   corner_k = attrs[tris[id][k]]                       rast.py:118-132
 *   img_a    = ((corner_0a*b0) + (corner_1a*b1)) + (corner_2a*b2)   rast.py:137-139
 *   alpha    = clamp(((2*b0) + (2*b1)) + (2*b2), 0, 1)  rast.py:145-146
 *   out_a    = (alpha*img_a) + ((1-alpha)*bg_a)         rast.py:149-150
 * Uncovered pixels carry id 0 and b = 0, so they read triangle 0's corners times zero and
 * come out as exactly the background.
 */
void pmr_oracle_interp_forward(const float *attrs, const int32_t *tris,
                               const int32_t *ids, const float *bary, const float *bg,
                               int A, int64_t P, float *out)
{
    for (int64_t p = 0; p < P; ++p) {
        const int32_t t = ids[p];
        const float b0 = bary[3 * p], b1 = bary[3 * p + 1], b2 = bary[3 * p + 2];
        const float *c0 = attrs + (size_t)A * (size_t)tris[3 * t + 0];
        const float *c1 = attrs + (size_t)A * (size_t)tris[3 * t + 1];
        const float *c2 = attrs + (size_t)A * (size_t)tris[3 * t + 2];
        float alpha = 2.0f * b0 + 2.0f * b1 + 2.0f * b2;
        alpha = alpha < 0.0f ? 0.0f : (alpha > 1.0f ? 1.0f : alpha);
        const float one_minus = 1.0f - alpha;
        for (int a = 0; a < A; ++a) {
            const float img = c0[a] * b0 + c1[a] * b1 + c2[a] * b2;
            out[(size_t)p * A + a] = alpha * img + one_minus * bg[a];
        }
    }
}

/*
 * Sum of n contiguous fp32 terms in the order torch's CPU sum kernel uses for a reduction
 * over the innermost contiguous dimension (torch 2.x, AVX2 dispatch, 8 float lanes;
 * ATen/native/cpu/SumKernel.cpp: vectorized_inner_sum / scalar_inner_sum / row_sum with
 * ilp_factor 4; valid while n/8 (or n) stays below the 16-element cascade step, i.e.
 * n < 512).  This is the order in which the reference's autograd reduces
 * d(weighted)/d(bary) over the attribute axis (backward of torch.mul at rast.py:137-138),
 * probed bit-for-bit for A in {3,4,5,8,9,12,13,16,17,32,33}.  Not part of the reference
 * source: it is the behaviour of the torch build the reference runs on.
 */
static float torch_inner_sum(const float *v, int n)
{
    if (n < 8) {
        /* scalar path: four interleaved partial sums, tail into partial 0 */
        float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const int q = n / 4;
        for (int i = 0; i < q; ++i)
            for (int k = 0; k < 4; ++k) part[k] += v[4 * i + k];
        for (int i = 4 * q; i < n; ++i) part[0] += v[i];
        part[0] += part[1];
        part[0] += part[2];
        part[0] += part[3];
        return part[0];
    }
    /* vector path: 8 lanes; the vectors themselves are summed with the same 4-way ilp */
    float lanes[4][8];
    memset(lanes, 0, sizeof(lanes));
    const int nvec = n / 8;
    const int q = nvec / 4;
    for (int i = 0; i < q; ++i)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) lanes[k][l] += v[8 * (4 * i + k) + l];
    for (int i = 4 * q; i < nvec; ++i)
        for (int l = 0; l < 8; ++l) lanes[0][l] += v[8 * i + l];
    for (int k = 1; k < 4; ++k)
        for (int l = 0; l < 8; ++l) lanes[0][l] += lanes[k][l];
    float acc = 0.0f;
    for (int i = 8 * nvec; i < n; ++i) acc += v[i];
    for (int l = 0; l < 8; ++l) acc += lanes[0][l];
    return acc;
}

/*
 * Autograd of the chain above for ONE image, in the order a single-threaded torch run
 * produces it:
 *   d_img_a     = g_a * alpha
 *   d_corner_ka = d_img_a * b_k, index_put(accumulate) into attrs row tris[id][k],
 *                 visiting pixels ascending, corners 0..2 (rast.py:130-132 backward)
 *   d_b_k       = sum_a d_img_a * corner_ka, reduced in torch's inner-sum order (see
 *                 torch_inner_sum)  (+ 2*d_alpha where the clamp is not
 *                 saturated: 0 <= 2*sum(b) <= 1, i.e. only on uncovered pixels, whose
 *                 d_b the rasterizer backward then ignores, K.cpp:162)
 * dattrs [V,A] is zeroed here; dbary [P,3] is fully written.
 */
void pmr_oracle_interp_backward(const float *g, const float *attrs, const int32_t *tris,
                                const int32_t *ids, const float *bary, const float *bg,
                                int V, int A, int64_t P, float *dattrs, float *dbary)
{
    float *prod = (float *)malloc(sizeof(float) * (size_t)(A > 0 ? A : 1));
    memset(dattrs, 0, (size_t)V * (size_t)A * sizeof(float));
    for (int64_t p = 0; p < P; ++p) {
        const int32_t t = ids[p];
        const float b[3] = {bary[3 * p], bary[3 * p + 1], bary[3 * p + 2]};
        const float s = 2.0f * b[0] + 2.0f * b[1] + 2.0f * b[2];
        const float alpha = s < 0.0f ? 0.0f : (s > 1.0f ? 1.0f : s);
        const float *gp = g + (size_t)p * A;
        float d_alpha = 0.0f;
        if (s >= 0.0f && s <= 1.0f) {
            /* d out / d alpha = img_a - bg_a, through the two products of rast.py:149-150 */
            float from_img = 0.0f, from_bg = 0.0f;
            for (int a = 0; a < A; ++a) {
                const float *c0 = attrs + (size_t)A * (size_t)tris[3 * t + 0];
                const float *c1 = attrs + (size_t)A * (size_t)tris[3 * t + 1];
                const float *c2 = attrs + (size_t)A * (size_t)tris[3 * t + 2];
                const float img = c0[a] * b[0] + c1[a] * b[1] + c2[a] * b[2];
                from_img += gp[a] * img;
                from_bg += gp[a] * bg[a];
            }
            d_alpha = from_img - from_bg;
        }
        for (int k = 0; k < 3; ++k) {
            const size_t row = (size_t)tris[3 * t + k];
            const float *ck = attrs + (size_t)A * row;
            for (int a = 0; a < A; ++a) {
                const float d_img = gp[a] * alpha;
                dattrs[row * A + a] += d_img * b[k];
                prod[a] = d_img * ck[a];
            }
            dbary[3 * p + k] = torch_inner_sum(prod, A) + 2.0f * d_alpha;
        }
    }
    free(prod);
}

/*
 * compute_vertex_normals for ONE mesh -- src/common/meshes.py:3-35.
 *
 * The reference runs three index_add_ passes, one per triangle corner c (meshes.py:23-33), each over the
 * triangles in ascending order, adding cross(p[c+1] - p[c], p[c+2] - p[c]) (corner indices mod 3) to the row
 * of vertex triangles[t][c]; then torch.nn.functional.normalize(eps=1e-6, p=2, dim=-1) (meshes.py:34).
 * torch's CPU kernels contract: cross = fma(a_p, b_q, -(a_q * b_p)) (the second product rounded first) and
 * the squared norm of a contiguous row = fma(z, z, fma(y, y, x * x)) -- probed on torch 2.11 and pinned against
 * the reference function itself by tests/test_oracle_golden.py.  This file is compiled with
 * -ffp-contract=off, so the fused operations are the explicit fmaf calls and nothing else.
 * raw [V,3] (may be NULL) receives the un-normalised sums.
 */
static void cross_torch(const float *a, const float *b, float *o)
{
    o[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    o[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    o[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}

void pmr_oracle_vertex_normals(const float *verts, const int32_t *tris, int V, int T, float *raw, float *normals)
{
    float *sum = (float *)calloc((size_t)(V > 0 ? V : 1) * 3, sizeof(float));
    for (int c = 0; c < 3; ++c) {
        for (int t = 0; t < T; ++t) {
            const int32_t i0 = tris[3 * t + c], i1 = tris[3 * t + (c + 1) % 3], i2 = tris[3 * t + (c + 2) % 3];
            float a[3], b[3], n[3];
            for (int k = 0; k < 3; ++k) {
                a[k] = verts[3 * (size_t)i1 + k] - verts[3 * (size_t)i0 + k];
                b[k] = verts[3 * (size_t)i2 + k] - verts[3 * (size_t)i0 + k];
            }
            cross_torch(a, b, n);
            for (int k = 0; k < 3; ++k) sum[3 * (size_t)i0 + k] += n[k];
        }
    }
    for (int v = 0; v < V; ++v) {
        const float *s = sum + 3 * (size_t)v;
        const float norm = sqrtf(fmaf(s[2], s[2], fmaf(s[1], s[1], s[0] * s[0])));
        const float d = norm > 1e-6f ? norm : 1e-6f;
        for (int k = 0; k < 3; ++k) {
            if (raw) raw[3 * (size_t)v + k] = s[k];
            normals[3 * (size_t)v + k] = s[k] / d;
        }
    }
    free(sum);
}
