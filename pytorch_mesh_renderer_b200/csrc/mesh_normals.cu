// mesh_normals.cu -- vertex normals of a triangle mesh (reference src/common/meshes.py:3-35
// compute_vertex_normals) and their backward, for fitting loops that move the geometry every step
// (SURVEY.md section 8f row 4).
//
// The reference adds, per image, three index_add_ passes (one per triangle corner, each in ascending
// triangle order) of the corner's cross product into the vertex rows, then normalises.  Written as a
// scatter that is the same contention pattern as the rasterizer's backward.  Here the mesh topology is
// turned ONCE into a vertex -> (corner, triangle) incidence table whose rows are sorted by (corner,
// triangle) -- exactly the order in which the reference's three passes touch a vertex -- and both
// directions become gathers: one thread per (image, vertex), no atomics, no zero-fill, and the fp32 sums
// are bit-reproducible and in the reference's order.
//
// Arithmetic follows what torch's CPU kernels do (probed; the test oracle restates it and is pinned against the
// reference function): cross = fma(a_p, b_q, -(a_q * b_p)), |n|^2 = fma(z, z, fma(y, y, x * x)),
// n / max(|n|, 1e-6).  The library is compiled with -fmad=false, so the fused operations are explicit.
#include "pmr_internal.cuh"

namespace pmr {

constexpr int kCornerShift = 30;                       // incidence code = corner << 30 | triangle
constexpr unsigned kTriangleMask = (1u << kCornerShift) - 1u;
constexpr float kNormalizeEps = 1e-6f;                 // meshes.py:34

// ---- topology: vertex -> incident (corner, triangle) table ------------------------------------------

__global__ void __launch_bounds__(256)
incidence_count_kernel(const int *__restrict__ tris, int n_corners, int V, int *__restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_corners) return;
  const int v = __ldg(tris + i);
  if (v >= 0 && v < V) atomicAdd(counts + v, 1);
}

// Exclusive scan of counts[0..V) into offsets[0..V] and cursor[0..V); one CTA walks the array in chunks.
__global__ void __launch_bounds__(1024)
incidence_scan_kernel(const int *__restrict__ counts, int V, int *__restrict__ offsets, int *__restrict__ cursor) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < V; base += 1024) {
    const int i = base + threadIdx.x;
    const int c = i < V ? counts[i] : 0;
    int incl = c;
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += up;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (warp ? warp_sums[warp - 1] : 0) + incl - c;
    if (i < V) {
      offsets[i] = excl;
      cursor[i] = excl;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[V] = carry_s;
}

__global__ void __launch_bounds__(256)
incidence_fill_kernel(const int *__restrict__ tris, int n_corners, int V, int *__restrict__ cursor,
                      unsigned *__restrict__ unsorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_corners) return;
  const int v = __ldg(tris + i);
  if (v < 0 || v >= V) return;
  const unsigned t = (unsigned)i / 3u, c = (unsigned)i - 3u * t;
  unsorted[atomicAdd(cursor + v, 1)] = (c << kCornerShift) | t;
}

// Every (corner, triangle) entry finds its rank inside its vertex's row by counting the smaller codes:
// quadratic in the valence, but spread over the row's own entries (a pole of a UV sphere has hundreds).
__global__ void __launch_bounds__(256)
incidence_rank_kernel(const int *__restrict__ tris, int n_corners, int V, const int *__restrict__ offsets,
                      const unsigned *__restrict__ unsorted, unsigned *__restrict__ incidence) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_corners) return;
  const int v = __ldg(tris + i);
  if (v < 0 || v >= V) return;
  const unsigned t = (unsigned)i / 3u, c = (unsigned)i - 3u * t;
  const unsigned code = (c << kCornerShift) | t;
  const int begin = offsets[v], end = offsets[v + 1];
  int rank = 0;
  for (int e = begin; e < end; ++e) rank += unsorted[e] < code;
  incidence[begin + rank] = code;
}

// ---- arithmetic shared by both directions -------------------------------------------------------------

struct Vec3 {
  float x, y, z;
};

__device__ __forceinline__ Vec3 load3(const float *p) { return {__ldg(p), __ldg(p + 1), __ldg(p + 2)}; }
__device__ __forceinline__ Vec3 sub3(const Vec3 &a, const Vec3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 add3(const Vec3 &a, const Vec3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }

// torch.cross on the CPU (CrossKernel.cpp, compiled with FMA contraction): a_p * b_q - a_q * b_p with the
// second product rounded first.
__device__ __forceinline__ Vec3 cross_torch(const Vec3 &a, const Vec3 &b) {
  Vec3 o;
  o.x = __fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y));
  o.y = __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z));
  o.z = __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x));
  return o;
}

__device__ __forceinline__ float norm_torch(const Vec3 &s) {
  return __fsqrt_rn(__fmaf_rn(s.z, s.z, __fmaf_rn(s.y, s.y, __fmul_rn(s.x, s.x))));
}

// ---- forward ---------------------------------------------------------------------------------------------

// One thread per (image, vertex): the row of the incidence table is walked in (corner, triangle) order.
__global__ void __launch_bounds__(256)
vertex_normals_forward_kernel(const float *__restrict__ verts, const int *__restrict__ tris,
                              const int *__restrict__ offsets, const unsigned *__restrict__ incidence, int V,
                              float *__restrict__ raw, float *__restrict__ normals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float *vb = verts + (size_t)blockIdx.y * V * 3;
  const Vec3 p = load3(vb + (size_t)v * 3);
  Vec3 sum = {0.0f, 0.0f, 0.0f};
  const int end = offsets[v + 1];
  for (int e = offsets[v]; e < end; ++e) {
    const unsigned code = __ldg(incidence + e);
    const unsigned c = code >> kCornerShift, t = code & kTriangleMask;
    const int *tri = tris + 3 * (size_t)t;
    const int i1 = __ldg(tri + (c == 2 ? 0 : c + 1)), i2 = __ldg(tri + (c == 0 ? 2 : c - 1));
    const Vec3 n = cross_torch(sub3(load3(vb + (size_t)i1 * 3), p), sub3(load3(vb + (size_t)i2 * 3), p));
    sum = add3(sum, n);
  }
  const size_t at = ((size_t)blockIdx.y * V + v) * 3;
  if (raw != nullptr) {
    raw[at] = sum.x; raw[at + 1] = sum.y; raw[at + 2] = sum.z;
  }
  const float denom = fmaxf(norm_torch(sum), kNormalizeEps);
  normals[at] = sum.x / denom; normals[at + 1] = sum.y / denom; normals[at + 2] = sum.z / denom;
}

// ---- backward --------------------------------------------------------------------------------------------

// Through the normalisation (torch autograd of x / clamp_min(|x|, eps)): g / d - x (g.x) / (d^2 |x|) where the
// clamp is inactive, g / d where it is active.
__global__ void __launch_bounds__(256)
vertex_normals_unnormalize_kernel(const float *__restrict__ grad_normals, const float *__restrict__ raw,
                                  long long n_rows, float *__restrict__ grad_raw) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const Vec3 s = load3(raw + r * 3), g = load3(grad_normals + r * 3);
  const float norm = norm_torch(s);
  const float d = fmaxf(norm, kNormalizeEps);
  Vec3 o = {g.x / d, g.y / d, g.z / d};
  if (norm >= kNormalizeEps && norm > 0.0f) {
    const float dd = d * d;
    const float grad_d = -(g.x * s.x / dd + g.y * s.y / dd + g.z * s.z / dd);
    const float scale = grad_d / norm;
    o.x += scale * s.x; o.y += scale * s.y; o.z += scale * s.z;
  }
  grad_raw[r * 3] = o.x; grad_raw[r * 3 + 1] = o.y; grad_raw[r * 3 + 2] = o.z;
}

// d p_c = sum over incident (t, c) of (G_0 + G_1 + G_2) x (p_{c+2} - p_{c+1}), G_k the gradient of the raw
// normal of corner k's vertex: the three cross products of a triangle collapse into one per incident corner.
__global__ void __launch_bounds__(256)
vertex_normals_backward_kernel(const float *__restrict__ grad_raw, const float *__restrict__ verts,
                               const int *__restrict__ tris, const int *__restrict__ offsets,
                               const unsigned *__restrict__ incidence, int V, float *__restrict__ d_verts) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const size_t image = (size_t)blockIdx.y * V * 3;
  const float *vb = verts + image, *gb = grad_raw + image;
  const Vec3 g_own = load3(gb + (size_t)v * 3);
  Vec3 sum = {0.0f, 0.0f, 0.0f};
  const int end = offsets[v + 1];
  for (int e = offsets[v]; e < end; ++e) {
    const unsigned code = __ldg(incidence + e);
    const unsigned c = code >> kCornerShift, t = code & kTriangleMask;
    const int *tri = tris + 3 * (size_t)t;
    const int i1 = __ldg(tri + (c == 2 ? 0 : c + 1)), i2 = __ldg(tri + (c == 0 ? 2 : c - 1));
    const Vec3 g = add3(add3(g_own, load3(gb + (size_t)i1 * 3)), load3(gb + (size_t)i2 * 3));
    const Vec3 edge = sub3(load3(vb + (size_t)i2 * 3), load3(vb + (size_t)i1 * 3));
    Vec3 n;
    n.x = g.y * edge.z - g.z * edge.y;
    n.y = g.z * edge.x - g.x * edge.z;
    n.z = g.x * edge.y - g.y * edge.x;
    sum = add3(sum, n);
  }
  float *o = d_verts + image + (size_t)v * 3;
  o[0] = sum.x; o[1] = sum.y; o[2] = sum.z;
}

// ---- launchers -------------------------------------------------------------------------------------------

int vertex_incidence_impl(Context *ctx, const int32_t *tris, int T, int V, int32_t *offsets, int32_t *incidence,
                          cudaStream_t stream) {
  const int n = 3 * T;
  // scratch: counts [V] | cursor [V] | unsorted [3T]
  if (ctx->scratch.reserve(ctx, ((size_t)2 * V + (size_t)n) * sizeof(int))) return PMR_ERR_CUDA;
  int *counts = static_cast<int *>(ctx->scratch.ptr);
  int *cursor = counts + V;
  unsigned *unsorted = reinterpret_cast<unsigned *>(cursor + V);
  PMR_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)V * sizeof(int), stream));
  const int blocks = (n + 255) / 256;
  if (n > 0) incidence_count_kernel<<<blocks, 256, 0, stream>>>(tris, n, V, counts);
  incidence_scan_kernel<<<1, 1024, 0, stream>>>(counts, V, offsets, cursor);
  if (n > 0) {
    incidence_fill_kernel<<<blocks, 256, 0, stream>>>(tris, n, V, cursor, unsorted);
    incidence_rank_kernel<<<blocks, 256, 0, stream>>>(tris, n, V, offsets, unsorted,
                                                      reinterpret_cast<unsigned *>(incidence));
  }
  ctx->launches += n > 0 ? 4 : 1;
  return check_launch(ctx, "vertex incidence kernels");
}

int vertex_normals_forward_impl(Context *ctx, const float *verts, const int32_t *tris, const int32_t *offsets,
                                const int32_t *incidence, int B, int V, float *raw, float *normals,
                                cudaStream_t stream) {
  vertex_normals_forward_kernel<<<dim3((V + 255) / 256, B), 256, 0, stream>>>(
      verts, tris, offsets, reinterpret_cast<const unsigned *>(incidence), V, raw, normals);
  ctx->launches += 1;
  return check_launch(ctx, "vertex_normals_forward_kernel");
}

int vertex_normals_backward_impl(Context *ctx, const float *grad_normals, const float *raw, const float *verts,
                                 const int32_t *tris, const int32_t *offsets, const int32_t *incidence, int B, int V,
                                 float *grad_raw, float *d_verts, cudaStream_t stream) {
  const long long rows = (long long)B * V;
  vertex_normals_unnormalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(grad_normals, raw, rows,
                                                                                        grad_raw);
  vertex_normals_backward_kernel<<<dim3((V + 255) / 256, B), 256, 0, stream>>>(
      grad_raw, verts, tris, offsets, reinterpret_cast<const unsigned *>(incidence), V, d_verts);
  ctx->launches += 2;
  return check_launch(ctx, "vertex_normals_backward kernels");
}

}  // namespace pmr
