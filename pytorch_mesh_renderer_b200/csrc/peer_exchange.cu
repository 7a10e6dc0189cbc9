// peer_exchange.cu -- multi-GPU step of the shared-mesh path: the backward of the vertex
// stage and the exchange of its result between the GPUs of one box, as ONE pair of kernels over peer memory
// instead of "kernel, then NCCL all-reduce".
//
// When B_local views per GPU look at one mesh, every rank reduces its views' clip-space gradients to a
// world-space partial [V,3] (transform_backward_kernel) and the partials must be summed over the ranks.  Here
// the kernel that computes the partial stores it straight into a slot of EVERY peer's exchange buffer (plain
// stores over NVLink / NVSwitch; the buffers are mapped with CUDA IPC), the last CTA raises a per-rank flag on
// every peer, and a second kernel on each rank waits for the world's flags and adds the slots in rank order.
// No collective library call, no host synchronisation, one hop over the switch, and -- unlike a ring
// all-reduce -- every rank adds the same numbers in the same order, so all ranks hold bit-identical sums.
//
// Exchange buffer of a rank (pmr_peer_alloc, zero-initialised):
//   line 0                  CTA ticket counter of the push kernel and, 64 bytes in, the step counter (the epoch of the
//                           last push launched here: `epoch` = 0 in the call means "the next one", which keeps
//                           the pair of kernels free of per-step arguments -- capturable in a CUDA graph)
//   line 1                  status word (1 = a wait timed out; sticky) and, 64 bytes in, the local "go" word
//                           (epoch whose partials have all arrived; written by CTA 0 of the reduction)
//   lines 2 ..              flags[parity][peer], one 128-byte line each: last epoch peer `peer` has delivered
//   then                    slots[parity][peer][n_pad] floats
// Two parities alternate by epoch.  A rank can be at most one step ahead of a peer (its own reduction of step
// k+1 needs that peer's flag k+1, raised after the peer's reduction of step k on the peer's stream), so the
// slots of parity k & 1 are never overwritten while a peer still reads step k.
//
// A wait that gives up (a peer died or fell more than the limit behind: PMR_PEER_WAIT_SECONDS, default 10 s) is
// LOUD: the status word is set and stays set, and this and every later reduction on the buffer writes NaN into
// its output instead of a sum of stale slots; pmr_peer_status / SharedGradientExchange.close() report it.
#include "pmr_internal.cuh"

namespace pmr {

constexpr int kLine = 128;                               // bytes per flag line
constexpr int kHeaderLines = 2 + 2 * PMR_MAX_PEERS;
constexpr size_t kHeaderBytes = (size_t)kHeaderLines * kLine;
constexpr int kGoOffset = kLine + 64;                    // byte offset of the "go" word
constexpr int kStepOffset = 64;                          // byte offset of the step counter

struct PeerTable {
  char *base[PMR_MAX_PEERS];
};

__host__ __device__ inline long long padded(long long n) { return (n + 3) & ~3ll; }

__device__ __forceinline__ int *flag_of(char *base, int parity, int peer) {
  return reinterpret_cast<int *>(base + (size_t)(2 + parity * PMR_MAX_PEERS + peer) * kLine);
}
__device__ __forceinline__ float *slot_of(char *base, int parity, int peer, int world, long long n_pad) {
  return reinterpret_cast<float *>(base + kHeaderBytes) + ((size_t)parity * world + peer) * n_pad;
}

__device__ __forceinline__ void store_release_system(int *p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int load_acquire_system(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// d_world partial of this rank (sum over its B views, as transform_backward_kernel with shared = 1), written into
// slot [parity][rank] of every peer; the last CTA to finish raises flag [parity][rank] = epoch on every peer.
__global__ void __launch_bounds__(256)
transform_backward_push_kernel(const float *__restrict__ matrices, const float4 *__restrict__ d_clip, int B, int V,
                               PeerTable peers, int rank, int world, int epoch_arg, long long n_pad) {
  extern __shared__ float ms[];                          // [B][12]: columns 0..2 of every row of M_b
  __shared__ bool last_cta;
  // the step: given, or the successor of the last one pushed from this buffer (written back by the last CTA, so
  // every CTA of this launch reads the same value)
  int *step_counter = reinterpret_cast<int *>(peers.base[rank] + kStepOffset);
  const int epoch = epoch_arg > 0 ? epoch_arg : ((*reinterpret_cast<volatile int *>(step_counter) + 1) & 0x3fffffff);
  const int parity = epoch & 1;
  for (int i = threadIdx.x; i < B * 12; i += blockDim.x) {
    const int bb = i / 12, r = (i % 12) / 3, k = i % 3;
    ms[i] = matrices[(size_t)bb * 16 + r * 4 + k];
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) {
    float gx = 0.0f, gy = 0.0f, gz = 0.0f;
    for (int i = 0; i < B; ++i) {
      const float4 g = __ldg(d_clip + (size_t)i * V + v);
      const float *m = ms + i * 12;
      gx += m[0] * g.x + m[3] * g.y + m[6] * g.z + m[9] * g.w;
      gy += m[1] * g.x + m[4] * g.y + m[7] * g.z + m[10] * g.w;
      gz += m[2] * g.x + m[5] * g.y + m[8] * g.z + m[11] * g.w;
    }
    for (int r = 0; r < world; ++r) {
      float *o = slot_of(peers.base[r], parity, rank, world, n_pad) + (size_t)v * 3;
      o[0] = gx; o[1] = gy; o[2] = gz;
    }
  }
  // Every thread's remote stores are ordered before this CTA's ticket; the CTA that draws the last ticket has
  // therefore (cumulatively) all stores of the grid before its flag stores.
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    int *counter = reinterpret_cast<int *>(peers.base[rank]);
    last_cta = atomicAdd(counter, 1) == (int)gridDim.x - 1;
    if (last_cta) {
      __threadfence_system();                            // acquire side of the ticket: the other CTAs' stores
      *counter = 0;                                      // next launch on this stream starts from zero
      *step_counter = epoch;                             // every CTA has read the old value by now (it drew a ticket)
    }
  }
  __syncthreads();
  if (last_cta && (int)threadIdx.x < world)
    store_release_system(flag_of(peers.base[threadIdx.x], parity, rank), epoch);
}

// out[i] = sum over ranks r = 0 .. world-1 (in that order) of slot[parity][r][i], once every rank's flag shows
// `epoch`.  The slots live in this GPU's memory; they were written by the peers over NVLink, so they are read past
// L1 (ld.global.cg) after the acquiring flag load.  Only CTA 0 polls the peers' flags (system scope); it then
// publishes the epoch in the local "go" word, which the other CTAs of the grid wait for (device scope) -- CTA 0
// is in the first wave of every launch, so they never wait for a CTA that cannot run.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(char *base, int world, int epoch_arg, long long n, long long n_pad,
                       long long wait_cycles, float *__restrict__ out) {
  __shared__ int failed;
  // the step of the push kernel launched just before this one on the same stream
  const int epoch = epoch_arg > 0 ? epoch_arg : *reinterpret_cast<volatile int *>(base + kStepOffset);
  const int parity = epoch & 1;
  int *status = reinterpret_cast<int *>(base + kLine);
  int *go = reinterpret_cast<int *>(base + kGoOffset);
  if (threadIdx.x == 0) failed = 0;
  __syncthreads();
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < world) {
      const int *flag = flag_of(base, parity, threadIdx.x);
      const long long t0 = clock64();
      while (load_acquire_system(flag) < epoch) {
        if (clock64() - t0 > wait_cycles) {
          atomicExch(status, 1);
          break;
        }
        __nanosleep(200);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();                                   // the peers' partials (acquired above) before the go word
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(go), "r"(epoch) : "memory");
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int seen;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(go) : "memory");
      if (seen == epoch) break;
      if (clock64() - t0 > 2 * wait_cycles) {
        atomicExch(status, 1);
        break;
      }
      __nanosleep(100);
    } while (true);
  }
  if (threadIdx.x == 0) failed = *reinterpret_cast<volatile int *>(status);
  __syncthreads();
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  float4 acc;
  if (failed) {
    acc.x = acc.y = acc.z = acc.w = __int_as_float(0x7fc00000);      // never a plausible gradient
  } else {
    const float4 *s0 = reinterpret_cast<const float4 *>(slot_of(base, parity, 0, world, n_pad));
    acc = __ldcg(s0 + i4);
    for (int r = 1; r < world; ++r) {
      const float4 x = __ldcg(reinterpret_cast<const float4 *>(slot_of(base, parity, r, world, n_pad)) + i4);
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
  }
  if (i4 * 4 + 3 < n) {
    float *o = out + i4 * 4;                             // out is [V,3] floats: 4-byte alignment only
    o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
  } else {
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
    for (int k = 0; i4 * 4 + k < n; ++k) out[i4 * 4 + k] = a[k];
  }
}

size_t peer_exchange_bytes(long long n_floats, int world) {
  return kHeaderBytes + (size_t)2 * world * padded(n_floats) * sizeof(float);
}

int transform_backward_exchange_impl(Context *ctx, const float *matrices, const float *d_clip, int B, int V,
                                     void *const *peers, int rank, int world, long long epoch, float *d_world,
                                     cudaStream_t stream) {
  PeerTable table;
  for (int r = 0; r < PMR_MAX_PEERS; ++r) table.base[r] = r < world ? static_cast<char *>(peers[r]) : nullptr;
  const long long n = (long long)V * 3, n_pad = padded(n);
  const int stamp = (int)(epoch & 0x3fffffff);             // 0: the kernels take the step from the buffer
  const size_t smem = (size_t)B * 12 * sizeof(float);
  if (smem > 48 * 1024) return set_error(ctx, PMR_ERR_SIZE, "too many views for one transform_backward launch");
  transform_backward_push_kernel<<<(V + 255) / 256, 256, smem, stream>>>(
      matrices, reinterpret_cast<const float4 *>(d_clip), B, V, table, rank, world, stamp, n_pad);
  reduce_partials_kernel<<<(unsigned)((n_pad / 4 + 255) / 256), 256, 0, stream>>>(
      table.base[rank], world, stamp, n, n_pad, ctx->peer_wait_cycles, d_world);
  ctx->launches += 2;
  return check_launch(ctx, "transform_backward_push_kernel / reduce_partials_kernel");
}

}  // namespace pmr
