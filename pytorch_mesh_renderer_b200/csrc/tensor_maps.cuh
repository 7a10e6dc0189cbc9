// Tensor maps (TMA) over the dense per-pixel arrays of the path, and the device-side copies that use them.
//
// Every per-pixel array of the path -- ids [B,H,W], z [B,H,W], barycentrics [B,H,W,3], attribute image / its
// gradient [B,H,W,A] -- is a 2-D array [B*H rows][W*channels] of 4-byte elements.  A warp works on a block of
// 8 x 4 pixels: a box of 8*channels x 4 elements, which one cp.async.bulk.tensor.2d moves between global and
// shared memory (SASS UTMALDG / UTMASTG) from ONE lane, where per-lane loads / stores of 12- or 36-byte pixels
// cost an instruction per 4 bytes and lane.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmr {

__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Box at element (c0, row c1) of the mapped array to shared memory at `dst` (128-byte aligned); completion is counted
// in bytes on the mbarrier at `bar` (the caller has announced them with arrive.expect_tx); the part of a box outside
// the array arrives as zeros.  REFILLING a box: the copy engine writes through the async proxy and is ordered with
// this warp's shared-memory LOADS of the previous contents by nothing -- not by program order, not by __syncwarp
// (which ptxas drops in converged code).  Issue the refill only after an instruction has consumed what was loaded
// (then the loads have returned), as backward_blocks_kernel does after its row stores.
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// cuTensorMapEncodeTiled through the runtime (the library does not link libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_tiled() {
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
static inline bool make_block_map(CUtensorMap *map, CUtensorMapDataType type, const void *base, long long inner,
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


// Box of the shared-memory array at `src` (dense [4][box_inner], 128-byte aligned) to the array at element
// (c0, row c1); elements outside the array are not written.  The issuing lane commits and waits for the reads
// (cp.async.bulk.commit_group + wait_group.read 0) before the shared memory is rewritten or the CTA exits, and
// the writers of the shared memory execute fence.proxy.async.shared::cta before the store is issued.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, unsigned src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}

}  // namespace pmr
