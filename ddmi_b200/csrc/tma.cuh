// TMA (cp.async.bulk.tensor) helpers: tensor maps built on the host through the driver entry point the runtime hands out
// (no link-time dependency on libcuda), tiled loads issued by one thread and completed on an mbarrier.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace ddmi {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Planes as the VAE decoder emits them: fp32 (batch, C, H, W) -> a 3-D tensor {W, H, batch * C}; box {bw, bh, bc}, no
// swizzle, out-of-bounds elements read as 0.  Needs a 16-byte aligned base and W % 4 == 0 (strides are multiples of 16 B).
static inline bool make_plane_map(CUtensorMap* map, const float* base, int batch, int C, int H, int W, int bw, int bh, int bc) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || ((uintptr_t)base & 15) || (W & 3) || bw > 256 || bh > 256 || bc > 256 || ((bw * 4) & 15)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch * C};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one thread: box at element coordinates (x, y, c) -> shared memory at `dst` (128-byte aligned); completes `bytes of the box`
// transactions on the mbarrier.  x must be a multiple of 4 (the box's first element on a 16-byte boundary): anything else is an
// illegal-instruction fault, not an error code.
__device__ __forceinline__ void load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int c, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(x), "r"(y), "r"(c)
               : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace tma
}  // namespace ddmi
