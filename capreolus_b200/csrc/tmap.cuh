// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point, so the
// library does not link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace capr {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor [rows, cols] (cols contiguous) -> map with a {64 elements = 128 bytes, box_rows} box, SWIZZLE_128B.
// box_rows = 1 is the form cp.async.bulk.tensor ... tile::gather4 expects (four such rows per instruction).
inline int make_bf16_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  CAPR_REQUIRE(fn, CAPR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CAPR_REQUIRE(r == CUDA_SUCCESS, CAPR_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows=%llu cols=%llu box_rows=%u)", (int)r,
               (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return CAPR_OK;
}

inline int make_gather_map(CUtensorMap* m, const void* table_plane, int V, int pitch) { return make_bf16_map(m, table_plane, (uint64_t)V, (uint64_t)pitch, 1); }

}  // namespace tc
}  // namespace capr
