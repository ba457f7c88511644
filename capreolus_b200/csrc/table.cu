// Embedding-table preparation.
//
// Replaces, once per set of embedding weights, the per-batch norm()/divide of
// SimilarityMatrix.cosine_similarity_matrix (capreolus/reranker/common.py:161-165):
//     sim = a.b / ((|a| + 1e-9) (|b| + 1e-9))    ==    (a / (|a| + 1e-9)) . (b / (|b| + 1e-9))
// so the scoring kernels gather rows that are already scaled and only take dot products.  The row pitch
// is padded to a multiple of 16 floats (zero filled) so every K-chunk the kernels stage is a whole number
// of aligned 16-byte segments.  One warp per row; HBM-bound streaming pass (read V*E*4, write V*pitch*4).
#include "common.cuh"

namespace capr {

__global__ void __launch_bounds__(256) table_prepare_kernel(const float* __restrict__ emb, int V, int E,
                                                            float* __restrict__ out, int pitch) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= V) return;
  const float* row = emb + (size_t)warp * E;
  float ss = 0.f;
  for (int e = lane; e < E; e += 32) {
    float x = row[e];
    ss = fmaf(x, x, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / (sqrtf(ss) + 1e-9f);
  float* o = out + (size_t)warp * pitch;
  for (int e = lane; e < pitch; e += 32) o[e] = e < E ? row[e] * inv : 0.f;
}

}  // namespace capr

extern "C" {

int capr_table_pitch(int E) { return E <= 0 ? 0 : ((E + 15) / 16) * 16; }

int capr_table_prepare(const float* emb, int V, int E, float* table, int pitch, capr_stream_t stream) {
  capr::DeviceGuard device_guard(emb);  // act on the device that owns the caller's buffers
  CAPR_REQUIRE(V > 0 && E > 0, CAPR_ERR_BAD_SHAPE, "capr_table_prepare: V=%d E=%d must be positive", V, E);
  CAPR_REQUIRE(pitch >= E && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "capr_table_prepare: pitch=%d must be a multiple of 16 and >= E=%d", pitch, E);
  CAPR_REQUIRE(emb && table, CAPR_ERR_BAD_POINTER, "capr_table_prepare: null pointer");
  CAPR_REQUIRE(((uintptr_t)table & 15) == 0, CAPR_ERR_BAD_POINTER, "capr_table_prepare: table must be 16-byte aligned");
  const int warps_per_block = 8;
  const int blocks = (V + warps_per_block - 1) / warps_per_block;
  capr::table_prepare_kernel<<<blocks, warps_per_block * 32, 0, (cudaStream_t)stream>>>(emb, V, E, table, pitch);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

}  // extern "C"
