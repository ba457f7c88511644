// K9 / K10 -- the callers either side of the scoring kernels.
//
//   EmbedText.id2vec + padlist        capreolus/extractor/embedtext.py:128-162, utils/common.py:99-111
//   PredSampler.generate_samples      capreolus/sampler/__init__.py:222-233       (qid, docid) -> padded id rows
//   PytorchTrainer.predict            capreolus/trainer/pytorch.py:338-348        fp16 rounding of the scores
//   Searcher.write_trec_run           capreolus/searcher/__init__.py:48-58        per query: sort by score, descending
//
// The reference builds every (query, doc) feature row in Python (list slicing, dict lookups: micro- to milliseconds per
// pair) and sorts the run in Python.  At millions of pairs per second both ends become the bottleneck, so:
//   capr_assemble_pairs  the tokenised queries / documents live ONCE in HBM as a packed id store (flat int32 ids + int64
//                        offsets); a batch is described by two int32 index vectors and the padded int64 [N,Q] / [N,D] rows
//                        the rerankers consume are gathered on the device (truncate to Q / D, pad with 0 = padlist).
//                        Pure HBM streaming: reads 4 B, writes 8 B per token.
//   capr_rank_by_query   scores -> float16 rounding (pytorch.py:347) -> per query a stable descending order (bitonic sort
//                        of (score desc, position asc) keys in shared memory; one CTA per query).
#include "common.cuh"
#include <cuda_fp16.h>

namespace capr {

struct AssembleArgs {
  const int* q_store;
  const long long* q_off;
  const int* d_store;
  const long long* d_off;
  const float* idf_store;  // nullable, parallel to q_store
  const int* qidx;
  const int* didx;
  int N, Q, D, n_queries, n_docs;
  long long* query_out;
  long long* doc_out;
  float* idf_out;  // nullable
};

// One warp per output row chunk: rows are written with coalesced 8-byte stores; the source segment is contiguous.
__global__ void __launch_bounds__(256) assemble_pairs_kernel(const AssembleArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pair = warp0; pair < a.N; pair += nwarps) {
    const int qi = a.qidx[pair], di = a.didx[pair];
    {
      long long beg = 0, len = 0;
      if (qi >= 0 && qi < a.n_queries) beg = a.q_off[qi], len = a.q_off[qi + 1] - beg;
      long long* out = a.query_out + pair * a.Q;
      for (int i = lane; i < a.Q; i += 32) {
        out[i] = i < len ? (long long)a.q_store[beg + i] : 0;
        if (a.idf_out) a.idf_out[pair * a.Q + i] = (i < len && a.idf_store) ? a.idf_store[beg + i] : 0.f;
      }
    }
    {
      long long beg = 0, len = 0;
      if (di >= 0 && di < a.n_docs) beg = a.d_off[di], len = a.d_off[di + 1] - beg;
      long long* out = a.doc_out + pair * a.D;
      for (int i = lane; i < a.D; i += 32) out[i] = i < len ? (long long)a.d_store[beg + i] : 0;
    }
  }
}

// ---- per-query ranking -------------------------------------------------------------------------------------------
// key = (monotone image of the fp16-rounded score) << 32 | (0xffffffff - position): sorting keys DESCENDING gives
// score descending, ties by ascending position = Python's stable sorted(..., reverse=True) over insertion order.
__device__ __forceinline__ unsigned int float_order_bits(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int CAP>
__global__ void __launch_bounds__(CAP / 2 > 1024 ? 1024 : (CAP / 2 < 32 ? 32 : CAP / 2))
rank_by_query_kernel(const float* __restrict__ scores, const long long* __restrict__ seg_off, int nq, float* __restrict__ rounded,
                     int* __restrict__ order) {
  __shared__ unsigned long long keys[CAP];
  for (int q = blockIdx.x; q < nq; q += gridDim.x) {
    const long long beg = seg_off[q];
    const int n = (int)(seg_off[q + 1] - beg);
    for (int i = threadIdx.x; i < CAP; i += blockDim.x) {
      unsigned long long k = 0;  // padding sorts last
      if (i < n) {
        const float r = __half2float(__float2half_rn(scores[beg + i]));  // score.astype(np.float16)
        if (rounded) rounded[beg + i] = r;
        k = ((unsigned long long)float_order_bits(r) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
      }
      keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= CAP; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < CAP / 2; t += blockDim.x) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool desc = (lo & size) == 0;  // first half of every size-block descending -> whole array descending
          const unsigned long long x = keys[lo], y = keys[hi];
          if ((x < y) == desc) keys[lo] = y, keys[hi] = x;
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[beg + i] = (int)(0xffffffffu - (unsigned)(keys[i] & 0xffffffffu));
    __syncthreads();
  }
}

}  // namespace capr

using namespace capr;

extern "C" {

int capr_assemble_pairs(const int32_t* q_store, const int64_t* q_off, int n_queries, const int32_t* d_store, const int64_t* d_off, int n_docs,
                        const float* idf_store, const int32_t* qidx, const int32_t* didx, int N, int Q, int D, int64_t* query_out,
                        int64_t* doc_out, float* idf_out, capr_stream_t stream) {
  const char* fn = "capr_assemble_pairs";
  CAPR_REQUIRE(N >= 0 && Q > 0 && D > 0 && n_queries >= 0 && n_docs >= 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape N=%d Q=%d D=%d", fn, N, Q, D);
  if (N == 0) return CAPR_OK;
  CAPR_REQUIRE(q_store && q_off && d_store && d_off && qidx && didx && query_out && doc_out, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  AssembleArgs a{q_store, (const long long*)q_off, d_store, (const long long*)d_off, idf_store, qidx, didx, N, Q, D, n_queries, n_docs,
                 (long long*)query_out, (long long*)doc_out, idf_out};
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const long long want = ((long long)N + 7) / 8;
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  assemble_pairs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

int capr_rank_by_query(const float* scores, const int64_t* seg_off, int n_queries, int max_segment, float* rounded, int32_t* order,
                       capr_stream_t stream) {
  const char* fn = "capr_rank_by_query";
  CAPR_REQUIRE(n_queries >= 0 && max_segment >= 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape n_queries=%d max_segment=%d", fn, n_queries, max_segment);
  if (n_queries == 0 || max_segment == 0) return CAPR_OK;
  CAPR_REQUIRE(scores && seg_off && order, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(max_segment <= 4096, CAPR_ERR_UNSUPPORTED, "%s: %d candidates for one query > 4096 is not supported", fn, max_segment);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const int grid = n_queries < sms * 4 ? n_queries : sms * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const long long* so = (const long long*)seg_off;
  if (max_segment <= 128) rank_by_query_kernel<128><<<grid, 64, 0, st>>>(scores, so, n_queries, rounded, order);
  else if (max_segment <= 1024) rank_by_query_kernel<1024><<<grid, 512, 0, st>>>(scores, so, n_queries, rounded, order);
  else rank_by_query_kernel<4096><<<grid, 1024, 0, st>>>(scores, so, n_queries, rounded, order);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

}  // extern "C"
