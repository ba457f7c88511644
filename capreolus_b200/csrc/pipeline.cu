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

// ---- BERT passage rows ---------------------------------------------------------------------------------------------------
// BertPassage._get_sliding_window_passages + _prepare_bert_input (capreolus/extractor/bertpassage.py:203-232, 268-284) on
// WordPiece ids: passage p = doc[p*stride : p*stride + passagelen] (the first `P` windows; exhausted documents give the one-token
// pad passage), row = [CLS] query [SEP] passage [SEP] [PAD]...; mask = (token != pad) on the written part, 0 on the padding;
// segment = 0 for [CLS] query [SEP], 1 from there to the end INCLUDING the padding.  One warp per (pair, passage) row.
struct BertAssembleArgs {
  const int* q_store;
  const long long* q_off;
  const int* d_store;
  const long long* d_off;
  const int* qidx;
  const int* didx;
  int N, P, L, maxqlen, padq, passagelen, stride, n_queries, n_docs, cls_id, sep_id, pad_id;
  long long* ids;
  long long* mask;
  long long* seg;
};

__global__ void __launch_bounds__(256) assemble_bert_kernel(const BertAssembleArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < (long long)a.N * a.P; row += nwarps) {
    const int n = (int)(row / a.P), p = (int)(row - (long long)n * a.P);
    const int qi = a.qidx[n], di = a.didx[n];
    long long qbeg = 0, qlen = 0, dbeg = 0, dlen = 0;
    if (qi >= 0 && qi < a.n_queries) qbeg = a.q_off[qi], qlen = a.q_off[qi + 1] - qbeg;
    if (di >= 0 && di < a.n_docs) dbeg = a.d_off[di], dlen = a.d_off[di + 1] - dbeg;
    if (qlen > a.maxqlen) qlen = a.maxqlen;                   // "Truncating query" (l.270-272)
    const int qeff = a.padq ? a.maxqlen : (int)qlen;          // padq pads the query to maxqlen with [PAD] (l.274-275)
    const long long start = (long long)p * a.stride;
    const bool real = start < dlen;                           // else: the pad passage [pad_tok] (l.230)
    long long plen = real ? dlen - start : 1;
    if (plen > a.passagelen && real) plen = a.passagelen;
    const int room = a.L - qeff - 3;                          // psg_toks[: maxseqlen - len(query_toks) - 3] (l.276)
    if (plen > room) plen = room > 0 ? room : 0;
    const int sep1 = qeff + 1, pbeg = qeff + 2, sep2 = pbeg + (int)plen;
    for (int t = lane; t < a.L; t += 32) {
      long long tok = a.pad_id;
      bool written = false;
      if (t == 0) tok = a.cls_id, written = true;
      else if (t < sep1) tok = (t - 1 < qlen) ? a.q_store[qbeg + t - 1] : a.pad_id, written = true;
      else if (t == sep1) tok = a.sep_id, written = true;
      else if (t < sep2) tok = real ? a.d_store[dbeg + start + (t - pbeg)] : a.pad_id, written = true;
      else if (t == sep2) tok = a.sep_id, written = true;
      const size_t o = (size_t)row * a.L + t;
      a.ids[o] = tok;
      a.mask[o] = (written && tok != a.pad_id) ? 1 : 0;
      a.seg[o] = t < pbeg ? 0 : 1;
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

// ---- narrow token ids -> the int64 rows the scoring kernels take --------------------------------------------------------
// The reference extractor emits np.long ids (embedtext.py:146-147): 8 bytes per token, 4 352 B per (|q|=32, |d|=512) pair,
// which is what bounds the host -> device leg of the predict loop (448 MB per 100 k pairs).  A 30 k-word vocabulary (and its
// negative OOV ids) fits int16, so the host side may ship ids as int16 / int32 and widen them here: 2 + 8 bytes of HBM
// traffic per token, sign-extending (OOV ids stay negative, 0 stays <pad>).
template <typename T>
__global__ void __launch_bounds__(256) widen_ids_kernel(const T* __restrict__ src, size_t n, long long* __restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (long long)src[i];
}

}  // namespace capr

using namespace capr;

extern "C" {

int capr_assemble_pairs(const int32_t* q_store, const int64_t* q_off, int n_queries, const int32_t* d_store, const int64_t* d_off, int n_docs,
                        const float* idf_store, const int32_t* qidx, const int32_t* didx, int N, int Q, int D, int64_t* query_out,
                        int64_t* doc_out, float* idf_out, capr_stream_t stream) {
  capr::DeviceGuard device_guard(q_store);  // act on the device that owns the caller's buffers
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

int capr_assemble_bert_pairs(const int32_t* q_store, const int64_t* q_off, int n_queries, const int32_t* d_store, const int64_t* d_off,
                             int n_docs, const int32_t* qidx, const int32_t* didx, int N, int P, int L, int maxqlen, int padq, int passagelen,
                             int stride, int cls_id, int sep_id, int pad_id, int64_t* ids, int64_t* mask, int64_t* seg, capr_stream_t stream) {
  capr::DeviceGuard device_guard(q_store);  // act on the device that owns the caller's buffers
  const char* fn = "capr_assemble_bert_pairs";
  CAPR_REQUIRE(N >= 0 && P > 0 && L > 3 && maxqlen >= 0 && passagelen > 0 && stride > 0 && n_queries >= 0 && n_docs >= 0, CAPR_ERR_BAD_SHAPE,
               "%s: bad shape N=%d P=%d L=%d maxqlen=%d passagelen=%d stride=%d", fn, N, P, L, maxqlen, passagelen, stride);
  CAPR_REQUIRE(maxqlen + 3 <= L, CAPR_ERR_BAD_SHAPE, "%s: maxqlen=%d does not fit in maxseqlen=%d", fn, maxqlen, L);
  if (N == 0) return CAPR_OK;
  CAPR_REQUIRE(q_store && q_off && d_store && d_off && qidx && didx && ids && mask && seg, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  BertAssembleArgs a{q_store, (const long long*)q_off, d_store, (const long long*)d_off, qidx, didx, N, P, L, maxqlen, padq, passagelen, stride,
                     n_queries, n_docs, cls_id, sep_id, pad_id, (long long*)ids, (long long*)mask, (long long*)seg};
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const long long want = ((long long)N * P + 7) / 8;
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  assemble_bert_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

int capr_rank_by_query(const float* scores, const int64_t* seg_off, int n_queries, int max_segment, float* rounded, int32_t* order,
                       capr_stream_t stream) {
  capr::DeviceGuard device_guard(scores);  // act on the device that owns the caller's buffers
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

int capr_widen_ids(const void* src, int src_bytes, size_t n, int64_t* dst, capr_stream_t stream) {
  capr::DeviceGuard device_guard(dst);  // act on the device that owns the caller's buffers
  const char* fn = "capr_widen_ids";
  CAPR_REQUIRE(src_bytes == 2 || src_bytes == 4, CAPR_ERR_BAD_SHAPE, "%s: src_bytes=%d must be 2 (int16) or 4 (int32)", fn, src_bytes);
  if (n == 0) return CAPR_OK;
  CAPR_REQUIRE(src && dst, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(((uintptr_t)src & (uintptr_t)(src_bytes - 1)) == 0 && ((uintptr_t)dst & 7) == 0, CAPR_ERR_BAD_POINTER, "%s: misaligned pointer", fn);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const size_t want = (n + 255) / 256;
  const int grid = (int)(want < (size_t)sms * 16 ? want : (size_t)sms * 16);
  if (src_bytes == 2) widen_ids_kernel<short><<<grid, 256, 0, (cudaStream_t)stream>>>((const short*)src, n, (long long*)dst);
  else widen_ids_kernel<int><<<grid, 256, 0, (cudaStream_t)stream>>>((const int*)src, n, (long long*)dst);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

}  // extern "C"
