// K7 -- fused DRMMTKS scoring kernel, tensor-core cosine tile + top-k pooling.
//
//   DRMMTKS_class.forward      capreolus/reranker/DRMMTKS.py:50-63   cos_mat -> topk(k) per query term -> ffw -> gate -> score
//   DRMMTKS_class._term_gate   capreolus/reranker/DRMMTKS.py:32-48   softmax term gate (IDF; the TV branch cannot run in the
//                                                                    reference: it feeds int64 token ids to a Linear(E,1))
//   SimilarityMatrix           capreolus/reranker/common.py:143-182  (producer: simtc.cuh)
//
// Same pipeline as knrm_tc_kernel (gather -> tcgen05 -> TMEM -> drain warps -> pooling warps); only the pooling differs:
// each pooling thread owns one query row and one 32-column slice of every half tile and keeps the k largest cosines it has
// seen in a sorted register list; at the end of the pair the 8 slices park their lists in the last half tile and each warp
// merges the 8 x k candidates of 4 rows by k rounds of warp-wide max selection (duplicates -- several exact matches, or
// the zeros of padded columns, which compete like in torch.topk over the zero-padded matrix -- are removed one instance
// per round).  Then Linear(k,1) + tanh per row, softmax gate over the query terms (pads at -1e7), Linear(1,1).
#include "simtc.cuh"

namespace capr {

struct TksArgs {
  simtc::Problem pr;
  int topk;
  const float* idf;
  const float *ffw_w, *ffw_b, *gate_w, *out_w, *out_b;
  float* scores;
  float* topk_out;
};

template <int T>
__global__ void __launch_bounds__(simtc::THREADS_PIPE, 1) drmmtks_tc_kernel(const TksArgs a) {
  using namespace simtc;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem s = carve(smem_raw, (a.pr.pitch + ATOM_K - 1) / ATOM_K, a.pr.deep != 0);
  float* z = spare_scratch(s);  // [QT] ffw output per query term
  const uint32_t tmem_base = setup(s, tid, THREADS_PIPE, MMA_WARP_PIPE);
  if (is_producer_warp(warp)) {
    producer_loop(s, a.pr, producer_index(warp) * 32 + lane);
  } else if (warp == MMA_WARP_PIPE) {
    mma_loop(s, a.pr, tmem_base);
  } else if (is_drain_warp(warp)) {
    drain_loop(s, a.pr, tmem_base, warp, lane);
  } else if (is_pool_warp(warp)) {
    const int pw = pool_index(warp);
    constexpr int SLICE = NT_DOCS / POOL_WARPS;  // 32 columns per warp and half tile
    static_assert(T <= SLICE, "a warp parks its k best values of a row in its own 32-column slice");
    constexpr int CAND = POOL_WARPS * T, PER = (CAND + 31) / 32;
    const int halves = halves_of(a.pr);
    const int K = a.topk;
    PoolSync ps;
    int unit = 0, it = 0;
    for (int pair = blockIdx.x; pair < a.pr.B; pair += gridDim.x, ++it) {
      float t[T];  // the K largest values of (row = lane, this warp's columns), descending
#pragma unroll
      for (int j = 0; j < T; ++j) t[j] = -INFINITY;
      int ub = 0;
      for (int h = 0; h < halves; ++h, ++unit) {
        ub = unit & 1;
        ps.wait_full(s, ub);
        const int nvalid = min(NT_DOCS, a.pr.D - h * NT_DOCS) - pw * SLICE;  // columns of this slice that exist
        const float4* row = reinterpret_cast<const float4*>(half_tile(s, ub) + lane * HALF_PITCH + pw * SLICE);
        for (int g = 0; g < SLICE / 4; ++g) {
          if (g * 4 >= nvalid) break;  // warp-uniform
          const float4 x = row[g];
          const float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // sorted insert without a serial bubble: every slot is rebuilt from the OLD list, t'[i] = max(min(t[i-1], v), t[i])
            // (v <= t[i]: unchanged; t[i] < v <= t[i-1]: v lands here; v > t[i-1]: the old t[i-1] shifts down).  Branch-free
            // and two operations deep, so consecutive cosines overlap; columns past the end enter as -inf (no effect).
            const float nv = (g * 4 + j < nvalid) ? v[j] : -INFINITY;
            float prev = t[0];
            t[0] = fmaxf(prev, nv);
#pragma unroll
            for (int i = 1; i < T; ++i) {
              const float cur = t[i];
              t[i] = fmaxf(fminf(prev, nv), cur);
              prev = cur;
            }
          }
        }
        if (h + 1 < halves) ps.release(s, ub, lane);  // the last half tile is kept for the merge
      }
      float* tile = half_tile(s, ub);
      {
        float* mine = tile + lane * HALF_PITCH + pw * SLICE;
#pragma unroll
        for (int j = 0; j < T; ++j) mine[j] = t[j];
      }
      epi_barrier();  // the 8 pooling warps
      // merge: warp pw owns rows 4pw..4pw+3; lane i holds candidates i, i+32, ... of the row's 8 x T list
      float c[4][PER];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int qrow = pw * 4 + r;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          const int idx = lane + 32 * e;
          c[r][e] = idx < CAND ? tile[qrow * HALF_PITCH + (idx / T) * SLICE + (idx % T)] : -INFINITY;
        }
      }
      ps.release(s, ub, lane);
      const int pp = it & 1;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int qrow = pw * 4 + r;
        float mine = 0.f;  // lane j ends up with the j-th largest value of the row
        for (int round = 0; round < K; ++round) {
          float m = c[r][0];
#pragma unroll
          for (int e = 1; e < PER; ++e) m = fmaxf(m, c[r][e]);
          const float M = warp_max(m);
          const unsigned who = __ballot_sync(0xffffffffu, m == M);
          if (lane == round) mine = M;
          if (lane == __ffs(who) - 1) {  // remove ONE instance of the maximum
            bool done = false;
#pragma unroll
            for (int e = 0; e < PER; ++e) {
              const bool hit = !done && c[r][e] == M;
              c[r][e] = hit ? -INFINITY : c[r][e];
              done = done || hit;
            }
          }
        }
        if (qrow < a.pr.Q) {
          if (a.topk_out && lane < K) a.topk_out[((size_t)pair * a.pr.Q + qrow) * K + lane] = mine;
          const float acc = warp_sum(lane < K ? a.ffw_w[lane] * mine : 0.f) + a.ffw_b[0];  // ffw: Linear(k,1) + tanh
          if (lane == 0) z[qrow] = tanhf(acc);
        }
      }
      epi_barrier();
      if (pw == 0) {
        // term gate (DRMMTKS.py:38-47): softmax over the Q query positions of w_g*idf, pads at -1e7; then output_layer
        float logit = -INFINITY;
        if (lane < a.pr.Q) {
          const float pad_bias = (s.qid[pp * QT + lane] == 0) ? -1e7f : 0.f;
          logit = a.gate_w[0] * a.idf[(size_t)pair * a.pr.Q + lane] + pad_bias;
        }
        const float m = warp_max(logit);
        const float ex = lane < a.pr.Q ? expf(logit - m) : 0.f;
        const float den = warp_sum(ex);
        const float x = warp_sum(lane < a.pr.Q ? (ex / den) * z[lane] : 0.f);
        if (lane == 0) a.scores[pair] = fmaf(a.out_w[0], x, a.out_b[0]);
      }
      epi_barrier();  // z is rewritten by the next pair
    }
  }
  teardown(s, tmem_base, tid, MMA_WARP_PIPE);
}

}  // namespace capr

using namespace capr;

extern "C" int capr_drmmtks_forward_tc(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D, const void* table_hi,
                                       const void* table_lo, int V, int E, int pitch, int topk, const float* ffw_w, const float* ffw_b,
                                       const float* gate_w, const float* out_w, const float* out_b, float* scores, float* topk_out,
                                       capr_stream_t stream) {
  capr::DeviceGuard device_guard(table_hi);  // act on the device that owns the caller's buffers
  const char* fn = "capr_drmmtks_forward_tc";
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0 && E > 0 && topk > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d E=%d topk=%d", fn, B, Q, D, V, E, topk);
  CAPR_REQUIRE(topk <= D, CAPR_ERR_BAD_SHAPE, "%s: topk=%d is out of range for maxdoclen=%d (torch.topk raises in the reference)", fn, topk, D);
  CAPR_REQUIRE(pitch >= E && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: pitch=%d must be a multiple of 16 and >= E (capr_table_pitch_bf16)", fn, pitch);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(D <= simtc::DEEP_DCAP && pitch <= simtc::MAX_ATOMS * simtc::ATOM_K, CAPR_ERR_UNSUPPORTED, "%s: needs maxdoclen <= %d and emb dim <= %d", fn, simtc::DEEP_DCAP, simtc::MAX_ATOMS * simtc::ATOM_K);
  CAPR_REQUIRE(topk <= 32, CAPR_ERR_UNSUPPORTED, "%s: topk=%d > 32 is not supported", fn, topk);
  CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table of %d x %d elements is too large for 32-bit row offsets", fn, V, pitch);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(query && doc && idf && table_hi && table_lo && ffw_w && ffw_b && gate_w && out_w && out_b && scores, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE((((uintptr_t)table_hi | (uintptr_t)table_lo) & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: table planes must be 16-byte aligned", fn);
  TksArgs a{simtc::Problem{(const long long*)query, (const long long*)doc, B, Q, D, V, (const __nv_bfloat16*)table_hi, (const __nv_bfloat16*)table_lo, pitch, E, 0},
            topk, idf, ffw_w, ffw_b, gate_w, out_w, out_b, scores, topk_out};
  const int atoms = (pitch + simtc::ATOM_K - 1) / simtc::ATOM_K;
  const char* ring_env = getenv("CAPR_SIM_RING");  // see capr_knrm_forward_tc
  a.pr.deep = (D > DT || (atoms >= 3 && !(ring_env && ring_env[0] == '2'))) ? 1 : 0;  // maxdoclen > 512 needs the deep layout's id arrays
  const size_t smem = simtc::smem_bytes(atoms, 0, a.pr.deep != 0);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const int grid = B < sms ? B : sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (topk <= 10) {
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(drmmtks_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    drmmtks_tc_kernel<10><<<grid, simtc::THREADS_PIPE, smem, st>>>(a);
  } else {
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(drmmtks_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    drmmtks_tc_kernel<32><<<grid, simtc::THREADS_PIPE, smem, st>>>(a);
  }
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}
