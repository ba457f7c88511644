// K1 (engine 2) -- fused KNRM scoring with the cosine tile on tcgen05 tensor cores.
//
//   KNRM_class.forward            capreolus/reranker/KNRM.py:39-55
//   RbfKernelBank                 capreolus/reranker/common.py:224-250
//   SimilarityMatrix              capreolus/reranker/common.py:143-182  (producer: simtc.cuh)
//
// Same math and same quirks as knrm.cu (sum over ALL doc positions, live-row test on the cosine row-sum, log(S+1e-6),
// combine); only the producer of the cosine tile differs: gather -> UMMA -> TMEM instead of gather -> FFMA.  The
// pooling epilogue (MUFU ex2 bound, 180 k exponentials per pair) runs on 8 warps while the producer and MMA warps are
// already working on the next pair.
#include "simtc.cuh"

namespace capr {

struct KnrmTcArgs {
  simtc::Problem pr;
  int K, hidden, flags;
  const float* mu;
  const float* sigma;
  const float *w1, *b1, *w2, *b2;
  float* scores;
  float* feats;
};

template <int KT>
__global__ void __launch_bounds__(simtc::THREADS, 1) knrm_tc_kernel(const KnrmTcArgs a) {
  using namespace simtc;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem s = carve(smem_raw, (a.pr.pitch + ATOM_K - 1) / ATOM_K);
  float* sPart = s.extra;  // [EPI_WARPS][KT] per-warp partial features
  const uint32_t tmem_base = setup(s, tid);

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    producer_loop(s, a.pr, tid - EPI_THREADS);
  } else if (warp == EPI_WARPS + PROD_WARPS) {
    mma_loop(s, a.pr, tmem_base);
  } else {
    // ===================== epilogue: 8 warps, 4 query rows each =====================
    float mu[KT], cc[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float sg = k < a.K ? a.sigma[k] : 1.f;
      mu[k] = k < a.K ? a.mu[k] : 0.f;
      cc[k] = -0.5f * 1.4426950408889634f / (sg * sg);
    }
    constexpr int ROWS_PER_WARP = QT / EPI_WARPS;  // 4
    uint32_t acc_phase[2] = {0, 0};
    int unit = 0;
    for (int pair = blockIdx.x; pair < a.pr.B; pair += gridDim.x, unit += halves_of(a.pr)) {
      drain_pair(s, a.pr, tmem_base, pair, unit, acc_phase, tid, (a.flags & CAPR_DEBUG_SKIP_DRAIN) != 0);
      // lane k accumulates this warp's share of R_k = sum over its live rows of log(S_k + 1e-6)   (KNRM.py:50-53)
      float R_part = 0.f;
      if (!(a.flags & CAPR_DEBUG_SKIP_POOL)) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
          const float* row = s.sim + (warp * ROWS_PER_WARP + r) * SIM_PITCH;
          float S[KT], rs = 0.f;
#pragma unroll
          for (int k = 0; k < KT; ++k) S[k] = 0.f;
          for (int c = lane; c < a.pr.D; c += 32) {
            const float v = row[c];
            rs += v;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              const float adj = v - mu[k];
              S[k] += ex2_approx(cc[k] * adj * adj);
            }
          }
          rs = warp_sum(rs);
          float mine = 0.f;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const float t = warp_sum(S[k]);
            mine = lane == k ? t : mine;
          }
          if (rs != 0.0f && warp * ROWS_PER_WARP + r < a.pr.Q) R_part += logf(mine + 1e-6f);  // KNRM.py:51-52
        }
      }
      if (lane < KT) sPart[warp * KT + lane] = R_part;
      epi_barrier();
      if (warp == 0) {
        float R = 0.f;
        if (lane < a.K) {
#pragma unroll
          for (int w = 0; w < EPI_WARPS; ++w) R += sPart[w * KT + lane];  // fixed order over the 8 row groups
          if (a.feats) a.feats[(size_t)pair * a.K + lane] = R;
        }
        if (a.scores) {
          float out;
          if (a.hidden == 0) {
            out = warp_sum(lane < a.K ? a.w1[lane] * R : 0.f) + a.b1[0];
          } else {
            float p = 0.f;
            for (int h = 0; h < a.hidden; ++h) {
              const float acc = warp_sum(lane < a.K ? a.w1[h * a.K + lane] * R : 0.f) + a.b1[h];
              p = fmaf(a.w2[h], tanhf(acc), p);
            }
            out = p + a.b2[0];
          }
          if (a.flags & CAPR_KNRM_SCORETANH) out = tanhf(out);
          if (lane == 0) a.scores[pair] = out;
        }
      }
      // the next drain_pair starts with an epi_barrier, which orders these reads before the next writes of s.sim / sS
    }
  }
  teardown(s, tmem_base, tid);
}

// hi/lo bf16 planes of the L2-normalised table (see table.cu for the fp32 variant)
__global__ void __launch_bounds__(256) table_prepare_bf16_kernel(const float* __restrict__ emb, int V, int E, __nv_bfloat16* __restrict__ hi,
                                                                 __nv_bfloat16* __restrict__ lo, int pitch) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= V) return;
  const float* row = emb + (size_t)warp * E;
  float ss = 0.f;
  for (int e = lane; e < E; e += 32) {
    const float x = row[e];
    ss = fmaf(x, x, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / (sqrtf(ss) + 1e-9f);
  for (int e = lane; e < pitch; e += 32) {
    const float y = e < E ? row[e] * inv : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(y);
    hi[(size_t)warp * pitch + e] = h;
    lo[(size_t)warp * pitch + e] = __float2bfloat16_rn(y - __bfloat162float(h));
  }
}

}  // namespace capr

using namespace capr;

extern "C" {

int capr_table_pitch_bf16(int E) { return E <= 0 ? 0 : ((E + 15) / 16) * 16; }

int capr_table_prepare_bf16(const float* emb, int V, int E, void* hi, void* lo, int pitch, capr_stream_t stream) {
  CAPR_REQUIRE(V > 0 && E > 0, CAPR_ERR_BAD_SHAPE, "capr_table_prepare_bf16: V=%d E=%d must be positive", V, E);
  CAPR_REQUIRE(pitch >= E && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "capr_table_prepare_bf16: pitch=%d must be a multiple of 16 and >= E=%d", pitch, E);
  CAPR_REQUIRE(emb && hi && lo, CAPR_ERR_BAD_POINTER, "capr_table_prepare_bf16: null pointer");
  CAPR_REQUIRE((((uintptr_t)hi | (uintptr_t)lo) & 15) == 0, CAPR_ERR_BAD_POINTER, "capr_table_prepare_bf16: planes must be 16-byte aligned");
  table_prepare_bf16_kernel<<<(V + 7) / 8, 256, 0, (cudaStream_t)stream>>>(emb, V, E, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, pitch);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

int capr_knrm_forward_tc(const int64_t* query, const int64_t* doc, int B, int Q, int D, const void* table_hi, const void* table_lo, int V,
                         int E, int pitch, const float* mu, const float* sigma, int K, const float* w1, const float* b1, int hidden,
                         const float* w2, const float* b2, int flags, float* scores, float* feats, capr_stream_t stream) {
  const char* fn = "capr_knrm_forward_tc";
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0 && E > 0 && K > 0 && hidden >= 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d E=%d K=%d", fn, B, Q, D, V, E, K);
  CAPR_REQUIRE(pitch >= E && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: pitch=%d must be a multiple of 16 and >= E (capr_table_pitch_bf16)", fn, pitch);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(D <= DT, CAPR_ERR_UNSUPPORTED, "%s: maxdoclen=%d > %d: use capr_knrm_forward (doc-tiled FFMA engine)", fn, D, DT);
  CAPR_REQUIRE(pitch <= simtc::MAX_ATOMS * simtc::ATOM_K, CAPR_ERR_UNSUPPORTED, "%s: embedding dim > %d: use capr_knrm_forward", fn, simtc::MAX_ATOMS * simtc::ATOM_K);
  CAPR_REQUIRE(K <= 16, CAPR_ERR_UNSUPPORTED, "%s: K=%d > 16 kernels: use capr_knrm_forward", fn, K);
  CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table of %d x %d elements is too large for 32-bit row offsets", fn, V, pitch);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(query && doc && table_hi && table_lo && mu && sigma, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE((((uintptr_t)table_hi | (uintptr_t)table_lo) & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: table planes must be 16-byte aligned", fn);
  CAPR_REQUIRE(scores || feats, CAPR_ERR_BAD_POINTER, "%s: no output requested", fn);
  if (scores) {
    CAPR_REQUIRE(w1 && b1, CAPR_ERR_BAD_POINTER, "%s: scores requested without combine weights", fn);
    CAPR_REQUIRE(hidden == 0 || (w2 && b2), CAPR_ERR_BAD_POINTER, "%s: hidden=%d needs w2/b2", fn, hidden);
  }
  KnrmTcArgs a{};
  a.pr = simtc::Problem{(const long long*)query, (const long long*)doc, B, Q, D, V, (const __nv_bfloat16*)table_hi, (const __nv_bfloat16*)table_lo, pitch, E, flags & 0xF00};
  a.K = K, a.hidden = hidden, a.flags = flags, a.mu = mu, a.sigma = sigma, a.w1 = w1, a.b1 = b1, a.w2 = w2, a.b2 = b2, a.scores = scores, a.feats = feats;
  const int KT = K <= 11 ? 11 : 16;
  const size_t smem = simtc::smem_bytes((pitch + simtc::ATOM_K - 1) / simtc::ATOM_K, (size_t)(simtc::EPI_WARPS * KT) * sizeof(float));
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const int grid = B < sms ? B : sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (KT == 11) {
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(knrm_tc_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knrm_tc_kernel<11><<<grid, simtc::THREADS, smem, st>>>(a);
  } else {
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(knrm_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knrm_tc_kernel<16><<<grid, simtc::THREADS, smem, st>>>(a);
  }
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

}  // extern "C"
