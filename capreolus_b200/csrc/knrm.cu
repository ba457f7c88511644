// K1 / K4 -- fused KNRM scoring kernel (and its training-statistics variant).
//
//   KNRM_class.forward            capreolus/reranker/KNRM.py:39-55
//   RbfKernel / RbfKernelBank     capreolus/reranker/common.py:224-250
//   SimilarityMatrix              capreolus/reranker/common.py:143-182  (producer: simtile.cuh)
//
// One persistent CTA per (query, doc) pair: gather -> cosine tile (smem) -> K Gaussian kernels ->
// sum over ALL doc positions (pads count with s=0, KNRM.py:50) -> a query row is live iff its cosine row
// sums to something != 0 (KNRM.py:51) -> log(S + 1e-6) summed over live rows (KNRM.py:52-53) -> combine.
// Nothing of the reference's [B,K,Q,D] temporaries ever leaves the SM; HBM sees ids in, one float out.
//
// exp(-0.5 (s-mu)^2 / sigma^2) is evaluated as ex2(c (s-mu)^2) with c = -0.5 log2(e) / sigma^2 (one MUFU op).
// Reductions over the doc axis are warp-shuffle butterflies in a fixed order, so scores are bit-reproducible
// and independent of how pairs are sharded over CTAs or GPUs.
#include "simtile.cuh"

namespace capr {

struct KnrmArgs {
  const long long* q;
  const long long* d;
  int B, Q, D, V, pitch, K, hidden, flags;
  const float* table;
  const float* mu;
  const float* sigma;
  const float *w1, *b1, *w2, *b2;
  float* scores;
  float* feats;
  float* stats;
};

template <int KT, bool TRAIN>
__global__ void __launch_bounds__(NT, 1) knrm_kernel(const KnrmArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SimTile s = carve_sim_tile(smem_raw, a.pitch);
  float* sS = reinterpret_cast<float*>(smem_raw + sim_tile_bytes(a.pitch));  // [QT][KT] soft-TF
  float* sT1 = sS + QT * KT;                                                  // [QT][KT] sum K (s-mu)
  float* sT2 = sT1 + QT * KT;                                                 // [QT][KT] sum K (s-mu)^2
  float* sRow = sT2 + QT * KT;                                                // [QT] row sums of s
  float* sFeat = sRow + QT;                                                   // [KT]
  clear_sim_tile(s, tid);

  float mu[KT], cc[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const float sg = k < a.K ? a.sigma[k] : 1.f;
    mu[k] = k < a.K ? a.mu[k] : 0.f;
    cc[k] = -0.5f * 1.4426950408889634f / (sg * sg);
  }
  __syncthreads();

  constexpr int ROWS_PER_WARP = QT / (NT / 32);  // 4
  for (int pair = blockIdx.x; pair < a.B; pair += gridDim.x) {
    float S[ROWS_PER_WARP][KT], T1[TRAIN ? ROWS_PER_WARP : 1][KT], T2[TRAIN ? ROWS_PER_WARP : 1][KT], rs[ROWS_PER_WARP];
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
      rs[r] = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        S[r][k] = 0.f;
        if (TRAIN) T1[r][k] = 0.f, T2[r][k] = 0.f;
      }
    }
    const long long* qids = a.q + (size_t)pair * a.Q;
    const long long* dids = a.d + (size_t)pair * a.D;
    for (int d0 = 0; d0 < a.D; d0 += DT) {
      build_sim_tile(s, a.table, a.pitch, a.V, qids, a.Q, dids, d0, a.D, d0 == 0, tid);
      const int ncols = min(DT, a.D - d0);
#pragma unroll
      for (int r = 0; r < ROWS_PER_WARP; ++r) {
        const float* row = s.sim + (warp * ROWS_PER_WARP + r) * SIM_PITCH;
        for (int c = lane; c < ncols; c += 32) {
          const float v = row[c];
          rs[r] += v;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const float adj = v - mu[k];
            const float e = ex2_approx(cc[k] * adj * adj);
            S[r][k] += e;
            if (TRAIN) {
              const float ea = e * adj;
              T1[r][k] += ea;
              T2[r][k] = fmaf(ea, adj, T2[r][k]);
            }
          }
        }
      }
      __syncthreads();  // sim / id buffers are rewritten by the next tile or pair
    }
    // butterfly over the 32 lanes of the warp (fixed order)
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
      rs[r] = warp_sum(rs[r]);
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        S[r][k] = warp_sum(S[r][k]);
        if (TRAIN) T1[r][k] = warp_sum(T1[r][k]), T2[r][k] = warp_sum(T2[r][k]);
      }
      if (lane == 0) {
        const int q = warp * ROWS_PER_WARP + r;
        sRow[q] = rs[r];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          sS[q * KT + k] = S[r][k];
          if (TRAIN) sT1[q * KT + k] = T1[r][k], sT2[q * KT + k] = T2[r][k];
        }
      }
    }
    __syncthreads();
    if (tid < a.K) {
      const int k = tid;
      float R = 0.f, A = 0.f, C = 0.f;
      for (int q = 0; q < a.Q; ++q) {
        if (sRow[q] != 0.0f) {  // KNRM.py:51 -- "which query terms are not padding?"
          const float tf = sS[q * KT + k] + 1e-6f;
          R += logf(tf);
          if (TRAIN) {
            const float inv = 1.0f / tf;
            A = fmaf(inv, sT1[q * KT + k], A);
            C = fmaf(inv, sT2[q * KT + k], C);
          }
        }
      }
      sFeat[k] = R;
      if (a.feats) a.feats[(size_t)pair * a.K + k] = R;
      if (TRAIN && a.stats) {
        // dR_k/dmu_k = A / sigma^2 ; dR_k/dsigma_k = C / sigma^3   
        const float sg = a.sigma[k];
        a.stats[((size_t)pair * 2 + 0) * a.K + k] = A / (sg * sg);
        a.stats[((size_t)pair * 2 + 1) * a.K + k] = C / (sg * sg * sg);
      }
    }
    __syncthreads();
    if (a.scores && warp == 0) {
      float out;
      if (a.hidden == 0) {  // singlefc: Linear(K,1)   KNRM.py:28-29
        float p = lane < a.K ? a.w1[lane] * sFeat[lane] : 0.f;
        if (KT > 32) for (int k = lane + 32; k < a.K; k += 32) p = fmaf(a.w1[k], sFeat[k], p);
        out = warp_sum(p) + a.b1[0];
      } else {  // Linear(K,H) -> tanh -> Linear(H,1)   KNRM.py:31
        float p = 0.f;
        for (int h = lane; h < a.hidden; h += 32) {
          float acc = a.b1[h];
          for (int k = 0; k < a.K; ++k) acc = fmaf(a.w1[h * a.K + k], sFeat[k], acc);
          p = fmaf(a.w2[h], tanhf(acc), p);
        }
        out = warp_sum(p) + a.b2[0];
      }
      if (a.flags & CAPR_KNRM_SCORETANH) out = tanhf(out);  // KNRM.py:32-33
      if (lane == 0) a.scores[pair] = out;
    }
    // sFeat / sS are rewritten only after the next pair's __syncthreads()s
  }
}

// Debug / test kernel: materialise SimilarityMatrix.forward for a batch (common.py:170-182).
__global__ void __launch_bounds__(NT, 1) simmat_kernel(const long long* q, const long long* d, int B, int Q, int D,
                                                       const float* table, int V, int pitch, float* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  SimTile s = carve_sim_tile(smem_raw, pitch);
  clear_sim_tile(s, tid);
  __syncthreads();
  for (int pair = blockIdx.x; pair < B; pair += gridDim.x) {
    for (int d0 = 0; d0 < D; d0 += DT) {
      build_sim_tile(s, table, pitch, V, q + (size_t)pair * Q, Q, d + (size_t)pair * D, d0, D, d0 == 0, tid);
      const int ncols = min(DT, D - d0);
      for (int i = tid; i < Q * ncols; i += NT) {
        int r = i / ncols, c = i - r * ncols;
        out[((size_t)pair * Q + r) * D + d0 + c] = s.sim[r * SIM_PITCH + c];
      }
      __syncthreads();
    }
  }
}

static int check_common(const char* fn, const void* q, const void* d, int B, int Q, int D, const float* table, int V, int pitch) {
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d", fn, B, Q, D, V);
  CAPR_REQUIRE(pitch > 0 && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: table pitch %d must be a positive multiple of 16 (capr_table_pitch)", fn, pitch);
  CAPR_REQUIRE(B == 0 || (q && d && table), CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(((uintptr_t)table & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: table must be 16-byte aligned", fn);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(pitch <= MAX_PITCH, CAPR_ERR_UNSUPPORTED, "%s: embedding dim > %d is not supported by the fused kernels yet", fn, MAX_PITCH);
  return CAPR_OK;
}

template <typename Kern>
static int launch_persistent(Kern kern, size_t smem, int B, cudaStream_t st, const char* what, int& grid) {
  CAPR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", what);
  grid = B < sms ? B : sms;
  return CAPR_OK;
}

}  // namespace capr

using namespace capr;

extern "C" {

int capr_simmat_forward(const int64_t* query, const int64_t* doc, int B, int Q, int D, const float* table, int V,
                        int pitch, float* sim, capr_stream_t stream) {
  capr::DeviceGuard device_guard(table);  // act on the device that owns the caller's buffers
  int rc = check_common("capr_simmat_forward", query, doc, B, Q, D, table, V, pitch);
  if (rc) return rc;
  CAPR_REQUIRE(sim, CAPR_ERR_BAD_POINTER, "capr_simmat_forward: null output");
  if (B == 0) return CAPR_OK;
  size_t smem = sim_tile_bytes(pitch);
  int grid = 0;
  rc = launch_persistent(simmat_kernel, smem, B, (cudaStream_t)stream, "capr_simmat_forward", grid);
  if (rc) return rc;
  simmat_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>((const long long*)query, (const long long*)doc, B, Q, D, table, V, pitch, sim);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

int capr_knrm_forward(const int64_t* query, const int64_t* doc, int B, int Q, int D, const float* table, int V,
                      int pitch, const float* mu, const float* sigma, int K, const float* w1, const float* b1,
                      int hidden, const float* w2, const float* b2, int flags, float* scores, float* feats,
                      float* stats, capr_stream_t stream) {
  capr::DeviceGuard device_guard(table);  // act on the device that owns the caller's buffers
  int rc = check_common("capr_knrm_forward", query, doc, B, Q, D, table, V, pitch);
  if (rc) return rc;
  CAPR_REQUIRE(K > 0 && hidden >= 0, CAPR_ERR_BAD_SHAPE, "capr_knrm_forward: K=%d hidden=%d", K, hidden);
  CAPR_REQUIRE(K <= 32, CAPR_ERR_UNSUPPORTED, "capr_knrm_forward: more than 32 kernels (K=%d) is not supported", K);
  CAPR_REQUIRE(mu && sigma, CAPR_ERR_BAD_POINTER, "capr_knrm_forward: null mu/sigma");
  CAPR_REQUIRE(B == 0 || scores || feats || stats, CAPR_ERR_BAD_POINTER, "capr_knrm_forward: no output requested");
  if (scores) {
    CAPR_REQUIRE(w1 && b1, CAPR_ERR_BAD_POINTER, "capr_knrm_forward: scores requested without combine weights");
    CAPR_REQUIRE(hidden == 0 || (w2 && b2), CAPR_ERR_BAD_POINTER, "capr_knrm_forward: hidden=%d needs w2/b2", hidden);
  }
  if (B == 0) return CAPR_OK;
  KnrmArgs a{(const long long*)query, (const long long*)doc, B, Q, D, V, pitch, K, hidden, flags, table, mu, sigma, w1, b1, w2, b2, scores, feats, stats};
  const bool train = stats != nullptr;
  const int KT = K <= 11 ? 11 : (K <= 16 ? 16 : 32);
  size_t smem = sim_tile_bytes(pitch) + (size_t)(3 * QT * KT + QT + KT) * sizeof(float);
  int grid = 0;
  cudaStream_t st = (cudaStream_t)stream;
#define CAPR_LAUNCH_KNRM(KT_, TR_)                                                              \
  do {                                                                                          \
    rc = launch_persistent(knrm_kernel<KT_, TR_>, smem, B, st, "capr_knrm_forward", grid);      \
    if (rc) return rc;                                                                          \
    knrm_kernel<KT_, TR_><<<grid, NT, smem, st>>>(a);                                           \
  } while (0)
  if (KT == 11) { if (train) CAPR_LAUNCH_KNRM(11, true); else CAPR_LAUNCH_KNRM(11, false); }
  else if (KT == 16) { if (train) CAPR_LAUNCH_KNRM(16, true); else CAPR_LAUNCH_KNRM(16, false); }
  else { if (train) CAPR_LAUNCH_KNRM(32, true); else CAPR_LAUNCH_KNRM(32, false); }
#undef CAPR_LAUNCH_KNRM
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

}  // extern "C"
