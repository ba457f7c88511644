// K1 (engine 3) -- fused KNRM scoring: term-frequency documents, cosine tile on tcgen05, kernel pooling straight from TMEM.
//
//   KNRM_class.forward            capreolus/reranker/KNRM.py:39-55
//   RbfKernelBank                 capreolus/reranker/common.py:224-250
//   SimilarityMatrix              capreolus/reranker/common.py:143-182  (producer / MMA roles: simtc3.cuh)
//
// Same math and the same quirks as knrm.cu / knrm_tc.cu: the kernel sums run over ALL doc positions (padded ones have s = 0 and
// still contribute exp(-mu^2 / 2 sigma^2), KNRM.py:50), a query row is live iff its cosine row-sum is not exactly 0 (KNRM.py:51),
// log(S + 1e-6), combine.  The sum over doc positions is evaluated in term-frequency form,
//     S_k[i] = sum over DISTINCT doc tokens t of  count(t) * exp(-(s_it - mu_k)^2 / 2 sigma_k^2),
// which is the same sum with equal terms grouped (identical tokens have identical cosines; <pad> is a token with s = 0).
#include "simtc3.cuh"

namespace capr {

// ---------------------------------------------------------------------------------------------------------------------------
// pre-pass: doc [B,D] int64 -> (distinct token, count) lists in first-occurrence order + their number
// ---------------------------------------------------------------------------------------------------------------------------
// One CTA of 128 threads per pair; open-addressing hash table in shared memory (2 048 slots for <= 1 024 tokens).  Determinism:
// a token's rank is the number of distinct tokens that first occur before it, which does not depend on the order in which the
// threads insert (the slot keeps the MINIMUM position).  HBM traffic: 8 B read + 6 B written per token (~1 % of a scoring step).
constexpr int TF_THREADS = 128;
constexpr int TF_SLOTS = 2048;
constexpr int TF_EMPTY = (int)0x80000000;  // id_as_int never returns INT_MIN

__global__ void __launch_bounds__(TF_THREADS) tf_dedup_kernel(const long long* __restrict__ doc, int B, int D, int dedup, int* __restrict__ out_ids,
                                                              unsigned short* __restrict__ out_cnt, int* __restrict__ out_nd) {
  __shared__ int keys[TF_SLOTS];
  __shared__ int pos[TF_SLOTS];
  __shared__ int cnt[TF_SLOTS];
  __shared__ int wcount[simtc3::IDS_PER_THREAD][TF_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int pair = blockIdx.x; pair < B; pair += gridDim.x) {
    const long long* d = doc + (size_t)pair * D;
    int* oi = out_ids + (size_t)pair * D;
    unsigned short* oc = out_cnt + (size_t)pair * D;
    if (!dedup) {  // identity: every position is its own "token" (A/B switch CAPR_KNRM_TF=0)
      for (int i = tid; i < D; i += TF_THREADS) oi[i] = id_as_int(d[i]), oc[i] = 1;
      if (tid == 0) out_nd[pair] = D;
      continue;
    }
    for (int sl = tid; sl < TF_SLOTS; sl += TF_THREADS) keys[sl] = TF_EMPTY, pos[sl] = 0x7fffffff, cnt[sl] = 0;
    __syncthreads();
    int id[simtc3::IDS_PER_THREAD], slot[simtc3::IDS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < simtc3::IDS_PER_THREAD; ++j) {
      const int i = tid + TF_THREADS * j;
      id[j] = 0, slot[j] = -1;
      if (i < D) {
        id[j] = id_as_int(d[i]);
        unsigned h = ((unsigned)id[j] * 2654435761u) >> 21;  // 11 bits
        while (true) {
          const int old = atomicCAS(&keys[h], TF_EMPTY, id[j]);
          if (old == TF_EMPTY || old == id[j]) break;
          h = (h + 1) & (TF_SLOTS - 1);
        }
        slot[j] = (int)h;
        atomicMin(&pos[h], i);
        atomicAdd(&cnt[h], 1);
      }
    }
    __syncthreads();
    unsigned ballots[simtc3::IDS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < simtc3::IDS_PER_THREAD; ++j) {
      const int i = tid + TF_THREADS * j;
      const bool first = slot[j] >= 0 && pos[slot[j]] == i;
      ballots[j] = __ballot_sync(0xffffffffu, first);
      if (lane == 0) wcount[j][warp] = __popc(ballots[j]);
    }
    __syncthreads();
    // positions are ordered (j, warp, lane): i = 128 j + 32 warp + lane; rank = distinct tokens that first occur before position i
    int total = 0;
#pragma unroll
    for (int j = 0; j < simtc3::IDS_PER_THREAD; ++j) {
      int before = total;
#pragma unroll
      for (int w = 0; w < TF_THREADS / 32; ++w) {
        const int c = wcount[j][w];
        before += w < warp ? c : 0;
        total += c;
      }
      if ((ballots[j] >> lane) & 1u) {
        const int rank = before + __popc(ballots[j] & ((1u << lane) - 1u));
        oi[rank] = id[j];
        oc[rank] = (unsigned short)cnt[slot[j]];
      }
    }
    for (int r = total + tid; r < D; r += TF_THREADS) oi[r] = 0, oc[r] = 0;
    if (tid == 0) out_nd[pair] = total;
    __syncthreads();  // the table is cleared for the next pair
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// scoring kernel
// ---------------------------------------------------------------------------------------------------------------------------
struct KnrmTc3Args {
  simtc3::Problem pr;
  int K, hidden, flags;
  const float* mu;
  const float* sigma;
  const float *w1, *b1, *w2, *b2;
  float* scores;
  float* feats;
};

template <int KT>
__global__ void __launch_bounds__(simtc3::THREADS, 1) knrm_tc3_kernel(const KnrmTc3Args a) {
  using namespace simtc3;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Problem& pr = a.pr;
  constexpr int RED_FLOATS = POOL_WARPS * (KT + 1) * 32;
  Smem s = carve(smem_raw, pr, (pr.pitch + ATOM_K - 1) / ATOM_K, RED_FLOATS);
  const uint32_t tmem_base = setup(s, pr, tid);

  if (warp >= POOL_WARPS && warp < MMA_WARP) {
    producer_loop(s, pr, (warp - POOL_WARPS) * 32 + lane);
  } else if (warp == MMA_WARP) {
    mma_loop(s, pr, tmem_base, lane);
  } else if (warp < POOL_WARPS) {
    // ===================== pooling: lane = query row; warp = (lane quarter, 64-column half) of the quarter's units ===========
    // Each thread keeps the K running sums of ITS row over ITS columns (K + 1 accumulators, no shuffles in the loop); per
    // column: w = count, rs += w s, S_k += w exp2(c_k (s - mu_k)^2) with c_k = -1/2 log2(e) / sigma_k^2 from the live parameters.
    const int quarter = warp & 3, chalf = warp >> 2;
    float mu[KT], cc[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float sg = k < a.K ? a.sigma[k] : 1.f;
      mu[k] = k < a.K ? a.mu[k] : 0.f;
      cc[k] = -0.5f * 1.4426950408889634f / (sg * sg);
    }
    int it = 0, g0 = 0;
    const int tr_role = warp == 0 ? 2 : 3;
    int tr_n = (lane == 0 && (warp == 0 || warp == 5)) ? 0 : 1024;
    (void)tr_role, (void)tr_n;
    for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
      const int pp = it & 1;
      SIM3_TR(tr_role, 20);
      tc::mbar_wait(&s.ids_full[pp], (uint32_t)((it >> 1) & 1));
      SIM3_TR(tr_role, 21);
      const int nd = s.nd[pp];
      const int qi = s.qid[pp * QT + lane];
      const int units = units_of(nd);
      float S[KT], rs = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k) S[k] = 0.f;
      for (int u = 0; u < units; ++u) {
        const int g = g0 + u;
        if ((g & 3) != quarter) continue;
        tc::mbar_wait(&s.acc_full[quarter], (uint32_t)((g >> 2) & 1));
        SIM3_TR(tr_role, 22);
        tc::tc_fence_after();
        const int live = live_cols(nd, u);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int col0 = chalf * 64 + c * 32;
          if (col0 >= live || CAPR_DBG(pr.debug & 8)) break;
          const int doc0 = pp * pr.dcap + u * U_DOCS + col0;
          float v[32];
          load_cosines(tmem_base, quarter, col0, live, qi, s.did + doc0, v);
          const unsigned short* cw = s.cnt + doc0;
          if (CAPR_DBG(pr.debug & 1)) {
            rs += v[lane & 31];
            continue;
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {  // (fully unrolled: v[] must stay in registers)
            const uint2 c4 = *reinterpret_cast<const uint2*>(cw + 4 * j4);  // 4 counts (uint16), warp-wide broadcast
            const float w[4] = {count_as_float(c4.x & 0xffffu), count_as_float(c4.x >> 16), count_as_float(c4.y & 0xffffu), count_as_float(c4.y >> 16)};
            const float x[4] = {v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]};
            rs += (w[0] * x[0] + w[1] * x[1]) + (w[2] * x[2] + w[3] * x[3]);
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              float e[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float adj = x[j] - mu[k];
                e[j] = ex2_approx(cc[k] * adj * adj);
              }
              S[k] += (w[0] * e[0] + w[1] * e[1]) + (w[2] * e[2] + w[3] * e[3]);
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&s.acc_empty[quarter]);
        SIM3_TR(tr_role, 23);
      }
      // hand the per-row partial sums to the finisher: red[slot][k][lane], slot = (PAIR-LOCAL unit index mod 4, column half) --
      // the quarter that pooled the pair's unit u is (g0 + u) % 4, which depends on the pairs this CTA scored before; indexing
      // the partials by u % 4 makes the order of the final sum, and therefore the bits of the score, a function of the pair alone
      tc::mbar_wait(s.red_empty, (uint32_t)((it & 1) ^ 1));
      SIM3_TR(tr_role, 24);
      const int slot = ((quarter - g0) & 3) + 4 * chalf;  // g0 here = first global unit of the pair
      float* mine = s.red + slot * (KT + 1) * 32 + lane;
#pragma unroll
      for (int k = 0; k < KT; ++k) mine[k * 32] = S[k];
      mine[KT * 32] = rs;
      g0 += units;
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(s.red_full);
        tc::mbar_arrive(&s.ids_empty[pp]);
      }
    }
  } else {
    // ===================== finisher: lane = query row ======================================================================
    // t_k = sum over the 8 partial-sum slots (pair-local unit index mod 4, column half) in a fixed order; live rows (cosine row-sum != 0, KNRM.py:51) contribute log(t_k + 1e-6);
    // R_k = butterfly sum over the rows (fixed order -> bit-reproducible); combine by lanes k < K.
    int it = 0;
    int tr_n = lane == 0 ? 0 : 1024;
    (void)tr_n;
    for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
      tc::mbar_wait(s.red_full, (uint32_t)(it & 1));
      SIM3_TR(4, 30);
      float t[KT + 1];
#pragma unroll
      for (int k = 0; k <= KT; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < POOL_WARPS; ++w) acc += s.red[(w * (KT + 1) + k) * 32 + lane];
        t[k] = acc;
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(s.red_empty);
      const bool live = t[KT] != 0.0f && lane < pr.Q;
      float R = 0.f;  // lane k ends up with R_k
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const float r = warp_sum(live ? logf(t[k] + 1e-6f) : 0.f);
        R = lane == k ? r : R;
      }
      if (lane < a.K && a.feats) a.feats[(size_t)pair * a.K + lane] = R;
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
      SIM3_TR(4, 31);
    }
  }
  teardown(s, tmem_base, tid);
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace capr

using namespace capr;

extern "C" {

int capr_tf_dedup(const int64_t* doc, int B, int D, int32_t* ids, uint16_t* counts, int32_t* n_distinct, capr_stream_t stream) {
  capr::DeviceGuard device_guard(doc);
  const char* fn = "capr_tf_dedup";
  CAPR_REQUIRE(B >= 0 && D > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d D=%d", fn, B, D);
  CAPR_REQUIRE(D <= simtc3::MAX_DCAP, CAPR_ERR_UNSUPPORTED, "%s: maxdoclen=%d > %d", fn, D, simtc3::MAX_DCAP);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(doc && ids && counts && n_distinct, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  tf_dedup_kernel<<<B < sms * 8 ? B : sms * 8, TF_THREADS, 0, (cudaStream_t)stream>>>((const long long*)doc, B, D, 1, ids, counts, n_distinct);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

size_t capr_tf_workspace_bytes(int B, int D) {
  if (B <= 0 || D <= 0) return 0;
  return align256((size_t)B * D * 4) + align256((size_t)B * D * 2) + align256((size_t)B * 4);
}

int capr_knrm_forward_tf(const int64_t* query, const int64_t* doc, int B, int Q, int D, const void* table_hi, const void* table_lo, int V,
                         int E, int pitch, const float* mu, const float* sigma, int K, const float* w1, const float* b1, int hidden,
                         const float* w2, const float* b2, int flags, float* scores, float* feats, void* workspace, size_t workspace_bytes,
                         capr_stream_t stream) {
  capr::DeviceGuard device_guard(table_hi);
  const char* fn = "capr_knrm_forward_tf";
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0 && E > 0 && K > 0 && hidden >= 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d E=%d K=%d", fn, B, Q, D, V, E, K);
  CAPR_REQUIRE(pitch >= E && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: pitch=%d must be a multiple of 16 and >= E (capr_table_pitch_bf16)", fn, pitch);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(D <= simtc3::MAX_DCAP, CAPR_ERR_UNSUPPORTED, "%s: maxdoclen=%d > %d: use capr_knrm_forward (doc-tiled FFMA engine)", fn, D, simtc3::MAX_DCAP);
  CAPR_REQUIRE(pitch <= simtc3::MAX_ATOMS * simtc3::ATOM_K, CAPR_ERR_UNSUPPORTED, "%s: embedding dim > %d: use capr_knrm_forward", fn, simtc3::MAX_ATOMS * simtc3::ATOM_K);
  CAPR_REQUIRE(K <= 16, CAPR_ERR_UNSUPPORTED, "%s: K=%d > 16 kernels: use capr_knrm_forward", fn, K);
  CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table of %d x %d elements is too large for 32-bit row offsets", fn, V, pitch);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(query && doc && table_hi && table_lo && mu && sigma && workspace, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE((((uintptr_t)table_hi | (uintptr_t)table_lo) & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: table planes must be 16-byte aligned", fn);
  CAPR_REQUIRE(((uintptr_t)workspace & 255) == 0, CAPR_ERR_BAD_POINTER, "%s: workspace must be 256-byte aligned", fn);
  CAPR_REQUIRE(scores || feats, CAPR_ERR_BAD_POINTER, "%s: no output requested", fn);
  if (scores) {
    CAPR_REQUIRE(w1 && b1, CAPR_ERR_BAD_POINTER, "%s: scores requested without combine weights", fn);
    CAPR_REQUIRE(hidden == 0 || (w2 && b2), CAPR_ERR_BAD_POINTER, "%s: hidden=%d needs w2/b2", fn, hidden);
  }
  // pairs per chunk the workspace can hold (any workspace >= one pair works: the call loops)
  size_t chunk = (size_t)B;
  while (chunk > 1 && capr_tf_workspace_bytes((int)chunk, D) > workspace_bytes) chunk = (chunk + 1) / 2;
  CAPR_REQUIRE(capr_tf_workspace_bytes((int)chunk, D) <= workspace_bytes, CAPR_ERR_BAD_SHAPE, "%s: workspace too small (capr_tf_workspace_bytes(1, D) = %zu bytes at least)", fn,
               capr_tf_workspace_bytes(1, D));
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const int KT = K <= 11 ? 11 : 16;
  const int atoms = (pitch + simtc3::ATOM_K - 1) / simtc3::ATOM_K;
  const int dcap = (D + simtc3::U_DOCS - 1) / simtc3::U_DOCS * simtc3::U_DOCS;
  const int red_floats = simtc3::POOL_WARPS * (KT + 1) * 32;
  // two query buffers (no bubble while the next pair's query block is gathered) or one (three more stages in flight)
  const char* qb_env = getenv("CAPR_SIM3_QBUFS");
  int n_qbufs = (qb_env && qb_env[0] == '2') ? 2 : 1;
  int n_stages = simtc3::stages_that_fit(atoms, n_qbufs, dcap, red_floats);
  if (n_stages < 2 && n_qbufs == 2) n_qbufs = 1, n_stages = simtc3::stages_that_fit(atoms, 1, dcap, red_floats);
  CAPR_REQUIRE(n_stages >= 2, CAPR_ERR_UNSUPPORTED, "%s: shared-memory budget exceeded (D=%d, pitch=%d)", fn, D, pitch);
  if (const char* st_env = getenv("CAPR_SIM3_STAGES")) {
    const int want = atoi(st_env);
    if (want >= 2 && want < n_stages) n_stages = want;
  }
  const size_t smem = simtc3::smem_bytes(atoms, n_stages, n_qbufs, dcap, red_floats);
  const char* tf_env = getenv("CAPR_KNRM_TF");
  const int dedup = !(tf_env && tf_env[0] == '0');
  cudaStream_t st = (cudaStream_t)stream;
  if (KT == 11) CAPR_CHECK_CUDA(cudaFuncSetAttribute(knrm_tc3_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else CAPR_CHECK_CUDA(cudaFuncSetAttribute(knrm_tc3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (size_t lo = 0; lo < (size_t)B; lo += chunk) {
    const int n = (int)((size_t)B - lo < chunk ? (size_t)B - lo : chunk);
    unsigned char* ws = (unsigned char*)workspace;
    int* tf_ids = (int*)ws;
    unsigned short* tf_cnt = (unsigned short*)(ws + align256((size_t)chunk * D * 4));
    int* tf_nd = (int*)(ws + align256((size_t)chunk * D * 4) + align256((size_t)chunk * D * 2));
    const int pre_grid = n < sms * 8 ? n : sms * 8;
    tf_dedup_kernel<<<pre_grid, TF_THREADS, 0, st>>>((const long long*)doc + lo * D, n, D, dedup, tf_ids, tf_cnt, tf_nd);
    CAPR_CHECK_CUDA(cudaGetLastError());
    KnrmTc3Args a{};
    a.pr = simtc3::Problem{(const long long*)query + lo * Q, tf_ids, tf_cnt, tf_nd, n, Q, D, V, (const __nv_bfloat16*)table_hi, (const __nv_bfloat16*)table_lo,
                           pitch, E, n_stages, n_qbufs, dcap, 0, nullptr};
#ifdef CAPR_DEBUG_BUILD
    if (const char* de = getenv("CAPR_SIM3_DEBUG")) a.pr.debug = atoi(de);
    if (const char* te = getenv("CAPR_SIM3_TRACE")) a.pr.trace = (long long*)strtoull(te, nullptr, 10);
#endif
    a.K = K, a.hidden = hidden, a.flags = flags, a.mu = mu, a.sigma = sigma, a.w1 = w1, a.b1 = b1, a.w2 = w2, a.b2 = b2;
    a.scores = scores ? scores + lo : nullptr;
    a.feats = feats ? feats + lo * K : nullptr;
    const int grid = n < sms ? n : sms;
    if (KT == 11) knrm_tc3_kernel<11><<<grid, simtc3::THREADS, smem, st>>>(a);
    else knrm_tc3_kernel<16><<<grid, simtc3::THREADS, smem, st>>>(a);
    CAPR_CHECK_CUDA(cudaGetLastError());
  }
  return CAPR_OK;
}

}  // extern "C"
