// K1 (engine 2) -- fused KNRM scoring with the cosine tile on tcgen05 tensor cores.
//
//   KNRM_class.forward            capreolus/reranker/KNRM.py:39-55
//   RbfKernelBank                 capreolus/reranker/common.py:224-250
//   SimilarityMatrix              capreolus/reranker/common.py:143-182  (producer: simtc.cuh)
//
// Same math and same quirks as knrm.cu (sum over ALL doc positions, live-row test on the cosine row-sum, log(S+1e-6),
// combine); only the producer of the cosine tile differs: gather -> UMMA -> TMEM instead of gather -> FFMA.  The
// epilogue is a pipeline of its own (simtc.cuh): 4 warps drain TMEM into half tiles of 32 x 256 cosines, 8 warps pool
// them (MUFU ex2 bound, 180 k exponentials per pair), while the producer and MMA warps work on the next pair.
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
__global__ void __launch_bounds__(simtc::THREADS_PIPE, 1) knrm_tc_kernel(const KnrmTcArgs a) {
  using namespace simtc;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem s = carve(smem_raw, (a.pr.pitch + ATOM_K - 1) / ATOM_K, a.pr.deep != 0);
  float* sPart = s.extra;  // [2][POOL_WARPS][KT] per-warp partial features, double-buffered by pair parity
  const uint32_t tmem_base = setup(s, tid, THREADS_PIPE, MMA_WARP_PIPE);

  if (is_producer_warp(warp)) {
    producer_loop(s, a.pr, producer_index(warp) * 32 + lane);
  } else if (warp == MMA_WARP_PIPE) {
    mma_loop(s, a.pr, tmem_base);
  } else if (is_drain_warp(warp)) {
    drain_loop(s, a.pr, tmem_base, warp, lane);
  } else if (is_pool_warp(warp)) {
    // ===================== pooling: 8 warps; lane = query row, warp = a 32-column slice of every half tile ============
    // Each thread keeps the K running sums of ITS row over ITS columns (K + 1 accumulators, no shuffles in the loop; the
    // one-row-per-lane float4 reads are conflict-free like the drain's stores).  At the end of the pair the 8 column
    // slices are combined through shared memory: every warp parks its [32 rows][K+1] partials in its own slice of the
    // last half tile (only this warp ever read those columns), then warp w reduces rows 4w..4w+3 in a fixed order.
    const int pw = pool_index(warp);
    float mu[KT], cc[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float sg = k < a.K ? a.sigma[k] : 1.f;
      mu[k] = k < a.K ? a.mu[k] : 0.f;
      cc[k] = -0.5f * 1.4426950408889634f / (sg * sg);
    }
    constexpr int ROWS_PER_WARP = QT / POOL_WARPS;  // 4 (second stage)
    constexpr int SLICE = NT_DOCS / POOL_WARPS;     // 32 columns per warp and half
    static_assert(KT + 1 <= SLICE, "partials of one row must fit in the warp's slice");
    const int halves = halves_of(a.pr);
    PoolSync ps;
    int unit = 0, it = 0;
    for (int pair = blockIdx.x; pair < a.pr.B; pair += gridDim.x, ++it) {
      float S[KT], rs = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k) S[k] = 0.f;
      int ub = 0;
      for (int h = 0; h < halves; ++h, ++unit) {
        ub = unit & 1;
        ps.wait_full(s, ub);
        const int nvalid = min(NT_DOCS, a.pr.D - h * NT_DOCS) - pw * SLICE;  // columns of this slice that exist
        if (!CAPR_DBG(a.flags & 0x100 /*CAPR_DEBUG_SKIP_POOL*/) && nvalid > 0) {
          const float4* row = reinterpret_cast<const float4*>(half_tile(s, ub) + lane * HALF_PITCH + pw * SLICE);
#pragma unroll 2
          for (int g = 0; g < SLICE / 4; ++g) {
            const float4 x = row[g];
            float v[4] = {x.x, x.y, x.z, x.w};
            if (nvalid < SLICE) {  // warp-uniform: ragged tail of the doc tile
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const bool valid = g * 4 + j < nvalid;
                rs += valid ? v[j] : 0.f;
                v[j] = valid ? v[j] : 1e18f;  // every kernel underflows to exactly 0
              }
            } else {
              rs += (v[0] + v[1]) + (v[2] + v[3]);
            }
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              float e[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float adj = v[j] - mu[k];
                e[j] = ex2_approx(cc[k] * adj * adj);
              }
              S[k] += (e[0] + e[1]) + (e[2] + e[3]);
            }
          }
        }
        if (h + 1 < halves) ps.release(s, ub, lane);  // the last half tile is kept for the cross-warp reduction
      }
      float* tile = half_tile(s, ub);
      {
        float* mine = tile + lane * HALF_PITCH + pw * SLICE;
#pragma unroll
        for (int k = 0; k < KT; ++k) mine[k] = S[k];
        mine[KT] = rs;
      }
      epi_barrier();  // the 8 pooling warps
      // lane k accumulates this warp's share of R_k = sum over its live rows of log(S_k + 1e-6)   (KNRM.py:50-53)
      float R_part = 0.f;
#pragma unroll
      for (int r = 0; r < ROWS_PER_WARP; ++r) {
        const int qrow = pw * ROWS_PER_WARP + r;
        float t = 0.f;
        if (lane <= KT) {
#pragma unroll
          for (int w = 0; w < POOL_WARPS; ++w) t += tile[qrow * HALF_PITCH + w * SLICE + lane];  // fixed order over the slices
        }
        const float rsum = __shfl_sync(0xffffffffu, t, KT);
        if (rsum != 0.0f && qrow < a.pr.Q) R_part += logf(t + 1e-6f);  // KNRM.py:51-52
      }
      ps.release(s, ub, lane);
      float* part = sPart + (it & 1) * POOL_WARPS * KT;
      if (lane < KT) part[pw * KT + lane] = R_part;
      epi_barrier();  // part[] of this parity is rewritten two pairs (= at least two barriers) later
      if (pw == 0) {
        float R = 0.f;
        if (lane < a.K) {
#pragma unroll
          for (int w = 0; w < POOL_WARPS; ++w) R += part[w * KT + lane];  // fixed order over the 8 row groups
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
    }
  }
  teardown(s, tmem_base, tid, MMA_WARP_PIPE);
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
  capr::DeviceGuard device_guard(emb);  // act on the device that owns the caller's buffers
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
  capr::DeviceGuard device_guard(table_hi);  // act on the device that owns the caller's buffers
  const char* fn = "capr_knrm_forward_tc";
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0 && E > 0 && K > 0 && hidden >= 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d E=%d K=%d", fn, B, Q, D, V, E, K);
  CAPR_REQUIRE(pitch >= E && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: pitch=%d must be a multiple of 16 and >= E (capr_table_pitch_bf16)", fn, pitch);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(D <= simtc::DEEP_DCAP, CAPR_ERR_UNSUPPORTED, "%s: maxdoclen=%d > %d: use capr_knrm_forward (doc-tiled FFMA engine)", fn, D, simtc::DEEP_DCAP);
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
  const int atoms = (pitch + simtc::ATOM_K - 1) / simtc::ATOM_K;
  // deep ring (one query buffer, three doc stages) whenever the second query buffer is worth less than a doc stage;
  // CAPR_SIM_RING=2 forces the default layout (A/B tests)
  const char* ring_env = getenv("CAPR_SIM_RING");
  a.pr.deep = (D > DT || (atoms >= 3 && !(ring_env && ring_env[0] == '2'))) ? 1 : 0;  // maxdoclen > 512 needs the deep layout's id arrays
  const size_t smem = simtc::smem_bytes(atoms, (size_t)(2 * simtc::POOL_WARPS * KT) * sizeof(float), a.pr.deep != 0);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const int grid = B < sms ? B : sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (KT == 11) {
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(knrm_tc_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knrm_tc_kernel<11><<<grid, simtc::THREADS_PIPE, smem, st>>>(a);
  } else {
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(knrm_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knrm_tc_kernel<16><<<grid, simtc::THREADS_PIPE, smem, st>>>(a);
  }
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

}  // extern "C"
