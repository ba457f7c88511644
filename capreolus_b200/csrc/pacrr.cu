// K3 -- fused PACRR scoring kernel.
//
//   PACRRConvMax2dModule.forward   capreolus/reranker/PACRR.py:73-82   pad -> Conv2d(1,F,n) -> ReLU -> max_f -> top-k_d
//   PACRR_class.forward            capreolus/reranker/PACRR.py:43-54   n = mingram..maxgram, + softmax(idf), 3-layer MLP
//   SimilarityMatrix               capreolus/reranker/common.py:143-182 (producer: simtile.cuh)
//
// The cosine tile stays in shared memory with a zero halo (the reference pads bottom/right with zeros,
// PACRR.py:64), so each n x n window is n*n conflict-free LDS.  The F x n x n filter taps live in
// __constant__ memory: the inner loop is FFMA with a constant-bank operand, no loads at all.
// relu(max_f x_f) == max_f relu(x_f), so ReLU is applied once after the filter max.  Each lane keeps a
// running top-k of its columns; lanes are merged with k rounds of warp max.  The reference's [B,F,Q,D]
// conv output (2 MB per pair and n-gram) never exists.
#include "simtc.cuh"

namespace capr {

constexpr int MAX_NGRAM = 5;     // largest conv window
constexpr int MAX_FILTERS = 64;
constexpr int MAX_KMAX = 8;
constexpr int MAX_GRAMS = 5;     // number of n-gram modules
constexpr int MAX_COMBINE = 128;
constexpr int CONV_FLOATS = MAX_FILTERS * (1 + 4 + 9 + 16 + 25);
// Window size n owns a FIXED slot of constant memory, so that with N and the filter index unrolled every tap is an
// immediate constant-bank operand of its FFMA (no LDC, no address arithmetic in the inner loop).
__host__ __device__ constexpr int conv_w_slot(int n) { return MAX_FILTERS * ((n - 1) * n * (2 * n - 1) / 6); }  // sum_{m<n} m^2
__host__ __device__ constexpr int conv_b_slot(int n) { return MAX_FILTERS * (n - 1); }

__constant__ float c_conv_w[CONV_FLOATS];
__constant__ float c_conv_b[MAX_NGRAM * MAX_FILTERS];

struct PacrrArgs {
  const long long* q;
  const long long* d;
  const float* idf;
  int B, Q, D, V, pitch, mingram, maxgram, F, kmax, combine, nonlin;
  int w_off[MAX_GRAMS];
  const float* table;
  const float *l1w, *l1b, *l2w, *l2b, *l3w, *l3b;
  float* scores;
  float* topk_out;
  simtc::Problem pr;  // tensor-core engine only
};

struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct EpiSync {
  __device__ __forceinline__ void operator()() const { simtc::epi_barrier(); }
};

__device__ __forceinline__ float act(float x, int nonlin) { return nonlin == 1 ? fmaxf(x, 0.f) : (nonlin == 2 ? tanhf(x) : x); }

// One n-gram module over the rows of this warp.  feat layout: [QT][qterm], this module writes columns
// [col0, col0 + kmax).
// The 4 query rows of the warp are convolved together and every lane takes TWO doc columns per step (c and c + 32),
// packed in the two halves of FFMA2 operands: a filter tap (a uniform-register scalar, broadcast by the instruction)
// feeds 4 FFMA2 = 8 multiply-adds, and the windows of the 4 rows share their shared-memory loads.
template <int N, int FT, int KM, int R>
__device__ __forceinline__ void ngram_rows(const float* sim, const PacrrArgs& a, int F, float* feat, int qterm, int col0, int q0, int lane) {
  constexpr int w_off = conv_w_slot(N), b_off = conv_b_slot(N);
  if (q0 >= a.Q) return;  // warp-uniform
  float top[R][KM];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < KM; ++k) top[r][k] = -INFINITY;
  for (int c = lane; c < a.D; c += 64) {
    float2 win[R + N - 1][N];  // .x: column c, .y: column c + 32 (inside the zero halo when past the doc)
#pragma unroll
    for (int u = 0; u < R + N - 1; ++u)
#pragma unroll
      for (int v = 0; v < N; ++v) {
        const float* p = sim + (q0 + u) * SIM_PITCH + c + v;
        win[u][v] = make_float2(p[0], p[32]);
      }
    float2 best[R];
#pragma unroll
    for (int r = 0; r < R; ++r) best[r] = make_float2(-INFINITY, -INFINITY);
    auto one_filter = [&](int f) {
      float2 x[R];
      const float bias = c_conv_b[b_off + f];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = make_float2(bias, bias);
#pragma unroll
      for (int u = 0; u < N; ++u)
#pragma unroll
        for (int v = 0; v < N; ++v) {
          const float w = c_conv_w[w_off + f * N * N + u * N + v];
#pragma unroll
          for (int r = 0; r < R; ++r) x[r] = fma2_bcast(w, win[r + u][v], x[r]);
        }
#pragma unroll
      for (int r = 0; r < R; ++r) best[r] = make_float2(fmaxf(best[r].x, x[r].x), fmaxf(best[r].y, x[r].y));
    };
    if (FT > 0) {
#pragma unroll
      for (int f = 0; f < FT; ++f) one_filter(f);
    } else {
      for (int f = 0; f < F; ++f) one_filter(f);
    }
    const bool second = c + 32 < a.D;
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        // ReLU (PACRR.py:78) commutes with the filter max (PACRR.py:79)
        float v = half == 0 ? fmaxf(best[r].x, 0.f) : (second ? fmaxf(best[r].y, 0.f) : -INFINITY);
#pragma unroll
        for (int k = 0; k < KM; ++k) {  // insert into the lane-local descending top-k
          if (k < a.kmax && v > top[r][k]) {
            const float t = top[r][k];
            top[r][k] = v;
            v = t;
          }
        }
      }
    }
  }
  // merge the 32 lane-local lists of every row: kmax rounds of (warp max, owner pops its head)
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (q0 + r >= a.Q) break;  // warp-uniform
    for (int k = 0; k < a.kmax; ++k) {
      const float m = warp_max(top[r][0]);
      const unsigned owners = __ballot_sync(0xffffffffu, top[r][0] == m);
      if (lane == (__ffs(owners) - 1)) {
#pragma unroll
        for (int j = 0; j < KM - 1; ++j) top[r][j] = top[r][j + 1];
        top[r][KM - 1] = -INFINITY;
      }
      if (lane == 0) feat[(q0 + r) * qterm + col0 + k] = m;
    }
  }
}

// 4 rows at a time for the windows the reference uses (n <= 3); 2 + 2 for the larger ones (register budget)
template <int N, int FT, int KM>
__device__ __forceinline__ void ngram_pass(const float* sim, const PacrrArgs& a, int F, float* feat, int qterm, int col0, int warp, int lane) {
  constexpr int ROWS = QT / (NT / 32);
  if (N <= 3) {
    ngram_rows<N, FT, KM, ROWS>(sim, a, F, feat, qterm, col0, warp * ROWS, lane);
  } else {
    ngram_rows<N, FT, KM, ROWS / 2>(sim, a, F, feat, qterm, col0, warp * ROWS, lane);
    ngram_rows<N, FT, KM, ROWS / 2>(sim, a, F, feat, qterm, col0, warp * ROWS + ROWS / 2, lane);
  }
}

template <int FT, int KM>
__device__ __forceinline__ void ngram_dispatch_n(int n, const float* s, const PacrrArgs& a, float* feat, int qterm, int col0, int warp, int lane) {
  switch (n) {
    case 1: ngram_pass<1, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane); break;
    case 2: ngram_pass<2, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane); break;
    case 3: ngram_pass<3, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane); break;
    case 4: ngram_pass<4, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane); break;
    default: ngram_pass<5, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane); break;
  }
}

template <int FT>
__device__ __forceinline__ void ngram_dispatch(int n, const float* s, const PacrrArgs& a, float* feat, int qterm, int col0, int warp, int lane) {
  if (a.kmax <= 2) ngram_dispatch_n<FT, 2>(n, s, a, feat, qterm, col0, warp, lane);  // the reference default (kmax = 2)
  else ngram_dispatch_n<FT, MAX_KMAX>(n, s, a, feat, qterm, col0, warp, lane);
}

// Everything after the cosine tile: n-gram conv/max/top-k passes, softmax(idf) channel, 3-layer combine MLP.
// Runs on 8 warps (tid 0..255); `sync` is the barrier of exactly those warps.
template <class Sync>
__device__ __forceinline__ void pacrr_epilogue(const float* sim, const PacrrArgs& a, int pair, float* feat, float* h1, float* h2, int tid,
                                               Sync sync) {
  const int lane = tid & 31, warp = tid >> 5;
  const int ngrams = a.maxgram - a.mingram + 1;
  const int qterm = ngrams * a.kmax + (a.idf ? 1 : 0);
  for (int g = 0; g < ngrams; ++g) {
    const int n = a.mingram + g;
    if (a.F == 32) ngram_dispatch<32>(n, sim, a, feat, qterm, g * a.kmax, warp, lane);
    else ngram_dispatch<0>(n, sim, a, feat, qterm, g * a.kmax, warp, lane);
  }
  if (a.idf && warp == 0) {
    // softmax over the query axis of the raw idf vector, pads included (PACRR.py:47-50)
    const float v = lane < a.Q ? a.idf[(size_t)pair * a.Q + lane] : -INFINITY;
    const float m = warp_max(v);
    const float e = lane < a.Q ? expf(v - m) : 0.f;
    const float den = warp_sum(e);
    if (lane < a.Q) feat[lane * qterm + qterm - 1] = e / den;
  }
  sync();
  if (a.topk_out) {
    const int tk = ngrams * a.kmax;
    for (int i = tid; i < a.Q * tk; i += 256) {
      int r = i / tk, c = i - r * tk;
      a.topk_out[((size_t)pair * a.Q + r) * tk + c] = feat[r * qterm + c];
    }
  }
  // combine: Linear(Q*qterm, C) -> act -> Linear(C, C) -> act -> Linear(C, 1)   (PACRR.py:29-40,53)
  const int in1 = a.Q * qterm;
  for (int o = warp; o < a.combine; o += 8) {
    const float* w = a.l1w + (size_t)o * in1;
    float p = 0.f;
    for (int i = lane; i < in1; i += 32) p = fmaf(w[i], feat[i], p);  // feat is [Q][qterm] contiguous == flattened
    p = warp_sum(p);
    if (lane == 0) h1[o] = act(p + a.l1b[o], a.nonlin);
  }
  sync();
  for (int o = warp; o < a.combine; o += 8) {
    const float* w = a.l2w + (size_t)o * a.combine;
    float p = 0.f;
    for (int i = lane; i < a.combine; i += 32) p = fmaf(w[i], h1[i], p);
    p = warp_sum(p);
    if (lane == 0) h2[o] = act(p + a.l2b[o], a.nonlin);
  }
  sync();
  if (warp == 0) {
    float p = 0.f;
    for (int i = lane; i < a.combine; i += 32) p = fmaf(a.l3w[i], h2[i], p);
    p = warp_sum(p);
    if (lane == 0) a.scores[pair] = p + a.l3b[0];
  }
  sync();
}

__global__ void __launch_bounds__(NT, 1) pacrr_kernel(const PacrrArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  SimTile s = carve_sim_tile(smem_raw, a.pitch);
  float* feat = reinterpret_cast<float*>(smem_raw + sim_tile_bytes(a.pitch));  // [QT][qterm]
  float* h1 = feat + QT * (MAX_GRAMS * MAX_KMAX + 1);                          // [MAX_COMBINE]
  float* h2 = h1 + MAX_COMBINE;                                                // [MAX_COMBINE]
  clear_sim_tile(s, tid);
  __syncthreads();
  for (int pair = blockIdx.x; pair < a.B; pair += gridDim.x) {
    build_sim_tile(s, a.table, a.pitch, a.V, a.q + (size_t)pair * a.Q, a.Q, a.d + (size_t)pair * a.D, 0, a.D, true, tid);
    pacrr_epilogue(s.sim, a, pair, feat, h1, h2, tid, BlockSync());
  }
}

// Engine 2: cosine tile from the tcgen05 producer (simtc.cuh); the conv/top-k/MLP epilogue is unchanged.
__global__ void __launch_bounds__(simtc::THREADS, 1) pacrr_tc_kernel(const PacrrArgs a) {
  using namespace simtc;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem s = carve(smem_raw, (a.pr.pitch + ATOM_K - 1) / ATOM_K);
  float* feat = s.extra;  // [QT][qterm]
  float* h1 = feat + QT * ((a.maxgram - a.mingram + 1) * a.kmax + (a.idf ? 1 : 0));
  float* h2 = h1 + MAX_COMBINE;
  const uint32_t tmem_base = setup(s, tid);
  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    producer_loop(s, a.pr, tid - EPI_THREADS);
  } else if (warp == EPI_WARPS + PROD_WARPS) {
    mma_loop(s, a.pr, tmem_base);
  } else {
    uint32_t acc_phase = 0;
    int unit = 0;
    for (int pair = blockIdx.x; pair < a.B; pair += gridDim.x, unit += halves_of(a.pr)) {
      drain_pair(s, a.pr, tmem_base, pair, unit, acc_phase, tid);
      pacrr_epilogue(s.sim, a, pair, feat, h1, h2, tid, EpiSync());
    }
  }
  teardown(s, tmem_base, tid);
}

}  // namespace capr

using namespace capr;

static int pacrr_run(const char* fn, bool tc_engine, const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                     const void* table, const void* table_lo, int V, int E, int pitch, int mingram, int maxgram, int nfilters, int kmax,
                     const float* const* conv_w, const float* const* conv_b, const float* l1w, const float* l1b, const float* l2w,
                     const float* l2b, const float* l3w, const float* l3b, int combine, int nonlin, float* scores, float* topk_out,
                     capr_stream_t stream) {
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d", fn, B, Q, D, V);
  CAPR_REQUIRE(mingram >= 1 && maxgram >= mingram && nfilters > 0 && kmax > 0 && combine > 0, CAPR_ERR_BAD_SHAPE, "%s: bad config mingram=%d maxgram=%d nfilters=%d kmax=%d combine=%d", fn, mingram, maxgram, nfilters, kmax, combine);
  CAPR_REQUIRE(nonlin >= 0 && nonlin <= 2, CAPR_ERR_BAD_SHAPE, "%s: nonlinearity must be none, relu or tanh", fn);
  CAPR_REQUIRE(pitch > 0 && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: table pitch %d must be a positive multiple of 16", fn, pitch);
  CAPR_REQUIRE(B == 0 || (query && doc && table && (!tc_engine || table_lo) && conv_w && conv_b && l1w && l1b && l2w && l2b && l3w && l3b && scores), CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(((uintptr_t)table & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: table must be 16-byte aligned", fn);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(D <= DT, CAPR_ERR_UNSUPPORTED, "%s: maxdoclen=%d > %d is not supported by the PACRR kernel yet", fn, D, DT);
  CAPR_REQUIRE(D >= kmax, CAPR_ERR_BAD_SHAPE, "%s: kmax=%d exceeds maxdoclen=%d", fn, kmax, D);
  CAPR_REQUIRE(pitch <= (tc_engine ? simtc::MAX_ATOMS * simtc::ATOM_K : MAX_PITCH), CAPR_ERR_UNSUPPORTED, "%s: embedding dim > %d is not supported by this engine", fn, tc_engine ? simtc::MAX_ATOMS * simtc::ATOM_K : MAX_PITCH);
  CAPR_REQUIRE(maxgram <= MAX_NGRAM && maxgram - mingram + 1 <= MAX_GRAMS, CAPR_ERR_UNSUPPORTED, "%s: maxgram=%d > %d is not supported", fn, maxgram, MAX_NGRAM);
  CAPR_REQUIRE(nfilters <= MAX_FILTERS && kmax <= MAX_KMAX && combine <= MAX_COMBINE, CAPR_ERR_UNSUPPORTED, "%s: nfilters<=%d, kmax<=%d, combine<=%d", fn, MAX_FILTERS, MAX_KMAX, MAX_COMBINE);
  if (B == 0) return CAPR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PacrrArgs a{};
  a.q = (const long long*)query; a.d = (const long long*)doc; a.idf = idf;
  a.B = B; a.Q = Q; a.D = D; a.V = V; a.pitch = pitch; a.mingram = mingram; a.maxgram = maxgram; a.F = nfilters;
  a.kmax = kmax; a.combine = combine; a.nonlin = nonlin; a.table = tc_engine ? nullptr : (const float*)table;
  if (tc_engine) a.pr = simtc::Problem{(const long long*)query, (const long long*)doc, B, Q, D, V, (const __nv_bfloat16*)table, (const __nv_bfloat16*)table_lo, pitch, E, 0};
  a.l1w = l1w; a.l1b = l1b; a.l2w = l2w; a.l2b = l2b; a.l3w = l3w; a.l3b = l3b; a.scores = scores; a.topk_out = topk_out;
  // Stage the filter taps in constant memory (stream-ordered device-to-device copies; the constant bank
  // is per device, so concurrent PACRR calls with different weights must share one stream).
  for (int g = 0; g <= maxgram - mingram; ++g) {
    const int n = mingram + g;
    CAPR_REQUIRE(conv_w[g] && conv_b[g], CAPR_ERR_BAD_POINTER, "%s: null conv weight %d", fn, g);
    a.w_off[g] = conv_w_slot(n);
    CAPR_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_conv_w, conv_w[g], sizeof(float) * nfilters * n * n, sizeof(float) * conv_w_slot(n), cudaMemcpyDeviceToDevice, st));
    CAPR_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_conv_b, conv_b[g], sizeof(float) * nfilters, sizeof(float) * conv_b_slot(n), cudaMemcpyDeviceToDevice, st));
  }
  const size_t extra = (size_t)(QT * (MAX_GRAMS * MAX_KMAX + 1) + 2 * MAX_COMBINE) * sizeof(float);
  const size_t extra_tc = (size_t)(QT * ((maxgram - mingram + 1) * kmax + (idf ? 1 : 0)) + 2 * MAX_COMBINE) * sizeof(float);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  if (tc_engine) {
    const size_t smem = simtc::smem_bytes((pitch + simtc::ATOM_K - 1) / simtc::ATOM_K, extra_tc);
    CAPR_REQUIRE(smem <= simtc::MAX_DYN_SMEM, CAPR_ERR_UNSUPPORTED, "%s: ngrams*kmax too large for the tensor-core engine's shared memory: use capr_pacrr_forward", fn);
    CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table too large for 32-bit row offsets", fn);
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(pacrr_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pacrr_tc_kernel<<<B < sms ? B : sms, simtc::THREADS, smem, st>>>(a);
  } else {
    const size_t smem = sim_tile_bytes(pitch) + extra;
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(pacrr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pacrr_kernel<<<B < sms ? B : sms, NT, smem, st>>>(a);
  }
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

extern "C" int capr_pacrr_forward(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                                  const float* table, int V, int pitch, int mingram, int maxgram, int nfilters, int kmax,
                                  const float* const* conv_w, const float* const* conv_b, const float* l1w,
                                  const float* l1b, const float* l2w, const float* l2b, const float* l3w, const float* l3b,
                                  int combine, int nonlin, float* scores, float* topk_out, capr_stream_t stream) {
  return pacrr_run("capr_pacrr_forward", false, query, doc, idf, B, Q, D, table, nullptr, V, pitch, pitch, mingram, maxgram, nfilters, kmax,
                   conv_w, conv_b, l1w, l1b, l2w, l2b, l3w, l3b, combine, nonlin, scores, topk_out, stream);
}

// Engine 2 (tensor cores): same contract; table given as bf16 (hi, lo) planes (capr_table_prepare_bf16).
extern "C" int capr_pacrr_forward_tc(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                                     const void* table_hi, const void* table_lo, int V, int E, int pitch, int mingram, int maxgram,
                                     int nfilters, int kmax, const float* const* conv_w, const float* const* conv_b, const float* l1w,
                                     const float* l1b, const float* l2w, const float* l2b, const float* l3w, const float* l3b,
                                     int combine, int nonlin, float* scores, float* topk_out, capr_stream_t stream) {
  return pacrr_run("capr_pacrr_forward_tc", true, query, doc, idf, B, Q, D, table_hi, table_lo, V, E, pitch, mingram, maxgram, nfilters,
                   kmax, conv_w, conv_b, l1w, l1b, l2w, l2b, l3w, l3b, combine, nonlin, scores, topk_out, stream);
}
