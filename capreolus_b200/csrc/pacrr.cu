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
#include <mutex>

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
  const float* conv_w[MAX_GRAMS];  // device pointers (conv-on-tensor-cores epilogue builds its filter tile from them)
  const float* conv_b[MAX_GRAMS];
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
// ncols: conv positions of this doc tile that are eligible for the top-k (the whole tile, or -- doc tiling, D > 512 -- the positions
// whose windows lie inside the tile); merge: fold the tile's top-k into the lists already in `feat` (second and later tiles).
template <int N, int FT, int KM, int R>
__device__ __forceinline__ void ngram_rows(const float* sim, const PacrrArgs& a, int F, float* feat, int qterm, int col0, int q0, int lane, int ncols,
                                           bool merge) {
  constexpr int w_off = conv_w_slot(N), b_off = conv_b_slot(N);
  if (q0 >= a.Q) return;  // warp-uniform
  float top[R][KM];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < KM; ++k) top[r][k] = -INFINITY;
  for (int c = lane; c < ncols; c += 64) {
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
    const bool second = c + 32 < ncols;
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
    float* out = feat + (q0 + r) * qterm + col0;
    float old[KM];  // descending top-k of the earlier doc tiles
#pragma unroll
    for (int k = 0; k < KM; ++k) old[k] = (merge && k < a.kmax) ? out[k] : -INFINITY;
    int taken = 0;  // how many of `old` have been emitted
    for (int k = 0; k < a.kmax; ++k) {
      const float m = warp_max(top[r][0]);  // largest value of this tile not emitted yet
      float o = -INFINITY;
#pragma unroll
      for (int j = 0; j < KM; ++j) o = j == taken ? old[j] : o;
      if (o >= m) {  // warp-uniform: the earlier tiles' next value wins; the tile's candidate stays for the next round
        ++taken;
        if (lane == 0) out[k] = o;
        continue;
      }
      const unsigned owners = __ballot_sync(0xffffffffu, top[r][0] == m);
      if (lane == (__ffs(owners) - 1)) {
#pragma unroll
        for (int j = 0; j < KM - 1; ++j) top[r][j] = top[r][j + 1];
        top[r][KM - 1] = -INFINITY;
      }
      if (lane == 0) out[k] = m;
    }
  }
}

// 4 rows at a time for the windows the reference uses (n <= 3); 2 + 2 for the larger ones (register budget)
template <int N, int FT, int KM>
__device__ __forceinline__ void ngram_pass(const float* sim, const PacrrArgs& a, int F, float* feat, int qterm, int col0, int warp, int lane, int ncols,
                                           bool merge) {
  constexpr int ROWS = QT / (NT / 32);
  if (N <= 3) {
    ngram_rows<N, FT, KM, ROWS>(sim, a, F, feat, qterm, col0, warp * ROWS, lane, ncols, merge);
  } else {
    ngram_rows<N, FT, KM, ROWS / 2>(sim, a, F, feat, qterm, col0, warp * ROWS, lane, ncols, merge);
    ngram_rows<N, FT, KM, ROWS / 2>(sim, a, F, feat, qterm, col0, warp * ROWS + ROWS / 2, lane, ncols, merge);
  }
}

template <int FT, int KM>
__device__ __forceinline__ void ngram_dispatch_n(int n, const float* s, const PacrrArgs& a, float* feat, int qterm, int col0, int warp, int lane, int ncols,
                                                 bool merge) {
  switch (n) {
    case 1: ngram_pass<1, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane, ncols, merge); break;
    case 2: ngram_pass<2, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane, ncols, merge); break;
    case 3: ngram_pass<3, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane, ncols, merge); break;
    case 4: ngram_pass<4, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane, ncols, merge); break;
    default: ngram_pass<5, FT, KM>(s, a, a.F, feat, qterm, col0, warp, lane, ncols, merge); break;
  }
}

template <int FT>
__device__ __forceinline__ void ngram_dispatch(int n, const float* s, const PacrrArgs& a, float* feat, int qterm, int col0, int warp, int lane, int ncols,
                                               bool merge) {
  if (a.kmax <= 2) ngram_dispatch_n<FT, 2>(n, s, a, feat, qterm, col0, warp, lane, ncols, merge);  // the reference default (kmax = 2)
  else ngram_dispatch_n<FT, MAX_KMAX>(n, s, a, feat, qterm, col0, warp, lane, ncols, merge);
}

template <class Sync>
__device__ __forceinline__ void pacrr_tail(const PacrrArgs& a, int pair, float* feat, float* h1, float* h2, int tid, Sync sync);

// Everything after the cosine tile: n-gram conv/max/top-k passes, softmax(idf) channel, 3-layer combine MLP.
// Runs on 8 warps (tid 0..255); `sync` is the barrier of exactly those warps.
// The n-gram conv / max / top-k passes over one doc tile of `ncols` eligible positions (8 warps, tid 0..255).
__device__ __forceinline__ void pacrr_convs(const float* sim, const PacrrArgs& a, float* feat, int tid, int ncols, bool merge) {
  const int lane = tid & 31, warp = tid >> 5;
  const int ngrams = a.maxgram - a.mingram + 1;
  const int qterm = ngrams * a.kmax + (a.idf ? 1 : 0);
  for (int g = 0; g < ngrams; ++g) {
    const int n = a.mingram + g;
    if (a.F == 32) ngram_dispatch<32>(n, sim, a, feat, qterm, g * a.kmax, warp, lane, ncols, merge);
    else ngram_dispatch<0>(n, sim, a, feat, qterm, g * a.kmax, warp, lane, ncols, merge);
  }
}

template <class Sync>
__device__ __forceinline__ void pacrr_epilogue(const float* sim, const PacrrArgs& a, int pair, float* feat, float* h1, float* h2, int tid,
                                               Sync sync) {
  pacrr_convs(sim, a, feat, tid, a.D, false);
  pacrr_tail(a, pair, feat, h1, h2, tid, sync);
}

// softmax(idf) channel + 3-layer combine MLP on the [Q][qterm] feature block (8 warps, tid 0..255).
template <class Sync>
__device__ __forceinline__ void pacrr_tail(const PacrrArgs& a, int pair, float* feat, float* h1, float* h2, int tid, Sync sync) {
  const int lane = tid & 31, warp = tid >> 5;
  const int ngrams = a.maxgram - a.mingram + 1;
  const int qterm = ngrams * a.kmax + (a.idf ? 1 : 0);
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
  // Doc tiling (maxdoclen > 512; the reference extractor's default is 800): tiles of 512 columns start every S = 512 - (maxgram - 1)
  // columns, so that every conv window that starts in [0, S) of a tile lies inside it; only those positions are eligible for the
  // tile's top-k (the last tile: all of its columns -- there the zero padding past the document is the reference's own,
  // PACRR.py:64), and the per-row top-k lists are carried from tile to tile in `feat`.
  const int S = DT - (a.maxgram - 1);
  const int tiles = a.D <= DT ? 1 : 1 + (a.D - DT + S - 1) / S;
  for (int pair = blockIdx.x; pair < a.B; pair += gridDim.x) {
    for (int t = 0; t < tiles; ++t) {
      const int d0 = t * S;
      build_sim_tile(s, a.table, a.pitch, a.V, a.q + (size_t)pair * a.Q, a.Q, a.d + (size_t)pair * a.D, d0, a.D, t == 0, tid);
      pacrr_convs(s.sim, a, feat, tid, t + 1 < tiles ? S : a.D - d0, t > 0);
      __syncthreads();  // the tile is rebuilt / `feat` is read by the tail
    }
    pacrr_tail(a, pair, feat, h1, h2, tid, BlockSync());
  }
}

// Engine 2: cosine tile from the tcgen05 producer (simtc.cuh); the conv/top-k/MLP epilogue is unchanged.
// IDLE_SLEEP: the gather and MMA warps are idle most of the time here (the convolutions bound the kernel) and sleep between polls
// of their barriers instead of spinning (tc::mbar_wait_idle).
template <bool IDLE_SLEEP>
__global__ void __launch_bounds__(simtc::THREADS, 1) pacrr_tc_kernel(const PacrrArgs a) {
  using namespace simtc;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem s = carve(smem_raw, (a.pr.pitch + ATOM_K - 1) / ATOM_K);
  float* feat = s.extra;  // [QT][qterm]
  float* h1 = feat + QT * ((a.maxgram - a.mingram + 1) * a.kmax + (a.idf ? 1 : 0));
  float* h2 = h1 + MAX_COMBINE;
  const uint32_t tmem_base = setup(s, tid, THREADS, MMA_WARP);
  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    producer_loop<IDLE_SLEEP>(s, a.pr, tid - EPI_THREADS);
  } else if (warp == EPI_WARPS + PROD_WARPS) {
    mma_loop<IDLE_SLEEP>(s, a.pr, tmem_base);
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

// ---- Engine 3: the convolutions on tensor cores too --------------------------------------------------------------------
// relu(max_f(b_f + sum_taps w_f[tap] * S[i+u][j+v])) for the 1x1, 2x2 and 3x3 windows is ONE skinny GEMM per block of 128
// cells: A = im2col rows of the 3x3 window (the smaller windows are its top-left corners: zero weights elsewhere), B = the
// 3 x 32 filters, bias folded in as two constant-one K columns.  fp32 accuracy comes from the same split as everywhere
// else: every cosine S and every weight W is (hi, lo) bf16 and the K axis carries S_hi*W_hi + S_lo*W_hi + S_hi*W_lo per
// tap -> K = 9 taps x 3 + 2 bias = 29 <= 32, i.e. two tcgen05.mma (M=128, N=32*ngrams, K=16) per 128 cells instead of
// 448 FMAs per cell.  The 8 epilogue warps build the A tile (no-swizzle K-major core-matrix layout, 16-byte stores),
// one elected thread issues the MMAs, and while they run the same warps reduce the PREVIOUS block: tcgen05.ld of its
// [128 x 32*ngrams] accumulator, max over the 32 filters of each window size, ReLU, lane-local running top-k of the row;
// at the end of each query row the 256 lane-local lists are merged through shared memory.
// STATUS: parity-green, opt-in (CAPR_PACRR_CONV=tc3), NOT the default: measured 1.1 M pairs/s against 2.5 M for the FFMA2
// epilogue -- see the note in pacrr_run and DESIGN.md.
// Resources: the cosine-tile producer runs in `single` mode (one query buffer, one TMEM accumulator buffer), which
// frees shared memory for two 8 KB A stages + the 6 KB filter tile + the merge scratch, and TMEM columns 256-511 for
// two conv accumulators.  Used when nfilters == 32, maxgram <= 3, kmax <= 4 (the reference defaults); other configs
// keep the FFMA2 epilogue above.
constexpr int C3_A_STAGE = 128 * 64;            // 128 cells x 32 bf16
constexpr int C3_B_BYTES = 96 * 64;             // up to 96 filter columns x 32 bf16
constexpr int C3_KM = 4;                        // kmax limit of this path
constexpr int C3_SCRATCH = 3 * 256 * C3_KM * 4; // merge scratch: [ngram][thread][k]
constexpr int C3_BYTES = 2 * C3_A_STAGE + C3_B_BYTES + C3_SCRATCH;  // 34 816 B

__device__ __forceinline__ uint32_t c3_off(int row, int chunk) { return (uint32_t)((row >> 3) * 512 + chunk * 128 + (row & 7) * 16); }

__device__ __forceinline__ void c3_split(float x, uint32_t& h, uint32_t& l) {
  const __nv_bfloat16 hb = __float2bfloat16_rn(x);
  h = (uint32_t)__bfloat16_as_ushort(hb);
  l = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(hb)));
}

template <int KM>
__device__ __forceinline__ void c3_insert(float (&top)[KM], float v) {
#pragma unroll
  for (int k = 0; k < KM; ++k) {
    if (v > top[k]) {
      const float t = top[k];
      top[k] = v;
      v = t;
    }
  }
}

template <int KM>
__global__ void __launch_bounds__(simtc::THREADS, 1) pacrr_tc3_kernel(const PacrrArgs a) {
  using namespace simtc;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int atoms = (a.pr.pitch + ATOM_K - 1) / ATOM_K;
  Smem s = carve(smem_raw, atoms);
  const int ngrams = a.maxgram - a.mingram + 1;
  const int qterm = ngrams * a.kmax + (a.idf ? 1 : 0);
  float* feat = s.extra;  // [QT][qterm]
  float* h1 = feat + QT * qterm;
  float* h2 = h1 + MAX_COMBINE;
  uint64_t* conv_full = reinterpret_cast<uint64_t*>(h2 + MAX_COMBINE);  // [2]
  // conv buffers: the unused second query buffer when it is large enough, else behind the barriers in the extra region
  unsigned char* cbuf = atoms * Q_ATOM_BYTES >= C3_BYTES ? s.qbuf(1) : reinterpret_cast<unsigned char*>(conv_full + 2);
  cbuf += (128u - (tc::smem_u32(cbuf) & 127u)) & 127u;
  unsigned char* a_stage = cbuf;                         // 2 x 8 KB
  unsigned char* b_tile = cbuf + 2 * C3_A_STAGE;          // 6 KB
  float* scratch = reinterpret_cast<float*>(b_tile + C3_B_BYTES);  // [ngram][256][KM]
  if (tid == 0) {
    tc::mbar_init(&conv_full[0], 1);
    tc::mbar_init(&conv_full[1], 1);
  }
  const uint32_t tmem_base = setup(s, tid, THREADS, MMA_WARP);  // fence.mbarrier_init + __syncthreads inside
  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    producer_loop(s, a.pr, tid - EPI_THREADS);
  } else if (warp == EPI_WARPS + PROD_WARPS) {
    mma_loop(s, a.pr, tmem_base);
  } else {
    const int Ntot = ngrams * 32;
    // ---- filter tile, once: column n = g*32 + f; K slots: taps 0-4 -> 3t+{0,1,2} = (W_hi, W_hi, W_lo), slot 15 = 0;
    //      taps 5-8 -> 16 + 3(t-5) + {0,1,2}; slots 28, 29 = (b_hi, b_lo); 30, 31 = 0
    for (int idx = tid; idx < Ntot * 32; idx += EPI_THREADS) {
      const int n = idx >> 5, slot = idx & 31;
      const int g = n >> 5, f = n & 31, win = a.mingram + g;
      float val = 0.f;
      int part = -1;  // 0: hi, 1: lo
      if (slot < 15 || (slot >= 16 && slot < 28)) {
        const int t = slot < 15 ? slot / 3 : 5 + (slot - 16) / 3;
        const int r = slot < 15 ? slot % 3 : (slot - 16) % 3;
        const int u = t / 3, v = t % 3;
        if (u < win && v < win) {
          val = a.conv_w[g][(size_t)f * win * win + u * win + v];
          part = r == 2 ? 1 : 0;
        }
      } else if (slot == 28 || slot == 29) {
        val = a.conv_b[g][f];
        part = slot - 28;
      }
      uint32_t h = 0, l = 0;
      if (part >= 0) c3_split(val, h, l);
      *reinterpret_cast<unsigned short*>(b_tile + c3_off(n, slot >> 3) + (slot & 7) * 2) = (unsigned short)(part == 1 ? l : h);
    }
    tc::fence_proxy_async();
    epi_barrier();
    const uint32_t idesc = tc::make_instr_desc(tc::FMT_BF16, 128, Ntot);
    const uint32_t a_addr = tc::smem_u32(a_stage), b_addr = tc::smem_u32(b_tile);
    const int grp = warp >> 2, quarter = warp & 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int NBJ = (a.D + 127) / 128;
    const int NB = a.Q * NBJ;
    uint32_t acc_phase = 0, conv_phase = 0;  // one bit per buffer
    int unit = 0;
    const int dbg = CAPR_DBG(a.pr.debug);  // profiling only (CAPR_PACRR_DEBUG): 0x10000 no build, 0x20000 no MMA, 0x40000 no reduce, 0x80000 no fence
    for (int pair = blockIdx.x; pair < a.B; pair += gridDim.x, unit += halves_of(a.pr)) {
      drain_pair(s, a.pr, tmem_base, pair, unit, acc_phase, tid);  // -> s.sim (fp32, zero halo); ends with epi_barrier
      float top[3][KM];
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int k = 0; k < KM; ++k) top[g][k] = -INFINITY;
      for (int b = 0; b <= ((dbg & 0x200000) ? -1 : NB); ++b) {
        if (b < NB && !(dbg & 0x10000)) {
          // ---- build the A tile of block b: cell = tid & 127, this thread writes K chunks 2*half, 2*half+1
          const int i = b / NBJ, jb = b - i * NBJ;
          const int cell = tid & 127, half = tid >> 7;
          const float* p = s.sim + i * SIM_PITCH + jb * 128 + cell;
          uint32_t w[8];
          if (half == 0) {
            uint32_t h0, l0, h1_, l1, h2_, l2, h3, l3, h4, l4;
            c3_split(p[0], h0, l0);                      // tap (0,0)
            c3_split(p[1], h1_, l1);                     // (0,1)
            c3_split(p[2], h2_, l2);                     // (0,2)
            c3_split(p[SIM_PITCH], h3, l3);              // (1,0)
            c3_split(p[SIM_PITCH + 1], h4, l4);          // (1,1)
            w[0] = h0 | (l0 << 16), w[1] = h0 | (h1_ << 16), w[2] = l1 | (h1_ << 16), w[3] = h2_ | (l2 << 16);
            w[4] = h2_ | (h3 << 16), w[5] = l3 | (h3 << 16), w[6] = h4 | (l4 << 16), w[7] = h4;
          } else {
            uint32_t h5, l5, h6, l6, h7, l7, h8, l8;
            c3_split(p[SIM_PITCH + 2], h5, l5);          // (1,2)
            c3_split(p[2 * SIM_PITCH], h6, l6);          // (2,0)
            c3_split(p[2 * SIM_PITCH + 1], h7, l7);      // (2,1)
            c3_split(p[2 * SIM_PITCH + 2], h8, l8);      // (2,2)
            w[0] = h5 | (l5 << 16), w[1] = h5 | (h6 << 16), w[2] = l6 | (h6 << 16), w[3] = h7 | (l7 << 16);
            w[4] = h7 | (h8 << 16), w[5] = l8 | (h8 << 16), w[6] = 0x3F803F80u, w[7] = 0u;  // bias columns: 1.0, 1.0
          }
          unsigned char* st = a_stage + (b & 1) * C3_A_STAGE;
          *reinterpret_cast<uint4*>(st + c3_off(cell, 2 * half)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(st + c3_off(cell, 2 * half + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
          if (!(dbg & 0x80000)) tc::fence_proxy_async();
        }
        if (!(dbg & 0x400000)) tc::tc_fence_before();  // the previous iteration's tcgen05.ld of the accumulator this block's MMAs overwrite
        if (!(dbg & 0x800000)) epi_barrier();
        if (b < NB && warp == 0 && !(dbg & 0x20000)) {
          tc::tc_fence_after();
          if (tc::elect_one()) {
            const uint32_t st = a_addr + (uint32_t)((b & 1) * C3_A_STAGE);
            const uint32_t d_tmem = tmem_base + (uint32_t)(256 + (b & 1) * 128);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::umma_f16(d_tmem, tc::make_nosw_kmajor_desc(st + ks * 256, 128, 512), tc::make_nosw_kmajor_desc(b_addr + ks * 256, 128, 512),
                           idesc, ks != 0);
            tc::umma_commit(&conv_full[b & 1]);
          }
          __syncwarp();
        }
        if (b >= 1) {
          // ---- reduce block b-1: max over the 32 filters of each window size, ReLU, running top-k of the row
          const int pb = b - 1, buf = pb & 1;
          const int i = pb / NBJ, jb = pb - i * NBJ;
          if (!(dbg & 0x20000)) tc::mbar_wait(&conv_full[buf], (conv_phase >> buf) & 1);
          conv_phase ^= 1u << buf;
          if (!(dbg & 0x1000000)) tc::tc_fence_after();
          const int j = jb * 128 + quarter * 32 + lane;
          // which window sizes this warp group takes for this block (balanced 64 + 32 columns, alternating)
          int g_lo, g_hi;
          if (ngrams == 3) {
            if (((pb ^ grp) & 1) == 0) g_lo = 0, g_hi = 2; else g_lo = 2, g_hi = 3;
          } else {
            g_lo = grp, g_hi = grp < ngrams ? grp + 1 : grp;
          }
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            if (g >= g_lo && g < g_hi && !(dbg & 0x40000)) {  // warp-uniform
              float v[32];
              tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(256 + buf * 128 + g * 32), v);
              tc::tmem_ld_wait();
              float m = v[0];
#pragma unroll
              for (int f = 1; f < 32; ++f) m = fmaxf(m, v[f]);
              m = j < a.D ? fmaxf(m, 0.f) : -INFINITY;  // ReLU commutes with the filter max; columns past the doc do not exist
              c3_insert<KM>(top[g], m);
            }
          }
          if (jb == NBJ - 1 && !(dbg & 0x100000)) {
            // ---- end of query row i: merge the 256 lane-local lists of every window size
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
              for (int k = 0; k < KM; ++k) {
                if (g < ngrams) scratch[(g * 256 + tid) * KM + k] = top[g][k];
                top[g][k] = -INFINITY;
              }
            epi_barrier();
            if (warp < ngrams) {
              float c[8 * KM];
#pragma unroll
              for (int w = 0; w < 8; ++w)
#pragma unroll
                for (int k = 0; k < KM; ++k) c[w * KM + k] = scratch[(warp * 256 + w * 32 + lane) * KM + k];
              for (int k = 0; k < a.kmax; ++k) {
                float m = c[0];
#pragma unroll
                for (int e = 1; e < 8 * KM; ++e) m = fmaxf(m, c[e]);
                const float M = warp_max(m);
                const unsigned who = __ballot_sync(0xffffffffu, m == M);
                if (lane == __ffs(who) - 1) {  // remove ONE instance of the maximum
                  bool done = false;
#pragma unroll
                  for (int e = 0; e < 8 * KM; ++e) {
                    const bool hit = !done && c[e] == M;
                    c[e] = hit ? -INFINITY : c[e];
                    done = done || hit;
                  }
                }
                if (lane == 0) feat[i * qterm + warp * a.kmax + k] = M;
              }
            }
            epi_barrier();  // scratch is rewritten at the end of the next row; feat row i is complete
          }
        }
      }
      pacrr_tail(a, pair, feat, h1, h2, tid, EpiSync());
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
  capr::DeviceGuard device_guard(table);  // act on the device that owns the caller's buffers
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d", fn, B, Q, D, V);
  CAPR_REQUIRE(mingram >= 1 && maxgram >= mingram && nfilters > 0 && kmax > 0 && combine > 0, CAPR_ERR_BAD_SHAPE, "%s: bad config mingram=%d maxgram=%d nfilters=%d kmax=%d combine=%d", fn, mingram, maxgram, nfilters, kmax, combine);
  CAPR_REQUIRE(nonlin >= 0 && nonlin <= 2, CAPR_ERR_BAD_SHAPE, "%s: nonlinearity must be none, relu or tanh", fn);
  CAPR_REQUIRE(pitch > 0 && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: table pitch %d must be a positive multiple of 16", fn, pitch);
  CAPR_REQUIRE(B == 0 || (query && doc && table && (!tc_engine || table_lo) && conv_w && conv_b && l1w && l1b && l2w && l2b && l3w && l3b && scores), CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(((uintptr_t)table & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: table must be 16-byte aligned", fn);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(D <= (tc_engine ? DT : 8 * DT), CAPR_ERR_UNSUPPORTED, "%s: maxdoclen=%d > %d is not supported by this engine%s", fn, D, tc_engine ? DT : 8 * DT,
               tc_engine ? " (capr_pacrr_forward tiles the document)" : "");
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
  if (tc_engine) {
    a.pr = simtc::Problem{(const long long*)query, (const long long*)doc, B, Q, D, V, (const __nv_bfloat16*)table, (const __nv_bfloat16*)table_lo, pitch, E, 0};
  }
  a.l1w = l1w; a.l1b = l1b; a.l2w = l2w; a.l2b = l2b; a.l3w = l3w; a.l3b = l3b; a.scores = scores; a.topk_out = topk_out;
  // Stage the filter taps in constant memory (stream-ordered device-to-device copies).  The constant bank is one per device, so
  // calls are serialised per device: a call on another stream (or thread) first waits for the event recorded behind the previous
  // call's kernel before it overwrites the taps, and the host-side section between the copies and the launch holds a mutex.
  static std::mutex conv_bank_mutex;
  static cudaEvent_t conv_bank_free[64] = {};
  std::lock_guard<std::mutex> conv_bank_lock(conv_bank_mutex);
  int cur_dev = 0;
  CAPR_CHECK_CUDA(cudaGetDevice(&cur_dev));
  CAPR_REQUIRE(cur_dev >= 0 && cur_dev < 64, CAPR_ERR_UNSUPPORTED, "%s: device index %d out of range", fn, cur_dev);
  if (conv_bank_free[cur_dev]) CAPR_CHECK_CUDA(cudaStreamWaitEvent(st, conv_bank_free[cur_dev], 0));
  else CAPR_CHECK_CUDA(cudaEventCreateWithFlags(&conv_bank_free[cur_dev], cudaEventDisableTiming));
  for (int g = 0; g <= maxgram - mingram; ++g) {
    const int n = mingram + g;
    CAPR_REQUIRE(conv_w[g] && conv_b[g], CAPR_ERR_BAD_POINTER, "%s: null conv weight %d", fn, g);
    a.w_off[g] = conv_w_slot(n);
    a.conv_w[g] = conv_w[g], a.conv_b[g] = conv_b[g];
    CAPR_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_conv_w, conv_w[g], sizeof(float) * nfilters * n * n, sizeof(float) * conv_w_slot(n), cudaMemcpyDeviceToDevice, st));
    CAPR_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_conv_b, conv_b[g], sizeof(float) * nfilters, sizeof(float) * conv_b_slot(n), cudaMemcpyDeviceToDevice, st));
  }
  const size_t extra = (size_t)(QT * (MAX_GRAMS * MAX_KMAX + 1) + 2 * MAX_COMBINE) * sizeof(float);
  const size_t extra_tc = (size_t)(QT * ((maxgram - mingram + 1) * kmax + (idf ? 1 : 0)) + 2 * MAX_COMBINE) * sizeof(float);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  // CAPR_PACRR_CONV=tc3 selects the experimental conv-on-tensor-cores epilogue (engine 3).  It is parity-green but measured
  // SLOWER than the FFMA2 epilogue (1.1 M vs 2.5 M pairs/s): the im2col build + TMEM reduce + top-k cost ~430 dependent
  // thread-instructions per cell on 8 warps (ncu source view), no fewer than the 350 high-ILP FFMA2 / FMNMX of the default path.
  const char* conv_env = getenv("CAPR_PACRR_CONV");
  if (tc_engine && nfilters == 32 && maxgram <= 3 && kmax <= C3_KM && conv_env && conv_env[0] == 't') {
    // engine 3: convolutions on tensor cores as well
    const int atoms = (pitch + simtc::ATOM_K - 1) / simtc::ATOM_K;
    const size_t conv_extra = (size_t)atoms * simtc::Q_ATOM_BYTES >= (size_t)C3_BYTES ? 0 : (size_t)C3_BYTES + 128;
    const size_t smem = simtc::smem_bytes(atoms, extra_tc + 2 * sizeof(uint64_t) + 8 + conv_extra);
    CAPR_REQUIRE(smem <= simtc::MAX_DYN_SMEM, CAPR_ERR_UNSUPPORTED, "%s: shared memory budget exceeded", fn);
    CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table too large for 32-bit row offsets", fn);
    a.pr.single = 1;
#ifdef CAPR_DEBUG_BUILD
    if (const char* de = getenv("CAPR_PACRR_DEBUG")) a.pr.debug = (int)strtol(de, nullptr, 0) & 0x1FF0000;
#endif
    if (kmax <= 2) {
      CAPR_CHECK_CUDA(cudaFuncSetAttribute(pacrr_tc3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pacrr_tc3_kernel<2><<<B < sms ? B : sms, simtc::THREADS, smem, st>>>(a);
    } else {
      CAPR_CHECK_CUDA(cudaFuncSetAttribute(pacrr_tc3_kernel<C3_KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pacrr_tc3_kernel<C3_KM><<<B < sms ? B : sms, simtc::THREADS, smem, st>>>(a);
    }
  } else if (tc_engine) {
    const size_t smem = simtc::smem_bytes((pitch + simtc::ATOM_K - 1) / simtc::ATOM_K, extra_tc);
    CAPR_REQUIRE(smem <= simtc::MAX_DYN_SMEM, CAPR_ERR_UNSUPPORTED, "%s: ngrams*kmax too large for the tensor-core engine's shared memory: use capr_pacrr_forward", fn);
    CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table too large for 32-bit row offsets", fn);
    const char* sleep_env = getenv("CAPR_PACRR_IDLE_SLEEP");  // A/B switch, default on
    if (sleep_env && sleep_env[0] == '0') {
      CAPR_CHECK_CUDA(cudaFuncSetAttribute(pacrr_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pacrr_tc_kernel<false><<<B < sms ? B : sms, simtc::THREADS, smem, st>>>(a);
    } else {
      CAPR_CHECK_CUDA(cudaFuncSetAttribute(pacrr_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pacrr_tc_kernel<true><<<B < sms ? B : sms, simtc::THREADS, smem, st>>>(a);
    }
  } else {
    const size_t smem = sim_tile_bytes(pitch) + extra;
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(pacrr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pacrr_kernel<<<B < sms ? B : sms, NT, smem, st>>>(a);
  }
  CAPR_CHECK_CUDA(cudaGetLastError());
  CAPR_CHECK_CUDA(cudaEventRecord(conv_bank_free[cur_dev], st));
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
