// K2 -- fused DRMM scoring kernel.
//
//   DRMM_class._hist_map     capreolus/reranker/DRMM.py:41-81    matching histogram (CH / NH / LCH)
//   DRMM_class._term_gate    capreolus/reranker/DRMM.py:83-99    softmax term gate (IDF / TV)
//   DRMM_class.forward       capreolus/reranker/DRMM.py:101-116
//   SimilarityMatrix         capreolus/reranker/common.py:143-182 (producer: simtile.cuh)
//
// The reference counts `sim < upper_bound_i` for each of the nbins bounds, differences neighbouring
// counts and copies them into a CPU tensor (30 device syncs per batch).  Here every cosine is binned
// once: b(s) = the smallest i with s < ub[i] (an arithmetic guess fixed up against the exact fp32 bounds
// the host passes in, i.e. torch.linspace(-1,1,nbins+1)[1:]), which is the same partition.  Quirks kept:
// padded doc columns fall in no bin (DRMM.py:59); padded query rows still histogram and are removed only
// by the -1e7 gate mask (DRMM.py:89); the last slot counts 0.999 < s < 1.001 and overlaps bin nbins-1.
// Identical tokens: the fp32 reference puts them in the last regular bin or not depending on the rounding of
// a.a/(|a|+1e-9)^2 (a coin flip per token); we follow the exact-arithmetic value (< 1.0 -> in the bin), which is what
// the reference computes when its cosines are evaluated in fp64 (tests/test_oracle.py pins this).
// Counting is integer work: __match_any_sync groups the lanes of a warp by bin and the group leader adds
// the group's size to a per-row shared-memory counter (no atomics, deterministic).
#include "simtc.cuh"

namespace capr {

constexpr int MAX_SLOTS = 64;     // nbins + 1 (FFMA engine)
constexpr int MAX_SLOTS_TC = 32;  // nbins + 1 (tensor-core engine: shared memory is nearly full)
constexpr int CNT_PITCH_TC = 33;  // counter row stride there: rows land in different banks for the same bin

struct DrmmArgs {
  const long long* q;
  const long long* d;
  const float* idf;
  int B, Q, D, V, pitch, E, nbins, hist_type, gate_type, nodes;
  const float* table;
  const float* raw_emb;
  const float* bin_ub;
  const float *ffw_w1, *ffw_b1, *ffw_w2, *ffw_b2, *gate_w, *out_w, *out_b;
  float* scores;
  float* hist_out;
  simtc::Problem pr;  // tensor-core engine only
  int pool_mode;      // tensor-core engine: DRMM_POOL_* (private byte histograms by default)
};

constexpr int DRMM_POOL_PRIVATE = 0;  // one byte histogram per (pooling warp, query row): plain LDS/STS, no atomics
constexpr int DRMM_POOL_ATOMIC = 1;   // shared 32-bit counters per row, integer atomics from the 8 pooling warps (A/B reference)
constexpr int DRMM_POOL_NOADD = 2;    // profiling only (results invalid): bins computed, nothing counted
constexpr int DRMM_POOL_SKIP = 3;     // profiling only (results invalid): no pooling work at all
constexpr int HIST_PITCH_B = 36;      // bytes per private histogram row: word stride 9 is odd, so equal bins of the 32 rows hit 32 banks
constexpr int HIST_WARP_B = QT * HIST_PITCH_B;  // 1 152 B per pooling warp
constexpr int HIST_WARPS_IN_SPARE = 6;           // 6 912 B behind the half tiles; warps 6, 7 + bounds + z live in the extra region
constexpr size_t DRMM_TC_EXTRA_BYTES = (size_t)(simtc::POOL_WARPS - HIST_WARPS_IN_SPARE) * HIST_WARP_B + (MAX_SLOTS_TC + QT) * sizeof(float);

struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct EpiSync {
  __device__ __forceinline__ void operator()() const { simtc::epi_barrier(); }
};

// Bin the `ncols` cosines of every query row of the tile (8 warps x 4 rows; row stride PITCH floats, at most COLS
// columns).  `did_smem` = the doc ids of these columns.
template <int SLOTS, int PITCH, int COLS>
__device__ __forceinline__ void drmm_count_tile(const float* sim, const int* qid, const int* did_smem, int ncols, const DrmmArgs& a,
                                                const float* ub, int* cnt, int warp, int lane) {
  const float guess_scale = 0.5f * (float)a.nbins;
  // this lane's doc ids (columns lane, lane+32, ...), fetched once for the 4 query rows of the warp
  int dd[COLS / 32];
#pragma unroll
  for (int t = 0; t < COLS / 32; ++t) {
    const int c = lane + 32 * t;
    dd[t] = c < ncols ? did_smem[c] : 0;
  }
  for (int r = 0; r < 4; ++r) {
    const int qrow = warp * 4 + r;
    const float* row = sim + qrow * PITCH;
    int* c_row = cnt + qrow * SLOTS;
    const int qi = qid[qrow];
#pragma unroll
    for (int t = 0; t < COLS / 32; ++t) {
      if (t * 32 >= ncols) break;  // warp-uniform
      const int c = lane + 32 * t;
      const int did = dd[t];
      const bool real = did != 0;  // padded columns are pushed to +1e7: no bin (DRMM.py:59)
      const float v = real ? row[c] : 0.f;
      // bin = the smallest i with v < ub[i].  The bounds are torch.linspace(-1,1,nbins+1)[1:], so the arithmetic guess
      // is off by at most one (only when v sits within rounding of an edge); one comparison on each side against the
      // exact fp32 bounds settles it.  v >= ub[nbins-1] = 1.0 -> nbins = "no bin".
      const int g = max(0, min((int)floorf((v + 1.0f) * guess_scale), a.nbins - 1));
      const float hi = ub[g];
      const float lo = ub[max(g - 1, 0)];
      int b = g + ((v >= hi) ? 1 : 0) - ((g > 0 && v < lo) ? 1 : 0);
      // identical in-vocabulary tokens are stored as exactly 1.0f (simtile.cuh); their exact-arithmetic cosine
      // 1 - 2e-9/|a| is < 1.0, i.e. inside the last regular bin [ub[nbins-2], 1.0) as well as the exact slot
      if (real && v == 1.0f && qi > 0 && qi == did) b = a.nbins - 1;
      if (!real) b = a.nbins + 1;  // sentinel group, never stored
      const unsigned peers = __match_any_sync(0xffffffffu, b);
      if (b < a.nbins && lane == (__ffs(peers) - 1)) c_row[b] += __popc(peers);
      const unsigned exact = __ballot_sync(0xffffffffu, real && v > 0.999f && v < 1.001f);  // DRMM.py:66
      if (lane == 0 && exact) c_row[a.nbins] += __popc(exact);
      __syncwarp();
    }
  }
}

// Tensor-core engine: lane = query row, warp = a 32-column slice of the half tile (the one-row-per-lane float4 reads are
// conflict-free, the doc ids of the slice are warp-uniform broadcasts).  Counting (MODE): by default each (pooling warp,
// row) owns a byte histogram, so the lane is the only writer of its counters and plain loads / stores do; the A/B mode
// adds into one per-row counter array with shared-memory integer atomics (integer adds commute: deterministic either
// way).  Odd word strides put equal bins of different rows in different banks.  The pooling is latency-bound, not
// issue-bound: computing the four bins of a float4 branch-free (overlapping their F2I -> LDS -> compare chains) took the
// kernel from 9.4 M to 10.3 M pairs/s; packing the counters into registers (no shared traffic, +16 ALU ops per cosine) or
// moving the producers to their own sub-partition both lost.
template <int MODE>
__device__ __forceinline__ void drmm_count_slice(const float* tile, int pitch, int col0, int nvalid, int qi, const int* did_slice,
                                                 const DrmmArgs& a, const float* ub, int* cnt_row, unsigned char* hist_row, int& sink) {
  const float guess_scale = 0.5f * (float)a.nbins;
  const float4* row = reinterpret_cast<const float4*>(tile + (threadIdx.x & 31) * pitch + col0);
  const int4* dids = reinterpret_cast<const int4*>(did_slice);
  for (int g = 0; g < 8; ++g) {
    if (g * 4 >= nvalid) break;  // warp-uniform
    const float4 x = row[g];
    const int4 d4 = dids[g];
    const float vv[4] = {x.x, x.y, x.z, x.w};
    const int dd[4] = {d4.x, d4.y, d4.z, d4.w};
    // the four bins are computed branch-free so that their dependent chains (F2I -> bound lookups -> compares) overlap
    int bin[4];
    bool exact[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = vv[j];
      const bool real = g * 4 + j < nvalid && dd[j] != 0;  // padded columns are pushed to +1e7: no bin (DRMM.py:59)
      // same binning as drmm_count_tile: arithmetic guess, one comparison on each side against the exact fp32 bounds
      const int g0 = max(0, min((int)floorf((v + 1.0f) * guess_scale), a.nbins - 1));
      const float hi = ub[g0];
      const float lo = ub[max(g0 - 1, 0)];
      int b = g0 + ((v >= hi) ? 1 : 0) - ((g0 > 0 && v < lo) ? 1 : 0);
      if (v == 1.0f && qi > 0 && qi == dd[j]) b = a.nbins - 1;  // identical in-vocabulary tokens (see drmm_count_tile)
      bin[j] = real ? b : a.nbins;                              // nbins = "no bin" (also v >= 1.0)
      exact[j] = real && v > 0.999f && v < 1.001f;              // DRMM.py:66
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = bin[j];
      if (MODE == DRMM_POOL_PRIVATE) {
        // this lane's own row of this warp's own histogram: at most 32 columns x 4 units = 128 increments per pair (fits a byte)
        if (b < a.nbins) hist_row[b] = (unsigned char)(hist_row[b] + 1);
        if (exact[j]) hist_row[a.nbins] = (unsigned char)(hist_row[a.nbins] + 1);
      } else if (MODE == DRMM_POOL_ATOMIC) {
        if (b < a.nbins) atomicAdd(cnt_row + b, 1);
        if (exact[j]) atomicAdd(cnt_row + a.nbins, 1);
      } else {
        sink += b;
      }
    }
  }
}

// counts -> +1 -> CH/NH/LCH -> 2-layer tanh feed-forward per query row -> softmax term gate -> score
// Counter accessors for drmm_finish: a [row][SLOTS] int array, or the sum of the 8 private byte histograms.
template <int SLOTS>
struct ArrayCounts {
  const int* cnt;
  __device__ __forceinline__ int operator()(int qrow, int slot) const { return cnt[qrow * SLOTS + slot]; }
};
struct PrivateCounts {
  const unsigned char* spare;  // pooling warps 0..HIST_WARPS_IN_SPARE-1
  const unsigned char* extra;  // the rest
  __device__ __forceinline__ const unsigned char* warp_hist(int pw) const {
    return pw < HIST_WARPS_IN_SPARE ? spare + pw * HIST_WARP_B : extra + (pw - HIST_WARPS_IN_SPARE) * HIST_WARP_B;
  }
  __device__ __forceinline__ int operator()(int qrow, int slot) const {
    int c = 0;
#pragma unroll
    for (int pw = 0; pw < simtc::POOL_WARPS; ++pw) c += warp_hist(pw)[qrow * HIST_PITCH_B + slot];
    return c;
  }
};

template <int SLOTS, class Counts, class Sync>
__device__ __forceinline__ void drmm_finish(const DrmmArgs& a, int pair, const Counts cnt, float* z, int warp, int lane, Sync sync) {
  const int nslots = a.nbins + 1;
  const long long* qids = a.q + (size_t)pair * a.Q;
  for (int r = 0; r < 4; ++r) {
    const int qrow = warp * 4 + r;
    if (qrow >= a.Q) continue;  // warp-uniform
    float h0 = lane < nslots ? (float)(cnt(qrow, lane) + 1) : 0.f;
    float h1 = (SLOTS > 32 && lane + 32 < nslots) ? (float)(cnt(qrow, (lane + 32) % SLOTS) + 1) : 0.f;
    if (a.hist_type == CAPR_DRMM_NH) {
      const float tot = warp_sum(h0 + h1);
      h0 /= tot;
      h1 /= tot;
    } else if (a.hist_type == CAPR_DRMM_LCH) {
      h0 = lane < nslots ? logf(h0) : 0.f;
      h1 = lane + 32 < nslots ? logf(h1) : 0.f;
    }
    if (a.hist_out) {
      float* ho = a.hist_out + ((size_t)pair * a.Q + qrow) * nslots;
      if (lane < nslots) ho[lane] = h0;
      if (lane + 32 < nslots) ho[lane + 32] = h1;
    }
    float acc2 = 0.f;
    for (int n = 0; n < a.nodes; ++n) {
      const float* w = a.ffw_w1 + n * nslots;
      float p = lane < nslots ? w[lane] * h0 : 0.f;
      if (lane + 32 < nslots) p = fmaf(w[lane + 32], h1, p);
      p = warp_sum(p) + a.ffw_b1[n];
      acc2 = fmaf(a.ffw_w2[n], tanhf(p), acc2);
    }
    if (lane == 0) z[qrow] = tanhf(acc2 + a.ffw_b2[0]);
  }
  sync();
  if (warp == 0) {
    // term gate: softmax over the Q query positions of w_g*idf (IDF) or w_g.emb[q] (TV), pads at -1e7
    float logit = -INFINITY;
    if (lane < a.Q) {
      const long long qid = qids[lane];
      const float pad_bias = (qid == 0) ? -1e7f : 0.f;  // (1 - q_mask) * -1e7   DRMM.py:89
      float g;
      if (a.gate_type == CAPR_DRMM_GATE_IDF) {
        g = a.gate_w[0] * a.idf[(size_t)pair * a.Q + lane];
      } else {
        // DRMM.py:109 indexes the table with the raw query ids (the reference raises on OOV ids; we read <pad>)
        const float* e = a.raw_emb + (size_t)table_row(qid, a.V) * a.E;
        g = 0.f;
        for (int k = 0; k < a.E; ++k) g = fmaf(a.gate_w[k], e[k], g);
      }
      logit = g + pad_bias;
    }
    const float m = warp_max(logit);
    const float ex = lane < a.Q ? expf(logit - m) : 0.f;
    const float den = warp_sum(ex);
    const float x = warp_sum(lane < a.Q ? (ex / den) * z[lane] : 0.f);
    if (lane == 0) a.scores[pair] = fmaf(a.out_w[0], x, a.out_b[0]);  // output_layer  DRMM.py:114
  }
}

__global__ void __launch_bounds__(NT, 1) drmm_kernel(const DrmmArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SimTile s = carve_sim_tile(smem_raw, a.pitch);
  int* cnt = reinterpret_cast<int*>(smem_raw + sim_tile_bytes(a.pitch));  // [QT][MAX_SLOTS]
  float* ub = reinterpret_cast<float*>(cnt + QT * MAX_SLOTS);             // [MAX_SLOTS]
  float* z = ub + MAX_SLOTS;                                              // [QT] ffw output per query term
  clear_sim_tile(s, tid);
  if (tid < a.nbins) ub[tid] = a.bin_ub[tid];
  __syncthreads();
  for (int pair = blockIdx.x; pair < a.B; pair += gridDim.x) {
    for (int i = tid; i < QT * MAX_SLOTS; i += NT) cnt[i] = 0;
    const long long* qids = a.q + (size_t)pair * a.Q;
    const long long* dids = a.d + (size_t)pair * a.D;
    for (int d0 = 0; d0 < a.D; d0 += DT) {
      build_sim_tile(s, a.table, a.pitch, a.V, qids, a.Q, dids, d0, a.D, d0 == 0, tid);  // also orders the cnt reset
      drmm_count_tile<MAX_SLOTS, SIM_PITCH, DT>(s.sim, s.qid, s.did, min(DT, a.D - d0), a, ub, cnt, warp, lane);
      __syncthreads();
    }
    drmm_finish<MAX_SLOTS>(a, pair, ArrayCounts<MAX_SLOTS>{cnt}, z, warp, lane, BlockSync());
    __syncthreads();  // z / cnt are rewritten by the next pair
  }
}

// Engine 2: cosine tile from the tcgen05 producer (simtc.cuh, pipelined epilogue); counting + finish on the 8 pooling
// warps, one half tile (256 docs) at a time.  MODE = DRMM_POOL_*: by default every (pooling warp, query row) owns a byte
// histogram (the lane is the only writer of its row: plain loads and stores), summed over the 8 warps by drmm_finish.
template <int MODE>
__global__ void __launch_bounds__(simtc::THREADS_PIPE, 1) drmm_tc_kernel(const DrmmArgs a) {
  using namespace simtc;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem s = carve(smem_raw, (a.pr.pitch + ATOM_K - 1) / ATOM_K, a.pr.deep != 0);
  constexpr bool PRIV = MODE != DRMM_POOL_ATOMIC;
  static_assert((QT * CNT_PITCH_TC + MAX_SLOTS_TC + QT) <= SPARE_FLOATS, "DRMM scratch must fit behind the two half tiles");
  static_assert(HIST_WARPS_IN_SPARE * HIST_WARP_B + (MAX_SLOTS_TC + QT) * 4 <= SPARE_FLOATS * 4, "private histograms must fit behind the two half tiles");
  static_assert(MAX_SLOTS_TC <= HIST_PITCH_B && HIST_PITCH_B % 4 == 0, "histogram row holds every slot, word-aligned");
  unsigned char* spare = reinterpret_cast<unsigned char*>(spare_scratch(s));
  unsigned char* extra = reinterpret_cast<unsigned char*>(s.extra);  // DRMM_TC_EXTRA_BYTES
  int* cnt = reinterpret_cast<int*>(spare);                          // atomic mode: [QT][CNT_PITCH_TC]
  float* ub = PRIV ? reinterpret_cast<float*>(spare + HIST_WARPS_IN_SPARE * HIST_WARP_B) : reinterpret_cast<float*>(cnt + QT * CNT_PITCH_TC);  // [MAX_SLOTS_TC]
  float* z = ub + MAX_SLOTS_TC;                                      // [QT]
  const uint32_t tmem_base = setup(s, tid, THREADS_PIPE, MMA_WARP_PIPE);
  if (is_producer_warp(warp)) {
    producer_loop(s, a.pr, producer_index(warp) * 32 + lane);
  } else if (warp == MMA_WARP_PIPE) {
    mma_loop(s, a.pr, tmem_base);
  } else if (is_drain_warp(warp)) {
    drain_loop(s, a.pr, tmem_base, warp, lane);
  } else if (is_pool_warp(warp)) {
    const int pw = pool_index(warp), ptid = pw * 32 + lane;
    const int halves = halves_of(a.pr);
    if (ptid < a.nbins) ub[ptid] = a.bin_ub[ptid];
    const PrivateCounts priv{spare, extra};
    unsigned char* hist_row = const_cast<unsigned char*>(priv.warp_hist(pw)) + lane * HIST_PITCH_B;
    PoolSync ps;
    int unit = 0, it = 0, sink = 0;
    for (int pair = blockIdx.x; pair < a.B; pair += gridDim.x, ++it) {
      if (PRIV) {
#pragma unroll
        for (int w = 0; w < HIST_PITCH_B / 4; ++w) reinterpret_cast<uint32_t*>(hist_row)[w] = 0u;
      } else {
        for (int i = ptid; i < QT * CNT_PITCH_TC; i += POOL_WARPS * 32) cnt[i] = 0;
      }
      epi_barrier();  // the 8 pooling warps: counters (and, first time, the bounds) are in place
      for (int h = 0; h < halves; ++h, ++unit) {
        const int ub_i = unit & 1;
        ps.wait_full(s, ub_i);  // also orders the drain warps' id writes of this pair before the reads below
        if (MODE != DRMM_POOL_SKIP) {
          const int qi = s.qid[(it & 1) * QT + lane];
          drmm_count_slice<MODE>(half_tile(s, ub_i), HALF_PITCH, pw * 32, min(NT_DOCS, a.D - h * NT_DOCS) - pw * 32, qi,
                                 s.did + (it & 1) * s.dcap + h * NT_DOCS + pw * 32, a, ub, cnt + lane * CNT_PITCH_TC, hist_row, sink);
        }
        ps.release(s, ub_i, lane);
      }
      epi_barrier();
      if (PRIV)
        drmm_finish<HIST_PITCH_B>(a, pair, priv, z, pw, lane, EpiSync());
      else
        drmm_finish<CNT_PITCH_TC>(a, pair, ArrayCounts<CNT_PITCH_TC>{cnt}, z, pw, lane, EpiSync());
      epi_barrier();  // z / the counters are rewritten by the next pair
    }
    if (MODE == DRMM_POOL_NOADD && sink == 0x7fffffff) z[0] = 1.f;  // keeps the profiling-only mode's bin arithmetic alive
  }
  teardown(s, tmem_base, tid, MMA_WARP_PIPE);
}

}  // namespace capr

using namespace capr;

extern "C" int capr_drmm_forward(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                                 const float* table, int V, int pitch, const float* raw_emb, int E, int nbins,
                                 const float* bin_ub, int hist_type, int gate_type, const float* ffw_w1,
                                 const float* ffw_b1, int nodes, const float* ffw_w2, const float* ffw_b2,
                                 const float* gate_w, const float* out_w, const float* out_b, float* scores,
                                 float* hist_out, capr_stream_t stream) {
  capr::DeviceGuard device_guard(table);  // act on the device that owns the caller's buffers
  const char* fn = "capr_drmm_forward";
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0 && nodes > 0 && nbins > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d nodes=%d nbins=%d", fn, B, Q, D, V, nodes, nbins);
  CAPR_REQUIRE(pitch > 0 && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: table pitch %d must be a positive multiple of 16", fn, pitch);
  CAPR_REQUIRE(B == 0 || (query && doc && table && bin_ub && ffw_w1 && ffw_b1 && ffw_w2 && ffw_b2 && gate_w && out_w && out_b && scores), CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(((uintptr_t)table & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: table must be 16-byte aligned", fn);
  CAPR_REQUIRE(hist_type >= 0 && hist_type <= 2, CAPR_ERR_BAD_SHAPE, "%s: histType should be CH, NH or LCH", fn);
  CAPR_REQUIRE(gate_type == 0 || gate_type == 1, CAPR_ERR_BAD_SHAPE, "%s: gateType should be IDF or TV", fn);
  CAPR_REQUIRE(gate_type != CAPR_DRMM_GATE_IDF || idf, CAPR_ERR_BAD_POINTER, "%s: IDF gate needs idf", fn);
  CAPR_REQUIRE(gate_type != CAPR_DRMM_GATE_TV || (raw_emb && E > 0), CAPR_ERR_BAD_POINTER, "%s: TV gate needs the raw embedding table", fn);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(pitch <= MAX_PITCH, CAPR_ERR_UNSUPPORTED, "%s: embedding dim > %d is not supported yet", fn, MAX_PITCH);
  CAPR_REQUIRE(nbins + 1 <= MAX_SLOTS, CAPR_ERR_UNSUPPORTED, "%s: nbins=%d > %d is not supported", fn, nbins, MAX_SLOTS - 1);
  if (B == 0) return CAPR_OK;
  DrmmArgs a{(const long long*)query, (const long long*)doc, idf, B, Q, D, V, pitch, E, nbins, hist_type, gate_type, nodes,
             table, raw_emb, bin_ub, ffw_w1, ffw_b1, ffw_w2, ffw_b2, gate_w, out_w, out_b, scores, hist_out, simtc::Problem{}, 0};
  size_t smem = sim_tile_bytes(pitch) + (size_t)QT * MAX_SLOTS * sizeof(int) + (MAX_SLOTS + QT) * sizeof(float);
  CAPR_CHECK_CUDA(cudaFuncSetAttribute(drmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  drmm_kernel<<<B < sms ? B : sms, NT, smem, (cudaStream_t)stream>>>(a);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

// Engine 2 (tensor cores): same contract as capr_drmm_forward; table given as bf16 (hi, lo) planes (capr_table_prepare_bf16).
extern "C" int capr_drmm_forward_tc(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                                    const void* table_hi, const void* table_lo, int V, int pitch, const float* raw_emb, int E, int nbins,
                                    const float* bin_ub, int hist_type, int gate_type, const float* ffw_w1, const float* ffw_b1,
                                    int nodes, const float* ffw_w2, const float* ffw_b2, const float* gate_w, const float* out_w,
                                    const float* out_b, float* scores, float* hist_out, capr_stream_t stream) {
  capr::DeviceGuard device_guard(table_hi);  // act on the device that owns the caller's buffers
  const char* fn = "capr_drmm_forward_tc";
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0 && E > 0 && nodes > 0 && nbins > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d Q=%d D=%d V=%d E=%d nodes=%d nbins=%d", fn, B, Q, D, V, E, nodes, nbins);
  CAPR_REQUIRE(pitch >= E && pitch % 16 == 0, CAPR_ERR_BAD_SHAPE, "%s: pitch=%d must be a multiple of 16 and >= E (capr_table_pitch_bf16)", fn, pitch);
  CAPR_REQUIRE(hist_type >= 0 && hist_type <= 2, CAPR_ERR_BAD_SHAPE, "%s: histType should be CH, NH or LCH", fn);
  CAPR_REQUIRE(gate_type == 0 || gate_type == 1, CAPR_ERR_BAD_SHAPE, "%s: gateType should be IDF or TV", fn);
  CAPR_REQUIRE(Q <= QT, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d > %d is not supported by the fused kernels yet", fn, Q, QT);
  CAPR_REQUIRE(D <= simtc::DEEP_DCAP && pitch <= simtc::MAX_ATOMS * simtc::ATOM_K && nbins + 1 <= MAX_SLOTS_TC, CAPR_ERR_UNSUPPORTED,
               "%s: needs maxdoclen <= %d, emb dim <= %d, nbins <= %d: use capr_drmm_forward", fn, simtc::DEEP_DCAP, simtc::MAX_ATOMS * simtc::ATOM_K, MAX_SLOTS_TC - 1);
  CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table of %d x %d elements is too large for 32-bit row offsets", fn, V, pitch);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(query && doc && table_hi && table_lo && bin_ub && ffw_w1 && ffw_b1 && ffw_w2 && ffw_b2 && gate_w && out_w && out_b && scores, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(gate_type != CAPR_DRMM_GATE_IDF || idf, CAPR_ERR_BAD_POINTER, "%s: IDF gate needs idf", fn);
  CAPR_REQUIRE(gate_type != CAPR_DRMM_GATE_TV || raw_emb, CAPR_ERR_BAD_POINTER, "%s: TV gate needs the raw embedding table", fn);
  DrmmArgs a{(const long long*)query, (const long long*)doc, idf, B, Q, D, V, pitch, E, nbins, hist_type, gate_type, nodes,
             nullptr, raw_emb, bin_ub, ffw_w1, ffw_b1, ffw_w2, ffw_b2, gate_w, out_w, out_b, scores, hist_out,
             simtc::Problem{(const long long*)query, (const long long*)doc, B, Q, D, V, (const __nv_bfloat16*)table_hi, (const __nv_bfloat16*)table_lo, pitch, E, 0}, 0};
  const int atoms = (pitch + simtc::ATOM_K - 1) / simtc::ATOM_K;
  const char* ring_env = getenv("CAPR_SIM_RING");  // see capr_knrm_forward_tc
  a.pr.deep = (D > DT || (atoms >= 3 && !(ring_env && ring_env[0] == '2'))) ? 1 : 0;  // maxdoclen > 512 needs the deep layout's id arrays
  const size_t smem = simtc::smem_bytes(atoms, DRMM_TC_EXTRA_BYTES, a.pr.deep != 0);
  CAPR_REQUIRE(smem <= simtc::MAX_DYN_SMEM, CAPR_ERR_UNSUPPORTED, "%s: %zu bytes of shared memory needed", fn, smem);
  // CAPR_DRMM_POOL = atomic | noadd | skip selects the A/B reference and the profiling-only ablations (results invalid for the last two)
  const char* pool_env = getenv("CAPR_DRMM_POOL");
  a.pool_mode = !pool_env ? DRMM_POOL_PRIVATE : pool_env[0] == 'a' ? DRMM_POOL_ATOMIC : pool_env[0] == 'n' ? DRMM_POOL_NOADD : pool_env[0] == 's' ? DRMM_POOL_SKIP : DRMM_POOL_PRIVATE;
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const int grid = B < sms ? B : sms;
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, simtc::THREADS_PIPE, smem, (cudaStream_t)stream>>>(a);
    return cudaSuccess;
  };
  switch (a.pool_mode) {
    case DRMM_POOL_ATOMIC: CAPR_CHECK_CUDA(launch(drmm_tc_kernel<DRMM_POOL_ATOMIC>)); break;
    case DRMM_POOL_NOADD: CAPR_CHECK_CUDA(launch(drmm_tc_kernel<DRMM_POOL_NOADD>)); break;
    case DRMM_POOL_SKIP: CAPR_CHECK_CUDA(launch(drmm_tc_kernel<DRMM_POOL_SKIP>)); break;
    default: CAPR_CHECK_CUDA(launch(drmm_tc_kernel<DRMM_POOL_PRIVATE>)); break;
  }
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}
