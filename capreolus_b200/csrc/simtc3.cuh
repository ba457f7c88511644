// Engine 3 of the KNRM-family kernels: the query x doc cosine tile is pooled STRAIGHT FROM TENSOR MEMORY, documents are scored
// in their TERM-FREQUENCY form, and the freed shared memory goes into the gather ring.
//
//   SimilarityMatrix.forward      capreolus/reranker/common.py:170-182   (cosine tile + OOV exact match)
//   KNRM_class.forward            capreolus/reranker/KNRM.py:39-55       (sum over ALL doc positions of K Gaussian kernels)
//
// What the ncu capture of engine 2 said (profiles/r01_v12_knrm_tc_kernel_ncu_full.json): nothing saturated -- L2 29 %, XU 51 %,
// tensor 44 %, issue 60 % -- the kernel waits on row gathers with 96 KB in flight per SM, and 74 KB of shared memory are a
// staging tile between TMEM and the pooling warps.  Engine 3 changes three things:
//
//  1. TERM FREQUENCIES.  KNRM (and DRMM) pool with a SUM over doc positions, so a document is its multiset of tokens: a pre-pass
//     (tf_dedup_kernel) turns each doc into (distinct token, count) pairs in first-occurrence order, <pad> included as a token of
//     its own (its row is all-zero, cosine exactly 0, weight = number of pads: KNRM.py:50 sums over padded positions too).  Each
//     DISTINCT token is gathered once and its cosine column is weighted by the count: a 512-token zipf document has ~310 distinct
//     tokens -> ~40 % fewer gathered rows, MMA columns and exponentials, bit-reproducibly (fixed first-occurrence order).
//  2. POOLING FROM TMEM.  A warp may only read the 32 TMEM lanes of its quarter (warp % 4), so the 32 query rows of a work
//     unit are steered into lane quarter  u % 4  by moving the START ADDRESS of the A-operand descriptor back by 4 096 bytes
//     per quarter: MMA row 32q + j then reads query row j, the other 96 rows read whatever precedes / follows the tile (finite
//     or not, they only produce accumulator lanes nobody reads).  The three products q_hi.d_hi + q_lo.d_hi + q_hi.d_lo are
//     accumulated into the SAME accumulator rows (3 MMAs per K step, N <= 128), so a TMEM lane holds finished cosines and the
//     pooling warps of quarter q read them with tcgen05.ld.32x32b (lane = query row) -- no drain warps, no staging tile, and
//     four 128-column accumulators decouple MMA from pooling four units deep.
//  3. A DEEPER RING.  The 74 KB tile and the drain hand-offs are gone: 32 KB stages (128 doc rows x one 64-element K atom, hi and
//     lo planes), 5 stages with one query buffer (default) or 3 with two (CAPR_SIM3_QBUFS=2), i.e. up to 160 KB of gathers in flight.
//
// Work unit = (pair, 128 distinct doc tokens).  Roles (14 warps): 0-7 pooling (quarter = warp % 4, column half = warp / 4),
// 8-11 gather producers, 12 MMA issuer + TMEM owner, 13 finisher (cross-warp reduction, log, combine).
#pragma once
#include "simtile.cuh"
#include "tc_common.cuh"

namespace capr {
namespace simtc3 {

constexpr int POOL_WARPS = 8, PROD_WARPS = 4;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int THREADS = (POOL_WARPS + PROD_WARPS + 2) * 32;  // 448
constexpr int MMA_WARP = POOL_WARPS + PROD_WARPS;            // 12
constexpr int FIN_WARP = MMA_WARP + 1;                       // 13
constexpr int ATOM_K = 64;                                   // bf16 elements per 128-byte swizzle row
constexpr int MAX_ATOMS = 5;                                 // pitch <= 320
constexpr int U_DOCS = 128;                                  // distinct doc tokens per work unit (= max MMA N)
constexpr int PLANE_BYTES = U_DOCS * 128;                    // one plane of 128 rows x one K atom = 16 KB
constexpr int STAGE_BYTES = 2 * PLANE_BYTES;                 // a stage = hi plane + lo plane of one K atom = 32 KB: the hand-off of a stage costs
                                                             // ~450-700 cycles whatever its size (producer arrive -> MMA wait -> commit -> producer
                                                             // wait; all-off ablation), so a unit is 5 hand-offs instead of 10
constexpr int Q_ATOM_BYTES = 64 * 128;                       // per K atom: rows 0-31 = q_hi, rows 32-63 = q_lo
constexpr int Q_PLANE_BYTES = 32 * 128;                      // 4 KB: the A descriptor of quarter q starts q * 4 KB before the tile
constexpr int MAX_STAGES = 6;
constexpr int MAX_DCAP = 1024;                               // maxdoclen <= 1024
constexpr int IDS_PER_THREAD = MAX_DCAP / PROD_THREADS;      // 8
constexpr size_t MAX_DYN_SMEM = 232448;                      // 227 KB
constexpr int N_BARS = 2 + 2 + 2 * MAX_STAGES + 4 + 4 + 2 + 2 + 2;

struct Problem {
  const long long* q;          // [B,Q] raw query ids (int64, reference layout)
  const int* tf_ids;           // [B,D] distinct doc tokens of each pair in first-occurrence order, 0 beyond tf_nd
  const unsigned short* tf_cnt;  // [B,D] their multiplicities (0 beyond tf_nd)
  const int* tf_nd;            // [B]   number of distinct tokens (<pad> counts as one)
  int B, Q, D, V;
  const __nv_bfloat16* hi;     // [V][pitch]
  const __nv_bfloat16* lo;
  int pitch, E;
  int n_stages, n_qbufs, dcap;  // shared-memory layout chosen on the host (dcap = D rounded up to 128)
  int debug;                    // debug build only (CAPR_SIM3_DEBUG; results invalid): 1 no pooling math, 2 no MMAs, 4 no gathers, 8 no TMEM loads
  long long* trace;             // debug build only (CAPR_SIM3_TRACE=<device pointer>): CTA 0 records (clock64 << 8 | tag) per role, 1024 slots each
};

// trace roles: 0 producer thread 0, 1 MMA lane 0, 2 pooling warp 0 lane 0, 3 pooling warp 5 lane 0, 4 finisher lane 0
#ifdef CAPR_DEBUG_BUILD
#define SIM3_TR(role, tag)                                                                                          \
  do {                                                                                                              \
    if (pr.trace && blockIdx.x == 0 && tr_n < 1024) pr.trace[(role) * 1024 + tr_n++] = (clock64() << 8) | (tag); \
  } while (0)
#else
#define SIM3_TR(role, tag) \
  do {                     \
  } while (0)
#endif

struct Smem {
  unsigned char* ring;   // n_stages x 16 KB
  unsigned char* q0;     // n_qbufs x atoms x 8 KB
  int q_stride;
  int* qid;              // [2][QT]      ids of the pair (by pair parity): exact-match rules in the pooling warps
  int* did;              // [2][dcap]
  unsigned short* cnt;   // [2][dcap]
  int* nd;               // [2]
  int* qrow;             // [QT]         table rows (producer only)
  int* drow;             // [dcap]
  float* red;            // [POOL_WARPS][KT+1][32]
  uint64_t *q_full, *q_empty, *d_full, *d_empty, *acc_full, *acc_empty, *ids_full, *ids_empty, *red_full, *red_empty;
  uint32_t* tmem_slot;
  int dcap;
  __device__ __forceinline__ unsigned char* qbuf(int b) const { return q0 + b * q_stride; }
  __device__ __forceinline__ unsigned char* stage(int i) const { return ring + i * STAGE_BYTES; }
};

__host__ __device__ inline size_t fixed_bytes(int atoms, int n_qbufs, int dcap, int red_floats) {
  return 1024 /*alignment*/ + (size_t)n_qbufs * atoms * Q_ATOM_BYTES + (size_t)(2 * QT + 2 * dcap) * 4 + (size_t)2 * dcap * 2 + 2 * 4 + (size_t)(QT + dcap) * 4 +
         (size_t)red_floats * 4 + N_BARS * 8 + 64;
}
__host__ inline int stages_that_fit(int atoms, int n_qbufs, int dcap, int red_floats) {
  const size_t fixed = fixed_bytes(atoms, n_qbufs, dcap, red_floats);
  if (fixed >= MAX_DYN_SMEM) return 0;
  const int n = (int)((MAX_DYN_SMEM - fixed) / STAGE_BYTES);
  return n > MAX_STAGES ? MAX_STAGES : n;
}
__host__ __device__ inline size_t smem_bytes(int atoms, int n_stages, int n_qbufs, int dcap, int red_floats) {
  return fixed_bytes(atoms, n_qbufs, dcap, red_floats) + (size_t)n_stages * STAGE_BYTES;
}

__device__ __forceinline__ Smem carve(unsigned char* raw, const Problem& pr, int atoms, int red_floats) {
  Smem s;
  unsigned char* p = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);  // (offset arithmetic: keeps the shared address space)
  s.ring = p;
  p += pr.n_stages * STAGE_BYTES;  // the ring comes FIRST: >= 12 KB must precede the query tile (quarter-3 descriptors start 12 KB before it)
  s.q0 = p;
  s.q_stride = atoms * Q_ATOM_BYTES;
  p += pr.n_qbufs * atoms * Q_ATOM_BYTES;
  s.dcap = pr.dcap;
  s.qid = reinterpret_cast<int*>(p);
  s.did = s.qid + 2 * QT;
  p += (2 * QT + 2 * pr.dcap) * 4;
  s.cnt = reinterpret_cast<unsigned short*>(p);
  p += 2 * pr.dcap * 2;
  s.nd = reinterpret_cast<int*>(p);
  s.qrow = s.nd + 2;
  s.drow = s.qrow + QT;
  p += (2 + QT + pr.dcap) * 4;
  s.red = reinterpret_cast<float*>(p);
  p += red_floats * 4;
  p += (8 - (tc::smem_u32(p) & 7)) & 7;
  uint64_t* b = reinterpret_cast<uint64_t*>(p);
  s.q_full = b, s.q_empty = b + 2, s.d_full = b + 4, s.d_empty = b + 4 + MAX_STAGES, s.acc_full = b + 4 + 2 * MAX_STAGES,
  s.acc_empty = s.acc_full + 4, s.ids_full = s.acc_empty + 4, s.ids_empty = s.ids_full + 2, s.red_full = s.ids_empty + 2, s.red_empty = s.red_full + 1;
  s.tmem_slot = reinterpret_cast<uint32_t*>(b + N_BARS);
  return s;
}

__device__ __forceinline__ void prod_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(PROD_THREADS) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int units_of(int nd) { return nd <= U_DOCS ? 1 : (nd + U_DOCS - 1) / U_DOCS; }
__device__ __forceinline__ int live_cols(int nd, int u) { return min(U_DOCS, max(0, nd - u * U_DOCS)); }
__device__ __forceinline__ int mma_n(int live) { return live <= 16 ? 16 : (live + 15) & ~15; }  // UMMA N: multiple of 16 in [16, 256] for M = 128

// Common prologue: barriers + TMEM.  Call from all threads; returns the TMEM base.
__device__ __forceinline__ uint32_t setup(const Smem& s, const Problem& pr, int tid) {
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s.q_full[i], PROD_THREADS);
      tc::mbar_init(&s.q_empty[i], 1);
      tc::mbar_init(&s.ids_full[i], PROD_THREADS);
      tc::mbar_init(&s.ids_empty[i], POOL_WARPS + 1);  // the pooling warps + the MMA warp
    }
    for (int i = 0; i < MAX_STAGES; ++i) {
      tc::mbar_init(&s.d_full[i], PROD_THREADS);
      tc::mbar_init(&s.d_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&s.acc_full[i], 1);
      tc::mbar_init(&s.acc_empty[i], 2);  // the two pooling warps of the quarter
    }
    tc::mbar_init(s.red_full, POOL_WARPS);
    tc::mbar_init(s.red_empty, 1);
    tc::fence_barrier_init();
  }
  if ((tid >> 5) == MMA_WARP) tc::tmem_alloc(s.tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  return *s.tmem_slot;
}

__device__ __forceinline__ void teardown(const Smem& s, uint32_t tmem_base, int tid) {
  tc::tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---- producer warps (4): ids -> shared memory, query block and doc stages -> the operand layouts ------------------------------
// Rows are copied with 16-byte cp.async straight into the SWIZZLE_128B K-major layout (8 lanes fetch one 128-byte row segment:
// fully coalesced) and completed on the stage's mbarrier with cp.async.mbarrier.arrive.noinc, so the issuing threads never wait
// for data.  <pad> / OOV tokens (table row 0) are zero-filled without touching global memory (src-size 0); rows past the unit's
// live tokens (rounded up to the MMA's N) are not written at all.
__device__ __forceinline__ void producer_loop(const Smem& s, const Problem& pr, int ptid /*0..127*/) {
  const int atoms = (pr.pitch + ATOM_K - 1) / ATOM_K;
  const int last_chunks = (pr.pitch - (atoms - 1) * ATOM_K) / 8;  // 16-byte chunks that exist in the last atom
  const int sub = ptid & 7;    // 16-byte chunk inside the 128-byte row segment
  const int rsub = ptid >> 3;  // 0..15: this thread serves rows rsub + 16*j
  int stage = 0, it = 0;
  uint32_t d_phase = 0;
  // ids of the NEXT pair are fetched into registers while this pair's gathers are being issued (they stream from HBM)
  long long q_next = 0;
  int nd_next = 0, d_next[IDS_PER_THREAD];
  unsigned short c_next[IDS_PER_THREAD];
  auto fetch_ids = [&](int pair) {
    const bool have = pair < pr.B;
    q_next = (have && ptid < pr.Q) ? pr.q[(size_t)pair * pr.Q + ptid] : 0;  // ptid < Q <= QT
    nd_next = have ? pr.tf_nd[pair] : 0;
#pragma unroll
    for (int j = 0; j < IDS_PER_THREAD; ++j) {
      const int i = ptid + PROD_THREADS * j;
      const bool in = have && i < pr.D;
      d_next[j] = in ? pr.tf_ids[(size_t)pair * pr.D + i] : 0;
      c_next[j] = in ? pr.tf_cnt[(size_t)pair * pr.D + i] : (unsigned short)0;
    }
  };
  fetch_ids(blockIdx.x);
  int tr_n = ptid == 0 ? 0 : 1024;
  (void)tr_n;
  for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
    const int pp = it & 1;
    prod_barrier();  // every producer thread is done reading the previous pair's qrow / drow
    SIM3_TR(0, 1);
    tc::mbar_wait(&s.ids_empty[pp], (uint32_t)(((it >> 1) & 1) ^ 1));  // pooling + MMA are done with the ids of pair it-2
    SIM3_TR(0, 2);
    const int nd = nd_next;
    if (ptid < QT) {
      s.qid[pp * QT + ptid] = id_as_int(q_next);
      s.qrow[ptid] = table_row(q_next, pr.V);
    }
    if (ptid == 0) s.nd[pp] = nd;
#pragma unroll
    for (int j = 0; j < IDS_PER_THREAD; ++j) {
      const int i = ptid + PROD_THREADS * j;
      if (i < pr.dcap) {
        s.did[pp * pr.dcap + i] = d_next[j];
        s.cnt[pp * pr.dcap + i] = c_next[j];
        s.drow[i] = table_row((long long)d_next[j], pr.V);
      }
    }
    tc::mbar_arrive(&s.ids_full[pp]);  // (release: the plain stores above are visible to the waiters)
    prod_barrier();
    SIM3_TR(0, 3);
    fetch_ids(pair + gridDim.x);
    const int units = units_of(nd);
    // query block: per K atom a 64-row tile, rows 0-31 = hi plane, rows 32-63 = lo plane of the 32 query tokens
    const int b = pr.n_qbufs == 2 ? pp : 0;
    const int qn = pr.n_qbufs == 2 ? (it >> 1) : it;
    tc::mbar_wait(&s.q_empty[b], (uint32_t)((qn & 1) ^ 1));
    SIM3_TR(0, 4);
    {
      const uint32_t qbase = tc::smem_u32(s.qbuf(b));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rsub + 16 * j;  // 0..63
        const int trow = s.qrow[r & 31];
        const __nv_bfloat16* src = (r < 32 ? pr.hi : pr.lo) + (size_t)trow * pr.pitch + sub * 8;
        const uint32_t dst = qbase + r * 128 + ((sub ^ (r & 7)) << 4);
        const uint32_t nbytes = trow != 0 ? 16u : 0u;
        for (int a = 0; a < atoms; ++a)
          if (a + 1 < atoms || sub < last_chunks)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + a * Q_ATOM_BYTES), "l"(src + a * ATOM_K), "r"(nbytes) : "memory");
      }
    }
    cp_async_arrive_noinc(&s.q_full[b]);
    SIM3_TR(0, 5);
    for (int u = 0; u < units; ++u) {
      SIM3_TR(0, 6);
      const int n16 = mma_n(live_cols(nd, u));
      unsigned off[8];   // element offsets of this thread's 8 rows (V * pitch < 2^31 is checked on the host)
      unsigned live = 0;  // bit j: row j is a real table row
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int trow = s.drow[u * U_DOCS + rsub + 16 * j];
        off[j] = (unsigned)trow * (unsigned)pr.pitch + (unsigned)(sub * 8);
        live |= (trow != 0 ? 1u : 0u) << j;
      }
      for (int a = 0; a < atoms; ++a) {
        tc::mbar_wait(&s.d_empty[stage], d_phase ^ 1);
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          const __nv_bfloat16* tab = (plane == 0 ? pr.hi : pr.lo) + a * ATOM_K;
          const uint32_t base = tc::smem_u32(s.stage(stage)) + plane * PLANE_BYTES;
          if (a + 1 < atoms || sub < last_chunks) {  // tail chunks of a partial last atom are never read
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = rsub + 16 * j;
              cp_async16_pred(base + r * 128 + ((sub ^ (r & 7)) << 4), tab + off[j], ((live >> j) & 1u) << 4, r < n16 && !CAPR_DBG(pr.debug & 4));
            }
          }
        }
        cp_async_arrive_noinc(&s.d_full[stage]);
        if (++stage == pr.n_stages) stage = 0, d_phase ^= 1;
      }
    }
    SIM3_TR(0, 7);
  }
  cp_async_commit();
  cp_async_wait<0>();  // nothing may still be landing in shared memory when the CTA tears down
}

// ---- MMA issuer: the WHOLE warp runs the loop (waits are warp-wide), one elected lane issues --------------------------------
// Unit g of this CTA accumulates into TMEM columns [128 (g%4), +N) with the query rows in lane quarter g%4.
__device__ __forceinline__ void mma_loop(const Smem& s, const Problem& pr, uint32_t tmem_base, int lane) {
  const int atoms = (pr.pitch + ATOM_K - 1) / ATOM_K;
  int stage = 0, it = 0, g = 0;
  uint32_t d_phase = 0;
  int tr_n = lane == 0 ? 0 : 1024;
  (void)tr_n;
  for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
    const int pp = it & 1;
    SIM3_TR(1, 10);
    tc::mbar_wait(&s.ids_full[pp], (uint32_t)((it >> 1) & 1));
    SIM3_TR(1, 11);
    const int nd = s.nd[pp];
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&s.ids_empty[pp]);  // this warp only needs the count
    const int units = units_of(nd);
    const int b = pr.n_qbufs == 2 ? pp : 0;
    const int qn = pr.n_qbufs == 2 ? (it >> 1) : it;
    tc::mbar_wait(&s.q_full[b], (uint32_t)(qn & 1));
    SIM3_TR(1, 12);
    const uint32_t qaddr = tc::smem_u32(s.qbuf(b));
    for (int u = 0; u < units; ++u, ++g) {
      const int buf = g & 3;
      tc::mbar_wait(&s.acc_empty[buf], (uint32_t)(((g >> 2) & 1) ^ 1));
      SIM3_TR(1, 13);
      tc::tc_fence_after();
      const uint32_t idesc = tc::make_instr_desc(tc::FMT_BF16, 128, mma_n(live_cols(nd, u)));
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * U_DOCS);
      const uint32_t a_base = qaddr - (uint32_t)(buf * Q_PLANE_BYTES);  // MMA rows [32 buf, 32 buf + 32) read the query rows
      for (int a = 0; a < atoms; ++a) {
        const uint64_t a_hi = tc::make_sw128_kmajor_desc(a_base + a * Q_ATOM_BYTES);
        const uint64_t a_lo = tc::make_sw128_kmajor_desc(a_base + a * Q_ATOM_BYTES + Q_PLANE_BYTES);
        const int ksteps = CAPR_DBG(pr.debug & 2) ? 0 : min(ATOM_K, pr.pitch - a * ATOM_K) / 16;  // a partial last atom has fewer K steps
        // one stage = d_hi and d_lo of this K atom: q_hi.d_hi + q_lo.d_hi + q_hi.d_lo
        tc::mbar_wait(&s.d_full[stage], d_phase);
        if (a == 0) SIM3_TR(1, 14);
        tc::tc_fence_after();
        {
          const uint64_t bd_hi = tc::make_sw128_kmajor_desc(tc::smem_u32(s.stage(stage)));
          const uint64_t bd_lo = tc::make_sw128_kmajor_desc(tc::smem_u32(s.stage(stage)) + PLANE_BYTES);
          if (tc::elect_one()) {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t koff = (uint64_t)(k * 2);  // 32 bytes per K=16 step, in 16-byte units
              tc::umma_f16(d_tmem, a_hi + koff, bd_hi + koff, idesc, (a | k) != 0);
              tc::umma_f16(d_tmem, a_lo + koff, bd_hi + koff, idesc, true);
              tc::umma_f16(d_tmem, a_hi + koff, bd_lo + koff, idesc, true);
            }
            tc::umma_commit(&s.d_empty[stage]);
          }
          __syncwarp();
          if (++stage == pr.n_stages) stage = 0, d_phase ^= 1;
        }
      }
      if (tc::elect_one()) {
        tc::umma_commit(&s.acc_full[buf]);
        if (u + 1 == units) tc::umma_commit(&s.q_empty[b]);
      }
      __syncwarp();
      SIM3_TR(1, 15);
    }
  }
}

// ---- pooling side: one 32-column chunk of finished cosines of this lane's query row ------------------------------------------
// Loads TMEM columns [col0, col0 + 32) of accumulator `buf` (lane quarter `buf`), applies the exact-match rules of
// simtile.cuh::store_sim_tile (identical OOV ids: +1; identical in-vocabulary ids: snap to exactly 1.0) and zeroes the columns
// past the unit's live tokens (stale TMEM).  did: the 32 token ids of these columns in shared memory.
__device__ __forceinline__ void load_cosines(uint32_t tmem_base, int buf, int col0, int live, int qi, const int* did, float (&v)[32]) {
  tc::tmem_ld_32x32(tmem_base + ((uint32_t)(buf * 32) << 16) + (uint32_t)(buf * U_DOCS + col0), v);
  tc::tmem_ld_wait();
  if (col0 + 32 > live) {  // warp-uniform: ragged tail of the unit
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (col0 + j < live) ? v[j] : 0.f;
  }
  if (qi != 0) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const int4 d4 = *reinterpret_cast<const int4*>(did + 4 * j4);
      const int dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool same = dd[j] == qi;
        float x = v[4 * j4 + j];
        x = (same && qi < 0) ? x + 1.0f : x;
        x = (same && qi > 0 && x > 0.5f) ? 1.0f : x;
        v[4 * j4 + j] = x;
      }
    }
  }
}

// count (uint16) -> float without an I2F (conversions share the XU pipe with ex2)
__device__ __forceinline__ float count_as_float(unsigned c) { return __uint_as_float(0x4B000000u | c) - 8388608.0f; }

}  // namespace simtc3
}  // namespace capr
