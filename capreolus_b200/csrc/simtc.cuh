// Tensor-core producer of the query x doc cosine tile (engine 2 of the KNRM-family kernels).
//
//   SimilarityMatrix.forward      capreolus/reranker/common.py:170-182   (the a_emb.bmm(b_emb^T) of l.165)
//
// The cosine tile of one pair is a skinny GEMM: 32 query rows x 512 doc rows x E=300.  fp32 FFMA makes it the
// limiter of the whole kernel (profiles/r01_v1_*: FMA pipe 41 %, issue-bound).  Here it runs on tcgen05 instead:
//   * the prepared table is stored as two bf16 planes, hi = bf16(e), lo = bf16(e - hi) (16 mantissa bits together),
//     row pitch = E rounded up to 16 elements, so a gathered row is the same 1.2 KB as in fp32;
//   * the QUERY block is the M operand: A = [q_hi (rows 0-31); q_lo (rows 32-63)] (the MMA is issued with M = 128; rows
//     64-127 read whatever follows in shared memory and produce accumulator rows nobody reads), 256 DOC rows are the N
//     operand.  Per K step two MMAs: B = d_hi gives rows 0-31 = q_hi.d_hi and rows 32-63 = q_lo.d_hi, B = d_lo adds
//     q_hi.d_lo to rows 0-31 (and the negligible q_lo.d_lo to rows 32-63).  cos[i][d] = D[i][d] + D[32+i][d].
//     (The first orientation of this kernel, docs as M=128 and the query block as N=64/32, needed 152 MMAs per pair at
//     ~61 cycles each -- a tcgen05.mma costs max(~60, N/2) cycles, scripts/mma_bench.py -- and was MMA-chain bound at
//     11 M pairs/s; N=256 needs 76 MMAs of 128 cycles that run the pipe at its full rate.)  CPU emulation of the three-product arithmetic against the
//     goldens: KNRM 8e-7, PACRR 2e-5, DRMM 0 bin flips (tests/emulate.py);
//   * rows are gathered with 16-byte cp.async straight into the canonical SWIZZLE_128B K-major layout (8 lanes fetch
//     one 128-byte row segment: fully coalesced) and completed on mbarriers with cp.async.mbarrier.arrive.noinc, so
//     the issuing threads never wait for data (see producer_loop for the two alternatives that were measured);
//   * accumulators live in TMEM: one buffer = 256 doc columns, two buffers, work unit = (pair, half of the doc tile), so
//     the drain of unit u (TMEM -> cosine tile in smem) and the model-specific pooling overlap gather + MMA of u+1.
//
// Warp roles: warps 0-7 epilogue (warps 0/4 read the q_hi rows = TMEM lanes 0-31, warps 1/5 the q_lo rows = lanes
// 32-63), warps 8-11 gather producers, warp 12 = MMA issuer + TMEM allocator (416 threads: PACRR, whose conv epilogue
// needs the whole tile with its halo and runs drain -> conv sequentially on warps 0-7).
// KNRM and DRMM pool column-additively, so they use the PIPELINED epilogue (544 threads): warps 0,1,4,5 only drain, into
// two half-tile buffers of 32 x 256 cosines, and 8 pooling warps consume them -- drain of half h+1 overlaps pooling of
// half h, across pair boundaries.  A warp runs on sub-partition warp % 4 and pooling is MUFU / issue bound, so the roles
// are laid out two pooling warps per sub-partition: pooling = 2,3,6,7,8,9,12,13; producers = 10,11,14,15; MMA = 16.
#pragma once
#include "simtile.cuh"
#include "tc_common.cuh"

namespace capr {
namespace simtc {

constexpr int EPI_WARPS = 8, PROD_WARPS = 4;
constexpr int EPI_THREADS = EPI_WARPS * 32, PROD_THREADS = PROD_WARPS * 32;
constexpr int THREADS = EPI_THREADS + PROD_THREADS + 32;
constexpr int ATOM_K = 64;                    // bf16 elements per 128-byte swizzle row
constexpr int MAX_ATOMS = 5;                  // pitch <= 320
constexpr int NT_DOCS = 256;                  // doc rows per MMA (N) = per work unit
constexpr int Q_ATOM_BYTES = 64 * 128;        // [q_hi;q_lo] 64 rows x 128 B
constexpr int D_STAGE_BYTES = NT_DOCS * 128;  // one plane (hi or lo) of 256 doc rows x one 64-element K atom = 32 KB
constexpr int D_STAGES = 2;                   // ring depth of the default layout
constexpr int DEEP_DCAP = 1024;               // "deep" layout: doc-id capacity per pair (maxdoclen up to 1024 = four 256-doc units)
constexpr int MAX_D_STAGES = 3;               // "deep" layout: one query buffer, three doc stages (96 KB of gathers in flight)
constexpr int ACC_COLS = NT_DOCS;             // TMEM columns per accumulator buffer
constexpr size_t MAX_DYN_SMEM = 232448;       // 227 KB
constexpr int POOL_WARPS = 8;                 // pipelined epilogue: pooling warps
#ifndef CAPR_MMA_WARP_PIPE
#define CAPR_MMA_WARP_PIPE 16
#endif
constexpr int MMA_WARP_PIPE = CAPR_MMA_WARP_PIPE;  // pipelined layout: 16 (scheduler 0, next to two drain and two pooling warps) or, A/B, 18
                                                   // (scheduler 2, next to two producer and two pooling warps; warps 16 and 17 then idle)
constexpr int THREADS_PIPE = (MMA_WARP_PIPE + 1) * 32;  // 17 warps
constexpr int MMA_WARP = EPI_WARPS + PROD_WARPS;  // sequential layout (PACRR)
constexpr int HALF_PITCH = NT_DOCS + 4;       // 260 floats: one-row-per-lane float4 stores are conflict-free (260 % 32 == 4)
constexpr int HALF_FLOATS = QT * HALF_PITCH;  // one half tile = 33 280 B; two of them fit in the full-tile region
constexpr int SPARE_FLOATS = SIM_ROWS * SIM_PITCH - 2 * HALF_FLOATS;  // 1 936 floats left over there (DRMM counters)

struct Smem {
  unsigned char* q0;          // 2 x [atoms][64 rows][128 B]
  int q_stride;
  unsigned char* d0;          // D_STAGES x 32 KB
  __device__ __forceinline__ unsigned char* qbuf(int b) const { return q0 + b * q_stride; }
  __device__ __forceinline__ unsigned char* dbuf(int i) const { return d0 + i * D_STAGE_BYTES; }
  float* sim;                 // [SIM_ROWS][SIM_PITCH]
  int* qrow;                  // [QT]   table rows of the pair being gathered (producer)
  int dcap;                   // doc ids per pair the id arrays below can hold (DT, or DEEP_DCAP in the deep layout)
  int* drow;                  // [dcap]
  int* qid;                   // [2][QT]   ids of the pair being drained / pooled (epilogue; pipelined mode: by pair parity)
  int* did;                   // [2][dcap]
  uint64_t *q_full, *q_empty, *d_full, *d_empty, *acc_full, *acc_empty, *half_full, *half_empty;
  uint32_t* tmem_slot;
  float* extra;               // model-specific scratch
};

// deep = false: two query buffers + two doc stages (default).  deep = true: ONE query buffer + THREE doc stages -- the gather
// is latency bound (Little: bytes in flight / loaded L2 latency), so a third 32 KB stage buys more than the second query buffer.
__host__ __device__ inline size_t smem_bytes(int atoms, size_t extra_bytes, bool deep = false) {
  return 1024 + (size_t)(deep ? 1 : 2) * atoms * Q_ATOM_BYTES + (size_t)(deep ? MAX_D_STAGES : D_STAGES) * D_STAGE_BYTES +
         (size_t)SIM_ROWS * SIM_PITCH * 4 + (size_t)(3 * QT + 3 * (deep ? DEEP_DCAP : DT)) * 4 + 18 * 8 + 16 + extra_bytes;
}

__device__ __forceinline__ Smem carve(unsigned char* raw, int atoms, bool deep = false) {
  Smem s;
  // (offset arithmetic instead of rounding the pointer as an integer: the compiler keeps the shared address space and
  // emits LDS/STS instead of generic loads)
  unsigned char* p = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);
  s.q0 = p;
  s.q_stride = atoms * Q_ATOM_BYTES;
  p += (deep ? 1 : 2) * atoms * Q_ATOM_BYTES;
  s.d0 = p;
  p += (deep ? MAX_D_STAGES : D_STAGES) * D_STAGE_BYTES;
  s.sim = reinterpret_cast<float*>(p);
  p += SIM_ROWS * SIM_PITCH * 4;
  s.dcap = deep ? DEEP_DCAP : DT;
  s.qrow = reinterpret_cast<int*>(p);
  s.drow = s.qrow + QT;
  s.qid = s.drow + s.dcap;
  s.did = s.qid + 2 * QT;
  p += (3 * QT + 3 * s.dcap) * 4;
  uint64_t* b = reinterpret_cast<uint64_t*>(p);
  s.q_full = b, s.q_empty = b + 2, s.d_full = b + 4, s.d_empty = b + 4 + MAX_D_STAGES, s.acc_full = b + 4 + 2 * MAX_D_STAGES,
  s.acc_empty = b + 6 + 2 * MAX_D_STAGES;
  s.half_full = b + 8 + 2 * MAX_D_STAGES, s.half_empty = b + 10 + 2 * MAX_D_STAGES;  // 18 barriers in all
  p += 18 * 8;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  s.extra = reinterpret_cast<float*>(p + 16);
  return s;
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }
__device__ __forceinline__ void prod_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(PROD_THREADS) : "memory"); }
__device__ __forceinline__ void drain_barrier() { asm volatile("bar.sync 3, 128;" ::: "memory"); }  // epilogue warps 0,1,4,5

struct Problem {
  const long long* q;
  const long long* d;
  int B, Q, D, V;
  const __nv_bfloat16* hi;  // [V][pitch]
  const __nv_bfloat16* lo;
  int pitch;                // elements per table row, multiple of 16 (the last 64-element K atom may be partial)
  int E;
  int debug;                // CAPR_DEBUG_* profiling switches (0 in production)
  int single;               // 1: use only query buffer 0 and TMEM accumulator buffer 0 (PACRR's conv-on-tensor-cores epilogue
                            // needs the second query buffer's shared memory and half of TMEM for itself)
  int deep;                 // 1: the "deep" shared-memory layout (carve(..., true)): one query buffer, three doc stages
};

__device__ __forceinline__ int ring_depth(const Problem& pr) { return pr.deep ? MAX_D_STAGES : D_STAGES; }
__device__ __forceinline__ bool one_qbuf(const Problem& pr) { return pr.single || pr.deep; }

__device__ __forceinline__ int halves_of(const Problem& pr) { return (pr.D + NT_DOCS - 1) / NT_DOCS; }

// Common prologue: barriers + TMEM.  Call from all threads; returns the TMEM base.
__device__ __forceinline__ uint32_t setup(const Smem& s, int tid, int nthreads = THREADS, int mma_warp = MMA_WARP) {
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s.q_full[i], PROD_THREADS);
      tc::mbar_init(&s.q_empty[i], 1);
      tc::mbar_init(&s.acc_full[i], 1);
      tc::mbar_init(&s.acc_empty[i], 4);  // the four draining warps
      tc::mbar_init(&s.half_full[i], 2);  // the two q_hi draining warps
      tc::mbar_init(&s.half_empty[i], POOL_WARPS);
    }
    for (int i = 0; i < MAX_D_STAGES; ++i) {
      tc::mbar_init(&s.d_full[i], PROD_THREADS);
      tc::mbar_init(&s.d_empty[i], 1);
    }
    tc::fence_barrier_init();
  }
  for (int i = tid; i < SIM_ROWS * SIM_PITCH; i += nthreads) s.sim[i] = 0.f;
  if (warp == mma_warp) tc::tmem_alloc(s.tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  return *s.tmem_slot;
}

__device__ __forceinline__ void teardown(const Smem& s, uint32_t tmem_base, int tid, int mma_warp = MMA_WARP) {
  tc::tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == mma_warp) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---- producer warps: gather the query block and the doc stages of every pair of this CTA ----------------------------
// 4 warps.  Work items, in order, per pair: the Q block, then for every (half, K atom): hi plane, lo plane of 256 doc
// rows.  Rows are copied with 16-byte cp.async straight into the SWIZZLE_128B operand layout (8 lanes fetch one
// 128-byte row segment: fully coalesced) and each thread posts cp.async.mbarrier.arrive.noinc on the stage's barrier,
// which fires when ITS copies have landed -- the issuing thread never waits for data, so every stage is in flight.
// (This is the cp.async -> UMMA hand-off CUTLASS uses in sm100_mma_cpasync_warpspecialized.hpp.  Two alternatives were
// measured and dropped: wait_group + fence.proxy.async + arrive serialises on the fence, 7.4 M pairs/s ceiling; TMA
// tile::gather4 of 128-byte rows costs ~80 cycles per instruction, 2.6 M pairs/s.)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <bool IDLE_SLEEP = false>
__device__ __forceinline__ void producer_loop(const Smem& s, const Problem& pr, int ptid /*0..127*/) {
  const int atoms = (pr.pitch + ATOM_K - 1) / ATOM_K;
  const int last_chunks = (pr.pitch - (atoms - 1) * ATOM_K) / 8;  // 16-byte chunks that exist in the last atom
  const int halves = halves_of(pr);
  const int sub = ptid & 7;    // 16-byte chunk inside the 128-byte row segment
  const int rsub = ptid >> 3;  // 0..15: this thread serves rows rsub + 16*j
  uint32_t q_phase = 0, d_phase = 0;  // q_phase: one bit per buffer
  int d_stage = 0, it = 0;
  const int nst = ring_depth(pr);
  // The ids of the NEXT pair are fetched into registers while this pair's gathers are being issued: the ids stream from
  // HBM (one use each), and a load -> barrier -> first gather chain at the top of every pair would leave the ring
  // draining for a DRAM round trip.
  constexpr int IDS_PER_THREAD = DEEP_DCAP / PROD_THREADS;  // 8 doc ids per producer thread at most
  long long q_next = 0, d_next[IDS_PER_THREAD];
  auto fetch_ids = [&](int pair) {
    const bool have = pair < pr.B;
    q_next = (have && ptid < pr.Q) ? pr.q[(size_t)pair * pr.Q + ptid] : 0;  // ptid < Q <= QT
#pragma unroll
    for (int j = 0; j < IDS_PER_THREAD; ++j) {
      const int i = ptid + PROD_THREADS * j;
      d_next[j] = (have && i < pr.D) ? pr.d[(size_t)pair * pr.D + i] : 0;
    }
  };
  fetch_ids(blockIdx.x);
  for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
    const int b = one_qbuf(pr) ? 0 : (it & 1);
    prod_barrier();  // every producer thread is done reading the previous pair's rows
    if (ptid < QT) s.qrow[ptid] = table_row(q_next, pr.V);
#pragma unroll
    for (int j = 0; j < IDS_PER_THREAD; ++j) {
      const int i = ptid + PROD_THREADS * j;
      if (i < halves * NT_DOCS) s.drow[i] = table_row(d_next[j], pr.V);
    }
    prod_barrier();
    fetch_ids(pair + gridDim.x);
    // query block: per atom a 64-row tile, rows 0-31 = hi plane, rows 32-63 = lo plane of the 32 query tokens
    tc::mbar_wait_idle<IDLE_SLEEP>(&s.q_empty[b], ((q_phase >> b) & 1) ^ 1);
    q_phase ^= 1u << b;
    {
      const uint32_t qbase = tc::smem_u32(s.qbuf(b));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rsub + 16 * j;  // 0..63
        const int trow = s.qrow[r & 31];
        const __nv_bfloat16* src = (r < 32 ? pr.hi : pr.lo) + (size_t)trow * pr.pitch + sub * 8;
        const uint32_t dst = qbase + r * 128 + ((sub ^ (r & 7)) << 4);
        const uint32_t nbytes = trow != 0 ? 16u : 0u;  // <pad> / OOV: zero-fill, no global read (see the doc stages below)
        for (int a = 0; a < atoms; ++a)
          if (a + 1 < atoms || sub < last_chunks)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + a * Q_ATOM_BYTES), "l"(src + a * ATOM_K), "r"(nbytes) : "memory");
      }
    }
    cp_async_arrive_noinc(&s.q_full[b]);
    for (int h = 0; h < halves; ++h) {
      // Rows of <pad> / OOV tokens (table row 0) are not read at all: cp.async with src-size 0 zero-fills the 16 bytes.  Their
      // cosine is exactly 0 whatever emb[0] holds (the reference masks them, common.py:149-153), ragged batches only gather
      // their real tokens, and 148 SMs do not hammer the one L2 line of row 0.
      unsigned off[16];  // element offsets of this thread's 16 rows (V * pitch < 2^31 is checked on the host)
      unsigned live = 0;  // bit j: row j is a real table row
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int trow = s.drow[h * NT_DOCS + rsub + 16 * j];
        off[j] = (unsigned)trow * (unsigned)pr.pitch + (unsigned)(sub * 8);
        live |= (trow != 0 ? 1u : 0u) << j;
      }
      for (int a = 0; a < atoms; ++a) {
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          const __nv_bfloat16* tab = (plane == 0 ? pr.hi : pr.lo) + a * ATOM_K;
          tc::mbar_wait_idle<IDLE_SLEEP>(&s.d_empty[d_stage], d_phase ^ 1);
          const uint32_t base = tc::smem_u32(s.dbuf(d_stage));
          if ((a + 1 < atoms || sub < last_chunks) && !CAPR_DBG(pr.debug & 0x800)) {  // tail chunks of a partial last atom are never read
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int r = rsub + 16 * j;
              asm volatile(CAPR_GATHER_CP " [%0], [%1], 16, %2;" ::"r"(base + r * 128 + ((sub ^ (r & 7)) << 4)), "l"(tab + off[j]),
                           "r"(((live >> j) & 1u) << 4)
                           : "memory");
            }
          }
          cp_async_arrive_noinc(&s.d_full[d_stage]);
          if (++d_stage == nst) d_stage = 0, d_phase ^= 1;
        }
      }
    }
  }
  cp_async_commit();
  cp_async_wait<0>();  // nothing may still be landing in shared memory when the CTA tears down
}

// ---- MMA issuer: the WHOLE warp runs the loop (waits are warp-wide), one elected lane issues ----------------------------
// Keeping the control flow warp-uniform lets ptxas hold descriptors / addresses in uniform registers; issuing from a
// divergent `if (lane == 0)` region instead costs ~150 cycles per tcgen05.mma (R2UR chains + a serialising
// ELECT/BRA.U.ANY loop around every UTCHMMA).
template <bool IDLE_SLEEP = false>
__device__ __forceinline__ void mma_loop(const Smem& s, const Problem& pr, uint32_t tmem_base) {
  const int atoms = (pr.pitch + ATOM_K - 1) / ATOM_K;
  const int halves = halves_of(pr);
  const uint32_t idesc = tc::make_instr_desc(tc::FMT_BF16, 128, NT_DOCS);
  const bool skip = CAPR_DBG(pr.debug & 0x400) != 0;
  uint32_t q_phase = 0, acc_phase = 0, d_phase = 0;  // q/acc: one bit per buffer
  int d_stage = 0, it = 0, unit = 0;
  const int nst = ring_depth(pr);
  for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
    const int b = one_qbuf(pr) ? 0 : (it & 1);
    tc::mbar_wait_idle<IDLE_SLEEP>(&s.q_full[b], (q_phase >> b) & 1);
    q_phase ^= 1u << b;
    const uint64_t q_desc = tc::make_sw128_kmajor_desc(tc::smem_u32(s.qbuf(b)));
    for (int h = 0; h < halves; ++h, ++unit) {
      const int ab = pr.single ? 0 : (unit & 1);
      tc::mbar_wait_idle<IDLE_SLEEP>(&s.acc_empty[ab], ((acc_phase >> ab) & 1) ^ 1);
      acc_phase ^= 1u << ab;
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(ab * ACC_COLS);
      for (int a = 0; a < atoms; ++a) {
        const uint64_t aq = q_desc + (uint64_t)((a * Q_ATOM_BYTES) >> 4);
        const int ksteps = skip ? 0 : min(ATOM_K, pr.pitch - a * ATOM_K) / 16;  // a partial last atom has fewer K steps
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          tc::mbar_wait_idle<IDLE_SLEEP>(&s.d_full[d_stage], d_phase);
          tc::tc_fence_after();
          const uint64_t bd = tc::make_sw128_kmajor_desc(tc::smem_u32(s.dbuf(d_stage)));
          if (tc::elect_one()) {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t koff = (uint64_t)(k * 2);  // 32 bytes per K=16 step, in 16-byte units
              tc::umma_f16(d_tmem, aq + koff, bd + koff, idesc, (a | plane | k) != 0);
            }
            tc::umma_commit(&s.d_empty[d_stage]);
          }
          __syncwarp();
          if (++d_stage == nst) d_stage = 0, d_phase ^= 1;
        }
      }
      if (tc::elect_one()) {
        tc::umma_commit(&s.acc_full[ab]);
        if (h + 1 == halves) tc::umma_commit(&s.q_empty[b]);
      }
      __syncwarp();
    }
  }
}

// ---- epilogue helper: drain the accumulators of one pair into s.sim -------------------------------------------------
// Called by the 256 epilogue threads with the index of the pair's first work unit.  Warps 0 and 4 hold the q_hi rows
// (TMEM lanes 0-31, one query per lane), warps 1 and 5 the q_lo rows (lanes 32-63); warps 0/1 take doc columns 0-127 of
// each half, warps 4/5 columns 128-255.  The lo warp parks its partial products in the tile, the hi warp adds its own,
// applies the exact-match rules of simtile.cuh::store_sim_tile and writes the cosine back.  Row stride 516 floats makes
// the one-row-per-lane 16-byte stores conflict-free.  Ends with epi_barrier: afterwards s.sim holds the whole tile and
// s.qid / s.did the ids of the pair.
__device__ __forceinline__ void drain_pair(const Smem& s, const Problem& pr, uint32_t tmem_base, int pair, int first_unit,
                                           uint32_t& acc_phase /* one bit per buffer */, int etid, bool skip_stores = false) {
  const int warp = etid >> 5, lane = etid & 31;
  const int halves = halves_of(pr);
  if (etid < QT) s.qid[etid] = id_as_int(etid < pr.Q ? pr.q[(size_t)pair * pr.Q + etid] : 0);
  for (int i = etid; i < DT; i += EPI_THREADS) s.did[i] = id_as_int(i < pr.D ? pr.d[(size_t)pair * pr.D + i] : 0);
  epi_barrier();
  if ((warp & 2) == 0) {  // warps 0,1,4,5 drain
    const bool is_lo = (warp & 1) != 0;
    const int col_half = warp >> 2;  // 0: columns 0-127, 1: columns 128-255 of the unit
    const uint32_t lane_off = (uint32_t)((warp & 1) * 32) << 16;
    const int qi = s.qid[lane];
    for (int h = 0; h < halves; ++h) {
      const int ab = pr.single ? 0 : ((first_unit + h) & 1);
      tc::mbar_wait(&s.acc_full[ab], (acc_phase >> ab) & 1);
      tc::tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col = col_half * 128 + cc * 32;  // column inside the unit
        const int doc0 = h * NT_DOCS + col;
        float v[32];
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(ab * ACC_COLS + col), v);
        tc::tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(s.sim + lane * SIM_PITCH + doc0);
        if (is_lo && !skip_stores) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        drain_barrier();  // lo partials are in the tile
        if (!is_lo && !skip_stores) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 l = dst[j];
            v[4 * j] += l.x, v[4 * j + 1] += l.y, v[4 * j + 2] += l.z, v[4 * j + 3] += l.w;
          }
          if (qi != 0) {  // identical ids: OOV exact match (+1) or in-vocabulary snap to 1.0
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool same = s.did[doc0 + j] == qi;
              v[j] = (same && qi < 0) ? v[j] + 1.0f : v[j];
              v[j] = (same && qi > 0 && v[j] > 0.5f) ? 1.0f : v[j];
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s.acc_empty[ab]);
      acc_phase ^= 1u << ab;
    }
  } else {
    for (int h = 0; h < halves; ++h) acc_phase ^= 1u << (pr.single ? 0 : ((first_unit + h) & 1));  // same phase bookkeeping in every warp
  }
  epi_barrier();
}

// ---- pipelined epilogue (KNRM, DRMM) -----------------------------------------------------------------------------------
// warp -> role.  Drain warps must sit on TMEM lane quarters 0 and 1 (warp % 4 in {0,1}); the rest is balanced per
// sub-partition (warp % 4): SMSP0 = drain 0,4 + pool 8,12 + MMA 16; SMSP1 = drain 1,5 + pool 9,13; SMSP2 = pool 2,6 +
// producers 10,14; SMSP3 = pool 3,7 + producers 11,15.
__device__ __forceinline__ bool is_drain_warp(int warp) { return warp < 8 && (warp & 2) == 0; }
__device__ __forceinline__ bool is_pool_warp(int warp) { return warp < 16 && ((warp < 8) == ((warp & 2) != 0)); }
__device__ __forceinline__ bool is_producer_warp(int warp) { return warp >= 8 && warp < 16 && (warp & 2) != 0; }
__device__ __forceinline__ int pool_index(int warp) { return (warp >> 2) * 2 + (warp & 1); }      // 2,3,6,7,8,9,12,13 -> 0..7
__device__ __forceinline__ int producer_index(int warp) { return ((warp - 8) >> 2) * 2 + (warp & 1); }  // 10,11,14,15 -> 0..3
__device__ __forceinline__ float* half_tile(const Smem& s, int hb) { return s.sim + hb * HALF_FLOATS; }
__device__ __forceinline__ float* spare_scratch(const Smem& s) { return s.sim + 2 * HALF_FLOATS; }

// Drain warps (0,1,4,5): all pairs of this CTA.  Half-tile buffer and TMEM buffer of a unit share the unit's parity.
__device__ __forceinline__ void drain_loop(const Smem& s, const Problem& pr, uint32_t tmem_base, int warp, int lane) {
  const int halves = halves_of(pr);
  const bool is_lo = (warp & 1) != 0;
  const int col_half = warp >> 2;
  const int dtid = (warp >> 2) * 64 + (warp & 1) * 32 + lane;  // 0..127 over the four drain warps
  const uint32_t lane_off = (uint32_t)((warp & 1) * 32) << 16;
  const bool skip_stores = CAPR_DBG(pr.debug & 0x200 /*CAPR_DEBUG_SKIP_DRAIN*/) != 0;
  uint32_t full_phase = 0, empty_phase = 0;  // one bit per buffer
  int unit = 0, it = 0;
  for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
    const int pp = it & 1;
    int* qid = s.qid + pp * QT;
    int* did = s.did + pp * s.dcap;
    int qi = 0;
    for (int h = 0; h < halves; ++h, ++unit) {
      const int ub = unit & 1;
      tc::mbar_wait(&s.half_empty[ub], ((empty_phase >> ub) & 1) ^ 1);  // pooling is done with this half-tile buffer
      empty_phase ^= 1u << ub;
      if (h == 0) {
        // ids of this pair.  Buffer pp was last read by the pooling of pair it-2, and that pooling released every one of
        // its half tiles before the wait above could complete (buffer ub was last used by unit-2, which belongs to pair
        // it-1 when halves == 2 and to pair it-2 when halves == 1).  The writes are ordered before half_full by the
        // barrier below + the hi warps' arrive.
        if (dtid < QT) qid[dtid] = id_as_int(dtid < pr.Q ? pr.q[(size_t)pair * pr.Q + dtid] : 0);
        for (int i = dtid; i < halves * NT_DOCS; i += 128) did[i] = id_as_int(i < pr.D ? pr.d[(size_t)pair * pr.D + i] : 0);
        drain_barrier();
        qi = qid[lane];
      }
      tc::mbar_wait(&s.acc_full[ub], (full_phase >> ub) & 1);
      full_phase ^= 1u << ub;
      tc::tc_fence_after();
      float* tile = half_tile(s, ub);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col = col_half * 128 + cc * 32;
        float v[32];
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(ub * ACC_COLS + col), v);
        tc::tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(tile + lane * HALF_PITCH + col);
        if (is_lo && !skip_stores) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        drain_barrier();  // lo partials are in the tile
        if (!is_lo && !skip_stores) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 l = dst[j];
            v[4 * j] += l.x, v[4 * j + 1] += l.y, v[4 * j + 2] += l.z, v[4 * j + 3] += l.w;
          }
          if (qi != 0) {  // identical ids: OOV exact match (+1) or in-vocabulary snap to 1.0 (simtile.cuh rules)
            const int* dd = did + h * NT_DOCS + col;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool same = dd[j] == qi;
              v[j] = (same && qi < 0) ? v[j] + 1.0f : v[j];
              v[j] = (same && qi > 0 && v[j] > 0.5f) ? 1.0f : v[j];
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&s.acc_empty[ub]);
        if (!is_lo) tc::mbar_arrive(&s.half_full[ub]);
      }
    }
  }
}

// Pooling-side handshake: wait for half `ub` of the current unit / hand the buffer back.
struct PoolSync {
  uint32_t phase = 0;  // one bit per buffer
  __device__ __forceinline__ void wait_full(const Smem& s, int ub) {
    tc::mbar_wait(&s.half_full[ub], (phase >> ub) & 1);
    phase ^= 1u << ub;
  }
  __device__ __forceinline__ void release(const Smem& s, int ub, int lane) {
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&s.half_empty[ub]);
  }
};

}  // namespace simtc
}  // namespace capr
