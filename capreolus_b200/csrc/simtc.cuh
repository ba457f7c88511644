// Tensor-core producer of the query x doc cosine tile (engine 2 of the KNRM-family kernels).
//
//   SimilarityMatrix.forward      capreolus/reranker/common.py:170-182   (the a_emb.bmm(b_emb^T) of l.165)
//
// The cosine tile of one pair is a skinny GEMM: 512 doc rows x 32 query rows x E=300.  fp32 FFMA makes it the
// limiter of the whole kernel (profiles/r01_v1_*: FMA pipe 41 %, issue-bound).  Here it runs on tcgen05 instead:
//   * the prepared table is stored as two bf16 planes, hi = bf16(e), lo = bf16(e - hi) (16 mantissa bits together),
//     row pitch padded to a multiple of 64 elements, so a gathered row is the same 1.2 KB as in fp32;
//   * docs are the M=128 operand (4 M-tiles per 512-doc tile), the query block is the N operand: the B tile stacks
//     [q_hi (32 rows); q_lo (32 rows)], so one MMA with N=64 yields d_hi.q_hi (columns 0-31) and d_hi.q_lo (columns
//     32-63), and a second MMA with N=32 adds d_lo.q_hi into columns 0-31.  cos = col[i] + col[32+i]: the three
//     products of the (hi+lo)(hi+lo) expansion, the dropped lo.lo term is ~2^-18 relative.  CPU emulation of this
//     arithmetic against the goldens: KNRM 8e-7, PACRR 2e-5, DRMM 0 bin flips (tests/emulate.py);
//   * rows are gathered by the TMA unit (cp.async.bulk.tensor ... tile::gather4: four table rows per instruction)
//     straight into the canonical SWIZZLE_128B K-major layout and completed on mbarriers -- the issuing warp never
//     waits for data.  (A first version used 16-byte cp.async + fence.proxy.async: the fence drains every outstanding
//     copy of the thread, which serialised the ring to one L2 round trip per stage: 7.4 M pairs/s ceiling.)
//   * accumulators live in TMEM (4 M-tiles x 64 columns per pair, double buffered = 512 columns), so the epilogue
//     of pair p (TMEM -> cosine tile in smem -> model-specific pooling) overlaps the gather + MMAs of pair p+1.
//
// Warp roles (320 threads): warps 0-7 epilogue (warp % 4 = the TMEM lane quarter it may read), warp 8 = TMA gather
// producer, warp 9 = MMA issuer + TMEM allocator.
#pragma once
#include "simtile.cuh"
#include "tc_common.cuh"

namespace capr {
namespace simtc {

constexpr int EPI_WARPS = 8, PROD_WARPS = 1;
constexpr int EPI_THREADS = EPI_WARPS * 32, PROD_THREADS = PROD_WARPS * 32;
constexpr int THREADS = EPI_THREADS + PROD_THREADS + 32;
constexpr int ATOM_K = 64;                        // bf16 elements per 128-byte swizzle row
constexpr int MAX_ATOMS = 5;                      // pitch <= 320
constexpr int MT = 128;                           // docs per M tile
constexpr int Q_ATOM_BYTES = 64 * 128;            // [q_hi;q_lo] 64 rows x 128 B
constexpr int D_STAGE_BYTES = MT * 128;           // one plane (hi or lo) of 128 doc rows x one 64-element K atom = 16 KB
constexpr int D_STAGES = 4;                       // ring depth
constexpr int ACC_COLS_PER_MT = 64, ACC_COLS_PER_PAIR = 256;

struct Smem {
  unsigned char* q[2];        // [atoms][64 rows][128 B]
  unsigned char* d[D_STAGES];
  float* sim;                 // [SIM_ROWS][SIM_PITCH]
  int* qrow;                  // [2][QT]   table rows of the pair being gathered (producer)
  int* drow;                  // [2][DT]
  int* qid;                   // [QT]      ids of the pair being drained (epilogue)
  uint64_t *q_full, *q_empty, *d_full, *d_empty, *acc_full, *acc_empty;
  uint32_t* tmem_slot;
  float* extra;               // model-specific scratch
};

__host__ __device__ inline size_t smem_bytes(int atoms, size_t extra_bytes) {
  return 1024 + (size_t)2 * atoms * Q_ATOM_BYTES + (size_t)D_STAGES * D_STAGE_BYTES + (size_t)SIM_ROWS * SIM_PITCH * 4 +
         (size_t)(2 * QT + 2 * DT + QT) * 4 + 16 * 8 + 16 + extra_bytes;
}

__device__ __forceinline__ Smem carve(unsigned char* raw, int atoms) {
  Smem s;
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  s.q[0] = p;
  s.q[1] = p + atoms * Q_ATOM_BYTES;
  p += 2 * atoms * Q_ATOM_BYTES;
  for (int i = 0; i < D_STAGES; ++i) s.d[i] = p + i * D_STAGE_BYTES;
  p += D_STAGES * D_STAGE_BYTES;
  s.sim = reinterpret_cast<float*>(p);
  p += SIM_ROWS * SIM_PITCH * 4;
  s.qrow = reinterpret_cast<int*>(p);
  s.drow = s.qrow + 2 * QT;
  s.qid = s.drow + 2 * DT;
  p += (2 * QT + 2 * DT + QT) * 4;
  uint64_t* b = reinterpret_cast<uint64_t*>(p);
  s.q_full = b, s.q_empty = b + 2, s.d_full = b + 4, s.d_empty = b + 4 + D_STAGES, s.acc_full = b + 4 + 2 * D_STAGES,
  s.acc_empty = b + 6 + 2 * D_STAGES;
  p += 16 * 8;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  s.extra = reinterpret_cast<float*>(p + 16);
  return s;
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

struct Problem {
  const long long* q;
  const long long* d;
  int B, Q, D, V;
  const __nv_bfloat16* hi;  // [V][pitch]
  const __nv_bfloat16* lo;
  int pitch;                // elements, multiple of 64
  int E;
};

// Common prologue: barriers + TMEM.  Call from all threads; returns the TMEM base.
__device__ __forceinline__ uint32_t setup(const Smem& s, int tid) {
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s.q_full[i], 1);
      tc::mbar_init(&s.q_empty[i], 1);
      tc::mbar_init(&s.acc_full[i], 1);
      tc::mbar_init(&s.acc_empty[i], EPI_WARPS);
    }
    for (int i = 0; i < D_STAGES; ++i) {
      tc::mbar_init(&s.d_full[i], 1);
      tc::mbar_init(&s.d_empty[i], 1);
    }
    tc::fence_barrier_init();
  }
  for (int i = tid; i < SIM_ROWS * SIM_PITCH; i += THREADS) s.sim[i] = 0.f;
  if (warp == EPI_WARPS + PROD_WARPS) tc::tmem_alloc(s.tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  return *s.tmem_slot;
}

__device__ __forceinline__ void teardown(const Smem& s, uint32_t tmem_base, int tid) {
  tc::tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == EPI_WARPS + PROD_WARPS) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---- producer warp: gather the query block and the doc stages of every pair of this CTA -----------------------------
// One warp.  Work items, in order, per pair: the Q block, then for every (M tile, K atom): hi plane, lo plane.  An item
// of 128 doc rows is 32 cp.async.bulk.tensor tile::gather4 instructions -- one per lane, 4 table rows each -- that the
// TMA unit writes straight into the SWIZZLE_128B operand layout and completes on the stage's mbarrier (async proxy:
// no fences, no waiting in the issuing thread).  All D_STAGES stages can be in flight at once.
__device__ __forceinline__ void producer_loop(const Smem& s, const Problem& pr, const CUtensorMap* tm_hi, const CUtensorMap* tm_lo, int lane) {
  const int atoms = pr.pitch / ATOM_K;
  const int n_mt = (pr.D + MT - 1) / MT;
  uint32_t q_phase[2] = {0, 0}, d_phase = 0;
  int d_stage = 0, it = 0;
  for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
    const int b = it & 1;
    int* qrow = s.qrow + b * QT;
    int* drow = s.drow + b * DT;
    qrow[lane] = table_row(lane < pr.Q ? pr.q[(size_t)pair * pr.Q + lane] : 0, pr.V);
    for (int i = lane; i < DT; i += 32) drow[i] = table_row(i < pr.D ? pr.d[(size_t)pair * pr.D + i] : 0, pr.V);
    __syncwarp();
    // query block: per atom a 64-row tile, rows 0-31 = hi plane, rows 32-63 = lo plane of the 32 query tokens
    tc::mbar_wait(&s.q_empty[b], q_phase[b] ^ 1);
    q_phase[b] ^= 1;
    if (lane == 0) tc::mbar_expect_tx(&s.q_full[b], (uint32_t)(atoms * Q_ATOM_BYTES));
    __syncwarp();
    if (lane < 16) {
      const int plane = lane >> 3, g = lane & 7;
      const CUtensorMap* tm = plane ? tm_lo : tm_hi;
      const int r0 = qrow[4 * g], r1 = qrow[4 * g + 1], r2 = qrow[4 * g + 2], r3 = qrow[4 * g + 3];
      for (int a = 0; a < atoms; ++a)
        tc::tma_gather4(s.q[b] + a * Q_ATOM_BYTES + (plane * 32 + 4 * g) * 128, tm, &s.q_full[b], a * ATOM_K, r0, r1, r2, r3);
    }
    for (int mt = 0; mt < n_mt; ++mt) {
      const int* rows = drow + mt * MT + 4 * lane;
      const int r0 = rows[0], r1 = rows[1], r2 = rows[2], r3 = rows[3];
      for (int a = 0; a < atoms; ++a) {
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          tc::mbar_wait(&s.d_empty[d_stage], d_phase ^ 1);
          if (lane == 0) tc::mbar_expect_tx(&s.d_full[d_stage], (uint32_t)D_STAGE_BYTES);
          __syncwarp();
          tc::tma_gather4(s.d[d_stage] + 4 * lane * 128, plane ? tm_lo : tm_hi, &s.d_full[d_stage], a * ATOM_K, r0, r1, r2, r3);
          if (++d_stage == D_STAGES) d_stage = 0, d_phase ^= 1;
        }
      }
    }
  }
}

// ---- MMA issuer (one thread) ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_loop(const Smem& s, const Problem& pr, uint32_t tmem_base) {
  const int atoms = pr.pitch / ATOM_K;
  const int n_mt = (pr.D + MT - 1) / MT;
  const uint32_t idesc64 = tc::make_instr_desc(tc::FMT_BF16, MT, 64);
  const uint32_t idesc32 = tc::make_instr_desc(tc::FMT_BF16, MT, 32);
  uint32_t q_phase[2] = {0, 0}, acc_phase[2] = {0, 0}, d_phase = 0;
  int d_stage = 0, it = 0;
  for (int pair = blockIdx.x; pair < pr.B; pair += gridDim.x, ++it) {
    const int b = it & 1;
    tc::mbar_wait(&s.acc_empty[b], acc_phase[b] ^ 1);
    acc_phase[b] ^= 1;
    tc::mbar_wait(&s.q_full[b], q_phase[b]);
    q_phase[b] ^= 1;
    tc::tc_fence_after();
    const uint32_t qaddr = tc::smem_u32(s.q[b]);
    for (int mt = 0; mt < n_mt; ++mt) {
      const uint32_t d_tmem = tmem_base + (uint32_t)(b * ACC_COLS_PER_PAIR + mt * ACC_COLS_PER_MT);
      for (int a = 0; a < atoms; ++a) {
        const uint64_t bq = tc::make_sw128_kmajor_desc(qaddr + a * Q_ATOM_BYTES);
        const int ksteps = min(ATOM_K, pr.E - a * ATOM_K + 15) / 16;  // skip the all-zero tail of the last atom
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          tc::mbar_wait(&s.d_full[d_stage], d_phase);
          tc::tc_fence_after();
          const uint64_t ad = tc::make_sw128_kmajor_desc(tc::smem_u32(s.d[d_stage]));
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t koff = (uint64_t)(k * 2);  // 32 bytes per K=16 step, in 16-byte units
            if (plane == 0) tc::umma_f16(d_tmem, ad + koff, bq + koff, idesc64, (a | k) != 0);  // [d_hi.q_hi | d_hi.q_lo]
            else tc::umma_f16(d_tmem, ad + koff, bq + koff, idesc32, true);                      //  += d_lo.q_hi
          }
          tc::umma_commit(&s.d_empty[d_stage]);
          if (++d_stage == D_STAGES) d_stage = 0, d_phase ^= 1;
        }
      }
    }
    tc::umma_commit(&s.q_empty[b]);
    tc::umma_commit(&s.acc_full[b]);
  }
}

// ---- epilogue helper: drain the accumulators of one pair into s.sim -------------------------------------------------
// Called by the 256 epilogue threads.  After it returns (it ends with epi_barrier) s.sim holds the cosine tile.
__device__ __forceinline__ void drain_pair(const Smem& s, const Problem& pr, uint32_t tmem_base, int pair, int b, uint32_t acc_parity,
                                           int etid, bool skip_stores = false) {
  const int warp = etid >> 5, lane = etid & 31, quarter = warp & 3, half = warp >> 2;
  const int n_mt = (pr.D + MT - 1) / MT;
  if (etid < QT) s.qid[etid] = id_as_int(etid < pr.Q ? pr.q[(size_t)pair * pr.Q + etid] : 0);
  epi_barrier();
  tc::mbar_wait(&s.acc_full[b], acc_parity);
  tc::tc_fence_after();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int mt = half * 2 + h;
    if (mt < n_mt && !skip_stores) {
      const int doc = mt * MT + quarter * 32 + lane;
      const int did = id_as_int(doc < pr.D ? pr.d[(size_t)pair * pr.D + doc] : 0);
      const uint32_t taddr = tmem_base + (uint32_t)(b * ACC_COLS_PER_PAIR + mt * ACC_COLS_PER_MT) + ((uint32_t)(quarter * 32) << 16);
      float hh[32], hl[32];
      tc::tmem_ld_32x32(taddr, hh);
      tc::tmem_ld_32x32(taddr + 32, hl);
      tc::tmem_ld_wait();
      // exact-match rules of simtile.cuh::store_sim_tile, branch-free: a doc token matches at most the few query
      // positions that hold the same id, so test the cheap "any match" first (warp-uniform skip in the common case)
      bool any = false;
#pragma unroll
      for (int i = 0; i < 32; ++i) any |= (s.qid[i] == did);
      any = any && did != 0;
      if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = hh[i] + hl[i];
          const int qi = s.qid[i];
          const bool same = qi == did;
          v = (same && qi < 0) ? v + 1.0f : v;
          v = (same && qi > 0 && v > 0.5f) ? 1.0f : v;
          s.sim[i * SIM_PITCH + doc] = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) s.sim[i * SIM_PITCH + doc] = hh[i] + hl[i];
      }
    }
  }
  tc::tc_fence_before();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(&s.acc_empty[b]);
  epi_barrier();
}

}  // namespace simtc
}  // namespace capr
