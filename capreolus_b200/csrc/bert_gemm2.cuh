// K5b -- the encoder GEMM on CTA PAIRS: tcgen05.mma.cta_group::2, 256 x 256 output tiles per cluster of two CTAs.
//
// Why: the ncu capture of gemm_kernel<3> (profiles/r01_v6_gemm_kernel_ncu_full.json) shows the 1-CTA kernel bound by
// the L2 -> SM fabric, not by the tensor pipe: every 128 x 256 tile streams (128 + 256) rows x 2 planes per K block =
// 96 KB per 1 536 cycles of MMA work = 62.5 B/clk/SM, the chip delivers 10.8 TB/s (~47 B/clk/SM) and the tensor pipe
// idles 26 % of the time.  With cta_group::2 the pair shares one 256-wide B tile: each CTA loads its own 128 rows of A and
// only HALF of B (128 weight rows), the MMA reads both halves (M = 256 over the two CTAs' A tiles, N = 256), so a CTA
// ingests 64 KB per 1 536 cycles = 41.7 B/clk -- under what the fabric delivers -- and has room for 3 pipeline stages.
//
// Protocol (CTA 0 of the pair = leader; same shared-memory layout in both CTAs):
//   TMA producer (warp 0 of BOTH CTAs)  cp.async.bulk.tensor ... cta_group::2 into the CTA's own shared memory, completing
//                                       transaction bytes on the LEADER's `full` barrier (barrier address with the peer bit
//                                       cleared); the leader's producer posts expect_tx for both CTAs' bytes.
//   MMA issuer (warp 1 of the leader)   waits `full`, issues tcgen05.mma.cta_group::2, frees the stage in BOTH CTAs with
//                                       tcgen05.commit ... multicast::cluster (mask 0b11), publishes accumulators likewise.
//   epilogue (warps 4-11 of both CTAs)  each CTA drains ITS 128 accumulator rows (its own TMEM): two warps per TMEM lane
//                                       quarter, one per 128-column half of the tile (the erf-GELU + bf16 split of FFN1 costs
//                                       ~45 instructions per element: with one warp per scheduler the epilogue of a K = 768
//                                       tile outlasted its 18.4 k cycles of MMAs, launch list r01_v10: 784 us vs 606 us for
//                                       FFN2); all 16 warps of the pair arrive on the leader's acc_empty barrier (remote
//                                       mbarrier.arrive for the peer).
//   TMEM                                tcgen05.alloc.cta_group::2 by warp 2 of both CTAs; cluster barriers bracket the kernel.
#pragma once
#include "bert_gemm.cuh"

namespace capr {
namespace bert {

constexpr int G2_THREADS = 384;                  // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int G2_BN = 256;                       // columns per pair tile (UMMA N); each CTA holds half of the B tile
constexpr int G2_B_HALF_BYTES = (G2_BN / 2) * BK * 2;  // 16 KB
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA

template <int MODE>
struct Gemm2Smem {
  static constexpr int PLANES = MODE == 3 ? 2 : 1;
  static constexpr int STAGE_BYTES = PLANES * (A_TILE_BYTES + G2_B_HALF_BYTES);  // 64 KB (MODE 3) / 32 KB (MODE 1)
  static constexpr int STAGES = MODE == 3 ? 3 : 6;
  static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

namespace tc2 {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// box -> this CTA's shared memory; transaction bytes -> the LEADER CTA's barrier at the same offset
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(map), "r"(tc::smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrive on `bar` (same offset) in BOTH CTAs once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tc::smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
}  // namespace tc2

// tm_b_*: maps of the weight planes with a {64, 128} box (half of the pair's B tile per CTA).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
             const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const GemmArgs g) {
  using S = Gemm2Smem<MODE>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* full = bars;                           // [STAGES]  TMA (both CTAs) -> MMA; used in the leader only
  uint64_t* empty = bars + S::STAGES;              // [STAGES]  MMA -> TMA, one copy per CTA (multicast commit)
  uint64_t* acc_full = bars + 2 * S::STAGES;       // [2]  MMA -> epilogue, one copy per CTA (multicast commit)
  uint64_t* acc_empty = bars + 2 * S::STAGES + 2;  // [2]  epilogues of both CTAs -> MMA; used in the leader only
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_n = g.N / G2_BN;
  const int tiles_m = (g.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = tiles_m * tiles_n;
  const int k_blocks = g.K / BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(&tm_a_hi);
    tc::prefetch_tensormap(&tm_b_hi);
    if (MODE == 3) {
      tc::prefetch_tensormap(&tm_a_lo);
      tc::prefetch_tensormap(&tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      tc::mbar_init(&full[s], 1);   // the leader producer's expect_tx arrive (+ the transaction bytes of both CTAs)
      tc::mbar_init(&empty[s], 1);  // one multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 16);  // 8 epilogue warps in each CTA of the pair
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc2::tmem_alloc2(tmem_base_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc2::cluster_sync();  // both CTAs' barriers are initialised before any remote arrive / transaction can reach them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const uint32_t stage_tx = 2u * (uint32_t)S::STAGE_BYTES;  // bytes of both CTAs land on the leader's barrier
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < n_tiles; t += n_pairs) {
      const int m0 = (t / tiles_n) * 2 * BM + (int)rank * BM;              // this CTA's 128 rows of A
      const int n0 = (t % tiles_n) * G2_BN + (int)rank * (G2_BN / 2);      // this CTA's half of the B tile
      for (int kb = 0; kb < k_blocks; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        unsigned char* st = smem + stage * S::STAGE_BYTES;
        if (tc::elect_one()) {
          if (rank == 0) tc::mbar_expect_tx(&full[stage], stage_tx);
          tc2::tma_load_2d_pair(st, &tm_a_hi, &full[stage], kb * BK, m0);
          tc2::tma_load_2d_pair(st + A_TILE_BYTES, &tm_b_hi, &full[stage], kb * BK, n0);
          if (MODE == 3) {
            tc2::tma_load_2d_pair(st + A_TILE_BYTES + G2_B_HALF_BYTES, &tm_a_lo, &full[stage], kb * BK, m0);
            tc2::tma_load_2d_pair(st + 2 * A_TILE_BYTES + G2_B_HALF_BYTES, &tm_b_lo, &full[stage], kb * BK, n0);
          }
        }
        __syncwarp();
        if (++stage == S::STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc = tc::make_instr_desc(tc::FMT_BF16, 2 * BM, G2_BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair; t < n_tiles; t += n_pairs) {
      tc::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G2_BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        const uint32_t st = tc::smem_u32(smem + stage * S::STAGE_BYTES);
        const uint64_t a_hi = tc::make_sw128_kmajor_desc(st);
        const uint64_t b_hi = tc::make_sw128_kmajor_desc(st + A_TILE_BYTES);
        const uint64_t a_lo = tc::make_sw128_kmajor_desc(st + A_TILE_BYTES + G2_B_HALF_BYTES);
        const uint64_t b_lo = tc::make_sw128_kmajor_desc(st + 2 * A_TILE_BYTES + G2_B_HALF_BYTES);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per K=16 step, in 16-byte units
            tc2::umma2_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, (kb | k) != 0);
            if (MODE == 3) {
              tc2::umma2_f16(d_tmem, a_lo + koff, b_hi + koff, idesc, true);
              tc2::umma2_f16(d_tmem, a_hi + koff, b_lo + koff, idesc, true);
            }
          }
          tc2::umma2_commit_both(&empty[stage]);  // the stage is reusable in both CTAs once these MMAs have read it
        }
        __syncwarp();
        if (++stage == S::STAGES) stage = 0, phase ^= 1;
      }
      if (tc::elect_one()) tc2::umma2_commit_both(&acc_full[acc]);  // accumulator complete (both CTAs' halves)
      __syncwarp();
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: each drains its own 128 rows) =====================
    const int quarter = warp & 3;
    const int c_begin = ((warp - 4) >> 2) * (G2_BN / 2);  // this warp's 128-column half of the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair; t < n_tiles; t += n_pairs) {
      const int m0 = (t / tiles_n) * 2 * BM + (int)rank * BM, n0 = (t % tiles_n) * G2_BN;
      const int row = m0 + quarter * 32 + lane;
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t)(acc * G2_BN) + ((uint32_t)(quarter * 32) << 16);
      for (int c = c_begin; c < c_begin + G2_BN / 2; c += 32) {
        float v[32];
        tc::tmem_ld_32x32(t_row + (uint32_t)c, v);
        tc::tmem_ld_wait();
        if (row < g.M) {
          const int col = n0 + c;
          const float4* b4 = reinterpret_cast<const float4*>(g.bias + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            v[4 * i + 0] += b.x, v[4 * i + 1] += b.y, v[4 * i + 2] += b.z, v[4 * i + 3] += b.w;
          }
          const size_t off = (size_t)row * g.N + col;
          if (g.epi == EPI_BIAS_GELU_SPLIT || g.epi == EPI_BIAS_SPLIT) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float x0 = v[2 * i], x1 = v[2 * i + 1];
              if (g.epi == EPI_BIAS_GELU_SPLIT) {
                x0 = gelu_erf(x0);  // erf GELU (HF "gelu")
                x1 = gelu_erf(x1);
              }
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(x0, h0, l0);
              split_bf16(x1, h1, l1);
              hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            uint4* ph = reinterpret_cast<uint4*>(g.out_hi + off);
            uint4* pl = reinterpret_cast<uint4*>(g.out_lo + off);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ph[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              if (MODE == 3 || g.epi == EPI_BIAS_SPLIT) pl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          } else {
            if (g.epi == EPI_BIAS_RESID_F32) {
              const float4* r4 = reinterpret_cast<const float4*>(g.resid + off);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 r = r4[i];
                v[4 * i + 0] += r.x, v[4 * i + 1] += r.y, v[4 * i + 2] += r.z, v[4 * i + 3] += r.w;
              }
            }
            float4* o4 = reinterpret_cast<float4*>(g.out_f32 + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc2::mbar_arrive_leader(&acc_empty[acc]);
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
  }
  // teardown: nobody may exit (or free TMEM) while the peer can still signal its barriers or the pair's MMAs are in flight
  tc::tc_fence_before();
  __syncthreads();
  tc2::cluster_sync();
  if (warp == 2) {
    tc::tc_fence_after();
    tc2::tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace bert
}  // namespace capr
