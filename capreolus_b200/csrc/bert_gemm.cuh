// K5 -- tcgen05 GEMM of the BERT encoder:  C[M,N] = A[M,K] . W[N,K]^T (+ bias, + GELU | + residual)
//
//   HF BertSelfAttention / BertSelfOutput / BertIntermediate / BertOutput Linear layers as called by
//   PTBERTMaxP_Class.predict_step -> self.bert(...)            capreolus/reranker/ptBERTMaxP.py:82
//
// A = activations (row = token, K contiguous), W = nn.Linear weight [out,in] (K contiguous): both operands are
// K-major, which is the natural layout for TMA boxes of {64 bf16, rows} with SWIZZLE_128B.
//
// Precision: plain bf16 operands give 2e-2 relative error on the logits, so the parity mode feeds
// every fp32 operand as a (hi, lo) pair of bf16 planes and accumulates hi.hi + lo.hi + hi.lo in the fp32 TMEM
// accumulator (3 tcgen05.mma per K step, ~2^-17 relative operand error).  MODE = 1 is the plain bf16 product.
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles):
//   warp 0   : TMA producer  -- one elected lane, cp.async.bulk.tensor into a ring of smem stages, mbarrier tx
//   warp 1   : MMA issuer    -- one elected lane, tcgen05.mma.cta_group::1.kind::f16 M=128 N=BN K=16, tcgen05.commit
//   warp 2   : TMEM allocator (512 columns = 2 accumulator buffers of up to 256 fp32 columns)
//   warps 4-7: epilogue      -- tcgen05.ld 32x32b (one output row per thread), bias / erf-GELU / residual, stores
// The TMEM accumulator is double buffered, so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace capr {
namespace bert {

constexpr int BM = 128;        // rows per output tile (UMMA M)
constexpr int BK = 64;         // bf16 elements per K block = one 128-byte swizzle row
constexpr int MAX_BN = 256;    // columns per output tile (UMMA N), runtime value <= MAX_BN
constexpr int GEMM_THREADS = 256;
constexpr int A_TILE_BYTES = BM * BK * 2;      // 16 KB
constexpr int B_TILE_BYTES = MAX_BN * BK * 2;  // 32 KB

enum Epilogue : int {
  EPI_BIAS_F32 = 0,         // out_f32 = acc + bias
  EPI_BIAS_GELU_SPLIT = 1,  // (out_hi, out_lo) = split_bf16(gelu_erf(acc + bias))
  EPI_BIAS_RESID_F32 = 2,   // out_f32 = acc + bias + resid
  EPI_BIAS_SPLIT = 3,       // (out_hi, out_lo) = split_bf16(acc + bias)   (Q/K/V for the tensor-core attention)
};

struct GemmArgs {
  int M, N, K, BN, epi;
  const float* bias;      // [N]
  const float* resid;     // [M,N] (EPI_BIAS_RESID_F32)
  float* out_f32;         // [M,N]
  __nv_bfloat16* out_hi;  // [M,N]
  __nv_bfloat16* out_lo;  // [M,N]
};

template <int MODE>
struct GemmSmem {
  static constexpr int PLANES = MODE == 3 ? 2 : 1;
  static constexpr int STAGE_BYTES = PLANES * (A_TILE_BYTES + B_TILE_BYTES);  // 96 KB (MODE 3) / 48 KB (MODE 1)
  static constexpr int STAGES = MODE == 3 ? 2 : 4;
  static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// erf-GELU (HF "gelu"): 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): one MUFU.RCP, one
// MUFU.EX2 and 8 FMA-pipe instructions instead of erff's two-branch ~30.  The FFN1 epilogue evaluates 3 072 of these per token and
// was the reason its GEMM ran 12 % behind FFN2 (launch list r02: 675 us vs 604 us); measured against a float64 GELU over
// [-12, 12] the result is as accurate as torch's own fp32 erf path (4.4e-7 absolute, scripts in DESIGN.md).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-(z * z) * 1.4426950408889634f);
  const float erf_abs = fmaf(-p, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const GemmArgs g) {
  using S = GemmSmem<MODE>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + S::STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * S::STAGES;       // [2]  MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * S::STAGES + 2;  // [2]  epilogue -> MMA
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = g.N / g.BN;
  const int tiles_m = (g.M + BM - 1) / BM;
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
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 4);  // one arrive per epilogue warp
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_base_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {
      const uint32_t stage_tx = (uint32_t)S::PLANES * (uint32_t)(A_TILE_BYTES + g.BN * BK * 2);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * g.BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + stage * S::STAGE_BYTES;
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&full[stage], stage_tx);
            tc::tma_load_2d(st, &tm_a_hi, &full[stage], kb * BK, m0);
            tc::tma_load_2d(st + A_TILE_BYTES, &tm_b_hi, &full[stage], kb * BK, n0);
            if (MODE == 3) {
              tc::tma_load_2d(st + A_TILE_BYTES + B_TILE_BYTES, &tm_a_lo, &full[stage], kb * BK, m0);
              tc::tma_load_2d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &tm_b_lo, &full[stage], kb * BK, n0);
            }
          }
          __syncwarp();
          if (++stage == S::STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop and one elected lane issues: with warp-uniform control flow ptxas keeps the
    // descriptors in uniform registers (a divergent `if (lane == 0)` issuer costs ~150 cycles per tcgen05.mma).
    {
      const uint32_t idesc = tc::make_instr_desc(tc::FMT_BF16, BM, g.BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        tc::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MAX_BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::tc_fence_after();
          const uint32_t st = tc::smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t a_hi = tc::make_sw128_kmajor_desc(st);
          const uint64_t b_hi = tc::make_sw128_kmajor_desc(st + A_TILE_BYTES);
          const uint64_t a_lo = tc::make_sw128_kmajor_desc(st + A_TILE_BYTES + B_TILE_BYTES);
          const uint64_t b_lo = tc::make_sw128_kmajor_desc(st + 2 * A_TILE_BYTES + B_TILE_BYTES);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per K=16 step, in 16-byte units
              tc::umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, (kb | k) != 0);
              if (MODE == 3) {
                tc::umma_f16(d_tmem, a_lo + koff, b_hi + koff, idesc, true);
                tc::umma_f16(d_tmem, a_hi + koff, b_lo + koff, idesc, true);
              }
            }
            tc::umma_commit(&empty[stage]);  // smem stage reusable once these MMAs have read it
          }
          __syncwarp();
          if (++stage == S::STAGES) stage = 0, phase ^= 1;
        }
        if (tc::elect_one()) tc::umma_commit(&acc_full[acc]);  // accumulator complete
        __syncwarp();
        if (++acc == 2) acc = 0, acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * g.BN;
      const int row = m0 + quarter * 32 + lane;
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t)(acc * MAX_BN) + ((uint32_t)(quarter * 32) << 16);
      for (int c = 0; c < g.BN; c += 32) {
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
      if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
  }
  // teardown
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace bert
}  // namespace capr
