// K6b -- masked self-attention of the BERT encoder on tcgen05, second generation: 256 queries per CTA.
//
//   HF BertSelfAttention: softmax(Q K^T / sqrt(dh) + key_mask) V        as called by ptBERTMaxP.py:82
//
// Same arithmetic as bert_attn.cuh (bf16 hi/lo planes, three products per K step for S = Q K^T and O = P V, one query row
// per softmax thread), restructured because the v1 kernel was bound by its four softmax warps (one per scheduler, nothing
// to hide the tcgen05.ld / MUFU / mbarrier latencies behind), not by the tensor pipe (ncu launch list r01_v5: 868 us per
// layer against ~250 us of MMA time):
//   * one CTA owns TWO blocks of 128 queries of the same (sequence, head): 8 softmax warps (two per scheduler) that share
//     every K/V tile (half the K/V TMA traffic) and ping-pong on the tensor pipe;
//   * the output accumulator STAYS IN TMEM across all key tiles (tcgen05.mma accumulates in place).  The running max used
//     inside exp2 is only raised when a tile's max exceeds it by more than 8 (log2 units), in which case the warp rescales
//     its rows of O in TMEM (tcgen05.ld -> multiply -> tcgen05.st) and the row sum; otherwise nothing is read back until the
//     end.  exp2(s - m_ref) <= 256, so nothing overflows and O / l is the same softmax-weighted mean;
//   * P is packed with cvt.rn.bf16x2.f32 (two values per instruction).
// Warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = softmax of query block 0, 8-11 = softmax of block 1.
// TMEM: S[block][2 buffers][64 cols] at 0..255, O[block][64 cols] at 256..383.
// K/V tiles go through a THREE-stage ring: the clock64 trace of a CTA (scripts/attn_trace.py) showed the two-stage ring exposing
// the whole TMA round trip (~2 200 cycles) every tile -- the stage of tile t is only released when P.V(t) of the second block
// completes, and the MMA warp then sat on kv_full(t+2) with P(t+1) already published.  The key mask is kept as 16 ballot words
// (64 B instead of a 2 KB float bias array) to make room for the third stage.
#pragma once
#include "bert_attn.cuh"

namespace capr {
namespace bert {

constexpr int A2_BLOCKS = 2;                    // query blocks of AT_BQ rows per CTA
constexpr int A2_THREADS = 384;
constexpr int A2_KV_STAGES = 3;               // the K/V TMA round trip (~2 200 cycles, traced) needs two tiles of prefetch to hide
constexpr float A2_RESCALE_THRESHOLD = 8.0f;    // log2 units
constexpr size_t A2_SMEM = 1024 + A2_BLOCKS * 2 * AT_Q_BYTES + A2_KV_STAGES * AT_KV_STAGE_BYTES + A2_BLOCKS * 2 * AT_P_BYTES + AT_MAX_L / 8 + 512;
static_assert(A2_SMEM <= 232448, "attention_tc2_kernel: shared memory budget");

struct Attn2Args {
  int L, H, heads, n_seq;
  long long total_rows;  // rows of the Q/K/V planes (padded token count)
  float scale_log2e;
  const long long* mask;  // [n_seq, L]
  __nv_bfloat16* ctx_hi;  // [T, H]
  __nv_bfloat16* ctx_lo;
  int debug;              // profiling only (CAPR_ATTN_DEBUG; results invalid): 1 no softmax math, 2 no P.V MMAs, 4 no K/V reloads, 8 no Q.K MMAs
  long long* trace;       // profiling only (CAPR_ATTN_TRACE): CTA 0 of the grid records clock64 stamps, [role][event] (see TR_* below)
  int qk_products;        // attention_tc4_kernel only: bf16 products of S = Q.K^T: 3 = q_hi.k_hi + q_lo.k_hi + q_hi.k_lo (round 1), 2 = without
                          // q_hi.k_lo, 1 = q_hi.k_hi alone.  The scores only feed a softmax: scripts/bert_precision_probe2.py measures the
                          // final score error of BERT-base at 2.8e-5 (3), 1.4e-4 (2), 2.3e-4 (1) against the 1e-3 bar; every Linear layer
                          // and P.V need all three products (dropping one costs >= 6.7e-4).
  int q_blocks;           // attention_tc4_kernel only: 256-query blocks per (sequence, head) to compute, 0 = all of them.  The last
                          // encoder layer of a classification forward needs the [CLS] row alone (ptBERTMaxP.py:82 reads logits of
                          // the pooled row 0), i.e. block 0.
};

// trace layout: 64 slots per role; roles: 0 producer, 1 MMA, 2 softmax block 0 (warp 4), 3 softmax block 1 (warp 8)
#ifdef CAPR_DEBUG_BUILD
#define CAPR_TR(role, slot) do { if (a.trace && blockIdx.x == 148 && lane == 0 && (slot) < 64) a.trace[(role) * 64 + (slot)] = clock64(); } while (0)
#else
#define CAPR_TR(role, slot) do { } while (0)
#endif

// (hi, lo) bf16 split of two values at once; returns the packed words {lo16 = a, hi16 = b}
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hw, uint32_t& lw) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hw = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hw << 16), hb = __uint_as_float(hw & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lw = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(A2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                     const __grid_constant__ CUtensorMap tm_kv_hi, const __grid_constant__ CUtensorMap tm_kv_lo, const Attn2Args a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                                          // [block][hi | lo]
  unsigned char* sKV = sQ + A2_BLOCKS * 2 * AT_Q_BYTES;              // [stage][K_hi, K_lo, V_hi, V_lo]
  unsigned char* sP = sKV + A2_KV_STAGES * AT_KV_STAGE_BYTES;        // [block][hi | lo]
  uint32_t* kmask = reinterpret_cast<uint32_t*>(sP + A2_BLOCKS * 2 * AT_P_BYTES);  // [AT_MAX_L / 32] bit j = key j is attended to
  uint64_t* bars = reinterpret_cast<uint64_t*>(kmask + AT_MAX_L / 32);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]
  uint64_t* s_full = bars + 7;        // [block][2]
  uint64_t* s_empty = bars + 11;      // [block][2]
  uint64_t* p_full = bars + 15;       // [block]
  uint64_t* p_empty = bars + 17;      // [block]
  uint64_t* o_full = bars + 19;       // [block]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  int* s_kv_len = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (warp == 0) CAPR_TR(0, 62);  // kernel entry
  const int qblocks = (a.L + A2_BLOCKS * AT_BQ - 1) / (A2_BLOCKS * AT_BQ);
  const int qb = blockIdx.x % qblocks, head = (blockIdx.x / qblocks) % a.heads, seq = blockIdx.x / (qblocks * a.heads);
  const int tok0 = seq * a.L;
  const long long* mrow = a.mask + (size_t)seq * a.L;

  if (tid == 0) {
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < A2_KV_STAGES; ++i) tc::mbar_init(&kv_full[i], 1), tc::mbar_init(&kv_empty[i], 1);
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&s_empty[i], 4);  // one arrive per softmax warp of the block
    }
    for (int g = 0; g < A2_BLOCKS; ++g) {
      tc::mbar_init(&p_full[g], 128);  // every softmax thread publishes its own row of P
      tc::mbar_init(&p_empty[g], 1);
      tc::mbar_init(&o_full[g], 1);
    }
    tc::fence_barrier_init();
    *s_kv_len = 0;
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) CAPR_TR(0, 61);  // barriers + TMEM ready
  // key mask as a bitmask (one ballot per 32 keys) and the position after the last attended key
  for (int w = warp; w < AT_MAX_L / 32; w += A2_THREADS / 32) {
    const int j = w * 32 + lane;
    const unsigned bits = __ballot_sync(0xffffffffu, j < a.L && mrow[j] != 0);
    if (lane == 0) {
      kmask[w] = bits;
      if (bits) atomicMax(s_kv_len, w * 32 + 32 - __clz(bits));
    }
  }
  __syncthreads();
  const int n_tiles = (*s_kv_len + AT_BK - 1) / AT_BK;
  const int col_q = head * AT_DH, col_k = a.H + head * AT_DH, col_v = 2 * a.H + head * AT_DH;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    CAPR_TR(0, 0);  // prologue done
    if (n_tiles > 0) {
      if (tc::elect_one()) {
        tc::mbar_expect_tx(q_full, A2_BLOCKS * 2 * AT_Q_BYTES);
        for (int g = 0; g < A2_BLOCKS; ++g) {
          long long row = (long long)tok0 + (qb * A2_BLOCKS + g) * AT_BQ;
          if (row >= a.total_rows) row = (long long)tok0 + qb * A2_BLOCKS * AT_BQ;  // block entirely past the data: its rows are never stored
          tc::tma_load_2d(sQ + (g * 2) * AT_Q_BYTES, &tm_q_hi, q_full, col_q, (int)row);
          tc::tma_load_2d(sQ + (g * 2 + 1) * AT_Q_BYTES, &tm_q_lo, q_full, col_q, (int)row);
        }
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        tc::mbar_wait(&kv_empty[stage], phase ^ 1);
        CAPR_TR(0, 1 + t);  // producer: stage free, issuing loads of tile t
        unsigned char* st = sKV + stage * AT_KV_STAGE_BYTES;
        const int row = tok0 + t * AT_BK;
        if (CAPR_DBG(a.debug & 4) && t >= A2_KV_STAGES) {
          if (tc::elect_one()) tc::mbar_arrive(&kv_full[stage]);
        } else if (tc::elect_one()) {
          tc::mbar_expect_tx(&kv_full[stage], AT_KV_STAGE_BYTES);
          tc::tma_load_2d(st, &tm_kv_hi, &kv_full[stage], col_k, row);
          tc::tma_load_2d(st + AT_T_BYTES, &tm_kv_lo, &kv_full[stage], col_k, row);
          tc::tma_load_2d(st + 2 * AT_T_BYTES, &tm_kv_hi, &kv_full[stage], col_v, row);
          tc::tma_load_2d(st + 3 * AT_T_BYTES, &tm_kv_lo, &kv_full[stage], col_v, row);
        }
        __syncwarp();
        if (++stage == A2_KV_STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    if (n_tiles > 0) {
      const uint32_t idesc_qk = tc::make_instr_desc(tc::FMT_BF16, AT_BQ, AT_BK);
      const uint32_t idesc_pv = tc::make_instr_desc(tc::FMT_BF16, AT_BQ, AT_DH) | (1u << 16);  // B is MN-major (V: dims contiguous)
      tc::mbar_wait(q_full, 0);
      CAPR_TR(1, 0);  // MMA: Q landed
      auto issue_qk = [&](int g, int t) {
        const int stage = t % A2_KV_STAGES, buf = t & 1;
        tc::mbar_wait(&kv_full[stage], (uint32_t)((t / A2_KV_STAGES) & 1));
        tc::mbar_wait(&s_empty[g * 2 + buf], (uint32_t)(((t >> 1) & 1) ^ 1));
        CAPR_TR(1, 1 + 4 * t + g);  // MMA: issuing QK(g, t)
        tc::tc_fence_after();
        const uint32_t q_hi = tc::smem_u32(sQ + (g * 2) * AT_Q_BYTES), q_lo = q_hi + AT_Q_BYTES;
        const uint32_t k_hi = tc::smem_u32(sKV + stage * AT_KV_STAGE_BYTES), k_lo = k_hi + AT_T_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(g * 128 + buf * AT_BK);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < (CAPR_DBG(a.debug & 8) ? 0 : AT_DH / 16); ++k) {
            const uint32_t ko = k * 32;
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_hi + ko), tc::make_sw128_kmajor_desc(k_hi + ko), idesc_qk, k != 0);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_lo + ko), tc::make_sw128_kmajor_desc(k_hi + ko), idesc_qk, true);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_hi + ko), tc::make_sw128_kmajor_desc(k_lo + ko), idesc_qk, true);
          }
          tc::umma_commit(&s_full[g * 2 + buf]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int t) {
        const int stage = t % A2_KV_STAGES;
        tc::mbar_wait(&p_full[g], (uint32_t)(t & 1));
        CAPR_TR(1, 1 + 4 * t + 2 + g);  // MMA: issuing PV(g, t)
        tc::tc_fence_after();
        const uint32_t v_hi = tc::smem_u32(sKV + stage * AT_KV_STAGE_BYTES + 2 * AT_T_BYTES), v_lo = v_hi + AT_T_BYTES;
        const uint32_t p_hi = tc::smem_u32(sP + (g * 2) * AT_P_BYTES), p_lo = p_hi + AT_P_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(256 + g * AT_DH);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < (CAPR_DBG(a.debug & 2) ? 0 : AT_BK / 16); ++k) {
            const uint32_t pk = k * 32;        // 16 keys = 32 bytes along P's K-major rows
            const uint32_t vk = k * 16 * 128;  // 16 keys = 16 rows of the V tile
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(p_hi + pk), make_sw128_mnmajor_desc(v_hi + vk), idesc_pv, (t | k) != 0);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(p_lo + pk), make_sw128_mnmajor_desc(v_hi + vk), idesc_pv, true);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(p_hi + pk), make_sw128_mnmajor_desc(v_lo + vk), idesc_pv, true);
          }
          tc::umma_commit(&o_full[g]);
          tc::umma_commit(&p_empty[g]);
          if (g == A2_BLOCKS - 1) tc::umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
      };
      issue_qk(0, 0);
      issue_qk(1, 0);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_qk(0, t + 1);
        issue_pv(0, t);
        if (t + 1 < n_tiles) issue_qk(1, t + 1);
        issue_pv(1, t);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / output: one query row per thread =====================
    const int g = (warp - 4) >> 2, quarter = warp & 3;
    const int row_in_blk = quarter * 32 + lane;
    const int qrow = (qb * A2_BLOCKS + g) * AT_BQ + row_in_blk;  // position inside the sequence
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + (uint32_t)(256 + g * AT_DH);
    float m_ref = -INFINITY, l_run = 0.f;
    unsigned char* prow_hi = sP + (g * 2) * AT_P_BYTES + row_in_blk * 128;
    unsigned char* prow_lo = prow_hi + AT_P_BYTES;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      tc::mbar_wait(&s_full[g * 2 + buf], (uint32_t)((t >> 1) & 1));
      if (quarter == 0) CAPR_TR(2 + g, 4 * t);  // softmax: S(t) visible
      tc::tc_fence_after();
      float s[AT_BK];
      {
        float lo32[32], hi32[32];
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(g * 128 + buf * AT_BK), lo32);
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(g * 128 + buf * AT_BK + 32), hi32);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = lo32[i], s[32 + i] = hi32[i];
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_empty[g * 2 + buf]);
      if (CAPR_DBG(a.debug & 1)) {
        tc::mbar_wait(&p_empty[g], (uint32_t)((t & 1) ^ 1));
        tc::mbar_arrive(&p_full[g]);
        continue;
      }
      float mx = -INFINITY;
      const uint32_t km0 = kmask[2 * t], km1 = kmask[2 * t + 1];
      if ((km0 & km1) == 0xffffffffu) {  // CTA-uniform fast path: every key of the tile is attended to
#pragma unroll
        for (int i = 0; i < AT_BK; ++i) s[i] *= a.scale_log2e;
      } else {
#pragma unroll
        for (int i = 0; i < AT_BK; ++i) s[i] = (((i < 32 ? km0 : km1) >> (i & 31)) & 1u) ? s[i] * a.scale_log2e : -INFINITY;
      }
#pragma unroll
      for (int i = 0; i < AT_BK / 4; ++i) mx = fmaxf(mx, fmaxf(fmaxf(s[4 * i], s[4 * i + 1]), fmaxf(s[4 * i + 2], s[4 * i + 3])));
      if (t == 0) {
        m_ref = mx;
      } else {
        const bool need = mx > m_ref + A2_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          // raise the reference max of the rows that need it: O and l were accumulated against the old one
          tc::mbar_wait(&o_full[g], (uint32_t)((t - 1) & 1));  // P.V of tile t-1 has landed; tile t is not issued before our p_full
          tc::tc_fence_after();
          const float factor = need ? ex2_approx(m_ref - mx) : 1.0f;
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            float o[32];
            tc::tmem_ld_32x32(o_tmem + (uint32_t)(hlf * 32), o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= factor;
            tc::tmem_st_32x32(o_tmem + (uint32_t)(hlf * 32), o);
          }
          tc::tmem_st_wait();
          tc::tc_fence_before();
          l_run *= factor;
          m_ref = need ? mx : m_ref;
        }
      }
      const bool dead = m_ref == -INFINITY;
      float rs = 0.f;
      if (quarter == 0) CAPR_TR(2 + g, 4 * t + 1);  // softmax: max known, waiting for the P buffer
      tc::mbar_wait(&p_empty[g], (uint32_t)((t & 1) ^ 1));  // P.V of tile t-1 is done reading the P buffer
      if (quarter == 0) CAPR_TR(2 + g, 4 * t + 2);  // softmax: P buffer free
#pragma unroll
      for (int c = 0; c < 8; ++c) {  // 8 keys per 16-byte chunk
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p0 = dead ? 0.f : ex2_approx(s[c * 8 + 2 * j] - m_ref);
          const float p1 = dead ? 0.f : ex2_approx(s[c * 8 + 2 * j + 1] - m_ref);
          rs += p0 + p1;
          split2_bf16(p0, p1, hw[j], lw[j]);
        }
        const int pos = (c ^ (row_in_blk & 7)) << 4;  // SWIZZLE_128B: chunk index XOR (row % 8)
        *reinterpret_cast<uint4*>(prow_hi + pos) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(prow_lo + pos) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      tc::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      tc::mbar_arrive(&p_full[g]);
      if (quarter == 0) CAPR_TR(2 + g, 4 * t + 3);  // softmax: P(t) published
      l_run += rs;
    }
    if (n_tiles > 0) {
      tc::mbar_wait(&o_full[g], (uint32_t)((n_tiles - 1) & 1));
      tc::tc_fence_after();
    }
    if (quarter == 0) CAPR_TR(2 + g, 60);  // softmax: last P.V landed
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    const size_t off = (size_t)(tok0 + qrow) * a.H + head * AT_DH;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      float o[32];
      if (n_tiles > 0) {
        tc::tmem_ld_32x32(o_tmem + (uint32_t)(hlf * 32), o);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
      if (qrow < a.L) {
        uint4* ph4 = reinterpret_cast<uint4*>(a.ctx_hi + off + hlf * 32);
        uint4* pl4 = reinterpret_cast<uint4*>(a.ctx_lo + off + hlf * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split2_bf16(o[c * 8 + 2 * j] * inv, o[c * 8 + 2 * j + 1] * inv, hw[j], lw[j]);
          ph4[c] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          pl4[c] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) CAPR_TR(0, 63);  // all roles done
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}


}  // namespace bert
}  // namespace capr
