// K6 -- fused masked self-attention of the BERT encoder on tcgen05 (flash-style, one pass over the keys).
//
//   HF BertSelfAttention: softmax(Q K^T / sqrt(dh) + key_mask) V        as called by ptBERTMaxP.py:82
//
// One CTA per (sequence, head, block of 128 queries).  Q, K, V arrive as bf16 (hi, lo) planes (the QKV GEMM epilogue
// writes them), so both GEMMs of attention keep the 3-product precision of the Linear layers:
//   S = Q K^T : A = Q tile [128 x 64] K-major, B = K tile [64 keys x 64] K-major             -> TMEM, 64 fp32 columns
//   O += P V  : A = P tile [128 x 64 keys] K-major (written by the softmax warps as bf16 hi/lo),
//               B = V tile [64 keys x 64 dims] as loaded (dims contiguous = MN-major B)         -> TMEM, 64 fp32 columns
// Warp roles: warp 0 = TMA producer (Q once, then K/V tiles of 64 keys through a 3-stage ring), warp 1 = MMA issuer
// (QK of tile t+1 is issued before PV of tile t, so the tensor core works while the softmax warps are busy),
// warp 2 = TMEM allocator, warps 4-7 = softmax: ONE QUERY ROW PER THREAD (tcgen05.ld 32x32b), so the row max / sum of the
// online softmax need no shuffles; P goes to shared memory in the SWIZZLE_128B operand layout (st.shared +
// fence.proxy.async), the running output stays in registers and is corrected by exp2(m_old - m_new) per key tile.
// Key tiles past the last unmasked key are skipped; masked keys get weight exactly 0 (HF adds finfo.min before softmax).
#pragma once
#include "bert_gemm.cuh"

namespace capr {
namespace bert {

constexpr int AT_BQ = 128;                      // queries per CTA (UMMA M)
constexpr int AT_BK = 64;                       // keys per tile
constexpr int AT_DH = 64;                       // head dim (one 128-byte swizzle row of bf16)
constexpr int AT_THREADS = 256;
constexpr int AT_KV_STAGES = 3;
constexpr int AT_Q_BYTES = AT_BQ * 128;         // one plane of the Q tile: 16 KB
constexpr int AT_T_BYTES = AT_BK * 128;         // one plane of a K or V tile: 8 KB
constexpr int AT_KV_STAGE_BYTES = 4 * AT_T_BYTES;  // K_hi, K_lo, V_hi, V_lo
constexpr int AT_P_BYTES = AT_BQ * 128;         // one plane of a P tile [128 x 64 keys]: 16 KB
constexpr int AT_MAX_L = 512;
constexpr size_t AT_SMEM = 1024 + 2 * AT_Q_BYTES + AT_KV_STAGES * AT_KV_STAGE_BYTES + 2 * 2 * AT_P_BYTES + AT_MAX_L * 4 + 256;

// MN-major (N contiguous) B operand in SWIZZLE_128B: 64 N-elements (128 bytes) x 8 K-rows per 1024-byte atom; successive
// 8-row K groups are `stride` bytes apart.  Used for V: rows = keys (K of the P.V product), 128 bytes of dims per row.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;  // leading byte offset: next 64-element block along N (unused: N = 64)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset: next group of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct AttnArgs {
  int L, H, heads, n_seq;
  float scale_log2e;
  const long long* mask;  // [n_seq, L]
  __nv_bfloat16* ctx_hi;  // [T, H]
  __nv_bfloat16* ctx_lo;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                    const __grid_constant__ CUtensorMap tm_kv_hi, const __grid_constant__ CUtensorMap tm_kv_lo, const AttnArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                                    // hi | lo
  unsigned char* sKV = sQ + 2 * AT_Q_BYTES;                    // [stage][K_hi, K_lo, V_hi, V_lo]
  unsigned char* sP = sKV + AT_KV_STAGES * AT_KV_STAGE_BYTES;  // [buf][hi | lo]
  float* kbias = reinterpret_cast<float*>(sP + 2 * 2 * AT_P_BYTES);  // [AT_MAX_L]  0 or -inf per key
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbias + AT_MAX_L);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]
  uint64_t* s_full = bars + 7;        // [2]
  uint64_t* s_empty = bars + 9;       // [2]
  uint64_t* p_full = bars + 11;       // [2]
  uint64_t* p_empty = bars + 13;      // [2]
  uint64_t* o_full = bars + 15;       // [2]
  uint64_t* o_empty = bars + 17;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  int* s_kv_len = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int qblocks = (a.L + AT_BQ - 1) / AT_BQ;
  const int qb = blockIdx.x % qblocks, head = (blockIdx.x / qblocks) % a.heads, seq = blockIdx.x / (qblocks * a.heads);
  const int tok0 = seq * a.L;
  const long long* mrow = a.mask + (size_t)seq * a.L;

  if (tid == 0) {
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < AT_KV_STAGES; ++i) tc::mbar_init(&kv_full[i], 1), tc::mbar_init(&kv_empty[i], 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&s_empty[i], 4);   // one arrive per softmax warp
      tc::mbar_init(&p_full[i], 128);  // every softmax thread publishes its own row of P
      tc::mbar_init(&p_empty[i], 1);
      tc::mbar_init(&o_full[i], 1);
      tc::mbar_init(&o_empty[i], 4);
    }
    tc::fence_barrier_init();
    *s_kv_len = 0;
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // key bias + last unmasked key (all threads)
  int last = 0;
  for (int j = tid; j < AT_MAX_L; j += AT_THREADS) {
    const bool on = j < a.L && mrow[j] != 0;
    kbias[j] = on ? 0.f : -INFINITY;
    if (on) last = j + 1;
  }
  if (last) atomicMax(s_kv_len, last);
  __syncthreads();
  const int n_tiles = (*s_kv_len + AT_BK - 1) / AT_BK;
  const int col_q = head * AT_DH, col_k = a.H + head * AT_DH, col_v = 2 * a.H + head * AT_DH;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, 2 * AT_Q_BYTES);
      tc::tma_load_2d(sQ, &tm_q_hi, q_full, col_q, tok0 + qb * AT_BQ);
      tc::tma_load_2d(sQ + AT_Q_BYTES, &tm_q_lo, q_full, col_q, tok0 + qb * AT_BQ);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < n_tiles; ++t) {
      tc::mbar_wait(&kv_empty[stage], phase ^ 1);
      unsigned char* st = sKV + stage * AT_KV_STAGE_BYTES;
      const int row = tok0 + t * AT_BK;
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&kv_full[stage], AT_KV_STAGE_BYTES);
        tc::tma_load_2d(st, &tm_kv_hi, &kv_full[stage], col_k, row);
        tc::tma_load_2d(st + AT_T_BYTES, &tm_kv_lo, &kv_full[stage], col_k, row);
        tc::tma_load_2d(st + 2 * AT_T_BYTES, &tm_kv_hi, &kv_full[stage], col_v, row);
        tc::tma_load_2d(st + 3 * AT_T_BYTES, &tm_kv_lo, &kv_full[stage], col_v, row);
      }
      __syncwarp();
      if (++stage == AT_KV_STAGES) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    if (n_tiles > 0) {
      const uint32_t idesc_qk = tc::make_instr_desc(tc::FMT_BF16, AT_BQ, AT_BK);
      const uint32_t idesc_pv = tc::make_instr_desc(tc::FMT_BF16, AT_BQ, AT_DH) | (1u << 16);  // B is MN-major (V: dims contiguous)
      const uint32_t q_hi = tc::smem_u32(sQ), q_lo = q_hi + AT_Q_BYTES;
      tc::mbar_wait(q_full, 0);
      auto issue_qk = [&](int t) {
        const int stage = t % AT_KV_STAGES, buf = t & 1;
        tc::mbar_wait(&kv_full[stage], (uint32_t)((t / AT_KV_STAGES) & 1));
        tc::mbar_wait(&s_empty[buf], (uint32_t)(((t >> 1) & 1) ^ 1));
        tc::tc_fence_after();
        const uint32_t k_hi = tc::smem_u32(sKV + stage * AT_KV_STAGE_BYTES), k_lo = k_hi + AT_T_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * AT_BK);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DH / 16; ++k) {
            const uint32_t ko = k * 32;
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_hi + ko), tc::make_sw128_kmajor_desc(k_hi + ko), idesc_qk, k != 0);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_lo + ko), tc::make_sw128_kmajor_desc(k_hi + ko), idesc_qk, true);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_hi + ko), tc::make_sw128_kmajor_desc(k_lo + ko), idesc_qk, true);
          }
          tc::umma_commit(&s_full[buf]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t) {
        const int stage = t % AT_KV_STAGES, buf = t & 1;
        tc::mbar_wait(&p_full[buf], (uint32_t)((t >> 1) & 1));
        tc::mbar_wait(&o_empty[buf], (uint32_t)(((t >> 1) & 1) ^ 1));
        tc::tc_fence_after();
        const uint32_t v_hi = tc::smem_u32(sKV + stage * AT_KV_STAGE_BYTES + 2 * AT_T_BYTES), v_lo = v_hi + AT_T_BYTES;
        const uint32_t p_hi = tc::smem_u32(sP + buf * 2 * AT_P_BYTES), p_lo = p_hi + AT_P_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(128 + buf * AT_DH);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_BK / 16; ++k) {
            const uint32_t pk = k * 32;           // 16 keys = 32 bytes along P's K-major rows
            const uint32_t vk = k * 16 * 128;     // 16 keys = 16 rows of the V tile
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(p_hi + pk), make_sw128_mnmajor_desc(v_hi + vk), idesc_pv, k != 0);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(p_lo + pk), make_sw128_mnmajor_desc(v_hi + vk), idesc_pv, true);
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(p_hi + pk), make_sw128_mnmajor_desc(v_lo + vk), idesc_pv, true);
          }
          tc::umma_commit(&o_full[buf]);
          tc::umma_commit(&kv_empty[stage]);
          tc::umma_commit(&p_empty[buf]);
        }
        __syncwarp();
      };
      issue_qk(0);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_qk(t + 1);
        issue_pv(t);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / output: one query row per thread =====================
    const int quarter = warp & 3;
    const int row_in_blk = quarter * 32 + lane;
    const int qrow = qb * AT_BQ + row_in_blk;  // position inside the sequence
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float o[AT_DH];
#pragma unroll
    for (int i = 0; i < AT_DH; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const uint32_t ph = (uint32_t)((t >> 1) & 1);
      tc::mbar_wait(&s_full[buf], ph);
      tc::tc_fence_after();
      float s[AT_BK];
      {
        float lo32[32], hi32[32];
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(buf * AT_BK), lo32);
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(buf * AT_BK + 32), hi32);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = lo32[i], s[32 + i] = hi32[i];
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_empty[buf]);
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < AT_BK; ++i) {
        s[i] = fmaf(s[i], a.scale_log2e, kbias[t * AT_BK + i]);
        mx = fmaxf(mx, s[i]);
      }
      const float m_new = fmaxf(m_run, mx);
      const bool dead = m_new == -INFINITY;
      const float corr = dead ? 1.f : ex2_approx(m_run - m_new);
      float rs = 0.f;
      tc::mbar_wait(&p_empty[buf], ph ^ 1);
      unsigned char* prow_hi = sP + buf * 2 * AT_P_BYTES + row_in_blk * 128;
      unsigned char* prow_lo = prow_hi + AT_P_BYTES;
#pragma unroll
      for (int c = 0; c < 8; ++c) {  // 8 keys per 16-byte chunk
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p0 = dead ? 0.f : ex2_approx(s[c * 8 + 2 * j] - m_new);
          const float p1 = dead ? 0.f : ex2_approx(s[c * 8 + 2 * j + 1] - m_new);
          rs += p0 + p1;
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(p0, h0, l0);
          split_bf16(p1, h1, l1);
          hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        const int pos = (c ^ (row_in_blk & 7)) << 4;  // SWIZZLE_128B: chunk index XOR (row % 8)
        *reinterpret_cast<uint4*>(prow_hi + pos) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(prow_lo + pos) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      tc::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      tc::mbar_arrive(&p_full[buf]);
      l_run = l_run * corr + rs;
      // fold in the P.V product of the previous tile (it was computed against m_run), then rescale to m_new
      if (t > 0) {
        const int pb = (t - 1) & 1;
        tc::mbar_wait(&o_full[pb], (uint32_t)(((t - 1) >> 1) & 1));
        tc::tc_fence_after();
        float a0[32], a1[32];
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(128 + pb * AT_DH), a0);
        tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(128 + pb * AT_DH + 32), a1);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&o_empty[pb]);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = (o[i] + a0[i]) * corr, o[32 + i] = (o[32 + i] + a1[i]) * corr;
      }
      m_run = m_new;
    }
    if (n_tiles > 0) {
      const int pb = (n_tiles - 1) & 1;
      tc::mbar_wait(&o_full[pb], (uint32_t)(((n_tiles - 1) >> 1) & 1));
      tc::tc_fence_after();
      float a0[32], a1[32];
      tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(128 + pb * AT_DH), a0);
      tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(128 + pb * AT_DH + 32), a1);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] += a0[i], o[32 + i] += a1[i];
    }
    if (qrow < a.L) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      const size_t off = (size_t)(tok0 + qrow) * a.H + head * AT_DH;
      uint4* ph4 = reinterpret_cast<uint4*>(a.ctx_hi + off);
      uint4* pl4 = reinterpret_cast<uint4*>(a.ctx_lo + off);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(o[c * 8 + 2 * j] * inv, h0, l0);
          split_bf16(o[c * 8 + 2 * j + 1] * inv, h1, l1);
          hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        ph4[c] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        pl4[c] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace bert
}  // namespace capr
