// K6c -- persistent masked self-attention: attention_tc2_kernel's pipeline, looping over (sequence, head, query block) items.
//
// The clock64 trace of attention_tc2_kernel (scripts/attn_trace.py, three-stage K/V ring) splits a CTA's 46.8 k cycles into
// 28 k of steady-state tiles and ~17 k of everything else: CTA launch gap (~6 k), barrier init + TMEM alloc + mask scan
// (~2 k), first Q/K round trip (~2.3 k) and the output epilogue (~6.5 k) -- none of it overlapped, because one CTA owns an SM.
// Here one CTA per SM stays resident and every role loops over work items, so that
//   * barriers and TMEM are set up once;
//   * a dedicated warp scans the key mask of item i+1 (16 ballot words + the tile count) while item i is computed;
//   * the producer loads Q of item i+1 as soon as the last Q.K^T of item i has been issued and streams K/V tiles through the
//     same three-stage ring without a break;
//   * the softmax warps' epilogue (O / l -> bf16 hi/lo -> global) overlaps the first tiles of the next item: the MMA warp only
//     waits for their tcgen05.ld of O (o_free) before the first P.V of the next item overwrites the accumulator.
// All buffer indices and mbarrier parities are derived from running counters (global tile index G, non-empty item count qi),
// so nothing resets at item boundaries.  Arithmetic is identical to attention_tc2_kernel.
#pragma once
#include "bert_attn2.cuh"

namespace capr {
namespace bert {

constexpr int A4_THREADS = 384;
constexpr size_t A4_SMEM = 1024 + A2_BLOCKS * 2 * AT_Q_BYTES + A2_KV_STAGES * AT_KV_STAGE_BYTES + A2_BLOCKS * 2 * AT_P_BYTES + 2 * (AT_MAX_L / 8) + 512;
static_assert(A4_SMEM <= 232448, "attention_tc4_kernel: shared memory budget");

__global__ void __launch_bounds__(A4_THREADS, 1)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                     const __grid_constant__ CUtensorMap tm_kv_hi, const __grid_constant__ CUtensorMap tm_kv_lo, const Attn2Args a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                                          // [block][hi | lo]
  unsigned char* sKV = sQ + A2_BLOCKS * 2 * AT_Q_BYTES;              // [stage][K_hi, K_lo, V_hi, V_lo]
  unsigned char* sP = sKV + A2_KV_STAGES * AT_KV_STAGE_BYTES;        // [block][hi | lo]
  uint32_t* kmask = reinterpret_cast<uint32_t*>(sP + A2_BLOCKS * 2 * AT_P_BYTES);  // [2 item parities][AT_MAX_L / 32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(kmask + 2 * (AT_MAX_L / 32));
  uint64_t* q_full = bars;            // [1]
  uint64_t* q_empty = bars + 1;       // [1]
  uint64_t* kv_full = bars + 2;       // [3]
  uint64_t* kv_empty = bars + 5;      // [3]
  uint64_t* s_full = bars + 8;        // [block][2]
  uint64_t* s_empty = bars + 12;      // [block][2]
  uint64_t* p_full = bars + 16;       // [block]
  uint64_t* p_empty = bars + 18;      // [block]
  uint64_t* o_full = bars + 20;       // [block]
  uint64_t* o_free = bars + 22;       // [block]  the softmax warps have read the finished accumulator
  uint64_t* km_full = bars + 24;      // [item parity]  mask words + tile count of the item are in shared memory
  uint64_t* km_empty = bars + 26;     // [item parity]  every consumer warp is done with them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  int* ntiles = reinterpret_cast<int*>(tmem_slot + 2);  // [2 item parities]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int all_qblocks = (a.L + A2_BLOCKS * AT_BQ - 1) / (A2_BLOCKS * AT_BQ);
  const int qblocks = (a.q_blocks > 0 && a.q_blocks < all_qblocks) ? a.q_blocks : all_qblocks;
  const int n_items = a.n_seq * a.heads * qblocks;

  if (tid == 0) {
    tc::mbar_init(q_full, 1);
    tc::mbar_init(q_empty, 1);
    for (int i = 0; i < A2_KV_STAGES; ++i) tc::mbar_init(&kv_full[i], 1), tc::mbar_init(&kv_empty[i], 1);
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&s_empty[i], 4);  // one arrive per softmax warp of the block
    }
    for (int g = 0; g < A2_BLOCKS; ++g) {
      tc::mbar_init(&p_full[g], 128);  // every softmax thread publishes its own row of P
      tc::mbar_init(&p_empty[g], 1);
      tc::mbar_init(&o_full[g], 1);
      tc::mbar_init(&o_free[g], 4);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&km_full[i], 1);
      tc::mbar_init(&km_empty[i], 10);  // producer warp + MMA warp + 8 softmax warps
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_coords = [&](int item, int& qb, int& head, int& seq) {
    qb = item % qblocks;
    head = (item / qblocks) % a.heads;
    seq = item / (qblocks * a.heads);
  };

  if (warp == 3) {
    // ===================== mask scanner: one item ahead of everybody else =====================
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int par = it & 1;
      int qb, head, seq;
      item_coords(item, qb, head, seq);
      const long long* mrow = a.mask + (size_t)seq * a.L;
      long long v[AT_MAX_L / 32];
#pragma unroll
      for (int w = 0; w < AT_MAX_L / 32; ++w) {
        const int j = w * 32 + lane;
        v[w] = j < a.L ? mrow[j] : 0;
      }
      tc::mbar_wait(&km_empty[par], (uint32_t)(((it >> 1) & 1) ^ 1));
      int len = 0;
#pragma unroll
      for (int w = 0; w < AT_MAX_L / 32; ++w) {
        const unsigned bits = __ballot_sync(0xffffffffu, v[w] != 0);
        if (lane == 0) kmask[par * (AT_MAX_L / 32) + w] = bits;
        if (bits) len = w * 32 + 32 - __clz(bits);
      }
      if (lane == 0) {
        ntiles[par] = (len + AT_BK - 1) / AT_BK;
        tc::mbar_arrive(&km_full[par]);  // (release: the plain stores above are visible to the waiters)
      }
      __syncwarp();
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    int it = 0, qi = 0, stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int par = it & 1;
      int qb, head, seq;
      item_coords(item, qb, head, seq);
      tc::mbar_wait(&km_full[par], (uint32_t)((it >> 1) & 1));
      const int n = ntiles[par];
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&km_empty[par]);
      if (n == 0) continue;
      const int tok0 = seq * a.L;
      const int col_q = head * AT_DH, col_k = a.H + head * AT_DH, col_v = 2 * a.H + head * AT_DH;
      tc::mbar_wait(q_empty, (uint32_t)((qi & 1) ^ 1));  // every Q.K^T of the previous item has read the Q tiles
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
      for (int t = 0; t < n; ++t) {
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
        if (++stage == A2_KV_STAGES) stage = 0, phase ^= 1;
      }
      ++qi;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loops, one elected lane issues) =====================
    const uint32_t idesc_qk = tc::make_instr_desc(tc::FMT_BF16, AT_BQ, AT_BK);
    const uint32_t idesc_pv = tc::make_instr_desc(tc::FMT_BF16, AT_BQ, AT_DH) | (1u << 16);  // B is MN-major (V: dims contiguous)
    int it = 0, qi = 0, gt = 0;  // gt: global index of the item's first tile
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int par = it & 1;
      tc::mbar_wait(&km_full[par], (uint32_t)((it >> 1) & 1));
      const int n = ntiles[par];
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&km_empty[par]);
      if (n == 0) continue;
      tc::mbar_wait(q_full, (uint32_t)(qi & 1));
      auto issue_qk = [&](int g, int t) {  // t: tile inside the item
        const int G = gt + t;
        const int stage = G % A2_KV_STAGES, buf = G & 1;
        tc::mbar_wait(&kv_full[stage], (uint32_t)((G / A2_KV_STAGES) & 1));
        tc::mbar_wait(&s_empty[g * 2 + buf], (uint32_t)(((G >> 1) & 1) ^ 1));
        tc::tc_fence_after();
        const uint32_t q_hi = tc::smem_u32(sQ + (g * 2) * AT_Q_BYTES), q_lo = q_hi + AT_Q_BYTES;
        const uint32_t k_hi = tc::smem_u32(sKV + stage * AT_KV_STAGE_BYTES), k_lo = k_hi + AT_T_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(g * 128 + buf * AT_BK);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DH / 16; ++k) {
            const uint32_t ko = k * 32;
            tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_hi + ko), tc::make_sw128_kmajor_desc(k_hi + ko), idesc_qk, k != 0);
            if (a.qk_products >= 2) tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_lo + ko), tc::make_sw128_kmajor_desc(k_hi + ko), idesc_qk, true);
            if (a.qk_products >= 3) tc::umma_f16(d_tmem, tc::make_sw128_kmajor_desc(q_hi + ko), tc::make_sw128_kmajor_desc(k_lo + ko), idesc_qk, true);
          }
          tc::umma_commit(&s_full[g * 2 + buf]);
          if (g == A2_BLOCKS - 1 && t == n - 1) tc::umma_commit(q_empty);  // the item's last Q.K^T: the Q tiles may be replaced
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int t) {
        const int G = gt + t;
        const int stage = G % A2_KV_STAGES;
        tc::mbar_wait(&p_full[g], (uint32_t)(G & 1));
        if (t == 0) tc::mbar_wait(&o_free[g], (uint32_t)((qi & 1) ^ 1));  // the previous item's output has been read out of TMEM
        tc::tc_fence_after();
        const uint32_t v_hi = tc::smem_u32(sKV + stage * AT_KV_STAGE_BYTES + 2 * AT_T_BYTES), v_lo = v_hi + AT_T_BYTES;
        const uint32_t p_hi = tc::smem_u32(sP + (g * 2) * AT_P_BYTES), p_lo = p_hi + AT_P_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(256 + g * AT_DH);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_BK / 16; ++k) {
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
      for (int t = 0; t < n; ++t) {
        if (t + 1 < n) issue_qk(0, t + 1);
        issue_pv(0, t);
        if (t + 1 < n) issue_qk(1, t + 1);
        issue_pv(1, t);
      }
      gt += n;
      ++qi;
    }
  } else if (warp >= 4) {
    // ===================== softmax / output: one query row per thread =====================
    const int g = (warp - 4) >> 2, quarter = warp & 3;
    const int row_in_blk = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + (uint32_t)(256 + g * AT_DH);
    unsigned char* prow_hi = sP + (g * 2) * AT_P_BYTES + row_in_blk * 128;
    unsigned char* prow_lo = prow_hi + AT_P_BYTES;
    int it = 0, gt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int par = it & 1;
      int qb, head, seq;
      item_coords(item, qb, head, seq);
      const int tok0 = seq * a.L;
      const int qrow = (qb * A2_BLOCKS + g) * AT_BQ + row_in_blk;  // position inside the sequence
      tc::mbar_wait(&km_full[par], (uint32_t)((it >> 1) & 1));
      const int n = ntiles[par];
      const uint32_t* km = kmask + par * (AT_MAX_L / 32);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int t = 0; t < n; ++t) {
        const int G = gt + t, buf = G & 1;
        tc::mbar_wait(&s_full[g * 2 + buf], (uint32_t)((G >> 1) & 1));
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
        float mx = -INFINITY;
        const uint32_t km0 = km[2 * t], km1 = km[2 * t + 1];
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
            tc::mbar_wait(&o_full[g], (uint32_t)((G - 1) & 1));  // P.V of the previous tile has landed; this tile's is not issued before our p_full
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
        tc::mbar_wait(&p_empty[g], (uint32_t)((G & 1) ^ 1));  // the previous P.V (possibly of the previous item) is done reading the P buffer
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
        l_run += rs;
      }
      // ---- output of the item: O / l -> bf16 (hi, lo) -> ctx
      float o0[32], o1[32];
      if (n > 0) {
        tc::mbar_wait(&o_full[g], (uint32_t)((gt + n - 1) & 1));
        tc::tc_fence_after();
        tc::tmem_ld_32x32(o_tmem, o0);
        tc::tmem_ld_32x32(o_tmem + 32u, o1);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&o_free[g]);  // the next item's first P.V may overwrite the accumulator
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o0[i] = 0.f, o1[i] = 0.f;
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&km_empty[par]);
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
            const float x0 = (c < 4 ? o0[c * 8 + 2 * j] : o1[(c - 4) * 8 + 2 * j]) * inv;
            const float x1 = (c < 4 ? o0[c * 8 + 2 * j + 1] : o1[(c - 4) * 8 + 2 * j + 1]) * inv;
            split2_bf16(x0, x1, hw[j], lw[j]);
          }
          ph4[c] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          pl4[c] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      gt += n;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace bert
}  // namespace capr
