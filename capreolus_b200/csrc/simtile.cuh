// Shared producer of the KNRM / DRMM / PACRR kernels: one CTA builds the query x doc similarity tile
// of one (query, doc) pair in shared memory.
//
//   SimilarityMatrix.forward                      capreolus/reranker/common.py:170-182
//     cosine_similarity_matrix                    :160-167   gather, norms, bmm, divide, pad zeroing
//     exact_match_matrix                          :155-158   OOV (negative id) exact match
//
// Math: s_ij = e(q_i).e(d_j) + [q_i < 0 and q_i == d_j], with e(t) the prepared
// (L2-normalised, see table.cu) row of token max(t,0); row 0 is all-zero, so pads and OOV ids give a
// cosine of exactly 0 without any masking pass.
//
// Tile: QT=32 query rows x DT=512 doc columns, 256 threads, 8x8 accumulators per thread (fp32 FFMA --
// the sigma=0.001 exact-match kernel of KNRM needs fp32-accurate cosines).  The query
// rows stay resident in shared memory for the whole pair; doc rows are gathered from the (L2-resident)
// table in K-chunks of 16 floats with cp.async into a 2-stage ring, 16-byte coalesced per row segment.
#pragma once
#include "common.cuh"

namespace capr {

constexpr int QT = 32;             // query rows per tile
constexpr int DT = 512;            // doc columns per tile
constexpr int KC = 16;             // embedding dims per pipeline stage
constexpr int DP = KC + 4;         // smem pitch of a staged doc row (80 B -> conflict-free LDS.128)
constexpr int NT = 256;            // threads per CTA
constexpr int SIM_ROWS = QT + 4;   // + zero halo rows (PACRR pads bottom/right with zeros, PACRR.py:64)
constexpr int SIM_PITCH = DT + 4;  // + zero halo columns
constexpr int MAX_PITCH = 448;     // widest prepared-table row the smem budget allows

struct SimTile {
  float* qs;   // [QT][pitch+4]    resident query rows
  float* ds;   // [2][DT][DP]      doc-row K-chunk ring
  float* sim;  // [SIM_ROWS][SIM_PITCH]
  int* qid;    // [QT]  token ids (for exact match / masks)
  int* did;    // [DT]
  int* qrow;   // [QT]  table rows
  int* drow;   // [DT]
};

__host__ __device__ inline size_t sim_tile_bytes(int pitch) {
  return ((size_t)QT * (pitch + 4) + 2 * DT * DP + SIM_ROWS * SIM_PITCH) * sizeof(float) + (2 * QT + 2 * DT) * sizeof(int);
}

__device__ __forceinline__ SimTile carve_sim_tile(unsigned char* base, int pitch) {
  SimTile s;
  float* f = reinterpret_cast<float*>(base);
  s.qs = f;
  f += QT * (pitch + 4);
  s.ds = f;
  f += 2 * DT * DP;
  s.sim = f;
  f += SIM_ROWS * SIM_PITCH;
  int* i = reinterpret_cast<int*>(f);
  s.qid = i;
  s.qrow = i + QT;
  s.did = i + 2 * QT;
  s.drow = i + 2 * QT + DT;
  return s;
}

// Zero the whole sim buffer once per CTA; the halo (rows >= QT, columns >= DT) is never written again.
__device__ __forceinline__ void clear_sim_tile(const SimTile& s, int tid) {
  for (int i = tid; i < SIM_ROWS * SIM_PITCH; i += NT) s.sim[i] = 0.f;
}

__device__ __forceinline__ void stage_query_ids(const SimTile& s, const long long* __restrict__ q, int Q, int V, int tid) {
  if (tid < QT) {
    long long id = tid < Q ? q[tid] : 0;
    s.qid[tid] = id_as_int(id);
    s.qrow[tid] = table_row(id, V);
  }
}
__device__ __forceinline__ void stage_doc_ids(const SimTile& s, const long long* __restrict__ d, int d0, int D, int V, int tid) {
  for (int i = tid; i < DT; i += NT) {
    int col = d0 + i;
    long long id = col < D ? d[col] : 0;
    s.did[i] = id_as_int(id);
    s.drow[i] = table_row(id, V);
  }
}

// cp.async the QT query rows (whole rows) into s.qs.  Caller commits.
__device__ __forceinline__ void issue_query_rows(const SimTile& s, const float* __restrict__ table, int pitch, int tid) {
  const int segs = pitch / 4;  // 16-byte segments per row
  const int qp = pitch + 4;
  for (int i = tid; i < QT * segs; i += NT) {
    int r = i / segs, c = i - r * segs;
    cp_async16_zfill(s.qs + r * qp + c * 4, table + (size_t)s.qrow[r] * pitch + c * 4, s.qrow[r] != 0);
  }
}

// GEMM over the staged ids: acc[i][j] = e(q[ty*8+i]) . e(d[tx + 64 j]).
// Precondition: ids staged + __syncthreads(); query rows issued and committed (any group).
__device__ __forceinline__ void sim_tile_gemm(const SimTile& s, const float* __restrict__ table, int pitch, int tid,
                                              float (&acc)[8][8]) {
  const int ty = tid >> 6;   // 0..3   query rows ty*8 .. ty*8+7   (warp-uniform)
  const int tx = tid & 63;   // 0..63  doc columns tx + 64*j
  const int qp = pitch + 4;
  const int nchunks = pitch / KC;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // this thread copies segment `seg` (16 B) of doc rows r0 + 64*i of every chunk
  const int r0 = tid >> 2, seg = tid & 3;
  const float* src[8];
  unsigned live = 0;  // bit i: row i is a real table row (pad / OOV rows are zero-filled, not read)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int trow = s.drow[r0 + 64 * i];
    src[i] = table + (size_t)trow * pitch + seg * 4;
    live |= (trow != 0 ? 1u : 0u) << i;
  }
  float* dst0 = s.ds + r0 * DP + seg * 4;

  auto issue_chunk = [&](int c) {
    float* dst = dst0 + (c & 1) * (DT * DP);
#pragma unroll
    for (int i = 0; i < 8; ++i) cp_async16_zfill(dst + i * 64 * DP, src[i] + c * KC, (live >> i) & 1u);
  };

  issue_chunk(0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      issue_chunk(c + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* qb = s.qs + (ty * 8) * qp + c * KC;
    const float* db = s.ds + (c & 1) * (DT * DP) + tx * DP;
#pragma unroll
    for (int kk = 0; kk < KC / 4; ++kk) {
      float4 qa[8], da[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) qa[i] = *reinterpret_cast<const float4*>(qb + i * qp + kk * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) da[j] = *reinterpret_cast<const float4*>(db + j * 64 * DP + kk * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = acc[i][j];
          a = fmaf(qa[i].x, da[j].x, a);
          a = fmaf(qa[i].y, da[j].y, a);
          a = fmaf(qa[i].z, da[j].z, a);
          a = fmaf(qa[i].w, da[j].w, a);
          acc[i][j] = a;
        }
    }
    __syncthreads();  // stage (c&1) is overwritten by chunk c+2
  }
}

// Store the accumulators into s.sim, adding the OOV exact match (common.py:155-158,179-181).
//
// Identical in-vocabulary tokens: the reference's fp32 self-cosine a.a / (|a|+1e-9)^2 is 1 +- 1 ulp with the sign
// set by rounding (measured on the N(0,1) table: 43 % below, 15 % equal, 43 % above 1.0), while its exact value is
// 1 - 2e-9/|a|.  We store exactly 1.0f for such cells (when the row is not a zero vector), so that every engine and
// every summation order agrees on them; DRMM, whose last regular bin tests `s < 1.0`, bins these cells by their
// exact-arithmetic value (drmm.cu).  See DESIGN.md "Exact matches".
__device__ __forceinline__ void store_sim_tile(const SimTile& s, int tid, const float (&acc)[8][8]) {
  const int ty = tid >> 6, tx = tid & 63;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int qi = s.qid[ty * 8 + i];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = tx + 64 * j;
      float v = acc[i][j];
      if (qi == s.did[col]) {
        if (qi < 0) v += 1.0f;
        else if (v > 0.5f) v = 1.0f;
      }
      s.sim[(ty * 8 + i) * SIM_PITCH + col] = v;
    }
  }
}

// Whole producer for one doc tile of one pair.  On return (after the trailing __syncthreads) s.sim holds
// the tile; columns >= D - d0 and rows >= Q are exactly 0.
__device__ __forceinline__ void build_sim_tile(const SimTile& s, const float* __restrict__ table, int pitch, int V,
                                               const long long* __restrict__ qids, int Q, const long long* __restrict__ dids,
                                               int d0, int D, bool load_query, int tid) {
  if (load_query) stage_query_ids(s, qids, Q, V, tid);
  stage_doc_ids(s, dids, d0, D, V, tid);
  __syncthreads();
  if (load_query) {
    issue_query_rows(s, table, pitch, tid);
    cp_async_commit();
  }
  float acc[8][8];
  sim_tile_gemm(s, table, pitch, tid, acc);
  store_sim_tile(s, tid, acc);
  __syncthreads();
}

}  // namespace capr
