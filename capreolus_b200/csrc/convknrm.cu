// K8 -- ConvKNRM scoring.
//
//   ConvKNRM_class.forward          capreolus/reranker/ConvKNRM.py:43-77
//   StackedSimilarityMatrix         capreolus/reranker/common.py:187-221   cosine of n-gram representations, pads zeroed
//   RbfKernelBank                   capreolus/reranker/common.py:224-250   (pooling: knrm_tc_kernel)
//
// The reference embeds the tokens, runs Conv1d(E -> F, n) for n = 1..maxngram over the right-zero-padded sequence (no
// activation), builds maxngram^2 cosine matrices between query and document n-grams and kernel-pools each like KNRM.
//
// B200 formulation (three kernels, all HBM / L2 streaming work except the pooling):
//   1. capr_convknrm_project (one-time per weight version): the convolution is linear and the embedding is frozen
//      (ConvKNRM.py:18 non_trainable=True), so   conv_n(emb)[t] = b_n + sum_{u<n} P_{n,u}[tok[t+u]]   with the
//      PROJECTED TABLES  P_{n,u} = emb . W_n[:, :, u]^T  ([V,F] each, maxngram(maxngram+1)/2 of them, stored as one
//      [V, S*F] fp32 table).  250 MFLOP of convolution per pair become six 512-byte row gathers per token.
//   2. convknrm_reps_kernel: one warp per token position; sums the gathered rows + bias for every n, L2-normalises
//      (1/(|x|+1e-9), common.py:206-207) and writes the vector as bf16 (hi, lo) planes into a per-chunk REP TABLE in exactly
//      the layout knrm_tc_kernel gathers from (capr_table_prepare_bf16), plus the id arrays that index it (0 = <pad> ->
//      the all-zero row 0, so padded positions give cosine 0 like common.py:213-215).
//   3. one knrm_tc_kernel launch per view (nq, nd) with feats-only output, then convknrm_combine_kernel: reorder to the
//      reference's feature index k*VIEWS + view (ConvKNRM.py:66-76) and apply `combine`.
// The rep table is streamed through HBM (835 KB per pair for 3 x 544 rows of 128 dims): ~3.3 MB of traffic per pair.
#include "common.cuh"
#include <cuda_bf16.h>

extern "C" int capr_knrm_forward_tc(const int64_t*, const int64_t*, int, int, int, const void*, const void*, int, int, int, const float*,
                                    const float*, int, const float*, const float*, int, const float*, const float*, int, float*, float*,
                                    capr_stream_t);

namespace capr {

constexpr int CONV_MAX_NGRAM = 4;
__host__ __device__ inline int conv_slot(int n /*1-based*/, int u) { return n * (n - 1) / 2 + u; }
__host__ __device__ inline int conv_slots(int maxngram) { return maxngram * (maxngram + 1) / 2; }

// ---- 1. projected tables: P[v][slot(n,u)*F + f] = sum_e emb[v][e] * W_n[f][e][u] ----------------------------------------------
// Plain fp32 (one-time, 13.8 GFLOP for V=30k, E=300, 768 output columns): 64x64 output tile per CTA, 16x16 threads x 4x4
// register tile, K chunks of 16 through shared memory.
struct ProjArgs {
  const float* emb;  // [V,E]
  const float* w[CONV_MAX_NGRAM];  // convs.{n-1}.0.weight [F,E,n]
  int V, E, F, maxngram;
  float* proj;  // [V, S*F]
};

__global__ void __launch_bounds__(256) convknrm_project_kernel(const ProjArgs a) {
  __shared__ float As[16][64 + 1];  // [k][row]
  __shared__ float Bs[16][64 + 1];  // [k][col]
  const int N = conv_slots(a.maxngram) * a.F;
  const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.E; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      const int v = row0 + r, e = k0 + k;
      As[k][r] = (v < a.V && e < a.E) ? a.emb[(size_t)v * a.E + e] : 0.f;
    }
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int c = i >> 4, k = i & 15;
      const int col = col0 + c, e = k0 + k;
      float w = 0.f;
      if (col < N && e < a.E) {
        const int slot = col / a.F, f = col - slot * a.F;
        int n = 1;
        while (conv_slot(n + 1, 0) <= slot) ++n;
        const int u = slot - conv_slot(n, 0);
        w = a.w[n - 1][((size_t)f * a.E + e) * n + u];
      }
      Bs[k][c] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i], bv[i] = Bs[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = row0 + ty * 4 + i;
    if (v >= a.V) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col < N) a.proj[(size_t)v * N + col] = acc[i][j];
    }
  }
}

// ---- 2. n-gram representations of one token matrix [B,L] -> rows of the rep table ----------------------------------------------
struct RepArgs {
  const long long* toks;  // [B,L]
  int B, L, V, F, maxngram, pitch;
  const float* proj;                  // [V, S*F]
  const float* bias[CONV_MAX_NGRAM];  // convs.{n-1}.0.bias [F]
  __nv_bfloat16* hi;                  // rep table planes [rows, pitch]
  __nv_bfloat16* lo;
  long long row_base;                 // row of (n=0, b=0, t=0); row(n,b,t) = row_base + (n*B + b)*L + t
  long long* ids;                     // [maxngram][B][L]: row index, or 0 where toks == 0 (<pad>)
};

// One warp per (b, t).  Lane owns floats 4*lane + 128*j of the F-vector (float4 loads: 512 contiguous bytes per warp and
// row segment).  MAXJ = ceil(pitch / 128).
template <int MAXJ>
__global__ void __launch_bounds__(256) convknrm_reps_kernel(const RepArgs a) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)a.B * a.L) return;
  const int b = (int)(wid / a.L), t = (int)(wid - (long long)b * a.L);
  const long long* row_toks = a.toks + (size_t)b * a.L;
  const int SF = conv_slots(a.maxngram) * a.F;
  const long long tok0 = row_toks[t];
  // table rows of the tokens under the widest window (positions past the end are the conv's zero padding: no contribution)
  const float* src[CONV_MAX_NGRAM];
#pragma unroll
  for (int u = 0; u < CONV_MAX_NGRAM; ++u) {
    src[u] = nullptr;
    if (u < a.maxngram && t + u < a.L) {
      const long long tk = row_toks[t + u];  // the reference indexes the table with the raw id (0 = <pad> row included)
      if (tk >= 0 && tk < (long long)a.V) src[u] = a.proj + (size_t)tk * SF;
    }
  }
  if (lane == 0) {
    for (int n = 0; n < a.maxngram; ++n)
      a.ids[((size_t)n * a.B + b) * a.L + t] = tok0 != 0 ? a.row_base + ((long long)n * a.B + b) * a.L + t : 0;
  }
  if (tok0 == 0) return;  // <pad>: the view kernels read the zero row 0 instead
#pragma unroll
  for (int n = 1; n <= CONV_MAX_NGRAM; ++n) {
    if (n > a.maxngram) break;
    float4 x[MAXJ];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int f = 4 * lane + 128 * j;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < a.F) {
        x[j] = *reinterpret_cast<const float4*>(a.bias[n - 1] + f);
#pragma unroll
        for (int u = 0; u < CONV_MAX_NGRAM; ++u) {
          if (u < n && src[u] != nullptr) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(src[u] + conv_slot(n, u) * a.F + f));
            x[j].x += p.x, x[j].y += p.y, x[j].z += p.z, x[j].w += p.w;
          }
        }
        ss = fmaf(x[j].x, x[j].x, ss), ss = fmaf(x[j].y, x[j].y, ss), ss = fmaf(x[j].z, x[j].z, ss), ss = fmaf(x[j].w, x[j].w, ss);
      }
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / (sqrtf(ss) + 1e-9f);  // common.py:206-207
    const size_t row = (size_t)(a.row_base + ((long long)(n - 1) * a.B + b) * a.L + t) * a.pitch;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int f = 4 * lane + 128 * j;
      if (f < a.pitch) {  // columns F..pitch-1 are written as zeros
        const float y[4] = {x[j].x * inv, x[j].y * inv, x[j].z * inv, x[j].w * inv};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h[i] = __float2bfloat16_rn(y[i]);
          l[i] = __float2bfloat16_rn(y[i] - __bfloat162float(h[i]));
        }
        *reinterpret_cast<uint2*>(a.hi + row + f) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(a.lo + row + f) = *reinterpret_cast<const uint2*>(l);
      }
    }
  }
}

__global__ void zero_row_kernel(__nv_bfloat16* hi, __nv_bfloat16* lo, int pitch) {
  for (int i = threadIdx.x; i < pitch; i += blockDim.x) hi[i] = __float2bfloat16_rn(0.f), lo[i] = __float2bfloat16_rn(0.f);
}

// ---- 3. combine: feats_v [VIEWS][B][K] -> x[b][k*VIEWS + v] -> Linear (-> tanh -> Linear) (-> tanh) ---------------------------------
struct CombineArgs {
  const float* feats_v;
  int B, K, VIEWS, hidden, flags;
  const float *w1, *b1, *w2, *b2;
  float* scores;
  float* feats_out;  // [B, K*VIEWS] nullable
};

__global__ void __launch_bounds__(128) convknrm_combine_kernel(const CombineArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.B) return;
  const int NF = a.K * a.VIEWS;
  auto feat = [&](int i) {  // i = k*VIEWS + v
    const int k = i / a.VIEWS, v = i - k * a.VIEWS;
    return a.feats_v[((size_t)v * a.B + warp) * a.K + k];
  };
  if (a.feats_out)
    for (int i = lane; i < NF; i += 32) a.feats_out[(size_t)warp * NF + i] = feat(i);
  if (!a.scores) return;
  float out;
  if (a.hidden == 0) {
    float p = 0.f;
    for (int i = lane; i < NF; i += 32) p = fmaf(a.w1[i], feat(i), p);
    out = warp_sum(p) + a.b1[0];
  } else {
    float acc = 0.f;
    for (int h = 0; h < a.hidden; ++h) {
      float p = 0.f;
      for (int i = lane; i < NF; i += 32) p = fmaf(a.w1[(size_t)h * NF + i], feat(i), p);
      acc = fmaf(a.w2[h], tanhf(warp_sum(p) + a.b1[h]), acc);
    }
    out = acc + a.b2[0];
  }
  if (a.flags & CAPR_KNRM_SCORETANH) out = tanhf(out);
  if (lane == 0) a.scores[warp] = out;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct ConvWs {  // carve of the workspace for a chunk of `c` pairs
  size_t rows, plane_bytes, ids_q, ids_d, feats, total;
  ConvWs(int c, int Q, int D, int maxngram, int pitch, int K, int views) {
    rows = 1 + (size_t)maxngram * c * (Q + D);
    plane_bytes = align256(rows * pitch * sizeof(__nv_bfloat16));
    ids_q = align256((size_t)maxngram * c * Q * 8);
    ids_d = align256((size_t)maxngram * c * D * 8);
    feats = align256((size_t)views * c * K * 4);
    total = 2 * plane_bytes + ids_q + ids_d + feats;
  }
};

}  // namespace capr

using namespace capr;

extern "C" {

int capr_convknrm_proj_cols(int maxngram, int F) { return (maxngram <= 0 || F <= 0) ? 0 : conv_slots(maxngram) * F; }

int capr_convknrm_project(const float* emb, int V, int E, const float* const* conv_w, int maxngram, int F, float* proj, capr_stream_t stream) {
  capr::DeviceGuard device_guard(emb);  // act on the device that owns the caller's buffers
  const char* fn = "capr_convknrm_project";
  CAPR_REQUIRE(V > 0 && E > 0 && F > 0 && maxngram > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape V=%d E=%d F=%d maxngram=%d", fn, V, E, F, maxngram);
  CAPR_REQUIRE(maxngram <= CONV_MAX_NGRAM, CAPR_ERR_UNSUPPORTED, "%s: maxngram=%d > %d is not supported", fn, maxngram, CONV_MAX_NGRAM);
  CAPR_REQUIRE(emb && conv_w && proj, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  ProjArgs a{};
  a.emb = emb, a.V = V, a.E = E, a.F = F, a.maxngram = maxngram, a.proj = proj;
  for (int n = 0; n < maxngram; ++n) {
    CAPR_REQUIRE(conv_w[n], CAPR_ERR_BAD_POINTER, "%s: conv_w[%d] is null", fn, n);
    a.w[n] = conv_w[n];
  }
  const int N = conv_slots(maxngram) * F;
  dim3 grid((N + 63) / 64, (V + 63) / 64);
  convknrm_project_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

size_t capr_convknrm_workspace_bytes(int chunk, int Q, int D, int maxngram, int F, int K, int crossmatch) {
  if (chunk <= 0 || Q <= 0 || D <= 0 || maxngram <= 0 || F <= 0 || K <= 0) return 0;
  const int pitch = ((F + 15) / 16) * 16;
  return ConvWs(chunk, Q, D, maxngram, pitch, K, crossmatch ? maxngram * maxngram : maxngram).total;
}

int capr_convknrm_forward(const int64_t* query, const int64_t* doc, int B, int Q, int D, const float* proj, int V, int maxngram, int F,
                          const float* const* conv_b, int crossmatch, const float* mu, const float* sigma, int K, const float* w1,
                          const float* b1, int hidden, const float* w2, const float* b2, int flags, float* scores, float* feats_out,
                          void* workspace, size_t workspace_bytes, capr_stream_t stream) {
  capr::DeviceGuard device_guard(proj);  // act on the device that owns the caller's buffers
  const char* fn = "capr_convknrm_forward";
  CAPR_REQUIRE(B >= 0 && Q > 0 && D > 0 && V > 0 && F > 0 && maxngram > 0 && K > 0 && hidden >= 0, CAPR_ERR_BAD_SHAPE,
               "%s: bad shape B=%d Q=%d D=%d V=%d F=%d maxngram=%d K=%d", fn, B, Q, D, V, F, maxngram, K);
  CAPR_REQUIRE(maxngram <= CONV_MAX_NGRAM, CAPR_ERR_UNSUPPORTED, "%s: maxngram=%d > %d is not supported", fn, maxngram, CONV_MAX_NGRAM);
  CAPR_REQUIRE(F % 4 == 0 && F <= 320, CAPR_ERR_UNSUPPORTED, "%s: filters=%d must be a multiple of 4 and <= 320", fn, F);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(query && doc && proj && conv_b && mu && sigma && workspace, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(scores || feats_out, CAPR_ERR_BAD_POINTER, "%s: no output requested", fn);
  CAPR_REQUIRE(!scores || (w1 && b1 && (hidden == 0 || (w2 && b2))), CAPR_ERR_BAD_POINTER, "%s: scores requested without combine weights", fn);
  CAPR_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)proj & 15) == 0, CAPR_ERR_BAD_POINTER, "%s: workspace must be 256-byte, proj 16-byte aligned", fn);
  const int pitch = ((F + 15) / 16) * 16;
  const int views = crossmatch ? maxngram * maxngram : maxngram;
  // largest chunk the workspace holds (rep-table offsets are 32-bit in the pooling kernel: rows * pitch < 2^31)
  const long long max_rows = ((1ll << 31) - 1) / pitch - 1;
  long long lo_c = 0, hi_c = max_rows / ((long long)maxngram * (Q + D));
  if (hi_c > B) hi_c = B;
  while (lo_c < hi_c) {  // binary search: ConvWs(c).total is monotone in c
    const long long mid = (lo_c + hi_c + 1) / 2;
    if (ConvWs((int)mid, Q, D, maxngram, pitch, K, views).total <= workspace_bytes) lo_c = mid; else hi_c = mid - 1;
  }
  const long long chunk = lo_c;
  CAPR_REQUIRE(chunk >= 1, CAPR_ERR_BAD_SHAPE, "%s: workspace of %zu bytes is too small (capr_convknrm_workspace_bytes)", fn, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += (int)chunk) {
    const int c = (int)((B - b0 < chunk) ? B - b0 : chunk);
    ConvWs ws(c, Q, D, maxngram, pitch, K, views);
    unsigned char* p = (unsigned char*)workspace;
    __nv_bfloat16* hi = (__nv_bfloat16*)p;
    __nv_bfloat16* lo = (__nv_bfloat16*)(p + ws.plane_bytes);
    long long* ids_q = (long long*)(p + 2 * ws.plane_bytes);
    long long* ids_d = (long long*)(p + 2 * ws.plane_bytes + ws.ids_q);
    float* feats_v = (float*)(p + 2 * ws.plane_bytes + ws.ids_q + ws.ids_d);
    zero_row_kernel<<<1, 128, 0, st>>>(hi, lo, pitch);
    for (int side = 0; side < 2; ++side) {
      RepArgs r{};
      r.toks = (const long long*)(side == 0 ? query + (size_t)b0 * Q : doc + (size_t)b0 * D);
      r.B = c, r.L = side == 0 ? Q : D, r.V = V, r.F = F, r.maxngram = maxngram, r.pitch = pitch, r.proj = proj, r.hi = hi, r.lo = lo;
      for (int n = 0; n < maxngram; ++n) {
        CAPR_REQUIRE(conv_b[n], CAPR_ERR_BAD_POINTER, "%s: conv_b[%d] is null", fn, n);
        r.bias[n] = conv_b[n];
      }
      r.row_base = side == 0 ? 1 : 1 + (long long)maxngram * c * Q;
      r.ids = side == 0 ? ids_q : ids_d;
      const long long warps = (long long)c * r.L;
      const unsigned blocks = (unsigned)((warps + 7) / 8);
      if (pitch <= 128) convknrm_reps_kernel<1><<<blocks, 256, 0, st>>>(r);
      else if (pitch <= 256) convknrm_reps_kernel<2><<<blocks, 256, 0, st>>>(r);
      else convknrm_reps_kernel<3><<<blocks, 256, 0, st>>>(r);
      CAPR_CHECK_CUDA(cudaGetLastError());
    }
    for (int v = 0; v < views; ++v) {
      const int nq = crossmatch ? v / maxngram : v, nd = crossmatch ? v % maxngram : v;
      const int rc = capr_knrm_forward_tc((const int64_t*)(ids_q + (size_t)nq * c * Q), (const int64_t*)(ids_d + (size_t)nd * c * D), c, Q, D, hi, lo,
                                          (int)ws.rows, F, pitch, mu, sigma, K, nullptr, nullptr, 0, nullptr, nullptr, flags & 0xF00, nullptr,
                                          feats_v + (size_t)v * c * K, stream);
      if (rc != CAPR_OK) return rc;
    }
    CombineArgs ca{feats_v, c, K, views, hidden, flags, w1, b1, w2, b2, scores ? scores + b0 : nullptr,
                   feats_out ? feats_out + (size_t)b0 * K * views : nullptr};
    convknrm_combine_kernel<<<(c + 3) / 4, 128, 0, st>>>(ca);
    CAPR_CHECK_CUDA(cudaGetLastError());
  }
  return CAPR_OK;
}

}  // extern "C"
