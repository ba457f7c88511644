// K11 -- CEDR-KNRM head: masked cosine matrices of BERT hidden states + kernel pooling + combine.
//
//   CEDRKNRM_Class.masked_simmats / _cos_simmat   capreolus/reranker/CEDRKNRM.py:85-108
//   CEDRKNRM_Class.knrm                           capreolus/reranker/CEDRKNRM.py:110-136
//   CEDRKNRM_Class.forward (cls feature, combine) capreolus/reranker/CEDRKNRM.py:138-171
//
// The encoder (capr_bert_forward_hidden) leaves the fp32 hidden states of the requested layers in HBM, [n_layers][T][H].
// Per (passage, layer) the similarity work is 2 * 33 * 511 * 768 = 26 MFLOP against 96.6 GFLOP of encoder per passage
// (0.35 % for all 13 layers), so this head is plain fp32 CUDA-core code:
//   cedr_pool_kernel    one CTA per (passage, layer): query rows (first maxqlen+1 positions after [CLS]) are normalised
//                       into shared memory; each warp takes two document tokens at a time in registers (24 floats per lane
//                       each), dots them with every live query row (float4 LDS, warp-shuffle reduction) and lane k adds
//                       kernel k's exp(-(s-mu_k)^2 / (2 sigma_k^2)) to a per-warp [query][kernel] accumulator.
//                       Masked doc tokens (mask * (seg == 1) == 0) and masked query rows contribute exactly 0 (l.126).
//   cedr_finish_kernel  per (doc, layer): sum the partial soft-TFs over the doc's passages, log(clamp(., 1e-10)) * 0.01,
//                       sum over ALL maxqlen+1 query rows (masked rows contribute log(1e-10) * 0.01 like the reference);
//                       also the [CLS] feature (mean / max over passages of the last hidden state).
//   linear_rows / linear_out   `combine`: Linear(F, hidden) -> Linear(hidden, 1) (no activation in between), or Linear(F, 1).
#include "common.cuh"

namespace capr {

constexpr int CEDR_MAXQ = 64;    // maxqlen + 1 query rows
constexpr int CEDR_MAXK = 16;    // kernels (lane k and lane 16+k evaluate kernel k for the two doc tokens in flight)
constexpr int CEDR_MAXH = 1024;  // hidden size (lane owns 4 floats per 128-float slab: <= 8 slabs)
constexpr int CEDR_WARPS = 16;  // one CTA per SM (the query block is ~100 KB): 16 warps hide the global-load / shuffle latencies

struct CedrPoolArgs {
  const float* hidden;   // [n_layers][T][H]
  const long long* mask; // [n_seq][L]
  const long long* seg;  // [n_seq][L]
  int n_seq, L, H, P, Qm, K, n_layers;
  const float* mu;
  const float* sigma;
  float* partial;        // [n_layers][n_seq][Qm][K]
};

template <int SLABS>  // H <= 128 * SLABS
__global__ void __launch_bounds__(CEDR_WARPS * 32) cedr_pool_kernel(const CedrPoolArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qv = reinterpret_cast<float*>(smem_raw);                  // [Qm][SLABS*128] normalised query rows (zero padded)
  float* acc = qv + (size_t)a.Qm * SLABS * 128;                    // [CEDR_WARPS][Qm][CEDR_MAXK]
  float* qmask_own = acc + (size_t)CEDR_WARPS * a.Qm * CEDR_MAXK;  // [Qm] this passage's query mask (zeroes the cosine)
  float* qmask_p0 = qmask_own + CEDR_MAXQ;                         // [Qm] passage 0's query mask (multiplies the kernels, l.121,126)
  const int s = blockIdx.x, layer = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t T = (size_t)a.n_seq * a.L;
  const float* x = a.hidden + ((size_t)layer * T + (size_t)s * a.L) * a.H;
  const long long* mrow = a.mask + (size_t)s * a.L;
  const long long* srow = a.seg + (size_t)s * a.L;
  const int s0 = (s / a.P) * a.P;  // first passage of this document
  const int HP = SLABS * 128;

  for (int i = threadIdx.x; i < CEDR_WARPS * a.Qm * CEDR_MAXK; i += blockDim.x) acc[i] = 0.f;
  if (threadIdx.x < a.Qm) {
    const int tok = 1 + threadIdx.x;  // [CLS] is skipped (CEDRKNRM.py:112)
    const bool in = tok < a.L;
    qmask_own[threadIdx.x] = (in && mrow[tok] != 0 && srow[tok] == 0) ? 1.f : 0.f;
    qmask_p0[threadIdx.x] = (in && a.mask[(size_t)s0 * a.L + tok] != 0 && a.seg[(size_t)s0 * a.L + tok] == 0) ? 1.f : 0.f;
  }
  __syncthreads();
  // normalised query rows: x / (|x| + 1e-9) (CEDRKNRM.py:89-91); masked rows are zero vectors -> cosine 0
  for (int i = warp; i < a.Qm; i += CEDR_WARPS) {
    const float* row = x + (size_t)(1 + i) * a.H;
    const bool live = qmask_own[i] != 0.f;
    float4 v[SLABS];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < SLABS; ++c) {
      const int f = 4 * lane + 128 * c;
      v[c] = (live && f < a.H) ? *reinterpret_cast<const float4*>(row + f) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss = fmaf(v[c].x, v[c].x, ss), ss = fmaf(v[c].y, v[c].y, ss), ss = fmaf(v[c].z, v[c].z, ss), ss = fmaf(v[c].w, v[c].w, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / (sqrtf(ss) + 1e-9f);
#pragma unroll
    for (int c = 0; c < SLABS; ++c)
      *reinterpret_cast<float4*>(qv + (size_t)i * HP + 4 * lane + 128 * c) = make_float4(v[c].x * inv, v[c].y * inv, v[c].z * inv, v[c].w * inv);
  }
  __syncthreads();
  float mu = 0.f, cc = 0.f;
  if ((lane & 15) < a.K) {  // lanes k and 16+k both hold kernel k (K <= 16 on this path, checked on the host)
    const float sg = a.sigma[lane & 15];
    mu = a.mu[lane & 15];
    cc = -0.5f * 1.4426950408889634f / (sg * sg);
  }
  float* my_acc = acc + (size_t)warp * a.Qm * CEDR_MAXK;
  // document tokens 1..L-1 (position 0 is [CLS]); two per warp iteration share every query-row load
  for (int j0 = 1 + 2 * warp; j0 < a.L; j0 += 2 * CEDR_WARPS) {
    const int j1 = j0 + 1;
    const bool live0 = mrow[j0] != 0 && srow[j0] == 1;
    const bool live1 = j1 < a.L && mrow[j1] != 0 && srow[j1] == 1;
    if (!live0 && !live1) continue;  // warp-uniform
    float4 d0[SLABS], d1[SLABS];
    float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
    for (int c = 0; c < SLABS; ++c) {
      const int f = 4 * lane + 128 * c;
      const bool inH = f < a.H;
      d0[c] = (live0 && inH) ? *reinterpret_cast<const float4*>(x + (size_t)j0 * a.H + f) : make_float4(0.f, 0.f, 0.f, 0.f);
      d1[c] = (live1 && inH) ? *reinterpret_cast<const float4*>(x + (size_t)j1 * a.H + f) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss0 = fmaf(d0[c].x, d0[c].x, ss0), ss0 = fmaf(d0[c].y, d0[c].y, ss0), ss0 = fmaf(d0[c].z, d0[c].z, ss0), ss0 = fmaf(d0[c].w, d0[c].w, ss0);
      ss1 = fmaf(d1[c].x, d1[c].x, ss1), ss1 = fmaf(d1[c].y, d1[c].y, ss1), ss1 = fmaf(d1[c].z, d1[c].z, ss1), ss1 = fmaf(d1[c].w, d1[c].w, ss1);
    }
    ss0 = warp_sum(ss0), ss1 = warp_sum(ss1);
    const float inv0 = 1.0f / (sqrtf(ss0) + 1e-9f), inv1 = 1.0f / (sqrtf(ss1) + 1e-9f);
    for (int i = 0; i < a.Qm; ++i) {
      const float gate = qmask_p0[i];
      if (gate == 0.f) continue;  // warp-uniform: the kernel values of this row are multiplied by 0 (CEDRKNRM.py:126)
      const float4* q = reinterpret_cast<const float4*>(qv + (size_t)i * HP + 4 * lane);
      float p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int c = 0; c < SLABS; ++c) {
        const float4 qq = q[32 * c];
        p0 = fmaf(qq.x, d0[c].x, p0), p0 = fmaf(qq.y, d0[c].y, p0), p0 = fmaf(qq.z, d0[c].z, p0), p0 = fmaf(qq.w, d0[c].w, p0);
        p1 = fmaf(qq.x, d1[c].x, p1), p1 = fmaf(qq.y, d1[c].y, p1), p1 = fmaf(qq.z, d1[c].z, p1), p1 = fmaf(qq.w, d1[c].w, p1);
      }
      // two reductions for the price of one: after the first exchange lanes 0-15 carry p0, lanes 16-31 carry p1
      const bool upper = (lane & 16) != 0;
      float v = upper ? p1 : p0;
      v += __shfl_xor_sync(0xffffffffu, upper ? p0 : p1, 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const float cosv = v * (upper ? inv1 : inv0);  // cosine (0 when this passage's own query mask zeroed the row)
      // lane k (p0) and lane 16+k (p1) evaluate kernel k
      const float adj = cosv - mu;
      float e = ((upper ? live1 : live0) && (lane & 15) < a.K) ? ex2_approx(cc * adj * adj) : 0.f;
      e += __shfl_down_sync(0xffffffffu, e, 16);
      if (lane < a.K) my_acc[i * CEDR_MAXK + lane] += e;
    }
  }
  __syncthreads();
  float* out = a.partial + ((size_t)layer * a.n_seq + s) * a.Qm * a.K;
  for (int idx = threadIdx.x; idx < a.Qm * a.K; idx += blockDim.x) {
    const int i = idx / a.K, k = idx - i * a.K;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < CEDR_WARPS; ++w) t += acc[((size_t)w * a.Qm + i) * CEDR_MAXK + k];  // fixed order
    out[idx] = t;
  }
}

struct CedrFinishArgs {
  const float* partial;  // [n_layers][n_seq][Qm][K]
  const float* last_hidden;  // [T][H] (nullable: no CLS feature)
  int B, P, L, H, Qm, K, n_layers, cls_mode;  // cls_mode: 0 none, 1 avg, 2 max
  float* feats;  // [B][F], F = (cls ? H : 0) + n_layers*K
};

__global__ void __launch_bounds__(256) cedr_finish_kernel(const CedrFinishArgs a) {
  const int b = blockIdx.x;
  const int cls_dim = a.cls_mode ? a.H : 0;
  const int F = cls_dim + a.n_layers * a.K;
  float* out = a.feats + (size_t)b * F;
  const int n_seq = a.B * a.P;
  for (int idx = threadIdx.x; idx < a.n_layers * a.K; idx += blockDim.x) {
    const int l = idx / a.K, k = idx - l * a.K;
    float tot = 0.f;
    for (int i = 0; i < a.Qm; ++i) {
      float sft = 0.f;
      for (int p = 0; p < a.P; ++p) sft += a.partial[(((size_t)l * n_seq + (size_t)b * a.P + p) * a.Qm + i) * a.K + k];
      tot += logf(fmaxf(sft, 1e-10f)) * 0.01f;  // CEDRKNRM.py:130
    }
    out[cls_dim + idx] = tot;
  }
  if (a.cls_mode) {
    for (int h = threadIdx.x; h < a.H; h += blockDim.x) {
      float v = a.cls_mode == 2 ? -INFINITY : 0.f;
      for (int p = 0; p < a.P; ++p) {
        const float c = a.last_hidden[((size_t)b * a.P + p) * a.L * a.H + h];  // position 0 = [CLS]
        v = a.cls_mode == 2 ? fmaxf(v, c) : v + c;
      }
      out[h] = a.cls_mode == 2 ? v : v / (float)a.P;
    }
  }
}

// `combine`: hvals[b][h] = W1[h] . x_b + b1[h] (one warp per (b, h); W1 rows are shared by all b through L2), then
// out[b] = w2 . hvals[b] + b2 (hidden > 0) or hvals[b][0] (hidden == 0, W1 [1,F]).  Fixed summation order.
__global__ void __launch_bounds__(256) linear_rows_kernel(const float* __restrict__ x, int F, const float* __restrict__ w1, const float* __restrict__ b1,
                                                          int rows, float* __restrict__ hvals) {
  const int h = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y, lane = threadIdx.x & 31;
  if (h >= rows) return;
  const float* w = w1 + (size_t)h * F;
  const float* xb = x + (size_t)b * F;
  float p = 0.f;
  for (int i = lane; i < F; i += 32) p = fmaf(w[i], xb[i], p);
  p = warp_sum(p);
  if (lane == 0) hvals[(size_t)b * rows + h] = p + b1[h];
}

__global__ void __launch_bounds__(128) linear_out_kernel(const float* __restrict__ hvals, int rows, int hidden, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, int B, float* __restrict__ out) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  if (hidden == 0) {
    if (lane == 0) out[b] = hvals[b];
    return;
  }
  float p = 0.f;
  for (int h = lane; h < rows; h += 32) p = fmaf(w2[h], hvals[(size_t)b * rows + h], p);
  p = warp_sum(p);
  if (lane == 0) out[b] = p + b2[0];
}

}  // namespace capr

using namespace capr;

extern "C" {

int capr_cedrknrm_feature_dim(int H, int n_layers, int K, int cls_mode) { return (cls_mode ? H : 0) + n_layers * K; }

static size_t cedr_partial_bytes(int n_seq, int maxqlen, int n_layers, int K) {
  return (((size_t)n_layers * n_seq * (maxqlen + 1) * K * sizeof(float)) + 255) & ~(size_t)255;
}

size_t capr_cedrknrm_workspace_bytes(int n_seq, int maxqlen, int n_layers, int K, int combine_hidden) {
  if (n_seq <= 0 || maxqlen <= 0 || n_layers < 0 || K <= 0 || combine_hidden < 0) return 0;
  return cedr_partial_bytes(n_seq, maxqlen, n_layers, K) + (size_t)n_seq * (combine_hidden > 0 ? combine_hidden : 1) * sizeof(float);
}

int capr_cedrknrm_head(const float* hidden, int n_layers, const float* last_hidden, const int64_t* mask, const int64_t* seg, int B, int P, int L,
                       int H, int maxqlen, const float* mu, const float* sigma, int K, int cls_mode, const float* w1, const float* b1,
                       int combine_hidden, const float* w2, const float* b2, float* feats, float* scores, void* workspace,
                       size_t workspace_bytes, capr_stream_t stream) {
  capr::DeviceGuard device_guard(mask);  // act on the device that owns the caller's buffers
  const char* fn = "capr_cedrknrm_head";
  CAPR_REQUIRE(B >= 0 && P > 0 && L > 1 && H > 0 && maxqlen > 0 && K > 0 && n_layers >= 0 && combine_hidden >= 0, CAPR_ERR_BAD_SHAPE,
               "%s: bad shape B=%d P=%d L=%d H=%d maxqlen=%d K=%d layers=%d", fn, B, P, L, H, maxqlen, K, n_layers);
  CAPR_REQUIRE(cls_mode >= 0 && cls_mode <= 2, CAPR_ERR_BAD_SHAPE, "%s: cls must be None, avg or max", fn);
  CAPR_REQUIRE(cls_mode != 0 || n_layers > 0, CAPR_ERR_BAD_SHAPE, "%s: invalid config: no simmat layers and no cls feature", fn);
  const int Qm = maxqlen + 1;  // [SEP] is counted as a query position (CEDRKNRM.py:78-80)
  CAPR_REQUIRE(Qm <= CEDR_MAXQ && K <= CEDR_MAXK && H <= CEDR_MAXH && H % 4 == 0, CAPR_ERR_UNSUPPORTED,
               "%s: needs maxqlen < %d, kernels <= %d, hidden <= %d and a multiple of 4", fn, CEDR_MAXQ, CEDR_MAXK, CEDR_MAXH);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(mask && seg && feats && (n_layers == 0 || (hidden && mu && sigma && workspace)) && (cls_mode == 0 || last_hidden), CAPR_ERR_BAD_POINTER,
               "%s: null pointer", fn);
  CAPR_REQUIRE(!scores || (w1 && b1 && (combine_hidden == 0 || (w2 && b2))), CAPR_ERR_BAD_POINTER, "%s: scores requested without combine weights", fn);
  const int n_seq = B * P;
  CAPR_REQUIRE(workspace && workspace_bytes >= capr_cedrknrm_workspace_bytes(n_seq, maxqlen, n_layers, K, combine_hidden), CAPR_ERR_BAD_SHAPE,
               "%s: workspace too small (capr_cedrknrm_workspace_bytes)", fn);
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  if (n_layers > 0) {
    CedrPoolArgs pa{hidden, (const long long*)mask, (const long long*)seg, n_seq, L, H, P, Qm, K, n_layers, mu, sigma, partial};
    const int slabs = (H + 127) / 128;
    const size_t smem = ((size_t)Qm * slabs * 128 + (size_t)CEDR_WARPS * Qm * CEDR_MAXK + 2 * CEDR_MAXQ) * sizeof(float);
    CAPR_REQUIRE(smem <= 232448, CAPR_ERR_UNSUPPORTED, "%s: maxqlen=%d x hidden=%d query block does not fit in shared memory", fn, maxqlen, H);
    dim3 grid(n_seq, n_layers);
#define CAPR_CEDR_LAUNCH(S)                                                                                              \
  do {                                                                                                                   \
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(cedr_pool_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    cedr_pool_kernel<S><<<grid, CEDR_WARPS * 32, smem, st>>>(pa);                                                        \
  } while (0)
    if (slabs <= 1) CAPR_CEDR_LAUNCH(1);
    else if (slabs <= 2) CAPR_CEDR_LAUNCH(2);
    else if (slabs <= 4) CAPR_CEDR_LAUNCH(4);
    else if (slabs <= 6) CAPR_CEDR_LAUNCH(6);
    else CAPR_CEDR_LAUNCH(8);
#undef CAPR_CEDR_LAUNCH
    CAPR_CHECK_CUDA(cudaGetLastError());
  }
  CedrFinishArgs fa{partial, last_hidden, B, P, L, H, Qm, K, n_layers, cls_mode, feats};
  cedr_finish_kernel<<<B, 256, 0, st>>>(fa);
  CAPR_CHECK_CUDA(cudaGetLastError());
  if (scores) {
    const int F = capr_cedrknrm_feature_dim(H, n_layers, K, cls_mode);
    const int rows = combine_hidden > 0 ? combine_hidden : 1;
    float* hvals = (float*)((unsigned char*)workspace + cedr_partial_bytes(n_seq, maxqlen, n_layers, K));
    linear_rows_kernel<<<dim3((rows + 7) / 8, B), 256, 0, st>>>(feats, F, w1, b1, rows, hvals);
    CAPR_CHECK_CUDA(cudaGetLastError());
    linear_out_kernel<<<(B + 3) / 4, 128, 0, st>>>(hvals, rows, combine_hidden, w2, b2, B, scores);
    CAPR_CHECK_CUDA(cudaGetLastError());
  }
  return CAPR_OK;
}

}  // extern "C"
