// K5/K6 -- monoBERT / BERT-MaxP encoder behind the C ABI.
//
//   PTBERTMaxP_Class.predict_step                 capreolus/reranker/ptBERTMaxP.py:67-96
//     self.bert(ids, attention_mask, token_type_ids)[0][:, 1]                      :82
//   = HF transformers BertForSequenceClassification (third party; math restated in oracle/restated.py):
//     embeddings (word + position + token_type, LayerNorm eps 1e-12)
//     L x { QKV Linear, softmax(QK^T/sqrt(dh) + key mask) V, out Linear, +residual LayerNorm,
//           Linear H->I, erf-GELU, Linear I->H, +residual LayerNorm }
//     pooler tanh(Linear(x[CLS])), classifier Linear(H, 2)
//
// Kernels in this file: embedding+LayerNorm, fp32->(hi,lo) bf16 split, residual LayerNorm, fused masked attention
// (fp32 FFMA flash-style, v1), pooler+classifier; the Linear layers run on tcgen05 (bert_gemm.cuh).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "bert_attn.cuh"
#include "bert_attn2.cuh"
#include "bert_attn3.cuh"
#include "bert_gemm2.cuh"
#include "tmap.cuh"

namespace capr {
namespace bert {

// ---------------------------------------------------------------------------------------------------------------
// elementwise / normalisation kernels (HBM-bound, one warp per token row)
// ---------------------------------------------------------------------------------------------------------------
__global__ void split_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 h, l;
    split_bf16(x[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// LayerNorm of one row held as `per` values per lane (H <= 32*MAXPER); biased variance, eps inside the sqrt (torch).
template <int MAXPER>
__device__ __forceinline__ void warp_layernorm_store(float (&v)[MAXPER], int H, int lane, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                                                     __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPER; ++i) s += (lane + 32 * i < H) ? v[i] : 0.f;
  const float mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPER; ++i) {
    const float d = (lane + 32 * i < H) ? v[i] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < MAXPER; ++i) {
    const int c = lane + 32 * i;
    if (c < H) {
      const float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
      out_f32[c] = y;
      __nv_bfloat16 h, l;
      split_bf16(y, h, l);
      out_hi[c] = h;
      out_lo[c] = l;
    }
  }
}

constexpr int LN_MAXPER = 32;  // hidden size <= 1024

// x = LayerNorm(word[id] + pos[t] + type[seg]);  T = n_seq * L tokens, position = token index inside its sequence.
__global__ void __launch_bounds__(256) embed_ln_kernel(const long long* __restrict__ ids, const long long* __restrict__ seg, int T, int L,
                                                       int H, int vocab, int max_pos, int type_vocab, const float* __restrict__ word,
                                                       const float* __restrict__ pos, const float* __restrict__ type,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                       float* __restrict__ x, __nv_bfloat16* __restrict__ x_hi,
                                                       __nv_bfloat16* __restrict__ x_lo) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= T) return;
  long long id = ids[row], sg = seg[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  sg = sg < 0 ? 0 : (sg >= type_vocab ? type_vocab - 1 : sg);
  const int p = row % L;
  const float* w = word + (size_t)id * H;
  const float* pe = pos + (size_t)(p < max_pos ? p : max_pos - 1) * H;
  const float* te = type + (size_t)sg * H;
  float v[LN_MAXPER];
#pragma unroll
  for (int i = 0; i < LN_MAXPER; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < H ? (w[c] + te[c]) + pe[c] : 0.f;  // HF: inputs_embeds + token_type_embeddings, then + position_embeddings
  }
  warp_layernorm_store<LN_MAXPER>(v, H, lane, gamma, beta, eps, x + (size_t)row * H, x_hi + (size_t)row * H, x_lo + (size_t)row * H);
}

// x = LayerNorm(y)  (y already holds dense(out) + bias + residual from the GEMM epilogue)
__global__ void __launch_bounds__(256) ln_kernel(const float* __restrict__ y, int T, int H, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float eps, float* __restrict__ x,
                                                 __nv_bfloat16* __restrict__ x_hi, __nv_bfloat16* __restrict__ x_lo) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= T) return;
  const float* yr = y + (size_t)row * H;
  float v[LN_MAXPER];
#pragma unroll
  for (int i = 0; i < LN_MAXPER; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < H ? yr[c] : 0.f;
  }
  warp_layernorm_store<LN_MAXPER>(v, H, lane, gamma, beta, eps, x + (size_t)row * H, x_hi + (size_t)row * H, x_lo + (size_t)row * H);
}

// ---------------------------------------------------------------------------------------------------------------
// attention: ctx = softmax(Q K^T / sqrt(dh) + key_mask) V, fp32, one CTA per (sequence, head, 64-query block)
// ---------------------------------------------------------------------------------------------------------------
// qkv [T, 3H] fp32 (Q | K | V, each head-major inside H); mask [n_seq, L] (1 = attend); output split into bf16 planes
// [T, H] for the out-projection GEMM.  Online softmax over key tiles of 64; tiles past the last unmasked key are
// skipped (masked keys get weight exactly 0, as torch's finfo.min additive mask does after softmax).
constexpr int ATT_BQ = 64, ATT_BK = 64, ATT_THREADS = 256;

template <int DH>
constexpr size_t attention_smem_bytes() { return (size_t)(3 * DH * 64 + 64 * 64 + 64) * sizeof(float) + 16; }

template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const float* __restrict__ qkv, const long long* __restrict__ mask, int L,
                                                                int H, int heads, float scale_log2e, __nv_bfloat16* __restrict__ ctx_hi,
                                                                __nv_bfloat16* __restrict__ ctx_lo) {
  constexpr int DPT = DH / 16;  // output dims per thread
  extern __shared__ __align__(16) float att_smem[];
  float (*Qt)[ATT_BQ] = reinterpret_cast<float (*)[ATT_BQ]>(att_smem);                    // Q^T [k][row], pre-scaled by scale*log2(e)
  float (*Kt)[ATT_BK] = reinterpret_cast<float (*)[ATT_BK]>(att_smem + DH * ATT_BQ);      // K^T [k][key ^ swz(k)]
  float (*Vs)[DH] = reinterpret_cast<float (*)[DH]>(att_smem + 2 * DH * 64);              // V   [key][d]
  float (*Pt)[ATT_BQ] = reinterpret_cast<float (*)[ATT_BQ]>(att_smem + 3 * DH * 64);      // P^T [key][row ^ swz(key)]
  float* kbias = att_smem + 3 * DH * 64 + 64 * 64;
  int* s_kv_len = reinterpret_cast<int*>(kbias + 64);

  const int qblocks = (L + ATT_BQ - 1) / ATT_BQ;
  const int qb = blockIdx.x % qblocks, head = (blockIdx.x / qblocks) % heads, seq = blockIdx.x / (qblocks * heads);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // rows ty*4..+3, keys / dims tx*4..+3
  const size_t tok0 = (size_t)seq * L;
  const int ld = 3 * H;
  const float* Qg = qkv + tok0 * ld + head * DH;
  const float* Kg = Qg + H;
  const float* Vg = Qg + 2 * H;
  const long long* mrow = mask + (size_t)seq * L;

  if (tid == 0) *s_kv_len = 0;
  __syncthreads();
  int last = 0;
  for (int j = tid; j < L; j += ATT_THREADS)
    if (mrow[j] != 0) last = j + 1;
  if (last) atomicMax(s_kv_len, last);
  for (int i = tid; i < ATT_BQ * DH; i += ATT_THREADS) {
    const int r = i / DH, k = i - r * DH;
    const int row = qb * ATT_BQ + r;
    Qt[k][r] = row < L ? Qg[(size_t)row * ld + k] * scale_log2e : 0.f;
  }
  __syncthreads();
  const int kv_len = *s_kv_len;

  float m_run[4], l_run[4], o[4][DPT];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int d = 0; d < DPT; ++d) o[i][d] = 0.f;
  }

  for (int k0 = 0; k0 < kv_len; k0 += ATT_BK) {
    // coalesced global reads (consecutive threads = consecutive dims of one key); the XOR on the key's 4-group index
    // spreads the transposed stores over the banks while keeping 4 consecutive keys contiguous for the float4 reads
    for (int i = tid; i < ATT_BK * DH; i += ATT_THREADS) {
      const int c = i / DH, k = i - c * DH;
      const int key = k0 + c;
      const bool in = key < L;
      Kt[k][c ^ ((k & 15) << 2)] = in ? Kg[(size_t)key * ld + k] : 0.f;
      Vs[c][k] = in ? Vg[(size_t)key * ld + k] : 0.f;
    }
    if (tid < ATT_BK) {
      const int key = k0 + tid;
      kbias[tid] = (key < L && mrow[key] != 0) ? 0.f : -INFINITY;
    }
    __syncthreads();
    // S = Q K^T (already in log2 units)
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < DH; ++k) {
      const float4 q4 = *reinterpret_cast<const float4*>(&Qt[k][ty * 4]);
      const float4 k4 = *reinterpret_cast<const float4*>(&Kt[k][(tx ^ (k & 15)) << 2]);
      const float qa[4] = {q4.x, q4.y, q4.z, q4.w}, ka[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
    }
    // online softmax: a row is shared by the 16 threads with the same ty (one half warp)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] += kbias[tx * 4 + j];
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o_ = 8; o_ > 0; o_ >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o_));
      const float m_new = fmaxf(m_run[i], mx);
      const bool dead = m_new == -INFINITY;  // every key so far is masked
      const float corr = dead ? 1.f : ex2_approx(m_run[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = dead ? 0.f : ex2_approx(s[i][j] - m_new);
        s[i][j] = p;
        rs += p;
      }
#pragma unroll
      for (int o_ = 8; o_ > 0; o_ >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o_);
      l_run[i] = l_run[i] * corr + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int d = 0; d < DPT; ++d) o[i][d] *= corr;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int key = tx * 4 + j;
      *reinterpret_cast<float4*>(&Pt[key][(ty ^ tx) << 2]) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    }
    __syncthreads();
    // O += P V
#pragma unroll 8
    for (int c = 0; c < ATT_BK; ++c) {
      const float4 p4 = *reinterpret_cast<const float4*>(&Pt[c][(ty ^ (c >> 2)) << 2]);
      const float pa[4] = {p4.x, p4.y, p4.z, p4.w};
      float va[DPT];
#pragma unroll
      for (int d = 0; d < DPT; ++d) va[d] = Vs[c][tx * DPT + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < DPT; ++d) o[i][d] = fmaf(pa[i], va[d], o[i][d]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = qb * ATT_BQ + ty * 4 + i;
    if (row >= L) continue;
    const float inv = l_run[i] > 0.f ? 1.0f / l_run[i] : 0.f;
    const size_t off = (tok0 + row) * H + head * DH + tx * DPT;
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
      __nv_bfloat16 h, l;
      split_bf16(o[i][d] * inv, h, l);
      ctx_hi[off + d] = h;
      ctx_lo[off + d] = l;
    }
  }
}

template <int DH>
static int launch_attention(int grid, const float* qkv, const long long* mask, int L, int H, int heads, float scale_log2e,
                            __nv_bfloat16* ctx_hi, __nv_bfloat16* ctx_lo, cudaStream_t st) {
  CAPR_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attention_smem_bytes<DH>()));
  attention_kernel<DH><<<grid, ATT_THREADS, attention_smem_bytes<DH>(), st>>>(qkv, mask, L, H, heads, scale_log2e, ctx_hi, ctx_lo);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

// [CLS] rows (token 0 of every sequence) of the attention context planes and of the fp32 residual stream -> compact [n_seq, H]
// buffers: the input of the CLS-only tail of the last encoder layer.
__global__ void __launch_bounds__(256) gather_cls_kernel(const __nv_bfloat16* __restrict__ ctx_hi, const __nv_bfloat16* __restrict__ ctx_lo,
                                                         const float* __restrict__ x, int L, int H, __nv_bfloat16* __restrict__ c_hi,
                                                         __nv_bfloat16* __restrict__ c_lo, float* __restrict__ cx) {
  const size_t src = (size_t)blockIdx.x * L * H, dst = (size_t)blockIdx.x * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    c_hi[dst + i] = ctx_hi[src + i];
    c_lo[dst + i] = ctx_lo[src + i];
    cx[dst + i] = x[src + i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pooler + classifier on the [CLS] row of every sequence: logits[n, :] = Wc tanh(Wp x[n*L] + bp) + bc
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pooler_classifier_kernel(const float* __restrict__ x, int L, int H, const float* __restrict__ wp,
                                                                const float* __restrict__ bp, const float* __restrict__ wc,
                                                                const float* __restrict__ bc, int n_labels, float* __restrict__ logits) {
  extern __shared__ float sm[];  // [H] cls row, [H] pooled
  float* cls = sm;
  float* pooled = sm + H;
  const int seq = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xr = x + (size_t)seq * L * H;
  for (int i = tid; i < H; i += blockDim.x) cls[i] = xr[i];
  __syncthreads();
  for (int j = warp; j < H; j += blockDim.x >> 5) {
    const float* w = wp + (size_t)j * H;
    float p = 0.f;
    for (int k = lane; k < H; k += 32) p = fmaf(w[k], cls[k], p);
    p = warp_sum(p);
    if (lane == 0) pooled[j] = tanhf(p + bp[j]);
  }
  __syncthreads();
  for (int c = warp; c < n_labels; c += blockDim.x >> 5) {
    const float* w = wc + (size_t)c * H;
    float p = 0.f;
    for (int k = lane; k < H; k += 32) p = fmaf(w[k], pooled[k], p);
    p = warp_sum(p);
    if (lane == 0) logits[(size_t)seq * n_labels + c] = p + bc[c];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return tc::make_bf16_map(m, base, rows, cols, box_rows);
}

static int pick_bn(int N) {
  for (int bn = MAX_BN; bn >= 32; bn >>= 1)
    if (N % bn == 0) return bn;
  return 0;
}

struct Linear {  // one nn.Linear prepared for the tensor cores
  int N = 0, K = 0, BN = 0;
  __nv_bfloat16 *w_hi = nullptr, *w_lo = nullptr;
  float* bias = nullptr;
  CUtensorMap map_hi, map_lo;
  bool has_pair_maps = false;      // N % 256 == 0: {64, 128} boxes for the CTA-pair kernel (each CTA loads half of the B tile)
  CUtensorMap pair_hi, pair_lo;
};

struct Layer {
  Linear qkv, attn_out, ffn1, ffn2;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};

struct Model {
  capr_bert_config cfg;
  int mode;  // 1 = bf16, 3 = bf16x3
  float *word = nullptr, *pos = nullptr, *type = nullptr, *emb_g = nullptr, *emb_b = nullptr;
  std::vector<Layer> layers;
  float *pool_w = nullptr, *pool_b = nullptr, *cls_w = nullptr, *cls_b = nullptr;
  std::vector<void*> owned;
  int sms = 0;
  int device = -1;              // the device the weights live on (capr_bert_create); every later call runs there
  bool ffma_attention = false;  // CAPR_BERT_ATTENTION=ffma: force the fp32 CUDA-core attention (A/B tests)
  bool attention_v1 = false;    // CAPR_BERT_ATTENTION=v1: the first tensor-core attention (128 queries per CTA), for A/B tests
  bool attention_v2 = false;    // CAPR_BERT_ATTENTION=v2: one CTA per (sequence, head, 256 queries) instead of the persistent kernel
  bool gemm_pairs = true;       // CAPR_BERT_GEMM=1cta: force the one-CTA GEMM (A/B tests)
  int max_pairs = 0;            // co-resident CTA pairs of gemm2_kernel (cudaOccupancyMaxActiveClusters)
  int qk_products = 3;          // bf16 products of S = Q.K^T in attention_tc4_kernel (CAPR_BERT_QK_PRODUCTS=1|2|3, see bert_attn2.cuh)
};

static int dev_alloc(Model* m, void** p, size_t bytes) {
  CAPR_CHECK_CUDA(cudaMalloc(p, bytes));
  m->owned.push_back(*p);
  return CAPR_OK;
}
static int dev_copy_f32(Model* m, float** dst, const float* src, size_t n, cudaStream_t st) {
  int rc = dev_alloc(m, (void**)dst, n * sizeof(float));
  if (rc) return rc;
  CAPR_CHECK_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return CAPR_OK;
}

// Build one Linear from up to three stacked [Ni, K] fp32 weights (QKV fusion) and their biases.
static int make_linear(Model* m, Linear* lin, int K, const float* const* ws, const float* const* bs, const int* ns, int parts, cudaStream_t st) {
  int N = 0;
  for (int i = 0; i < parts; ++i) N += ns[i];
  lin->N = N;
  lin->K = K;
  lin->BN = pick_bn(N);
  CAPR_REQUIRE(lin->BN > 0 && K % BK == 0, CAPR_ERR_UNSUPPORTED, "capr_bert_create: Linear %dx%d needs N %% 32 == 0 and K %% 64 == 0", N, K);
  int rc;
  if ((rc = dev_alloc(m, (void**)&lin->w_hi, (size_t)N * K * 2))) return rc;
  if ((rc = dev_alloc(m, (void**)&lin->w_lo, (size_t)N * K * 2))) return rc;
  if ((rc = dev_alloc(m, (void**)&lin->bias, (size_t)N * 4))) return rc;
  size_t row = 0;
  for (int i = 0; i < parts; ++i) {
    const size_t n = (size_t)ns[i] * K;
    split_kernel<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(ws[i], n, lin->w_hi + row * K, lin->w_lo + row * K);
    CAPR_CHECK_CUDA(cudaGetLastError());
    CAPR_CHECK_CUDA(cudaMemcpyAsync(lin->bias + row, bs[i], (size_t)ns[i] * 4, cudaMemcpyDeviceToDevice, st));
    row += ns[i];
  }
  if ((rc = make_map(&lin->map_hi, lin->w_hi, N, K, lin->BN))) return rc;
  if ((rc = make_map(&lin->map_lo, lin->w_lo, N, K, lin->BN))) return rc;
  if (N % G2_BN == 0) {
    if ((rc = make_map(&lin->pair_hi, lin->w_hi, N, K, G2_BN / 2))) return rc;
    if ((rc = make_map(&lin->pair_lo, lin->w_lo, N, K, G2_BN / 2))) return rc;
    lin->has_pair_maps = true;
  }
  return CAPR_OK;
}

// Co-resident clusters of gemm2_kernel<MODE> (the persistent grid must not exceed it: tiles are statically partitioned).
template <int MODE>
static int pair_capacity(int sms) {
  static int cached = -1;
  if (cached >= 0) return cached;
  if (cudaFuncSetAttribute(gemm2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem<MODE>::BYTES) != cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sms & ~1), 1, 1);
  cfg.blockDim = dim3(G2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Gemm2Smem<MODE>::BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm2_kernel<MODE>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return cached = n;
}

template <int MODE>
static int launch_gemm(const Model* m, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const Linear& lin, int M, int epi, const float* resid,
                       float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t st) {
  GemmArgs g{M, lin.N, lin.K, lin.BN, epi, lin.bias, resid, out_f32, out_hi, out_lo};
  if (m->gemm_pairs && lin.has_pair_maps && M > BM) {
    const int cap = pair_capacity<MODE>(m->sms);
    if (cap > 0) {
      const int pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * (lin.N / G2_BN);
      const int pairs = pair_tiles < cap ? pair_tiles : cap;
      gemm2_kernel<MODE><<<2 * pairs, G2_THREADS, Gemm2Smem<MODE>::BYTES, st>>>(a_hi, a_lo, lin.pair_hi, lin.pair_lo, g);
      CAPR_CHECK_CUDA(cudaGetLastError());
      return CAPR_OK;
    }
  }
  const int tiles = ((M + BM - 1) / BM) * (lin.N / lin.BN);
  const int grid = tiles < m->sms ? tiles : m->sms;
  CAPR_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<MODE>::BYTES));
  gemm_kernel<MODE><<<grid, GEMM_THREADS, GemmSmem<MODE>::BYTES, st>>>(a_hi, a_lo, lin.map_hi, lin.map_lo, g);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

static int gemm(const Model* m, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const Linear& lin, int M, int epi, const float* resid,
                float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t st) {
  return m->mode == 3 ? launch_gemm<3>(m, a_hi, a_lo, lin, M, epi, resid, out_f32, out_hi, out_lo, st)
                      : launch_gemm<1>(m, a_hi, a_lo, lin, M, epi, resid, out_f32, out_hi, out_lo, st);
}

struct Workspace {
  float *x, *y, *qkv;  // qkv: fp32 [Tp,3H] (FFMA attention) or, aliased, two bf16 planes [Tp,3H] (tensor-core attention)
  __nv_bfloat16 *x_hi, *x_lo, *ctx_hi, *ctx_lo, *ffn_hi, *ffn_lo;
  // compact [Np, .] copies of the [CLS] rows (Np = n_seq rounded up to 128) for the CLS-only tail of the last layer
  float *cx, *cy;
  __nv_bfloat16 *cx_hi, *cx_lo, *cctx_hi, *cctx_lo, *cffn_hi, *cffn_lo;
};
static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
static size_t carve(const capr_bert_config& c, size_t T, size_t n_seq, unsigned char* base, Workspace* w) {
  const size_t Tp = (T + BM - 1) / BM * BM;
  const size_t Np = (n_seq + BM - 1) / BM * BM;
  const size_t H = c.hidden, I = c.intermediate;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  float* x = (float*)take(Tp * H * 4);
  float* y = (float*)take(Tp * H * 4);
  float* qkv = (float*)take(Tp * 3 * H * 4);
  __nv_bfloat16* x_hi = (__nv_bfloat16*)take(Tp * H * 2);
  __nv_bfloat16* x_lo = (__nv_bfloat16*)take(Tp * H * 2);
  __nv_bfloat16* ctx_hi = (__nv_bfloat16*)take(Tp * H * 2);
  __nv_bfloat16* ctx_lo = (__nv_bfloat16*)take(Tp * H * 2);
  __nv_bfloat16* ffn_hi = (__nv_bfloat16*)take(Tp * I * 2);
  __nv_bfloat16* ffn_lo = (__nv_bfloat16*)take(Tp * I * 2);
  float* cx = (float*)take(Np * H * 4);
  float* cy = (float*)take(Np * H * 4);
  __nv_bfloat16* cx_hi = (__nv_bfloat16*)take(Np * H * 2);
  __nv_bfloat16* cx_lo = (__nv_bfloat16*)take(Np * H * 2);
  __nv_bfloat16* cctx_hi = (__nv_bfloat16*)take(Np * H * 2);
  __nv_bfloat16* cctx_lo = (__nv_bfloat16*)take(Np * H * 2);
  __nv_bfloat16* cffn_hi = (__nv_bfloat16*)take(Np * I * 2);
  __nv_bfloat16* cffn_lo = (__nv_bfloat16*)take(Np * I * 2);
  if (w) *w = Workspace{x, y, qkv, x_hi, x_lo, ctx_hi, ctx_lo, ffn_hi, ffn_lo, cx, cy, cx_hi, cx_lo, cctx_hi, cctx_lo, cffn_hi, cffn_lo};
  return off;
}

}  // namespace bert
}  // namespace capr

using namespace capr;
using namespace capr::bert;

extern "C" {

int capr_bert_num_weights(const capr_bert_config* cfg) { return cfg ? 5 + 16 * cfg->layers + 4 : 0; }

int capr_bert_create(const capr_bert_config* cfg, const float* const* weights, int n_weights, int precision_mode, capr_stream_t stream,
                     capr_bert_t* out) {
  const char* fn = "capr_bert_create";
  CAPR_REQUIRE(cfg && weights && out, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(cfg->hidden > 0 && cfg->layers > 0 && cfg->heads > 0 && cfg->intermediate > 0 && cfg->vocab > 0 && cfg->max_pos > 0 &&
                   cfg->type_vocab > 0 && cfg->n_labels > 0,
               CAPR_ERR_BAD_SHAPE, "%s: bad config", fn);
  CAPR_REQUIRE(n_weights == capr_bert_num_weights(cfg), CAPR_ERR_BAD_SHAPE, "%s: expected %d weight pointers, got %d", fn,
               capr_bert_num_weights(cfg), n_weights);
  CAPR_REQUIRE(precision_mode == CAPR_BERT_BF16 || precision_mode == CAPR_BERT_BF16X3, CAPR_ERR_BAD_SHAPE, "%s: unknown precision mode %d", fn, precision_mode);
  CAPR_REQUIRE(cfg->hidden % cfg->heads == 0, CAPR_ERR_BAD_SHAPE, "%s: hidden %% heads != 0", fn);
  const int dh = cfg->hidden / cfg->heads;
  CAPR_REQUIRE(dh == 16 || dh == 32 || dh == 64, CAPR_ERR_UNSUPPORTED, "%s: head dim %d not in {16,32,64}", fn, dh);
  CAPR_REQUIRE(cfg->hidden <= 32 * LN_MAXPER && cfg->hidden % 64 == 0 && cfg->intermediate % 64 == 0, CAPR_ERR_UNSUPPORTED,
               "%s: hidden must be a multiple of 64 and <= %d, intermediate a multiple of 64", fn, 32 * LN_MAXPER);
  for (int i = 0; i < n_weights; ++i) CAPR_REQUIRE(weights[i], CAPR_ERR_BAD_POINTER, "%s: weight %d is null", fn, i);
  cudaStream_t st = (cudaStream_t)stream;
  capr::DeviceGuard device_guard(weights[0]);  // allocate and build the snapshot on the device that owns the weights
  Model* m = new (std::nothrow) Model();
  CAPR_REQUIRE(m, CAPR_ERR_CUDA, "%s: out of host memory", fn);
  m->cfg = *cfg;
  m->device = capr::device_of(weights[0]);
  m->mode = precision_mode == CAPR_BERT_BF16X3 ? 3 : 1;
  m->sms = sm_count();
  {
    const char* e = getenv("CAPR_BERT_ATTENTION");
    m->ffma_attention = e && e[0] == 'f';
    m->attention_v1 = e && e[0] == 'v' && e[1] == '1';
    m->attention_v2 = e && e[0] == 'v' && e[1] == '2';
    const char* ge = getenv("CAPR_BERT_GEMM");
    m->gemm_pairs = !(ge && ge[0] == '1');
    const char* qe = getenv("CAPR_BERT_QK_PRODUCTS");
    m->qk_products = (qe && qe[0] >= '1' && qe[0] <= '3') ? qe[0] - '0' : 3;
  }
  const size_t H = cfg->hidden, I = cfg->intermediate;
  int rc = CAPR_OK;
#define CAPR_TRY(expr)            \
  do {                            \
    if ((rc = (expr)) != CAPR_OK) { \
      capr_bert_destroy((capr_bert_t)m); \
      return rc;                  \
    }                             \
  } while (0)
  if (m->sms <= 0) {
    delete m;
    set_error("%s: no CUDA device", fn);
    return CAPR_ERR_NO_DEVICE;
  }
  const float* const* w = weights;
  CAPR_TRY(dev_copy_f32(m, &m->word, w[0], (size_t)cfg->vocab * H, st));
  CAPR_TRY(dev_copy_f32(m, &m->pos, w[1], (size_t)cfg->max_pos * H, st));
  CAPR_TRY(dev_copy_f32(m, &m->type, w[2], (size_t)cfg->type_vocab * H, st));
  CAPR_TRY(dev_copy_f32(m, &m->emb_g, w[3], H, st));
  CAPR_TRY(dev_copy_f32(m, &m->emb_b, w[4], H, st));
  m->layers.resize(cfg->layers);
  for (int l = 0; l < cfg->layers; ++l) {
    const float* const* p = w + 5 + 16 * l;
    Layer& L = m->layers[l];
    const float* qkv_w[3] = {p[0], p[2], p[4]};
    const float* qkv_b[3] = {p[1], p[3], p[5]};
    const int hs[3] = {(int)H, (int)H, (int)H};
    CAPR_TRY(make_linear(m, &L.qkv, (int)H, qkv_w, qkv_b, hs, 3, st));
    const int h1[1] = {(int)H}, i1[1] = {(int)I};
    CAPR_TRY(make_linear(m, &L.attn_out, (int)H, p + 6, p + 7, h1, 1, st));
    CAPR_TRY(dev_copy_f32(m, &L.ln1_g, p[8], H, st));
    CAPR_TRY(dev_copy_f32(m, &L.ln1_b, p[9], H, st));
    CAPR_TRY(make_linear(m, &L.ffn1, (int)H, p + 10, p + 11, i1, 1, st));
    CAPR_TRY(make_linear(m, &L.ffn2, (int)I, p + 12, p + 13, h1, 1, st));
    CAPR_TRY(dev_copy_f32(m, &L.ln2_g, p[14], H, st));
    CAPR_TRY(dev_copy_f32(m, &L.ln2_b, p[15], H, st));
  }
  const float* const* t = w + 5 + 16 * cfg->layers;
  CAPR_TRY(dev_copy_f32(m, &m->pool_w, t[0], H * H, st));
  CAPR_TRY(dev_copy_f32(m, &m->pool_b, t[1], H, st));
  CAPR_TRY(dev_copy_f32(m, &m->cls_w, t[2], (size_t)cfg->n_labels * H, st));
  CAPR_TRY(dev_copy_f32(m, &m->cls_b, t[3], cfg->n_labels, st));
#undef CAPR_TRY
  *out = (capr_bert_t)m;
  return CAPR_OK;
}

void capr_bert_destroy(capr_bert_t h) {
  Model* m = (Model*)h;
  if (!m) return;
  for (void* p : m->owned) cudaFree(p);
  delete m;
}

size_t capr_bert_workspace_bytes(capr_bert_t h, int n_seq, int L) {
  Model* m = (Model*)h;
  if (!m || n_seq <= 0 || L <= 0) return 0;
  return carve(m->cfg, (size_t)n_seq * L, (size_t)n_seq, nullptr, nullptr);
}

// Shared body of capr_bert_forward / capr_bert_forward_hidden.  hidden_layers[i] in [0, layers]: 0 = embedding output,
// l = output of encoder layer l (HF `hidden_states[l]`); the fp32 activations are copied to hidden_out[i] ([T,H] each).
// x_in (nullable): fp32 [T,H] input of the first encoder layer; replaces the embedding stage (ids / seg are then unused) --
// the PARADE aggregator feeds passage [CLS] vectors to two stand-alone BertLayers (ptparade.py:55-68).
static int bert_run(const char* fn, capr_bert_t h, const int64_t* ids, const int64_t* mask, const int64_t* seg, int n_seq, int L, float* logits,
                    const int* hidden_layers, int n_hidden, float* hidden_out, void* workspace, size_t workspace_bytes, capr_stream_t stream,
                    const float* x_in = nullptr) {
  Model* m = (Model*)h;
  CAPR_REQUIRE(m, CAPR_ERR_BAD_POINTER, "%s: null handle", fn);
  capr::DeviceGuard device_guard(m->device);
  CAPR_REQUIRE(n_seq >= 0 && L > 0 && n_hidden >= 0, CAPR_ERR_BAD_SHAPE, "%s: n_seq=%d L=%d", fn, n_seq, L);
  if (n_seq == 0) return CAPR_OK;
  CAPR_REQUIRE(x_in || L <= m->cfg.max_pos, CAPR_ERR_BAD_SHAPE, "%s: sequence length %d exceeds max_position_embeddings %d", fn, L, m->cfg.max_pos);
  CAPR_REQUIRE((x_in || (ids && seg)) && mask && (logits || n_hidden > 0) && workspace, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(n_hidden == 0 || (hidden_layers && hidden_out), CAPR_ERR_BAD_POINTER, "%s: hidden states requested without buffers", fn);
  for (int i = 0; i < n_hidden; ++i)
    CAPR_REQUIRE(hidden_layers[i] >= 0 && hidden_layers[i] <= m->cfg.layers, CAPR_ERR_BAD_SHAPE, "%s: hidden layer %d not in [0, %d]", fn, hidden_layers[i], m->cfg.layers);
  CAPR_REQUIRE(((uintptr_t)workspace & 255) == 0, CAPR_ERR_BAD_POINTER, "%s: workspace must be 256-byte aligned", fn);
  const size_t T = (size_t)n_seq * L;
  CAPR_REQUIRE(T < (size_t)1 << 31, CAPR_ERR_BAD_SHAPE, "%s: too many tokens in one call (%zu); split the batch", fn, T);
  Workspace ws;
  const size_t need = carve(m->cfg, T, (size_t)n_seq, (unsigned char*)workspace, &ws);
  CAPR_REQUIRE(workspace_bytes >= need, CAPR_ERR_BAD_SHAPE, "%s: workspace too small (%zu < %zu)", fn, workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int H = m->cfg.hidden, I = m->cfg.intermediate, heads = m->cfg.heads, dh = H / heads;
  const int Ti = (int)T;
  const size_t Tp = (T + BM - 1) / BM * BM;
  int rc;
  CUtensorMap x_hi, x_lo, c_hi, c_lo, f_hi, f_lo;
  if ((rc = make_map(&x_hi, ws.x_hi, Tp, H, BM))) return rc;
  if ((rc = make_map(&x_lo, ws.x_lo, Tp, H, BM))) return rc;
  if ((rc = make_map(&c_hi, ws.ctx_hi, Tp, H, BM))) return rc;
  if ((rc = make_map(&c_lo, ws.ctx_lo, Tp, H, BM))) return rc;
  if ((rc = make_map(&f_hi, ws.ffn_hi, Tp, I, BM))) return rc;
  if ((rc = make_map(&f_lo, ws.ffn_lo, Tp, I, BM))) return rc;
  if (Tp > T) {  // rows past T of the A operands are read by TMA: keep them finite
    CAPR_CHECK_CUDA(cudaMemsetAsync(ws.x_hi + T * H, 0, (Tp - T) * H * 2, st));
    CAPR_CHECK_CUDA(cudaMemsetAsync(ws.x_lo + T * H, 0, (Tp - T) * H * 2, st));
    CAPR_CHECK_CUDA(cudaMemsetAsync(ws.ctx_hi + T * H, 0, (Tp - T) * H * 2, st));
    CAPR_CHECK_CUDA(cudaMemsetAsync(ws.ctx_lo + T * H, 0, (Tp - T) * H * 2, st));
    CAPR_CHECK_CUDA(cudaMemsetAsync(ws.ffn_hi + T * I, 0, (Tp - T) * I * 2, st));
    CAPR_CHECK_CUDA(cudaMemsetAsync(ws.ffn_lo + T * I, 0, (Tp - T) * I * 2, st));
  }
  const int row_blocks = (Ti + 7) / 8;  // 8 warps (rows) per 256-thread CTA
  if (x_in) {
    CAPR_CHECK_CUDA(cudaMemcpyAsync(ws.x, x_in, T * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const size_t n = T * H;
    split_kernel<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(x_in, n, ws.x_hi, ws.x_lo);
  } else {
    embed_ln_kernel<<<row_blocks, 256, 0, st>>>((const long long*)ids, (const long long*)seg, Ti, L, H, m->cfg.vocab, m->cfg.max_pos,
                                                m->cfg.type_vocab, m->word, m->pos, m->type, m->emb_g, m->emb_b, m->cfg.ln_eps, ws.x, ws.x_hi,
                                                ws.x_lo);
  }
  CAPR_CHECK_CUDA(cudaGetLastError());
  auto emit_hidden = [&](int layer_index) -> int {
    for (int i = 0; i < n_hidden; ++i)
      if (hidden_layers[i] == layer_index)
        CAPR_CHECK_CUDA(cudaMemcpyAsync(hidden_out + (size_t)i * T * H, ws.x, T * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return CAPR_OK;
  };
  if ((rc = emit_hidden(0))) return rc;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)dh);
  const int att_grid = n_seq * heads * ((L + ATT_BQ - 1) / ATT_BQ);
  // tensor-core attention: head dim 64, L <= 512; Q/K/V then live as bf16 (hi, lo) planes aliased onto the fp32 qkv buffer
  const bool tc_attention = dh == AT_DH && L <= AT_MAX_L && !m->ffma_attention;
  __nv_bfloat16* qkv_hi = reinterpret_cast<__nv_bfloat16*>(ws.qkv);
  __nv_bfloat16* qkv_lo = qkv_hi + Tp * 3 * H;
  CUtensorMap q_hi, q_lo, kv_hi, kv_lo;
  if (tc_attention) {
    if ((rc = make_map(&q_hi, qkv_hi, Tp, 3 * H, AT_BQ))) return rc;
    if ((rc = make_map(&q_lo, qkv_lo, Tp, 3 * H, AT_BQ))) return rc;
    if ((rc = make_map(&kv_hi, qkv_hi, Tp, 3 * H, AT_BK))) return rc;
    if ((rc = make_map(&kv_lo, qkv_lo, Tp, 3 * H, AT_BK))) return rc;
    if (Tp > T) {  // K/V tiles may reach into the pad rows: they are multiplied by P = 0, so they must be finite
      CAPR_CHECK_CUDA(cudaMemsetAsync(qkv_hi + T * 3 * H, 0, (Tp - T) * 3 * H * 2, st));
      CAPR_CHECK_CUDA(cudaMemsetAsync(qkv_lo + T * 3 * H, 0, (Tp - T) * 3 * H * 2, st));
    }
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM));
    CAPR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A4_SMEM));
  }
#ifdef CAPR_DEBUG_BUILD
  const char* adbg = getenv("CAPR_ATTN_DEBUG");
  // CAPR_ATTN_TRACE=<device pointer, decimal>: 256 int64 clock stamps of one CTA of attention_tc2_kernel (scripts/attn_trace.py)
  const char* atr = getenv("CAPR_ATTN_TRACE");
#else
  const char* adbg = nullptr;
  const char* atr = nullptr;
#endif
  const Attn2Args at2{L, H, heads, n_seq, (long long)Tp, scale_log2e, (const long long*)mask, ws.ctx_hi, ws.ctx_lo, adbg ? atoi(adbg) : 0,
                      atr ? (long long*)strtoull(atr, nullptr, 10) : nullptr, m->qk_products, 0};
  const int att_tc2_grid = n_seq * heads * ((L + A2_BLOCKS * AT_BQ - 1) / (A2_BLOCKS * AT_BQ));
  const AttnArgs at{L, H, heads, n_seq, scale_log2e, (const long long*)mask, ws.ctx_hi, ws.ctx_lo};
  const int att_tc_grid = n_seq * heads * ((L + AT_BQ - 1) / AT_BQ);
  // CLS-only tail: a classification forward consumes row 0 of the last layer only (pooler -> classifier; ptBERTMaxP.py:82,
  // SURVEY.md App. B "Only row 0 ([CLS]) of the last layer's output is consumed").  The last layer still needs K / V of every
  // token, but its attention runs for the first query block alone and out-projection, both LayerNorms and the FFN (10 of the 12
  // weight-matrix units of a layer) run on the n_seq [CLS] rows gathered into compact buffers.  Off when hidden states of the
  // last layer are requested (CEDR-KNRM, PARADE) or when an A/B attention kernel is selected; CAPR_BERT_CLS_ONLY=0 disables it.
  const char* cls_env = getenv("CAPR_BERT_CLS_ONLY");
  const bool cls_only = logits && n_hidden == 0 && tc_attention && !m->attention_v1 && !m->attention_v2 &&
                        !(cls_env && cls_env[0] == '0');
  const size_t Np = ((size_t)n_seq + BM - 1) / BM * BM;
  CUtensorMap cx_hi, cx_lo, cc_hi, cc_lo, cf_hi, cf_lo;
  if (cls_only) {
    if ((rc = make_map(&cx_hi, ws.cx_hi, Np, H, BM))) return rc;
    if ((rc = make_map(&cx_lo, ws.cx_lo, Np, H, BM))) return rc;
    if ((rc = make_map(&cc_hi, ws.cctx_hi, Np, H, BM))) return rc;
    if ((rc = make_map(&cc_lo, ws.cctx_lo, Np, H, BM))) return rc;
    if ((rc = make_map(&cf_hi, ws.cffn_hi, Np, I, BM))) return rc;
    if ((rc = make_map(&cf_lo, ws.cffn_lo, Np, I, BM))) return rc;
    if (Np > (size_t)n_seq) {  // pad rows of the compact A operands are read by TMA: keep them finite
      const size_t pad = Np - (size_t)n_seq, at = (size_t)n_seq;
      CAPR_CHECK_CUDA(cudaMemsetAsync(ws.cx_hi + at * H, 0, pad * H * 2, st));
      CAPR_CHECK_CUDA(cudaMemsetAsync(ws.cx_lo + at * H, 0, pad * H * 2, st));
      CAPR_CHECK_CUDA(cudaMemsetAsync(ws.cctx_hi + at * H, 0, pad * H * 2, st));
      CAPR_CHECK_CUDA(cudaMemsetAsync(ws.cctx_lo + at * H, 0, pad * H * 2, st));
      CAPR_CHECK_CUDA(cudaMemsetAsync(ws.cffn_hi + at * I, 0, pad * I * 2, st));
      CAPR_CHECK_CUDA(cudaMemsetAsync(ws.cffn_lo + at * I, 0, pad * I * 2, st));
    }
  }
  const int cls_row_blocks = (n_seq + 7) / 8;
  for (int l = 0; l < m->cfg.layers; ++l) {
    const Layer& ly = m->layers[l];
    const bool tail = cls_only && l + 1 == m->cfg.layers;
    if (tc_attention) {
      if ((rc = gemm(m, x_hi, x_lo, ly.qkv, Ti, EPI_BIAS_SPLIT, nullptr, nullptr, qkv_hi, qkv_lo, st))) return rc;
      if (m->attention_v1) attention_tc_kernel<<<att_tc_grid, AT_THREADS, AT_SMEM, st>>>(q_hi, q_lo, kv_hi, kv_lo, at);
      else if (m->attention_v2) attention_tc2_kernel<<<att_tc2_grid, A2_THREADS, A2_SMEM, st>>>(q_hi, q_lo, kv_hi, kv_lo, at2);
      else if (tail) {
        Attn2Args at_cls = at2;
        at_cls.q_blocks = 1;  // only block 0 holds the [CLS] query
        const int items = n_seq * heads;
        attention_tc4_kernel<<<items < m->sms ? items : m->sms, A4_THREADS, A4_SMEM, st>>>(q_hi, q_lo, kv_hi, kv_lo, at_cls);
      } else attention_tc4_kernel<<<att_tc2_grid < m->sms ? att_tc2_grid : m->sms, A4_THREADS, A4_SMEM, st>>>(q_hi, q_lo, kv_hi, kv_lo, at2);
      CAPR_CHECK_CUDA(cudaGetLastError());
    } else {
      if ((rc = gemm(m, x_hi, x_lo, ly.qkv, Ti, EPI_BIAS_F32, nullptr, ws.qkv, nullptr, nullptr, st))) return rc;
      if (dh == 64) rc = launch_attention<64>(att_grid, ws.qkv, (const long long*)mask, L, H, heads, scale_log2e, ws.ctx_hi, ws.ctx_lo, st);
      else if (dh == 32) rc = launch_attention<32>(att_grid, ws.qkv, (const long long*)mask, L, H, heads, scale_log2e, ws.ctx_hi, ws.ctx_lo, st);
      else rc = launch_attention<16>(att_grid, ws.qkv, (const long long*)mask, L, H, heads, scale_log2e, ws.ctx_hi, ws.ctx_lo, st);
      if (rc) return rc;
    }
    if (tail) {
      gather_cls_kernel<<<n_seq, 256, 0, st>>>(ws.ctx_hi, ws.ctx_lo, ws.x, L, H, ws.cctx_hi, ws.cctx_lo, ws.cx);
      CAPR_CHECK_CUDA(cudaGetLastError());
      if ((rc = gemm(m, cc_hi, cc_lo, ly.attn_out, n_seq, EPI_BIAS_RESID_F32, ws.cx, ws.cy, nullptr, nullptr, st))) return rc;
      ln_kernel<<<cls_row_blocks, 256, 0, st>>>(ws.cy, n_seq, H, ly.ln1_g, ly.ln1_b, m->cfg.ln_eps, ws.cx, ws.cx_hi, ws.cx_lo);
      CAPR_CHECK_CUDA(cudaGetLastError());
      if ((rc = gemm(m, cx_hi, cx_lo, ly.ffn1, n_seq, EPI_BIAS_GELU_SPLIT, nullptr, nullptr, ws.cffn_hi, ws.cffn_lo, st))) return rc;
      if ((rc = gemm(m, cf_hi, cf_lo, ly.ffn2, n_seq, EPI_BIAS_RESID_F32, ws.cx, ws.cy, nullptr, nullptr, st))) return rc;
      ln_kernel<<<cls_row_blocks, 256, 0, st>>>(ws.cy, n_seq, H, ly.ln2_g, ly.ln2_b, m->cfg.ln_eps, ws.cx, ws.cx_hi, ws.cx_lo);
      CAPR_CHECK_CUDA(cudaGetLastError());
      break;
    }
    if ((rc = gemm(m, c_hi, c_lo, ly.attn_out, Ti, EPI_BIAS_RESID_F32, ws.x, ws.y, nullptr, nullptr, st))) return rc;
    ln_kernel<<<row_blocks, 256, 0, st>>>(ws.y, Ti, H, ly.ln1_g, ly.ln1_b, m->cfg.ln_eps, ws.x, ws.x_hi, ws.x_lo);
    CAPR_CHECK_CUDA(cudaGetLastError());
    if ((rc = gemm(m, x_hi, x_lo, ly.ffn1, Ti, EPI_BIAS_GELU_SPLIT, nullptr, nullptr, ws.ffn_hi, ws.ffn_lo, st))) return rc;
    if ((rc = gemm(m, f_hi, f_lo, ly.ffn2, Ti, EPI_BIAS_RESID_F32, ws.x, ws.y, nullptr, nullptr, st))) return rc;
    ln_kernel<<<row_blocks, 256, 0, st>>>(ws.y, Ti, H, ly.ln2_g, ly.ln2_b, m->cfg.ln_eps, ws.x, ws.x_hi, ws.x_lo);
    CAPR_CHECK_CUDA(cudaGetLastError());
    if ((rc = emit_hidden(l + 1))) return rc;
  }
  if (logits) {
    // (CLS-only tail: the final hidden [CLS] rows are the compact buffer, one row per sequence)
    pooler_classifier_kernel<<<n_seq, 256, 2 * H * sizeof(float), st>>>(cls_only ? ws.cx : ws.x, cls_only ? 1 : L, H, m->pool_w, m->pool_b, m->cls_w, m->cls_b,
                                                                        m->cfg.n_labels, logits);
    CAPR_CHECK_CUDA(cudaGetLastError());
  }
  return CAPR_OK;
}

int capr_bert_forward(capr_bert_t h, const int64_t* ids, const int64_t* mask, const int64_t* seg, int n_seq, int L, float* logits,
                      void* workspace, size_t workspace_bytes, capr_stream_t stream) {
  CAPR_REQUIRE(logits || n_seq == 0, CAPR_ERR_BAD_POINTER, "capr_bert_forward: null pointer");
  return bert_run("capr_bert_forward", h, ids, mask, seg, n_seq, L, logits, nullptr, 0, nullptr, workspace, workspace_bytes, stream);
}

int capr_bert_forward_hidden(capr_bert_t h, const int64_t* ids, const int64_t* mask, const int64_t* seg, int n_seq, int L, const int* hidden_layers,
                             int n_hidden, float* hidden_out, float* logits, void* workspace, size_t workspace_bytes, capr_stream_t stream) {
  return bert_run("capr_bert_forward_hidden", h, ids, mask, seg, n_seq, L, logits, hidden_layers, n_hidden, hidden_out, workspace, workspace_bytes, stream);
}

// ---- PARADE aggregation head: ptparade.py:55-78 ---------------------------------------------------
}  // extern "C"

namespace capr {
namespace bert {

// merged[b][0] = initial_cls + pos[0];  merged[b][1+p] = last_hidden[(b*P+p)*L + 0] + pos[1+p]      (ptparade.py:57-62)
__global__ void __launch_bounds__(256) parade_merge_kernel(const float* __restrict__ last_hidden, int B, int P, int L, int H,
                                                           const float* __restrict__ initial_cls, const float* __restrict__ pos,
                                                           float* __restrict__ merged, long long* __restrict__ ones) {
  const int row = blockIdx.x;  // b*(P+1) + j
  const int b = row / (P + 1), j = row - b * (P + 1);
  const float* src = j == 0 ? initial_cls : last_hidden + ((size_t)b * P + (j - 1)) * L * H;
  for (int h = threadIdx.x; h < H; h += blockDim.x) merged[(size_t)row * H + h] = src[h] + pos[(size_t)j * H + h];
  if (threadIdx.x == 0) ones[row] = 1;
}

// score[b] = w . x[b*(P+1) + 0] + bias      (ptparade.py:66-67,78)
__global__ void __launch_bounds__(128) parade_score_kernel(const float* __restrict__ x, int B, int rows_per_doc, int H, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ scores) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = x + (size_t)b * rows_per_doc * H;
  float p = 0.f;
  for (int h = lane; h < H; h += 32) p = fmaf(w[h], row[h], p);
  p = warp_sum(p);
  if (lane == 0) scores[b] = p + bias[0];
}

static size_t parade_extra_bytes(const capr_bert_config& c, size_t rows) {
  return align256(rows * c.hidden * 4) * 2 + align256(rows * 8);  // merged input, aggregator output, all-ones mask
}

}  // namespace bert
}  // namespace capr

extern "C" {

size_t capr_parade_workspace_bytes(capr_bert_t agg, int B, int P) {
  Model* m = (Model*)agg;
  if (!m || B <= 0 || P <= 0) return 0;
  const size_t rows = (size_t)B * (P + 1);
  return carve(m->cfg, rows, (size_t)B, nullptr, nullptr) + parade_extra_bytes(m->cfg, rows);
}

int capr_parade_head(capr_bert_t agg, const float* last_hidden, int B, int P, int L, const float* initial_cls, const float* pos_emb,
                     const float* lin_w, const float* lin_b, float* scores, float* aggregated, void* workspace, size_t workspace_bytes,
                     capr_stream_t stream) {
  const char* fn = "capr_parade_head";
  Model* m = (Model*)agg;
  CAPR_REQUIRE(m, CAPR_ERR_BAD_POINTER, "%s: null handle", fn);
  capr::DeviceGuard device_guard(m->device);
  CAPR_REQUIRE(B >= 0 && P > 0 && L > 0, CAPR_ERR_BAD_SHAPE, "%s: bad shape B=%d P=%d L=%d", fn, B, P, L);
  if (B == 0) return CAPR_OK;
  CAPR_REQUIRE(last_hidden && initial_cls && pos_emb && lin_w && lin_b && scores && workspace, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(((uintptr_t)workspace & 255) == 0, CAPR_ERR_BAD_POINTER, "%s: workspace must be 256-byte aligned", fn);
  const int H = m->cfg.hidden;
  const size_t rows = (size_t)B * (P + 1);
  const size_t enc = carve(m->cfg, rows, (size_t)B, nullptr, nullptr);
  CAPR_REQUIRE(workspace_bytes >= enc + parade_extra_bytes(m->cfg, rows), CAPR_ERR_BAD_SHAPE, "%s: workspace too small (capr_parade_workspace_bytes)", fn);
  unsigned char* p = (unsigned char*)workspace + enc;
  float* merged = (float*)p;
  float* out = (float*)(p + align256(rows * H * 4));
  long long* ones = (long long*)(p + 2 * align256(rows * H * 4));
  cudaStream_t st = (cudaStream_t)stream;
  parade_merge_kernel<<<(unsigned)rows, 256, 0, st>>>(last_hidden, B, P, L, H, initial_cls, pos_emb, merged, ones);
  CAPR_CHECK_CUDA(cudaGetLastError());
  const int last_layer[1] = {m->cfg.layers};
  // the two BertLayers see no attention mask (ptparade.py:64-65): all-ones mask, sequences of P+1 vectors
  const int rc = bert_run(fn, agg, nullptr, (const int64_t*)ones, nullptr, B, P + 1, nullptr, last_layer, 1, out, workspace, enc, stream, merged);
  if (rc) return rc;
  parade_score_kernel<<<(B + 3) / 4, 128, 0, st>>>(out, B, P + 1, H, lin_w, lin_b, scores);
  CAPR_CHECK_CUDA(cudaGetLastError());
  if (aggregated)  // transformer_out_2[:, 0, :] (tests)
    CAPR_CHECK_CUDA(cudaMemcpy2DAsync(aggregated, (size_t)H * 4, out, (size_t)(P + 1) * H * 4, (size_t)H * 4, B, cudaMemcpyDeviceToDevice, st));
  return CAPR_OK;
}

#ifdef CAPR_DEBUG_BUILD
// Debug / test entry: C = A . W^T + bias through the same tcgen05 kernel (A [M,K], W [N,K], fp32 device buffers).
int capr_gemm_test(const float* a, const float* w, const float* bias, int M, int N, int K, int precision_mode, float* c, capr_stream_t stream) {
  const char* fn = "capr_gemm_test";
  CAPR_REQUIRE(a && w && bias && c, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(M > 0 && N > 0 && K > 0 && K % BK == 0 && pick_bn(N) > 0, CAPR_ERR_BAD_SHAPE, "%s: need K %% 64 == 0 and N %% 32 == 0", fn);
  cudaStream_t st = (cudaStream_t)stream;
  capr::DeviceGuard device_guard(a);
  Model m;
  m.mode = precision_mode == CAPR_BERT_BF16X3 ? 3 : 1;
  m.sms = sm_count();
  CAPR_REQUIRE(m.sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  Linear lin;
  const float* ws[1] = {w};
  const float* bs[1] = {bias};
  const int ns[1] = {N};
  int rc = make_linear(&m, &lin, K, ws, bs, ns, 1, st);
  const size_t Mp = ((size_t)M + BM - 1) / BM * BM;
  __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr;
  if (!rc) rc = dev_alloc(&m, (void**)&a_hi, Mp * K * 2);
  if (!rc) rc = dev_alloc(&m, (void**)&a_lo, Mp * K * 2);
  CUtensorMap ma_hi, ma_lo;
  if (!rc) {
    cudaMemsetAsync(a_hi, 0, Mp * K * 2, st);
    cudaMemsetAsync(a_lo, 0, Mp * K * 2, st);
    split_kernel<<<1024, 256, 0, st>>>(a, (size_t)M * K, a_hi, a_lo);
    rc = make_map(&ma_hi, a_hi, Mp, K, BM);
  }
  if (!rc) rc = make_map(&ma_lo, a_lo, Mp, K, BM);
  if (!rc) rc = gemm(&m, ma_hi, ma_lo, lin, M, EPI_BIAS_F32, nullptr, c, nullptr, nullptr, st);
  cudaError_t e = cudaStreamSynchronize(st);
  for (void* p : m.owned) cudaFree(p);
  if (!rc && e != cudaSuccess) return cuda_fail(e, "capr_gemm_test");
  return rc;
}
#endif  // CAPR_DEBUG_BUILD

}  // extern "C"
