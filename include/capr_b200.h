/* capr_b200.h -- C ABI of the B200-native reranker scoring engine (libcapr_b200.so).
 *
 * This is the drop-in boundary of the repo: plain pointers and sizes, no torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the capreolus repo, commit
 * 789288c).  The reference has no FFI of its own -- it is pure Python calling torch ops -- so the
 * "binding a maintainer would add" is a ctypes stub; see INTEGRATION.md and capreolus_b200/_lib.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the parameter is documented as "host";
 *   - token ids are int64, row-major [B,Q] / [B,D]:  id > 0 in-vocabulary, 0 = <pad>, < 0 = OOV
 *     (capreolus/reranker/common.py:174, capreolus/extractor/embedtext.py:142-151);
 *   - work is enqueued on `stream` (a cudaStream_t) and NOT synchronised; buffers are borrowed for
 *     the duration of the enqueued work only;
 *   - every function returns CAPR_OK or a negative capr_status; capr_last_error() gives the text for
 *     the calling thread.  No function falls back to a CPU path.
 *
 * Multi-GPU.  Every entry point works on the device that owns its buffers (a device guard switches to it
 * for the call) and knows nothing about other ranks: pairs are independent, so a caller shards the
 * candidate list, calls the same entry point once per rank on its slice and concatenates the scores.  That
 * one collective -- an all-gather of ceil(N/R) floats per rank, <= 500 KB at BASELINE.json configs[3] -- is
 * DELIBERATELY host-side (capreolus_b200/sharding.py: torch.distributed.all_gather_into_tensor over NCCL)
 * and not part of this ABI: it is latency-bound, follows the last kernel of the step and has nothing to
 * overlap with, and keeping it out keeps the library free of an NCCL / communicator dependency (the
 * reference has no multi-GPU scoring path to mirror; DESIGN.md section 8, measured cost 0.05 ms per step).
 */
#ifndef CAPR_B200_H_
#define CAPR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* capr_stream_t; /* cudaStream_t */

typedef enum {
  CAPR_OK = 0,
  CAPR_ERR_BAD_SHAPE = -1,   /* a dimension is <= 0 or inconsistent */
  CAPR_ERR_BAD_POINTER = -2, /* null or misaligned pointer */
  CAPR_ERR_UNSUPPORTED = -3, /* valid in the reference, not implemented by the kernels (message says what) */
  CAPR_ERR_CUDA = -4,        /* a CUDA runtime call failed; message carries cudaGetErrorString */
  CAPR_ERR_NO_DEVICE = -5    /* no sm_100 device visible */
} capr_status;

#define CAPR_ABI_VERSION 1
int capr_abi_version(void);
const char* capr_last_error(void);

/* Number of SMs / compute capability major*10+minor of the current device (0 if none). */
int capr_device_sm_count(void);
int capr_device_arch(void);

/* ---- embedding table --------------------------------------------------------------------------
 * Replaces create_emb_layer + the two norm()/divide steps of SimilarityMatrix.cosine_similarity_matrix
 * (capreolus/reranker/common.py:279-288, 161-165).  The kernels gather from a *prepared* table:
 * row v = emb[v] / (||emb[v]||_2 + 1e-9), row pitch `capr_table_pitch(E)` floats (E rounded up to a
 * multiple of 16, zero filled), so cos(q,d) is a plain dot product of two gathered rows.  Must be
 * re-run whenever the embedding weights change (finetune=True). */
int capr_table_pitch(int E);
int capr_table_prepare(const float* emb /*[V,E]*/, int V, int E, float* table /*[V,pitch]*/, int pitch,
                       capr_stream_t stream);

/* ---- similarity matrix (debug / tests) ----------------------------------------------------------
 * SimilarityMatrix.forward (capreolus/reranker/common.py:170-182): sim[B,Q,D] fp32. */
int capr_simmat_forward(const int64_t* query, const int64_t* doc, int B, int Q, int D, const float* table, int V,
                        int pitch, float* sim /*[B,Q,D]*/, capr_stream_t stream);

/* ---- KNRM -----------------------------------------------------------------------------------------
 * KNRM_class.forward (capreolus/reranker/KNRM.py:39-55) with RbfKernelBank (common.py:224-250) fused:
 * gather -> cosine tile -> K Gaussian kernels -> sum over doc -> masked log -> sum over query -> combine.
 *   mu, sigma  [K]                     kernels.kernels.{k}.mu / .sigma
 *   w1 [H,K], b1 [H]                   combine.0 ; H = 1 when hidden == 0 (singlefc=True)
 *   w2 [1,hidden], b2 [1]              combine.2 (only when hidden > 0, i.e. singlefc=False)
 *   flags                              CAPR_KNRM_SCORETANH = final tanh (scoretanh=True)
 *   scores [B]       (nullable)        what KNRM.test / KNRM.score return per pair (KNRM.py:87-101)
 *   feats  [B,K]     (nullable)        the log soft-TF features fed to `combine` (KNRM.py:53)
 *   stats  [B,2,K]   (nullable)        backward statistics for d/dmu, d/dsigma
 */
#define CAPR_KNRM_SCORETANH 1
#ifdef CAPR_DEBUG_BUILD
/* profiling aids of the DEBUG build (libcapr_b200_dbg.so) for capr_knrm_forward_tc only; results are NOT valid when set: skip
 * the pooling loop / the TMEM drain / the MMAs / the gathers.  The product library ignores these bits. */
#define CAPR_DEBUG_SKIP_POOL 0x100
#define CAPR_DEBUG_SKIP_DRAIN 0x200
#define CAPR_DEBUG_SKIP_MMA 0x400
#define CAPR_DEBUG_SKIP_GATHER 0x800
#endif
int capr_knrm_forward(const int64_t* query, const int64_t* doc, int B, int Q, int D, const float* table, int V,
                      int pitch, const float* mu, const float* sigma, int K, const float* w1, const float* b1,
                      int hidden, const float* w2, const float* b2, int flags, float* scores, float* feats,
                      float* stats, capr_stream_t stream);

/* Engine 2 (tensor cores): same contract as capr_knrm_forward (inference outputs only), but the cosine tile is
 * computed by tcgen05.mma from a table stored as two bf16 planes, hi = bf16(e) and lo = bf16(e - hi) of the same
 * L2-normalised rows (capr_table_prepare_bf16; pitch = capr_table_pitch_bf16(E), E rounded up to a multiple of 16), with
 * the three products hi.hi + hi.lo + lo.hi accumulated in fp32.  Limits: D <= 1024, pitch <= 320, K <= 16 (else
 * CAPR_ERR_UNSUPPORTED -> use capr_knrm_forward). */
int capr_table_pitch_bf16(int E);
int capr_table_prepare_bf16(const float* emb /*[V,E]*/, int V, int E, void* table_hi /*bf16 [V,pitch]*/,
                            void* table_lo /*bf16 [V,pitch]*/, int pitch, capr_stream_t stream);
int capr_knrm_forward_tc(const int64_t* query, const int64_t* doc, int B, int Q, int D, const void* table_hi,
                         const void* table_lo, int V, int E, int pitch, const float* mu, const float* sigma, int K,
                         const float* w1, const float* b1, int hidden, const float* w2, const float* b2, int flags,
                         float* scores, float* feats, capr_stream_t stream);

/* Engine 3 (tensor cores, term-frequency documents, pooling from tensor memory; csrc/simtc3.cuh): same contract and outputs as
 * capr_knrm_forward_tc.  A pre-pass rewrites every document as its (distinct token, count) list in first-occurrence order --
 * KNRM's kernel sums run over all doc positions (KNRM.py:50), so identical tokens contribute identical terms -- and the scoring
 * kernel gathers each distinct token once.  workspace: capr_tf_workspace_bytes(B, D) bytes hold all B documents at once; a
 * smaller workspace (at least capr_tf_workspace_bytes(1, D)) makes the call loop over chunks of pairs.  256-byte aligned.
 * Limits: Q <= 32, D <= 1024, pitch <= 320, K <= 16 (else CAPR_ERR_UNSUPPORTED -> capr_knrm_forward_tc / capr_knrm_forward). */
size_t capr_tf_workspace_bytes(int B, int D);
/* The pre-pass on its own: doc [B,D] int64 -> ids [B,D] int32 (the distinct tokens of each document in first-occurrence order,
 * <pad> = 0 and OOV (< 0) ids included as tokens, 0 beyond n_distinct[b]), counts [B,D] uint16 (multiplicities, 0 beyond),
 * n_distinct [B].  sum(counts[b]) == D; ids outside int32 are clamped like everywhere else. */
int capr_tf_dedup(const int64_t* doc, int B, int D, int32_t* ids, uint16_t* counts, int32_t* n_distinct, capr_stream_t stream);
int capr_knrm_forward_tf(const int64_t* query, const int64_t* doc, int B, int Q, int D, const void* table_hi,
                         const void* table_lo, int V, int E, int pitch, const float* mu, const float* sigma, int K,
                         const float* w1, const float* b1, int hidden, const float* w2, const float* b2, int flags,
                         float* scores, float* feats, void* workspace, size_t workspace_bytes, capr_stream_t stream);

/* ---- DRMM -----------------------------------------------------------------------------------------
 * DRMM_class.forward (capreolus/reranker/DRMM.py:101-116): _hist_map (41-81) + ffw + _term_gate (83-99).
 *   idf [B,Q] fp32; bin_ub [nbins] = torch.linspace(-1,1,nbins+1)[1:] (device, the exact fp32 values)
 *   hist_type: 0 = CH, 1 = NH, 2 = LCH;  gate_type: 0 = IDF (gate_w [1]), 1 = TV (gate_w [E], raw_emb [V,E])
 *   ffw_w1 [nodes,nbins+1], ffw_b1 [nodes], ffw_w2 [nodes], ffw_b2 [1], out_w [1], out_b [1]
 *   hist_out [B,Q,nbins+1] (nullable) is the transformed histogram _hist_map returns. */
#define CAPR_DRMM_CH 0
#define CAPR_DRMM_NH 1
#define CAPR_DRMM_LCH 2
#define CAPR_DRMM_GATE_IDF 0
#define CAPR_DRMM_GATE_TV 1
int capr_drmm_forward(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                      const float* table, int V, int pitch, const float* raw_emb, int E, int nbins,
                      const float* bin_ub, int hist_type, int gate_type, const float* ffw_w1, const float* ffw_b1,
                      int nodes, const float* ffw_w2, const float* ffw_b2, const float* gate_w, const float* out_w,
                      const float* out_b, float* scores, float* hist_out, capr_stream_t stream);

/* Engine 2 (tensor cores), see capr_knrm_forward_tc.  Limits: D <= 1024, pitch <= 320, nbins <= 31. */
int capr_drmm_forward_tc(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                         const void* table_hi, const void* table_lo, int V, int pitch, const float* raw_emb, int E,
                         int nbins, const float* bin_ub, int hist_type, int gate_type, const float* ffw_w1,
                         const float* ffw_b1, int nodes, const float* ffw_w2, const float* ffw_b2, const float* gate_w,
                         const float* out_w, const float* out_b, float* scores, float* hist_out, capr_stream_t stream);

/* ---- PACRR ----------------------------------------------------------------------------------------
 * PACRR_class.forward (capreolus/reranker/PACRR.py:43-54) with PACRRConvMax2dModule (57-82) fused.
 *   conv_w / conv_b: HOST arrays of (maxgram-mingram+1) device pointers, ngrams.{i}.conv.weight [F,1,n,n] / .bias [F]
 *   idf [B,Q] or NULL (config idf=False);  nonlin: 0 none, 1 relu, 2 tanh
 *   l1w [C, Q*(ngrams*kmax + (idf?1:0))], l1b [C], l2w [C,C], l2b [C], l3w [1,C], l3b [1]
 *   topk_out [B,Q,ngrams*kmax] (nullable). */
int capr_pacrr_forward(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                       const float* table, int V, int pitch, int mingram, int maxgram, int nfilters, int kmax,
                       const float* const* conv_w, const float* const* conv_b, const float* l1w, const float* l1b,
                       const float* l2w, const float* l2b, const float* l3w, const float* l3b, int combine,
                       int nonlin, float* scores, float* topk_out, capr_stream_t stream);

/* Engine 2 (tensor cores), see capr_knrm_forward_tc.  Limits: D <= 512, pitch <= 320. */
int capr_pacrr_forward_tc(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D,
                          const void* table_hi, const void* table_lo, int V, int E, int pitch, int mingram, int maxgram,
                          int nfilters, int kmax, const float* const* conv_w, const float* const* conv_b, const float* l1w,
                          const float* l1b, const float* l2w, const float* l2b, const float* l3w, const float* l3b,
                          int combine, int nonlin, float* scores, float* topk_out, capr_stream_t stream);

/* ---- DRMMTKS ------------------------------------------------------------------
 * DRMMTKS_class.forward (capreolus/reranker/DRMMTKS.py:50-63): cosine matrix -> torch.topk(k) over the doc axis per query
 * term -> ffw = Linear(k,1)+tanh -> IDF softmax term gate (DRMMTKS.py:32-48) -> output_layer.  Tensor-core engine only
 * (table as bf16 hi/lo planes, capr_table_prepare_bf16).  gateType='TV' is not offered: the reference feeds int64 token
 * ids to a Linear(E,1) there and raises.
 *   ffw_w [1,topk], ffw_b [1], gate_w [1], out_w [1], out_b [1];  idf [B,Q];  topk_out [B,Q,topk] (nullable, descending).
 * Limits: Q <= 32, D <= 1024, pitch <= 320, topk <= min(32, D). */
int capr_drmmtks_forward_tc(const int64_t* query, const int64_t* doc, const float* idf, int B, int Q, int D, const void* table_hi,
                            const void* table_lo, int V, int E, int pitch, int topk, const float* ffw_w, const float* ffw_b,
                            const float* gate_w, const float* out_w, const float* out_b, float* scores, float* topk_out,
                            capr_stream_t stream);

/* ---- ConvKNRM -----------------------------------------------------------------
 * ConvKNRM_class.forward (capreolus/reranker/ConvKNRM.py:43-77) with StackedSimilarityMatrix (common.py:187-221) and
 * RbfKernelBank.  The embedding is frozen (ConvKNRM.py:18) and Conv1d is linear, so the n-gram encoders are folded into
 * PROJECTED TABLES once per weight version:
 *   capr_convknrm_project: proj [V, S*F] fp32, S = maxngram(maxngram+1)/2, column block slot(n,u) = n(n-1)/2 + u holds
 *   emb . convs.{n-1}.0.weight[:, :, u]^T   (conv_w: HOST array of maxngram DEVICE pointers, weight [F,E,n]).
 * capr_convknrm_forward then needs per token only row gathers + adds:  conv_n(t) = bias_n + sum_u proj[tok[t+u]][slot(n,u)].
 *   conv_b: HOST array of maxngram DEVICE pointers (convs.{n-1}.0.bias [F]);  crossmatch: 1 = all maxngram^2 (nq, nd) views,
 *   0 = nq == nd only;  mu, sigma [K];  w1 [H, K*VIEWS], b1 [H] (H = 1 when hidden == 0), w2 [1,hidden], b2 [1];
 *   flags: CAPR_KNRM_SCORETANH;  feats_out [B, K*VIEWS] (nullable) = the tensor fed to `combine`, index k*VIEWS + view.
 *   workspace: capr_convknrm_workspace_bytes(chunk, ...) bytes for `chunk` pairs at a time (any chunk >= 1 works; the
 *   call loops over as many pairs as fit), 256-byte aligned.
 * Token ids: 0 = <pad> (extractor.pad, ConvKNRM.py:17); ids < 0 or >= V make the reference raise IndexError -- here they
 * contribute a zero vector.  Limits: Q <= 32, D <= 512, F % 4 == 0, F <= 320, maxngram <= 4, K <= 16. */
int capr_convknrm_proj_cols(int maxngram, int F);
int capr_convknrm_project(const float* emb /*[V,E]*/, int V, int E, const float* const* conv_w, int maxngram, int F,
                          float* proj /*[V, capr_convknrm_proj_cols]*/, capr_stream_t stream);
size_t capr_convknrm_workspace_bytes(int chunk, int Q, int D, int maxngram, int F, int K, int crossmatch);
int capr_convknrm_forward(const int64_t* query, const int64_t* doc, int B, int Q, int D, const float* proj, int V, int maxngram,
                          int F, const float* const* conv_b, int crossmatch, const float* mu, const float* sigma, int K,
                          const float* w1, const float* b1, int hidden, const float* w2, const float* b2, int flags,
                          float* scores, float* feats_out, void* workspace, size_t workspace_bytes, capr_stream_t stream);

/* ---- predict-loop plumbing ------------------------------------------------
 * capr_assemble_pairs replaces EmbedText.id2vec / padlist per pair (capreolus/extractor/embedtext.py:128-162,
 * capreolus/utils/common.py:99-111) as driven by PredSampler.generate_samples (capreolus/sampler/__init__.py:222-233):
 * the tokenised queries and documents live once on the device as packed id stores (q_store / d_store: flat int32 token ids,
 * q_off [n_queries+1] / d_off [n_docs+1]: int64 offsets; idf_store: optional fp32 parallel to q_store) and a batch is two
 * int32 index vectors.  Output rows are what the rerankers consume: query_out [N,Q], doc_out [N,D] int64, truncated to
 * Q / D tokens and right-padded with 0; idf_out [N,Q] (nullable).  An index outside the store gives an all-pad row. */
int capr_assemble_pairs(const int32_t* q_store, const int64_t* q_off, int n_queries, const int32_t* d_store, const int64_t* d_off,
                        int n_docs, const float* idf_store, const int32_t* qidx, const int32_t* didx, int N, int Q, int D,
                        int64_t* query_out, int64_t* doc_out, float* idf_out, capr_stream_t stream);
/* capr_assemble_bert_pairs: the same for the BERT rerankers -- BertPassage._get_sliding_window_passages + _prepare_bert_input
 * (capreolus/extractor/bertpassage.py:203-232, 268-284) on WordPiece ids: per pair P passage rows
 * [CLS] query [SEP] doc[p*stride : p*stride+passagelen] [SEP] [PAD]..., query truncated to maxqlen (padded to it when padq),
 * passage truncated to L - len(query) - 3, exhausted documents give the one-token pad passage; mask = 1 on written tokens
 * != pad_id; segment = 0 for [CLS] query [SEP] and 1 to the end including the padding.  ids / mask / seg: [N, P, L] int64. */
int capr_assemble_bert_pairs(const int32_t* q_store, const int64_t* q_off, int n_queries, const int32_t* d_store, const int64_t* d_off,
                             int n_docs, const int32_t* qidx, const int32_t* didx, int N, int P, int L, int maxqlen, int padq,
                             int passagelen, int stride, int cls_id, int sep_id, int pad_id, int64_t* ids, int64_t* mask, int64_t* seg,
                             capr_stream_t stream);
/* capr_widen_ids: the host side of the predict loop may ship token ids narrower than the reference's int64
 * (capreolus/extractor/embedtext.py:146-147 emits np.long; a 30 k vocabulary and its negative OOV ids fit int16): src is
 * n int16 (src_bytes 2) or int32 (src_bytes 4) ids, dst the sign-extended int64 ids every scoring entry point takes.  Cuts the
 * host -> device bytes of capreolus/trainer/pytorch.py:342 (`v.to(device)`) 4x / 2x. */
int capr_widen_ids(const void* src, int src_bytes, size_t n, int64_t* dst, capr_stream_t stream);
/* capr_rank_by_query replaces the float16 rounding of PytorchTrainer.predict (capreolus/trainer/pytorch.py:345-348) and the
 * per-query sort of Searcher.write_trec_run (capreolus/searcher/__init__.py:48-58).  Pairs of one query are contiguous:
 * seg_off [n_queries+1] int64.  rounded [N] (nullable) = float(float16(score)); order [N]: for query q, order[seg_off[q]+r]
 * is the position (inside the segment) of the rank-(r+1) document: score descending, ties in input order (Python's stable
 * sort).  max_segment = the largest segment length (<= 4096). */
int capr_rank_by_query(const float* scores, const int64_t* seg_off, int n_queries, int max_segment, float* rounded, int32_t* order,
                       capr_stream_t stream);

/* ---- pairwise losses (tests / training loop) ------------------------------------------------------
 * pair_hinge_loss (capreolus/reranker/common.py:7,101-103): loss[0] = mean(max(0, 1 - (pos - neg))),
 * grad_pos/grad_neg [B] (nullable) = d loss / d score. */
int capr_pair_hinge(const float* pos, const float* neg, int B, float* loss, float* grad_pos, float* grad_neg,
                    capr_stream_t stream);
/* pair_softmax_loss (capreolus/reranker/common.py:96-98; trainer config softmaxloss=True, trainer/pytorch.py:220-223):
 * loss[0] = mean(1 - softmax([pos, neg], dim=1)[:, 0]); grad_pos / grad_neg [B] (nullable) = d loss / d score. */
int capr_pair_softmax(const float* pos, const float* neg, int B, float* loss, float* grad_pos, float* grad_neg,
                      capr_stream_t stream);

/* ---- monoBERT / BERT-MaxP encoder -------------------------------------------------------------------
 * PTBERTMaxP_Class.predict_step (capreolus/reranker/ptBERTMaxP.py:67-96) calls
 * self.bert(ids, attention_mask, token_type_ids)[0], a HF BertForSequenceClassification; this handle is that
 * encoder on tcgen05 tensor cores.  capr_bert_create SNAPSHOTS the weights (it converts the Linear weights to
 * bf16 (hi, lo) planes); call it again after the torch parameters change.
 *
 * weights: HOST array of capr_bert_num_weights(cfg) = 5 + 16*layers + 4 DEVICE fp32 pointers, in this order
 * (HF state_dict names):
 *   bert.embeddings.word_embeddings.weight [vocab,H], position_embeddings.weight [max_pos,H],
 *   token_type_embeddings.weight [type_vocab,H], embeddings.LayerNorm.weight [H], .bias [H];
 *   per layer i: attention.self.query.weight [H,H], .bias, key.weight, .bias, value.weight, .bias,
 *     attention.output.dense.weight [H,H], .bias, attention.output.LayerNorm.weight, .bias,
 *     intermediate.dense.weight [I,H], .bias, output.dense.weight [H,I], .bias, output.LayerNorm.weight, .bias;
 *   bert.pooler.dense.weight [H,H], .bias, classifier.weight [n_labels,H], classifier.bias.
 * precision_mode: CAPR_BERT_BF16X3 = every fp32 operand as two bf16 planes, 3 tensor-core products per K step
 * (meets the 1e-3 parity bar); CAPR_BERT_BF16 = plain bf16 operands (3x fewer MMAs, ~2e-2 relative error). */
typedef struct {
  int hidden, layers, heads, intermediate, vocab, max_pos, type_vocab, n_labels;
  float ln_eps;
} capr_bert_config;
typedef struct capr_bert_opaque* capr_bert_t;
#define CAPR_BERT_BF16 1
#define CAPR_BERT_BF16X3 3
int capr_bert_num_weights(const capr_bert_config* cfg);
int capr_bert_create(const capr_bert_config* cfg, const float* const* weights, int n_weights, int precision_mode,
                     capr_stream_t stream, capr_bert_t* out);
void capr_bert_destroy(capr_bert_t handle);
/* Bytes of scratch capr_bert_forward needs for n_seq sequences of length L (256-byte aligned device buffer). */
size_t capr_bert_workspace_bytes(capr_bert_t handle, int n_seq, int L);
/* ids / mask / seg: [n_seq, L] int64 (mask 1 = real token; capreolus/extractor/bertpassage.py:268-284);
 * logits [n_seq, n_labels] fp32 = the classifier output; the passage score is logits[:, 1]. */
int capr_bert_forward(capr_bert_t handle, const int64_t* ids, const int64_t* mask, const int64_t* seg, int n_seq, int L,
                      float* logits, void* workspace, size_t workspace_bytes, capr_stream_t stream);
/* Same encoder, additionally copying fp32 hidden states out (HF `output_hidden_states=True`, as CEDRKNRM.py:19-38 asks for):
 * hidden_layers [n_hidden] (HOST ints in [0, layers]: 0 = embedding output, l = output of encoder layer l) ->
 * hidden_out [n_hidden, n_seq*L, H] fp32.  logits may be NULL (encoder without a classification head). */
int capr_bert_forward_hidden(capr_bert_t handle, const int64_t* ids, const int64_t* mask, const int64_t* seg, int n_seq, int L,
                             const int* hidden_layers, int n_hidden, float* hidden_out, float* logits, void* workspace,
                             size_t workspace_bytes, capr_stream_t stream);

/* ---- PARADE aggregation head -----------------------------------------------------------
 * PTParade_Class.aggregate_using_transformer + linear (capreolus/reranker/ptparade.py:55-78): the [CLS] vector of every passage
 * (row 0 of each passage in last_hidden [B*P*L, H], from capr_bert_forward_hidden of the passage encoder) is prefixed with
 * `initial_cls_embedding` [H], `full_position_embeddings` [(P+1), H] are added, the B sequences of P+1 vectors run through
 * transformer_layer_1/2 WITHOUT an attention mask and score[b] = linear(out[b, 0, :]).
 * `agg` is a capr_bert_t created from the two BertLayers' weights (layers = 2; its embedding / pooler / classifier slots
 * are never used: pass any finite tensors of the right shapes).  aggregated [B, H] (nullable) = transformer_out_2[:, 0, :]. */
size_t capr_parade_workspace_bytes(capr_bert_t agg, int B, int P);
int capr_parade_head(capr_bert_t agg, const float* last_hidden, int B, int P, int L, const float* initial_cls, const float* pos_emb,
                     const float* lin_w, const float* lin_b, float* scores, float* aggregated, void* workspace, size_t workspace_bytes,
                     capr_stream_t stream);

/* ---- CEDR-KNRM head ------------------------------------------------------------------
 * CEDRKNRM_Class.masked_simmats / knrm / forward (capreolus/reranker/CEDRKNRM.py:85-171) on the hidden states
 * capr_bert_forward_hidden produced for n_seq = B*P passages ([CLS] q [SEP] doc [SEP] pad; P passages per document):
 *   hidden [n_layers, B*P*L, H]   the layers listed in `simmat_layers` (n_layers may be 0: cls feature only)
 *   last_hidden [B*P*L, H]        hidden_states[-1] (its [CLS] rows give the cls feature; NULL when cls_mode == 0)
 *   query rows = positions 1..maxqlen+1 masked by mask*(seg==0); doc columns = positions 1..L-1 masked by mask*(seg==1);
 *   cosine = a.b / ((|a|+1e-9)(|b|+1e-9)); K kernels (mu, sigma [K]) summed over the unmasked doc tokens of all P passages,
 *   log(clamp(., 1e-10)) * 0.01, summed over the maxqlen+1 query rows -> K features per layer.
 *   cls_mode: 0 none, 1 avg, 2 max over the passages.   feats [B, F], F = capr_cedrknrm_feature_dim (cls first).
 *   combine: w1 [combine_hidden or 1, F], b1; w2 [1, combine_hidden], b2 (combine_hidden == 0: single Linear(F,1)).
 *   scores [B] (nullable).  workspace: capr_cedrknrm_workspace_bytes(B*P, maxqlen, n_layers, K, combine_hidden).
 * Limits: maxqlen < 64, K <= 16, H <= 1024 and a multiple of 4. */
int capr_cedrknrm_feature_dim(int H, int n_layers, int K, int cls_mode);
size_t capr_cedrknrm_workspace_bytes(int n_seq, int maxqlen, int n_layers, int K, int combine_hidden);
int capr_cedrknrm_head(const float* hidden, int n_layers, const float* last_hidden, const int64_t* mask, const int64_t* seg, int B,
                       int P, int L, int H, int maxqlen, const float* mu, const float* sigma, int K, int cls_mode, const float* w1,
                       const float* b1, int combine_hidden, const float* w2, const float* b2, float* feats, float* scores,
                       void* workspace, size_t workspace_bytes, capr_stream_t stream);
/* ---- debug build only (libcapr_b200_dbg.so = the same sources with -DCAPR_DEBUG_BUILD + the micro-benchmarks) --------------
 * None of the following is exported by the product library libcapr_b200.so. */
#ifdef CAPR_DEBUG_BUILD
/* Test hook: C[M,N] = A[M,K] . W[N,K]^T + bias through the encoder's tcgen05 GEMM kernel (synchronises the stream). */
int capr_gemm_test(const float* a, const float* w, const float* bias, int M, int N, int K, int precision_mode, float* c,
                   capr_stream_t stream);

/* Debug micro-benchmark (not on any product path): cycles[grid] = best-of-reps SM cycles for n_mma back-to-back
 * tcgen05.mma of shape M x N x 16 (bf16, shared-memory operands) cycling over n_acc accumulators. */
int capr_debug_mma_bench(int M, int N, int n_mma, int n_acc, int reps, int grid, long long* cycles, capr_stream_t stream);
/* Debug micro-benchmark: cycles[grid] = SM cycles for iters x 128 packed fp32 FMAs per thread (256 threads per CTA) with the
 * scalar operand from the constant bank (mode 0, the form PACRR's conv uses), from vector registers (mode 1), or as plain FFMA
 * pairs (mode 2).  scratch: >= 64 + grid*256 floats. */
int capr_debug_ffma2_bench(int mode, int iters, int grid, float* scratch, long long* cycles, capr_stream_t stream);
/* Debug micro-benchmark (csrc/bench/gather_bench.cu): the pure L2 -> SM row-gather rate of the KNRM-family producer's access
 * pattern -- 128-row units of a bf16 (hi, lo) table of `pitch` elements per row, rows[n_rows] int32 in [0, V), `stages` 16 KB
 * stages in flight per SM.  The caller times the launch; bytes moved = (n_rows rounded down to 128) * pitch * 2 planes * 2 B. */
int capr_debug_gather_bench(const void* table_hi, const void* table_lo, int V, int pitch, const int32_t* rows, int n_rows, int stages,
                            capr_stream_t stream);
/* Same measurement with the hand-off variants side by side: mode bit 0 = WARP-OWNED stages (one producer warp fills a whole
 * stage: 32 cp.async.mbarrier.arrive.noinc per stage instead of 128), bit 1 = no copies (hand-off skeleton only).  stages: 4, 8, 12. */
int capr_debug_gather_bench2(const void* table_hi, const void* table_lo, int V, int pitch, const int32_t* rows, int n_rows, int stages, int mode,
                             capr_stream_t stream);
/* ... and with 4, 8 or 16 producer warps sharing every stage (is the gather limited per warp or per SM?). */
int capr_debug_gather_bench3(const void* table_hi, const void* table_lo, int V, int pitch, const int32_t* rows, int n_rows, int stages, int prod_warps,
                             capr_stream_t stream);
#endif /* CAPR_DEBUG_BUILD */

#ifdef __cplusplus
}
#endif
#endif /* CAPR_B200_H_ */
