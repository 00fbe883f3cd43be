/* allrank_b200 -- C ABI of the B200 (sm_100a) scoring + listwise-loss + metric path.
 *
 * This is the drop-in boundary.  allRank is pure Python, so there is no
 * existing FFI to mirror; each entry point below replaces one *Python call surface* of the
 * reference and cites it.  The Python host layer (allrank_b200/{losses,metrics,model}.py) binds
 * these with ctypes and re-exports the reference's names/signatures; INTEGRATION.md shows the
 * stub a maintainer adds on the allRank side.
 *
 * Conventions
 *   - plain pointers + sizes only; no torch types.  All pointers are DEVICE pointers unless the
 *     parameter name ends in `_host`.  Tensors are row-major, contiguous, fp32 unless stated.
 *   - the library never allocates or frees device memory: outputs and workspaces are caller-owned
 *     (the host layer allocates them with PyTorch's caching allocator).
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) of the CURRENT device and returns without synchronising; calls are re-entrant.
 *   - return value: 0 = ok, <0 = ARB_E_* below; arb_last_error() gives a thread-local message.
 *   - inputs are const: the caller's y_pred / y_true / x / mask are never modified
 *     (the reference clones before masking: listNet.py:17-18, lambdaLoss.py:25-26, metrics.py:53-54).
 */
#ifndef ALLRANK_B200_H
#define ALLRANK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARB_OK 0
#define ARB_E_INVALID_ARG (-1)
#define ARB_E_UNSUPPORTED (-2)
#define ARB_E_CUDA (-3)
#define ARB_E_WORKSPACE (-4)

#define ARB_MAX_ATS 32

/* lambdaLoss weighing schemes: allrank/models/losses/lambdaLoss.py:84-114 (selected by name at :61) */
#define ARB_SCHEME_NONE 0
#define ARB_SCHEME_NDCGLOSS1 1
#define ARB_SCHEME_NDCGLOSS2 2
#define ARB_SCHEME_LAMBDARANK 3
#define ARB_SCHEME_NDCGLOSS2PP 4
#define ARB_SCHEME_RANKNET 5
#define ARB_SCHEME_RANKNET_GTDIFF 6
#define ARB_SCHEME_RANKNET_GTDIFF_POWED 7

#define ARB_REDUCTION_SUM 0
#define ARB_REDUCTION_MEAN 1
#define ARB_LOG_BINARY 0
#define ARB_LOG_NATURAL 1

#define ARB_GAIN_POW2 0     /* 2^x - 1 : default gain_function of metrics.dcg (metrics.py:41)          */
#define ARB_GAIN_IDENTITY 1 /* x       : what neuralNDCG passes when powered_relevancies=False (:58)   */

const char* arb_last_error(void);
/* 4 (round 2): packed rows (arb_set_pack_rows / arb_get_pack_rows; the scorer workspace layout depends on the call's
 * dropout rates), arb_set_attention_bwd_persistent; 3: general FC block, bf16 mode */
int32_t arb_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t arb_launch_count(void);

/* ---------------------------------------------------------------- metrics
 * Replaces allrank.models.metrics.{dcg,ndcg,mrr} (allrank/models/metrics.py:41-77, :7-28, :80-113),
 * called from train_utils.metric_on_batch (allrank/training/train_utils.py:32-34).
 * One launch computes any subset of the outputs (NULL pointer = not wanted):
 *   out_dcg   [B,n_ats]  DCG@at of labels ranked by y_pred          (ats_dcg_host: already min(at,S))
 *   out_idcg  [B,n_ats]  DCG@at of labels ranked by themselves
 *   out_ndcg  [B,n_ats]  out_dcg/out_idcg, `filler` where out_idcg == 0
 *   out_mrr   [B,n_ats]  reciprocal rank of the first max-label item if rank < at (ats_mrr_host, unclipped);
 *                        all zeros if the batch-wide sum of per-slate max labels is 0 (metrics.py:108-109)
 *   out_order [B,S] i32  the descending argsort of the masked scores (stable; bit-exact on tie-free input)
 * `discounts` is the [S] fp32 table 1/log2(j+2) that the reference evaluates on the HOST (metrics.py:64);
 * the host layer builds it the same way so DCG values can match bit for bit.
 * `mrr_scratch` : >= 2*B floats. */
int32_t arb_rank_metrics(const float* y_pred, const float* y_true, int32_t B, int32_t S,
                         const float* discounts, const int32_t* ats_dcg_host, const int32_t* ats_mrr_host,
                         int32_t n_ats, int32_t gain_mode, float pad_value, float filler,
                         float* out_dcg, float* out_idcg, float* out_ndcg, float* out_mrr, int32_t* out_order,
                         float* mrr_scratch, void* stream);

/* ---------------------------------------------------------------- losses (forward + backward in one launch)
 * Each replaces `loss_func(y_pred, y_true)` at allrank/training/train_utils.py:20 for one member of
 * allrank.models.losses.  `loss` is a device scalar; `grad` ([B,S], may be NULL for eval) receives
 * d loss / d y_pred.  `scratch` : >= 2*B floats. */

/* listNet(y_pred, y_true, eps, padded_value_indicator)            allrank/models/losses/listNet.py:8-30 */
int32_t arb_listnet(const float* y_pred, const float* y_true, int32_t B, int32_t S, float eps, float pad_value,
                    float* loss, float* grad, float* scratch, void* stream);

/* listMLE(...)                                                    allrank/models/losses/listMLE.py:7-38
 * `perm` [S] i64: the column shuffle the reference draws with torch.randperm (:17) -- drawn by the host
 * layer from the same global CPU RNG.  `order` (nullable [B,S] i32): debug hook feeding a realised sort
 * order of the shuffled labels; NULL = stable descending sort on the device. */
int32_t arb_listmle(const float* y_pred, const float* y_true, int32_t B, int32_t S, float eps, float pad_value,
                    const int64_t* perm, const int32_t* order, float* loss, float* grad, float* scratch,
                    void* stream);

/* approxNDCGLoss(y_pred, y_true, eps, padded_value_indicator, alpha)   .../losses/approxNDCG.py:7-53 */
int32_t arb_approx_ndcg(const float* y_pred, const float* y_true, int32_t B, int32_t S, float eps,
                        float pad_value, float alpha, float* loss, float* grad, float* scratch, void* stream);

/* lambdaLoss(y_pred, y_true, eps, pad, weighing_scheme, k, sigma, mu, reduction, reduction_log)
 *                                                                  .../losses/lambdaLoss.py:7-114
 * k <= 0 means "None" (no truncation). */
int32_t arb_lambda_loss(const float* y_pred, const float* y_true, int32_t B, int32_t S, float eps,
                        float pad_value, int32_t scheme, int32_t k, float sigma, float mu, int32_t reduction,
                        int32_t log_base, float* loss, float* grad, float* scratch, void* stream);

/* --- SURVEY.md 8(f) rank-1 "next" losses, built on the same machinery ---
 * rankNet / rankNet_weightByGTDiff / rankNet_weightByGTDiff_pow            .../losses/rankNet.py:9-79
 * weight_mode 0: unweighted, 1: |t_i - t_j|, 2: |t_i^2 - t_j^2|; mean over all selected pairs of the batch. */
int32_t arb_ranknet(const float* y_pred, const float* y_true, int32_t B, int32_t S, float pad_value,
                    int32_t weight_mode, float* loss, float* grad, float* scratch, void* stream);
/* mode 0: binary_listNet(eps)            .../losses/binary_listNet.py:8-33
 * mode 1: pointwise_rmse(no_of_levels = param)   .../losses/pointwise.py:6-32
 * mode 2: bce (y_pred are probabilities)          .../losses/bce.py:8-32 */
int32_t arb_pointwise_loss(const float* y_pred, const float* y_true, int32_t B, int32_t S, float pad_value,
                           int32_t mode, float param, float eps, float* loss, float* grad, float* scratch,
                           void* stream);

/* ordinal(y_pred, y_true, n, pad)                                          .../losses/ordinal.py:8-50
 * y_pred [B,S,n] are probabilities (the scorer's d_output = n, Sigmoid head); target level j of an item with
 * label t is 1[t >= j+1] (with_ordinals, :8-22).  loss = sum of the n*valid BCE terms / number of valid items. */
int32_t arb_ordinal(const float* y_pred, const float* y_true, int32_t B, int32_t S, int32_t n, float pad_value,
                    float* loss, float* grad, float* scratch, void* stream);

/* neuralNDCG(y_pred, y_true, pad, temperature, powered_relevancies, k, stochastic=False)
 *                            .../losses/neuralNDCG.py:10-70 + loss_utils.py:8-67 (NeuralSort, Sinkhorn)
 * max_iter / tol are the Sinkhorn parameters the reference hard-codes to 50 / 1e-6 (neuralNDCG.py:41-42).
 * `discounts`: the same host-evaluated [S] table 1/log2(j+2) as arb_rank_metrics (neuralNDCG.py:52).
 * `workspace`: arb_neural_ndcg_workspace_bytes(B,S,max_iter) bytes (0 when every slate fits in shared memory). */
size_t arb_neural_ndcg_workspace_bytes(int32_t B, int32_t S, int32_t max_iter);
int32_t arb_neural_ndcg(const float* y_pred, const float* y_true, int32_t B, int32_t S,
                        const float* discounts, float pad_value, float temperature, int32_t powered_relevancies, int32_t k, int32_t max_iter, float tol,
                        float* loss, float* grad, float* scratch, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Parity hook for deterministic_neural_sort / sinkhorn_scaling (allrank/models/losses/loss_utils.py:34-67, :8-31), which
 * arb_neural_ndcg fuses: the kernel's own matrices for slates of at most 128 items.  p0_out, p_out: zero-filled
 * [B,S,S] buffers; entries [b, rank j, item i] of the real items receive NeuralSort's P_hat and its Sinkhorn scaling
 * (max_iter iterations at most, per-slate tolerance test).  Slates without a relevant item are left untouched.
 * scratch: >= 2*B floats. */
int32_t arb_neural_sort_debug(const float* y_pred, const float* y_true, int32_t B, int32_t S, const float* discounts,
                              float pad_value, float temperature, int32_t max_iter, float tol, float* p0_out,
                              float* p_out, float* scratch, void* stream);

/* ---------------------------------------------------------------- scorer: LTRModel(x, mask, indices) -> scores
 * Replaces allrank.models.model.LTRModel.forward / .score (allrank/models/model.py:72-92) for the model family
 * make_model builds (model.py:131-151): one input Linear (FCModel, model.py:12-44) -> N pre-norm Transformer
 * encoder blocks (allrank/models/transformer.py:28-227: custom LayerNorm, MultiHeadedAttention with key-padding
 * mask, PositionwiseFeedForward, residuals) -> final LayerNorm -> Linear(d_model -> 1) head (model.py:95-128).
 * Called from allrank/training/train_utils.py:20 (`model(xb, mask, indices)`) and :34 (`model.score`).
 *
 * Every matrix product runs on the tcgen05 tensor cores in TF32 with fp32 accumulation (csrc/gemm_tf32.cu);
 * LayerNorm / softmax / head are fp32 SIMT kernels (csrc/scorer_kernels.cu).  Dropout masks are a counter-based
 * hash of (seed, layer, site, element index) fused into the producing kernels and regenerated in backward
 * (csrc/dropout.cuh); `seed` must be the same in the forward and the backward call of a step.
 *
 *  Parameters live in ONE flat fp32 buffer (the host layer makes the nn.Parameters views of it), laid out as
 *   per FC layer i: fc_w[s_i, s_{i-1}] fc_b[s_i] (s_{-1} = F) | in_norm_w[F] in_norm_b[F] if fc_input_norm |
 *   (pad to a multiple of 8 elements) | per layer: wq wk wv [3d,d] bq bk bv [3d] wo[d,d] bo[d] w1[dff,d] b1[dff] w2[d,dff] b2[d]
 *   ln1_a ln1_b ln2_a ln2_b [d each] | lnf_a lnf_b [d] head_w[d] head_b[1] (pad to 4) | pe[pe_rows,d] if pe_mode == 2
 * (a fixed sinusoidal table, pe_mode == 1, is a buffer, not a parameter: it is passed separately as `pe_table`)
 * arb_scorer_param_count() gives the total; gradients use the same layout and are ACCUMULATED into `grads`.
 */
#define ARB_ACT_NONE 0
#define ARB_ACT_TANH 1
#define ARB_ACT_SIGMOID 2
#define ARB_ACT_RELU 3

typedef struct arb_scorer_config {
  int32_t n_features;   /* F (row pitch of x; must be a multiple of 4 -- the host layer pads)            */
  int32_t d_model;      /* fc_model.sizes[-1] (model.py:142); multiple of 32 * n_heads                    */
  int32_t n_layers;     /* transformer N; 0 = FC-only model (no encoder, no final LayerNorm)              */
  int32_t n_heads;      /* h                                                                             */
  int32_t d_ff;         /* PositionwiseFeedForward hidden width                                           */
  int32_t out_act;      /* ARB_ACT_*: post_model.output_activation (model.py:105-107)                     */
  float ln_eps;         /* 1e-6 (transformer.py:63)                                                       */
  float dropout;        /* transformer.dropout: on attention probabilities, both sublayer outputs and the FFN
                           hidden layer (transformer.py:105,155,227); applied only when training != 0          */
  float fc_dropout;     /* fc_model.dropout on the input FC output (model.py:43)                          */
  int32_t pe_mode;      /* positional encoding (allrank/models/positional.py): 0 none, 1 fixed table, 2 learned     */
  int32_t pe_rows;      /* rows of the table = max_indices + 1; the last row is the padding row              */
  int32_t d_output;     /* post_model.d_output (model.py:104,108): outputs per item; 0 or 1 = one score per item.
                           > 1 (ordinal loss): scores are [B,S,d_output], head weight [d_output,d_model]           */
  /* --- ABI v3: the general FCModel input block (model.py:16-44): [nn.LayerNorm(F)] -> n x dropout(act(Linear)) */
  int32_t n_fc_layers;  /* len(fc_model.sizes), 1..ARB_MAX_FC_LAYERS; 0 = one layer of width d_model (v2 behaviour)  */
  int32_t fc_sizes[8];  /* fc_model.sizes (multiples of 4, <= 8192); fc_sizes[n_fc_layers-1] must equal d_model       */
  int32_t fc_act;       /* ARB_ACT_*: fc_model.activation applied after every FC linear, before its dropout          */
  int32_t fc_input_norm;/* fc_model.input_norm: nn.LayerNorm(n_features) (eps 1e-5, biased variance) on x first;
                           needs n_features % 4 == 0 (no feature padding)                                             */
  int32_t bf16;         /* 1: bf16 mode of the encoder (BASELINE config 3): every encoder linear runs as a tcgen05
                           kind::f16 product of bfloat16 operands with fp32 accumulation -- weights from a bfloat16
                           shadow of the fp32 master parameters (refreshed by every forward call), activations that only
                           feed products (LayerNorm outputs, attention context, FFN hidden layer) and the gradients that
                           only feed products stored as bfloat16; residual stream, LayerNorm statistics, softmax,
                           attention scores (TF32 products on fp32 Q/K/V), head, loss and parameter gradients stay fp32.
                           Needs the fused attention kernels (slate_length <= 256, head width 16 or 32) and
                           d_model, d_ff multiples of 8.  0: TF32 products on fp32 data everywhere.                  */
} arb_scorer_config;
#define ARB_MAX_FC_LAYERS 8

int64_t arb_scorer_param_count(const arb_scorer_config* cfg);
/* floats of activation workspace for a [B,S] batch; `training` != 0 keeps what backward needs.  Query it with the
 * configuration of the call itself (the dropout rates in particular: a call without dropout may run over packed rows,
 * arb_set_pack_rows, whose buffers differ). */
int64_t arb_scorer_workspace_floats(const arb_scorer_config* cfg, int32_t B, int32_t S, int32_t training);
/* x [B,S,F] fp32, mask [B,S] uint8 (1 = padded, train_utils.py:19) -> scores [B,S] fp32.
 * indices [B,S] int64 (original item ranks, -1 = padded; positional.py:45-50) and pe_table are only read when
 * cfg->pe_mode != 0 (pe_table: the fixed table for mode 1; ignored for mode 2, whose table lives in params). */
int32_t arb_scorer_forward(const arb_scorer_config* cfg, const float* params, const float* x, const uint8_t* mask,
                           const int64_t* indices, const float* pe_table,
                           int32_t B, int32_t S, float* scores, float* workspace, int64_t workspace_floats,
                           int32_t training, uint64_t seed, void* stream);
/* d_scores [B,S] -> grads += d loss / d params.  `workspace` is the one the training forward filled;
 * `scratch`: arb_scorer_backward_scratch_floats(cfg,B,S) floats. */
int64_t arb_scorer_backward_scratch_floats(const arb_scorer_config* cfg, int32_t B, int32_t S);
int32_t arb_scorer_backward(const arb_scorer_config* cfg, const float* params, const float* x, const uint8_t* mask,
                            const int64_t* indices, int32_t B, int32_t S, const float* scores, const float* d_scores, float* grads,
                            float* workspace, int64_t workspace_floats, float* scratch, int64_t scratch_floats,
                            uint64_t seed, void* stream);

/* 0: unfused attention (materialised logits + generic GEMMs); 1: fused tcgen05 attention forward kernel, unfused
 * backward; 2 (default): fused forward and backward kernels.  Fused kernels serve slates of <= 256 items (backward:
 * head width <= 32); other shapes use the unfused path automatically.  Process-wide; exists for A/B tests. */
void arb_set_attention_mode(int32_t mode);

/* 1 (default): the fused attention kernels skip the 128-item tiles that lie entirely inside a slate's padding -- keys
 * beyond the last real item have probability exactly 0 and rows beyond the last item that is real or carries a score
 * gradient have exactly zero activation gradients, so results are unchanged; 0: dense tiles.  For A/B measurements. */
void arb_set_attention_skip_padding(int32_t on);

/* Packed rows (padding removal).  1: the encoder -- every LayerNorm, linear, attention and head kernel and every
 * gradient product -- runs over the items below each slate's extent only (extent = last unmasked item + 1, rounded up to
 * 16 rows; the live row count stays on the device, nothing synchronises).  Exact for every real item's score and every
 * parameter gradient; the scores of the items beyond a slate's extent are then 0 (and carry no gradient) instead of
 * what the network computes for a padded feature row -- values every consumer in allRank masks (losses.py: the
 * padded_value_indicator masks, metrics.py:31-35, inference_utils.py:51).  Applies to calls with a transformer whose
 * attention runs in the fused kernels (slate_length <= 256, head width 16 / 32), no dropout, no positional encoding
 * and d_output = 1; other calls use the dense layout.  0: dense [B*S] rows everywhere, i.e. padded items scored like
 * the reference does.  Process-wide. */
void arb_set_pack_rows(int32_t on);

/* Fused attention backward: 1: one CTA per SM walks the (slate, head) items as one stream of tile iterations (an
 * item's first loads and products run behind the previous item's last iteration); 0: one CTA per item.  Same results.
 * Process-wide; exists for A/B measurements. */
void arb_set_attention_bwd_persistent(int32_t on);
int32_t arb_get_pack_rows(void);

/* 1 (default): the kernels of a step are chained with programmatic dependent launch -- a kernel's prologue (barrier
 * init, TMEM allocation, tensor-map prefetch) overlaps its predecessor's last wave, and it blocks in griddepcontrol.wait
 * before its first global-memory access; 0: every launch fully serialised.  For A/B measurements. */
void arb_set_pdl(int32_t on);

/* 1 (default): the fused attention forward runs as the two-pass / two-CTAs-per-SM kernel (head width <= 32);
 * 0: the single-pass kernel that keeps the whole S x S tile in TMEM (one CTA per SM). For A/B measurements. */
void arb_set_attention_fwd_two_pass(int32_t on);

/* GEMM kernel choice.  0: one CTA per output tile everywhere; 1: the persistent, decoupled-pipeline kernel (one CTA per
 * SM walking all tiles) wherever it is supported; 2 (default): persistent for every unbatched, non-split product except
 * short-K ones (K < 256) with a residual / mask tile -- the per-launch measurements of profiles/r2.  Process-wide; exists
 * for A/B measurements. */
void arb_set_gemm_persistent(int32_t on);

/* 1 (default): MMA operands are rounded fp32 -> tf32 by the TMA unit (TFLOAT32 tensor maps); 0: the tensor core
 * truncates.  Process-wide; exists for the precision tests. */
void arb_set_tf32_round_on_load(int32_t enable);

/* Building block exposed for tests: C = epilogue(alpha * A op B) on row-major fp32 matrices (csrc/gemm_tf32.cu). */
int32_t arb_gemm_tf32(const float* A, const float* B, float* C, const float* aux, const float* bias, int32_t M,
                      int32_t N, int32_t K, int32_t a_mn, int32_t b_mn, int32_t batch, int64_t a_bstride,
                      int64_t b_bstride, int64_t c_bstride, int32_t block_n, int32_t flags, float alpha,
                      int32_t split_k, void* stream);

/* The same building block with bf16 operands (tcgen05 kind::f16, fp32 accumulation) -- the matrix products of the
 * scorer's bf16 mode (BASELINE config 3).  A and B are bfloat16; C (and aux) bfloat16 when out_bf16 != 0, else fp32;
 * flags as for arb_gemm_tf32 (the ARB_GEMM_* values below); colsum_out: optional [N] bias-gradient accumulator. */
int32_t arb_gemm_bf16(const void* A, const void* B, void* C, const void* aux, const float* bias, int32_t M, int32_t N,
                      int32_t K, int32_t a_mn, int32_t b_mn, int32_t block_n, int32_t flags, float alpha,
                      int32_t split_k, int32_t out_bf16, float* colsum_out, void* stream);

/* ---------------------------------------------------------------- optimiser + profiling helpers
 * Flat Adam over the scorer's flat parameter/gradient buffers: torch.optim.Adam semantics (the optimiser the
 * reference instantiates from its config, allrank/main.py:82), one launch.  grads are multiplied by grad_scale
 * first (1/world_size after a sum all-reduce).  `step` is the 1-based step count. */
int32_t arb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                      void* stream);
/* The same with the step counter on the device: state[0] (a float, initially 0) is incremented by a one-thread prep
 * launch and the step's bias corrections are computed there -- what a CUDA-graph replay of a training step needs
 * (allrank_b200.graph.GraphedTrainStep); state holds 3 floats. */
int32_t arb_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, float* state, float grad_scale,
                          void* stream);

/* ---------------------------------------------------------------- slate movers
 * arb_assemble_slates: the per-slate FixLength transform + ToTensor + DataLoader collation of
 * allrank/data/dataset_loading.py:32-93,:230-247 for a batch of queries, from a corpus resident in HBM:
 *   docs_x [N,F] fp32 and docs_y [N] fp32 grouped by query, offsets [n_queries+1] int64 (CSR), queries [B] int64
 *   (a query number outside [0, n_queries) produces an all-padding slate).
 * Query shorter than S: rows in order, then zero rows with y = -1 and index = -1 (bit-identical to the reference).
 * Otherwise: S items sampled uniformly without replacement in random order; a sample without any relevant item is
 * redrawn while the query has one (the query's only relevant item, if its labels sum to 1, replaces the last
 * sampled item instead) -- :55-74.  The sample stream is a counter hash of (seed, slate, attempt, item).
 * max_query_len: longest query of the corpus (sizes the shared-memory sort; arb_assemble_slates_smem_bytes).
 * Outputs: x_out [B,S,F], y_out [B,S] fp32, idx_out [B,S] int64 (positions inside the query, -1 = padded). */
size_t arb_assemble_slates_smem_bytes(int32_t max_query_len, int32_t S);
int32_t arb_assemble_slates(const float* docs_x, const float* docs_y, const int64_t* offsets, int64_t n_queries,
                            const int64_t* queries, int32_t B, int32_t S, int32_t F, int32_t max_query_len,
                            uint64_t seed, float* x_out, float* y_out, int64_t* idx_out, void* stream);
/* inference_utils.__rank_slates (allrank/inference/inference_utils.py:37-60): x_out[b,r,:] = x[b,order[b,r],:],
 * y_out[b,r] = y[b,order[b,r]], with `order` the descending score ranking (arb_rank_metrics' out_order). */
int32_t arb_gather_slates(const float* x, const float* y, const int32_t* order, int32_t B, int32_t S, int32_t F,
                          float* x_out, float* y_out, void* stream);

/* Per-launch device timing for bench.py's roofline: enable, run steps, collect per kernel class
 * (0 = tcgen05 GEMM [work = flops], 1 = scorer SIMT, 2 = losses, 3 = metrics, 4 = optimiser, 5 = slate assembly /
 * gather [work = bytes]). */
void arb_prof_enable(int32_t on);
int32_t arb_prof_collect(int32_t cls, double* total_ms, double* total_work, int64_t* launches);
/* algorithmic HBM bytes (operands + outputs, each counted once) summed by the last arb_prof_collect(cls, ...) */
double arb_prof_last_bytes(int32_t cls);
/* Per-kernel table since arb_prof_enable(1), one text line per distinct launch name:
 *   name \t class \t launches \t total_ms \t total_work \t total_algorithmic_bytes
 * (GEMM launches are named by shape and operand layout, the other kernels by their launcher).  Writes at most
 * cap - 1 bytes + NUL into buf and returns the count; buf == NULL returns the size needed. */
int64_t arb_prof_report(char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* ALLRANK_B200_H */
