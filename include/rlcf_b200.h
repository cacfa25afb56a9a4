/* rlcf_b200.h -- C ABI of librlcf_b200.so: the sm_100a kernels behind the RLCF test-time-adaptation hot path.
 *
 * The reference (mzhaoshuai/RLCF) is pure Python over PyTorch and has no FFI of its own; every entry point
 * below replaces the ATen / cuBLAS / cuDNN library call that the cited reference line dispatches to
 * (SURVEY.md section 2.2, rows K1-K15).  Conventions:
 *   - plain device pointers + int sizes + a cudaStream_t (passed as void*); no torch types;
 *   - every function returns 0 on success or an RLCF_ERR_* code; rlcf_last_error() gives the message;
 *   - stateless: the caller owns all memory (PyTorch in the shipped host code); kernels are enqueued on the
 *     caller's stream and never synchronise; safe for one host thread per device;
 *   - matrices are row-major; "fp16" is IEEE binary16; accumulations and the residual stream are fp32.
 *
 * Row / parameter-set convention for the LayerNorm-family kernels: rows are grouped in contiguous runs of
 * `rows_per_set`; run g uses the affine parameters at gamma + g*param_stride (param_stride = 0 shares one
 * set).  This is how one batched launch serves many test images that each own a private copy of the
 * LayerNorm parameters (the reference adapts one image at a time, TPT/tune_cls_rl.py:192-222).
 */
#ifndef RLCF_B200_H_
#define RLCF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLCF_ABI_VERSION 1

enum { RLCF_OK = 0, RLCF_ERR_ARG = 1, RLCF_ERR_CUDA = 2, RLCF_ERR_DRIVER = 3 };

/* GEMM epilogues */
enum {
  RLCF_EPI_F16 = 0,          /* out16 = alpha*acc + bias                                   */
  RLCF_EPI_GELU_F16 = 1,     /* u = alpha*acc + bias; aux_out16 = u; out16 = u*sigmoid(1.702u)  (model.py:166-168) */
  RLCF_EPI_RESID_F32 = 2,    /* out32 = alpha*acc + bias + resid32                         (model.py:190-191) */
  RLCF_EPI_GELU_BWD_F16 = 3, /* out16 = alpha*acc * quickgelu'(aux_in16)                   */
  RLCF_EPI_F32 = 4,          /* out32 = alpha*acc + bias                                   */
  RLCF_EPI_ADAMW = 5         /* out32 (= the parameter) <- AdamW(param, m, v, alpha*acc): rlcf_gemm_wgrad_adamw only */
};

int rlcf_abi_version(void);
const char* rlcf_last_error(void);
/* Number of kernels this library has enqueued since load (bench.py's gpu_launches). */
uint64_t rlcf_launch_count(void);
/* 1 = one CTA per tile (UMMA 128x256), 2 = CTA pair per tile (cta_group::2, UMMA 256x256). Returns the value set. */
int rlcf_set_gemm_cta_group(int cta_group);
/* 1 = 4-CTA clusters: two CTA pairs share one weight tile through TMA multicast (512 x 256 cluster tile). */
int rlcf_set_gemm_multicast(int on);
/* Attention kernels (forward and backward): 0 = tcgen05/TMEM kernels (default; sequences beyond their TMEM layout --
 * forward L > 640 or causal L > 272, backward L > 384 -- use the warp-level kernels), 1 = warp-level mma.sync kernels. Returns the value set. */
int rlcf_set_attention_impl(int impl);

/* D[M,N] = A[M,K] * B[N,K]^T, fp16 operands (K contiguous), fp32 accumulate on tcgen05 tensor cores.
 * Replaces: nn.Conv2d patch embedding (TPT/clip/model.py:224), nn.MultiheadAttention in_proj / out_proj
 * (model.py:175,187), mlp.c_fc / c_proj (model.py:177-181) and their autograd dgrad (tpt_cls_rl.py:77).
 * N % 32 == 0, K % 8 == 0, lda/ldb/ldo % 8 == 0, 16-byte aligned pointers. */
int rlcf_gemm_f16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, float alpha,
                  const float* bias, const float* resid, const void* aux_in, void* aux_out, void* out, int ldo,
                  void* stream);

/* `groups` independent GEMMs of one shape in a single persistent launch (rank-3 TMA maps: k, row, group): group g
 * reads A + g*a_group_stride and B + g*b_group_stride, adds bias + g*bias_group_stride and writes
 * out / resid / aux + g*out_group_stride (strides in elements; A/B/out strides % 8 == 0, bias stride % 4 == 0).
 * Replaces the per-sample Linear / autograd calls once every test sample owns its weights: TTA steps >= 2 of full
 * encoder tuning (TPT/tpt_cls_rl.py:52-79 with custom_clip.py:477-479; retrieval/clip_ret_policy.py:84-103) and the
 * per-sample wgrad dW = dY^T X. */
int rlcf_gemm_f16_grouped(const void* A, int lda, int64_t a_group_stride, const void* B, int ldb, int64_t b_group_stride,
                          int groups, int M, int N, int K, int epilogue, float alpha, const float* bias,
                          int64_t bias_group_stride, const float* resid, const void* aux_in, void* aux_out, void* out,
                          int ldo, int64_t out_group_stride, void* stream);

/* images fp32 [*,C,H,W] -> patch rows fp16 [n_views*(H/p)*(W/p), k_pad] (column = c*p*p + ky*p + kx, zero padded
 * to k_pad); view_idx (device int32 [n_views], may be NULL = identity) picks the source views.
 * Together with rlcf_gemm_f16 replaces conv1 (model.py:224) and the inputs[selected_idx] gather (tpt_cls_rl.py:55,59). */
int rlcf_im2col_f16(const float* images, const int32_t* view_idx, int n_views, int C, int H, int W, int patch,
                    int k_pad, void* out, void* stream);

/* Token assembly + ln_pre (model.py:225-229): row (v,t) = LN((t==0 ? cls : patch_out[v*(L-1)+t-1]) + pos[t]).
 * x_pre (may be NULL) receives the pre-LN rows (needed by rlcf_layernorm_bwd for ln_pre). */
int rlcf_embed_lnpre(const float* patch_out, const float* cls, const float* pos, const float* gamma,
                     const float* beta, int64_t param_stride, int rows_per_set, int n_views, int L, int d, float eps,
                     float* x_pre, float* x, void* stream);

/* Text-side token assembly (model.py:343-345): x[(c,t)] = tok_emb[tokens[c,t]] + pos[t]. */
int rlcf_embed_text(const int64_t* tokens, const float* tok_emb, const float* pos, int n_seq, int L, int d, float* x,
                    void* stream);

/* Row LayerNorm in fp32 (model.py:157-163), eps inside the sqrt.  x rows are ldx floats apart.
 * out16 / out32 may each be NULL. */
int rlcf_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, int64_t param_stride,
                       int rows_per_set, int M, int d, float eps, void* out16, float* out32, void* stream);

/* LayerNorm backward.  dy is fp16 (dy_is_f32 = 0) or fp32 rows lddy apart; x is the saved LN input.
 *   dx_accum != NULL : dx_accum[row] (+)= dLN/dx   (accumulate = 1 adds to the residual-stream gradient)
 *   dx16 != NULL     : fp16 copy [rows, d] of the updated dx_accum rows (A operand of the next dgrad GEMM)
 *   partials         : (may be NULL when the LayerNorm is frozen) [n_sets, n_slots, p_total] fp32; block b of set g writes d(gamma) at
 *                      partials[g][b][p_off .. p_off+d) and d(beta) at [p_off+d .. p_off+2d).
 * n_slots blocks are launched per set. */
int rlcf_layernorm_bwd(const void* dy, int dy_is_f32, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                       int64_t param_stride, int rows_per_set, int n_sets, int d, float eps, float* dx_accum,
                       int64_t lddx, int accumulate, void* dx16, float* partials, int n_slots, int64_t p_total,
                       int64_t p_off, void* stream);

/* Fused multi-head attention core (nn.MultiheadAttention, model.py:185-187): qkv fp16 [n_seq*L, 3d] packed
 * q|k|v, head h = columns [h*64,(h+1)*64) of each third; out fp16 [n_seq*L, d].  causal != 0 applies the
 * text tower's additive -inf upper-triangular mask (model.py:328-334).  lse (may be NULL) receives the
 * per-row log-sum-exp [n_seq, heads, L] needed by the backward. head_dim is fixed at 64 (CLIP ViT/text). */
int rlcf_attention_fwd(const void* qkv, int n_seq, int L, int heads, int causal, void* out, float* lse, void* stream);
int rlcf_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int n_seq, int L,
                       int heads, int causal, void* dqkv, void* stream);

/* Attention output of ONE query row per sequence (row q_row; 0 = the class token): out fp16 [n_seq, d].  The last block
 * of VisionTransformer.forward feeds only x[:, 0, :] into ln_post (model.py:232-238), so an inference forward runs that
 * block's attention / out_proj / ln_2 / MLP on the class-token rows alone -- same values, 1/L of the work.  K and V
 * are read from every token of qkv [n_seq*L, 3d]; no mask (vision towers).  q_rows != NULL: the caller projected only
 * that row's query (fp16 [n_seq, d]) and qkv is k | v alone, [n_seq*L, 2d] -- the in_proj rows of the other tokens'
 * queries are not computed either.  x / x_row (both or neither): also copies row q_row of each sequence of the fp32
 * residual stream x [n_seq*L, d] into x_row [n_seq, d].  L <= 672. */
int rlcf_attention_row_fwd(const void* qkv, const void* q_rows, int n_seq, int L, int heads, int q_row, void* out,
                           const float* x, float* x_row, void* stream);

/* Head: LN(x[row]) @ proj -> L2 normalise -> logits = logit_scale * f . class_feat^T
 * (model.py:235-238, custom_clip.py:423-432 / clip_reward.py:130-137).
 * row of sequence v is row_idx[v] if row_idx != NULL else v*row_stride (in rows of d floats).
 * feat [n,E] gets the normalised features, inv_norm [n] 1/|f| (both may be NULL), logits [n,C] (NULL to skip). */
int rlcf_head_fwd(const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma, const float* beta,
                  int64_t param_stride, int seqs_per_set, const float* proj, const float* class_feat, float logit_scale,
                  int n, int d, int E, int C, float eps, float* feat, float* inv_norm, float* logits, void* stream);

/* select_confident_samples (tpt_cls_rl.py:32-35): per image, entropy of softmax(logits[v,:]) for its V views,
 * ascending order, first S kept.  sel [n_img,S] = view index inside the image; sel_global = img*V + sel;
 * entropy [n_img,V] (may be NULL). */
int rlcf_entropy_select(const float* logits, int n_img, int V, int C, int S, int32_t* sel, int32_t* sel_global,
                        float* entropy, void* stream);

/* top-K sampling + CLIPScore + reward post-processing + reward-weighted CE and its gradient
 * (tpt_cls_rl.py:63-71, clip_reward.py:111-128,152-165).  logits rows: row_idx[i] (NULL = i) for the
 * n_img*S selected views.  reward_img [n_img*S, Er] and reward_cls [C, Er] are L2-normalised.
 * dlogits [n_img*S, C] = loss_scale * dL/dlogits with L = mean over the image's S*K samples.
 * Optional outputs: topk_idx [n_img*S,K] int32, scores / rewards [n_img*S,K], loss [n_img]. */
int rlcf_reward_loss(const float* logits, const int32_t* row_idx, const float* reward_img, const float* reward_cls,
                     int n_img, int S, int K, int C, int Er, float clipscore_weight, int reward_process,
                     int process_batch, int amplify, float loss_scale, float* dlogits, int32_t* topk_idx,
                     float* scores, float* rewards, float* loss, void* stream);

/* TPT loss (config 1): marginal entropy of the S selected views (tpt_cls_rl.py:38-44) and its gradient. */
int rlcf_avg_entropy_loss(const float* logits, const int32_t* row_idx, int n_img, int S, int C, float loss_scale,
                          float* dlogits, float* loss, void* stream);
/* --min_entropy_reg (tpt_cls_rl.py:73-74): loss += weight * avg_entropy(output).  ACCUMULATES weight * loss_scale *
 * d avg_entropy / d logits into dlogits and weight * avg_entropy into loss (both already hold the RLCF term). */
int rlcf_avg_entropy_reg(const float* logits, const int32_t* row_idx, int n_img, int S, int C, float loss_scale,
                         float weight, float* dlogits, float* loss, void* stream);

/* Backward of rlcf_head_fwd for the S selected views of each image (one block per view):
 * dlogits -> d feat -> d(LN out) -> ln_post backward.  Writes dx into dres rows (row_idx as in head_fwd) and
 * view s's ln_post d(gamma), d(beta) into partials[g][s][p_off..p_off+2d)  (requires S <= n_slots). */
int rlcf_head_bwd(const float* dlogits, const float* x, const int32_t* row_idx, int64_t row_stride,
                  const float* gamma, int64_t param_stride, const float* proj, const float* class_feat,
                  float logit_scale, const float* feat, const float* inv_norm, int n_img, int S, int d, int E, int C,
                  float eps, float* dres, float* partials, int n_slots, int64_t p_total, int64_t p_off, void* stream);

/* Generalised head backward: sequence q of set g has dlogits[g*dl_set_stride + q*dl_seq_stride + k*dl_k_stride],
 * k < K, taken against other_feat + g*other_set_stride ([K,E]).  With the roles of image and text swapped this is the
 * backward of the TEXT head in prompt tuning (custom_clip.py:62-73,315-335): sequences = class prompts, K = selected
 * views, other_feat = that image's view features.  partials may be NULL (frozen LayerNorm). */
int rlcf_head_bwd_ex(const float* dlogits, int64_t dl_set_stride, int64_t dl_seq_stride, int64_t dl_k_stride,
                     const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma,
                     int64_t param_stride, const float* proj, const float* other_feat, int64_t other_set_stride,
                     float logit_scale, const float* feat, const float* inv_norm, int n_sets, int seqs_per_set, int d,
                     int E, int K, float eps, float* dres, float* partials, int n_slots, int64_t p_total, int64_t p_off,
                     const float* beta, float* y_out, float* df_out, void* stream);
/* (beta, y_out [n,d], df_out [n,E] may be NULL; when given, the LayerNorm output y and d(features before
 * normalisation) are kept for the projection's weight gradient.) */

/* Prompt assembly (PromptLearner.forward, class token at the end, custom_clip.py:198-232) + positional embedding:
 * x[(g,c,t)] = (1 <= t <= n_ctx ? ctx[g*ctx_stride + (t-1)*d ..] : tok_emb[tokens[c,t]]) + pos[t]. */
int rlcf_embed_prompts(const int64_t* tokens, const float* tok_emb, const float* pos, const float* ctx,
                       int64_t ctx_stride, int n_ctx, int n_sets, int n_cls, int L, int d, float* x, void* stream);

/* logits[g,s,c] = logit_scale * <img_feat[g,s,:], txt_feat[g*txt_set_stride + c*E ..]>  (custom_clip.py:325-335). */
int rlcf_pair_logits(const float* img_feat, const float* txt_feat, int64_t txt_set_stride, int n_sets, int S, int C,
                     int E, float logit_scale, float* logits, void* stream);

/* dctx[g,i,:] = sum over the n_cls prompts of set g of dx[(g*n_cls + c)*L + 1 + i, :]  (gradient of the learnable
 * context vectors, tpt_cls_rl.py:103-105,119-120). */
int rlcf_ctx_grad(const float* dx, int n_sets, int n_cls, int L, int n_ctx, int d, float* dctx, void* stream);
/* PromptLearner.forward for every layout (custom_clip.py:198-289: class token at the end, in the middle or at the front
 * of the context; learned class tokens, 209-221): src_map int32 [n_cls, L]: entry >= 0 = position inside the class's own
 * token row whose frozen embedding is used, entry < 0 = learnable vector -1 - entry of the set (vec + g*vec_stride:
 * context vectors, then one class vector per class).  x rows (g, c, t) = source + pos[t]. */
int rlcf_embed_prompts_map(const int64_t* tokens, const float* tok_emb, const float* pos, const float* vec,
                           int64_t vec_stride, const int32_t* src_map, int n_sets, int n_cls, int L, int d, float* x,
                           void* stream);
/* Its backward onto the learnable vectors: ctx_pos int32 [n_cls, n_ctx] = token position of context vector v in class c;
 * cls_pos int32 [n_cls] (NULL without learned class tokens) = position of class c's own vector.
 * dvec [n_sets, n_ctx (+ n_cls), d]. */
int rlcf_vec_grad_map(const float* dx, const int32_t* ctx_pos, const int32_t* cls_pos, int n_sets, int n_cls, int L,
                      int n_ctx, int d, float* dvec, void* stream);

/* Fused gradient reduction + AdamW (torch.optim.AdamW semantics, tune_cls_rl.py:79-81, tpt_cls_rl.py:76-79):
 * g = sum over slots of partials / loss_scale; decoupled weight decay; bias-corrected moments.
 * params/m/v: [n_sets, p_total]; grad_out (may be NULL) receives g. */
int rlcf_adamw_step(float* params, float* m, float* v, const float* partials, int n_sets, int n_slots,
                    int64_t p_total, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    float loss_scale, float* grad_out, void* stream);

/* AdamW reading the parameters from params_in + g*params_in_stride (0 = one shared initial copy) and writing them to
 * params[g]; fresh_state != 0 treats the moments as zero without reading them.  This is the first optimiser step
 * after the per-image reset (tune_cls_rl.py:210-213) for the 86 M-parameter full-tuning case without ever copying
 * the initial weights per image.  grads: [n_sets, n_slots, p_total]. */
int rlcf_adamw_step_from(float* params, float* m, float* v, const float* grads, int n_sets, int n_slots,
                         int64_t p_total, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                         float loss_scale, const float* params_in, int64_t params_in_stride, int fresh_state,
                         void* stream);

/* Per-set transposed fp16 copies for weight-gradient GEMMs: out[c, g*rows_pad + r] = in[row(g,r), c], zero padded to
 * rows_pad; skip_first = L drops the first of every L input rows (class token).  dW = dY^T X then runs on
 * rlcf_gemm_f16 with both operands K-major (autograd wgrad of model.py:175-181,224). */
int rlcf_transpose_blocks_f16(const void* in, int in_is_f32, int n_sets, int rows_per_set, int rows_pad, int cols,
                              int skip_first, int64_t in_set_stride_rows, void* out, int64_t ld_out, void* stream);
/* Bias gradients: out[g*out_stride + c] = sum over the set's rows of in[., c] (fp16 in, fp32 out). */
int rlcf_colsum_f16(const void* in, int n_sets, int rows_per_set, int cols, float* out, int64_t out_stride,
                    void* stream);
/* positional / class embedding gradient: out[g][t][:] = sum over the set's S sequences of dx[(g*S+s)*L + t][:]. */
int rlcf_seq_sum(const float* dx, int n_sets, int S, int L, int d, float* out, int64_t out_stride, void* stream);
/* projection gradient: out[g][i][j] = sum_s y[g*S+s][i] * df[g*S+s][j]. */
int rlcf_outer_sum(const float* y, const float* df, int n_sets, int S, int d, int E, float* out, int64_t out_stride,
                   void* stream);

/* model.reset() + optimizer.load_state_dict (tune_cls_rl.py:210-213) for the trainable slice only:
 * params[g] = init for every set g; m = v = 0. */
int rlcf_reset_params(const float* init, float* params, float* m, float* v, int n_sets, int64_t p_total,
                      void* stream);

/* fp32 -> fp16 with optional zero padding of each row from cols to ld_out (weight preparation). */
int rlcf_cast_f16(const float* in, int64_t rows, int64_t cols, int64_t ld_in, void* out, int64_t ld_out,
                  void* stream);
/* dst[l][j] = src[l][idx[j]] for l < n_layers, j < n, in blocks of seq_bytes bytes (one sequence's rows of one activation
 * tensor; sizes and pointers multiples of 4, 16-byte vectors when they allow).  The reference keeps the autograd graph of all 64 views and backpropagates
 * through output[selected_idx] (tpt_cls_rl.py:57-71); here the 64-view pass writes its per-layer activations to a store
 * and the selected views' are lifted out for the backward -- instead of running those views a second time. */
int rlcf_gather_seqs(const void* src, const int32_t* idx, void* dst, int64_t seq_bytes, int64_t src_layer_bytes,
                     int64_t dst_layer_bytes, int n_layers, int n, void* stream);
/* out16[c, r] = in32[r, c]  (W^T copies used as the B operand of dgrad GEMMs). */
int rlcf_transpose_cast_f16(const float* in, int rows, int cols, void* out, void* stream);

/* ---- per-sample weights (every test sample / query owns ALL parameters of the tuned tower) --------------------
 * Retrieval TTA tunes the whole image (or text) encoder for 8 steps per query (retrieval/clip_ret_policy.py:76-137,
 * custom_models.py:144-152) and full-encoder classification tuning does the same per image (custom_clip.py:477-479),
 * so after the first AdamW step the class/positional embeddings and the output projection differ per sample too.
 * These variants take the per-set stride of those tensors (in floats; 0 = shared). */
int rlcf_embed_lnpre_sets(const float* patch_out, const float* cls, const float* pos, int64_t embed_stride,
                          const float* gamma, const float* beta, int64_t param_stride, int rows_per_set, int n_views,
                          int L, int d, float eps, float* x_pre, float* x, void* stream);
int rlcf_head_fwd_sets(const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma,
                       const float* beta, int64_t param_stride, int seqs_per_set, const float* proj,
                       int64_t proj_stride, const float* class_feat, float logit_scale, int n, int d, int E, int C,
                       float eps, float* feat, float* inv_norm, float* logits, void* stream);
int rlcf_head_bwd_sets(const float* dlogits, int64_t dl_set_stride, int64_t dl_seq_stride, int64_t dl_k_stride,
                       const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma,
                       int64_t param_stride, const float* proj, int64_t proj_stride, const float* other_feat,
                       int64_t other_set_stride, float logit_scale, const float* feat, const float* inv_norm,
                       int n_sets, int seqs_per_set, int d, int E, int K, float eps, float* dres, float* partials,
                       int n_slots, int64_t p_total, int64_t p_off, const float* beta, float* y_out, float* df_out,
                       void* stream);
/* n_sets transposing casts in one launch: out16[g*out_stride + c*rows + r] = in32[g*in_stride + r*cols + c]. */
int rlcf_transpose_cast_f16_sets(const float* in, int rows, int cols, int n_sets, int64_t in_stride, void* out,
                                 int64_t out_stride, void* stream);

/* ---- retrieval TTA (retrieval/clip_ret_policy.py:76-137) ------------------------------------------------------
 * One query per row of `logits` [n_query, C] (row stride ld) against a gallery of C candidates.  Per query:
 * top-K candidates (torch.topk, clip_ret_policy.py:90/123), CLIPScore = max(0, w*<reward_gallery[idx],
 * reward_query>) (retrieval/clip_reward.py:143-168), rewards_post_process over the K samples, loss =
 * mean_k(r_k * CE(logits, idx_k)) (clip_ret_policy.py:97-98) and dlogits [n_query, C] = loss_scale * dL/dlogits.
 * K <= 32.  topk_idx [n_query,K] int32, scores / rewards [n_query,K], loss [n_query] may be NULL. */
int rlcf_retrieval_loss(const float* logits, int64_t ld, const float* reward_query, const float* reward_gallery,
                        int n_query, int K, int C, int Er, float clipscore_weight, int reward_process, int amplify,
                        float loss_scale, float* dlogits, int32_t* topk_idx, float* scores, float* rewards,
                        float* loss, void* stream);
/* partial[q, chunk, :] = sum over the chunk's candidates c of dlogits[q,c] * gallery[c,:]  (gallery [C,E] fp32).
 * The chunks are summed, in order, by rlcf_head_bwd_sets called with K = n_chunks and other_feat = partial, which
 * makes d(query feature) = dlogits @ gallery deterministic for any gallery size. */
int rlcf_dfeat_partial(const float* dlogits, const float* gallery, int n_query, int C, int E, int n_chunks,
                       float* partial, void* stream);
/* out[q*out_stride] = scale * sum_c a[q,c]*b[q,c]: the gradient of logit_scale (tuned in text->image retrieval,
 * custom_models.py:144-152) is sum_c dlogits[c]*logits[c]. */
int rlcf_rowdot(const float* a, const float* b, int n_rows, int C, float scale, float* out, int64_t out_stride,
                void* stream);

/* fp32 (CUDA-core) path of the once-per-dataset class text features (CLIP.encode_text, TPT/clip/model.py:342-356;
 * custom_clip.py:404-408; clip_reward.py:139-150): kept at the reference's precision because they are an input of every
 * per-image step.  out[M,N] = epi(A[M,K] W[N,K]^T + bias), all fp32; epilogue 0 = none, 1 = QuickGELU (model.py:166-168),
 * 2 = + resid (same layout as out).  N, K, lda, ldw, ldo multiples of 4. */
int rlcf_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, int M, int N, int K, int epilogue,
                  const float* bias, const float* resid, float* out, int64_t ldo, void* stream);
/* nn.MultiheadAttention core in fp32: qkv fp32 [n_seq*L, 3d] packed q|k|v -> out fp32 [n_seq*L, d]; causal as above. */
int rlcf_attention_f32(const float* qkv, int n_seq, int L, int heads, int causal, float* out, void* stream);

/* Top-1 / top-5 hit counters of `accuracy` (TPT/utils/tools.py:84-98) as tune_cls_rl.py:243-247 accumulates them:
 * hits[0] += #rows whose target is the arg-max, hits[1] += #rows whose target is among the 5 largest logits,
 * hits[2] += n (device int64 [3], accumulated with integer atomics; ties go to the lower class index). */
int rlcf_accuracy_count(const float* logits, const int64_t* target, int n, int C, int64_t* hits, void* stream);

/* ---- text->image retrieval: the text tower is tuned, including the caption's token-embedding rows, the positional
 * embedding and logit_scale (custom_models.py:144-152; tune_text, clip_ret_policy.py:106-137) ----
 * x[g*n + i] = a[g*a_stride + i] + b[g*b_stride + i]: per-query token rows + positional embedding (CLIP.encode_text,
 * open_clip model: x = token_embedding(text) + positional_embedding). */
int rlcf_add_rows(const float* a, int64_t a_stride, const float* b, int64_t b_stride, int n_sets, int64_t n, float* x,
                  void* stream);
/* out[q,c] = in[q,c] * exp(ls[q*ls_stride]): logits = logit_scale.exp() * cos with a per-query, trainable
 * logit_scale (forward), and d cos = exp(ls) * dlogits (backward). in == out allowed. */
int rlcf_scale_rows_exp(const float* in, const float* ls, int64_t ls_stride, int n_rows, int C, float* out,
                        void* stream);
/* Gradient of the per-query embedding rows from dx [n_sets*L, d]: g_pos[g][t] = dx[g,t]; g_tok[g][t] = sum of dx over
 * the positions of query g that hold the same token id as position t (weight tying of nn.Embedding rows). */
int rlcf_tied_rows_grad(const float* dx, const int64_t* tokens, int n_sets, int L, int d, float* g_tok, float* g_pos,
                        int64_t out_stride, void* stream);

/* torch.optim.AdamW over EVERY parameter of a tuned encoder, one parameter vector of p_total floats per sample
 * (TPT/tune_cls_rl.py:79-81 with custom_clip.py:477-479; retrieval/clip_ret_policy.py:235): streaming 16-byte pass,
 * params_in / fresh_state as in rlcf_adamw_step_from, and the fp16 copy of the first n16 parameters (the GEMM weights)
 * is written to w16 + g*w16_stride in the same pass (w16 may be NULL).  Sizes and strides are multiples of 4. */
int rlcf_adamw_full(float* params, float* m, float* v, const float* grads, int n_sets, int64_t p_total, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, float loss_scale,
                    const float* params_in, int64_t params_in_stride, int fresh_state, void* w16, int64_t w16_stride,
                    int64_t n16, int64_t set_stride, void* stream);
/* (set_stride: floats between consecutive samples' vectors in params / m / v / grads; 0 = p_total.  A larger stride
 * updates a sub-range of every sample's vector, e.g. everything after the GEMM weights that rlcf_gemm_wgrad_adamw owns.) */
/* out16[g][c][r] = in16[g][r][c] for n_sets matrices laid out set_stride elements apart (in and out share it). */
int rlcf_transpose_f16_sets(const void* in, int rows, int cols, void* out, int n_sets, int64_t set_stride,
                            void* stream);

/* ---- on-device view generation (TPT/data/datautils.py:76-128, TPT/data/augmix_ops.py) --------------------------
 * rlcf_resample_u8: n_views crops of ONE decoded uint8 image src [H,W,3], each resized to [out_h,out_w,3] exactly as
 * PIL's Image.resize does (libImaging/Resample.c: separable, 22-bit fixed-point taps, uint8 rounding after each pass).
 * Replaces transforms.Resize + CenterCrop (view 0) and RandomResizedCrop + RandomHorizontalFlip (datautils.py:89-92).
 * The host supplies what Pillow precomputes per call: hdr [n_views,8] int32 = {x0, y0, row_first, n_rows, flip,0,0,0}
 * (crop origin, first source row below y0 that is needed and how many), hb / vb [n_views,out,2] = (first tap, tap
 * count) and hk / vk [n_views,out,ks] = taps.  tmp: uint8 scratch [n_views,tmp_rows,out_w,3], tmp_rows >= max n_rows. */
int rlcf_resample_u8(const uint8_t* src, int H, int W, int n_views, const int32_t* hdr, const int32_t* hb,
                     const int32_t* hk, int ks_h, const int32_t* vb, const int32_t* vk, int ks_v, int out_h, int out_w,
                     uint8_t* tmp, int tmp_rows, uint8_t* out, void* stream);
/* AugMix (datautils.py:95-111) on n_views 224x224 uint8 crops x_orig [n_views,224,224,3]: per view up to 3 chains of
 * up to 3 operations -- type 0 autocontrast, 1 equalize, 2 posterize(param = bits), 3 solarize(param = threshold),
 * 4 affine with BILINEAR resampling (rotate / shear / translate; coefficients in mats) -- ToTensor + Normalize and
 * the blend m*x + (1-m)*sum_i w_i*aug_i, all with PIL's / torch's arithmetic.  vflag[v] = 0 emits the plain
 * normalised view.  wts [n_views,4] = (w0,w1,w2,m), omm [n_views] = float32(1-m), n_ops [n_views,3],
 * ops [n_views,3,3,2] int32, mats [n_views,3,3,6] double, out fp32 [n_views,3,224,224]. */
int rlcf_augmix_views(const uint8_t* x_orig, int n_views, const int32_t* vflag, const float* wts, const float* omm,
                      const int32_t* n_ops, const int32_t* ops, const double* mats, float mean0, float mean1,
                      float mean2, float std0, float std1, float std2, float* out, void* stream);

/* rlcf_transpose_blocks_f16 for fp16 input with the per-set column sums in the same pass: colsum[g*colsum_stride + c] =
 * sum_r in[row(g,r), c] (the bias gradient of the Linear whose output gradient is being laid out for the wgrad GEMM;
 * colsum may be NULL).  rows_pad, cols, ld_out even. */
int rlcf_transpose_blocks_colsum(const void* in, int n_sets, int rows_per_set, int rows_pad, int cols, int skip_first,
                                 int64_t in_set_stride_rows, void* out, int64_t ld_out, float* colsum,
                                 int64_t colsum_stride, void* stream);

/* The tap tables rlcf_resample_u8 consumes, computed on the device exactly as Pillow's precompute_coeffs +
 * normalize_coeffs_8bpc do (double precision, same operation order).  geom [n_views,8] int32 = {in_w, in_h, res_w,
 * res_h, lo_x, lo_y, filter (0 = BILINEAR, 1 = BICUBIC), 0}: a source region of in_w x in_h pixels is resized to
 * res_w x res_h, of which the window of out x out outputs at (lo_x, lo_y) is kept.  ks_h / ks_v >= the largest tap
 * count (ceil(support * max(scale, 1)) * 2 + 1).  Fills hb/hk/vb/vk and hdr[:,2:4] (hdr[:,0:2] and [:,4], the crop
 * origin and the flip flag, are the caller's). */
int rlcf_resample_taps(const int32_t* geom, int n_views, int out, int ks_h, int ks_v, int32_t* hdr, int32_t* hb,
                       int32_t* hk, int32_t* vb, int32_t* vk, void* stream);

/* Weight gradient + optimizer in ONE kernel: for every group g (test sample), dW_g = A_g @ B_g^T (A = dY^T [N_out, K =
 * rows], B = X^T [N_in, K]) is accumulated in TMEM and the epilogue applies torch.optim.AdamW to the sample's fp32 master
 * weight tile right there: it reads the parameter (from params_in + g*params_in_gs; pass the shared initial copy with
 * stride 0 for the first step) and the moments (skipped when fresh_state), writes parameter / moments back and the
 * fp16 copy the next forward reads -- the gradient never goes to HBM (8 B per weight per step less traffic than a
 * wgrad GEMM followed by rlcf_adamw_full).  The accumulator carries loss_scale; bias correction uses `step`.
 * Replaces autograd's wgrad + torch.optim.AdamW.step + autocast's cast for one Linear weight
 * (TPT/tpt_cls_rl.py:76-79, retrieval/clip_ret_policy.py:100-103,134-137). */
int rlcf_gemm_wgrad_adamw(const void* A, int lda, int64_t a_group_stride, const void* B, int ldb, int64_t b_group_stride,
                          int groups, int n_out, int n_in, int K, float* params, float* m, float* v, int ldp,
                          int64_t param_group_stride, const float* params_in, int64_t params_in_gs, int fresh_state,
                          void* w16, int64_t w16_group_stride, float lr, float beta1, float beta2, float eps,
                          float weight_decay, int step, float loss_scale, void* stream);

/* rlcf_reward_loss with an ensemble of up to 4 frozen reward models (CLIPRewardsMultiple, TPT/clip_reward.py:180-307):
 * model i contributes weight_i * max(0, w * <reward_cls_i[idx], reward_img_i>) to a sample's CLIPScore (weights = the
 * normalised confidences of clip_reward.py:207, or 1/n for weighted_scores = False).  Unused slots: NULL / 0. */
int rlcf_reward_loss_multi(const float* logits, const int32_t* row_idx, int n_models, const float* reward_img0,
                           const float* reward_img1, const float* reward_img2, const float* reward_img3,
                           const float* reward_cls0, const float* reward_cls1, const float* reward_cls2,
                           const float* reward_cls3, int er0, int er1, int er2, int er3, float weight0, float weight1,
                           float weight2, float weight3, int n_img, int S, int K, int C, float clipscore_weight,
                           int reward_process, int process_batch, int amplify, float loss_scale, float* dlogits,
                           int32_t* topk_idx, float* scores, float* rewards, float* loss, void* stream);

/* nn.functional.interpolate(images[view_idx], size=(oh, ow), mode="bicubic", align_corners=True): the resize of the
 * selected views to a reward model's own input resolution (TPT/clip_reward.py:133-134).  images fp32 [*, C, H, W],
 * view_idx int32 [n_views] or NULL, out fp32 [n_views, C, oh, ow]. */
int rlcf_bicubic_resize(const float* images, const int32_t* view_idx, int n_views, int C, int H, int W, int oh, int ow,
                        float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLCF_B200_H_ */
