// extern "C" surface of librlcf_b200.so (declared in include/rlcf_b200.h) plus the small amount of
// process-wide state the library keeps: last-error text, launch counter, GEMM CTA-group mode.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rlcf_internal.h"

namespace rlcf {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_cta_group{0};
static std::atomic<int> g_attn_impl{-1};
static std::atomic<int> g_gemm_mc{-1};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int gemm_cta_group() {
  int v = g_cta_group.load(std::memory_order_relaxed);
  if (v == 0) {
    const char* e = getenv("RLCF_GEMM_CTA_GROUP");
    v = (e != nullptr && atoi(e) == 1) ? 1 : 2;
    g_cta_group.store(v, std::memory_order_relaxed);
  }
  return v;
}

int gemm_multicast() {
  int v = g_gemm_mc.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("RLCF_GEMM_MULTICAST");
    v = (e != nullptr) ? (atoi(e) != 0) : 0;
    g_gemm_mc.store(v, std::memory_order_relaxed);
  }
  return v;
}

int attention_impl() {
  int v = g_attn_impl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("RLCF_ATTN_IMPL");
    v = (e != nullptr && atoi(e) == 1) ? 1 : 0;
    g_attn_impl.store(v, std::memory_order_relaxed);
  }
  return v;
}

PFN_encodeTiled get_encode_tiled() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<PFN_encodeTiled>(fn);
}

// implemented in kernels.cu / attention.cu
int im2col_f16(const float*, const int32_t*, int, int, int, int, int, int, __half*, cudaStream_t);
int embed_lnpre(const float*, const float*, const float*, const float*, const float*, long long, int, int, int, int,
                float, float*, float*, long long, cudaStream_t);
int embed_text(const long long*, const float*, const float*, int, int, int, float*, cudaStream_t);
int layernorm_fwd(const float*, long long, const float*, const float*, long long, int, int, int, float, __half*,
                  float*, cudaStream_t);
int layernorm_bwd(const void*, int, long long, const float*, long long, const float*, long long, int, int, int, float,
                  float*, long long, int, __half*, float*, int, long long, long long, cudaStream_t);
int attention_fwd(const __half*, int, int, int, int, __half*, float*, cudaStream_t);
int attention_row_fwd(const __half*, const __half*, int, int, int, int, __half*, const float*, float*, cudaStream_t);
int attention_bwd(const __half*, const __half*, const __half*, const float*, int, int, int, int, __half*,
                  cudaStream_t);
int head_fwd(const float*, const int32_t*, long long, const float*, const float*, long long, int, const float*,
             const float*, float, int, int, int, int, float, float*, float*, float*, long long, cudaStream_t);
int entropy_select(const float*, int, int, int, int, int32_t*, int32_t*, float*, cudaStream_t);
int reward_loss(const float*, const int32_t*, const float*, const float*, int, int, int, int, int, float, int, int,
                int, float, float*, int32_t*, float*, float*, float*, cudaStream_t);
int reward_loss_multi(const float*, const int32_t*, int, const float* const*, const float* const*, const int*,
                      const float*, int, int, int, int, float, int, int, int, float, float*, int32_t*, float*, float*,
                      float*, cudaStream_t);
int avg_entropy_loss(const float*, const int32_t*, int, int, int, float, float*, float*, cudaStream_t, float = 1.f,
                     int = 0);
int head_bwd(const float*, const float*, const int32_t*, long long, const float*, long long, const float*,
             const float*, float, const float*, const float*, int, int, int, int, int, float, float*, float*, int,
             long long, long long, long long, long long, long long, long long, const float*, float*, float*,
             long long, cudaStream_t);
int transpose_blocks(const void*, int, int, int, int, int, int, long long, __half*, long long, cudaStream_t);
int colsum_f16(const __half*, int, int, int, float*, long long, cudaStream_t);
int transpose_blocks_colsum(const __half*, int, int, int, int, int, long long, __half*, long long, float*, long long,
                            cudaStream_t);
int seq_sum(const float*, int, int, int, int, float*, long long, cudaStream_t);
int outer_sum(const float*, const float*, int, int, int, int, float*, long long, cudaStream_t);
int embed_prompts(const long long*, const float*, const float*, const float*, long long, int, int, int, int, int,
                  float*, cudaStream_t);
int pair_logits(const float*, const float*, long long, int, int, int, int, float, float*, cudaStream_t);
int ctx_grad(const float*, int, int, int, int, int, float*, cudaStream_t);
int embed_prompts_map(const long long*, const float*, const float*, const float*, long long, const int*, int, int, int,
                      int, float*, cudaStream_t);
int vec_grad_map(const float*, const int*, const int*, int, int, int, int, int, float*, cudaStream_t);
int adamw_step(float*, float*, float*, const float*, int, int, long long, float, float, float, float, float, int,
               float, float*, const float*, long long, int, cudaStream_t);
int reset_params(const float*, float*, float*, float*, int, long long, cudaStream_t);
int cast_f16(const float*, long long, long long, long long, __half*, long long, cudaStream_t);
int gather_seqs(const void*, const int*, void*, long long, long long, long long, int, int, cudaStream_t);
int transpose_cast_f16(const float*, int, int, __half*, int, long long, long long, cudaStream_t);
int retrieval_loss(const float*, long long, const float*, const float*, int, int, int, int, float, int, int, float,
                   float*, int32_t*, float*, float*, float*, cudaStream_t);
int dfeat_partial(const float*, const float*, int, int, int, int, float*, cudaStream_t);
int rowdot(const float*, const float*, int, int, float, float*, long long, cudaStream_t);
int adamw_full(float*, float*, float*, const float*, int, long long, float, float, float, float, float, int, float,
               const float*, long long, int, __half*, long long, long long, long long, cudaStream_t);
int transpose_f16(const __half*, int, int, __half*, int, long long, cudaStream_t);
int resample_u8(const uint8_t*, int, int, int, const int*, const int*, const int*, int, const int*, const int*, int, int,
                int, uint8_t*, int, uint8_t*, cudaStream_t);
int resample_taps(const int*, int, int, int, int, int*, int*, int*, int*, int*, cudaStream_t);
int bicubic_resize(const float*, const int32_t*, int, int, int, int, int, int, float*, cudaStream_t);
int augmix_views(const uint8_t*, int, const int*, const float*, const float*, const int*, const int*, const double*,
                 const float*, const float*, float*, cudaStream_t);
int add_rows(const float*, long long, const float*, long long, int, long long, float*, cudaStream_t);
int scale_rows_exp(const float*, const float*, long long, int, int, float*, cudaStream_t);
int tied_rows_grad(const float*, const long long*, int, int, int, float*, float*, long long, cudaStream_t);
int accuracy_count(const float*, const long long*, int, int, long long*, cudaStream_t);
int gemm_f32(const float*, long long, const float*, long long, int, int, int, int, const float*, const float*, float*,
             long long, cudaStream_t);
int attention_f32(const float*, int, int, int, int, float*, cudaStream_t);

}  // namespace rlcf

using namespace rlcf;
#define S(x) static_cast<cudaStream_t>(x)
#define H(x) static_cast<__half*>(x)
#define CH(x) static_cast<const __half*>(x)

extern "C" {

int rlcf_abi_version(void) { return RLCF_ABI_VERSION; }
const char* rlcf_last_error(void) { return g_err; }
uint64_t rlcf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int rlcf_set_gemm_cta_group(int cta_group) {
  if (cta_group == 1 || cta_group == 2) g_cta_group.store(cta_group, std::memory_order_relaxed);
  return gemm_cta_group();
}

int rlcf_set_gemm_multicast(int on) {
  if (on == 0 || on == 1) g_gemm_mc.store(on, std::memory_order_relaxed);
  return gemm_multicast();
}

int rlcf_set_attention_impl(int impl) {
  if (impl == 0 || impl == 1) g_attn_impl.store(impl, std::memory_order_relaxed);
  return attention_impl();
}

int rlcf_gemm_f16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, float alpha,
                  const float* bias, const float* resid, const void* aux_in, void* aux_out, void* out, int ldo,
                  void* stream) {
  if (A == nullptr || B == nullptr || out == nullptr) return set_error(RLCF_ERR_ARG, "gemm: null pointer");
  return gemm_f16(CH(A), lda, CH(B), ldb, M, N, K, epilogue, alpha, bias, resid, CH(aux_in), H(aux_out), out, ldo,
                  S(stream));
}

int rlcf_gemm_f16_grouped(const void* A, int lda, int64_t a_group_stride, const void* B, int ldb, int64_t b_group_stride,
                          int groups, int M, int N, int K, int epilogue, float alpha, const float* bias,
                          int64_t bias_group_stride, const float* resid, const void* aux_in, void* aux_out, void* out,
                          int ldo, int64_t out_group_stride, void* stream) {
  if (A == nullptr || B == nullptr || out == nullptr) return set_error(RLCF_ERR_ARG, "gemm: null pointer");
  return gemm_f16_grouped(CH(A), lda, a_group_stride, CH(B), ldb, b_group_stride, groups, M, N, K, epilogue, alpha, bias,
                          bias_group_stride, resid, CH(aux_in), H(aux_out), out, ldo, out_group_stride, S(stream));
}

int rlcf_im2col_f16(const float* images, const int32_t* view_idx, int n_views, int C, int H_, int W, int patch,
                    int k_pad, void* out, void* stream) {
  if (images == nullptr || out == nullptr) return set_error(RLCF_ERR_ARG, "im2col: null pointer");
  return im2col_f16(images, view_idx, n_views, C, H_, W, patch, k_pad, H(out), S(stream));
}

int rlcf_embed_lnpre(const float* patch_out, const float* cls, const float* pos, const float* gamma,
                     const float* beta, int64_t param_stride, int rows_per_set, int n_views, int L, int d, float eps,
                     float* x_pre, float* x, void* stream) {
  if (!patch_out || !cls || !pos || !gamma || !beta || !x) return set_error(RLCF_ERR_ARG, "embed_lnpre: null pointer");
  return embed_lnpre(patch_out, cls, pos, gamma, beta, param_stride, rows_per_set, n_views, L, d, eps, x_pre, x, 0,
                     S(stream));
}

int rlcf_embed_lnpre_sets(const float* patch_out, const float* cls, const float* pos, int64_t embed_stride,
                          const float* gamma, const float* beta, int64_t param_stride, int rows_per_set, int n_views,
                          int L, int d, float eps, float* x_pre, float* x, void* stream) {
  if (!patch_out || !cls || !pos || !gamma || !beta || !x)
    return set_error(RLCF_ERR_ARG, "embed_lnpre_sets: null pointer");
  return embed_lnpre(patch_out, cls, pos, gamma, beta, param_stride, rows_per_set, n_views, L, d, eps, x_pre, x,
                     embed_stride, S(stream));
}

int rlcf_embed_text(const int64_t* tokens, const float* tok_emb, const float* pos, int n_seq, int L, int d, float* x,
                    void* stream) {
  if (!tokens || !tok_emb || !pos || !x) return set_error(RLCF_ERR_ARG, "embed_text: null pointer");
  return embed_text(reinterpret_cast<const long long*>(tokens), tok_emb, pos, n_seq, L, d, x, S(stream));
}

int rlcf_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, int64_t param_stride,
                       int rows_per_set, int M, int d, float eps, void* out16, float* out32, void* stream) {
  if (!x || !gamma || !beta || (!out16 && !out32)) return set_error(RLCF_ERR_ARG, "layernorm_fwd: null pointer");
  return layernorm_fwd(x, ldx, gamma, beta, param_stride, rows_per_set, M, d, eps, H(out16), out32, S(stream));
}

int rlcf_layernorm_bwd(const void* dy, int dy_is_f32, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                       int64_t param_stride, int rows_per_set, int n_sets, int d, float eps, float* dx_accum,
                       int64_t lddx, int accumulate, void* dx16, float* partials, int n_slots, int64_t p_total,
                       int64_t p_off, void* stream) {
  if (!dy || !x || !gamma) return set_error(RLCF_ERR_ARG, "layernorm_bwd: null pointer");
  return layernorm_bwd(dy, dy_is_f32, lddy, x, ldx, gamma, param_stride, rows_per_set, n_sets, d, eps, dx_accum, lddx,
                       accumulate, H(dx16), partials, n_slots, p_total, p_off, S(stream));
}

int rlcf_attention_fwd(const void* qkv, int n_seq, int L, int heads, int causal, void* out, float* lse,
                       void* stream) {
  if (!qkv || !out) return set_error(RLCF_ERR_ARG, "attention_fwd: null pointer");
  return attention_fwd(CH(qkv), n_seq, L, heads, causal, H(out), lse, S(stream));
}

int rlcf_attention_row_fwd(const void* qkv, const void* q_rows, int n_seq, int L, int heads, int q_row, void* out,
                           const float* x, float* x_row, void* stream) {
  if (!qkv || !out) return set_error(RLCF_ERR_ARG, "attention_row_fwd: null pointer");
  return attention_row_fwd(CH(qkv), q_rows ? CH(q_rows) : nullptr, n_seq, L, heads, q_row, H(out), x, x_row, S(stream));
}

int rlcf_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int n_seq, int L,
                       int heads, int causal, void* dqkv, void* stream) {
  if (!qkv || !out || !dout || !dqkv) return set_error(RLCF_ERR_ARG, "attention_bwd: null pointer");
  return attention_bwd(CH(qkv), CH(out), CH(dout), lse, n_seq, L, heads, causal, H(dqkv), S(stream));
}

int rlcf_head_fwd(const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma, const float* beta,
                  int64_t param_stride, int seqs_per_set, const float* proj, const float* class_feat, float logit_scale,
                  int n, int d, int E, int C, float eps, float* feat, float* inv_norm, float* logits, void* stream) {
  if (!x || !gamma || !beta || !proj) return set_error(RLCF_ERR_ARG, "head_fwd: null pointer");
  return head_fwd(x, row_idx, row_stride, gamma, beta, param_stride, seqs_per_set, proj, class_feat, logit_scale, n, d,
                  E, C, eps, feat, inv_norm, logits, 0, S(stream));
}

int rlcf_head_fwd_sets(const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma,
                       const float* beta, int64_t param_stride, int seqs_per_set, const float* proj,
                       int64_t proj_stride, const float* class_feat, float logit_scale, int n, int d, int E, int C,
                       float eps, float* feat, float* inv_norm, float* logits, void* stream) {
  if (!x || !gamma || !beta || !proj) return set_error(RLCF_ERR_ARG, "head_fwd_sets: null pointer");
  return head_fwd(x, row_idx, row_stride, gamma, beta, param_stride, seqs_per_set, proj, class_feat, logit_scale, n, d,
                  E, C, eps, feat, inv_norm, logits, proj_stride, S(stream));
}

int rlcf_entropy_select(const float* logits, int n_img, int V, int C, int S_, int32_t* sel, int32_t* sel_global,
                        float* entropy, void* stream) {
  if (!logits || !sel) return set_error(RLCF_ERR_ARG, "entropy_select: null pointer");
  return entropy_select(logits, n_img, V, C, S_, sel, sel_global, entropy, S(stream));
}

int rlcf_reward_loss(const float* logits, const int32_t* row_idx, const float* reward_img, const float* reward_cls,
                     int n_img, int S_, int K, int C, int Er, float clipscore_weight, int reward_process,
                     int process_batch, int amplify, float loss_scale, float* dlogits, int32_t* topk_idx,
                     float* scores, float* rewards, float* loss, void* stream) {
  if (!logits || !reward_img || !reward_cls || !dlogits) return set_error(RLCF_ERR_ARG, "reward_loss: null pointer");
  return reward_loss(logits, row_idx, reward_img, reward_cls, n_img, S_, K, C, Er, clipscore_weight, reward_process,
                     process_batch, amplify, loss_scale, dlogits, topk_idx, scores, rewards, loss, S(stream));
}

int rlcf_avg_entropy_loss(const float* logits, const int32_t* row_idx, int n_img, int S_, int C, float loss_scale,
                          float* dlogits, float* loss, void* stream) {
  if (!logits || !dlogits) return set_error(RLCF_ERR_ARG, "avg_entropy_loss: null pointer");
  return avg_entropy_loss(logits, row_idx, n_img, S_, C, loss_scale, dlogits, loss, S(stream));
}

int rlcf_avg_entropy_reg(const float* logits, const int32_t* row_idx, int n_img, int S_, int C, float loss_scale,
                         float weight, float* dlogits, float* loss, void* stream) {
  if (!logits || !dlogits) return set_error(RLCF_ERR_ARG, "avg_entropy_reg: null pointer");
  return avg_entropy_loss(logits, row_idx, n_img, S_, C, loss_scale, dlogits, loss, S(stream), weight, 1);
}

int rlcf_head_bwd(const float* dlogits, const float* x, const int32_t* row_idx, int64_t row_stride,
                  const float* gamma, int64_t param_stride, const float* proj, const float* class_feat,
                  float logit_scale, const float* feat, const float* inv_norm, int n_img, int S_, int d, int E, int C,
                  float eps, float* dres, float* partials, int n_slots, int64_t p_total, int64_t p_off, void* stream) {
  if (!dlogits || !x || !gamma || !proj || !class_feat || !feat || !inv_norm || !dres || !partials)
    return set_error(RLCF_ERR_ARG, "head_bwd: null pointer");
  return head_bwd(dlogits, x, row_idx, row_stride, gamma, param_stride, proj, class_feat, logit_scale, feat, inv_norm,
                  n_img, S_, d, E, C, eps, dres, partials, n_slots, p_total, p_off, static_cast<long long>(S_) * C, C, 1,
                  0, nullptr, nullptr, nullptr, 0, S(stream));
}

int rlcf_head_bwd_ex(const float* dlogits, int64_t dl_set_stride, int64_t dl_seq_stride, int64_t dl_k_stride,
                     const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma,
                     int64_t param_stride, const float* proj, const float* other_feat, int64_t other_set_stride,
                     float logit_scale, const float* feat, const float* inv_norm, int n_sets, int seqs_per_set, int d,
                     int E, int K, float eps, float* dres, float* partials, int n_slots, int64_t p_total, int64_t p_off,
                     const float* beta, float* y_out, float* df_out, void* stream) {
  if (!dlogits || !x || !gamma || !proj || !other_feat || !feat || !inv_norm || !dres)
    return set_error(RLCF_ERR_ARG, "head_bwd_ex: null pointer");
  return head_bwd(dlogits, x, row_idx, row_stride, gamma, param_stride, proj, other_feat, logit_scale, feat, inv_norm,
                  n_sets, seqs_per_set, d, E, K, eps, dres, partials, n_slots, p_total, p_off, dl_set_stride,
                  dl_seq_stride, dl_k_stride, other_set_stride, beta, y_out, df_out, 0, S(stream));
}

int rlcf_head_bwd_sets(const float* dlogits, int64_t dl_set_stride, int64_t dl_seq_stride, int64_t dl_k_stride,
                       const float* x, const int32_t* row_idx, int64_t row_stride, const float* gamma,
                       int64_t param_stride, const float* proj, int64_t proj_stride, const float* other_feat,
                       int64_t other_set_stride, float logit_scale, const float* feat, const float* inv_norm,
                       int n_sets, int seqs_per_set, int d, int E, int K, float eps, float* dres, float* partials,
                       int n_slots, int64_t p_total, int64_t p_off, const float* beta, float* y_out, float* df_out,
                       void* stream) {
  if (!dlogits || !x || !gamma || !proj || !other_feat || !feat || !inv_norm || !dres)
    return set_error(RLCF_ERR_ARG, "head_bwd_sets: null pointer");
  return head_bwd(dlogits, x, row_idx, row_stride, gamma, param_stride, proj, other_feat, logit_scale, feat, inv_norm,
                  n_sets, seqs_per_set, d, E, K, eps, dres, partials, n_slots, p_total, p_off, dl_set_stride,
                  dl_seq_stride, dl_k_stride, other_set_stride, beta, y_out, df_out, proj_stride, S(stream));
}

int rlcf_embed_prompts(const int64_t* tokens, const float* tok_emb, const float* pos, const float* ctx,
                       int64_t ctx_stride, int n_ctx, int n_sets, int n_cls, int L, int d, float* x, void* stream) {
  if (!tokens || !tok_emb || !pos || !ctx || !x) return set_error(RLCF_ERR_ARG, "embed_prompts: null pointer");
  return embed_prompts(reinterpret_cast<const long long*>(tokens), tok_emb, pos, ctx, ctx_stride, n_ctx, n_sets, n_cls,
                       L, d, x, S(stream));
}

int rlcf_pair_logits(const float* img_feat, const float* txt_feat, int64_t txt_set_stride, int n_sets, int S_, int C,
                     int E, float logit_scale, float* logits, void* stream) {
  if (!img_feat || !txt_feat || !logits) return set_error(RLCF_ERR_ARG, "pair_logits: null pointer");
  return pair_logits(img_feat, txt_feat, txt_set_stride, n_sets, S_, C, E, logit_scale, logits, S(stream));
}

int rlcf_embed_prompts_map(const int64_t* tokens, const float* tok_emb, const float* pos, const float* vec,
                           int64_t vec_stride, const int32_t* src_map, int n_sets, int n_cls, int L, int d, float* x,
                           void* stream) {
  if (!tokens || !tok_emb || !pos || !vec || !src_map || !x)
    return set_error(RLCF_ERR_ARG, "embed_prompts_map: null pointer");
  return embed_prompts_map(reinterpret_cast<const long long*>(tokens), tok_emb, pos, vec, vec_stride, src_map, n_sets,
                           n_cls, L, d, x, S(stream));
}

int rlcf_vec_grad_map(const float* dx, const int32_t* ctx_pos, const int32_t* cls_pos, int n_sets, int n_cls, int L,
                      int n_ctx, int d, float* dvec, void* stream) {
  if (!dx || !ctx_pos || !dvec) return set_error(RLCF_ERR_ARG, "vec_grad_map: null pointer");
  return vec_grad_map(dx, ctx_pos, cls_pos, n_sets, n_cls, L, n_ctx, d, dvec, S(stream));
}

int rlcf_ctx_grad(const float* dx, int n_sets, int n_cls, int L, int n_ctx, int d, float* dctx, void* stream) {
  if (!dx || !dctx) return set_error(RLCF_ERR_ARG, "ctx_grad: null pointer");
  return ctx_grad(dx, n_sets, n_cls, L, n_ctx, d, dctx, S(stream));
}

int rlcf_adamw_step(float* params, float* m, float* v, const float* partials, int n_sets, int n_slots,
                    int64_t p_total, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    float loss_scale, float* grad_out, void* stream) {
  if (!params || !m || !v || !partials) return set_error(RLCF_ERR_ARG, "adamw_step: null pointer");
  return adamw_step(params, m, v, partials, n_sets, n_slots, p_total, lr, beta1, beta2, eps, weight_decay, step,
                    loss_scale, grad_out, nullptr, 0, 0, S(stream));
}

int rlcf_adamw_step_from(float* params, float* m, float* v, const float* grads, int n_sets, int n_slots,
                         int64_t p_total, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                         float loss_scale, const float* params_in, int64_t params_in_stride, int fresh_state,
                         void* stream) {
  if (!params || !m || !v || !grads || !params_in) return set_error(RLCF_ERR_ARG, "adamw_step_from: null pointer");
  return adamw_step(params, m, v, grads, n_sets, n_slots, p_total, lr, beta1, beta2, eps, weight_decay, step,
                    loss_scale, nullptr, params_in, params_in_stride, fresh_state, S(stream));
}

int rlcf_transpose_blocks_f16(const void* in, int in_is_f32, int n_sets, int rows_per_set, int rows_pad, int cols,
                              int skip_first, int64_t in_set_stride_rows, void* out, int64_t ld_out, void* stream) {
  if (!in || !out) return set_error(RLCF_ERR_ARG, "transpose_blocks: null pointer");
  return transpose_blocks(in, in_is_f32, n_sets, rows_per_set, rows_pad, cols, skip_first, in_set_stride_rows, H(out),
                          ld_out, S(stream));
}

int rlcf_colsum_f16(const void* in, int n_sets, int rows_per_set, int cols, float* out, int64_t out_stride,
                    void* stream) {
  if (!in || !out) return set_error(RLCF_ERR_ARG, "colsum: null pointer");
  return colsum_f16(CH(in), n_sets, rows_per_set, cols, out, out_stride, S(stream));
}

int rlcf_seq_sum(const float* dx, int n_sets, int S_, int L, int d, float* out, int64_t out_stride, void* stream) {
  if (!dx || !out) return set_error(RLCF_ERR_ARG, "seq_sum: null pointer");
  return seq_sum(dx, n_sets, S_, L, d, out, out_stride, S(stream));
}

int rlcf_outer_sum(const float* y, const float* df, int n_sets, int S_, int d, int E, float* out, int64_t out_stride,
                   void* stream) {
  if (!y || !df || !out) return set_error(RLCF_ERR_ARG, "outer_sum: null pointer");
  return outer_sum(y, df, n_sets, S_, d, E, out, out_stride, S(stream));
}

int rlcf_reset_params(const float* init, float* params, float* m, float* v, int n_sets, int64_t p_total,
                      void* stream) {
  if (!init || !params) return set_error(RLCF_ERR_ARG, "reset_params: null pointer");
  return reset_params(init, params, m, v, n_sets, p_total, S(stream));
}

int rlcf_cast_f16(const float* in, int64_t rows, int64_t cols, int64_t ld_in, void* out, int64_t ld_out,
                  void* stream) {
  if (!in || !out) return set_error(RLCF_ERR_ARG, "cast_f16: null pointer");
  return cast_f16(in, rows, cols, ld_in, H(out), ld_out, S(stream));
}

int rlcf_gather_seqs(const void* src, const int32_t* idx, void* dst, int64_t seq_bytes, int64_t src_layer_bytes,
                     int64_t dst_layer_bytes, int n_layers, int n, void* stream) {
  if (!src || !idx || !dst) return set_error(RLCF_ERR_ARG, "gather_seqs: null pointer");
  return gather_seqs(src, idx, dst, seq_bytes, src_layer_bytes, dst_layer_bytes, n_layers, n, S(stream));
}

int rlcf_transpose_cast_f16(const float* in, int rows, int cols, void* out, void* stream) {
  if (!in || !out) return set_error(RLCF_ERR_ARG, "transpose_cast_f16: null pointer");
  return transpose_cast_f16(in, rows, cols, H(out), 1, 0, 0, S(stream));
}

int rlcf_transpose_cast_f16_sets(const float* in, int rows, int cols, int n_sets, int64_t in_stride, void* out,
                                 int64_t out_stride, void* stream) {
  if (!in || !out) return set_error(RLCF_ERR_ARG, "transpose_cast_f16_sets: null pointer");
  return transpose_cast_f16(in, rows, cols, H(out), n_sets, in_stride, out_stride, S(stream));
}

int rlcf_retrieval_loss(const float* logits, int64_t ld, const float* reward_query, const float* reward_gallery,
                        int n_query, int K, int C, int Er, float clipscore_weight, int reward_process, int amplify,
                        float loss_scale, float* dlogits, int32_t* topk_idx, float* scores, float* rewards,
                        float* loss, void* stream) {
  if (!logits || !reward_query || !reward_gallery || !dlogits)
    return set_error(RLCF_ERR_ARG, "retrieval_loss: null pointer");
  return retrieval_loss(logits, ld, reward_query, reward_gallery, n_query, K, C, Er, clipscore_weight, reward_process,
                        amplify, loss_scale, dlogits, topk_idx, scores, rewards, loss, S(stream));
}

int rlcf_dfeat_partial(const float* dlogits, const float* gallery, int n_query, int C, int E, int n_chunks,
                       float* partial, void* stream) {
  if (!dlogits || !gallery || !partial) return set_error(RLCF_ERR_ARG, "dfeat_partial: null pointer");
  return dfeat_partial(dlogits, gallery, n_query, C, E, n_chunks, partial, S(stream));
}

int rlcf_rowdot(const float* a, const float* b, int n_rows, int C, float scale, float* out, int64_t out_stride,
                void* stream) {
  if (!a || !b || !out) return set_error(RLCF_ERR_ARG, "rowdot: null pointer");
  return rowdot(a, b, n_rows, C, scale, out, out_stride, S(stream));
}

int rlcf_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, int M, int N, int K, int epilogue,
                  const float* bias, const float* resid, float* out, int64_t ldo, void* stream) {
  if (!A || !W || !out) return set_error(RLCF_ERR_ARG, "gemm_f32: null pointer");
  return gemm_f32(A, lda, W, ldw, M, N, K, epilogue, bias, resid, out, ldo, S(stream));
}

int rlcf_attention_f32(const float* qkv, int n_seq, int L, int heads, int causal, float* out, void* stream) {
  if (!qkv || !out) return set_error(RLCF_ERR_ARG, "attention_f32: null pointer");
  return attention_f32(qkv, n_seq, L, heads, causal, out, S(stream));
}

int rlcf_accuracy_count(const float* logits, const int64_t* target, int n, int C, int64_t* hits, void* stream) {
  if (!logits || !target || !hits) return set_error(RLCF_ERR_ARG, "accuracy_count: null pointer");
  return accuracy_count(logits, reinterpret_cast<const long long*>(target), n, C, reinterpret_cast<long long*>(hits),
                        S(stream));
}

int rlcf_add_rows(const float* a, int64_t a_stride, const float* b, int64_t b_stride, int n_sets, int64_t n, float* x,
                  void* stream) {
  if (!a || !b || !x) return set_error(RLCF_ERR_ARG, "add_rows: null pointer");
  return add_rows(a, a_stride, b, b_stride, n_sets, n, x, S(stream));
}

int rlcf_scale_rows_exp(const float* in, const float* ls, int64_t ls_stride, int n_rows, int C, float* out,
                        void* stream) {
  if (!in || !ls || !out) return set_error(RLCF_ERR_ARG, "scale_rows_exp: null pointer");
  return scale_rows_exp(in, ls, ls_stride, n_rows, C, out, S(stream));
}

int rlcf_tied_rows_grad(const float* dx, const int64_t* tokens, int n_sets, int L, int d, float* g_tok, float* g_pos,
                        int64_t out_stride, void* stream) {
  if (!dx || !tokens || !g_tok || !g_pos) return set_error(RLCF_ERR_ARG, "tied_rows_grad: null pointer");
  return tied_rows_grad(dx, reinterpret_cast<const long long*>(tokens), n_sets, L, d, g_tok, g_pos, out_stride,
                        S(stream));
}

int rlcf_adamw_full(float* params, float* m, float* v, const float* grads, int n_sets, int64_t p_total, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, float loss_scale,
                    const float* params_in, int64_t params_in_stride, int fresh_state, void* w16, int64_t w16_stride,
                    int64_t n16, int64_t set_stride, void* stream) {
  if (!params || !m || !v || !grads || !params_in) return set_error(RLCF_ERR_ARG, "adamw_full: null pointer");
  return adamw_full(params, m, v, grads, n_sets, p_total, lr, beta1, beta2, eps, weight_decay, step, loss_scale,
                    params_in, params_in_stride, fresh_state, H(w16), w16_stride, n16, set_stride, S(stream));
}

int rlcf_transpose_f16_sets(const void* in, int rows, int cols, void* out, int n_sets, int64_t set_stride,
                            void* stream) {
  if (!in || !out) return set_error(RLCF_ERR_ARG, "transpose_f16_sets: null pointer");
  return transpose_f16(CH(in), rows, cols, H(out), n_sets, set_stride, S(stream));
}

int rlcf_resample_u8(const uint8_t* src, int H_, int W, int n_views, const int32_t* hdr, const int32_t* hb,
                     const int32_t* hk, int ks_h, const int32_t* vb, const int32_t* vk, int ks_v, int out_h, int out_w,
                     uint8_t* tmp, int tmp_rows, uint8_t* out, void* stream) {
  if (!src || !hdr || !hb || !hk || !vb || !vk || !tmp || !out) return set_error(RLCF_ERR_ARG, "resample_u8: null pointer");
  return resample_u8(src, H_, W, n_views, hdr, hb, hk, ks_h, vb, vk, ks_v, out_h, out_w, tmp, tmp_rows, out, S(stream));
}

int rlcf_augmix_views(const uint8_t* x_orig, int n_views, const int32_t* vflag, const float* wts, const float* omm,
                      const int32_t* n_ops, const int32_t* ops, const double* mats, float mean0, float mean1,
                      float mean2, float std0, float std1, float std2, float* out, void* stream) {
  if (!x_orig || !vflag || !wts || !omm || !n_ops || !ops || !mats || !out)
    return set_error(RLCF_ERR_ARG, "augmix_views: null pointer");
  const float mean[3] = {mean0, mean1, mean2}, stdv[3] = {std0, std1, std2};
  return augmix_views(x_orig, n_views, vflag, wts, omm, n_ops, ops, mats, mean, stdv, out, S(stream));
}

int rlcf_transpose_blocks_colsum(const void* in, int n_sets, int rows_per_set, int rows_pad, int cols, int skip_first,
                                 int64_t in_set_stride_rows, void* out, int64_t ld_out, float* colsum,
                                 int64_t colsum_stride, void* stream) {
  if (!in || !out) return set_error(RLCF_ERR_ARG, "transpose_blocks_colsum: null pointer");
  return transpose_blocks_colsum(CH(in), n_sets, rows_per_set, rows_pad, cols, skip_first, in_set_stride_rows, H(out),
                                 ld_out, colsum, colsum_stride, S(stream));
}

int rlcf_resample_taps(const int32_t* geom, int n_views, int out, int ks_h, int ks_v, int32_t* hdr, int32_t* hb,
                       int32_t* hk, int32_t* vb, int32_t* vk, void* stream) {
  if (!geom || !hdr || !hb || !hk || !vb || !vk) return set_error(RLCF_ERR_ARG, "resample_taps: null pointer");
  return resample_taps(geom, n_views, out, ks_h, ks_v, hdr, hb, hk, vb, vk, S(stream));
}

int rlcf_gemm_wgrad_adamw(const void* A, int lda, int64_t a_group_stride, const void* B, int ldb, int64_t b_group_stride,
                          int groups, int n_out, int n_in, int K, float* params, float* m, float* v, int ldp,
                          int64_t param_group_stride, const float* params_in, int64_t params_in_gs, int fresh_state,
                          void* w16, int64_t w16_group_stride, float lr, float beta1, float beta2, float eps,
                          float weight_decay, int step, float loss_scale, void* stream) {
  if (!A || !B || !params || !m || !v || !params_in) return set_error(RLCF_ERR_ARG, "gemm_wgrad_adamw: null pointer");
  if (step < 1 || loss_scale <= 0.f) return set_error(RLCF_ERR_ARG, "gemm_wgrad_adamw: bad step / loss_scale");
  if (groups > 1 && (params_in_gs % 4 != 0 || w16_group_stride % 4 != 0))
    return set_error(RLCF_ERR_ARG, "gemm_wgrad_adamw: group strides must be multiples of 4 elements");
  AdamwEpi o;
  o.m = m; o.v = v; o.p_in = params_in; o.p_in_gs = params_in_gs; o.w16 = H(w16); o.w16_gs = w16_group_stride;
  o.fresh = fresh_state; o.lr = lr; o.b1 = beta1; o.b2 = beta2; o.eps = eps; o.wd = weight_decay;
  o.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  o.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
  return gemm_f16_grouped(CH(A), lda, a_group_stride, CH(B), ldb, b_group_stride, groups, n_out, n_in, K, EPI_ADAMW,
                          1.0f / loss_scale, nullptr, 0, nullptr, nullptr, nullptr, params, ldp, param_group_stride,
                          S(stream), &o);
}

int rlcf_reward_loss_multi(const float* logits, const int32_t* row_idx, int n_models, const float* reward_img0,
                           const float* reward_img1, const float* reward_img2, const float* reward_img3,
                           const float* reward_cls0, const float* reward_cls1, const float* reward_cls2,
                           const float* reward_cls3, int er0, int er1, int er2, int er3, float weight0, float weight1,
                           float weight2, float weight3, int n_img, int S_, int K, int C, float clipscore_weight,
                           int reward_process, int process_batch, int amplify, float loss_scale, float* dlogits,
                           int32_t* topk_idx, float* scores, float* rewards, float* loss, void* stream) {
  if (!logits || !dlogits) return set_error(RLCF_ERR_ARG, "reward_loss_multi: null pointer");
  const float* img[4] = {reward_img0, reward_img1, reward_img2, reward_img3};
  const float* cls[4] = {reward_cls0, reward_cls1, reward_cls2, reward_cls3};
  const int er[4] = {er0, er1, er2, er3};
  const float wt[4] = {weight0, weight1, weight2, weight3};
  return reward_loss_multi(logits, row_idx, n_models, img, cls, er, wt, n_img, S_, K, C, clipscore_weight,
                           reward_process, process_batch, amplify, loss_scale, dlogits, topk_idx, scores, rewards, loss,
                           S(stream));
}

int rlcf_bicubic_resize(const float* images, const int32_t* view_idx, int n_views, int C, int H_, int W, int oh, int ow,
                        float* out, void* stream) {
  if (!images || !out) return set_error(RLCF_ERR_ARG, "bicubic_resize: null pointer");
  return bicubic_resize(images, view_idx, n_views, C, H_, W, oh, ow, out, S(stream));
}

}  // extern "C"
