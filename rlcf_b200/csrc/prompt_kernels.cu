// Kernels specific to prompt tuning (TPT/tpt_cls_rl.py with ClipTestTimeTuning, TPT/clip/custom_clip.py:76-344):
// prompt assembly with per-image learnable context vectors, image x text logits with per-image text features,
// and the reduction of the text-tower input gradient onto the context vectors.
#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

// PromptLearner.forward with class_token_position == "end" (custom_clip.py:198-232) + TextEncoder's positional add
// (custom_clip.py:62-63): row (g, c, t) = (1 <= t <= n_ctx ? ctx[g][t-1] : token_embedding[tokens[c][t]]) + pos[t].
// `g` is the parameter set (test image); ctx_stride = 0 shares one context.
__global__ void embed_prompts_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb,
                                     const float* __restrict__ pos, const float* __restrict__ ctx,
                                     long long ctx_stride, int n_ctx, int n_cls, int L, int d4, long long total,
                                     float* __restrict__ x) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % d4);
    const long long row = i / d4;
    const int t = static_cast<int>(row % L);
    const long long seq = row / L;
    const int c = static_cast<int>(seq % n_cls);
    const long long g = seq / n_cls;
    float4 a;
    if (t >= 1 && t <= n_ctx) {
      a = reinterpret_cast<const float4*>(ctx + g * ctx_stride)[static_cast<long long>(t - 1) * d4 + c4];
    } else {
      const long long tok = tokens[static_cast<long long>(c) * L + t];
      a = __ldg(reinterpret_cast<const float4*>(emb) + tok * d4 + c4);
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos) + static_cast<long long>(t) * d4 + c4);
    reinterpret_cast<float4*>(x)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

int embed_prompts(const long long* tokens, const float* tok_emb, const float* pos, const float* ctx,
                  long long ctx_stride, int n_ctx, int n_sets, int n_cls, int L, int d, float* x,
                  cudaStream_t stream) {
  if (n_sets <= 0 || n_cls <= 0 || L <= 0 || d % 4 || n_ctx < 0 || n_ctx + 1 >= L)
    return set_error(RLCF_ERR_ARG, "embed_prompts: bad shape");
  const long long total = static_cast<long long>(n_sets) * n_cls * L * (d / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  embed_prompts_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(tokens, tok_emb, pos, ctx, ctx_stride, n_ctx,
                                                                     n_cls, L, d / 4, total, x);
  RLCF_CHECK_LAUNCH("embed_prompts");
  return 0;
}

// General prompt layout (custom_clip.py:233-289: class token in the middle / at the front, custom_clip.py:209-221:
// learnable class tokens): src_map[c][t] >= 0 takes the frozen embedding of tokens[c][src_map[c][t]], src_map[c][t] < 0
// takes learnable vector v = -1 - src_map[c][t] of the parameter set, vec[g][v][:] (context vectors first, then -- with
// learned class tokens -- one vector per class).  Row (g, c, t) = source + pos[t].
__global__ void embed_prompts_map_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb,
                                         const float* __restrict__ pos, const float* __restrict__ vec,
                                         long long vec_stride, const int* __restrict__ src_map, int n_cls, int L, int d4,
                                         long long total, float* __restrict__ x) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % d4);
    const long long row = i / d4;
    const int t = static_cast<int>(row % L);
    const long long seq = row / L;
    const int c = static_cast<int>(seq % n_cls);
    const long long g = seq / n_cls;
    const int m = src_map[c * L + t];
    float4 a;
    if (m < 0) {
      a = reinterpret_cast<const float4*>(vec + g * vec_stride)[static_cast<long long>(-1 - m) * d4 + c4];
    } else {
      const long long tok = tokens[static_cast<long long>(c) * L + m];
      a = __ldg(reinterpret_cast<const float4*>(emb) + tok * d4 + c4);
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos) + static_cast<long long>(t) * d4 + c4);
    reinterpret_cast<float4*>(x)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

int embed_prompts_map(const long long* tokens, const float* tok_emb, const float* pos, const float* vec,
                      long long vec_stride, const int* src_map, int n_sets, int n_cls, int L, int d, float* x,
                      cudaStream_t stream) {
  if (n_sets <= 0 || n_cls <= 0 || L <= 0 || d % 4) return set_error(RLCF_ERR_ARG, "embed_prompts_map: bad shape");
  const long long total = static_cast<long long>(n_sets) * n_cls * L * (d / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  embed_prompts_map_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(tokens, tok_emb, pos, vec, vec_stride, src_map,
                                                                         n_cls, L, d / 4, total, x);
  RLCF_CHECK_LAUNCH("embed_prompts_map");
  return 0;
}

// Gradient of the learnable vectors under a general layout: context vector v (< n_ctx) appears once in every class, at
// token position ctx_pos[c][v]: d vec[g][v] = sum_c dx[g, c, ctx_pos[c][v]] (fixed order: deterministic); the class
// vector of class c (learned class tokens, cls_pos != NULL) sits at cls_pos[c]: d vec[g][n_ctx + c] = dx[g, c, cls_pos[c]].
__global__ void vec_grad_map_kernel(const float* __restrict__ dx, const int* __restrict__ ctx_pos,
                                    const int* __restrict__ cls_pos, int n_cls, int L, int n_ctx, int n_vec, int d4,
                                    long long total, float* __restrict__ dvec) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % d4);
    const int v = static_cast<int>((i / d4) % n_vec);
    const long long g = i / (static_cast<long long>(d4) * n_vec);
    const float4* base = reinterpret_cast<const float4*>(dx) + (g * n_cls) * L * d4 + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < n_ctx) {
      for (int c = 0; c < n_cls; ++c) {
        const float4 u = base[(static_cast<long long>(c) * L + ctx_pos[c * n_ctx + v]) * d4];
        acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
      }
    } else {
      const int c = v - n_ctx;
      acc = base[(static_cast<long long>(c) * L + cls_pos[c]) * d4];
    }
    reinterpret_cast<float4*>(dvec)[i] = acc;
  }
}

int vec_grad_map(const float* dx, const int* ctx_pos, const int* cls_pos, int n_sets, int n_cls, int L, int n_ctx,
                 int d, float* dvec, cudaStream_t stream) {
  if (n_sets <= 0 || n_cls <= 0 || n_ctx <= 0 || d % 4) return set_error(RLCF_ERR_ARG, "vec_grad_map: bad shape");
  const int n_vec = n_ctx + (cls_pos != nullptr ? n_cls : 0);
  const long long total = static_cast<long long>(n_sets) * n_vec * (d / 4);
  long long blocks = (total + 127) / 128;
  if (blocks > 148 * 32) blocks = 148 * 32;
  vec_grad_map_kernel<<<static_cast<int>(blocks), 128, 0, stream>>>(dx, ctx_pos, cls_pos, n_cls, L, n_ctx, n_vec, d / 4,
                                                                    total, dvec);
  RLCF_CHECK_LAUNCH("vec_grad_map");
  return 0;
}

// logits[g, s, c] = logit_scale * <img[g, s, :], txt[g, c, :]>   (ClipTestTimeTuning.inference, custom_clip.py:325-335)
// txt_stride = 0 shares one set of text features across images.  One warp per (g, s, c).
__global__ void __launch_bounds__(256)
pair_logits_kernel(const float* __restrict__ img, const float* __restrict__ txt, long long txt_stride, int S, int C,
                   int E, float scale, long long total, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  for (long long w = blockIdx.x * 8LL + (threadIdx.x >> 5); w < total; w += static_cast<long long>(gridDim.x) * 8) {
    const int c = static_cast<int>(w % C);
    const long long gs = w / C;
    const long long g = gs / S;
    const float4* a = reinterpret_cast<const float4*>(img + gs * E);
    const float4* b = reinterpret_cast<const float4*>(txt + g * txt_stride + static_cast<long long>(c) * E);
    float acc = 0.f;
    for (int j = lane; j < (E >> 2); j += 32) {
      const float4 u = __ldg(a + j), v = __ldg(b + j);
      acc = fmaf(u.x, v.x, fmaf(u.y, v.y, fmaf(u.z, v.z, fmaf(u.w, v.w, acc))));
    }
    acc = warp_sum(acc);
    if (lane == 0) logits[w] = scale * acc;
  }
}

int pair_logits(const float* img, const float* txt, long long txt_stride, int n_sets, int S, int C, int E, float scale,
                float* logits, cudaStream_t stream) {
  if (n_sets <= 0 || S <= 0 || C <= 0 || E <= 0 || E % 4) return set_error(RLCF_ERR_ARG, "pair_logits: bad shape");
  const long long total = static_cast<long long>(n_sets) * S * C;
  long long blocks = (total + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pair_logits_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(img, txt, txt_stride, S, C, E, scale, total, logits);
  RLCF_CHECK_LAUNCH("pair_logits");
  return 0;
}

// d ctx[g][i][:] = sum_c dx[(g * n_cls + c) * L + 1 + i][:]  -- the positional embedding and the frozen token
// embeddings take no gradient; the sum over classes runs in a fixed order (deterministic).
__global__ void ctx_grad_kernel(const float* __restrict__ dx, int n_cls, int L, int n_ctx, int d4, long long total,
                                float* __restrict__ dctx) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % d4);
    const int t = static_cast<int>((i / d4) % n_ctx);
    const long long g = i / (static_cast<long long>(d4) * n_ctx);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* base = reinterpret_cast<const float4*>(dx) + ((g * n_cls) * L + 1 + t) * d4 + c4;
    for (int c = 0; c < n_cls; ++c) {
      const float4 v = base[static_cast<long long>(c) * L * d4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dctx)[i] = acc;
  }
}

int ctx_grad(const float* dx, int n_sets, int n_cls, int L, int n_ctx, int d, float* dctx, cudaStream_t stream) {
  if (n_sets <= 0 || n_cls <= 0 || n_ctx <= 0 || d % 4) return set_error(RLCF_ERR_ARG, "ctx_grad: bad shape");
  const long long total = static_cast<long long>(n_sets) * n_ctx * (d / 4);
  long long blocks = (total + 127) / 128;
  ctx_grad_kernel<<<static_cast<int>(blocks), 128, 0, stream>>>(dx, n_cls, L, n_ctx, d / 4, total, dctx);
  RLCF_CHECK_LAUNCH("ctx_grad");
  return 0;
}

}  // namespace rlcf
