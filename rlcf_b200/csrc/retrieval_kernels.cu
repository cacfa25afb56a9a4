// Kernels specific to retrieval TTA (retrieval/clip_ret_policy.py:76-137): one query against a gallery of thousands
// of candidates, so the "class" axis is 5 000 - 25 000 wide and K = 12 / 20 candidates are sampled per step.
//   retrieval_loss_kernel   top-K of one score row, CLIPScore of the sampled pairs, rewards, reward-weighted CE and
//                           its gradient w.r.t. the row (clip_ret_policy.py:88-98 / 121-131)
//   dfeat_partial_kernel    d(query feature) = dlogits @ gallery, split over gallery chunks (deterministic two-stage
//                           reduction: the chunks are summed by head_bwd in a fixed order)
//   logit_scale_grad_kernel d(logit_scale) = sum_c dlogits[c] * logits[c]   (text->image tunes logit_scale too)
#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

constexpr int kRetThreads = 512;
constexpr int kRetMaxK = 32;

__device__ __forceinline__ float block_reduce_sum_ret(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < kRetThreads / 32; ++i) r += scratch[i];   // fixed order: every thread gets the same bits
  return r;
}

__device__ __forceinline__ float block_reduce_max_ret(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = -INFINITY;
  for (int i = 0; i < kRetThreads / 32; ++i) r = fmaxf(r, scratch[i]);
  return r;
}

// One block per query.  logits row q: [C] (stride ld).  Sampled candidates = the K largest scores, ties broken by
// the lower index (torch.topk on CUDA/CPU returns sorted values; equal values are not expected on real features).
__global__ void __launch_bounds__(kRetThreads)
retrieval_loss_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ r_query,
                      const float* __restrict__ r_gallery, int K, int C, int Er, float w, int reward_process,
                      int amplify, float loss_scale, float* __restrict__ dlogits, int32_t* __restrict__ topk_idx,
                      float* __restrict__ scores_out, float* __restrict__ rewards_out, float* __restrict__ loss_out) {
  __shared__ float scratch[kRetThreads / 32];
  __shared__ int scratch_i[kRetThreads / 32];
  __shared__ float sc[kRetMaxK], ce[kRetMaxK];
  __shared__ int idx[kRetMaxK];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* r = logits + q * ld;
  float mx = -INFINITY;
  for (int c = tid; c < C; c += kRetThreads) mx = fmaxf(mx, r[c]);
  mx = block_reduce_max_ret(mx, scratch);
  float se = 0.f;
  for (int c = tid; c < C; c += kRetThreads) se += expf(r[c] - mx);
  const float lse = logf(block_reduce_sum_ret(se, scratch));
  // top-K by K rounds of block-wide argmax; already chosen candidates are skipped
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = tid; c < C; c += kRetThreads) {
      bool taken = false;
      for (int j = 0; j < k; ++j) taken |= (idx[j] == c);
      const float val = r[c];
      if (!taken && (val > bv || (val == bv && c < bi))) { bv = val; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { scratch[warp] = bv; scratch_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < kRetThreads / 32; ++i)
        if (scratch[i] > bv || (scratch[i] == bv && scratch_i[i] < bi)) { bv = scratch[i]; bi = scratch_i[i]; }
      idx[k] = bi;
      ce[k] = (mx + lse) - bv;                 // cross-entropy of the sampled candidate
    }
    __syncthreads();
  }
  // CLIPScore = max(0, w * <reward gallery feature, reward query feature>)   (retrieval/clip_reward.py:143-168)
  for (int k = warp; k < K; k += kRetThreads / 32) {
    const float* t = r_gallery + static_cast<size_t>(idx[k]) * Er;
    const float* f = r_query + static_cast<size_t>(q) * Er;
    float acc = 0.f;
    for (int j = lane; j < Er; j += 32) acc = fmaf(__ldg(t + j), __ldg(f + j), acc);
    acc = warp_sum(acc);
    if (lane == 0) sc[k] = fmaxf(w * acc, 0.f);
  }
  __syncthreads();
  if (scores_out && tid < K) scores_out[static_cast<size_t>(q) * K + tid] = sc[tid];
  __syncthreads();
  // rewards_post_process over the K samples of this query (unbiased std, as torch.std)
  if (tid == 0 && reward_process && K > 1) {
    float m = 0.f;
    for (int k = 0; k < K; ++k) m += sc[k];
    m /= K;
    float sd = 1.f;
    if (amplify) {
      float v = 0.f;
      for (int k = 0; k < K; ++k) v += (sc[k] - m) * (sc[k] - m);
      sd = sqrtf(v / (K - 1)) + 1e-5f;
    }
    for (int k = 0; k < K; ++k) sc[k] = (sc[k] - m) / sd;
  }
  __syncthreads();
  const float inv_n = 1.f / K;
  if (tid == 0 && loss_out) {
    float l = 0.f;
    for (int k = 0; k < K; ++k) l += sc[k] * ce[k];
    loss_out[q] = l * inv_n;
  }
  if (tid < K) {
    if (rewards_out) rewards_out[static_cast<size_t>(q) * K + tid] = sc[tid];
    if (topk_idx) topk_idx[static_cast<size_t>(q) * K + tid] = idx[tid];
  }
  // dL/dlogit[c] = (1/K) * sum_k r[k] * (softmax[c] - [c == idx[k]])
  float rs = 0.f;
  for (int k = 0; k < K; ++k) rs += sc[k];
  const float off = mx + lse;
  float* o = dlogits + static_cast<size_t>(q) * C;
  for (int c = tid; c < C; c += kRetThreads) {
    float g = rs * expf(r[c] - off);
    for (int k = 0; k < K; ++k) g -= (idx[k] == c) ? sc[k] : 0.f;
    o[c] = g * inv_n * loss_scale;
  }
}

int retrieval_loss(const float* logits, long long ld, const float* r_query, const float* r_gallery, int n_query, int K,
                   int C, int Er, float w, int reward_process, int amplify, float loss_scale, float* dlogits,
                   int32_t* topk_idx, float* scores, float* rewards, float* loss, cudaStream_t stream) {
  if (n_query <= 0 || K <= 0 || K > kRetMaxK || K > C || Er <= 0 || ld < C)
    return set_error(RLCF_ERR_ARG, "retrieval_loss: bad shape (K must be 1..%d and <= gallery size)", kRetMaxK);
  retrieval_loss_kernel<<<n_query, kRetThreads, 0, stream>>>(logits, ld, r_query, r_gallery, K, C, Er, w,
                                                             reward_process, amplify, loss_scale, dlogits, topk_idx,
                                                             scores, rewards, loss);
  RLCF_CHECK_LAUNCH("retrieval_loss");
  return 0;
}

// partial[q, chunk, j] = sum_{c in chunk} dl[q, c] * gallery[c, j].  grid = (n_chunks, n_query), E/4 float4 columns
// spread over the threads, the chunk's rows over the remaining thread groups (summed through shared memory).
constexpr int kDfThreads = 256;
__global__ void __launch_bounds__(kDfThreads)
dfeat_partial_kernel(const float* __restrict__ dl, const float* __restrict__ gallery, int C, int E4, int chunk,
                     float* __restrict__ partial) {
  extern __shared__ float4 red[];   // [groups][E4]
  const int q = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x;
  const int groups = kDfThreads / E4 > 0 ? kDfThreads / E4 : 1;
  const int c0 = ch * chunk, c1 = min(C, c0 + chunk);
  const float* d = dl + static_cast<size_t>(q) * C;
  for (int j0 = 0; j0 < E4; j0 += kDfThreads) {
    const int j = j0 + (groups > 1 ? tid % E4 : tid);
    const int grp = groups > 1 ? tid / E4 : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < E4 && grp < groups) {
      for (int c = c0 + grp; c < c1; c += groups) {
        const float s = d[c];
        const float4 g = __ldg(reinterpret_cast<const float4*>(gallery) + static_cast<size_t>(c) * E4 + j);
        acc.x = fmaf(s, g.x, acc.x); acc.y = fmaf(s, g.y, acc.y); acc.z = fmaf(s, g.z, acc.z); acc.w = fmaf(s, g.w, acc.w);
      }
    }
    if (groups > 1) {
      if (grp < groups) red[grp * E4 + j] = acc;
      __syncthreads();
      if (tid < E4) {
        float4 t = red[tid];
        for (int g = 1; g < groups; ++g) {
          const float4 u = red[g * E4 + tid];
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        reinterpret_cast<float4*>(partial)[(static_cast<size_t>(q) * gridDim.x + ch) * E4 + tid] = t;
      }
      __syncthreads();
    } else if (j < E4) {
      reinterpret_cast<float4*>(partial)[(static_cast<size_t>(q) * gridDim.x + ch) * E4 + j] = acc;
    }
  }
}

int dfeat_partial(const float* dl, const float* gallery, int n_query, int C, int E, int n_chunks, float* partial,
                  cudaStream_t stream) {
  if (n_query <= 0 || C <= 0 || E <= 0 || E % 4 || n_chunks <= 0 || n_query > 65535)
    return set_error(RLCF_ERR_ARG, "dfeat_partial: bad shape");
  const int E4 = E / 4;
  const int chunk = (C + n_chunks - 1) / n_chunks;
  const int groups = kDfThreads / E4 > 0 ? kDfThreads / E4 : 1;
  const size_t smem = groups > 1 ? static_cast<size_t>(groups) * E4 * sizeof(float4) : 0;
  dim3 grid(n_chunks, n_query);
  dfeat_partial_kernel<<<grid, kDfThreads, smem, stream>>>(dl, gallery, C, E4, chunk, partial);
  RLCF_CHECK_LAUNCH("dfeat_partial");
  return 0;
}

// out[q] = scale * sum_c a[q, c] * b[q, c]   (d logit_scale = sum_c dlogits * logits; logits = exp(ls) * cos)
__global__ void __launch_bounds__(kRetThreads)
rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, float scale, float* __restrict__ out,
              long long out_stride) {
  __shared__ float scratch[kRetThreads / 32];
  const int q = blockIdx.x;
  float acc = 0.f;
  for (int c = threadIdx.x; c < C; c += kRetThreads)
    acc = fmaf(a[static_cast<size_t>(q) * C + c], b[static_cast<size_t>(q) * C + c], acc);
  acc = block_reduce_sum_ret(acc, scratch);
  if (threadIdx.x == 0) out[q * out_stride] = scale * acc;
}

int rowdot(const float* a, const float* b, int n_rows, int C, float scale, float* out, long long out_stride,
           cudaStream_t stream) {
  if (n_rows <= 0 || C <= 0) return set_error(RLCF_ERR_ARG, "rowdot: bad shape");
  rowdot_kernel<<<n_rows, kRetThreads, 0, stream>>>(a, b, C, scale, out, out_stride);
  RLCF_CHECK_LAUNCH("rowdot");
  return 0;
}

// ---- text->image: the caption's token-embedding rows, the positional embedding and logit_scale are tuned too
// (custom_models.py:144-152).  Each query keeps a private copy of the L embedding rows it uses; positions that hold
// the same token id are tied by giving every copy the SUM of their gradients (identical AdamW trajectories).

// x[g, i] = a[g*a_stride + i] + b[g*b_stride + i], i < n (n % 4 == 0): token rows + positional embedding.
__global__ void add_rows_kernel(const float* __restrict__ a, long long a_stride, const float* __restrict__ b,
                                long long b_stride, int n4, long long total, float* __restrict__ x) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long g = i / n4;
    const int j = static_cast<int>(i % n4);
    const float4 u = reinterpret_cast<const float4*>(a + g * a_stride)[j];
    const float4 v = reinterpret_cast<const float4*>(b + g * b_stride)[j];
    reinterpret_cast<float4*>(x)[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

int add_rows(const float* a, long long a_stride, const float* b, long long b_stride, int n_sets, long long n, float* x,
             cudaStream_t stream) {
  if (n_sets <= 0 || n <= 0 || n % 4 || a_stride % 4 || b_stride % 4) return set_error(RLCF_ERR_ARG, "add_rows: bad shape");
  const long long total = n_sets * (n / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_rows_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(a, a_stride, b, b_stride, static_cast<int>(n / 4), total, x);
  RLCF_CHECK_LAUNCH("add_rows");
  return 0;
}

// out[q, c] = in[q, c] * exp(ls[q*ls_stride])   (logits = logit_scale.exp() * cos, custom_models.py:70-73)
__global__ void scale_rows_exp_kernel(const float* __restrict__ in, const float* __restrict__ ls, long long ls_stride,
                                      int C, long long total, float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = in[i] * expf(ls[(i / C) * ls_stride]);
}

int scale_rows_exp(const float* in, const float* ls, long long ls_stride, int n_rows, int C, float* out,
                   cudaStream_t stream) {
  if (n_rows <= 0 || C <= 0) return set_error(RLCF_ERR_ARG, "scale_rows_exp: bad shape");
  const long long total = static_cast<long long>(n_rows) * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_rows_exp_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, ls, ls_stride, C, total, out);
  RLCF_CHECK_LAUNCH("scale_rows_exp");
  return 0;
}

// dx [n_sets*L, d] = gradient w.r.t. the text tower's input rows.  g_pos[g][t] = dx[g,t];
// g_tok[g][t] = sum over t' with tokens[g,t'] == tokens[g,t] of dx[g,t']  (ascending t': deterministic).
__global__ void tied_rows_grad_kernel(const float* __restrict__ dx, const long long* __restrict__ tokens, int L, int d4,
                                      float* __restrict__ g_tok, float* __restrict__ g_pos, long long out_stride) {
  const int g = blockIdx.y, t = blockIdx.x;
  const long long* tk = tokens + static_cast<long long>(g) * L;
  const long long mine = tk[t];
  const float4* base = reinterpret_cast<const float4*>(dx) + static_cast<long long>(g) * L * d4;
  for (int j = threadIdx.x; j < d4; j += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int u = 0; u < L; ++u) {
      if (tk[u] != mine) continue;
      const float4 v = base[static_cast<long long>(u) * d4 + j];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(g_tok + g * out_stride)[static_cast<long long>(t) * d4 + j] = acc;
    reinterpret_cast<float4*>(g_pos + g * out_stride)[static_cast<long long>(t) * d4 + j] =
        base[static_cast<long long>(t) * d4 + j];
  }
}

int tied_rows_grad(const float* dx, const long long* tokens, int n_sets, int L, int d, float* g_tok, float* g_pos,
                   long long out_stride, cudaStream_t stream) {
  if (n_sets <= 0 || n_sets > 65535 || L <= 0 || d % 4 || out_stride % 4)
    return set_error(RLCF_ERR_ARG, "tied_rows_grad: bad shape");
  dim3 grid(L, n_sets);
  tied_rows_grad_kernel<<<grid, 128, 0, stream>>>(dx, tokens, L, d / 4, g_tok, g_pos, out_stride);
  RLCF_CHECK_LAUNCH("tied_rows_grad");
  return 0;
}

}  // namespace rlcf
