// Bandwidth/latency-bound kernels of the RLCF hot path: im2col, token assembly + ln_pre, LayerNorm fwd/bwd,
// head (ln_post + proj + L2-norm + logits) fwd/bwd, entropy selection, top-K/CLIPScore/reward/CE-gradient,
// fused gradient-reduce + AdamW.  Warp-shuffle reductions, 128-bit global accesses where the layout allows.
// Reference lines are cited per kernel; see include/rlcf_b200.h for the ABI contract.
#include <cstdlib>

#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

// ------------------------------------------------------------------------------------------------ block reduce
template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* scratch /* >= kThreads/32 + 1 floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch from the previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += scratch[i];  // same order in every thread -> deterministic
  return t;
}

// ------------------------------------------------------------------------------------------------ im2col
// conv1 with stride == kernel (TPT/clip/model.py:224) is a GEMM over non-overlapping patches.
__global__ void im2col_kernel(const float* __restrict__ img, const int32_t* __restrict__ view_idx, int n_views, int C,
                              int H, int W, int p, int k_pad, __half* __restrict__ out) {
  const int gw = W / p, gh = H / p;
  const int k_real = C * p * p;
  const int chunks = k_pad / 8;
  const long long total = static_cast<long long>(n_views) * gh * gw * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(i % chunks);
    const long long row = i / chunks;
    const int px = static_cast<int>(row % gw);
    const int py = static_cast<int>((row / gw) % gh);
    const int v = static_cast<int>(row / (gw * gh));
    const int src = view_idx ? view_idx[v] : v;
    const float* base = img + static_cast<size_t>(src) * C * H * W;
    __align__(16) __half h[8];
    const int k0 = chunk * 8;
    if ((p % 8) == 0 && k0 + 8 <= k_real) {
      const int c = k0 / (p * p), rem = k0 % (p * p), ky = rem / p, kx = rem % p;
      const float4* s = reinterpret_cast<const float4*>(base + (static_cast<size_t>(c) * H + py * p + ky) * W + px * p + kx);
      const float4 a = __ldg(s), b = __ldg(s + 1);
      h[0] = __float2half_rn(a.x); h[1] = __float2half_rn(a.y); h[2] = __float2half_rn(a.z); h[3] = __float2half_rn(a.w);
      h[4] = __float2half_rn(b.x); h[5] = __float2half_rn(b.y); h[6] = __float2half_rn(b.z); h[7] = __float2half_rn(b.w);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        float val = 0.f;
        if (k < k_real) {
          const int c = k / (p * p), rem = k % (p * p), ky = rem / p, kx = rem % p;
          val = __ldg(base + (static_cast<size_t>(c) * H + py * p + ky) * W + px * p + kx);
        }
        h[e] = __float2half_rn(val);
      }
    }
    *reinterpret_cast<uint4*>(out + row * k_pad + k0) = *reinterpret_cast<const uint4*>(h);
  }
}

int im2col_f16(const float* images, const int32_t* view_idx, int n_views, int C, int H, int W, int patch, int k_pad,
               __half* out, cudaStream_t stream) {
  if (n_views <= 0) return set_error(RLCF_ERR_ARG, "im2col: n_views=%d", n_views);
  if (H % patch || W % patch) return set_error(RLCF_ERR_ARG, "im2col: %dx%d not divisible by patch %d", H, W, patch);
  if (k_pad % 8 || k_pad < C * patch * patch) return set_error(RLCF_ERR_ARG, "im2col: bad k_pad=%d", k_pad);
  if ((patch % 8) == 0 && (W % 4) != 0) return set_error(RLCF_ERR_ARG, "im2col: W must be a multiple of 4");
  const long long total = static_cast<long long>(n_views) * (H / patch) * (W / patch) * (k_pad / 8);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  im2col_kernel<<<static_cast<int>(blocks), threads, 0, stream>>>(images, view_idx, n_views, C, H, W, patch, k_pad, out);
  RLCF_CHECK_LAUNCH("im2col");
  return 0;
}

// ------------------------------------------------------------------------------------------------ LayerNorm fwd
// One warp per row, row kept in registers (d = 128*NV floats), two-pass mean / variance in fp32
// (TPT/clip/model.py:157-163: nn.LayerNorm on x.float(), eps = 1e-5).
template <int NV>
__device__ __forceinline__ void ln_row_stats(const float4 (&v)[NV], int d, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + b * b) + (c * c + e * e);
  }
  rstd = 1.0f / sqrtf(warp_sum(q) / d + eps);
}

template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
              const float* __restrict__ beta, long long pstride, int rows_per_set, int M, float eps,
              __half* __restrict__ out16, float* __restrict__ out32) {
  constexpr int d = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  float mean, rstd;
  ln_row_stats<NV>(v, d, eps, mean, rstd);
  const long long po = static_cast<long long>(row / rows_per_set) * pstride;
  const float4* g4 = reinterpret_cast<const float4*>(gamma + po);
  const float4* b4 = reinterpret_cast<const float4*>(beta + po);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    const size_t o = static_cast<size_t>(row) * d + (lane + 32 * i) * 4;
    if (out16) {
      __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      *reinterpret_cast<uint2*>(out16 + o) =
          make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
    if (out32) *reinterpret_cast<float4*>(out32 + o) = y;
  }
}

#define RLCF_DISPATCH_NV(d, CALL)                         \
  switch ((d) / 128) {                                    \
    case 1: { constexpr int NV = 1; CALL; } break;        \
    case 2: { constexpr int NV = 2; CALL; } break;        \
    case 3: { constexpr int NV = 3; CALL; } break;        \
    case 4: { constexpr int NV = 4; CALL; } break;        \
    case 5: { constexpr int NV = 5; CALL; } break;        \
    case 6: { constexpr int NV = 6; CALL; } break;        \
    case 7: { constexpr int NV = 7; CALL; } break;        \
    case 8: { constexpr int NV = 8; CALL; } break;        \
    default: return set_error(RLCF_ERR_ARG, "width %d unsupported (need multiple of 128, <= 1024)", (d)); \
  }

static int check_width(int d, const char* who) {
  if (d <= 0 || d % 128 != 0 || d > 1024)
    return set_error(RLCF_ERR_ARG, "%s: width %d unsupported (need multiple of 128, <= 1024)", who, d);
  return 0;
}

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, long long pstride,
                  int rows_per_set, int M, int d, float eps, __half* out16, float* out32, cudaStream_t stream) {
  if (int rc = check_width(d, "layernorm_fwd")) return rc;
  if (M <= 0 || rows_per_set <= 0 || ldx % 4) return set_error(RLCF_ERR_ARG, "layernorm_fwd: bad M/rows_per_set/ldx");
  const int blocks = (M + 7) / 8;
  RLCF_DISPATCH_NV(d, (ln_fwd_kernel<NV><<<blocks, 256, 0, stream>>>(x, ldx, gamma, beta, pstride, rows_per_set, M,
                                                                      eps, out16, out32)));
  RLCF_CHECK_LAUNCH("layernorm_fwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------ embed + ln_pre
// TPT/clip/model.py:225-229: cat(class_embedding, patches) + positional_embedding, then ln_pre.
template <int NV>
__global__ void __launch_bounds__(256)
embed_lnpre_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls, const float* __restrict__ pos,
                   const float* __restrict__ gamma, const float* __restrict__ beta, long long pstride,
                   int rows_per_set, int M, int L, float eps, float* __restrict__ x_pre, float* __restrict__ x,
                   long long embed_stride) {
  constexpr int d = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int v = row / L, t = row % L;
  const long long eo = static_cast<long long>(row / rows_per_set) * embed_stride;   // per-set class / positional rows
  const float4* src = t == 0 ? reinterpret_cast<const float4*>(cls + eo)
                             : reinterpret_cast<const float4*>(patch_out + (static_cast<size_t>(v) * (L - 1) + t - 1) * d);
  const float4* p4 = reinterpret_cast<const float4*>(pos + eo + static_cast<size_t>(t) * d);
  float4 val[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = src[lane + 32 * i], b = __ldg(p4 + lane + 32 * i);
    val[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
  if (x_pre) {
    float4* o = reinterpret_cast<float4*>(x_pre + static_cast<size_t>(row) * d);
#pragma unroll
    for (int i = 0; i < NV; ++i) o[lane + 32 * i] = val[i];
  }
  float mean, rstd;
  ln_row_stats<NV>(val, d, eps, mean, rstd);
  const long long po = static_cast<long long>(row / rows_per_set) * pstride;
  const float4* g4 = reinterpret_cast<const float4*>(gamma + po);
  const float4* b4 = reinterpret_cast<const float4*>(beta + po);
  float4* o = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * d);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    o[lane + 32 * i] = make_float4((val[i].x - mean) * rstd * g.x + b.x, (val[i].y - mean) * rstd * g.y + b.y,
                                   (val[i].z - mean) * rstd * g.z + b.z, (val[i].w - mean) * rstd * g.w + b.w);
  }
}

int embed_lnpre(const float* patch_out, const float* cls, const float* pos, const float* gamma, const float* beta,
                long long pstride, int rows_per_set, int n_views, int L, int d, float eps, float* x_pre, float* x,
                long long embed_stride, cudaStream_t stream) {
  if (int rc = check_width(d, "embed_lnpre")) return rc;
  if (embed_stride % 4) return set_error(RLCF_ERR_ARG, "embed_lnpre: embed_stride must be a multiple of 4 floats");
  if (n_views <= 0 || L < 2 || rows_per_set <= 0) return set_error(RLCF_ERR_ARG, "embed_lnpre: bad shape");
  const int M = n_views * L;
  const int blocks = (M + 7) / 8;
  RLCF_DISPATCH_NV(d, (embed_lnpre_kernel<NV><<<blocks, 256, 0, stream>>>(patch_out, cls, pos, gamma, beta, pstride,
                                                                           rows_per_set, M, L, eps, x_pre, x,
                                                                           embed_stride)));
  RLCF_CHECK_LAUNCH("embed_lnpre");
  return 0;
}

// TPT/clip/model.py:343-345: token_embedding(text) + positional_embedding
__global__ void embed_text_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb,
                                  const float* __restrict__ pos, int n_rows, int L, int d4, float* __restrict__ x) {
  const long long total = static_cast<long long>(n_rows) * d4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const long long row = i / d4;
    const int t = static_cast<int>(row % L);
    const long long tok = tokens[row];
    const float4 a = __ldg(reinterpret_cast<const float4*>(emb) + tok * d4 + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos) + static_cast<long long>(t) * d4 + c);
    reinterpret_cast<float4*>(x)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

int embed_text(const long long* tokens, const float* tok_emb, const float* pos, int n_seq, int L, int d, float* x,
               cudaStream_t stream) {
  if (n_seq <= 0 || L <= 0 || d % 4) return set_error(RLCF_ERR_ARG, "embed_text: bad shape");
  const long long total = static_cast<long long>(n_seq) * L * (d / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  embed_text_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(tokens, tok_emb, pos, n_seq * L, L, d / 4, x);
  RLCF_CHECK_LAUNCH("embed_text");
  return 0;
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma = sum dy * xhat;  dbeta = sum dy.
// grid (n_slots, n_sets): block b of set s reduces its contiguous chunk of the set's rows and writes one
// deterministic partial; the AdamW kernel sums the slots.
// kSmemAcc: the per-warp dgamma / dbeta accumulators live in shared memory (`acc`, [8 warps][2][d] floats) instead of 48
// registers per lane.  ln_bwd_kernel needs 168 registers, so only ONE 256-thread block fits an SM (8 warps, each with two
// dependent memory round trips per row) and a 32-image launch runs 7 waves; ln_bwd_smem_kernel is bounded to 128
// registers -> two blocks per SM.  Same additions in the same order per warp and the same slot reduction, so the
// results are bit-identical.  The shared-memory variant is the DEFAULT (measured 197.7 -> 130.0 us per launch at the
// 32-image policy geometry; bit-identical for widths 768 / 1024, fp16 and fp32 dy -- scripts/dump_ln_bwd.py,
// tests/test_kernels_gpu.py::test_layernorm_bwd_smem_variant_is_bit_identical); RLCF_LN_BWD_SMEM=0 selects the register
// variant.
template <int NV, bool kDyF32, bool kSmemAcc>
__device__ __forceinline__ void ln_bwd_body(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x,
                                            long long ldx, const float* __restrict__ gamma, long long pstride,
                                            int rows_per_set, float eps, float* __restrict__ dx, long long lddx,
                                            int accumulate, __half* __restrict__ dx16, float* __restrict__ partials,
                                            long long p_total, long long p_off, float* acc) {
  constexpr int d = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int set = blockIdx.y, slot = blockIdx.x, n_slots = gridDim.x;
  const int rpb = (rows_per_set + n_slots - 1) / n_slots;
  const int r_begin = slot * rpb;
  const int r_end = min(rows_per_set, r_begin + rpb);
  const float4* g4 = reinterpret_cast<const float4*>(gamma + set * pstride);
  float4* acc_g = reinterpret_cast<float4*>(acc) + (warp * 2) * (d / 4);   // kSmemAcc only
  float4* acc_b = acc_g + d / 4;
  float4 gam[kSmemAcc ? 1 : NV], dg[kSmemAcc ? 1 : NV], db[kSmemAcc ? 1 : NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if constexpr (kSmemAcc) {
      acc_g[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc_b[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      gam[i] = __ldg(g4 + lane + 32 * i);
      dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int r = r_begin + warp; r < r_end; r += 8) {
    const long long row = static_cast<long long>(set) * rows_per_set + r;
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[NV], g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
    if constexpr (kDyF32) {
      const float4* dr = reinterpret_cast<const float4*>(static_cast<const float*>(dy_) + row * lddy);
#pragma unroll
      for (int i = 0; i < NV; ++i) g[i] = dr[lane + 32 * i];
    } else {
      const uint2* dr = reinterpret_cast<const uint2*>(static_cast<const __half*>(dy_) + row * lddy);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint2 q = dr[lane + 32 * i];
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
        g[i] = make_float4(a.x, a.y, b.x, b.y);
      }
    }
    float mean, rstd;
    ln_row_stats<NV>(v, d, eps, mean, rstd);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x = (v[i].x - mean) * rstd; v[i].y = (v[i].y - mean) * rstd;
      v[i].z = (v[i].z - mean) * rstd; v[i].w = (v[i].w - mean) * rstd;
      if constexpr (kSmemAcc) {
        float4 ag = acc_g[lane + 32 * i], ab = acc_b[lane + 32 * i];
        ag.x += g[i].x * v[i].x; ag.y += g[i].y * v[i].y; ag.z += g[i].z * v[i].z; ag.w += g[i].w * v[i].w;
        ab.x += g[i].x; ab.y += g[i].y; ab.z += g[i].z; ab.w += g[i].w;
        acc_g[lane + 32 * i] = ag;
        acc_b[lane + 32 * i] = ab;
        const float4 gm = __ldg(g4 + lane + 32 * i);
        g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;
      } else {
        dg[i].x += g[i].x * v[i].x; dg[i].y += g[i].y * v[i].y; dg[i].z += g[i].z * v[i].z; dg[i].w += g[i].w * v[i].w;
        db[i].x += g[i].x; db[i].y += g[i].y; db[i].z += g[i].z; db[i].w += g[i].w;
        g[i].x *= gam[i].x; g[i].y *= gam[i].y; g[i].z *= gam[i].z; g[i].w *= gam[i].w;
      }
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
    if (dx) {
      float4* o = reinterpret_cast<float4*>(dx + row * lddx);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 t;
        t.x = rstd * (g[i].x - s1 - v[i].x * s2); t.y = rstd * (g[i].y - s1 - v[i].y * s2);
        t.z = rstd * (g[i].z - s1 - v[i].z * s2); t.w = rstd * (g[i].w - s1 - v[i].w * s2);
        if (accumulate) {
          const float4 old = o[lane + 32 * i];
          t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
        }
        o[lane + 32 * i] = t;
        if (dx16) {  // fp16 copy of the updated residual gradient = A operand of the next dgrad GEMM
          __half2 h0 = __floats2half2_rn(t.x, t.y), h1 = __floats2half2_rn(t.z, t.w);
          *reinterpret_cast<uint2*>(dx16 + row * static_cast<long long>(d) + (lane + 32 * i) * 4) =
              make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
      }
    }
  }
  if (partials == nullptr) return;  // parameters frozen (text tower in prompt tuning): only dx was needed
  float* part = partials + (static_cast<long long>(set) * n_slots + slot) * p_total + p_off;
  if constexpr (kSmemAcc) {
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      for (int c = threadIdx.x; c < d; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += acc[(w * 2 + pass) * d + c];
        part[pass * d + c] = s;
      }
    }
  } else {
    float(*red)[d] = reinterpret_cast<float(*)[d]>(acc);   // [8][d]
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < NV; ++i)
        reinterpret_cast<float4*>(red[warp])[lane + 32 * i] = pass == 0 ? dg[i] : db[i];
      __syncthreads();
      for (int c = threadIdx.x; c < d; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][c];
        part[pass * d + c] = s;
      }
    }
  }
}

template <int NV, bool kDyF32>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x, long long ldx,
              const float* __restrict__ gamma, long long pstride, int rows_per_set, float eps,
              float* __restrict__ dx, long long lddx, int accumulate, __half* __restrict__ dx16,
              float* __restrict__ partials, long long p_total, long long p_off) {
  __shared__ float red[8 * NV * 128];
  ln_bwd_body<NV, kDyF32, false>(dy_, lddy, x, ldx, gamma, pstride, rows_per_set, eps, dx, lddx, accumulate, dx16,
                                 partials, p_total, p_off, red);
}

template <int NV, bool kDyF32>
__global__ void __launch_bounds__(256, 2)
ln_bwd_smem_kernel(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x, long long ldx,
                   const float* __restrict__ gamma, long long pstride, int rows_per_set, float eps,
                   float* __restrict__ dx, long long lddx, int accumulate, __half* __restrict__ dx16,
                   float* __restrict__ partials, long long p_total, long long p_off) {
  extern __shared__ float4 ln_acc[];   // [8 warps][2: dgamma, dbeta][d / 4]
  ln_bwd_body<NV, kDyF32, true>(dy_, lddy, x, ldx, gamma, pstride, rows_per_set, eps, dx, lddx, accumulate, dx16,
                                partials, p_total, p_off, reinterpret_cast<float*>(ln_acc));
}

template <int NV, bool kDyF32>
static cudaError_t launch_ln_bwd_smem(dim3 grid, cudaStream_t stream, const void* dy, long long lddy, const float* x,
                                      long long ldx, const float* gamma, long long pstride, int rows_per_set, float eps,
                                      float* dx, long long lddx, int accumulate, __half* dx16, float* partials,
                                      long long p_total, long long p_off) {
  constexpr int smem = 8 * 2 * NV * 128 * static_cast<int>(sizeof(float));
  static DynSmemState st;
  if (cudaError_t e = ensure_dyn_smem(ln_bwd_smem_kernel<NV, kDyF32>, smem, st)) return e;
  ln_bwd_smem_kernel<NV, kDyF32><<<grid, 256, smem, stream>>>(dy, lddy, x, ldx, gamma, pstride, rows_per_set, eps, dx,
                                                                lddx, accumulate, dx16, partials, p_total, p_off);
  return cudaSuccess;
}

int layernorm_bwd(const void* dy, int dy_is_f32, long long lddy, const float* x, long long ldx, const float* gamma,
                  long long pstride, int rows_per_set, int n_sets, int d, float eps, float* dx, long long lddx,
                  int accumulate, __half* dx16, float* partials, int n_slots, long long p_total, long long p_off,
                  cudaStream_t stream) {
  if (int rc = check_width(d, "layernorm_bwd")) return rc;
  if (rows_per_set <= 0 || n_sets <= 0 || n_slots <= 0) return set_error(RLCF_ERR_ARG, "layernorm_bwd: bad shape");
  if (partials == nullptr && dx == nullptr) return set_error(RLCF_ERR_ARG, "layernorm_bwd: nothing to compute");
  if (dx16 != nullptr && dx == nullptr) return set_error(RLCF_ERR_ARG, "layernorm_bwd: dx16 needs dx_accum");
  dim3 grid(n_slots, n_sets);
  static const bool smem_acc = getenv("RLCF_LN_BWD_SMEM") == nullptr || atoi(getenv("RLCF_LN_BWD_SMEM")) != 0;
  if (smem_acc) {
    cudaError_t e = cudaSuccess;
    if (dy_is_f32) {
      RLCF_DISPATCH_NV(d, (e = launch_ln_bwd_smem<NV, true>(grid, stream, dy, lddy, x, ldx, gamma, pstride, rows_per_set,
                                                             eps, dx, lddx, accumulate, dx16, partials, p_total, p_off)));
    } else {
      RLCF_DISPATCH_NV(d, (e = launch_ln_bwd_smem<NV, false>(grid, stream, dy, lddy, x, ldx, gamma, pstride, rows_per_set,
                                                              eps, dx, lddx, accumulate, dx16, partials, p_total, p_off)));
    }
    if (e != cudaSuccess) return set_error(RLCF_ERR_CUDA, "layernorm_bwd attr: %s", cudaGetErrorString(e));
    RLCF_CHECK_LAUNCH("layernorm_bwd");
    return 0;
  }
  if (dy_is_f32) {
    RLCF_DISPATCH_NV(d, (ln_bwd_kernel<NV, true><<<grid, 256, 0, stream>>>(dy, lddy, x, ldx, gamma, pstride,
                                                                            rows_per_set, eps, dx, lddx, accumulate,
                                                                            dx16, partials, p_total, p_off)));
  } else {
    RLCF_DISPATCH_NV(d, (ln_bwd_kernel<NV, false><<<grid, 256, 0, stream>>>(dy, lddy, x, ldx, gamma, pstride,
                                                                             rows_per_set, eps, dx, lddx, accumulate,
                                                                             dx16, partials, p_total, p_off)));
  }
  RLCF_CHECK_LAUNCH("layernorm_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------ head fwd
// ln_post(x[:,0]) @ proj (model.py:235-238) -> f/|f| -> logit_scale * f @ class_feat^T (custom_clip.py:423-432).
// One block handles VPC sequences so that every proj / class_feat element fetched from L2 is used VPC times.
// Everything stays fp32 (the final features are never rounded to fp16).  Dynamic smem: yT[d][VPC] + 2 f[VPC][E].
constexpr int kHeadThreads = 256;
template <int VPC>
__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_idx, long long row_stride,
                const float* __restrict__ gamma, const float* __restrict__ beta, long long pstride, int seqs_per_set,
                const float* __restrict__ proj, const float* __restrict__ cls_feat, float logit_scale, int n, int d,
                int E, int C, float eps, float* __restrict__ feat, float* __restrict__ inv_norm,
                float* __restrict__ logits, long long proj_stride) {
  extern __shared__ float sm[];
  float* yT = sm;            // [d][VPC]
  float* f = sm + d * VPC;   // [VPC][E]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * VPC;
  proj += static_cast<long long>(n0 / seqs_per_set) * proj_stride;   // per-set projection (host: VPC | seqs_per_set)
  const int nv = min(VPC, n - n0);
  // LayerNorm of each sequence's row by one warp (two-pass statistics)
  for (int v = warp; v < VPC; v += kHeadThreads / 32) {
    if (v < nv) {
      const int s = n0 + v;
      const long long row = row_idx ? row_idx[s] : s * row_stride;
      const float* xr = x + row * d;
      float sum = 0.f;
      for (int i = lane; i < d; i += 32) sum += xr[i];
      const float mean = warp_sum(sum) / d;
      float q = 0.f;
      for (int i = lane; i < d; i += 32) { const float a = xr[i] - mean; q += a * a; }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / d + eps);
      const long long po = static_cast<long long>(s / seqs_per_set) * pstride;
      for (int i = lane; i < d; i += 32) yT[i * VPC + v] = (xr[i] - mean) * rstd * gamma[po + i] + beta[po + i];
    } else {
      for (int i = lane; i < d; i += 32) yT[i * VPC + v] = 0.f;
    }
  }
  __syncthreads();
  // f = y @ proj: thread (jq, half) accumulates 4 adjacent columns over half of the rows of proj with 16-byte loads
  // (8 independent loads in flight per thread), the two halves are summed through shared memory.
  {
    const int E4 = E >> 2;
    float* fpart = f + VPC * E;   // [VPC][E] partial sums of the upper half
    for (int jq = tid & 127; jq < E4; jq += 128) {
      const int half = tid >> 7;
      const int i0 = half * (d >> 1), i1 = i0 + (d >> 1);
      float acc[VPC][4];
#pragma unroll
      for (int v = 0; v < VPC; ++v) acc[v][0] = acc[v][1] = acc[v][2] = acc[v][3] = 0.f;
#pragma unroll 8
      for (int i = i0; i < i1; ++i) {
        const float4 pj = __ldg(reinterpret_cast<const float4*>(proj + static_cast<size_t>(i) * E) + jq);
#pragma unroll
        for (int v = 0; v < VPC; ++v) {
          const float yv = yT[i * VPC + v];
          acc[v][0] = fmaf(yv, pj.x, acc[v][0]); acc[v][1] = fmaf(yv, pj.y, acc[v][1]);
          acc[v][2] = fmaf(yv, pj.z, acc[v][2]); acc[v][3] = fmaf(yv, pj.w, acc[v][3]);
        }
      }
      float* dst = half == 0 ? f : fpart;
#pragma unroll
      for (int v = 0; v < VPC; ++v)
        *reinterpret_cast<float4*>(dst + v * E + jq * 4) = make_float4(acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
    }
    __syncthreads();
    for (int j = tid; j < VPC * E; j += kHeadThreads) f[j] += fpart[j];
  }
  __syncthreads();
  for (int v = warp; v < nv; v += kHeadThreads / 32) {
    float ss = 0.f;
    for (int j = lane; j < E; j += 32) ss += f[v * E + j] * f[v * E + j];
    const float inv = 1.0f / sqrtf(warp_sum(ss));
    for (int j = lane; j < E; j += 32) {
      const float val = f[v * E + j] * inv;
      f[v * E + j] = val;
      if (feat) feat[static_cast<size_t>(n0 + v) * E + j] = val;
    }
    if (inv_norm && lane == 0) inv_norm[n0 + v] = inv;
  }
  __syncthreads();
  if (logits) {
    for (int c = warp; c < C; c += kHeadThreads / 32) {
      const float* t = cls_feat + static_cast<size_t>(c) * E;
      float acc[VPC];
#pragma unroll
      for (int v = 0; v < VPC; ++v) acc[v] = 0.f;
      for (int j = lane; j < E; j += 32) {
        const float tj = __ldg(t + j);
#pragma unroll
        for (int v = 0; v < VPC; ++v) acc[v] = fmaf(f[v * E + j], tj, acc[v]);
      }
#pragma unroll
      for (int v = 0; v < VPC; ++v) {
        const float r = warp_sum(acc[v]);
        if (lane == 0 && v < nv) logits[static_cast<size_t>(n0 + v) * C + c] = logit_scale * r;
      }
    }
  }
}

int head_fwd(const float* x, const int32_t* row_idx, long long row_stride, const float* gamma, const float* beta,
             long long pstride, int seqs_per_set, const float* proj, const float* cls_feat, float logit_scale, int n,
             int d, int E, int C, float eps, float* feat, float* inv_norm, float* logits, long long proj_stride,
             cudaStream_t stream) {
  if (n <= 0 || d <= 0 || E <= 0 || seqs_per_set <= 0) return set_error(RLCF_ERR_ARG, "head_fwd: bad shape");
  if (logits && (cls_feat == nullptr || C <= 0)) return set_error(RLCF_ERR_ARG, "head_fwd: logits need class_feat");
  if (proj_stride % 4) return set_error(RLCF_ERR_ARG, "head_fwd: proj_stride must be a multiple of 4 floats");
  int vpc = n >= 8 * 148 / 2 ? 8 : (n >= 64 ? 4 : 1);
  while (proj_stride != 0 && seqs_per_set % vpc != 0) vpc >>= 1;   // a block must not straddle two projections
  if (vpc == 2) vpc = 1;
  if (E % 4 != 0 || d % 2 != 0) return set_error(RLCF_ERR_ARG, "head_fwd: E must be a multiple of 4 and d even");
  const size_t smem = static_cast<size_t>(vpc) * (d + 2 * E) * sizeof(float);
  if (smem > 227 * 1024) return set_error(RLCF_ERR_ARG, "head_fwd: width too large");
#define RLCF_HEAD_LAUNCH(V)                                                                                       \
  {                                                                                                               \
    static DynSmemState st;                                                                                       \
    if (smem > 48 * 1024)                                                                                         \
      if (cudaError_t e = ensure_dyn_smem(head_fwd_kernel<V>, smem, st))                                          \
        return set_error(RLCF_ERR_CUDA, "head_fwd attr: %s", cudaGetErrorString(e));                              \
    head_fwd_kernel<V><<<(n + V - 1) / V, kHeadThreads, smem, stream>>>(                                          \
        x, row_idx, row_stride, gamma, beta, pstride, seqs_per_set, proj, cls_feat, logit_scale, n, d, E, C, eps, \
        feat, inv_norm, logits, proj_stride);                                                                     \
  }
  if (vpc == 8) RLCF_HEAD_LAUNCH(8) else if (vpc == 4) RLCF_HEAD_LAUNCH(4) else RLCF_HEAD_LAUNCH(1)
#undef RLCF_HEAD_LAUNCH
  RLCF_CHECK_LAUNCH("head_fwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------ entropy select
// tpt_cls_rl.py:32-35: H = -(softmax * log_softmax).sum(1); argsort ascending; keep the first S.
__global__ void __launch_bounds__(256)
entropy_select_kernel(const float* __restrict__ logits, int V, int C, int S, int32_t* __restrict__ sel,
                      int32_t* __restrict__ sel_global, float* __restrict__ entropy) {
  extern __shared__ float Hs[];  // [V]
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = warp; v < V; v += 8) {
    const float* r = logits + (static_cast<size_t>(img) * V + v) * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, r[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(r[c] - mx);
    const float lse = logf(warp_sum(se));
    float h = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float lp = r[c] - mx - lse;
      h += expf(lp) * lp;
    }
    h = -warp_sum(h);
    if (lane == 0) {
      Hs[v] = h;
      if (entropy) entropy[static_cast<size_t>(img) * V + v] = h;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float hi = Hs[i];
    int rank = 0;
    for (int j = 0; j < V; ++j) {
      const float hj = Hs[j];
      rank += (hj < hi) || (hj == hi && j < i);
    }
    if (rank < S) {
      sel[img * S + rank] = i;
      if (sel_global) sel_global[img * S + rank] = img * V + i;
    }
  }
}

int entropy_select(const float* logits, int n_img, int V, int C, int S, int32_t* sel, int32_t* sel_global,
                   float* entropy, cudaStream_t stream) {
  if (n_img <= 0 || V <= 0 || C <= 0 || S <= 0 || S > V) return set_error(RLCF_ERR_ARG, "entropy_select: bad shape");
  entropy_select_kernel<<<n_img, 256, V * sizeof(float), stream>>>(logits, V, C, S, sel, sel_global, entropy);
  RLCF_CHECK_LAUNCH("entropy_select");
  return 0;
}

// ------------------------------------------------------------------------------------------------ reward + loss
// tpt_cls_rl.py:63-71 and clip_reward.py:111-128,152-165.  One block per image, one warp per selected view.
constexpr int kMaxK = 8;
constexpr int kMaxRewardModels = 4;
// One or several frozen reward models (CLIPRewards / CLIPRewardsMultiple, clip_reward.py:43-178 / 180-307): model i has
// image features img[i] [n_views_total, er[i]] and class features cls[i] [C, er[i]]; the sample's score is
// sum_i wt[i] * max(0, w * <cls_i, img_i>)  (wt = normalised confidences, or 1/n for the plain mean; n = 1: wt = 1).
struct RewardSet {
  const float* img[kMaxRewardModels];
  const float* cls[kMaxRewardModels];
  int er[kMaxRewardModels];
  float wt[kMaxRewardModels];
  int n;
};

__global__ void __launch_bounds__(256)
reward_loss_kernel(const float* __restrict__ logits, const int32_t* __restrict__ row_idx, const RewardSet rs, int S,
                   int K, int C, float w, int reward_process, int process_batch, int amplify, float loss_scale,
                   float* __restrict__ dlogits, int32_t* __restrict__ topk_idx, float* __restrict__ scores_out,
                   float* __restrict__ rewards_out, float* __restrict__ loss_out) {
  extern __shared__ float smf[];
  float* sc = smf;                  // [S*K] scores, then rewards
  float* lse_s = sc + S * K;        // [S]
  float* mx_s = lse_s + S;          // [S]
  float* ce = mx_s + S;             // [S*K]
  int* idx = reinterpret_cast<int*>(ce + S * K);  // [S*K]
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < S; s += 8) {
    const int n = img * S + s;
    const float* r = logits + static_cast<size_t>(row_idx ? row_idx[n] : n) * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, r[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(r[c] - mx);
    const float lse = logf(warp_sum(se));
    int chosen[kMaxK];
    for (int k = 0; k < K; ++k) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = lane; c < C; c += 32) {
        bool taken = false;
        for (int j = 0; j < k; ++j) taken |= (chosen[j] == c);
        const float val = r[c];
        if (!taken && (val > bv || (val == bv && c < bi))) { bv = val; bi = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      chosen[k] = bi;
      // CLIPScore = max(0, w * <t_cls, f_img>)   (clip_reward.py:119-126); ensembles: weighted sum over the models
      // (clip_reward.py:226-250)
      float score = 0.f;
      for (int mdl = 0; mdl < rs.n; ++mdl) {
        const int Er = rs.er[mdl];
        const float* t = rs.cls[mdl] + static_cast<size_t>(bi) * Er;
        const float* f = rs.img[mdl] + static_cast<size_t>(n) * Er;
        float acc = 0.f;
        for (int j = lane; j < Er; j += 32) acc = fmaf(__ldg(t + j), __ldg(f + j), acc);
        acc = warp_sum(acc);
        const float sm = fmaxf(w * acc, 0.f);
        score = rs.n == 1 ? sm : score + rs.wt[mdl] * sm;
      }
      if (lane == 0) {
        sc[s * K + k] = score;
        ce[s * K + k] = (mx + lse) - bv;
        idx[s * K + k] = bi;
      }
    }
    if (lane == 0) { lse_s[s] = lse; mx_s[s] = mx; }
  }
  __syncthreads();
  if (scores_out)
    for (int i = threadIdx.x; i < S * K; i += blockDim.x) scores_out[static_cast<size_t>(img) * S * K + i] = sc[i];
  __syncthreads();
  // rewards_post_process (clip_reward.py:152-165); torch.std is the unbiased estimator
  if (threadIdx.x == 0 && reward_process) {
    if (process_batch) {
      const int n = S * K;
      if (n > 1) {
        float m = 0.f;
        for (int i = 0; i < n; ++i) m += sc[i];
        m /= n;
        float sd = 1.f;
        if (amplify) {
          float q = 0.f;
          for (int i = 0; i < n; ++i) q += (sc[i] - m) * (sc[i] - m);
          sd = sqrtf(q / (n - 1)) + 1e-5f;
        }
        for (int i = 0; i < n; ++i) sc[i] = (sc[i] - m) / sd;
      }
    } else if (K > 1) {
      for (int s = 0; s < S; ++s) {
        float m = 0.f;
        for (int k = 0; k < K; ++k) m += sc[s * K + k];
        m /= K;
        float sd = 1.f;
        if (amplify) {
          float q = 0.f;
          for (int k = 0; k < K; ++k) q += (sc[s * K + k] - m) * (sc[s * K + k] - m);
          sd = sqrtf(q / (K - 1)) + 1e-5f;
        }
        for (int k = 0; k < K; ++k) sc[s * K + k] = (sc[s * K + k] - m) / sd;
      }
    }
  }
  __syncthreads();
  const float inv_n = 1.f / (S * K);
  if (threadIdx.x == 0 && loss_out) {
    float l = 0.f;
    for (int i = 0; i < S * K; ++i) l += sc[i] * ce[i];
    loss_out[img] = l * inv_n;
  }
  for (int i = threadIdx.x; i < S * K; i += blockDim.x) {
    if (rewards_out) rewards_out[static_cast<size_t>(img) * S * K + i] = sc[i];
    if (topk_idx) topk_idx[static_cast<size_t>(img) * S * K + i] = idx[i];
  }
  // dL/dlogit[s,c] = (1/(S K)) * sum_k r[s,k] * (softmax[s,c] - [c == idx[s,k]])
  for (int s = warp; s < S; s += 8) {
    const int n = img * S + s;
    const float* r = logits + static_cast<size_t>(row_idx ? row_idx[n] : n) * C;
    float rs = 0.f;
    for (int k = 0; k < K; ++k) rs += sc[s * K + k];
    const float off = mx_s[s] + lse_s[s];
    float* o = dlogits + static_cast<size_t>(n) * C;
    for (int c = lane; c < C; c += 32) {
      float g = rs * expf(r[c] - off);
      for (int k = 0; k < K; ++k) g -= (idx[s * K + k] == c) ? sc[s * K + k] : 0.f;
      o[c] = g * inv_n * loss_scale;
    }
  }
}

int reward_loss_multi(const float* logits, const int32_t* row_idx, int n_models, const float* const* r_img,
                      const float* const* r_cls, const int* er, const float* wt, int n_img, int S, int K, int C, float w,
                      int reward_process, int process_batch, int amplify, float loss_scale, float* dlogits,
                      int32_t* topk_idx, float* scores, float* rewards, float* loss, cudaStream_t stream) {
  if (n_img <= 0 || S <= 0 || K <= 0 || K > kMaxK || K > C || n_models < 1 || n_models > kMaxRewardModels)
    return set_error(RLCF_ERR_ARG, "reward_loss: bad shape (K must be 1..%d, 1..%d reward models)", kMaxK,
                     kMaxRewardModels);
  RewardSet rs{};
  rs.n = n_models;
  for (int i = 0; i < n_models; ++i) {
    if (r_img[i] == nullptr || r_cls[i] == nullptr || er[i] <= 0)
      return set_error(RLCF_ERR_ARG, "reward_loss: reward model %d has no features", i);
    rs.img[i] = r_img[i]; rs.cls[i] = r_cls[i]; rs.er[i] = er[i]; rs.wt[i] = wt[i];
  }
  const size_t smem = (static_cast<size_t>(3) * S * K + 2 * S) * sizeof(float);
  if (smem > 48 * 1024) return set_error(RLCF_ERR_ARG, "reward_loss: S*K too large");
  reward_loss_kernel<<<n_img, 256, smem, stream>>>(logits, row_idx, rs, S, K, C, w, reward_process, process_batch,
                                                   amplify, loss_scale, dlogits, topk_idx, scores, rewards, loss);
  RLCF_CHECK_LAUNCH("reward_loss");
  return 0;
}

int reward_loss(const float* logits, const int32_t* row_idx, const float* r_img, const float* r_cls, int n_img, int S,
                int K, int C, int Er, float w, int reward_process, int process_batch, int amplify, float loss_scale,
                float* dlogits, int32_t* topk_idx, float* scores, float* rewards, float* loss, cudaStream_t stream) {
  const float one = 1.f;
  return reward_loss_multi(logits, row_idx, 1, &r_img, &r_cls, &Er, &one, n_img, S, K, C, w, reward_process,
                           process_batch, amplify, loss_scale, dlogits, topk_idx, scores, rewards, loss, stream);
}

// ------------------------------------------------------------------------------------------------ TPT entropy loss
// tpt_cls_rl.py:38-44: avg_logits = logsumexp_s(log_softmax(x_s)) - log S ; loss = -(avg * exp(avg)).sum()
// With p[s,c] = softmax, Pbar[c] = sum_s p[s,c]: exp(avg_c) = Pbar_c / S and
// d loss / d x[s,c] = p[s,c] * (G_c / Pbar_c - sum_c' G_c' p[s,c'] / Pbar_c'),  G_c = -(1 + avg_c) * Pbar_c / S.
__global__ void __launch_bounds__(256)
avg_entropy_kernel(const float* __restrict__ logits, const int32_t* __restrict__ row_idx, int S, int C,
                   float loss_scale, float* __restrict__ dlogits, float* __restrict__ loss_out, float loss_weight,
                   int accumulate) {
  extern __shared__ float smf[];
  float* off = smf;        // [S]  max + lse per view
  float* q = off + S;      // [C]  G_c / Pbar_c
  float* t2 = q + C;       // [S]
  float* scratch = t2 + S; // [16]
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < S; s += 8) {
    const int n = img * S + s;
    const float* r = logits + static_cast<size_t>(row_idx ? row_idx[n] : n) * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, r[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(r[c] - mx);
    se = warp_sum(se);
    if (lane == 0) off[s] = mx + logf(se);
  }
  __syncthreads();
  float lsum = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float pb = 0.f;
    for (int s = 0; s < S; ++s) {
      const int n = img * S + s;
      pb += expf(logits[static_cast<size_t>(row_idx ? row_idx[n] : n) * C + c] - off[s]);
    }
    const float e = pb / S;
    const float avg = fmaxf(logf(e), -3.402823466e38f);
    lsum -= avg * e;
    q[c] = pb > 0.f ? -(1.f + avg) / S : 0.f;
  }
  const float total = block_sum<256>(lsum, scratch);
  if (threadIdx.x == 0 && loss_out) loss_out[img] = (accumulate ? loss_out[img] : 0.f) + loss_weight * total;
  __syncthreads();
  for (int s = warp; s < S; s += 8) {
    const int n = img * S + s;
    const float* r = logits + static_cast<size_t>(row_idx ? row_idx[n] : n) * C;
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a += q[c] * expf(r[c] - off[s]);
    a = warp_sum(a);
    if (lane == 0) t2[s] = a;
  }
  __syncthreads();
  for (int s = warp; s < S; s += 8) {
    const int n = img * S + s;
    const float* r = logits + static_cast<size_t>(row_idx ? row_idx[n] : n) * C;
    float* o = dlogits + static_cast<size_t>(n) * C;
    for (int c = lane; c < C; c += 32)
      o[c] = (accumulate ? o[c] : 0.f) + loss_scale * expf(r[c] - off[s]) * (q[c] - t2[s]);
  }
}

int avg_entropy_loss(const float* logits, const int32_t* row_idx, int n_img, int S, int C, float loss_scale,
                     float* dlogits, float* loss, cudaStream_t stream, float weight, int accumulate) {
  if (n_img <= 0 || S <= 0 || C <= 0) return set_error(RLCF_ERR_ARG, "avg_entropy_loss: bad shape");
  const size_t smem = (static_cast<size_t>(2) * S + C + 16) * sizeof(float);
  avg_entropy_kernel<<<n_img, 256, smem, stream>>>(logits, row_idx, S, C, loss_scale * weight, dlogits, loss, weight,
                                                   accumulate);
  RLCF_CHECK_LAUNCH("avg_entropy_loss");
  return 0;
}

// ------------------------------------------------------------------------------------------------ head bwd
// Reverse of head_fwd for one selected view per block (grid = (S, n_img)); view s of image g writes its ln_post
// d(gamma), d(beta) into gradient slot s of set g, so the reduction over views stays deterministic.
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ x, const int32_t* __restrict__ row_idx,
                long long row_stride, const float* __restrict__ gamma, long long pstride,
                const float* __restrict__ proj, const float* __restrict__ cls_feat, float logit_scale,
                const float* __restrict__ feat, const float* __restrict__ inv_norm, int S, int d, int E, int C,
                float eps, float* __restrict__ dres, float* __restrict__ partials, int n_slots, long long p_total,
                long long p_off, long long dl_set, long long dl_s, long long dl_k, long long cls_stride,
                const float* __restrict__ beta, float* __restrict__ y_out, float* __restrict__ df_out,
                long long proj_stride) {
  extern __shared__ float sm[];
  float* df = sm;            // [E]   (first: read with 16-byte loads, E % 4 == 0)
  float* dy = df + E;        // [d]
  float* xh = dy + d;        // [d]
  float* scratch = xh + d;   // [16]
  float* dl = scratch + 16;  // [C]
  const int img = blockIdx.y, s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* gam = gamma + img * pstride;
  const int n = img * S + s;
  cls_feat += img * cls_stride;
  proj += img * proj_stride;
  for (int c = tid; c < C; c += kHeadThreads) dl[c] = dlogits[img * dl_set + s * dl_s + c * dl_k];
  __syncthreads();
  // d fhat = logit_scale * dlogits @ class_feat   (8 independent loads in flight per thread)
  float dot = 0.f;
  for (int j = tid; j < E; j += kHeadThreads) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int c = 0;
    for (; c + 4 <= C; c += 4) {
      a0 = fmaf(dl[c + 0], __ldg(cls_feat + static_cast<size_t>(c + 0) * E + j), a0);
      a1 = fmaf(dl[c + 1], __ldg(cls_feat + static_cast<size_t>(c + 1) * E + j), a1);
      a2 = fmaf(dl[c + 2], __ldg(cls_feat + static_cast<size_t>(c + 2) * E + j), a2);
      a3 = fmaf(dl[c + 3], __ldg(cls_feat + static_cast<size_t>(c + 3) * E + j), a3);
    }
    for (; c < C; ++c) a0 = fmaf(dl[c], __ldg(cls_feat + static_cast<size_t>(c) * E + j), a0);
    const float acc = ((a0 + a1) + (a2 + a3)) * logit_scale;
    df[j] = acc;
    dot += acc * feat[static_cast<size_t>(n) * E + j];
  }
  dot = block_sum<kHeadThreads>(dot, scratch);
  const float inv = inv_norm[n];
  // d f = (d fhat - fhat * <fhat, d fhat>) / |f|
  for (int j = tid; j < E; j += kHeadThreads) {
    df[j] = (df[j] - feat[static_cast<size_t>(n) * E + j] * dot) * inv;
    if (df_out) df_out[static_cast<size_t>(n) * E + j] = df[j];   // needed for d proj = y^T d f (full tuning)
  }
  __syncthreads();
  // d y = d f @ proj^T  (warp per output row, 16-byte loads along E)
  for (int i = warp; i < d; i += kHeadThreads / 32) {
    const float4* pr = reinterpret_cast<const float4*>(proj + static_cast<size_t>(i) * E);
    float acc = 0.f;
    for (int j4 = lane; j4 < (E >> 2); j4 += 32) {
      const float4 w = __ldg(pr + j4);
      const float4 g = reinterpret_cast<const float4*>(df)[j4];
      acc = fmaf(g.x, w.x, fmaf(g.y, w.y, fmaf(g.z, w.z, fmaf(g.w, w.w, acc))));
    }
    acc = warp_sum(acc);
    if (lane == 0) dy[i] = acc;
  }
  // ln_post backward on the class-token row
  const long long row = row_idx ? row_idx[n] : n * row_stride;
  const float* xr = x + row * d;
  float sx = 0.f;
  for (int i = tid; i < d; i += kHeadThreads) { xh[i] = xr[i]; sx += xh[i]; }
  const float mean = block_sum<kHeadThreads>(sx, scratch) / d;
  float sq = 0.f;
  for (int i = tid; i < d; i += kHeadThreads) { const float a = xh[i] - mean; sq += a * a; }
  const float rstd = 1.0f / sqrtf(block_sum<kHeadThreads>(sq, scratch) / d + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid; i < d; i += kHeadThreads) {
    xh[i] = (xh[i] - mean) * rstd;
    const float g = dy[i] * gam[i];
    s1 += g;
    s2 += g * xh[i];
  }
  s1 = block_sum<kHeadThreads>(s1, scratch) / d;
  s2 = block_sum<kHeadThreads>(s2, scratch) / d;
  float* o = dres + row * d;
  float* part = partials ? partials + (static_cast<long long>(img) * n_slots + s) * p_total + p_off : nullptr;
  for (int i = tid; i < d; i += kHeadThreads) {
    const float g = dy[i] * gam[i];
    o[i] = rstd * (g - s1 - xh[i] * s2);
    if (part) {
      part[i] = dy[i] * xh[i];
      part[d + i] = dy[i];
    }
    if (y_out) y_out[static_cast<size_t>(n) * d + i] = xh[i] * gam[i] + beta[img * pstride + i];  // ln_post output
  }
}

int head_bwd(const float* dlogits, const float* x, const int32_t* row_idx, long long row_stride, const float* gamma,
             long long pstride, const float* proj, const float* cls_feat, float logit_scale, const float* feat,
             const float* inv_norm, int n_img, int S, int d, int E, int C, float eps, float* dres, float* partials,
             int n_slots, long long p_total, long long p_off, long long dl_set, long long dl_s, long long dl_k,
             long long cls_stride, const float* beta, float* y_out, float* df_out, long long proj_stride,
             cudaStream_t stream) {
  if (proj_stride % 4) return set_error(RLCF_ERR_ARG, "head_bwd: proj_stride must be a multiple of 4 floats");
  if (n_img <= 0 || S <= 0 || d <= 0 || d > 1024 || E <= 0 || E % 4 != 0 || C <= 0)
    return set_error(RLCF_ERR_ARG, "head_bwd: bad shape");
  if (y_out != nullptr && beta == nullptr) return set_error(RLCF_ERR_ARG, "head_bwd: y_out needs beta");
  if (partials != nullptr && S > n_slots)
    return set_error(RLCF_ERR_ARG, "head_bwd: %d views per image need at least %d gradient slots", S, S);
  const size_t smem = (static_cast<size_t>(C) + E + 2 * d + 16) * sizeof(float);
  if (smem > 48 * 1024) return set_error(RLCF_ERR_ARG, "head_bwd: C too large for shared memory");
  dim3 grid(S, n_img);
  head_bwd_kernel<<<grid, kHeadThreads, smem, stream>>>(dlogits, x, row_idx, row_stride, gamma, pstride, proj,
                                                        cls_feat, logit_scale, feat, inv_norm, S, d, E, C, eps, dres,
                                                        partials, n_slots, p_total, p_off, dl_set, dl_s, dl_k,
                                                        cls_stride, beta, y_out, df_out, proj_stride);
  RLCF_CHECK_LAUNCH("head_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------ AdamW
// torch.optim.AdamW single-tensor update order (decoupled decay first, then bias-corrected Adam).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ partials,
             int n_slots, long long p_total, long long total, float lr, float b1, float b2, float eps, float wd,
             float bc1, float bc2_sqrt, float inv_scale, float* __restrict__ grad_out,
             const float* __restrict__ p_in, long long p_in_stride, int fresh) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long set = i / p_total, j = i % p_total;
    const float* pp = partials + set * n_slots * p_total + j;
    float g = 0.f;
    for (int s = 0; s < n_slots; ++s) g += pp[s * p_total];
    g *= inv_scale;
    if (grad_out) grad_out[i] = g;
    // fresh: first step after reset -- parameters come from the (shared) initial copy, moments start at zero
    const float p0 = p_in ? p_in[set * p_in_stride + j] : p[i];
    const float m0 = fresh ? 0.f : m[i], v0 = fresh ? 0.f : v[i];
    float w = p0 * (1.f - lr * wd);
    const float mi = m0 + (g - m0) * (1.f - b1);  // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v0 * b2 + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    w -= (lr / bc1) * (mi / denom);
    p[i] = w;
  }
}

int adamw_step(float* params, float* m, float* v, const float* partials, int n_sets, int n_slots, long long p_total,
               float lr, float b1, float b2, float eps, float wd, int step, float loss_scale, float* grad_out,
               const float* params_in, long long params_in_stride, int fresh, cudaStream_t stream) {
  if (n_sets <= 0 || n_slots <= 0 || p_total <= 0 || step < 1) return set_error(RLCF_ERR_ARG, "adamw: bad shape");
  const long long total = n_sets * p_total;
  const double bc1 = 1.0 - pow(static_cast<double>(b1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(b2), step);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adamw_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(params, m, v, partials, n_slots, p_total, total, lr, b1,
                                                             b2, eps, wd, static_cast<float>(bc1),
                                                             static_cast<float>(sqrt(bc2)), 1.f / loss_scale, grad_out,
                                                             params_in, params_in_stride, fresh);
  RLCF_CHECK_LAUNCH("adamw");
  return 0;
}

// Streaming AdamW for a whole encoder per sample (full tuning / retrieval TTA: 86 M parameters x n_sets): 16-byte
// accesses, 4 independent float4 groups in flight per thread, and the fp16 GEMM copy of the first n16 parameters is
// written in the same pass (saves re-reading the fp32 masters for the cast).  HBM-bound: 16 B read (g, p, m, v) +
// 12 B written (p, m, v) + 2 B (fp16 copy) per parameter.
__device__ __forceinline__ float adamw_one(float p0, float g, float& m, float& v, float lr, float b1, float b2,
                                           float eps, float wd, float bc1, float bc2_sqrt) {
  return adamw_update(p0, g, m, v, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
}

__global__ void __launch_bounds__(256)
adamw_full_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ grads,
                  long long p4, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                  float inv_scale, const float* __restrict__ p_in, long long p_in_stride, int fresh,
                  __half* __restrict__ w16, long long w16_stride, long long n16_4, long long set_stride4) {
  const long long set = blockIdx.y;
  const long long base = set * set_stride4;   // in float4 units
  float4* P = reinterpret_cast<float4*>(p) + base;
  float4* M = reinterpret_cast<float4*>(m) + base;
  float4* V = reinterpret_cast<float4*>(v) + base;
  const float4* G = reinterpret_cast<const float4*>(grads) + base;
  const float4* PI = reinterpret_cast<const float4*>(p_in + set * p_in_stride);
  uint2* W = w16 ? reinterpret_cast<uint2*>(w16 + set * w16_stride) : nullptr;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  constexpr int U = 4;
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < p4; i0 += U * stride) {
    float4 g[U], q[U], mm[U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < p4) {
        g[u] = __ldcs(G + i);
        q[u] = fresh ? __ldg(PI + i) : __ldcs(PI + i);
        if (!fresh) { mm[u] = __ldcs(M + i); vv[u] = __ldcs(V + i); }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= p4) continue;
      if (fresh) { mm[u] = make_float4(0.f, 0.f, 0.f, 0.f); vv[u] = mm[u]; }
      float4 w;
      w.x = adamw_one(q[u].x, __fmul_rn(g[u].x, inv_scale), mm[u].x, vv[u].x, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
      w.y = adamw_one(q[u].y, __fmul_rn(g[u].y, inv_scale), mm[u].y, vv[u].y, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
      w.z = adamw_one(q[u].z, __fmul_rn(g[u].z, inv_scale), mm[u].z, vv[u].z, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
      w.w = adamw_one(q[u].w, __fmul_rn(g[u].w, inv_scale), mm[u].w, vv[u].w, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
      __stcs(P + i, w);
      __stcs(M + i, mm[u]);
      __stcs(V + i, vv[u]);
      if (W != nullptr && i < n16_4) {
        const __half2 h0 = __floats2half2_rn(w.x, w.y), h1 = __floats2half2_rn(w.z, w.w);
        W[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
    }
  }
}

int adamw_full(float* params, float* m, float* v, const float* grads, int n_sets, long long p_total, float lr,
               float b1, float b2, float eps, float wd, int step, float loss_scale, const float* params_in,
               long long params_in_stride, int fresh, __half* w16, long long w16_stride, long long n16,
               long long set_stride, cudaStream_t stream) {
  if (n_sets <= 0 || n_sets > 65535 || p_total <= 0 || step < 1) return set_error(RLCF_ERR_ARG, "adamw_full: bad shape");
  if (set_stride == 0) set_stride = p_total;
  if (p_total % 4 || params_in_stride % 4 || n16 % 4 || w16_stride % 4 || n16 > p_total || set_stride % 4 ||
      set_stride < p_total)
    return set_error(RLCF_ERR_ARG, "adamw_full: sizes and strides must be multiples of 4 elements");
  const double bc1 = 1.0 - pow(static_cast<double>(b1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(b2), step);
  const long long p4 = p_total / 4;
  long long bx = (p4 + 256 * 4 - 1) / (256 * 4);
  const long long cap = (148 * 8 + n_sets - 1) / n_sets;   // ~8 resident blocks per SM over all sets
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), n_sets);
  adamw_full_kernel<<<grid, 256, 0, stream>>>(params, m, v, grads, p4, lr, b1, b2, eps, wd, static_cast<float>(bc1),
                                              static_cast<float>(sqrt(bc2)), 1.f / loss_scale, params_in,
                                              params_in_stride, fresh, w16, w16_stride, n16 / 4, set_stride / 4);
  RLCF_CHECK_LAUNCH("adamw_full");
  return 0;
}

// out16[g][c][r] = in16[g][r][c]: transposed fp16 weight copies (dgrad B operands) from the fp16 copies the AdamW pass
// has just written -- 2 B read + 2 B written per weight instead of 4 + 2 from the fp32 masters.
__global__ void transpose_f16_kernel(const __half* __restrict__ in, int rows, int cols, __half* __restrict__ out,
                                     long long set_stride) {
  __shared__ __half tile[64][66];
  in += blockIdx.z * set_stride;
  out += blockIdx.z * set_stride;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 256 threads: 32 x 8
  for (int j = ty; j < 64; j += 8) {
    const int r = r0 + j, c = c0 + 2 * tx;
    __half2 val = __floats2half2_rn(0.f, 0.f);
    if (r < rows && c < cols) val = *reinterpret_cast<const __half2*>(in + static_cast<size_t>(r) * cols + c);
    tile[j][2 * tx] = __low2half(val);
    tile[j][2 * tx + 1] = __high2half(val);
  }
  __syncthreads();
  for (int j = ty; j < 64; j += 8) {
    const int c = c0 + j, r = r0 + 2 * tx;
    if (c < cols && r < rows)
      *reinterpret_cast<__half2*>(out + static_cast<size_t>(c) * rows + r) = __halves2half2(tile[2 * tx][j], tile[2 * tx + 1][j]);
  }
}

int transpose_f16(const __half* in, int rows, int cols, __half* out, int n_sets, long long set_stride,
                  cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || rows % 2 || cols % 2 || n_sets <= 0 || n_sets > 65535)
    return set_error(RLCF_ERR_ARG, "transpose_f16: bad shape (rows, cols even)");
  dim3 grid((cols + 63) / 64, (rows + 63) / 64, n_sets);
  transpose_f16_kernel<<<grid, 256, 0, stream>>>(in, rows, cols, out, set_stride);
  RLCF_CHECK_LAUNCH("transpose_f16");
  return 0;
}

__global__ void reset_params_kernel(const float* __restrict__ init, float* __restrict__ p, float* __restrict__ m,
                                    float* __restrict__ v, long long p_total, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    p[i] = init[i % p_total];
    if (m) m[i] = 0.f;
    if (v) v[i] = 0.f;
  }
}

int reset_params(const float* init, float* params, float* m, float* v, int n_sets, long long p_total,
                 cudaStream_t stream) {
  if (n_sets <= 0 || p_total <= 0) return set_error(RLCF_ERR_ARG, "reset_params: bad shape");
  const long long total = n_sets * p_total;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  reset_params_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(init, params, m, v, p_total, total);
  RLCF_CHECK_LAUNCH("reset_params");
  return 0;
}

// ------------------------------------------------------------------------------------------------ weight prep
__global__ void cast_f16_kernel(const float* __restrict__ in, long long rows, long long cols, long long ld_in,
                                __half* __restrict__ out, long long ld_out) {
  const long long total = rows * ld_out;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / ld_out, c = i % ld_out;
    out[i] = __float2half_rn(c < cols ? in[r * ld_in + c] : 0.f);
  }
}

int cast_f16(const float* in, long long rows, long long cols, long long ld_in, __half* out, long long ld_out,
             cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || ld_out < cols) return set_error(RLCF_ERR_ARG, "cast_f16: bad shape");
  long long blocks = (rows * ld_out + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_f16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, rows, cols, ld_in, out, ld_out);
  RLCF_CHECK_LAUNCH("cast_f16");
  return 0;
}

// dst[layer][j] = src[layer][idx[j]] for blocks ("sequences") of seq_bytes bytes: the activations of the selected
// views are lifted out of the all-views store of the 64-view pass, one launch per tensor (grid.y = layer).  16-byte
// vectors when sizes and pointers allow, else 4-byte words (the log-sum-exp rows of small towers).
template <typename T>
__global__ void __launch_bounds__(256) gather_seqs_kernel(const T* __restrict__ src, const int* __restrict__ idx,
                                                          T* __restrict__ dst, long long seq_vec, long long src_layer_vec,
                                                          long long dst_layer_vec, int n) {
  src += blockIdx.y * src_layer_vec;
  dst += blockIdx.y * dst_layer_vec;
  const long long total = static_cast<long long>(n) * seq_vec;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long j = i / seq_vec, o = i - j * seq_vec;
    dst[i] = src[static_cast<long long>(idx[j]) * seq_vec + o];
  }
}

int gather_seqs(const void* src, const int* idx, void* dst, long long seq_bytes, long long src_layer_bytes,
                long long dst_layer_bytes, int n_layers, int n, cudaStream_t stream) {
  if (n <= 0 || n_layers <= 0 || seq_bytes <= 0) return set_error(RLCF_ERR_ARG, "gather_seqs: bad shape");
  const uintptr_t bits = static_cast<uintptr_t>(seq_bytes | src_layer_bytes | dst_layer_bytes) |
                         reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
  if (bits & 3) return set_error(RLCF_ERR_ARG, "gather_seqs: sizes and pointers must be multiples of 4 bytes");
  const int vec = (bits & 15) ? 4 : 16;
  const long long total = static_cast<long long>(n) * (seq_bytes / vec);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const dim3 grid(static_cast<unsigned>(blocks), n_layers);
  if (vec == 16)
    gather_seqs_kernel<uint4><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(src), idx, static_cast<uint4*>(dst),
                                                        seq_bytes / 16, src_layer_bytes / 16, dst_layer_bytes / 16, n);
  else
    gather_seqs_kernel<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t*>(src), idx,
                                                           static_cast<uint32_t*>(dst), seq_bytes / 4, src_layer_bytes / 4,
                                                           dst_layer_bytes / 4, n);
  RLCF_CHECK_LAUNCH("gather_seqs");
  return 0;
}

__global__ void transpose_cast_kernel(const float* __restrict__ in, int rows, int cols, __half* __restrict__ out,
                                      long long in_stride, long long out_stride) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  in += blockIdx.z * in_stride;     // one parameter set per grid.z
  out += blockIdx.z * out_stride;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[static_cast<size_t>(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[static_cast<size_t>(c) * rows + r] = __float2half_rn(tile[threadIdx.x][j]);
  }
}

int transpose_cast_f16(const float* in, int rows, int cols, __half* out, int n_sets, long long in_stride,
                       long long out_stride, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || n_sets <= 0 || n_sets > 65535)
    return set_error(RLCF_ERR_ARG, "transpose_cast_f16: bad shape");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, n_sets), block(32, 8);
  transpose_cast_kernel<<<grid, block, 0, stream>>>(in, rows, cols, out, in_stride, out_stride);
  RLCF_CHECK_LAUNCH("transpose_cast_f16");
  return 0;
}

}  // namespace rlcf
