// fp32 (CUDA-core) kernels of the once-per-dataset text path: class features are an INPUT of the per-image loop
// (CLIPCLS_TTA.get_class_features, TPT/clip/custom_clip.py:404-408; CLIPRewards.extract_text_features,
// TPT/clip_reward.py:139-150; CLIP.encode_text, TPT/clip/model.py:342-356) and are computed once for C prompts of 77
// tokens, so they are kept at the reference's own precision: fp32 operands and fp32 accumulation end to end instead of
// the fp16 tensor-core operands of the per-image towers.  ~1.2 TFLOP for 200 prompts: tens of milliseconds on the FP32
// pipes, paid once per dataset.
#include "rlcf_internal.h"

namespace rlcf {

// ------------------------------------------------------------------------------------------------ SGEMM
// out[M,N] = epi(A[M,K] * W[N,K]^T + bias);  A rows lda apart, W rows ldw apart (both K-contiguous, the PyTorch Linear
// layout), out rows ldo apart.  128 x 128 tile per 256-thread block, 8 x 8 micro-tile per thread, K stepped by 16 with a
// register-staged double buffer.  epi: 0 none, 1 QuickGELU (x * sigmoid(1.702 x), model.py:166-168), 2 + resid.
constexpr int kSgBM = 128, kSgBN = 128, kSgBK = 16, kSgThreads = 256;

__device__ __forceinline__ float quick_gelu_f32(float u) { return u / (1.f + expf(-1.702f * u)); }

template <int kEpi>
__global__ void __launch_bounds__(kSgThreads)
sgemm_nt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw, int M, int N,
                int K, const float* __restrict__ bias, const float* __restrict__ resid, float* __restrict__ out,
                long long ldo) {
  __shared__ float As[2][kSgBK][kSgBM + 4];
  __shared__ float Ws[2][kSgBK][kSgBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kSgBM, n0 = blockIdx.x * kSgBN;
  // global -> smem staging: each thread moves 2 float4 of A and 2 of W per K step (128 rows x 16 k = 512 float4)
  const int ld_row = tid >> 2;          // 0..63 (+64 for the second)
  const int ld_k = (tid & 3) * 4;       // 0,4,8,12
  float4 ra[2], rw[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = ld_row + 64 * i;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + ld_k;
      ra[i] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + static_cast<long long>(gm) * lda + gk)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[i] = (gn < N && gk < K) ? *reinterpret_cast<const float4*>(W + static_cast<long long>(gn) * ldw + gk)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = ld_row + 64 * i;
      As[buf][ld_k + 0][r] = ra[i].x; As[buf][ld_k + 1][r] = ra[i].y;
      As[buf][ld_k + 2][r] = ra[i].z; As[buf][ld_k + 3][r] = ra[i].w;
      Ws[buf][ld_k + 0][r] = rw[i].x; Ws[buf][ld_k + 1][r] = rw[i].y;
      Ws[buf][ld_k + 2][r] = rw[i].z; Ws[buf][ld_k + 3][r] = rw[i].w;
    }
  };
  const int tx = tid & 15, ty = tid >> 4;   // thread's micro-tile: rows ty*4 + {0..3} and 64 + ty*4 + {0..3}; cols alike
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  const int n_k = (K + kSgBK - 1) / kSgBK;
  for (int kt = 0; kt < n_k; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_k) load_tiles((kt + 1) * kSgBK);
#pragma unroll
    for (int kk = 0; kk < kSgBK; ++kk) {
      float a[8], w[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < n_k) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int gn = n0 + jh * 64 + tx * 4;
      if (gn >= N) continue;            // N % 4 == 0: a float4 group is inside or outside as a whole
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float u = acc[i][jh * 4 + j] + (bias != nullptr ? bias[gn + j] : 0.f);
        if (kEpi == 1) u = quick_gelu_f32(u);
        if (kEpi == 2) u += resid[static_cast<long long>(gm) * ldo + gn + j];
        v[j] = u;
      }
      *reinterpret_cast<float4*>(out + static_cast<long long>(gm) * ldo + gn) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

int gemm_f32(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K, int epi,
             const float* bias, const float* resid, float* out, long long ldo, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N % 4) || (K % 4) || (lda % 4) || (ldw % 4) || (ldo % 4))
    return set_error(RLCF_ERR_ARG, "gemm_f32: N, K and the leading dimensions must be multiples of 4");
  if (epi == 2 && resid == nullptr) return set_error(RLCF_ERR_ARG, "gemm_f32: residual epilogue without resid");
  dim3 grid((N + kSgBN - 1) / kSgBN, (M + kSgBM - 1) / kSgBM);
  if (epi == 0) sgemm_nt_kernel<0><<<grid, kSgThreads, 0, stream>>>(A, lda, W, ldw, M, N, K, bias, resid, out, ldo);
  else if (epi == 1) sgemm_nt_kernel<1><<<grid, kSgThreads, 0, stream>>>(A, lda, W, ldw, M, N, K, bias, resid, out, ldo);
  else if (epi == 2) sgemm_nt_kernel<2><<<grid, kSgThreads, 0, stream>>>(A, lda, W, ldw, M, N, K, bias, resid, out, ldo);
  else return set_error(RLCF_ERR_ARG, "gemm_f32: unknown epilogue %d", epi);
  RLCF_CHECK_LAUNCH("gemm_f32");
  return 0;
}

// ------------------------------------------------------------------------------------------------ attention, fp32
// One block per (sequence, head): K and V rows of the head in shared memory (padded to 65 floats), one warp per query:
// lanes own keys for the scores (full 64-wide dot products), then dimensions for the P.V sum.  softmax in fp32 with
// the exact expf; causal = the text tower's additive -inf mask above the diagonal (model.py:328-334).
constexpr int kPaWarps = 8;

__global__ void __launch_bounds__(kPaWarps * 32)
attn_f32_kernel(const float* __restrict__ qkv, int L, int heads, int causal, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* Ks = sm;                         // [L][65]
  float* Vs = Ks + static_cast<size_t>(L) * 65;   // [L][65]
  float* Ps = Vs + static_cast<size_t>(L) * 65;   // [kPaWarps][L]
  float* Qs = Ps + static_cast<size_t>(kPaWarps) * L;   // [kPaWarps][64]
  const int h = blockIdx.x, s = blockIdx.y;
  const int d3 = 3 * heads * 64;
  const float* base = qkv + static_cast<size_t>(s) * L * d3;
  for (int i = threadIdx.x; i < L * 64; i += blockDim.x) {
    const int t = i >> 6, c = i & 63;
    Ks[t * 65 + c] = base[static_cast<size_t>(t) * d3 + heads * 64 + h * 64 + c];
    Vs[t * 65 + c] = base[static_cast<size_t>(t) * d3 + 2 * heads * 64 + h * 64 + c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* p = Ps + warp * L;
  float* q = Qs + warp * 64;
  for (int t = warp; t < L; t += kPaWarps) {
    q[lane] = base[static_cast<size_t>(t) * d3 + h * 64 + lane] * 0.125f;        // head_dim^-0.5
    q[lane + 32] = base[static_cast<size_t>(t) * d3 + h * 64 + lane + 32] * 0.125f;
    __syncwarp();
    const int n_keys = causal ? t + 1 : L;
    float mx = -INFINITY;
    for (int j = lane; j < n_keys; j += 32) {
      float dot = 0.f;
#pragma unroll 16
      for (int c = 0; c < 64; ++c) dot = fmaf(q[c], Ks[j * 65 + c], dot);
      p[j] = dot;
      mx = fmaxf(mx, dot);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < n_keys; j += 32) {
      const float e = expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < n_keys; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, Vs[j * 65 + lane], o0);
      o1 = fmaf(pj, Vs[j * 65 + lane + 32], o1);
    }
    const float inv = 1.f / sum;
    float* orow = out + (static_cast<size_t>(s) * L + t) * heads * 64 + h * 64;
    orow[lane] = o0 * inv;
    orow[lane + 32] = o1 * inv;
    __syncwarp();
  }
}

int attention_f32(const float* qkv, int n_seq, int L, int heads, int causal, float* out, cudaStream_t stream) {
  if (n_seq <= 0 || L <= 0 || heads <= 0) return set_error(RLCF_ERR_ARG, "attention_f32: bad shape");
  const size_t smem = (static_cast<size_t>(2) * L * 65 + static_cast<size_t>(kPaWarps) * L + kPaWarps * 64) * sizeof(float);
  if (smem > 227 * 1024) return set_error(RLCF_ERR_ARG, "attention_f32: sequence %d too long (the fp32 path serves the "
                                                        "text tower and short image sequences)", L);
  static DynSmemState st;
  if (smem > 48 * 1024)
    if (cudaError_t e = ensure_dyn_smem(attn_f32_kernel, smem, st))
      return set_error(RLCF_ERR_CUDA, "attention_f32 attr: %s", cudaGetErrorString(e));
  if (n_seq > 65535) return set_error(RLCF_ERR_ARG, "attention_f32: too many sequences per launch");
  dim3 grid(heads, n_seq);
  attn_f32_kernel<<<grid, kPaWarps * 32, smem, stream>>>(qkv, L, heads, causal, out);
  RLCF_CHECK_LAUNCH("attention_f32");
  return 0;
}

}  // namespace rlcf
