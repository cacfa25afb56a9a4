// Fused multi-head attention core for CLIP towers (head_dim 64, sequence <= 577 tokens): forward with
// online softmax and the exact backward (dQ, dK, dV), both on warp-level tensor-core MMAs
// (mma.sync m16n8k16, fp16 in / fp32 accumulate) with K/V (and Q/dO in the backward) resident in shared memory.
// Replaces nn.MultiheadAttention's bmm + softmax + bmm (TPT/clip/model.py:185-187) and its autograd.
// The whole sequence of a (view, head) fits one CTA's shared memory, so no KV tiling over HBM is needed:
// qkv is read once, the output written once.
#include <cstdlib>

#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

constexpr int kHd = 64;             // head dim
constexpr int kRowBytes = kHd * 2;  // 128 B per token row in smem

// 16-byte chunk `chunk` (0..7) of row `row` in a [rows][64] fp16 tile, XOR-swizzled against bank conflicts.
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {
  return static_cast<uint32_t>(row) * kRowBytes + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Copies rows [row0, row0+rows_pad) x 64 halfs of one head (global row stride ld halfs) into a swizzled tile;
// rows >= rows_valid (sequence-relative) are zero-filled.
__device__ __forceinline__ void load_tile(uint8_t* tile, const __half* g, long long ld, int row0, int rows_pad,
                                          int rows_valid, int tid, int nthreads) {
  for (int i = tid; i < rows_pad * 8; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < rows_valid) val = *reinterpret_cast<const uint4*>(g + (row0 + r) * ld + c * 8);
    *reinterpret_cast<uint4*>(tile + sw_off(r, c)) = val;
  }
}

// A-operand fragments (16 rows x 64 k) of rows [row0,row0+16) of a tile: 4 k-steps x 4 regs.
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], uint32_t tile, int row0, int lane) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], tile + sw_off(r, ks * 2 + (lane >> 4)));
}

// C[16 x 16] (two n-tiles) = A[16 x 64] * T[n0..n0+16][0..64]^T, T rows are the n index (K-major B operand).
__device__ __forceinline__ void mma_nt_16x16(float (&c)[2][4], const uint32_t (&a)[4][4], uint32_t tile, int n0,
                                             int lane) {
  const int r = n0 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t b[4];
    ldsm_x4(b, tile + sw_off(r, ks * 2 + ((lane >> 3) & 1)));
    mma16816(c[0], a[ks], b[0], b[1]);
    mma16816(c[1], a[ks], b[2], b[3]);
  }
}

// acc[16 x 64] += P[16 x 16] * T[k0..k0+16][0..64], T rows are the k index (needs transposed ldmatrix).
__device__ __forceinline__ void mma_nn_16x64(float (&acc)[8][4], const uint32_t (&p)[4], uint32_t tile, int k0,
                                             int lane) {
  const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t b[4];
    ldsm_x4_t(b, tile + sw_off(r, np * 2 + (lane >> 4)));
    mma16816(acc[2 * np], p, b[0], b[1]);
    mma16816(acc[2 * np + 1], p, b[2], b[3]);
  }
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int kFwdWarps = 4;
__global__ void __launch_bounds__(kFwdWarps * 32)
attn_fwd_kernel(const __half* __restrict__ qkv, int L, int Lp, int heads, int causal, __half* __restrict__ out,
                float* __restrict__ lse_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int d = heads * kHd;
  const long long ld = 3LL * d;
  const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (kFwdWarps * 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* sK = smem;
  uint8_t* sV = sK + Lp * kRowBytes;
  uint8_t* sQ = sV + Lp * kRowBytes;
  const __half* base = qkv + static_cast<long long>(seq) * L * ld + h * kHd;
  // causal: keys beyond the last query of this CTA are never needed
  const int kv_rows = causal ? min(Lp, ((min(q0 + kFwdWarps * 16, L) + 15) / 16) * 16) : Lp;
  load_tile(sK, base + d, ld, 0, kv_rows, L, tid, kFwdWarps * 32);
  load_tile(sV, base + 2 * d, ld, 0, kv_rows, L, tid, kFwdWarps * 32);
  load_tile(sQ, base, ld, q0, kFwdWarps * 16, L, tid, kFwdWarps * 32);
  __syncthreads();

  const int qw = q0 + warp * 16;  // first query row of this warp
  if (qw >= L) return;
  uint32_t qa[4][4];
  load_a_frags(qa, smem_u32(sQ), warp * 16, lane);

  const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int g = lane >> 2, t = lane & 3;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  const int k_end = causal ? min(kv_rows, ((qw + 16 + 15) / 16) * 16) : kv_rows;
  for (int k0 = 0; k0 < k_end; k0 += 16) {
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    mma_nt_16x16(s, qa, smem_u32(sK), k0, lane);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = k0 + nt * 8 + 2 * t + (e & 1);
        const int qrow = qw + g + (e >> 1) * 8;
        const bool ok = key < L && (!causal || key <= qrow);
        s[nt][e] = ok ? s[nt][e] * c : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      // m_new stays -inf only while every key so far is masked (padded query rows under the causal mask)
      corr[r] = m_new == -INFINITY ? 1.f : exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float mref = m_run[e >> 1];
        const float pv = mref == -INFINITY ? 0.f : exp2f(s[nt][e] - mref);
        s[nt][e] = pv;
        rs[e >> 1] += pv;
      }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    uint32_t pa[4] = {pack_h2(s[0][0], s[0][1]), pack_h2(s[0][2], s[0][3]), pack_h2(s[1][0], s[1][1]),
                      pack_h2(s[1][2], s[1][3])};
    mma_nn_16x64(o, pa, smem_u32(sV), k0, lane);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qrow = qw + g + r * 8;
    if (qrow >= L) continue;
    const float inv = 1.f / l_run[r];
    __half* orow = out + (static_cast<long long>(seq) * L + qrow) * d + h * kHd;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * t) = pack_h2(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
    if (lse_out && t == 0)
      lse_out[(static_cast<long long>(seq) * heads + h) * L + qrow] =
          m_run[r] * 0.6931471805599453f + logf(l_run[r]);  // natural-log LSE of the scaled scores
  }
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per (sequence, head).  Phase A: each warp owns 16 queries and produces dQ.  Phase B: each warp owns
// 16 keys and produces dK, dV.  Both recompute P from Q, K and the saved log-sum-exp.
constexpr int kBwdWarps = 13;  // 197 tokens = 13 blocks of 16 rows: one block per warp in each phase
__global__ void __launch_bounds__(kBwdWarps * 32)
attn_bwd_kernel(const __half* __restrict__ qkv, const __half* __restrict__ out, const __half* __restrict__ dout,
                const float* __restrict__ lse_in, int L, int Lp, int heads, int causal, __half* __restrict__ dqkv) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int d = heads * kHd;
  const long long ld = 3LL * d;
  const int seq = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Lp * kRowBytes;
  uint8_t* sV = sK + Lp * kRowBytes;
  uint8_t* sdO = sV + Lp * kRowBytes;
  float* sLse = reinterpret_cast<float*>(sdO + Lp * kRowBytes);  // scaled by log2(e)
  float* sD = sLse + Lp;
  const __half* base = qkv + static_cast<long long>(seq) * L * ld + h * kHd;
  const __half* obase = out + static_cast<long long>(seq) * L * d + h * kHd;
  const __half* dobase = dout + static_cast<long long>(seq) * L * d + h * kHd;
  load_tile(sQ, base, ld, 0, Lp, L, tid, kBwdWarps * 32);
  load_tile(sK, base + d, ld, 0, Lp, L, tid, kBwdWarps * 32);
  load_tile(sV, base + 2 * d, ld, 0, Lp, L, tid, kBwdWarps * 32);
  load_tile(sdO, dobase, d, 0, Lp, L, tid, kBwdWarps * 32);
  for (int r = tid; r < Lp; r += kBwdWarps * 32) {
    float dsum = 0.f, lv = 0.f;
    if (r < L) {
      const uint4* po = reinterpret_cast<const uint4*>(obase + static_cast<long long>(r) * d);
      const uint4* pd = reinterpret_cast<const uint4*>(dobase + static_cast<long long>(r) * d);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 a = po[c8], b = pd[c8];
        const __half2* ha = reinterpret_cast<const __half2*>(&a);
        const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __half22float2(ha[e]), fb = __half22float2(hb[e]);
          dsum += fa.x * fb.x + fa.y * fb.y;
        }
      }
      lv = lse_in[(static_cast<long long>(seq) * heads + h) * L + r] * 1.4426950408889634f;
    }
    sD[r] = dsum;
    sLse[r] = lv;
  }
  __syncthreads();

  const float scale = 0.125f;
  const float c = scale * 1.4426950408889634f;
  const int g = lane >> 2, t = lane & 3;
  const int nblk = Lp / 16;

  // ---------------- phase A: dQ
  for (int qb = warp; qb < nblk; qb += kBwdWarps) {
    const int q0 = qb * 16;
    if (q0 >= L) break;
    uint32_t qa[4][4], doa[4][4];
    load_a_frags(qa, smem_u32(sQ), q0, lane);
    load_a_frags(doa, smem_u32(sdO), q0, lane);
    const float lse_r[2] = {sLse[q0 + g], sLse[q0 + g + 8]};
    const float d_r[2] = {sD[q0 + g], sD[q0 + g + 8]};
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const int k_end = causal ? min(Lp, q0 + 16) : Lp;
    for (int k0 = 0; k0 < k_end; k0 += 16) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_nt_16x16(s, qa, smem_u32(sK), k0, lane);
      mma_nt_16x16(dp, doa, smem_u32(sV), k0, lane);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = k0 + nt * 8 + 2 * t + (e & 1);
          const int qrow = q0 + g + (e >> 1) * 8;
          const bool ok = key < L && qrow < L && (!causal || key <= qrow);
          const float p = ok ? exp2f(s[nt][e] * c - lse_r[e >> 1]) : 0.f;
          s[nt][e] = p * (dp[nt][e] - d_r[e >> 1]) * scale;  // dS (already times the softmax scale)
        }
      uint32_t dsa[4] = {pack_h2(s[0][0], s[0][1]), pack_h2(s[0][2], s[0][3]), pack_h2(s[1][0], s[1][1]),
                         pack_h2(s[1][2], s[1][3])};
      mma_nn_16x64(dq, dsa, smem_u32(sK), k0, lane);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qrow = q0 + g + r * 8;
      if (qrow >= L) continue;
      __half* o = dqkv + (static_cast<long long>(seq) * L + qrow) * ld + h * kHd;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<uint32_t*>(o + nt * 8 + 2 * t) = pack_h2(dq[nt][2 * r], dq[nt][2 * r + 1]);
    }
  }

  // ---------------- phase B: dK, dV
  for (int kb = warp; kb < nblk; kb += kBwdWarps) {
    const int k0 = kb * 16;
    if (k0 >= L) break;
    uint32_t ka[4][4], va[4][4];
    load_a_frags(ka, smem_u32(sK), k0, lane);
    load_a_frags(va, smem_u32(sV), k0, lane);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int q_begin = causal ? k0 : 0;  // queries before the key block never see it
    for (int q0 = q_begin; q0 < Lp; q0 += 16) {
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_nt_16x16(st, ka, smem_u32(sQ), q0, lane);    // S^T = K Q^T
      mma_nt_16x16(dpt, va, smem_u32(sdO), q0, lane);  // dP^T = V dO^T
      float pt[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qrow = q0 + nt * 8 + 2 * t + (e & 1);
          const int key = k0 + g + (e >> 1) * 8;
          const bool ok = key < L && qrow < L && (!causal || key <= qrow);
          const float p = ok ? exp2f(st[nt][e] * c - sLse[qrow]) : 0.f;
          pt[nt][e] = p;
          st[nt][e] = p * (dpt[nt][e] - sD[qrow]) * scale;
        }
      uint32_t pa[4] = {pack_h2(pt[0][0], pt[0][1]), pack_h2(pt[0][2], pt[0][3]), pack_h2(pt[1][0], pt[1][1]),
                        pack_h2(pt[1][2], pt[1][3])};
      uint32_t dsa[4] = {pack_h2(st[0][0], st[0][1]), pack_h2(st[0][2], st[0][3]), pack_h2(st[1][0], st[1][1]),
                         pack_h2(st[1][2], st[1][3])};
      mma_nn_16x64(dv, pa, smem_u32(sdO), q0, lane);  // dV += P^T dO
      mma_nn_16x64(dk, dsa, smem_u32(sQ), q0, lane);  // dK += dS^T Q
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = k0 + g + r * 8;
      if (key >= L) continue;
      __half* o = dqkv + (static_cast<long long>(seq) * L + key) * ld + h * kHd;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(o + d + nt * 8 + 2 * t) = pack_h2(dk[nt][2 * r], dk[nt][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(o + 2 * d + nt * 8 + 2 * t) = pack_h2(dv[nt][2 * r], dv[nt][2 * r + 1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Attention output of ONE query row per sequence (the class token).  In the LAST block of a ViT only the class
// token's row feeds ln_post and the projection (TPT/clip/model.py:235-238), so an inference forward needs the last
// block's attention, out_proj, ln_2 and MLP for that row alone; K and V still come from every token.  One warp per
// (sequence, head), fp32 throughout: lane j scores keys j, j+32, ...; then the lanes own two output dimensions each
// and walk the keys together.  Optionally copies the row of the fp32 residual stream to a compact [n_seq, d] buffer
// (the residual operand of the out_proj GEMM on those rows).
constexpr int kRowMaxT = 21;   // key slots per lane: L <= 672

__global__ void __launch_bounds__(128) attn_row_fwd_kernel(const __half* __restrict__ qkv, long long ld, int k_off, int v_off,
                                                           const __half* __restrict__ q_rows, int n_units, int L,
                                                           int heads, int q_row, __half* __restrict__ out,
                                                           const float* __restrict__ x, float* __restrict__ x_row) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * 4 + warp;
  if (u >= n_units) return;
  const int h = u % heads, seq = u / heads;
  const int d = heads * kHd;
  const __half* base = qkv + static_cast<long long>(seq) * L * ld + h * kHd;
  // the query row, whole, in every lane: from its own compact [n_seq, d] tensor when the caller projected only that row
  float2 q[32];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q_rows != nullptr ? q_rows + static_cast<long long>(seq) * d + h * kHd
                                                                       : base + q_row * ld);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const uint4 v = qp[c8];
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) q[c8 * 4 + e] = __half22float2(hv[e]);
    }
  }
  float s[kRowMaxT];
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < kRowMaxT; ++t) {
    const int j = t * 32 + lane;
    float acc = -INFINITY;
    if (t * 32 < L && j < L) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + j * ld + k_off);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 v = kp[c8];
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 kf = __half22float2(hv[e]);
          a0 = fmaf(q[c8 * 4 + e].x, kf.x, a0);
          a1 = fmaf(q[c8 * 4 + e].y, kf.y, a1);
        }
      }
      acc = a0 + a1;
    }
    s[t] = acc;
    m = fmaxf(m, acc);
  }
  m = warp_max(m);
  const float c = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
  float l = 0.f;
#pragma unroll
  for (int t = 0; t < kRowMaxT; ++t) {
    s[t] = (t * 32 + lane < L) ? exp2f((s[t] - m) * c) : 0.f;
    l += s[t];
  }
  l = warp_sum(l);
  float o0 = 0.f, o1 = 0.f;
  const __half* vbase = base + v_off + 2 * lane;
#pragma unroll
  for (int t = 0; t < kRowMaxT; ++t) {
    if (t * 32 < L) {
      const int n = min(32, L - t * 32);
      if (n == 32) {
#pragma unroll 8
        for (int jj = 0; jj < 32; ++jj) {
          const float pj = __shfl_sync(0xffffffffu, s[t], jj);
          const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(vbase + (t * 32 + jj) * ld));
          o0 = fmaf(pj, vf.x, o0);
          o1 = fmaf(pj, vf.y, o1);
        }
      } else {
        for (int jj = 0; jj < n; ++jj) {
          const float pj = __shfl_sync(0xffffffffu, s[t], jj);
          const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(vbase + (t * 32 + jj) * ld));
          o0 = fmaf(pj, vf.x, o0);
          o1 = fmaf(pj, vf.y, o1);
        }
      }
    }
  }
  const float inv = 1.f / l;
  *reinterpret_cast<__half2*>(out + static_cast<long long>(seq) * d + h * kHd + 2 * lane) = __floats2half2_rn(o0 * inv, o1 * inv);
  if (x_row != nullptr) {
    const float2 xv = *reinterpret_cast<const float2*>(x + (static_cast<long long>(seq) * L + q_row) * d + h * kHd + 2 * lane);
    *reinterpret_cast<float2*>(x_row + static_cast<long long>(seq) * d + h * kHd + 2 * lane) = xv;
  }
}

int attention_row_fwd(const __half* qkv, const __half* q_rows, int n_seq, int L, int heads, int q_row, __half* out,
                      const float* x, float* x_row, cudaStream_t stream) {
  if (n_seq <= 0 || L <= 0 || heads <= 0 || q_row < 0 || q_row >= L)
    return set_error(RLCF_ERR_ARG, "attention_row_fwd: bad shape");
  if (L > 32 * kRowMaxT) return set_error(RLCF_ERR_ARG, "attention_row_fwd: sequence %d longer than %d", L, 32 * kRowMaxT);
  if ((x == nullptr) != (x_row == nullptr)) return set_error(RLCF_ERR_ARG, "attention_row_fwd: x and x_row go together");
  if (((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(q_rows)) & 15) != 0)
    return set_error(RLCF_ERR_ARG, "attention_row_fwd: qkv / q_rows must be 16-byte aligned");
  const int n_units = n_seq * heads, d = heads * kHd;
  // q_rows given: qkv holds k | v only ([n_seq*L, 2d]); else the packed q | k | v of rlcf_attention_fwd
  const long long ld = q_rows != nullptr ? 2LL * d : 3LL * d;
  attn_row_fwd_kernel<<<(n_units + 3) / 4, 128, 0, stream>>>(qkv, ld, q_rows != nullptr ? 0 : d, q_rows != nullptr ? d : 2 * d,
                                                             q_rows, n_units, L, heads, q_row, out, x, x_row);
  RLCF_CHECK_LAUNCH("attention_row_fwd");
  return 0;
}

int attention_fwd_tc(const __half* qkv, int n_seq, int L, int heads, int causal, __half* out, float* lse,
                     cudaStream_t stream);
int attention_fwd_tc8(const __half* qkv, int n_seq, int L, int heads, int causal, __half* out, float* lse,
                      cudaStream_t stream);
int attention_fwd_tcl(const __half* qkv, int n_seq, int L, int heads, int causal, __half* out, float* lse,
                      cudaStream_t stream);

int attention_fwd(const __half* qkv, int n_seq, int L, int heads, int causal, __half* out, float* lse,
                  cudaStream_t stream) {
  if (n_seq <= 0 || L <= 0 || heads <= 0) return set_error(RLCF_ERR_ARG, "attention_fwd: bad shape");
  if (attention_impl() == 0) {  // tcgen05 kernels; -1 = shape not covered -> next kernel
    // eight softmax warps per team (attention_tc8.cu, L <= 208: 217 vs 243 us at 512 x 197 x 12) where the sequence
    // fits (200 us with its token), else four (attention_tc.cu, L <= 272); RLCF_ATTN_TC8=0 keeps the four-warp kernel everywhere
    static const int tc8 = getenv("RLCF_ATTN_TC8") != nullptr ? atoi(getenv("RLCF_ATTN_TC8")) : 1;
    // key-block kernel (attention_tcl.cu) for what the two above do not cover: 272 < L <= 640, not causal (ViT-L/14@336px:
    // 577 tokens).  RLCF_ATTN_TCL=0 leaves those to the mma.sync kernel, =2 tries it FIRST for every shape (A/B probe)
    static const int tcl = getenv("RLCF_ATTN_TCL") != nullptr ? atoi(getenv("RLCF_ATTN_TCL")) : 1;
    if (tcl == 2) {
      const int rcl = attention_fwd_tcl(qkv, n_seq, L, heads, causal, out, lse, stream);
      if (rcl != -1) return rcl;
    }
    if (tc8) {
      const int rc8 = attention_fwd_tc8(qkv, n_seq, L, heads, causal, out, lse, stream);
      if (rc8 != -1) return rc8;
    }
    const int rc = attention_fwd_tc(qkv, n_seq, L, heads, causal, out, lse, stream);
    if (rc != -1) return rc;
    if (tcl == 1) {
      const int rcl = attention_fwd_tcl(qkv, n_seq, L, heads, causal, out, lse, stream);
      if (rcl != -1) return rcl;
    }
  }
  const int Lp = (L + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>(2 * Lp + kFwdWarps * 16) * kRowBytes;
  if (smem > 227 * 1024) return set_error(RLCF_ERR_ARG, "attention_fwd: sequence %d too long for one CTA", L);
  static DynSmemState st;
  if (cudaError_t e = ensure_dyn_smem(attn_fwd_kernel, smem, st))
    return set_error(RLCF_ERR_CUDA, "attention_fwd attr: %s", cudaGetErrorString(e));
  dim3 grid((L + kFwdWarps * 16 - 1) / (kFwdWarps * 16), heads, n_seq);
  attn_fwd_kernel<<<grid, kFwdWarps * 32, smem, stream>>>(qkv, L, Lp, heads, causal, out, lse);
  RLCF_CHECK_LAUNCH("attention_fwd");
  return 0;
}

int attention_bwd_tc(const __half* qkv, const __half* out, const __half* dout, const float* lse, int n_seq, int L,
                     int heads, int causal, __half* dqkv, cudaStream_t stream);

int attention_bwd(const __half* qkv, const __half* out, const __half* dout, const float* lse, int n_seq, int L,
                  int heads, int causal, __half* dqkv, cudaStream_t stream) {
  if (n_seq <= 0 || L <= 0 || heads <= 0 || lse == nullptr) return set_error(RLCF_ERR_ARG, "attention_bwd: bad args");
  // tcgen05 kernel for every sequence it covers (L <= 384); rlcf_set_attention_impl(1) /
  // RLCF_ATTN_IMPL=1 forces the warp-MMA kernels (forward and backward)
  if (attention_impl() == 0) {
    const int rc = attention_bwd_tc(qkv, out, dout, lse, n_seq, L, heads, causal, dqkv, stream);
    if (rc != -1) return rc;
  }
  const int Lp = (L + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>(4 * Lp) * kRowBytes + 2 * Lp * sizeof(float);
  if (smem > 227 * 1024) return set_error(RLCF_ERR_ARG, "attention_bwd: sequence %d too long for one CTA", L);
  static DynSmemState st;
  if (cudaError_t e = ensure_dyn_smem(attn_bwd_kernel, smem, st))
    return set_error(RLCF_ERR_CUDA, "attention_bwd attr: %s", cudaGetErrorString(e));
  dim3 grid(heads, n_seq);
  attn_bwd_kernel<<<grid, kBwdWarps * 32, smem, stream>>>(qkv, out, dout, lse, L, Lp, heads, causal, dqkv);
  RLCF_CHECK_LAUNCH("attention_bwd");
  return 0;
}

}  // namespace rlcf
