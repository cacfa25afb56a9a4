// Persistent tcgen05 fused attention forward for CLIP towers (head_dim 64, L <= 272 tokens).
//
// Work unit ("tile") = 128 query rows of one (sequence, head).  The whole key/value range of the sequence fits in
// shared memory (L <= 257 for every configured tower), so there is no KV loop and no online-softmax rescaling:
//   TMA     : Q tile [128 x 64], K and V tiles [Lk x 64] (Lk = L rounded up to 16), 128-byte swizzle
//   UMMA #1 : S[128 x min(Lk,256)] = Q K^T   (SS, both K-major; fp32 accumulator in TMEM)
//   softmax : one thread per query row reads its S row from TMEM (tcgen05.ld), max / exp2 / sum in fp32, and writes
//             P as packed fp16 back into the TMEM columns it has already consumed (tcgen05.st); the at most 16 keys
//             beyond column 256 (key 256 of ViT-L/14's 257 tokens) are scored on CUDA cores so that a tile never
//             needs more than 256 TMEM columns
//   UMMA #2 : O[128 x 64] = P V              (TS: A = P from TMEM, B = V from smem as an MN-major operand)
//   epilogue: O / rowsum -> fp16 -> global (128 contiguous bytes per row); optional log-sum-exp for the backward
//
// One CTA per SM runs two independent "teams"; each team owns one shared-memory stage, 256 TMEM columns, a TMA
// thread, an MMA-issuing thread and four softmax warps, and walks every second tile of the CTA's tile list.  While one
// team is in its softmax, the other team's loads and MMAs proceed, so launch, allocation and load latencies are paid
// once per CTA instead of once per tile.  Replaces the bmm-softmax-bmm of nn.MultiheadAttention
// (TPT/clip/model.py:185-187).
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

struct AttnTcArgs {
  int L, Lk, heads, causal;
  int kv_box_rows, n_kv_boxes;  // TMA boxes covering the Lk key rows
  int n_mma;                    // UMMA N of S = min(Lk, 256); keys [n_mma, L) are scored on CUDA cores
  int o_off;                    // TMEM column (inside the team's 256) of the O accumulator
  int n_qt, n_units;            // query tiles per (sequence, head); number of (sequence, head) units
  int stage_bytes;
  int q_start;                  // first query row the 128-row tiles cover (1 when row 0 is the producer warp's row job)
  int row_job;                  // 1: L - 1 is a multiple of 128 (257 tokens): the tiles cover rows [1, L) exactly and the
                                // team's otherwise idle TMA warp computes row 0 on the CUDA cores from the K / V tiles in
                                // shared memory -- instead of a third tile with one live row out of 128
  int sum_mma;                  // 1: the row sums come out of the P V MMA (a block of ones appended to V: O gets 80 columns)
  int debug;                    // RLCF_ATTN_DEBUG bit mask (timing probes only): 1 skip max pass, 2 skip exp pass, 4 skip stores, 8 timeline, 16 no exp token, 32 per-thread O stores, 128 row sums on the CUDA cores; row job: 256 protocol only, 512 no P V, 1024 print its clock64 duration, 2048 release K / V before it
  const __half* qkv;
  __half* out;
  float* lse;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// MN-major, 128-byte-swizzled B operand that is wider than one 64-element atom along N: the atoms of the second
// 64-column block start `lbo_bytes` after the first block's (leading-dimension byte offset field, bits [16,30)).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

constexpr int kAttnThreads = 384;  // warps 0/3: TMA (team 0/1), warps 1/2: MMA (team 0/1), warps 4-7 / 8-11: softmax
constexpr int kMaxExtraKeys = 16;

// barrier indices inside a team's block of 10
enum { B_KVFULL = 0, B_KVFREE = 1, B_QFULL = 2 /*+buf*/, B_QFREE = 4 /*+buf*/, B_SREADY = 6, B_PREADY = 7, B_OREADY = 8,
       B_TMEMFREE = 9, B_PER_TEAM = 10 };

// max over the visible keys of one 32-column S chunk (two independent chains of 3-input maxima)
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], float m, int ch, bool full, int key_end) {
  float m2 = -INFINITY;
  if (full) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      m = fmaxf(m, __uint_as_float(v[j]));
      m2 = fmaxf(m2, __uint_as_float(v[16 + j]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      m = fmaxf(m, (ch * 32 + j < key_end) ? __uint_as_float(v[j]) : -INFINITY);
      m2 = fmaxf(m2, (ch * 32 + 16 + j < key_end) ? __uint_as_float(v[16 + j]) : -INFINITY);
    }
  }
  return fmaxf(m, m2);
}

// p = 2^(s c - m c) for one chunk; writes the packed fp16 P chunk to TMEM columns [16 ch, 16 ch + 16) -- S columns
// this row has already consumed -- and returns the chunk's row-sum contribution.  (Evaluating a fraction of the
// exponentials with a polynomial on the FMA pipe, as FlashAttention-4 does, was measured 8-14 % SLOWER here: the pass
// is paced by TMEM reads and per-warp issue latency, not by the MUFU lanes -- profiles/r1_attention_probes.txt.)
// kSum = false: the row sum is produced by the P V MMA (AttnTcArgs::sum_mma), the 32 additions per chunk disappear from
// this issue-latency-bound loop.
template <bool kSum>
__device__ __forceinline__ float chunk_exp(const uint32_t (&v)[32], uint32_t trow, int ch, bool full, int key_end,
                                           float c, float mc) {
  uint32_t pk[16];
  float l0 = 0.f, l1 = 0.f;
  if (full) {   // separate straight-line path: predicated-off masking instructions would still take issue slots
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = ex2_approx(fmaf(__uint_as_float(v[2 * j]), c, -mc));
      const float b = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), c, -mc));
      if constexpr (kSum) { if (j & 1) l1 += a + b; else l0 += a + b; }
      const __half2 hp = __floats2half2_rn(a, b);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hp);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = (ch * 32 + 2 * j < key_end) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), c, -mc)) : 0.f;
      const float b = (ch * 32 + 2 * j + 1 < key_end) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), c, -mc)) : 0.f;
      if constexpr (kSum) { if (j & 1) l1 += a + b; else l0 += a + b; }
      const __half2 hp = __floats2half2_rn(a, b);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hp);
    }
  }
  tmem_st_32x16(trow + ch * 16, pk);
  return l0 + l1;
}

// ---- query row 0 of one (sequence, head) by ONE warp, from the 128-byte-swizzled K / V tiles in shared memory.
// Warp-level tensor-core MMAs (mma.sync m16n8k16; the 16-row A operand carries the query in row 0 and zeros below it):
// a scalar fp32 version spent its time converting K and V to fp32 -- 1 150 half2 conversions per unit at a quarter of
// the FMA rate made the row take longer than the two tcgen05 tiles it rides along with (profiles/r2_attention_probes.txt).
__device__ __forceinline__ uint32_t r0_sw(int row, int chunk) {       // TMA's SWIZZLE_128B: 16-byte chunk ^ (row & 7)
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void r0_ldsm(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void r0_ldsm_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void r0_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kRow0Blocks = (256 + kMaxExtraKeys) / 16;   // 16-key blocks of the K / V tiles

// kFull: the tiles hold exactly kRow0Blocks blocks (257 tokens) -- no per-block branch, so the blocks' MMAs interleave
template <bool kFull>
__device__ __noinline__ void row0_attention(uint32_t sK, uint32_t sV, const __half* qg, int L, int Lk, int lane,
                                            __half* out_row, float* lse_row, bool skip_pv) {
  const int g = lane >> 2, t = lane & 3;
  // A fragments of [q; 0; ...; 0] (16 x 64): row g of the fragment lives in lanes 4g .. 4g+3, so only lanes 0-3 load
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    qa[ks][0] = g == 0 ? *reinterpret_cast<const uint32_t*>(qg + ks * 16 + 2 * t) : 0u;
    qa[ks][2] = g == 0 ? *reinterpret_cast<const uint32_t*>(qg + ks * 16 + 8 + 2 * t) : 0u;
    qa[ks][1] = qa[ks][3] = 0u;
  }
  // scores of row g against keys kb*16 + {2t, 2t+1, 8+2t, 9+2t}: every block's MMAs are independent of the others'
  float s[kRow0Blocks][4];
  float m = -INFINITY;
#pragma unroll
  for (int kb = 0; kb < kRow0Blocks; ++kb) {
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
    if (kFull || kb * 16 < Lk) {
      const int r = kb * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b[4];
        r0_ldsm(b, sK + r0_sw(r, ks * 2 + ((lane >> 3) & 1)));
        r0_mma(c0, qa[ks], b[0], b[1]);
        r0_mma(c1, qa[ks], b[2], b[3]);
      }
    }
    const int key = kb * 16 + 2 * t;
    s[kb][0] = key < L ? c0[0] : -INFINITY;
    s[kb][1] = key + 1 < L ? c0[1] : -INFINITY;
    s[kb][2] = key + 8 < L ? c1[0] : -INFINITY;
    s[kb][3] = key + 9 < L ? c1[1] : -INFINITY;
    m = fmaxf(fmaxf(m, fmaxf(s[kb][0], s[kb][1])), fmaxf(s[kb][2], s[kb][3]));
  }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  const float c = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
  const float mc = m * c;
  float l = 0.f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kb = 0; kb < kRow0Blocks; ++kb) {
    if (kFull || kb * 16 < Lk) {
      float pv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pv[e] = g == 0 ? ex2_approx(fmaf(s[kb][e], c, -mc)) : 0.f;      // masked keys: 2^-inf = 0
        l += pv[e];
      }
      if (!skip_pv) {
        __half2 h01 = __floats2half2_rn(pv[0], pv[1]), h23 = __floats2half2_rn(pv[2], pv[3]);
        const uint32_t pa[4] = {*reinterpret_cast<uint32_t*>(&h01), 0u, *reinterpret_cast<uint32_t*>(&h23), 0u};
        const int r = kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          r0_ldsm_t(b, sV + r0_sw(r, np * 2 + (lane >> 4)));
          r0_mma(o[2 * np], pa, b[0], b[1]);
          r0_mma(o[2 * np + 1], pa, b[2], b[3]);
        }
      }
    }
  }
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  if (g == 0) {
    const float inv = 1.f / l;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<__half2*>(out_row + nt * 8 + 2 * t) = __floats2half2_rn(o[nt][0] * inv, o[nt][1] * inv);
    if (t == 0 && lse_row != nullptr) *lse_row = m * 0.125f + logf(l);
  }
}

template <bool kSumMma>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV,
                   const __grid_constant__ CUtensorMap mapO, AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // sum_mma: one constant [Lk x 64] tile (128-byte rows, swizzled like V) whose column 0 is 1 and the rest 0, shared by
  // both teams: appended to V as a second 64-column block of the P V MMA's B operand, it makes O column 64 the row sum
  uint8_t* sOnes = smem + 2 * p.stage_bytes;
  const int ones_bytes = kSumMma ? p.Lk * 128 : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ones_bytes);        // [2 teams][B_PER_TEAM]
  uint64_t* tok = bars + 2 * B_PER_TEAM;                                   // [4 lane quarters][2 teams]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.heads * 64;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapKV);
    tma_prefetch_desc(&mapO);
    for (int t = 0; t < 2; ++t) {
      uint64_t* b = bars + t * B_PER_TEAM;
      for (int i = 0; i < B_PER_TEAM; ++i) mbar_init(&b[i], 1);
      mbar_init(&b[B_PREADY], 4);    // one arrive per softmax warp
      mbar_init(&b[B_TMEMFREE], 4);
      if (p.row_job) mbar_init(&b[B_KVFREE], 2);   // the MMA commit and the producer warp's row job both read K / V
      if (!(p.debug & 32)) {         // the O tile is staged in the Q buffer: its TMA store must have read it too
        mbar_init(&b[B_QFREE], 5);
        mbar_init(&b[B_QFREE + 1], 5);
      }
    }
    for (int i = 0; i < 8; ++i) mbar_init(&tok[i], 1);
    fence_barrier_init();
  }
  if constexpr (kSumMma) {
    for (int i = tid; i < p.Lk * 8; i += kAttnThreads) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // element (row r, column 0) lives in the row's logical 16-byte chunk 0 = physical chunk (0 ^ (r & 7))
    for (int r = tid; r < p.Lk; r += kAttnThreads)
      *reinterpret_cast<__half*>(sOnes + r * 128 + ((r & 7) << 4)) = __float2half(1.f);
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // role of this warp: team and kind
  const int team = warp >= 4 ? (warp - 4) >> 2 : (warp == 0 || warp == 1 ? 0 : 1);
  uint64_t* tb = bars + team * B_PER_TEAM;
  uint8_t* sQ = smem + team * p.stage_bytes;     // two Q buffers of 16 KB
  uint8_t* sK = sQ + 2 * 128 * 128;
  uint8_t* sV = sK + p.Lk * 128;
  const uint32_t tmem = tmem_base + team * 256;
  // units (sequence, head) of this CTA: u = blockIdx.x + i * gridDim.x; team t takes i = t, t + 2, ...
  // K and V of a unit stay in shared memory while the team walks the unit's n_qt query tiles.
  const int first = blockIdx.x + team * gridDim.x, stride = 2 * gridDim.x;

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------ TMA producer of this team (+ the row job)
    uint32_t uc = 0, tc = 0;
    for (int u = first; u < p.n_units; u += stride, ++uc) {
      const int h = u % p.heads, seq = u / p.heads;
      const int row_base = seq * p.L;
      if (lane == 0) {
        mbar_wait(&tb[B_KVFREE], (uc & 1) ^ 1);
        mbar_expect_tx(&tb[B_KVFULL], 2 * p.Lk * 128);
        for (int b = 0; b < p.n_kv_boxes; ++b) {
          tma_load_2d(sK + b * p.kv_box_rows * 128, &mapKV, &tb[B_KVFULL], d + h * 64, row_base + b * p.kv_box_rows);
          tma_load_2d(sV + b * p.kv_box_rows * 128, &mapKV, &tb[B_KVFULL], 2 * d + h * 64, row_base + b * p.kv_box_rows);
        }
        for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
          const int buf = tc & 1;
          mbar_wait(&tb[B_QFREE + buf], ((tc >> 1) & 1) ^ 1);
          mbar_expect_tx(&tb[B_QFULL + buf], 128 * 128);
          tma_load_2d(sQ + buf * 128 * 128, &mapQ, &tb[B_QFULL + buf], h * 64, row_base + p.q_start + qt * 128);
        }
      }
      if (p.row_job) {
        // query row 0 of this unit, whole warp, fp32: lane j scores keys j, j + 32, ... from the swizzled K tile, then
        // every lane owns two output dimensions and the warp walks the rows of the V tile together
        __syncwarp();
        mbar_wait(&tb[B_KVFULL], uc & 1);
        if (p.debug & 256) {            // timing probe: barrier protocol only, row 0 is NOT computed
          __syncwarp();
          if (lane == 0) mbar_arrive(&tb[B_KVFREE]);
          continue;
        }
        const __half* qg = reinterpret_cast<const __half*>(p.qkv) + static_cast<size_t>(row_base) * 3 * d + h * 64;
        __half* out_row = p.out + static_cast<size_t>(row_base) * d + h * 64;
        float* lse_row = (p.lse != nullptr && !(p.debug & 8)) ? p.lse + (static_cast<size_t>(seq) * p.heads + h) * p.L : nullptr;
        if ((p.debug & 2048) && lane == 0) mbar_arrive(&tb[B_KVFREE]);   // timing probe: release K / V BEFORE the row job (results undefined)
        const long long t_row0 = clock64();
        if (p.Lk == kRow0Blocks * 16)
          row0_attention<true>(smem_u32(sK), smem_u32(sV), qg, p.L, p.Lk, lane, out_row, lse_row, (p.debug & 512) != 0);
        else
          row0_attention<false>(smem_u32(sK), smem_u32(sV), qg, p.L, p.Lk, lane, out_row, lse_row, (p.debug & 512) != 0);
        __syncwarp();
        if ((p.debug & 1024) && blockIdx.x == 0 && lane == 0 && uc >= 2 && uc < 6)
          printf("team %d unit %u: row job %lld cycles, started %lld\n", team, uc, clock64() - t_row0, t_row0);
        if (lane == 0 && !(p.debug & 2048)) mbar_arrive(&tb[B_KVFREE]);     // this warp is done with the unit's K / V
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------ MMA issuer of this team
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(128, p.n_mma);
      const uint32_t idesc_o = umma_idesc_f16(128, kSumMma ? 80 : 64) | (1u << 16);  // B (= V [| ones]) is MN-major
      const uint64_t dk = umma_desc_k_sw128(smem_u32(sK));
      const uint64_t dv = kSumMma ? umma_desc_mn_sw128(smem_u32(sV), smem_u32(sOnes) - smem_u32(sV))
                                    : umma_desc_k_sw128(smem_u32(sV));
      const int ksteps = p.Lk >> 4;
      uint32_t uc = 0, tc = 0;
      for (int u = first; u < p.n_units; u += stride, ++uc) {
        mbar_wait(&tb[B_KVFULL], uc & 1);
        for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
          const int buf = tc & 1;
          const uint32_t ph = tc & 1;
          const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ + buf * 128 * 128));
          mbar_wait(&tb[B_QFULL + buf], (tc >> 1) & 1);
          mbar_wait(&tb[B_TMEMFREE], ph ^ 1);  // the previous tile's O has been read out of TMEM
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss<1>(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(&tb[B_SREADY]);
          mbar_wait(&tb[B_PREADY], ph);        // softmax has written P (and is done with sQ / sK)
          tc_fence_after();
          for (int j = 0; j < ksteps; ++j) umma_f16_ts(tmem + p.o_off, tmem + 8 * j, dv + 128 * j, idesc_o, j != 0);
          umma_commit(&tb[B_OREADY]);
          umma_commit(&tb[B_QFREE + buf]);
          if (qt == p.n_qt - 1) umma_commit(&tb[B_KVFREE]);  // all readers of this unit's K/V have retired
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue: thread = query row
    const int r = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const int n_chunks = (p.n_mma + 31) >> 5;
    const int n_extra = p.L - p.n_mma;              // keys scored on CUDA cores (<= kMaxExtraKeys), usually <= 0
    // Left alone the two teams fall into lockstep (same phase within 200 cycles: both in the exp pass on the same SM
    // sub-partitions, then both waiting for the tensor core).  A token per lane quarter lets only one of the two warps
    // that share a sub-partition into its exp pass at a time, strictly alternating, which shifts the teams by one exp
    // pass: one team's exponentials overlap the other team's MMAs, O read-out and loads (-10 % kernel time; one token
    // per team instead of per quarter measures the same).  Every warp takes the token once per tile (dead warps pass
    // it on) and the team with fewer tiles keeps passing it until the other one is done.
    const bool use_tok = !(p.debug & 16);
    const bool tma_out = !(p.debug & 32);
    uint64_t* tok_mine = &tok[(warp & 3) * 2 + team];
    uint64_t* tok_other = &tok[(warp & 3) * 2 + (team ^ 1)];
    auto units_of = [&](int t) {
      const int f = static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x);
      return f < p.n_units ? (p.n_units - f + stride - 1) / stride : 0;
    };
    const uint32_t n_tok = static_cast<uint32_t>(max(units_of(0), units_of(1)) * p.n_qt);
    uint32_t tc = 0;
    for (int u = first; u < p.n_units; u += stride) {
      const int h = u % p.heads, seq = u / p.heads;
      for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
        const uint32_t ph = tc & 1;
        const uint8_t* sQb = sQ + (tc & 1) * 128 * 128;
        const int q0 = p.q_start + qt * 128, qrow = q0 + r;
        const int key_end = p.causal ? min(p.L, qrow + 1) : p.L;  // keys [0, key_end) are visible to this row
        const bool warp_live = q0 + (warp & 3) * 32 < p.L;         // rows of a dead warp are never stored
        // chunks below `full_chunks` are visible to every row of this warp: no masking needed there
        const int warp_min_end = p.causal ? min(p.L, q0 + (warp & 3) * 32 + 1) : p.L;
        const int full_chunks = min(n_chunks, warp_min_end >> 5);
        // RLCF_ATTN_DEBUG & 8: timeline probe -- the first softmax warp of each team of CTA 0 writes clock64 stamps of
        // its first 64 tiles into the (oversized) lse buffer: [team][tile][8] = {start, S ready, max done, P written,
        // O ready, stored}
        const bool probe = (p.debug & 8) && blockIdx.x == 0 && (warp & 3) == 0 && lane == 0 && tc < 64 && p.lse != nullptr;
        long long* stamp = reinterpret_cast<long long*>(p.lse) + (team * 64 + (tc & 63)) * 8;
        if (probe) stamp[0] = clock64();
        mbar_wait(&tb[B_SREADY], ph);
        tc_fence_after();
        if (probe) stamp[1] = clock64();
        if (tma_out && tc > 0 && lane == 0) {   // the previous tile's O store has left its Q buffer
          bulk_wait_read_all();
          mbar_arrive(&tb[B_QFREE + ((tc - 1) & 1)]);
        }
        float m = -INFINITY, l = 0.f;
        bool tok_held = false;
        float sx[kMaxExtraKeys];
        if (warp_live) {
          if (n_extra > 0) {
            // scores of keys >= 256 from shared memory: q row r and key rows n_mma.. (both 128-byte swizzled rows)
            const uint4* qr = reinterpret_cast<const uint4*>(sQb + r * 128);
#pragma unroll
            for (int e = 0; e < kMaxExtraKeys; ++e) {
              float acc = 0.f;
              if (e < n_extra) {
                const int kr = p.n_mma + e;
                const uint4* kk = reinterpret_cast<const uint4*>(sK + kr * 128);
#pragma unroll
                for (int ck = 0; ck < 8; ++ck) {
                  const uint4 a = qr[ck ^ (r & 7)], b = kk[ck ^ (kr & 7)];
                  const __half2* ha = reinterpret_cast<const __half2*>(&a);
                  const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
                  for (int x = 0; x < 4; ++x) {
                    const float2 fa = __half22float2(ha[x]), fb = __half22float2(hb[x]);
                    acc = fmaf(fa.x, fb.x, fmaf(fa.y, fb.y, acc));
                  }
                }
                if (kr < key_end) m = fmaxf(m, acc);
              }
              sx[e] = acc;
            }
          }
          // ---- pass 1: row maximum (the TMEM load of the next chunk is in flight while a chunk is reduced)
          if (!(p.debug & 1)) {
            uint32_t va[32], vb[32];
            tmem_ld_32x32(trow, va);
            for (int ch = 0; ch < n_chunks; ch += 2) {
              tmem_ld_wait();
              if (ch + 1 < n_chunks) tmem_ld_32x32(trow + (ch + 1) * 32, vb);
              m = chunk_max(va, m, ch, ch < full_chunks, key_end);
              if (ch + 1 < n_chunks) {
                tmem_ld_wait();
                if (ch + 2 < n_chunks) tmem_ld_32x32(trow + (ch + 2) * 32, va);
                m = chunk_max(vb, m, ch + 1, ch + 1 < full_chunks, key_end);
              }
            }
          } else {
            m = 0.f;
          }
          // ---- pass 2: p = 2^(s*c - m*c), row sum, packed fp16 P back into consumed S columns
          const float mc = m * c;
          if (probe) stamp[2] = clock64();
          if (use_tok) {
            mbar_wait(tok_mine, team == 0 ? ((tc & 1) ^ 1) : (tc & 1));
            tok_held = true;
          }
          if (!(p.debug & 2)) {
            uint32_t va[32], vb[32];
            tmem_ld_32x32(trow, va);
            for (int ch = 0; ch < n_chunks; ch += 2) {
              tmem_ld_wait();
              if (ch + 1 < n_chunks) tmem_ld_32x32(trow + (ch + 1) * 32, vb);
              l += chunk_exp<!kSumMma>(va, trow, ch, ch < full_chunks, key_end, c, mc);
              if (ch + 1 < n_chunks) {
                tmem_ld_wait();
                if (ch + 2 < n_chunks) tmem_ld_32x32(trow + (ch + 2) * 32, va);
                l += chunk_exp<!kSumMma>(vb, trow, ch + 1, ch + 1 < full_chunks, key_end, c, mc);
              }
            }
          }
          if (p.Lk > p.n_mma) {  // P of the keys beyond column 256: 16 keys = 8 packed columns at [n_mma/2, n_mma/2 + 8)
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k0 = p.n_mma + 2 * j;
              const float a = (k0 < key_end) ? ex2_approx(fmaf(sx[2 * j], c, -mc)) : 0.f;
              const float b = (k0 + 1 < key_end) ? ex2_approx(fmaf(sx[2 * j + 1], c, -mc)) : 0.f;
              l += a + b;
              const __half2 hp = __floats2half2_rn(a, b);
              pk[j] = *reinterpret_cast<const uint32_t*>(&hp);
            }
            tmem_st_32x8(trow + (p.n_mma >> 1), pk);
          }
          tmem_st_wait();
        }
        if (use_tok && !tok_held) mbar_wait(tok_mine, team == 0 ? ((tc & 1) ^ 1) : (tc & 1));  // dead warp: pass it on
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (use_tok) mbar_arrive(tok_other);
          mbar_arrive(&tb[B_PREADY]);
        }
        if (probe) stamp[3] = clock64();
        // ---- epilogue
        mbar_wait(&tb[B_OREADY], ph);
        tc_fence_after();
        if (probe) stamp[4] = clock64();
        if (warp_live) {
          uint32_t o[64];
          tmem_ld_32x32(trow + p.o_off, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
          tmem_ld_32x32(trow + p.o_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
          if constexpr (kSumMma) {          // O column 64 = sum_j P_j * 1: the row sum of the fp16 P the MMA multiplied
            uint32_t lsum[16];
            tmem_ld_32x16(trow + p.o_off + 64, lsum);
            tmem_ld_wait();
            l = __uint_as_float(lsum[0]);
          }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tb[B_TMEMFREE]);
          const float inv = 1.f / l;
          uint4 ov[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __half2 h0 = __floats2half2_rn(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
            __half2 h1 = __floats2half2_rn(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
            __half2 h2 = __floats2half2_rn(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
            __half2 h3 = __floats2half2_rn(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
            ov[j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                               *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          }
          if (tma_out) {
            // Stage this warp's 32 x 64 O block in its rows of the tile's Q buffer (S = Q K^T has retired; the buffer
            // is handed back to the producer only after this store, see B_QFREE), 128-byte swizzled like a TMA box,
            // and let one TMA store write it: full 128-byte rows, rows >= L clipped by the tensor map.  One thread
            // per row storing its own 128 bytes cost 8 fully scattered store instructions per warp, which backed the
            // LSU up for ~2000 cycles per tile.
            if (!(p.debug & 4)) {
              uint4* srow = reinterpret_cast<uint4*>(const_cast<uint8_t*>(sQb) + r * 128);
#pragma unroll
              for (int j = 0; j < 8; ++j) srow[j ^ (r & 7)] = ov[j];
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&mapO, sQb + (warp & 3) * 32 * 128, h * 64, q0 + (warp & 3) * 32, seq);
                bulk_commit_group();
              }
            }
          } else if (qrow < p.L && !(p.debug & 4)) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(seq) * p.L + qrow) * d + h * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = ov[j];
          }
          if (qrow < p.L && !(p.debug & 4) && p.lse != nullptr && !(p.debug & 8))
            p.lse[(static_cast<size_t>(seq) * p.heads + h) * p.L + qrow] = m * 0.125f + logf(l);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tb[B_TMEMFREE]);
        }
        if (probe) stamp[5] = clock64();
      }
    }
    if (tma_out && lane == 0) bulk_wait_read_all();  // shared memory must outlive the last O store's read
    if (use_tok) {
      for (; tc < n_tok; ++tc) {                     // keep the other team's token moving
        mbar_wait(tok_mine, team == 0 ? ((tc & 1) ^ 1) : (tc & 1));
        __syncwarp();
        if (lane == 0) mbar_arrive(tok_other);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

static int make_tmap_rows64(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows) {
  static PFN_encodeTiled encode = get_encode_tiled();
  if (encode == nullptr) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled(attention) failed (%d)", static_cast<int>(r));
  return 0;
}

// Returns -1 when the shape is outside what this kernel covers (caller falls back to the mma.sync kernel).
int attention_fwd_tc(const __half* qkv, int n_seq, int L, int heads, int causal, __half* out, float* lse,
                     cudaStream_t stream) {
  const int Lk = (L + 15) / 16 * 16;
  if (Lk > 256 + kMaxExtraKeys || (reinterpret_cast<uintptr_t>(qkv) & 15) != 0) return -1;
  AttnTcArgs a{};
  a.L = L; a.Lk = Lk; a.heads = heads; a.causal = causal; a.out = out; a.lse = lse;
  a.n_mma = Lk < 256 ? Lk : 256;
  if (Lk <= 256) { a.kv_box_rows = Lk; a.n_kv_boxes = 1; }
  else { a.kv_box_rows = Lk / 2; a.n_kv_boxes = 2; }
  a.o_off = ((Lk / 2) + 31) / 32 * 32;   // behind P (Lk/2 packed columns), <= 160, so O ends at <= 224 < 256
  a.qkv = qkv;
  // 257 tokens = 2 x 128 + 1: the tiles take rows [1, L), the producer warp computes row 0 (RLCF_ATTN_ROWJOB=0: three tiles)
  static const int row_job = getenv("RLCF_ATTN_ROWJOB") != nullptr ? atoi(getenv("RLCF_ATTN_ROWJOB")) : 1;
  a.row_job = row_job && !causal && L > 128 && (L - 1) % 128 == 0;
  a.q_start = a.row_job ? 1 : 0;
  a.n_qt = (L - a.q_start + 127) / 128;
  a.n_units = heads * n_seq;
  a.stage_bytes = 2 * 128 * 128 + 2 * Lk * 128;
  static const int debug = getenv("RLCF_ATTN_DEBUG") ? atoi(getenv("RLCF_ATTN_DEBUG")) : 0;
  a.debug = debug;
  // Row sums from the tensor core (a constant tile of ones as a second B block of the P V MMA): needs Lk * 128 more bytes
  // of shared memory, room for 80 accumulator columns behind P, and every key inside the MMA (no CUDA-core extras).
  // RLCF_ATTN_DEBUG & 128 keeps the sums on the CUDA cores (A/B probe).
  const size_t smem_base = 1024 + 2 * static_cast<size_t>(a.stage_bytes) + (2 * B_PER_TEAM + 8) * 8 + 16;
  a.sum_mma = !(debug & 128) && Lk <= 256 && a.o_off + 80 <= 256 &&
              smem_base + static_cast<size_t>(Lk) * 128 <= 227 * 1024;
  const size_t smem = smem_base + (a.sum_mma ? static_cast<size_t>(Lk) * 128 : 0);
  auto kernel = a.sum_mma ? attn_fwd_tc_kernel<true> : attn_fwd_tc_kernel<false>;
  static DynSmemState st[2];    // one per kernel variant
  if (cudaError_t e = ensure_dyn_smem(kernel, smem, st[a.sum_mma]))
    return set_error(RLCF_ERR_CUDA, "attention_fwd_tc attr: %s", cudaGetErrorString(e));
  CUtensorMap mq, mkv;
  const long long rows = static_cast<long long>(n_seq) * L;
  if (int rc = make_tmap_rows64(&mq, qkv, rows, 3 * heads * 64, 128)) return rc;
  if (int rc = make_tmap_rows64(&mkv, qkv, rows, 3 * heads * 64, a.kv_box_rows)) return rc;
  // out as (column, token, sequence): a 64 x 32 box per softmax warp, tokens >= L clipped
  CUtensorMap mo;
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return -1;
  {
    static PFN_encodeTiled encode = get_encode_tiled();
    if (encode == nullptr) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
    const cuuint64_t dm = static_cast<cuuint64_t>(heads) * 64;
    cuuint64_t gdim[3] = {dm, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(n_seq)};
    cuuint64_t gstride[2] = {dm * 2, dm * 2 * static_cast<cuuint64_t>(L)};
    cuuint32_t box[3] = {64u, 32u, 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled(attention out) failed (%d)", static_cast<int>(r));
  }
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  kernel<<<grid, kAttnThreads, smem, stream>>>(mq, mkv, mo, a);
  RLCF_CHECK_LAUNCH("attention_fwd_tc");
  return 0;
}

}  // namespace rlcf
