// tcgen05 / TMEM backward of the fused attention core for CLIP towers (head_dim 64, L <= 384 tokens): dQ, dK, dV of
// one (sequence, head) unit from Q, K, V, the forward output O, its gradient dO and the saved log-sum-exp.
// Replaces autograd's backward of nn.MultiheadAttention's bmm-softmax-bmm (TPT/clip/model.py:185-187 under
// TPT/tpt_cls_rl.py:77) and the warp-MMA kernel of attention.cu for every tower that is ever tuned (ViT-B/32, ViT-B/16,
// the text towers, ViT-L/14 with its 257 tokens).
//
// Per unit the four [Lk x 64] fp16 tiles Q, K, V, dO are TMA-loaded once (128-byte swizzle) and every contraction runs
// on the tensor core with fp32 accumulators in TMEM.  With P = exp(S/8 - lse), D_i = sum_c dO_ic O_ic and
// dS = P o (dP - D) / 8:
//   phase A (rows = 128 queries):  S = Q K^T, dP = dO V^T (SS)  ->  threads: dS (fp16, back into TMEM)  ->  dQ = dS K (TS)
//   phase B (rows = 128 keys):     S^T = K Q^T, dP^T = V dO^T   ->  threads: P^T, dS^T                 ->  dV = P^T dO, dK = dS^T Q
// (the scores are recomputed in both orientations, as in the warp-MMA kernel: a TMEM accumulator cannot be transposed).
// One thread owns one accumulator row (lane); P / dS are packed to fp16 into TMEM columns that the row has already
// consumed, exactly like P in the forward kernel (attention_tc.cu).
//
// CTA = 6 warps: warp 0 producer (TMA of the next unit's tiles into the other shared-memory stage, and that unit's
// D / lse vectors), warp 1 TMEM allocation + MMA issuer, warps 2-5 the 128 accumulator rows (their TMEM chunk loads
// are issued one chunk ahead of the arithmetic).
// TMEM columns: [0,224) S / packed P, [224,448) dP / packed dS + dK accumulator, [448,512) dQ or dV accumulator.
// Sequences beyond 224 tokens do not fit S and dP side by side: their score columns are processed in blocks of 128
// (S in [0,128), dP in [224,352)) and the output MMAs accumulate over the blocks; everything else is unchanged.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

struct AttnBwdArgs {
  int L, Lk, heads, causal, n_units, n_rt;   // n_rt = row tiles of 128 per unit (1..3)
  int n_cb, cb_w;                            // score-column blocks per tile and their width (Lk, or 128 when Lk > 224)
  int vec_n;                                 // entries of the per-stage D / lse vectors (n_rt * 128, at least 256)
  int box_rows, n_boxes;                     // TMA boxes covering the Lk rows of a tile (a box has at most 256 rows)
  int tile_bytes;                            // smem bytes per operand tile (Lk rows; a row tile may read past it, see host)
  int n_stages;                              // shared-memory stages of {Q, K, V, dO, D, lse}: 2 when they fit, else 1
  int debug;                                 // RLCF_ATTN_BWD_DEBUG=1: CTA 0 prints a clock64 timeline of its first tiles
  const __half* out;                         // forward output O  [n_seq*L, d]
  const __half* dout;                        // dO                [n_seq*L, d]
  const float* lse;                          // [n_seq, heads, L]
  __half* dqkv;                              // [n_seq*L, 3d]
};

constexpr int kBwdTcThreads = 192;
constexpr int kColS = 0, kColDP = 224, kColOut = 448, kColDK = 352;   // TMEM column map (see header)
// barriers: FULL / FREE / DFULL exist per stage (index + stage)
enum { BB_FULL = 0, BB_FREE = 2, BB_DFULL = 4, BB_SREADY = 6, BB_PREADY = 7, BB_OREADY = 8, BB_TMEMFREE = 9, BB_COUNT = 10 };

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_fence() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// 64 fp32 accumulator columns of this thread's row -> 64 fp16 (32 packed registers) ...
__device__ __forceinline__ void pack_row64(uint32_t (&pk)[32], const uint32_t (&lo)[32], const uint32_t (&hi)[32]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    pk[j] = pack2(__uint_as_float(lo[2 * j]), __uint_as_float(lo[2 * j + 1]));
    pk[16 + j] = pack2(__uint_as_float(hi[2 * j]), __uint_as_float(hi[2 * j + 1]));
  }
}
// ... and those 128 contiguous bytes to global memory
__device__ __forceinline__ void store_row64(__half* dst, const uint32_t (&pk)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < 8; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

// kBlocks = false: the whole score row is one block (L <= 224; block loop and column offsets fold away).
template <bool kBlocks>
__global__ void __launch_bounds__(kBwdTcThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapDO, AttnBwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage s: tiles Q, K, V, dO at smem + s * 4 * tile_bytes; behind the last tile a pad of (n_rt * 128 - Lk) rows, so that a
  // row tile's 128-row A operand starting at row 128 / 256 stays inside the allocation (rows >= Lk are never stored)
  const int stage_bytes = 4 * p.tile_bytes;
  uint8_t* vec_base = smem + p.n_stages * stage_bytes + (p.n_rt * 128 - p.Lk) * 128;
  float* sLse0 = reinterpret_cast<float*>(vec_base);            // [stage][vec_n] lse * log2(e); 0 beyond L
  float* sD0 = sLse0 + 2 * p.vec_n;                              // [stage][vec_n] D_i = sum_c dO_ic O_ic; 0 beyond L
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD0 + 2 * p.vec_n);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BB_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.heads * 64;

  if (tid == 0) {
    tma_prefetch_desc(&mapQKV);
    tma_prefetch_desc(&mapDO);
    for (int i = 0; i < BB_COUNT; ++i)
      mbar_init(&bars[i], (i == BB_PREADY || i == BB_TMEMFREE) ? 4 : 1);   // 4 = one arrive per row warp
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tiles_per_unit = 2 * p.n_rt;     // phase A row tiles, then phase B row tiles
  const int n_cb = kBlocks ? p.n_cb : 1;     // score-column blocks per tile
  const int cb_w = kBlocks ? p.cb_w : p.Lk;  // and their width

  if (warp == 0) {
    // ------------------------------------------------------------ producer: tiles (TMA, lane 0) + D / lse (all lanes)
    uint32_t uc = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++uc) {
      const int h = u % p.heads, seq = u / p.heads;
      const int st = p.n_stages == 2 ? (uc & 1) : 0;
      const uint32_t sph = p.n_stages == 2 ? ((uc >> 1) & 1) : (uc & 1);   // phase of this stage's barriers
      const size_t row_base = static_cast<size_t>(seq) * p.L;
      uint8_t* sQ = smem + st * stage_bytes;
      // every MMA that read this stage has retired (the row threads' last use of its D / lse precedes those MMAs)
      mbar_wait(&bars[BB_FREE + st], sph ^ 1);
      if (lane == 0) {
        const int row = seq * p.L;
        mbar_expect_tx(&bars[BB_FULL + st], 4 * p.Lk * 128);
        for (int b = 0; b < p.n_boxes; ++b) {
          const int ro = b * p.box_rows;
          tma_load_2d(sQ + ro * 128, &mapQKV, &bars[BB_FULL + st], h * 64, row + ro);
          tma_load_2d(sQ + p.tile_bytes + ro * 128, &mapQKV, &bars[BB_FULL + st], d + h * 64, row + ro);
          tma_load_2d(sQ + 2 * p.tile_bytes + ro * 128, &mapQKV, &bars[BB_FULL + st], 2 * d + h * 64, row + ro);
          tma_load_2d(sQ + 3 * p.tile_bytes + ro * 128, &mapDO, &bars[BB_FULL + st], h * 64, row + ro);
        }
      }
      float* sLse = sLse0 + st * p.vec_n;
      float* sD = sD0 + st * p.vec_n;
      for (int rr = lane; rr < p.vec_n; rr += 32) {
        float dsum = 0.f, lv = 0.f;
        if (rr < p.L) {
          const uint4* po = reinterpret_cast<const uint4*>(p.out + (row_base + rr) * d + h * 64);
          const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (row_base + rr) * d + h * 64);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint4 a = po[c8], b = pd[c8];
            const __half2* ha = reinterpret_cast<const __half2*>(&a);
            const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fa = __half22float2(ha[e]), fb = __half22float2(hb[e]);
              dsum = fmaf(fa.x, fb.x, fmaf(fa.y, fb.y, dsum));
            }
          }
          lv = p.lse[(static_cast<size_t>(seq) * p.heads + h) * p.L + rr] * 1.4426950408889634f;
        }
        sD[rr] = dsum;
        sLse[rr] = lv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BB_DFULL + st]);    // release: the vectors above are visible to the waiters
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_o = umma_idesc_f16(128, 64) | (1u << 16);   // B operand is MN-major ([k][64] rows)
      uint32_t uc = 0, it = 0, ib = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++uc) {
        const int st = p.n_stages == 2 ? (uc & 1) : 0;
        const uint32_t sph = p.n_stages == 2 ? ((uc >> 1) & 1) : (uc & 1);
        uint8_t* sQ = smem + st * stage_bytes;
        uint8_t* sK = sQ + p.tile_bytes;
        uint8_t* sV = sK + p.tile_bytes;
        uint8_t* sdO = sV + p.tile_bytes;
        mbar_wait(&bars[BB_FULL + st], sph);
        tc_fence_after();
        for (int t = 0; t < tiles_per_unit; ++t, ++it) {
          const bool phase_b = t >= p.n_rt;
          const int r0 = (phase_b ? t - p.n_rt : t) * 128;
          // scores: A = this tile's 128 rows of (Q | K), B = a block of rows of (K | Q); dP alike with (dO | V) x (V | dO)
          const uint64_t a_s = umma_desc_k_sw128(smem_u32((phase_b ? sK : sQ) + r0 * 128));
          const uint64_t a_p = umma_desc_k_sw128(smem_u32((phase_b ? sV : sdO) + r0 * 128));
          mbar_wait(&bars[BB_TMEMFREE], (it & 1) ^ 1);    // the previous tile's accumulators have been read out
          tc_fence_after();
          for (int cb = 0; cb < n_cb; ++cb, ++ib) {
            const int c0 = kBlocks ? cb * cb_w : 0;                   // first score column (key in phase A, query in phase B)
            const int w = kBlocks ? min(cb_w, p.Lk - c0) : p.Lk;   // multiple of 16
            const uint32_t idesc_s = umma_idesc_f16(128, w);
            const uint64_t b_s = umma_desc_k_sw128(smem_u32((phase_b ? sQ : sK) + c0 * 128));
            const uint64_t b_p = umma_desc_k_sw128(smem_u32((phase_b ? sdO : sV) + c0 * 128));
            // (cb > 0: these overwrite the packed operands of the previous block's output MMAs, which the tensor core
            // executes first -- MMAs of one thread run in issue order -- and whose inputs the row threads have finished)
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss<1>(tmem + kColS, a_s + 2 * k, b_s + 2 * k, idesc_s, k != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss<1>(tmem + kColDP, a_p + 2 * k, b_p + 2 * k, idesc_s, k != 0);
            umma_commit(&bars[BB_SREADY]);
            mbar_wait(&bars[BB_PREADY], ib & 1);          // packed dS (and P^T) of this block are in TMEM
            tc_fence_after();
            const int ksteps = w >> 4, k0 = c0 >> 4;      // output MMAs accumulate over the blocks
            if (!phase_b) {
              const uint64_t bk = umma_desc_k_sw128(smem_u32(sK));
              for (int j = 0; j < ksteps; ++j)
                umma_ts(tmem + kColOut, tmem + kColS + 8 * j, bk + 128 * (k0 + j), idesc_o, (cb | j) != 0);
            } else {
              const uint64_t bdo = umma_desc_k_sw128(smem_u32(sdO));
              const uint64_t bq = umma_desc_k_sw128(smem_u32(sQ));
              for (int j = 0; j < ksteps; ++j)
                umma_ts(tmem + kColOut, tmem + kColS + 8 * j, bdo + 128 * (k0 + j), idesc_o, (cb | j) != 0);
              for (int j = 0; j < ksteps; ++j)
                umma_ts(tmem + kColDK, tmem + kColDP + 8 * j, bq + 128 * (k0 + j), idesc_o, (cb | j) != 0);
            }
          }
          umma_commit(&bars[BB_OREADY]);
          if (t == tiles_per_unit - 1) umma_commit(&bars[BB_FREE + st]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ accumulator rows: thread = one TMEM lane
    const int q4 = warp & 3;                              // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    const float scale = 0.125f;
    const float c = scale * 1.4426950408889634f;
    uint32_t it = 0, uc = 0, ib = 0;
    const bool probe = p.debug && blockIdx.x == 0 && warp == 2 && lane == 0;
    long long stamp[12][6];
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++uc) {
      const int h = u % p.heads, seq = u / p.heads;
      const size_t row_base = static_cast<size_t>(seq) * p.L;
      const int st = p.n_stages == 2 ? (uc & 1) : 0;
      const uint32_t sph = p.n_stages == 2 ? ((uc >> 1) & 1) : (uc & 1);
      const float* sLse = sLse0 + st * p.vec_n;
      const float* sD = sD0 + st * p.vec_n;
      mbar_wait(&bars[BB_DFULL + st], sph);               // this unit's D / lse vectors (written by the producer warp)
      for (int t = 0; t < tiles_per_unit; ++t, ++it) {
        const bool phase_b = t >= p.n_rt;
        const int r0 = (phase_b ? t - p.n_rt : t) * 128;
        const int grow = r0 + r;                          // query (phase A) or key (phase B) of this thread
        const uint32_t ph = it & 1;
        const float lse_r = sLse[grow], d_r = sD[grow];
        // d_rs = D * scale is folded into one FFMA: dS = p * (dP * scale - D * scale)
        const float d_rs = d_r * scale;
        const int wrow0 = r0 + q4 * 32;                   // first row (query in phase A, key in phase B) of this warp
        // rows >= L only produce accumulator rows that are never stored (every MMA here is row-independent in A): a warp
        // without a live row skips the arithmetic (the last row tile of ViT-L/14's 257 tokens holds ONE live row)
        const bool warp_live = wrow0 < p.L;
        if (probe && it < 12) stamp[it][0] = clock64();
        for (int cb = 0; cb < n_cb; ++cb, ++ib) {
          const int c0 = kBlocks ? cb * cb_w : 0;                     // first score column of this block
          const int n_chunks = ((kBlocks ? min(cb_w, p.Lk - c0) : p.Lk) + 31) >> 5;
          mbar_wait(&bars[BB_SREADY], ib & 1);
          tc_fence_after();
          if (probe && it < 12 && cb == 0) stamp[it][1] = clock64();
          if (warp_live) {
            // two register sets: the loads of chunk ch + 1 are in flight while chunk ch is processed
            uint32_t s0[32], dp0[32], s1[32], dp1[32];
            tmem_ld_32x32(trow + kColS, s0);
            tmem_ld_32x32(trow + kColDP, dp0);
            // Chunks whose 32 columns are all visible to all 32 rows of this warp take a straight-line path without any
            // predicate arithmetic (4.5 instructions per element instead of ~19; a lone warp per SM sub-partition issues
            // at ~0.3 IPC, so the instruction count is what the chunk loop costs): warp-uniform `full` below.
            auto process = [&](auto full_tag, const uint32_t (&s)[32], const uint32_t (&dp)[32], int ch) {
              constexpr bool kFull = decltype(full_tag)::value;
              uint32_t pk_ds[16], pk_p[16];
              const int col0 = c0 + ch * 32;              // global index of the chunk's first column
              if (!phase_b) {
                // columns = keys; this row's query is `grow`
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float v2[2];
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const float pr = ex2_approx(fmaf(__uint_as_float(s[2 * j + e]), c, -lse_r));
                    const float v = pr * fmaf(__uint_as_float(dp[2 * j + e]), scale, -d_rs);
                    if constexpr (kFull) {
                      v2[e] = v;
                    } else {
                      const int key = col0 + 2 * j + e;
                      v2[e] = (key < p.L && (!p.causal || key <= grow)) ? v : 0.f;
                    }
                  }
                  pk_ds[j] = pack2(v2[0], v2[1]);
                }
                tmem_st16(trow + kColS + ch * 16, pk_ds);
              } else {
                // columns = queries; this row's key is `grow`; lse and D of the 32 queries come from shared memory
                const float4* l4 = reinterpret_cast<const float4*>(sLse + col0);
                const float4* d4 = reinterpret_cast<const float4*>(sD + col0);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 lq = l4[j4], dq = d4[j4];
                  const float lqa[4] = {lq.x, lq.y, lq.z, lq.w}, dqa[4] = {dq.x, dq.y, dq.z, dq.w};
                  float pv[4], dv[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float pr = ex2_approx(fmaf(__uint_as_float(s[4 * j4 + e]), c, -lqa[e]));
                    const float v = pr * fmaf(__uint_as_float(dp[4 * j4 + e]), scale, -dqa[e] * scale);
                    if constexpr (kFull) {
                      pv[e] = pr;
                      dv[e] = v;
                    } else {
                      const int q = col0 + 4 * j4 + e;
                      const bool ok = q < p.L && (!p.causal || grow <= q);
                      pv[e] = ok ? pr : 0.f;
                      dv[e] = ok ? v : 0.f;
                    }
                  }
                  pk_p[2 * j4] = pack2(pv[0], pv[1]);
                  pk_p[2 * j4 + 1] = pack2(pv[2], pv[3]);
                  pk_ds[2 * j4] = pack2(dv[0], dv[1]);
                  pk_ds[2 * j4 + 1] = pack2(dv[2], dv[3]);
                }
                tmem_st16(trow + kColS + ch * 16, pk_p);
                tmem_st16(trow + kColDP + ch * 16, pk_ds);
              }
            };
            auto chunk = [&](const uint32_t (&s)[32], const uint32_t (&dp)[32], int ch) {
              // every column of the chunk is a real token, and (causal) visible to every row of this warp
              const int col0 = c0 + ch * 32;
              bool full = col0 + 32 <= p.L;
              if (p.causal) full = full && (phase_b ? (wrow0 + 31 <= col0) : (col0 + 31 <= wrow0));
              if (full) process(std::true_type{}, s, dp, ch); else process(std::false_type{}, s, dp, ch);
            };
            for (int ch = 0; ch < n_chunks; ch += 2) {
              tmem_ld_wait();                                // chunk ch is in s0 / dp0
              if (ch + 1 < n_chunks) {
                tmem_ld_32x32(trow + kColS + (ch + 1) * 32, s1);
                tmem_ld_32x32(trow + kColDP + (ch + 1) * 32, dp1);
              }
              chunk(s0, dp0, ch);
              if (ch + 1 < n_chunks) {
                tmem_ld_wait();                              // chunk ch + 1 is in s1 / dp1
                if (ch + 2 < n_chunks) {
                  tmem_ld_32x32(trow + kColS + (ch + 2) * 32, s0);
                  tmem_ld_32x32(trow + kColDP + (ch + 2) * 32, dp0);
                }
                chunk(s1, dp1, ch + 1);
              }
            }
            tmem_st_fence();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BB_PREADY]);
        }
        if (probe && it < 12) stamp[it][2] = stamp[it][3] = clock64();
        // ---- read the output accumulators of this tile
        mbar_wait(&bars[BB_OREADY], ph);
        tc_fence_after();
        if (probe && it < 12) stamp[it][4] = clock64();
        {
          // All accumulators of the tile go to registers first and TMEM is handed back at once: the 8 (16 in phase B)
          // scattered 16-byte stores per thread then drain while the tensor core already works on the next tile
          // (issued before the hand-back they sat on the critical path: 2500-3700 cycles per phase-B tile).
          uint32_t pa[32], pb[32];
          {
            uint32_t lo[32], hi[32];
            tmem_ld_32x32(trow + kColOut, lo);
            tmem_ld_32x32(trow + kColOut + 32, hi);
            tmem_ld_wait();
            pack_row64(pa, lo, hi);
            if (phase_b) {
              tmem_ld_32x32(trow + kColDK, lo);
              tmem_ld_32x32(trow + kColDK + 32, hi);
              tmem_ld_wait();
              pack_row64(pb, lo, hi);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BB_TMEMFREE]);
          __half* dst = p.dqkv + (row_base + grow) * (3 * static_cast<size_t>(d)) + h * 64;
          if (grow < p.L) {
            store_row64(dst + (phase_b ? 2 * d : 0), pa);         // dV (phase B) or dQ (phase A)
            if (phase_b) store_row64(dst + d, pb);                // dK
          }
        }
        if (probe && it < 12) stamp[it][5] = clock64();
      }
    }
    if (probe)
      for (uint32_t i = 0; i < (it < 12 ? it : 12); ++i)
        printf("attn_bwd_tc tile %u: wait S %lld | chunks %lld | st fence %lld | wait out %lld | read-out+store %lld | total %lld\n",
               i, stamp[i][1] - stamp[i][0], stamp[i][2] - stamp[i][1], stamp[i][3] - stamp[i][2],
               stamp[i][4] - stamp[i][3], stamp[i][5] - stamp[i][4], i ? stamp[i][5] - stamp[i - 1][5] : 0ll);
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

static int make_tmap_rows64_bwd(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows) {
  static PFN_encodeTiled encode = get_encode_tiled();
  if (encode == nullptr) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled(attention bwd) failed (%d)", static_cast<int>(r));
  return 0;
}

// Returns -1 when the shape is outside what this kernel covers (the caller falls back to the warp-MMA kernel).
int attention_bwd_tc(const __half* qkv, const __half* out, const __half* dout, const float* lse, int n_seq, int L,
                     int heads, int causal, __half* dqkv, cudaStream_t stream) {
  const int Lk = (L + 15) / 16 * 16;
  if (Lk > 384 || Lk < 16) return -1;
  if (((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) |
        reinterpret_cast<uintptr_t>(dqkv)) & 15) != 0)
    return -1;
  AttnBwdArgs a{};
  a.L = L; a.Lk = Lk; a.heads = heads; a.causal = causal; a.n_units = heads * n_seq;
  a.n_rt = (L + 127) / 128;
  // S and dP (fp32, Lk columns each) sit side by side in TMEM up to Lk = 224; longer rows go in blocks of 128 columns
  a.cb_w = Lk <= 224 ? Lk : 128;
  a.n_cb = (Lk + a.cb_w - 1) / a.cb_w;
  a.vec_n = a.n_rt * 128 < 256 ? 256 : a.n_rt * 128;
  a.n_boxes = Lk <= 256 ? 1 : 2;
  a.box_rows = Lk / a.n_boxes;           // Lk % 16 == 0: the halves of a 272-row tile are 136 rows (a multiple of 8)
  if (a.box_rows * a.n_boxes != Lk || a.box_rows % 8 != 0) return -1;
  a.tile_bytes = Lk * 128;               // multiple of 1024 (Lk % 16 == 0): every tile start keeps the swizzle alignment
  a.out = out; a.dout = dout; a.lse = lse; a.dqkv = dqkv;
  static const int debug = getenv("RLCF_ATTN_BWD_DEBUG") != nullptr ? atoi(getenv("RLCF_ATTN_BWD_DEBUG")) : 0;
  a.debug = debug;
  // A row tile's A operand spans 128 rows from row 0 / 128 / 256 of its tile, i.e. up to (n_rt * 128 - Lk) rows past the
  // tile's end: into the next tile, or -- for the last tile -- into a pad of that size.  Rows >= L only produce rows that
  // are never stored.
  auto smem_for = [&](int stages) {
    return 1024 + static_cast<size_t>(stages) * 4 * a.tile_bytes + static_cast<size_t>(a.n_rt * 128 - Lk) * 128 +
           4 * static_cast<size_t>(a.vec_n) * sizeof(float) + BB_COUNT * 8 + 16;
  };
  a.n_stages = smem_for(2) <= 227 * 1024 ? 2 : 1;
  const size_t smem = smem_for(a.n_stages);
  if (smem > 227 * 1024) return -1;
  const bool blocks = a.n_cb > 1;
  auto kernel = blocks ? attn_bwd_tc_kernel<true> : attn_bwd_tc_kernel<false>;
  static DynSmemState st[2];
  if (cudaError_t e = ensure_dyn_smem(kernel, smem, st[blocks]))
    return set_error(RLCF_ERR_CUDA, "attention_bwd_tc attr: %s", cudaGetErrorString(e));
  CUtensorMap mqkv, mdo;
  const long long rows = static_cast<long long>(n_seq) * L;
  if (int rc = make_tmap_rows64_bwd(&mqkv, qkv, rows, 3 * heads * 64, a.box_rows)) return rc;
  if (int rc = make_tmap_rows64_bwd(&mdo, dout, rows, heads * 64, a.box_rows)) return rc;
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  kernel<<<grid, kBwdTcThreads, smem, stream>>>(mqkv, mdo, a);
  RLCF_CHECK_LAUNCH("attention_bwd_tc");
  return 0;
}

}  // namespace rlcf
