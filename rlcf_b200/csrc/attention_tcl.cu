// tcgen05 fused attention forward for LONG sequences (272 < L <= 640 tokens, head_dim 64): ViT-L/14@336px, the largest
// reward model the reference lists (TPT/clip_reward.py:22-27), has 577 tokens -- more than attention_tc.cu's "all
// scores of a query tile in 256 TMEM columns" allows.  This kernel walks the keys in blocks.
//
// Work unit = one (sequence, head): its K and V (nb blocks of KB keys, nb * KB >= L) are TMA-loaded once and stay in
// shared memory while the CTA walks the unit's 128-row query tiles.  Per tile, two passes over the key blocks:
//   pass A : S_j = Q K_j^T (SS MMA into TMEM), the softmax threads only take the row maximum of each block;
//   pass B : S_j again (the tensor pipe is 85 % idle in this kernel, recomputing is cheaper than keeping 640 columns),
//            p = 2^((s - m) c) with the FINAL row maximum, written back as fp16 into the consumed columns, and
//            O += P_j [V_j | 1] (TS MMA; the block of ones makes O column 64 the row sum).
// Because pass B already knows the row maximum there is no online-softmax rescaling of O and the result does not depend
// on the block order.  Two softmax groups of four warps (two warps per SM sub-partition) each own one S buffer
// (TMEM columns [0, 208) and [208, 416); O at [416, 496)) and take every second key block, so the MMA of one block
// overlaps the softmax of the other and the row loops -- dependent-issue-latency bound, profiles/r2_attention_probes.txt
// -- have twice the warps to hide behind.  The two partial row maxima of a row meet in shared memory once per tile.
// Replaces the bmm-softmax-bmm of nn.MultiheadAttention (TPT/clip/model.py:185-187) for those sequence lengths; before
// this kernel they ran on the mma.sync kernel of attention.cu.
#include <cstdlib>

#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

struct AttnTclArgs {
  int L, heads;
  int KB, nb;            // keys per block (multiple of 16, <= 208), key blocks
  int n_qt, n_units;     // query tiles per unit; (sequence, head) units
  __half* out;
  float* lse;
};

constexpr int kTclThreads = 320;   // warp 0: TMA, warp 1: MMA, warps 2-5: softmax group 0, warps 6-9: group 1
constexpr int kTclColS1 = 208;     // second S buffer
constexpr int kTclColO = 416;      // O accumulator: 64 + 16 columns
enum { L_KVFULL = 0, L_KVFREE, L_QFULL /*+buf*/, L_QFREE = L_QFULL + 2 /*+buf*/, L_SREADY = L_QFREE + 2 /*+g*/,
       L_SDONE = L_SREADY + 2 /*+g*/, L_PREADY = L_SDONE + 2 /*+g*/, L_PVDONE = L_PREADY + 2 /*+g*/,
       L_OREADY = L_PVDONE + 2, L_OFREE, L_NBARS };

__device__ __forceinline__ void tcl_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ void tcl_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tcl_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tcl_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tcl_pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
// MN-major, 128-byte-swizzled B operand of two 64-column blocks: the second block's atoms start lbo_bytes after the first's
__device__ __forceinline__ uint64_t tcl_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// max over the first `valid` keys of a block's 32-column chunk `ch`
__device__ __forceinline__ float tcl_chunk_max(const uint32_t (&v)[32], float m, int ch, int valid) {
  float m2 = -INFINITY;
  if ((ch + 1) * 32 <= valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      m = fmaxf(m, __uint_as_float(v[j]));
      m2 = fmaxf(m2, __uint_as_float(v[16 + j]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      m = fmaxf(m, (ch * 32 + j < valid) ? __uint_as_float(v[j]) : -INFINITY);
      m2 = fmaxf(m2, (ch * 32 + 16 + j < valid) ? __uint_as_float(v[16 + j]) : -INFINITY);
    }
  }
  return fmaxf(m, m2);
}

// p = 2^(s c - m c) of one chunk, packed to fp16 into 16 TMEM columns at `dst`; keys >= valid get exactly 0
__device__ __forceinline__ void tcl_chunk_exp(const uint32_t (&v)[32], uint32_t dst, int ch, int valid, float c, float mc) {
  uint32_t pk[16];
  if ((ch + 1) * 32 <= valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = ex2_approx(fmaf(__uint_as_float(v[2 * j]), c, -mc));
      const float b = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), c, -mc));
      const __half2 hp = __floats2half2_rn(a, b);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hp);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = (ch * 32 + 2 * j < valid) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), c, -mc)) : 0.f;
      const float b = (ch * 32 + 2 * j + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), c, -mc)) : 0.f;
      const __half2 hp = __floats2half2_rn(a, b);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hp);
    }
  }
  tcl_st16(dst, pk);
}

__global__ void __launch_bounds__(kTclThreads, 1)
attn_fwd_tcl_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV,
                    const __grid_constant__ CUtensorMap mapO, AttnTclArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int blk_bytes = p.KB * 128;
  uint8_t* sQ = smem;                                   // two Q tiles of 16 KB (the O tile is staged in the same buffer)
  uint8_t* sK = sQ + 2 * 128 * 128;                     // [nb][KB][64] halves, 128-byte swizzled rows
  uint8_t* sV = sK + p.nb * blk_bytes;
  uint8_t* sOnes = sV + p.nb * blk_bytes;               // [KB][64] halves: column 0 = 1
  float* pmax = reinterpret_cast<float*>(sOnes + blk_bytes);              // [tile parity][group][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(pmax + 2 * 2 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L_NBARS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.heads * 64;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapKV);
    tma_prefetch_desc(&mapO);
    for (int i = 0; i < L_NBARS; ++i) mbar_init(&bars[i], 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars[L_SDONE + g], 4);      // one arrive per warp of the group
      mbar_init(&bars[L_PREADY + g], 4);
      mbar_init(&bars[L_QFREE + g], 4);      // the four warps that TMA-store the O tile out of the Q buffer
    }
    mbar_init(&bars[L_OFREE], 8);
    fence_barrier_init();
  }
  for (int i = tid; i < p.KB * 8; i += kTclThreads) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // element (row r, column 0) lives in the row's logical 16-byte chunk 0 = physical chunk (0 ^ (r & 7))
  for (int r = tid; r < p.KB; r += kTclThreads)
    *reinterpret_cast<__half*>(sOnes + r * 128 + ((r & 7) << 4)) = __float2half(1.f);
  fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int first = blockIdx.x, stride = gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t uc = 0, tc = 0;
      for (int u = first; u < p.n_units; u += stride, ++uc) {
        const int h = u % p.heads, seq = u / p.heads;
        const int row_base = seq * p.L;
        mbar_wait(&bars[L_KVFREE], (uc & 1) ^ 1);
        mbar_expect_tx(&bars[L_KVFULL], 2 * p.nb * blk_bytes);
        for (int b = 0; b < p.nb; ++b) {   // rows past the end of the tensor are zero-filled, rows of the next sequence masked
          tma_load_2d(sK + b * blk_bytes, &mapKV, &bars[L_KVFULL], d + h * 64, row_base + b * p.KB);
          tma_load_2d(sV + b * blk_bytes, &mapKV, &bars[L_KVFULL], 2 * d + h * 64, row_base + b * p.KB);
        }
        for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
          const int buf = tc & 1;
          mbar_wait(&bars[L_QFREE + buf], ((tc >> 1) & 1) ^ 1);
          mbar_expect_tx(&bars[L_QFULL + buf], 128 * 128);
          tma_load_2d(sQ + buf * 128 * 128, &mapQ, &bars[L_QFULL + buf], h * 64, row_base + qt * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(128, p.KB);
      const uint32_t idesc_o = umma_idesc_f16(128, 80) | (1u << 16);   // B (= V | ones) is MN-major
      // per S buffer: completed uses of each kind, and what the last use was (0 none, 1 pass A, 2 pass B)
      uint32_t nA[2] = {0, 0}, nB[2] = {0, 0}, nP[2] = {0, 0};
      int last[2] = {0, 0};
      uint32_t uc = 0, tc = 0;
      auto wait_free = [&](int g) {          // the previous contents of S buffer g have been consumed
        if (last[g] == 1) mbar_wait(&bars[L_SDONE + g], (nA[g] - 1) & 1);
        else if (last[g] == 2) mbar_wait(&bars[L_PVDONE + g], (nB[g] - 1) & 1);
        tc_fence_after();
      };
      for (int u = first; u < p.n_units; u += stride, ++uc) {
        mbar_wait(&bars[L_KVFULL], uc & 1);
        for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
          const int buf = tc & 1;
          const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ + buf * 128 * 128));
          mbar_wait(&bars[L_QFULL + buf], (tc >> 1) & 1);
          tc_fence_after();
          auto issue_s = [&](int j) {
            const int g = j & 1;
            const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + j * blk_bytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss<1>(tmem + g * kTclColS1, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            umma_commit(&bars[L_SREADY + g]);
          };
          auto issue_pv = [&](int j) {
            const int g = j & 1;
            mbar_wait(&bars[L_PREADY + g], nP[g] & 1);
            ++nP[g];
            if (j == 0) mbar_wait(&bars[L_OFREE], (tc & 1) ^ 1);   // the previous tile's O has been read out of TMEM
            tc_fence_after();
            const uint32_t sv = smem_u32(sV + j * blk_bytes);
            const uint64_t dv = tcl_desc_mn(sv, smem_u32(sOnes) - sv);
            const int keys = min(p.KB, p.L - j * p.KB);
            const int ksteps = (keys + 15) >> 4;
            for (int k = 0; k < ksteps; ++k)
              tcl_umma_ts(tmem + kTclColO, tmem + g * kTclColS1 + 8 * k, dv + 128 * k, idesc_o, (j | k) != 0);
            umma_commit(&bars[L_PVDONE + g]);
          };
          for (int j = 0; j < p.nb; ++j) {         // pass A: scores for the row maxima
            const int g = j & 1;
            wait_free(g);
            issue_s(j);
            ++nA[g];
            last[g] = 1;
          }
          for (int j = 0; j < p.nb; ++j) {         // pass B: scores again, then P V one block behind
            const int g = j & 1;
            wait_free(g);
            issue_s(j);
            ++nB[g];
            last[g] = 2;
            if (j >= 1) issue_pv(j - 1);
          }
          issue_pv(p.nb - 1);
          umma_commit(&bars[L_OREADY]);
          if (qt == p.n_qt - 1) umma_commit(&bars[L_KVFREE]);   // every reader of this unit's K / V has retired
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue: thread = (query row, key-block parity)
    const int g = (warp - 2) >> 2;                 // group = S buffer
    const int q4 = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    const uint32_t tS = trow + g * kTclColS1;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    uint32_t use = 0, tc = 0;                      // uses of this group's S buffer so far; tiles so far
    for (int u = first; u < p.n_units; u += stride) {
      const int h = u % p.heads, seq = u / p.heads;
      for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
        uint8_t* sQb = sQ + (tc & 1) * 128 * 128;
        const int q0 = qt * 128, qrow = q0 + r;
        const bool warp_live = q0 + q4 * 32 < p.L;   // rows of a dead warp are never stored
        // ---- pass A: row maximum over this group's key blocks
        float m = -INFINITY;
        for (int j = g; j < p.nb; j += 2, ++use) {
          mbar_wait(&bars[L_SREADY + g], use & 1);
          tc_fence_after();
          if (warp_live) {
            const int valid = min(p.KB, p.L - j * p.KB);
            const int n_ch = (valid + 31) >> 5;
            uint32_t va[32], vb[32];
            tmem_ld_32x32(tS, va);
            for (int ch = 0; ch < n_ch; ch += 2) {
              tmem_ld_wait();
              if (ch + 1 < n_ch) tmem_ld_32x32(tS + (ch + 1) * 32, vb);
              m = tcl_chunk_max(va, m, ch, valid);
              if (ch + 1 < n_ch) {
                tmem_ld_wait();
                if (ch + 2 < n_ch) tmem_ld_32x32(tS + (ch + 2) * 32, va);
                m = tcl_chunk_max(vb, m, ch + 1, valid);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[L_SDONE + g]);
        }
        // ---- the two partial maxima of a row meet (a group without blocks contributes -inf)
        float* pm = pmax + (tc & 1) * 256;
        pm[g * 128 + r] = m;
        tcl_pair_sync(1 + q4);
        m = fmaxf(m, pm[(g ^ 1) * 128 + r]);
        const float mc = m * c;
        // ---- pass B: P = 2^((s - m) c) as fp16 into the consumed S columns
        for (int j = g; j < p.nb; j += 2, ++use) {
          mbar_wait(&bars[L_SREADY + g], use & 1);
          tc_fence_after();
          if (warp_live) {
            const int valid = min(p.KB, p.L - j * p.KB);
            const int n_ch = (valid + 31) >> 5;
            uint32_t va[32], vb[32];
            tmem_ld_32x32(tS, va);
            for (int ch = 0; ch < n_ch; ch += 2) {
              tmem_ld_wait();
              if (ch + 1 < n_ch) tmem_ld_32x32(tS + (ch + 1) * 32, vb);
              tcl_chunk_exp(va, tS + 16 * ch, ch, valid, c, mc);
              if (ch + 1 < n_ch) {
                tmem_ld_wait();
                if (ch + 2 < n_ch) tmem_ld_32x32(tS + (ch + 2) * 32, va);
                tcl_chunk_exp(vb, tS + 16 * (ch + 1), ch + 1, valid, c, mc);
              }
            }
            tcl_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[L_PREADY + g]);
        }
        // ---- epilogue: group g normalises and stages O columns [32 g, 32 g + 32)
        mbar_wait(&bars[L_OREADY], tc & 1);
        tc_fence_after();
        if (warp_live) {
          uint32_t o[32], lsum[16];
          tmem_ld_32x32(trow + kTclColO + 32 * g, o);
          tcl_ld16(trow + kTclColO + 64, lsum);      // O column 64 = sum_j P_j * 1
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[L_OFREE]);
          const float l = __uint_as_float(lsum[0]);
          const float inv = 1.f / l;
          uint4* srow = reinterpret_cast<uint4*>(sQb + r * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __half2 h0 = __floats2half2_rn(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
            __half2 h1 = __floats2half2_rn(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
            __half2 h2 = __floats2half2_rn(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
            __half2 h3 = __floats2half2_rn(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
            srow[(4 * g + j) ^ (r & 7)] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                     *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          }
          if (g == 1 && qrow < p.L && p.lse != nullptr)
            p.lse[(static_cast<size_t>(seq) * p.heads + h) * p.L + qrow] = m * 0.125f + logf(l);
          fence_proxy_async();
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[L_OFREE]);
        }
        tcl_pair_sync(1 + q4);                      // both halves of the 32 staged rows are in shared memory
        if (g == 0) {
          if (lane == 0) {
            if (warp_live) {                        // rows >= L are clipped by the (column, token, sequence) map
              tma_store_3d(&mapO, sQb + q4 * 32 * 128, h * 64, q0 + q4 * 32, seq);
              bulk_commit_group();
              bulk_wait_read_all();                 // the store has read the buffer: the next Q tile may land in it
            }
            mbar_arrive(&bars[L_QFREE + (tc & 1)]);
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

static int tcl_tmap_rows64(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows) {
  static PFN_encodeTiled encode = get_encode_tiled();
  if (encode == nullptr) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled(attention tcl) failed (%d)", static_cast<int>(r));
  return 0;
}

// Returns -1 when the shape is outside what this kernel covers (the caller falls back to the mma.sync kernel).
int attention_fwd_tcl(const __half* qkv, int n_seq, int L, int heads, int causal, __half* out, float* lse,
                      cudaStream_t stream) {
  if (causal || L < 32) return -1;     // long sequences are the vision towers; the causal text towers have 77 tokens
  if (((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return -1;
  AttnTclArgs a{};
  a.L = L; a.heads = heads; a.out = out; a.lse = lse;
  a.nb = 2 * ((L + 2 * 208 - 1) / (2 * 208));                 // an even number of blocks of at most 208 keys
  a.KB = ((L + a.nb - 1) / a.nb + 15) / 16 * 16;
  if (a.KB > 208) return -1;
  if (static_cast<long long>(a.nb - 1) * a.KB >= L) return -1;   // every block holds at least one key
  a.n_qt = (L + 127) / 128;
  a.n_units = heads * n_seq;
  const size_t smem = 1024 + 2 * 128 * 128 + static_cast<size_t>(2 * a.nb + 1) * a.KB * 128 + 2 * 2 * 128 * sizeof(float) +
                      L_NBARS * 8 + 16;
  if (smem > 227 * 1024) return -1;
  static DynSmemState st;
  if (cudaError_t e = ensure_dyn_smem(attn_fwd_tcl_kernel, smem, st))
    return set_error(RLCF_ERR_CUDA, "attention_fwd_tcl attr: %s", cudaGetErrorString(e));
  CUtensorMap mq, mkv, mo;
  const long long rows = static_cast<long long>(n_seq) * L;
  if (int rc = tcl_tmap_rows64(&mq, qkv, rows, 3 * heads * 64, 128)) return rc;
  if (int rc = tcl_tmap_rows64(&mkv, qkv, rows, 3 * heads * 64, a.KB)) return rc;
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
    if (r != CUDA_SUCCESS) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled(attention tcl out) failed (%d)", static_cast<int>(r));
  }
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  attn_fwd_tcl_kernel<<<grid, kTclThreads, smem, stream>>>(mq, mkv, mo, a);
  RLCF_CHECK_LAUNCH("attention_fwd_tcl");
  return 0;
}

}  // namespace rlcf
