// Variant of the tcgen05 fused attention forward (attention_tc.cu) with EIGHT softmax warps per team: two warps per SM
// sub-partition and team, each owning half of the 32-key chunks of the same 32 query rows.
//
// Why: profiles/r2_attention_probes.txt -- the softmax loops of attention_tc.cu are bound by dependent-issue latency
// (~0.2 IPC with one warp per team and sub-partition), not by the TMEM port, the MUFU lanes or the instruction count.
// Twice the warps, each with half the chunks, halves the length of every softmax pass of a tile.  Everything else is the
// scheme of attention_tc.cu: two teams per CTA, TMA loads of Q / K / V, S = Q K^T (SS) into TMEM, P written back as fp16
// into TMEM, O = P [V | 1] (TS; the block of ones makes O column 64 the row sum), TMA store of O.
//
// What the column split needs:
//   * the row maximum is exchanged between the two warps of a row through shared memory (one named barrier per pair);
//   * the second half's P chunks cannot go to "columns the row has already consumed" -- those belong to the first half's
//     S chunks, which its exponential pass may still be reading -- so they are written to the 48 TMEM columns behind S
//     ([208, 256) of the team's 256); the P V MMA takes its A operand from the two places.  Hence Lk <= 208 (7 chunks:
//     4 + 3): ViT-B/32, ViT-B/16 and the text towers.  ViT-L/14 (257 tokens) stays on attention_tc.cu.
// Replaces the bmm-softmax-bmm of nn.MultiheadAttention (TPT/clip/model.py:185-187).
#include <cstdlib>

#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

struct AttnTc8Args {
  int L, Lk, heads, causal;
  int n_chunks, n_first;        // 32-key chunks of S; chunks [0, n_first) belong to the first warp of a row pair
  int n_qt, n_units;            // query tiles per (sequence, head); number of (sequence, head) units
  int stage_bytes;
  int use_tok;                  // exponential-pass token between the two teams (as in attention_tc.cu)
  __half* out;
  float* lse;
};

constexpr int kTc8Threads = 640;   // warps 0/3: TMA (team 0/1), warps 1/2: MMA (team 0/1), warps 4-11 / 12-19: softmax
constexpr int kTc8ColO = 128;      // O accumulator (64 + 16 columns) inside the team's 256 TMEM columns
constexpr int kTc8ColP2 = 208;     // packed P of the second chunk half
enum { T8_KVFULL = 0, T8_KVFREE = 1, T8_QFULL = 2 /*+buf*/, T8_QFREE = 4 /*+buf*/, T8_SREADY = 6, T8_PREADY = 7,
       T8_OREADY = 8, T8_TMEMFREE = 9, T8_PER_TEAM = 10 };

__device__ __forceinline__ void tc8_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ void tc8_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc8_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc8_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc8_pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
// MN-major, 128-byte-swizzled B operand of two 64-column blocks: the second block's atoms start lbo_bytes after the first's
__device__ __forceinline__ uint64_t tc8_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float tc8_chunk_max(const uint32_t (&v)[32], float m, int ch, bool full, int key_end) {
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

// p = 2^(s c - m c) of one chunk, packed to fp16 and written to 16 TMEM columns at `dst` (the row sum comes from the MMA)
__device__ __forceinline__ void tc8_chunk_exp(const uint32_t (&v)[32], uint32_t dst, int ch, bool full, int key_end,
                                              float c, float mc) {
  uint32_t pk[16];
  if (full) {
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
      const float a = (ch * 32 + 2 * j < key_end) ? ex2_approx(fmaf(__uint_as_float(v[2 * j]), c, -mc)) : 0.f;
      const float b = (ch * 32 + 2 * j + 1 < key_end) ? ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), c, -mc)) : 0.f;
      const __half2 hp = __floats2half2_rn(a, b);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hp);
    }
  }
  tc8_st16(dst, pk);
}

__global__ void __launch_bounds__(kTc8Threads, 1)
attn_fwd_tc8_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV,
                    const __grid_constant__ CUtensorMap mapO, AttnTc8Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOnes = smem + 2 * p.stage_bytes;                                // [Lk][64] halves: column 0 = 1
  float* pmax = reinterpret_cast<float*>(sOnes + p.Lk * 128);               // [tile parity][team][half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(pmax + 2 * 2 * 2 * 128);     // [2 teams][T8_PER_TEAM]
  uint64_t* tok = bars + 2 * T8_PER_TEAM;                                   // [4 lane quarters][2 teams]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.heads * 64;

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&tok[i], 2);     // both warps of a row pair pass the token on
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapKV);
    tma_prefetch_desc(&mapO);
    for (int t = 0; t < 2; ++t) {
      uint64_t* b = bars + t * T8_PER_TEAM;
      for (int i = 0; i < T8_PER_TEAM; ++i) mbar_init(&b[i], 1);
      mbar_init(&b[T8_PREADY], 8);     // one arrive per softmax warp
      mbar_init(&b[T8_TMEMFREE], 8);
      mbar_init(&b[T8_QFREE], 5);      // the MMA commit + the four warps whose TMA stores read the staged O tile
      mbar_init(&b[T8_QFREE + 1], 5);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < p.Lk * 8; i += kTc8Threads) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int r = tid; r < p.Lk; r += kTc8Threads)      // (row r, column 0) sits in physical 16-byte chunk (0 ^ (r & 7))
    *reinterpret_cast<__half*>(sOnes + r * 128 + ((r & 7) << 4)) = __float2half(1.f);
  fence_proxy_async();
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int team = warp >= 4 ? (warp - 4) >> 3 : (warp == 0 || warp == 1 ? 0 : 1);
  uint64_t* tb = bars + team * T8_PER_TEAM;
  uint8_t* sQ = smem + team * p.stage_bytes;     // two Q buffers of 16 KB
  uint8_t* sK = sQ + 2 * 128 * 128;
  uint8_t* sV = sK + p.Lk * 128;
  const uint32_t tmem = tmem_base + team * 256;
  const int first = blockIdx.x + team * gridDim.x, stride = 2 * gridDim.x;

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------ TMA producer of this team
    if (lane == 0) {
      uint32_t uc = 0, tc = 0;
      for (int u = first; u < p.n_units; u += stride, ++uc) {
        const int h = u % p.heads, seq = u / p.heads;
        const int row_base = seq * p.L;
        mbar_wait(&tb[T8_KVFREE], (uc & 1) ^ 1);
        mbar_expect_tx(&tb[T8_KVFULL], 2 * p.Lk * 128);
        tma_load_2d(sK, &mapKV, &tb[T8_KVFULL], d + h * 64, row_base);
        tma_load_2d(sV, &mapKV, &tb[T8_KVFULL], 2 * d + h * 64, row_base);
        for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
          const int buf = tc & 1;
          mbar_wait(&tb[T8_QFREE + buf], ((tc >> 1) & 1) ^ 1);
          mbar_expect_tx(&tb[T8_QFULL + buf], 128 * 128);
          tma_load_2d(sQ + buf * 128 * 128, &mapQ, &tb[T8_QFULL + buf], h * 64, row_base + qt * 128);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------ MMA issuer of this team
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(128, p.Lk);
      const uint32_t idesc_o = umma_idesc_f16(128, 80) | (1u << 16);   // B = [V | ones], MN-major
      const uint64_t dk = umma_desc_k_sw128(smem_u32(sK));
      const uint64_t dv = tc8_desc_mn(smem_u32(sV), smem_u32(sOnes) - smem_u32(sV));
      const int ksteps = p.Lk >> 4;
      const int k_first = 2 * p.n_first;       // 16-key steps whose P lives in the first half's columns
      uint32_t uc = 0, tc = 0;
      for (int u = first; u < p.n_units; u += stride, ++uc) {
        mbar_wait(&tb[T8_KVFULL], uc & 1);
        for (int qt = 0; qt < p.n_qt; ++qt, ++tc) {
          const int buf = tc & 1;
          const uint32_t ph = tc & 1;
          const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ + buf * 128 * 128));
          mbar_wait(&tb[T8_QFULL + buf], (tc >> 1) & 1);
          mbar_wait(&tb[T8_TMEMFREE], ph ^ 1);  // the previous tile's O has been read out of TMEM
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss<1>(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(&tb[T8_SREADY]);
          mbar_wait(&tb[T8_PREADY], ph);        // all eight warps have written their P chunks
          tc_fence_after();
          for (int j = 0; j < ksteps; ++j) {
            const uint32_t pa = j < k_first ? tmem + 8 * j : tmem + kTc8ColP2 + 8 * (j - k_first);
            tc8_umma_ts(tmem + kTc8ColO, pa, dv + 128 * j, idesc_o, j != 0);
          }
          umma_commit(&tb[T8_OREADY]);
          umma_commit(&tb[T8_QFREE + buf]);
          if (qt == p.n_qt - 1) umma_commit(&tb[T8_KVFREE]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue: two warps per 32 query rows
    const int wi = (warp - 4) & 7;
    const int q4 = wi & 3;                       // TMEM lane quarter (= warp % 4)
    const int half_id = wi >> 2;                 // 0: chunks [0, n_first), 1: chunks [n_first, n_chunks)
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const int ch_begin = half_id == 0 ? 0 : p.n_first;
    const int ch_end = half_id == 0 ? p.n_first : p.n_chunks;
    const int pair_bar = 1 + team * 4 + q4;      // named barrier of the two warps that share these 32 rows
    // token per lane quarter: only one team's warps of a sub-partition are in their exponential pass at a time, strictly
    // alternating (attention_tc.cu); every warp takes it once per tile, dead warps and the team with fewer tiles pass it on
    uint64_t* tok_mine = &tok[q4 * 2 + team];
    uint64_t* tok_other = &tok[q4 * 2 + (team ^ 1)];
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
        uint8_t* sQb = sQ + (tc & 1) * 128 * 128;
        const int q0 = qt * 128, qrow = q0 + r;
        const int key_end = p.causal ? min(p.L, qrow + 1) : p.L;  // keys [0, key_end) are visible to this row
        const bool warp_live = q0 + q4 * 32 < p.L;                 // rows of a dead warp pair are never stored
        const int warp_min_end = p.causal ? min(p.L, q0 + q4 * 32 + 1) : p.L;
        const int full_chunks = min(p.n_chunks, warp_min_end >> 5);
        float* pm = pmax + (((tc & 1) * 2 + team) * 2) * 128;     // [half][128] of this tile parity and team
        mbar_wait(&tb[T8_SREADY], ph);
        tc_fence_after();
        if (half_id == 0 && tc > 0 && lane == 0) {   // the previous tile's O store has left its Q buffer
          bulk_wait_read_all();
          mbar_arrive(&tb[T8_QFREE + ((tc - 1) & 1)]);
        }
        float m = -INFINITY;
        bool tok_held = false;
        if (warp_live) {
          // ---- pass 1: maximum over this warp's chunks, then over the pair.  One chunk buffer only: with 640 threads a
          // thread has 102 registers, and the four softmax warps per sub-partition hide the TMEM load latency by
          // themselves (the double-buffered form spilled 130 registers' worth)
          for (int ch = ch_begin; ch < ch_end; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(trow + ch * 32, v);
            tmem_ld_wait();
            m = tc8_chunk_max(v, m, ch, ch < full_chunks, key_end);
          }
          pm[half_id * 128 + r] = m;
          tc8_pair_sync(pair_bar);
          m = fmaxf(m, pm[(half_id ^ 1) * 128 + r]);
          // ---- pass 2: p = 2^(s*c - m*c) of this warp's chunks, packed fp16 P: the first half into S columns the
          // row pair's FIRST warp has consumed, the second half into the columns behind S
          const float mc = m * c;
          if (p.use_tok) {
            mbar_wait(tok_mine, team == 0 ? ((tc & 1) ^ 1) : (tc & 1));
            tok_held = true;
          }
          for (int ch = ch_begin; ch < ch_end; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(trow + ch * 32, v);
            tmem_ld_wait();
            tc8_chunk_exp(v, half_id == 0 ? trow + ch * 16 : trow + kTc8ColP2 + (ch - p.n_first) * 16, ch,
                          ch < full_chunks, key_end, c, mc);
          }
          tc8_st_wait();
        }
        if (p.use_tok && !tok_held) mbar_wait(tok_mine, team == 0 ? ((tc & 1) ^ 1) : (tc & 1));   // dead warp: pass it on
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (p.use_tok) mbar_arrive(tok_other);
          mbar_arrive(&tb[T8_PREADY]);
        }
        // ---- epilogue: this warp's 32 of the 64 output columns
        mbar_wait(&tb[T8_OREADY], ph);
        tc_fence_after();
        if (warp_live) {
          uint32_t o[32], ls[16];
          tmem_ld_32x32(trow + kTc8ColO + half_id * 32, o);
          tc8_ld16(trow + kTc8ColO + 64, ls);        // column 64 = sum_j P_j (the block of ones)
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tb[T8_TMEMFREE]);
          const float l = __uint_as_float(ls[0]);
          const float inv = 1.f / l;
          uint4 ov[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __half2 h0 = __floats2half2_rn(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
            __half2 h1 = __floats2half2_rn(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
            __half2 h2 = __floats2half2_rn(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
            __half2 h3 = __floats2half2_rn(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
            ov[j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                               *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          }
          // stage this warp's half of the 32 x 64 O block in the tile's (retired) Q buffer, 128-byte swizzled like a
          // TMA box; the pair's first warp issues one TMA store for the block (rows >= L clipped by the tensor map)
          uint4* srow = reinterpret_cast<uint4*>(sQb + r * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) srow[(half_id * 4 + j) ^ (r & 7)] = ov[j];
          fence_proxy_async();
          tc8_pair_sync(pair_bar);
          if (half_id == 0) {
            if (lane == 0) {
              tma_store_3d(&mapO, sQb + q4 * 32 * 128, h * 64, q0 + q4 * 32, seq);
              bulk_commit_group();
            }
            if (qrow < p.L && p.lse != nullptr)
              p.lse[(static_cast<size_t>(seq) * p.heads + h) * p.L + qrow] = m * 0.125f + logf(l);
          }
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tb[T8_TMEMFREE]);
        }
      }
    }
    if (half_id == 0 && lane == 0) bulk_wait_read_all();  // shared memory must outlive the last O store's read
    if (p.use_tok) {
      for (; tc < n_tok; ++tc) {                           // keep the other team's token moving
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

static int tc8_tmap_rows64(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows) {
  static PFN_encodeTiled encode = get_encode_tiled();
  if (encode == nullptr) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled(attention tc8) failed (%d)", static_cast<int>(r));
  return 0;
}

// Returns -1 when the shape is outside what this kernel covers (the caller uses attention_fwd_tc).
int attention_fwd_tc8(const __half* qkv, int n_seq, int L, int heads, int causal, __half* out, float* lse,
                      cudaStream_t stream) {
  const int Lk = (L + 15) / 16 * 16;
  if (Lk > 208 || Lk < 16) return -1;
  if (((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return -1;
  AttnTc8Args a{};
  a.L = L; a.Lk = Lk; a.heads = heads; a.causal = causal; a.out = out; a.lse = lse;
  a.n_chunks = (Lk + 31) / 32;
  a.n_first = (a.n_chunks + 1) / 2;
  if (a.n_chunks - a.n_first > 3 || 16 * a.n_first > kTc8ColO) return -1;   // second half's P: 48 columns behind S
  a.n_qt = (L + 127) / 128;
  a.n_units = heads * n_seq;
  a.stage_bytes = 2 * 128 * 128 + 2 * Lk * 128;
  // measured at 512 x 197 x 12: 218 us without the token, 200 us with it (RLCF_ATTN_TC8_TOKEN=0 disables it)
  static const int use_tok = getenv("RLCF_ATTN_TC8_TOKEN") != nullptr ? atoi(getenv("RLCF_ATTN_TC8_TOKEN")) : 1;
  a.use_tok = use_tok;
  const size_t smem = 1024 + 2 * static_cast<size_t>(a.stage_bytes) + static_cast<size_t>(Lk) * 128 +
                      2 * 2 * 2 * 128 * sizeof(float) + (2 * T8_PER_TEAM + 8) * 8 + 16;
  if (smem > 227 * 1024) return -1;
  static DynSmemState st;
  if (cudaError_t e = ensure_dyn_smem(attn_fwd_tc8_kernel, smem, st))
    return set_error(RLCF_ERR_CUDA, "attention_fwd_tc8 attr: %s", cudaGetErrorString(e));
  CUtensorMap mq, mkv, mo;
  const long long rows = static_cast<long long>(n_seq) * L;
  if (int rc = tc8_tmap_rows64(&mq, qkv, rows, 3 * heads * 64, 128)) return rc;
  if (int rc = tc8_tmap_rows64(&mkv, qkv, rows, 3 * heads * 64, Lk)) return rc;
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
    if (r != CUDA_SUCCESS) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled(attention tc8 out) failed (%d)", static_cast<int>(r));
  }
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  attn_fwd_tc8_kernel<<<grid, kTc8Threads, smem, stream>>>(mq, mkv, mo, a);
  RLCF_CHECK_LAUNCH("attention_fwd_tc8");
  return 0;
}

}  // namespace rlcf
