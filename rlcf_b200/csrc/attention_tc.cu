// tcgen05 fused attention forward for CLIP towers (head_dim 64, L <= 320 tokens).
//
// One CTA = 128 query rows of one (sequence, head).  The whole key/value range of the sequence is resident in
// shared memory (L <= 257 for every configured tower), so there is no KV loop and no online-softmax rescaling:
//   TMA     : Q tile [128 x 64], K and V tiles [Lk x 64] (Lk = L rounded up to 16), 128-byte swizzle
//   UMMA #1 : S[128 x Lk] = Q K^T            (SS, both K-major; fp32 accumulator in TMEM columns [0, Lk))
//   softmax : two threads per query row each read half of the S row from TMEM (tcgen05.ld), max / exp2 / sum in
//             fp32, and write P as packed fp16 back into TMEM columns [0, Lk/2) (tcgen05.st)
//   UMMA #2 : O[128 x 64] = P V              (TS: A = P from TMEM, B = V from smem, MN-major)
//   epilogue: O / rowsum -> fp16 -> global, 128 contiguous bytes per row; optional log-sum-exp for the backward
// Two CTAs are co-resident per SM (<= 84 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the
// other's loads and MMAs.  Replaces the bmm-softmax-bmm of nn.MultiheadAttention (TPT/clip/model.py:185-187).
#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

struct AttnTcArgs {
  int L, Lk, heads, causal;
  int kv_box_rows, n_kv_boxes;  // TMA boxes covering the Lk key rows
  int n0, n1;                   // UMMA N of the one or two key chunks of S (n0 + n1 == Lk)
  int o_off;                    // TMEM column of the O accumulator (behind P, inside the dead S region)
  int tmem_cols;
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kMaxChunksPerGroup = 5;  // 32-column S chunks per thread: Lk <= 320
constexpr int kAttnThreads = 256;  // two threads per query row: warps 0-3 and 4-7 split the key columns / O columns

__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV, AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 128 * 128;
  uint8_t* sV = sK + p.Lk * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.Lk * 128);  // [0] loads, [1] S ready, [2] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* xchg = reinterpret_cast<float*>(bars + 4);               // [2][128] partial row max, [2][128] partial row sum

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = warp >> 2;            // 0: first half of the key chunks / O columns [0,32); 1: the rest
  const int r = tid & 127;              // query row inside the tile (= TMEM lane)
  const int q0 = blockIdx.x * 128, h = blockIdx.y, seq = blockIdx.z;
  const int d = p.heads * 64;
  const int row_base = seq * p.L;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapKV);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<1>(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_expect_tx(&bars[0], 128 * 128 + 2 * p.Lk * 128);
    tma_load_2d(sQ, &mapQ, &bars[0], h * 64, row_base + q0);
    for (int b = 0; b < p.n_kv_boxes; ++b) {
      tma_load_2d(sK + b * p.kv_box_rows * 128, &mapKV, &bars[0], d + h * 64, row_base + b * p.kv_box_rows);
      tma_load_2d(sV + b * p.kv_box_rows * 128, &mapKV, &bars[0], 2 * d + h * 64, row_base + b * p.kv_box_rows);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    // S = Q K^T, one UMMA chain per key chunk (N <= 256 per instruction)
    const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
    int n_off = 0;
    for (int ch = 0; ch < 2; ++ch) {
      const int n = ch == 0 ? p.n0 : p.n1;
      if (n == 0) break;
      const uint64_t dk = umma_desc_k_sw128(smem_u32(sK) + n_off * 128);
      const uint32_t idesc = umma_idesc_f16(128, n);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss<1>(tmem + n_off, dq + 2 * k, dk + 2 * k, idesc, k != 0);
      n_off += n;
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();

  // ------------------------------------------------------------ softmax: two threads per query row
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int qrow = q0 + r;  // query index inside the sequence
  const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int n_chunks = (p.Lk + 31) >> 5;
  const int split = (n_chunks + 1) >> 1;
  const int ch_begin = grp == 0 ? 0 : split, ch_end = grp == 0 ? split : n_chunks;
  const int key_end = p.causal ? min(p.L, qrow + 1) : p.L;  // keys [0, key_end) are visible to this row
  // a warp whose 32 rows all lie beyond the sequence does no softmax work (its rows are never stored)
  const bool warp_live = q0 + (warp & 3) * 32 < p.L;
  float m = -INFINITY;
  if (warp_live) {
    for (int ch = ch_begin; ch < ch_end; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32(trow + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, (ch * 32 + j < key_end) ? __uint_as_float(v[j]) : -INFINITY);
    }
  }
  xchg[grp * 128 + r] = m;
  __syncthreads();
  m = fmaxf(m, xchg[(grp ^ 1) * 128 + r]);
  const float mc = m * c;
  float l = 0.f;
  // P chunk ch (packed fp16) goes to TMEM columns [16 ch, 16 ch + 16).  For group 0 these are S columns the same
  // thread has already consumed.  Group 1's P columns overlap S chunks that group 0 may still be reading, so
  // group 1 keeps its packed chunks in registers until group 0 has finished its second pass (barrier A).
  uint32_t pkbuf[kMaxChunksPerGroup][16];
  if (warp_live) {
#pragma unroll
    for (int i = 0; i < kMaxChunksPerGroup; ++i) {
      const int ch = ch_begin + i;
      if (ch < ch_end) {  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(trow + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = (ch * 32 + 2 * j < key_end) ? exp2f(fmaf(__uint_as_float(v[2 * j]), c, -mc)) : 0.f;
          const float b = (ch * 32 + 2 * j + 1 < key_end) ? exp2f(fmaf(__uint_as_float(v[2 * j + 1]), c, -mc)) : 0.f;
          // the row sum uses the fp16-rounded probabilities that the P V product actually sees
          const __half2 hp = __floats2half2_rn(a, b);
          const float2 back = __half22float2(hp);
          l += back.x + back.y;
          pkbuf[i][j] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        if (grp == 0) tmem_st_32x16(trow + ch * 16, pkbuf[i]);
      }
    }
    if (grp == 0) tmem_st_wait();
  }
  xchg[256 + grp * 128 + r] = l;
  __syncthreads();  // barrier A: group 0 has consumed all of its S columns
  if (grp == 1 && warp_live) {
#pragma unroll
    for (int i = 0; i < kMaxChunksPerGroup; ++i) {
      const int ch = ch_begin + i;
      if (ch < ch_end) tmem_st_32x16(trow + ch * 16, pkbuf[i]);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();  // barrier B: P complete in TMEM

  if (tid == 0) {
    tc_fence_after();
    // O = P V : A = P (TMEM, 8 columns per 16 keys), B = V rows [16 j, 16 j + 16) as an MN-major operand
    const uint32_t idesc = umma_idesc_f16(128, 64) | (1u << 16);
    const uint64_t dv = umma_desc_k_sw128(smem_u32(sV));
    const int ksteps = p.Lk >> 4;
    for (int j = 0; j < ksteps; ++j) umma_f16_ts(tmem + p.o_off, tmem + 8 * j, dv + 128 * j, idesc, j != 0);
    umma_commit(&bars[2]);
  }
  __syncwarp();
  l += xchg[256 + (grp ^ 1) * 128 + r];

  // ------------------------------------------------------------ epilogue: each thread stores 32 of the 64 columns
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  if (warp_live) {
    uint32_t o[32];
    tmem_ld_32x32(trow + p.o_off + grp * 32, o);
    tmem_ld_wait();
    if (qrow < p.L) {
      const float inv = 1.f / l;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(row_base + qrow)) * d + h * 64 + grp * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 h0 = __floats2half2_rn(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
        __half2 h1 = __floats2half2_rn(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
        __half2 h2 = __floats2half2_rn(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
        __half2 h3 = __floats2half2_rn(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
        dst[j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                            *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
      }
      if (p.lse != nullptr && grp == 0)
        p.lse[(static_cast<size_t>(seq) * p.heads + h) * p.L + qrow] = m * 0.125f + logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, p.tmem_cols);
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
  if (Lk > 32 * 2 * kMaxChunksPerGroup || (reinterpret_cast<uintptr_t>(qkv) & 15) != 0) return -1;
  AttnTcArgs a{};
  a.L = L; a.Lk = Lk; a.heads = heads; a.causal = causal; a.out = out; a.lse = lse;
  if (Lk <= 256) {
    a.kv_box_rows = Lk; a.n_kv_boxes = 1; a.n0 = Lk; a.n1 = 0;
  } else {
    a.kv_box_rows = Lk / 2; a.n_kv_boxes = 2;
    a.n0 = ((Lk / 2) + 15) / 16 * 16; a.n1 = Lk - a.n0;
  }
  a.o_off = ((Lk / 2) + 31) / 32 * 32;
  const int need = max(Lk, a.o_off + 64);
  a.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
  const size_t smem = 1024 + 128 * 128 + 2 * static_cast<size_t>(Lk) * 128 + 64 + 4 * 128 * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(RLCF_ERR_CUDA, "attention_fwd_tc attr: %s", cudaGetErrorString(e));
    configured = smem;
  }
  CUtensorMap mq, mkv;
  const long long rows = static_cast<long long>(n_seq) * L;
  if (int rc = make_tmap_rows64(&mq, qkv, rows, 3 * heads * 64, 128)) return rc;
  if (int rc = make_tmap_rows64(&mkv, qkv, rows, 3 * heads * 64, a.kv_box_rows)) return rc;
  dim3 grid((L + 127) / 128, heads, n_seq);
  attn_fwd_tc_kernel<<<grid, kAttnThreads, smem, stream>>>(mq, mkv, a);
  RLCF_CHECK_LAUNCH("attention_fwd_tc");
  return 0;
}

}  // namespace rlcf
