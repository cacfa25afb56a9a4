// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = A[M,K] (fp16, K-major) x B[N,K]^T (fp16, K-major), fp32 accumulation in TMEM,
//   fused epilogues (bias / QuickGELU / residual / QuickGELU-backward).
//
// This one kernel carries every dense contraction of the RLCF hot path (SURVEY.md 2.2 rows K1,K3,K5,K6,K7
// and their dgrad counterparts in K12): the reference reaches them through nn.Conv2d / nn.MultiheadAttention
// in-proj+out-proj / nn.Linear (TPT/clip/model.py:175-181,224) and autograd.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a kStages-deep smem ring
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (UMMA 128x256x16 or, as a CTA pair, 256x256x16)
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       : tcgen05.ld accumulator -> registers -> smem transpose -> coalesced global I/O
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  kCtaGroup == 2 pairs two SMs (cta_group::2): each CTA stages its own 128 rows of A
// and half (128 rows) of B, halving the shared-memory and L2 traffic per FLOP.
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

constexpr int kBM = 128;      // accumulator rows per CTA (TMEM lanes)
constexpr int kBN = 256;      // accumulator columns per tile
constexpr int kBK = 64;       // K elements per stage (= 128 bytes = one swizzle atom)
constexpr int kUmmaK = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int kGemmThreads = 384;   // 4 control warps + 8 epilogue warps
constexpr int kBarrierBytes = 256;
constexpr int kStagingBytes = 8 * 4096;  // epilogue transpose buffers, 32 x 32 fp32 per epilogue warp

template <int kCtaGroup>
struct GemmCfg {
  static constexpr int kBRows = kBN / kCtaGroup;                 // B rows staged by each CTA
  static constexpr int kABytes = kBM * kBK * 2;                  // 16 KB
  static constexpr int kBBytes = kBRows * kBK * 2;               // 32 KB / 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kCtaGroup == 1 ? 4 : 6;         // 192 KB ring either way
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + kBarrierBytes + kStagingBytes;
};

struct GemmArgs {
  int M, N, K;           // per group
  int G;                 // independent problems (groups); group g uses A/B/out/bias advanced by g group strides
  long long out_gs;      // elements between consecutive groups of out / resid / aux
  long long bias_gs;     // elements between consecutive groups of bias
  int epi;
  const float* bias;     // [N] or null
  const float* resid;    // EPI_RESID_F32: [M, ldo] fp32 (may alias out)
  const __half* aux_in;  // EPI_GELU_BWD_F16: pre-activation u [M, ldo] fp16
  __half* aux_out;       // EPI_GELU_F16: optional pre-activation save [M, ldo] fp16
  void* out;             // fp16 or fp32 [M, ldo]
  int ldo;               // leading dimension of out / resid / aux (elements)
  float alpha;           // scale applied to the accumulator before the epilogue
  int debug_nostore;     // RLCF_GEMM_DEBUG_NOSTORE=1: skip the epilogue's global traffic (timing probe only)
  int resid_prefetch;    // EPI_RESID_F32: the producer prefetches each tile's residual rows into L2 (RLCF_GEMM_RESID_PREFETCH=1; slower)
  AdamwEpi opt;          // EPI_ADAMW only
};

__device__ __forceinline__ float adamw_elem(float p0, float g, float& m, float& v, const AdamwEpi& o) {
  return adamw_update(p0, g, m, v, o.lr, o.b1, o.b2, o.eps, o.wd, o.bc1, o.bc2_sqrt);   // ptx.cuh: shared with adamw_full
}

// kMc = CTA pairs per cluster (cta_group::2 only).  With kMc == 2 a 4-CTA cluster computes a 512 x 256 super tile:
// the two pairs share the B (weight) tile, which pair 0 loads once and TMA-multicasts into both pairs' shared
// memory -- 25 % fewer L2->SM operand bytes per FLOP, which is what bounds the K = 768 shapes.
template <int kCtaGroup, int kEpi, int kMc>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs p) {
  using Cfg = GemmCfg<kCtaGroup>;
  static_assert(kMc == 1 || kCtaGroup == 2, "multicast needs CTA pairs");
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled UMMA/TMA tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cluster_rank = kCtaGroup == 2 ? cluster_ctarank() : 0;
  const uint32_t pair = cluster_rank >> 1;      // CTA pair inside the cluster
  const uint32_t cta_rank = cluster_rank & 1;   // rank inside the pair
  const bool leader = cta_rank == 0;

  const int tile_m = kBM * kCtaGroup * kMc;     // rows per cluster tile
  const int m_tiles = (p.M + tile_m - 1) / tile_m;
  const int n_tiles = (p.N + kBN - 1) / kBN;
  const int tiles_per_group = m_tiles * n_tiles;
  const int num_tiles = p.G * tiles_per_group;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int worker = blockIdx.x / (kCtaGroup * kMc);
  const int num_workers = gridDim.x / (kCtaGroup * kMc);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kMc);   // one tcgen05.commit arrive per pair that reads this stage
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8 * kCtaGroup);  // one arrive per epilogue warp per CTA
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kCtaGroup>(tmem_slot, 512);
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // In a CTA pair every load signals the leader's barrier (same smem offset, peer bit cleared).
      const uint32_t peer_mask = 0xFEFFFFFFu;
      for (int t = worker; t < num_tiles; t += num_workers) {
        const int g = t / tiles_per_group, tg = t - g * tiles_per_group;
        const int m_blk = tg / n_tiles, n_blk = tg % n_tiles;
        const int m0 = m_blk * tile_m + static_cast<int>(pair) * kBM * kCtaGroup + static_cast<int>(cta_rank) * kBM;
        const int n0 = n_blk * kBN + static_cast<int>(cta_rank) * Cfg::kBRows;
        // EPI_RESID_F32, opt-in experiment (RLCF_GEMM_RESID_PREFETCH=1): the epilogue of this tile reads 128 rows x 1 KB
        // of the fp32 residual stream, 4 KB per warp at a time with a DRAM round trip in front of every chunk.  Here the
        // producer asks L2 for those rows while it streams the operands (one 1 KB cp.async.bulk.prefetch.L2 per row, spread
        // over the first half of the K loop).  MEASURED SLOWER and therefore off by default: out_proj 635 -> 716 us,
        // c_proj 1462 -> 1782 us at M = 403 456 (profiles/r2_gemm_resid_prefetch.txt) -- the K = 768 main loop is already
        // L2-bandwidth-bound, and the prefetched lines compete with the operand tiles for it.
        int pf_row = 0, pf_rows = 0, pf_per_kb = 0;
        uint32_t pf_bytes = 0;
        const float* pf_base = nullptr;
        if constexpr (kEpi == EPI_RESID_F32) {
          if (p.resid_prefetch && m0 < p.M) {
            pf_rows = min(kBM, p.M - m0);
            pf_bytes = static_cast<uint32_t>(min(kBN, p.N - n_blk * kBN)) * 4u;
            pf_per_kb = (kBM + max(1, num_kb / 2) - 1) / max(1, num_kb / 2);
            pf_base = p.resid + static_cast<size_t>(g * p.out_gs) + static_cast<size_t>(m0) * p.ldo + n_blk * kBN;
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          if constexpr (kEpi == EPI_RESID_F32) {
            const int pf_end = min(pf_rows, pf_row + pf_per_kb);
            for (; pf_row < pf_end; ++pf_row) l2_prefetch_bulk(pf_base + static_cast<size_t>(pf_row) * p.ldo, pf_bytes);
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if constexpr (kCtaGroup == 1) {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_3d(sa, &tmA, &full_bar[stage], kb * kBK, m0, g);
            tma_load_3d(sb, &tmB, &full_bar[stage], kb * kBK, n0, g);
          } else {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const uint32_t bar = smem_u32(&full_bar[stage]) & peer_mask;
            tma_load_3d_cg2(sa, &tmA, bar, kb * kBK, m0, g);
            if constexpr (kMc == 1) {
              tma_load_3d_cg2(sb, &tmB, bar, kb * kBK, n0, g);
            } else if (pair == 0) {
              // this CTA's half of the B tile goes to the CTAs of the same in-pair rank in both pairs
              tma_load_3d_cg2_mc(sb, &tmB, bar, kb * kBK, n0, g, static_cast<uint16_t>(0b0101u << cta_rank));
            }
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kBM * kCtaGroup, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = worker; t < num_tiles; t += num_workers, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 32 bytes (16 fp16) along K inside the swizzle atom: +2 in the >>4 address field
            umma_f16_ss<kCtaGroup>(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          // the smem slot is released to every producer that writes into it (all CTAs of the cluster when the
          // B tile is multicast); the accumulator-ready signal goes to this pair's two CTAs only
          if constexpr (kCtaGroup == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_cg2(&empty_bar[stage], static_cast<uint16_t>((1u << (2 * kMc)) - 1));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (kCtaGroup == 1) umma_commit(&tfull_bar[as]);
        else umma_commit_cg2(&tfull_bar[as], static_cast<uint16_t>(0b11u << (2 * pair)));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (8 warps)
    // Warp w reads TMEM lane quarter (w & 3) and column half ((w - 4) >> 2) of the 128 x 256 accumulator in four
    // 32-column chunks.  Each chunk goes registers -> swizzled smem staging (4 KB per warp) -> registers in a
    // transposed assignment where 8 consecutive lanes cover 128 contiguous bytes of one output row, so that every
    // global access (residual / pre-activation read, output write) is fully coalesced.
    const int ew = warp & 3;
    const int half_id = (warp - 4) >> 2;
    float4* stg = reinterpret_cast<float4*>(smem + Cfg::kStages * Cfg::kStageBytes + kBarrierBytes) + (warp - 4) * 256;
    const int tr = lane >> 3;  // row within a group of 4 rows (transposed phase)
    const int tj = lane & 7;   // float4 column chunk (transposed phase)
    int it = 0;
    for (int t = worker; t < num_tiles; t += num_workers, ++it) {
      const int g = t / tiles_per_group, tg = t - g * tiles_per_group;
      const int m_blk = tg / n_tiles, n_blk = tg % n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * tile_m + static_cast<int>(pair) * kBM * kCtaGroup + static_cast<int>(cta_rank) * kBM + ew * 32;
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * kBN + half_id * 128;
      const int col_base = n_blk * kBN + half_id * 128;
      const int n_chunks = max(0, min(4, (p.N - col_base + 31) / 32));  // warp-uniform
      // This warp's 128 bias values live in registers (4 per lane) and are broadcast by shuffle: with ~230 KB of
      // shared memory per CTA there is no L1 left, so per-chunk bias loads would each pay an L2 round trip.
      float4 breg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr && col_base + 4 * lane < p.N)
        breg = __ldg(reinterpret_cast<const float4*>(p.bias + g * p.bias_gs + col_base) + lane);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (n_chunks == 0) {  // this warp's column half lies entirely beyond N: nothing to read, release at once
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtaGroup == 1) mbar_arrive(&tempty_bar[as]);
          else mbar_arrive_cluster(&tempty_bar[as], 2 * pair);
        }
      }
      // per-lane global offsets of the coalesced phase: row (row0 + tr + 4 i), columns col_base + 32 c + 4 tj .. +3
      const size_t goff0 = static_cast<size_t>(g * p.out_gs) + static_cast<size_t>(row0 + tr) * p.ldo + col_base + tj * 4;
      const size_t gstep = static_cast<size_t>(4) * p.ldo;
      const int rows_left = p.M - (row0 + tr);  // row i of this lane is valid iff 4 i < rows_left
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const size_t goff = goff0 + c * 32;
        // prefetch what the coalesced phase needs from global memory; the latency overlaps the TMEM load
        float4 z[kEpi == EPI_RESID_F32 ? 8 : 1];
        uint2 zu[kEpi == EPI_GELU_BWD_F16 ? 8 : 1];
        if constexpr (kEpi == EPI_RESID_F32) {
          const float* rp = p.resid + goff;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            z[i] = 4 * i < rows_left ? *reinterpret_cast<const float4*>(rp + i * gstep) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else if constexpr (kEpi == EPI_GELU_BWD_F16) {
          const __half* ap = p.aux_in + goff;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            zu[i] = 4 * i < rows_left ? *reinterpret_cast<const uint2*>(ap + i * gstep) : make_uint2(0u, 0u);
        }
        uint32_t r[32];
        tmem_ld_32x32(tacc + c * 32, r);
        tmem_ld_wait();
        if (c == n_chunks - 1) {
          // last TMEM read of this tile by this warp: hand the accumulator back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtaGroup == 1) mbar_arrive(&tempty_bar[as]);
            else mbar_arrive_cluster(&tempty_bar[as], 2 * pair);
          }
        }
        if (p.debug_nostore) { __syncwarp(); continue; }  // bring-up probe: main loop + TMEM drain only
        // phase 1 (thread = row): stage the raw accumulator chunk
#pragma unroll
        for (int j = 0; j < 8; ++j)
          stg[lane * 8 + (j ^ (lane & 7))] =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        // bias of this lane's 4 columns in the coalesced phase
        float4 b4;
        b4.x = __shfl_sync(0xffffffffu, breg.x, c * 8 + tj);
        b4.y = __shfl_sync(0xffffffffu, breg.y, c * 8 + tj);
        b4.z = __shfl_sync(0xffffffffu, breg.z, c * 8 + tj);
        b4.w = __shfl_sync(0xffffffffu, breg.w, c * 8 + tj);
        __syncwarp();
        // phase 2 (8 lanes = one 128-byte row segment): epilogue math + coalesced global stores
        if constexpr (kEpi == EPI_ADAMW) {
          // The accumulator chunk is the (loss-scaled) gradient of 32 columns x 32 rows of one sample's weight.  Optimizer
          // step on the fp32 master tile in place, 4 rows at a time so that 12 independent 16-byte loads are in flight
          // per lane (master, exp_avg, exp_avg_sq) before the first dependent instruction.
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 pw[4], mm[4], vv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = 4 * h + j;
              pw[j] = mm[j] = vv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (4 * i < rows_left) {
                const size_t e = goff + i * gstep;
                const size_t e_in = e - static_cast<size_t>(g * p.out_gs) + static_cast<size_t>(g * p.opt.p_in_gs);
                pw[j] = *reinterpret_cast<const float4*>(p.opt.p_in + e_in);
                if (!p.opt.fresh) {
                  mm[j] = *reinterpret_cast<const float4*>(p.opt.m + e);
                  vv[j] = *reinterpret_cast<const float4*>(p.opt.v + e);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = 4 * h + j;
              const int rr = tr + 4 * i;
              float4 q = stg[rr * 8 + (tj ^ (rr & 7))];
              q.x = __fmul_rn(q.x, p.alpha); q.y = __fmul_rn(q.y, p.alpha);
              q.z = __fmul_rn(q.z, p.alpha); q.w = __fmul_rn(q.w, p.alpha);
              if (4 * i < rows_left) {
                const size_t e = goff + i * gstep;
                float4 w;
                w.x = adamw_elem(pw[j].x, q.x, mm[j].x, vv[j].x, p.opt); w.y = adamw_elem(pw[j].y, q.y, mm[j].y, vv[j].y, p.opt);
                w.z = adamw_elem(pw[j].z, q.z, mm[j].z, vv[j].z, p.opt); w.w = adamw_elem(pw[j].w, q.w, mm[j].w, vv[j].w, p.opt);
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + e) = w;
                *reinterpret_cast<float4*>(p.opt.m + e) = mm[j];
                *reinterpret_cast<float4*>(p.opt.v + e) = vv[j];
                if (p.opt.w16 != nullptr) {
                  const size_t e16 = e - static_cast<size_t>(g * p.out_gs) + static_cast<size_t>(g * p.opt.w16_gs);
                  __half2 h0 = __floats2half2_rn(w.x, w.y), h1 = __floats2half2_rn(w.z, w.w);
                  *reinterpret_cast<uint2*>(p.opt.w16 + e16) =
                      make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
                }
              }
            }
          }
        } else {
          // All eight row groups of this lane go through each stage together, without a branch in between: the
          // eight independent load -> math -> convert chains overlap (a branch around each row's store used to
          // serialise them, which made the QuickGELU epilogue -- two dependent MUFU operations per value -- cost
          // 60 % more than the main loop of a K = 768 tile: 8.2 us per tile against 5.1; 6.7 us in this form, with the
          // plain fp16 epilogue at 6.1).  Only the stores are predicated.
          float4 q[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = tr + 4 * i;
            q[i] = stg[rr * 8 + (tj ^ (rr & 7))];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            q[i].x = fmaf(q[i].x, p.alpha, b4.x); q[i].y = fmaf(q[i].y, p.alpha, b4.y);
            q[i].z = fmaf(q[i].z, p.alpha, b4.z); q[i].w = fmaf(q[i].w, p.alpha, b4.w);
          }
          if constexpr (kEpi == EPI_RESID_F32 || kEpi == EPI_F32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if constexpr (kEpi == EPI_RESID_F32) { q[i].x += z[i].x; q[i].y += z[i].y; q[i].z += z[i].z; q[i].w += z[i].w; }
              if (4 * i < rows_left) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + goff + i * gstep) = q[i];
            }
          } else {
            if constexpr (kEpi == EPI_GELU_F16) {
              if (p.aux_out != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  __half2 u0 = __floats2half2_rn(q[i].x, q[i].y), u1 = __floats2half2_rn(q[i].z, q[i].w);
                  if (4 * i < rows_left)
                    *reinterpret_cast<uint2*>(p.aux_out + goff + i * gstep) =
                        make_uint2(*reinterpret_cast<uint32_t*>(&u0), *reinterpret_cast<uint32_t*>(&u1));
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                q[i].x = quick_gelu(q[i].x); q[i].y = quick_gelu(q[i].y);
                q[i].z = quick_gelu(q[i].z); q[i].w = quick_gelu(q[i].w);
              }
            } else if constexpr (kEpi == EPI_GELU_BWD_F16) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 ua = __half22float2(*reinterpret_cast<const __half2*>(&zu[i].x));
                const float2 ub = __half22float2(*reinterpret_cast<const __half2*>(&zu[i].y));
                q[i].x *= quick_gelu_grad(ua.x); q[i].y *= quick_gelu_grad(ua.y);
                q[i].z *= quick_gelu_grad(ub.x); q[i].w *= quick_gelu_grad(ub.y);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __half2 h0 = __floats2half2_rn(q[i].x, q[i].y), h1 = __floats2half2_rn(q[i].z, q[i].w);
              if (4 * i < rows_left)
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + goff + i * gstep) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
            }
          }
        }
        __syncwarp();  // staging buffer is reused by the next chunk; also reconverges for the .aligned TMEM load
      }
    }
  }

  __syncwarp();  // role branches diverge inside warps 0/1; the barriers below are .aligned
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<kCtaGroup>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- host side
// Rank-3 map (k, row, group): rows beyond `rows` inside a group read as zeros, so tiles never cross a group boundary.
static int make_tmap_f16(CUtensorMap* map, const void* base, int rows, int cols, int ld_elems, int groups,
                         long long gs_elems, int box_rows) {
  static PFN_encodeTiled encode = get_encode_tiled();
  if (encode == nullptr) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld_elems % 8) != 0)
    return set_error(RLCF_ERR_ARG, "gemm operand must be 16-byte aligned with a leading dimension multiple of 8");
  if (groups == 1) gs_elems = static_cast<long long>(rows) * ld_elems;
  if (gs_elems <= 0 || (gs_elems % 8) != 0)
    return set_error(RLCF_ERR_ARG, "gemm group stride must be a positive multiple of 8 elements");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld_elems) * 2, static_cast<cuuint64_t>(gs_elems) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(RLCF_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}

template <int kCtaGroup, int kEpi, int kMc>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, cudaStream_t stream) {
  using Cfg = GemmCfg<kCtaGroup>;
  constexpr int kClusterCtas = kCtaGroup * kMc;
  // per device (the attribute and the co-residency of clusters are properties of the device the launch goes to)
  static DynSmemState st;
  static int max_clusters_dev[32] = {};   // co-resident clusters: the kernel is persistent, so the grid must not exceed it
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 32) dev = 31;
  int& max_clusters = max_clusters_dev[dev];
  const bool configured = max_clusters > 0 && dev != 31;
  if (!configured) {
    cudaError_t e = ensure_dyn_smem(gemm_f16_kernel<kCtaGroup, kEpi, kMc>, Cfg::kSmemBytes, st);
    if (e != cudaSuccess) return set_error(RLCF_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
    max_clusters = sm_count() / kClusterCtas;
    if (kClusterCtas > 2) {
      // GPCs do not all hold a multiple of 4 SMs; ask the driver how many 4-CTA clusters fit at once
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(max_clusters * kClusterCtas);
      q.blockDim = dim3(kGemmThreads);
      q.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = kClusterCtas;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, gemm_f16_kernel<kCtaGroup, kEpi, kMc>, &q);
      if (e != cudaSuccess) return set_error(RLCF_ERR_CUDA, "cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e));
      if (n > 0 && n < max_clusters) max_clusters = n;
      if (getenv("RLCF_GEMM_VERBOSE")) fprintf(stderr, "rlcf gemm: %d-CTA clusters, %d co-resident\n", kClusterCtas, n);
    }
  }
  const int tile_m = kBM * kClusterCtas;
  const int tiles = args.G * ((args.M + tile_m - 1) / tile_m) * ((args.N + kBN - 1) / kBN);
  int workers = tiles < max_clusters ? tiles : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(workers * kClusterCtas);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClusterCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_f16_kernel<kCtaGroup, kEpi, kMc>, ta, tb, args);
  if (e != cudaSuccess) return set_error(RLCF_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

int gemm_f16(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, int epi, float alpha,
             const float* bias, const float* resid, const __half* aux_in, __half* aux_out, void* out, int ldo,
             cudaStream_t stream) {
  return gemm_f16_grouped(A, lda, 0, B, ldb, 0, 1, M, N, K, epi, alpha, bias, 0, resid, aux_in, aux_out, out, ldo, 0,
                          stream);
}

int gemm_f16_grouped(const __half* A, int lda, long long a_gs, const __half* B, int ldb, long long b_gs, int G, int M,
                     int N, int K, int epi, float alpha, const float* bias, long long bias_gs, const float* resid,
                     const __half* aux_in, __half* aux_out, void* out, int ldo, long long out_gs,
                     cudaStream_t stream, const AdamwEpi* adamw) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(RLCF_ERR_ARG, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  if ((epi == EPI_ADAMW) != (adamw != nullptr)) return set_error(RLCF_ERR_ARG, "gemm: EPI_ADAMW needs optimizer state");
  if (G <= 0) return set_error(RLCF_ERR_ARG, "gemm: G=%d groups", G);
  if (G > 1 && (out_gs % 8 != 0 || bias_gs % 4 != 0 || out_gs < 0 || bias_gs < 0))
    return set_error(RLCF_ERR_ARG, "gemm: group strides of out (%lld) / bias (%lld) must be multiples of 8 / 4", out_gs,
                     bias_gs);
  if (N % 32 != 0) return set_error(RLCF_ERR_ARG, "gemm: N=%d must be a multiple of 32", N);
  if (K % 8 != 0) return set_error(RLCF_ERR_ARG, "gemm: K=%d must be a multiple of 8", K);
  if (ldo % 8 != 0) return set_error(RLCF_ERR_ARG, "gemm: ldo=%d must be a multiple of 8", ldo);
  if (epi < 0 || epi >= EPI_COUNT) return set_error(RLCF_ERR_ARG, "gemm: unknown epilogue %d", epi);
  if (epi == EPI_RESID_F32 && resid == nullptr) return set_error(RLCF_ERR_ARG, "gemm: residual epilogue needs resid");
  if (epi == EPI_GELU_BWD_F16 && aux_in == nullptr) return set_error(RLCF_ERR_ARG, "gemm: gelu-bwd needs aux_in");
  const int cg = gemm_cta_group();
  CUtensorMap ta, tb;
  if (int rc = make_tmap_f16(&ta, A, M, K, lda, G, a_gs, kBM)) return rc;
  if (int rc = make_tmap_f16(&tb, B, N, K, ldb, G, b_gs, kBN / cg)) return rc;
  static const int debug_nostore = getenv("RLCF_GEMM_DEBUG_NOSTORE") != nullptr;
  static const int resid_prefetch = getenv("RLCF_GEMM_RESID_PREFETCH") != nullptr && atoi(getenv("RLCF_GEMM_RESID_PREFETCH")) != 0;
  // bulk prefetches need 16-byte aligned rows: resid rows are ldo floats apart (ldo % 8 == 0) from a 16-byte aligned base
  const int pf = resid_prefetch && epi == EPI_RESID_F32 && (reinterpret_cast<uintptr_t>(resid) & 15) == 0;
  GemmArgs args{M, N, K, G, G > 1 ? out_gs : 0, G > 1 ? bias_gs : 0, epi, bias, resid, aux_in, aux_out, out, ldo,
                alpha, debug_nostore, pf, adamw != nullptr ? *adamw : AdamwEpi{}};
  if (G == 1) { args.opt.p_in_gs = 0; args.opt.w16_gs = 0; }
  // multicast pays once there are at least two 256-row tiles per cluster slot; tiny problems keep 2-CTA clusters
  const bool mc = cg == 2 && gemm_multicast() && M > 2 * kBM * 2;
  switch (epi) {
#define RLCF_GEMM_CASE(E)                                              \
  case E:                                                              \
    if (mc) return launch_gemm<2, E, 2>(ta, tb, args, stream);         \
    return cg == 2 ? launch_gemm<2, E, 1>(ta, tb, args, stream) : launch_gemm<1, E, 1>(ta, tb, args, stream);
    RLCF_GEMM_CASE(EPI_F16)
    RLCF_GEMM_CASE(EPI_GELU_F16)
    RLCF_GEMM_CASE(EPI_RESID_F32)
    RLCF_GEMM_CASE(EPI_GELU_BWD_F16)
    RLCF_GEMM_CASE(EPI_F32)
    RLCF_GEMM_CASE(EPI_ADAMW)
#undef RLCF_GEMM_CASE
  }
  return set_error(RLCF_ERR_ARG, "gemm: unknown epilogue %d", epi);
}

}  // namespace rlcf
