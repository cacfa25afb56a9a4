// Support kernels for full image-encoder tuning (TPT/tune_cls_rl.py with --tune_norm 0, the default of
// scripts/rlcf-tune.sh): weight gradients are tensor-core GEMMs over TRANSPOSED fp16 copies of the activation and
// output-gradient matrices (dW[out,in] = dY^T X per test image), so the one K-major tcgen05 GEMM kernel serves
// forward, dgrad and wgrad.  This file holds the layout transforms and the small reductions around those GEMMs.
#include "ptx.cuh"
#include "rlcf_internal.h"

namespace rlcf {

// out[c, g * rows_pad + r] = in[row(g, r), c]  (fp16), zero for rows_per_set <= r < rows_pad.
// skip_first > 0: the input rows are grouped in runs of `skip_first` tokens whose first token is dropped
// (class token: patch rows of a ViT), i.e. row(g, r) = g*rows_in_set + (r / (L-1)) * L + 1 + r % (L-1).
template <typename Tin>
__global__ void transpose_blocks_kernel(const Tin* __restrict__ in, int rows_per_set, int rows_pad, int cols,
                                        int skip_first, long long in_set_stride_rows, __half* __restrict__ out,
                                        long long ld_out) {
  __shared__ float tile[32][33];
  const int g = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows_per_set && c < cols) {
      long long src = r;
      if (skip_first > 0) src = static_cast<long long>(r / (skip_first - 1)) * skip_first + 1 + r % (skip_first - 1);
      v = static_cast<float>(in[(g * in_set_stride_rows + src) * cols + c]);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows_pad) out[static_cast<long long>(c) * ld_out + static_cast<long long>(g) * rows_pad + r] =
        __float2half_rn(tile[threadIdx.x][j]);
  }
}

// fp16 fast path: 64 x 64 tiles, 4-byte accesses on both sides, one block per (column tile, set) walking down the set's
// rows, which also yields the column sums of the set (the bias gradient of the Linear whose dY is being transposed).
__global__ void __launch_bounds__(256)
transpose_blocks_f16_kernel(const __half* __restrict__ in, int rows_per_set, int rows_pad, int cols, int skip_first,
                            long long in_set_stride_rows, __half* __restrict__ out, long long ld_out,
                            float* __restrict__ colsum, long long colsum_stride) {
  __shared__ __half tile[64][66];
  __shared__ float part[8][64];
  const int g = blockIdx.y, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = c0 + 2 * tx;
  float s0 = 0.f, s1 = 0.f;
  for (int r0 = 0; r0 < rows_pad; r0 += 64) {
    for (int j = ty; j < 64; j += 8) {
      const int r = r0 + j;
      __half2 val = __floats2half2_rn(0.f, 0.f);
      if (r < rows_per_set && c < cols) {
        long long src = r;
        if (skip_first > 0) src = static_cast<long long>(r / (skip_first - 1)) * skip_first + 1 + r % (skip_first - 1);
        val = *reinterpret_cast<const __half2*>(in + (g * in_set_stride_rows + src) * cols + c);
        const float2 f = __half22float2(val);
        s0 += f.x; s1 += f.y;
      }
      tile[j][2 * tx] = __low2half(val);
      tile[j][2 * tx + 1] = __high2half(val);
    }
    __syncthreads();
    for (int j = ty; j < 64; j += 8) {
      const int cc = c0 + j, r = r0 + 2 * tx;
      if (cc < cols && r < rows_pad)
        *reinterpret_cast<__half2*>(out + static_cast<long long>(cc) * ld_out + static_cast<long long>(g) * rows_pad + r) =
            __halves2half2(tile[2 * tx][j], tile[2 * tx + 1][j]);
    }
    __syncthreads();
  }
  if (colsum == nullptr) return;
  part[ty][2 * tx] = s0;
  part[ty][2 * tx + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x];
    colsum[g * colsum_stride + c0 + threadIdx.x] = s;
  }
}

int transpose_blocks_colsum(const __half* in, int n_sets, int rows_per_set, int rows_pad, int cols, int skip_first,
                            long long in_set_stride_rows, __half* out, long long ld_out, float* colsum,
                            long long colsum_stride, cudaStream_t stream) {
  if (n_sets <= 0 || n_sets > 65535 || rows_per_set <= 0 || rows_pad < rows_per_set || rows_pad % 2 || cols <= 0 ||
      cols % 2 || skip_first == 1 || ld_out % 2)
    return set_error(RLCF_ERR_ARG, "transpose_blocks_colsum: bad shape");
  dim3 grid((cols + 63) / 64, n_sets);
  transpose_blocks_f16_kernel<<<grid, 256, 0, stream>>>(in, rows_per_set, rows_pad, cols, skip_first,
                                                        in_set_stride_rows, out, ld_out, colsum, colsum_stride);
  RLCF_CHECK_LAUNCH("transpose_blocks_colsum");
  return 0;
}

int transpose_blocks(const void* in, int in_is_f32, int n_sets, int rows_per_set, int rows_pad, int cols,
                     int skip_first, long long in_set_stride_rows, __half* out, long long ld_out,
                     cudaStream_t stream) {
  if (n_sets <= 0 || rows_per_set <= 0 || rows_pad < rows_per_set || cols <= 0 || skip_first == 1)
    return set_error(RLCF_ERR_ARG, "transpose_blocks: bad shape");
  dim3 grid((cols + 31) / 32, (rows_pad + 31) / 32, n_sets), block(32, 8);
  if (in_is_f32)
    transpose_blocks_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(in), rows_per_set, rows_pad,
                                                               cols, skip_first, in_set_stride_rows, out, ld_out);
  else
    transpose_blocks_kernel<__half><<<grid, block, 0, stream>>>(static_cast<const __half*>(in), rows_per_set, rows_pad,
                                                                cols, skip_first, in_set_stride_rows, out, ld_out);
  RLCF_CHECK_LAUNCH("transpose_blocks");
  return 0;
}

// Bias gradient: out[g * out_stride + c] = scale * sum_r in[(g * rows_per_set + r) * cols + c]   (fp16 in, fp32 out)
__global__ void __launch_bounds__(256)
colsum_kernel(const __half* __restrict__ in, int rows_per_set, int cols, float* __restrict__ out, long long out_stride) {
  const int g = blockIdx.y;
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);   // 64 columns per block, 4 row phases
  const int phase = threadIdx.x >> 6;
  __shared__ float part[4][64];
  float acc = 0.f;
  if (c < cols)
    for (int r = phase; r < rows_per_set; r += 4)
      acc += __half2float(in[(static_cast<long long>(g) * rows_per_set + r) * cols + c]);
  part[phase][threadIdx.x & 63] = acc;
  __syncthreads();
  if (phase == 0 && c < cols)
    out[g * out_stride + c] = (part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x]);
}

int colsum_f16(const __half* in, int n_sets, int rows_per_set, int cols, float* out, long long out_stride,
               cudaStream_t stream) {
  if (n_sets <= 0 || rows_per_set <= 0 || cols <= 0) return set_error(RLCF_ERR_ARG, "colsum: bad shape");
  dim3 grid((cols + 63) / 64, n_sets);
  colsum_kernel<<<grid, 256, 0, stream>>>(in, rows_per_set, cols, out, out_stride);
  RLCF_CHECK_LAUNCH("colsum");
  return 0;
}

// Gradient of positional_embedding (and, in its row 0, of class_embedding):
// out[g * out_stride + t * d + c] = sum over the S sequences of set g of dx[(g*S + s)*L + t, c]   (model.py:227-228)
__global__ void seq_sum_kernel(const float* __restrict__ dx, int S, int L, int d4, long long total,
                               float* __restrict__ out, long long out_stride) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long per_set = static_cast<long long>(L) * d4;
    const long long g = i / per_set, rem = i % per_set;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
      const float4 v = reinterpret_cast<const float4*>(dx)[(g * S + s) * per_set + rem];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out + g * out_stride)[rem] = acc;
  }
}

int seq_sum(const float* dx, int n_sets, int S, int L, int d, float* out, long long out_stride, cudaStream_t stream) {
  if (n_sets <= 0 || S <= 0 || L <= 0 || d % 4 || out_stride % 4) return set_error(RLCF_ERR_ARG, "seq_sum: bad shape");
  const long long total = static_cast<long long>(n_sets) * L * (d / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  seq_sum_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dx, S, L, d / 4, total, out, out_stride);
  RLCF_CHECK_LAUNCH("seq_sum");
  return 0;
}

// Gradient of visual.proj: out[g * out_stride + i * E + j] = sum_s y[(g*S + s), i] * df[(g*S + s), j]   (model.py:237-238)
__global__ void outer_sum_kernel(const float* __restrict__ y, const float* __restrict__ df, int S, int d, int E,
                                 long long total, float* __restrict__ out, long long out_stride) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long per_set = static_cast<long long>(d) * E;
    const long long g = idx / per_set, rem = idx % per_set;
    const int i = static_cast<int>(rem / E), j = static_cast<int>(rem % E);
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = fmaf(y[(g * S + s) * d + i], df[(g * S + s) * E + j], acc);
    out[g * out_stride + rem] = acc;
  }
}

int outer_sum(const float* y, const float* df, int n_sets, int S, int d, int E, float* out, long long out_stride,
              cudaStream_t stream) {
  if (n_sets <= 0 || S <= 0 || d <= 0 || E <= 0) return set_error(RLCF_ERR_ARG, "outer_sum: bad shape");
  const long long total = static_cast<long long>(n_sets) * d * E;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  outer_sum_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(y, df, S, d, E, total, out, out_stride);
  RLCF_CHECK_LAUNCH("outer_sum");
  return 0;
}

}  // namespace rlcf
