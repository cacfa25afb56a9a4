// Internal declarations shared by the .cu translation units of librlcf_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>

#include "../../include/rlcf_b200.h"

namespace rlcf {

// Error codes (RLCF_OK, RLCF_ERR_*) and epilogue ids (RLCF_EPI_*) come from the public header.
enum : int {
  EPI_F16 = RLCF_EPI_F16,
  EPI_GELU_F16 = RLCF_EPI_GELU_F16,
  EPI_RESID_F32 = RLCF_EPI_RESID_F32,
  EPI_GELU_BWD_F16 = RLCF_EPI_GELU_BWD_F16,
  EPI_F32 = RLCF_EPI_F32,
  EPI_ADAMW = RLCF_EPI_ADAMW,
  EPI_COUNT = 6,
};

// Optimizer state of the fused wgrad + AdamW epilogue (EPI_ADAMW); all pointers address the weight's [n_out, n_in] tile
// of group 0, group g is `*_gs` elements further.
struct AdamwEpi {
  float* m = nullptr;
  float* v = nullptr;
  const float* p_in = nullptr;   // parameters are read here (written to GemmArgs::out)
  long long p_in_gs = 0;
  __half* w16 = nullptr;
  long long w16_gs = 0;
  int fresh = 0;                 // 1: moments start from zero (first step after reset)
  float lr = 0, b1 = 0, b2 = 0, eps = 0, wd = 0, bc1 = 1, bc2_sqrt = 1;
};

int set_error(int code, const char* fmt, ...);
void count_launch();
int sm_count();
int gemm_cta_group();
int gemm_multicast();  // 1 = 4-CTA clusters with the B tile TMA-multicast across two CTA pairs
int attention_impl();  // 0 = tcgen05 forward kernel, 1 = warp-level mma.sync kernel

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

int gemm_f16(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, int epi, float alpha,
             const float* bias, const float* resid, const __half* aux_in, __half* aux_out, void* out, int ldo,
             cudaStream_t stream);
// G independent problems of the same shape in one launch: group g reads A + g*a_gs, B + g*b_gs, bias + g*bias_gs and
// writes out/resid/aux + g*out_gs (strides in elements).
int gemm_f16_grouped(const __half* A, int lda, long long a_gs, const __half* B, int ldb, long long b_gs, int G, int M,
                     int N, int K, int epi, float alpha, const float* bias, long long bias_gs, const float* resid,
                     const __half* aux_in, __half* aux_out, void* out, int ldo, long long out_gs,
                     cudaStream_t stream, const AdamwEpi* adamw = nullptr);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: the bookkeeping of "already raised
// to at least `smem` bytes" is kept per device, so a process that drives a second GPU configures the kernel there too.
struct DynSmemState {
  size_t configured[32] = {};
};
template <typename Kernel>
inline cudaError_t ensure_dyn_smem(Kernel kernel, size_t smem, DynSmemState& st) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 32) dev = 31;   // beyond the table: always (re)configure
  if (smem > st.configured[dev] || dev == 31) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    st.configured[dev] = smem;
  }
  return cudaSuccess;
}

// Checks the launch of the kernel that was just enqueued.
#define RLCF_CHECK_LAUNCH(name)                                                                 \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) return ::rlcf::set_error(RLCF_ERR_CUDA, name ": %s", cudaGetErrorString(e__)); \
    ::rlcf::count_launch();                                                                     \
  } while (0)

}  // namespace rlcf
