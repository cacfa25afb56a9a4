// Evaluation-side kernels of the per-image loop: the top-1 / top-5 hit counters of TPT/utils/tools.py:84-98
// (`accuracy`) accumulated on the device, so that a whole adaptation step -- prediction included -- stays inside one
// CUDA graph and no host round trip separates two steps (TPT/tune_cls_rl.py:243-247 calls .item() per image).
#include "rlcf_internal.h"

namespace rlcf {

// One warp per test image: rank of the target class = number of classes scored strictly higher (ties broken towards
// the lower class index, the order a stable descending sort gives).  hits[0] += rank < 1, hits[1] += rank < 5,
// hits[2] += 1.  Integer counters: sums are exact and order-independent.
__global__ void accuracy_count_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int n,
                                      int C, unsigned long long* __restrict__ hits) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= n) return;
  const float* row = logits + static_cast<size_t>(warp) * C;
  const long long t = target[warp];
  int rank = 0;
  if (t >= 0 && t < C) {
    const float ref = row[t];
    for (int c = lane; c < C; c += 32) {
      const float x = row[c];
      rank += (x > ref) || (x == ref && c < t);
    }
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  } else {
    rank = C;   // a label outside the logits' columns can never be hit
  }
  if (lane == 0) {
    if (rank < 1) atomicAdd(hits + 0, 1ull);
    if (rank < 5) atomicAdd(hits + 1, 1ull);
    atomicAdd(hits + 2, 1ull);
  }
}

int accuracy_count(const float* logits, const long long* target, int n, int C, long long* hits, cudaStream_t stream) {
  if (n <= 0 || C <= 0) return set_error(RLCF_ERR_ARG, "accuracy_count: bad shape");
  const int threads = 128;
  const int blocks = (n * 32 + threads - 1) / threads;
  accuracy_count_kernel<<<blocks, threads, 0, stream>>>(logits, target, n, C,
                                                        reinterpret_cast<unsigned long long*>(hits));
  RLCF_CHECK_LAUNCH("accuracy_count");
  return 0;
}

}  // namespace rlcf
