// Boosted aggregation of the bDNN-style multi-window prediction, plus small element-wise helpers.
//
// Reference (vad/predictor.py:238-258, :95): the model's log-probs for window i, slot k are
// scattered to boosted_outputs[positions[i,k], k] of a zero-initialised [L, W, 2] array, the
// class axis is soft-maxed and class 1 is taken, giving probs[L, W]; callers then average over
// W.  Slots never written keep (0, 0) -> probability exactly 0.5 (boosted_counts is unused).
//
// Done here in gather form (no atomics, deterministic): slot k of frame p was written by window
// i = p - half - rel[k] iff 0 <= i < n.  Since softmax(log_softmax(z))[1] == sigmoid(z1 - z0),
// the per-window probability computed by the classifier kernel is what lands in the slot.
#include "vadb_common.cuh"

namespace vadb {
namespace {

__device__ __forceinline__ int rel_of_slot(int k, int half, int jump) {
  int nl = (half + jump - 1) / jump;
  return (k < nl) ? (-half + k * jump) : (k == nl ? 0 : 1 + (k - nl - 1) * jump);
}

__global__ void boost_kernel(const float* __restrict__ prob_nW, int L, int half, int jump, int W,
                             float* __restrict__ probs_LW, float* __restrict__ mean_L) {
  const int n = L - 2 * half;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < L;
       p += (long)gridDim.x * blockDim.x) {
    float sum = 0.f;
    for (int k = 0; k < W; ++k) {
      const long i = p - half - rel_of_slot(k, half, jump);
      const float v = (i >= 0 && i < n) ? __ldg(prob_nW + i * W + k) : 0.5f;
      if (probs_LW) probs_LW[p * W + k] = v;
      sum += v;
    }
    if (mean_L) mean_L[p] = sum / (float)W;
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(in + i));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + i) = pk;
  }
  // tail (n % 4) handled by the first thread of the grid
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t j = n & ~(size_t)3; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
}

// [rows, cols] fp32 -> [rows, 128] bf16, zero-padded columns (front-end weight for the tensor-core path)
__global__ void pad_rows_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * 128) {
    const int r = i >> 7, c = i & 127;
    out[i] = __float2bfloat16_rn(c < cols ? in[(long)r * cols + c] : 0.f);
  }
}

}  // namespace

cudaError_t launch_pad_rows_bf16(const float* in, bf16* out, int rows, int cols, cudaStream_t s) {
  pad_rows_bf16_kernel<<<(rows * 128 + 255) / 256, 256, 0, s>>>(in, out, rows, cols);
  return cudaGetLastError();
}

cudaError_t launch_boost(const float* prob_nW, int L, int half, int jump, int W, float* probs_LW,
                         float* mean_L, cudaStream_t s) {
  if (L <= 0) return cudaSuccess;
  int blocks = (L + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  boost_kernel<<<blocks, 256, 0, s>>>(prob_nW, L, half, jump, W, probs_LW, mean_L);
  return cudaGetLastError();
}

cudaError_t launch_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks == 0) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace vadb
