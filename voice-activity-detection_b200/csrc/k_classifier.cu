// Final encoder LayerNorm + 2-class classifier + log-softmax / P(speech), one warp per frame.
//
// vad/modeling/transformer.py:33 (encoder.layer_norm), vad/models/self_attention.py:20-21,26-27
// (Linear(d_model, 2) + LogSoftmax(dim=2)) and the callers' softmax(...)[..., 1]
// (vad/predictor.py:225, :257-258), which equals sigmoid(z1 - z0).
#include "vadb_common.cuh"

namespace vadb {
namespace {

__global__ void __launch_bounds__(256) classifier_kernel(const float* __restrict__ h,
                                                         const float* __restrict__ g,
                                                         const float* __restrict__ b,
                                                         const float* __restrict__ wc,
                                                         const float* __restrict__ bc, int M,
                                                         float* __restrict__ prob,
                                                         float* __restrict__ logp) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  // per-lane constants: 4 consecutive features
  const float4 gam = __ldg(reinterpret_cast<const float4*>(g) + lane);
  const float4 bet = __ldg(reinterpret_cast<const float4*>(b) + lane);
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc) + lane);
  const float4 w1 = __ldg(reinterpret_cast<const float4*>(wc + D) + lane);
  const float b0 = __ldg(bc), b1 = __ldg(bc + 1);
  for (long row = warp; row < M; row += nwarps) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(h + row * D) + lane);
    float s = x.x + x.y + x.z + x.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    const float dx = x.x - mean, dy = x.y - mean, dz = x.z - mean, dw = x.w - mean;
    float q = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / D) + LN_EPS);
    const float y0 = dx * rstd * gam.x + bet.x, y1 = dy * rstd * gam.y + bet.y;
    const float y2 = dz * rstd * gam.z + bet.z, y3 = dw * rstd * gam.w + bet.w;
    float z0 = y0 * w0.x + y1 * w0.y + y2 * w0.z + y3 * w0.w;
    float z1 = y0 * w1.x + y1 * w1.y + y2 * w1.z + y3 * w1.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      z0 += __shfl_xor_sync(0xffffffffu, z0, o);
      z1 += __shfl_xor_sync(0xffffffffu, z1, o);
    }
    if (lane == 0) {
      z0 += b0; z1 += b1;
      // log_softmax([z0, z1]) computed stably
      const float mx = fmaxf(z0, z1);
      const float lse = mx + log1pf(expf(-fabsf(z1 - z0)));
      const float lp0 = z0 - lse, lp1 = z1 - lse;
      if (logp) { logp[row * 2] = lp0; logp[row * 2 + 1] = lp1; }
      if (prob) prob[row] = 1.0f / (1.0f + expf(z0 - z1));   // softmax(logp)[1]
    }
  }
}

}  // namespace

cudaError_t launch_classifier(const float* h, const float* g, const float* b, const float* wc,
                              const float* bc, int M, float* prob, float* logp, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  const int warps_per_block = 8;
  long blocks = ((long)M + warps_per_block - 1) / warps_per_block;
  if (blocks > 148 * 16) blocks = 148 * 16;
  classifier_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(h, g, b, wc, bc, M, prob, logp);
  return cudaGetLastError();
}

}  // namespace vadb
