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
                                                         float* __restrict__ logp, int tiled) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  // per-lane constants: 4 consecutive features
  const float4 gam = __ldg(reinterpret_cast<const float4*>(g) + lane);
  const float4 bet = __ldg(reinterpret_cast<const float4*>(b) + lane);
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc) + lane);
  const float4 w1 = __ldg(reinterpret_cast<const float4*>(wc + D) + lane);
  const float b0 = __ldg(bc), b1 = __ldg(bc + 1);
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();                 // weights above, activations below (vadb_common.cuh)
  // four rows per warp iteration: four independent 512-byte row loads in flight per warp
  for (long row0 = warp * 4; row0 < M; row0 += nwarps * 4) {
    float4 xr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      xr[u] = (row0 + u < M) ? __ldcg(reinterpret_cast<const float4*>(h) + h_quad_index(row0 + u, lane, tiled))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    float st[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) st[u] = (xr[u].x + xr[u].y) + (xr[u].z + xr[u].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) st[u] += __shfl_xor_sync(0xffffffffu, st[u], o);
    float d[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float mean = st[u] * (1.0f / D);
      d[u][0] = xr[u].x - mean; d[u][1] = xr[u].y - mean; d[u][2] = xr[u].z - mean; d[u][3] = xr[u].w - mean;
      st[u] = (d[u][0] * d[u][0] + d[u][1] * d[u][1]) + (d[u][2] * d[u][2] + d[u][3] * d[u][3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) st[u] += __shfl_xor_sync(0xffffffffu, st[u], o);
    float z0[4], z1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float rstd = 1.0f / sqrtf(st[u] * (1.0f / D) + LN_EPS);
      const float y0 = d[u][0] * rstd * gam.x + bet.x, y1 = d[u][1] * rstd * gam.y + bet.y;
      const float y2 = d[u][2] * rstd * gam.z + bet.z, y3 = d[u][3] * rstd * gam.w + bet.w;
      z0[u] = y0 * w0.x + y1 * w0.y + y2 * w0.z + y3 * w0.w;
      z1[u] = y0 * w1.x + y1 * w1.y + y2 * w1.z + y3 * w1.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        z0[u] += __shfl_xor_sync(0xffffffffu, z0[u], o);
        z1[u] += __shfl_xor_sync(0xffffffffu, z1[u], o);
      }
    if (lane < 4 && row0 + lane < M) {
      const long row = row0 + lane;
      const float a0 = (lane == 0 ? z0[0] : lane == 1 ? z0[1] : lane == 2 ? z0[2] : z0[3]) + b0;
      const float a1 = (lane == 0 ? z1[0] : lane == 1 ? z1[1] : lane == 2 ? z1[2] : z1[3]) + b1;
      // log_softmax([a0, a1]) computed stably
      const float mx = fmaxf(a0, a1);
      const float lse = mx + log1pf(expf(-fabsf(a1 - a0)));
      if (logp) { logp[row * 2] = a0 - lse; logp[row * 2 + 1] = a1 - lse; }
      if (prob) prob[row] = 1.0f / (1.0f + expf(a0 - a1));   // softmax(logp)[1]
    }
  }
}

}  // namespace

cudaError_t launch_classifier(const float* h, const float* g, const float* b, const float* wc,
                              const float* bc, int M, float* prob, float* logp, int tiled, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  const int warps_per_block = 8;
  long blocks = ((long)M + 4 * warps_per_block - 1) / (4 * warps_per_block);
  if (blocks > 148 * 16) blocks = 148 * 16;
  { cudaError_t e = launch_k(classifier_kernel, (unsigned)blocks, warps_per_block * 32, 0, s, h, g, b, wc, bc, M, prob, logp, tiled); if (e != cudaSuccess) return e; }
  return cudaGetLastError();
}

}  // namespace vadb
