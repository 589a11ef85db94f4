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
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < L;
       p += (long)gridDim.x * blockDim.x) {
    float sum = 0.f;
    for (int k = 0; k < W; ++k) {
      const long i = p - half - rel_of_slot(k, half, jump);
      const float v = (i >= 0 && i < n) ? __ldcg(prob_nW + i * W + k) : 0.5f;
      if (probs_LW) probs_LW[p * W + k] = v;
      sum += v;
    }
    if (mean_L) mean_L[p] = sum / (float)W;
  }
}

// Window gather AFTER the input projection.  The input layer is a per-frame Linear
// (vad/models/self_attention.py:12-16), so the [L, F] clip is projected once and the W-fold overlapping
// context windows (vad/predictor.py:182-218) are assembled from the projected rows: row m = i*W + k of
// the model input is  proj[half + i + rel[k]] + PE[k]/sqrt(d)  (the window is the model's whole sequence,
// so its positional slot is k).  One warp per output row: 512 B gathered (L2-resident: the projected
// clip is W-fold reused), residual row written in fp32 and LayerNorm_1 of layer 0 emitted in bf16 for
// the first Q/K/V GEMM -- the same two outputs the front-end GEMM produces on the clip path.
__global__ void __launch_bounds__(256)
window_gather_ln_kernel(const float* __restrict__ proj, const float* __restrict__ pe, float* __restrict__ out_h,
                        bf16* __restrict__ out_ln, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                        long n_rows, int W, int half, int jump, int tiled) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long n_warps = ((long)gridDim.x * blockDim.x) >> 5;
  const float4 gam = __ldg(reinterpret_cast<const float4*>(ln_g) + lane);
  const float4 bet = __ldg(reinterpret_cast<const float4*>(ln_b) + lane);
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (long m = warp; m < n_rows; m += n_warps) {
    const long i = m / W;
    const int k = (int)(m - i * W);
    const long src = half + i + rel_of_slot(k, half, jump);
    float4 v = *(reinterpret_cast<const float4*>(proj + src * D) + lane);     // written by the previous kernel
    const float4 e = __ldg(reinterpret_cast<const float4*>(pe + (long)k * D) + lane);
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    reinterpret_cast<float4*>(out_h)[h_quad_index(m, lane, tiled)] = v;
    float sum = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / D);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    float sq = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / D) + LN_EPS);       // biased variance, eps 1e-5 (nn.LayerNorm)
    __nv_bfloat162 lo = __floats2bfloat162_rn(dx * rstd * gam.x + bet.x, dy * rstd * gam.y + bet.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(dz * rstd * gam.z + bet.z, dw * rstd * gam.w + bet.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(out_ln + m * D) + lane, pk);
  }
}

// Row-major fp32 residual rows -> tiled layout + LayerNorm(row) in bf16: used when the front end ran on the
// CUDA-core fallback (feature sizes the tensor-core front end does not take) in bf16 mode.  One warp per row.
__global__ void __launch_bounds__(256)
retile_ln_kernel(const float* __restrict__ h_rows, float* __restrict__ h_tiled, bf16* __restrict__ out_ln,
                 const float* __restrict__ ln_g, const float* __restrict__ ln_b, long n_rows) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long n_warps = ((long)gridDim.x * blockDim.x) >> 5;
  const float4 gam = __ldg(reinterpret_cast<const float4*>(ln_g) + lane);
  const float4 bet = __ldg(reinterpret_cast<const float4*>(ln_b) + lane);
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (long m = warp; m < n_rows; m += n_warps) {
    const float4 v = *(reinterpret_cast<const float4*>(h_rows + m * D) + lane);
    reinterpret_cast<float4*>(h_tiled)[h_quad_index(m, lane, 1)] = v;
    float sum = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / D);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    float sq = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / D) + LN_EPS);
    __nv_bfloat162 lo = __floats2bfloat162_rn(dx * rstd * gam.x + bet.x, dy * rstd * gam.y + bet.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(dz * rstd * gam.z + bet.z, dw * rstd * gam.w + bet.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *(reinterpret_cast<uint2*>(out_ln + m * D) + lane) = pk;
  }
}

// Length-bucketed forward: gather the first Tk frames of the listed clips into a dense batch (16-byte vectors;
// row_bytes is a multiple of 8 and the buffers are 16-byte aligned, so a clip's Tk * row_bytes is a multiple of 16)
__global__ void gather_clips_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, const int32_t* __restrict__ ids,
                                    long clip_vec_in, long clip_vec_out, long total) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long c = i / clip_vec_out, r = i - c * clip_vec_out;
    out[i] = __ldg(x + (long)ids[c] * clip_vec_in + r);
  }
}

__global__ void scatter_clips_kernel(const float* __restrict__ prob_k, const float* __restrict__ logp_k,
                                     float* __restrict__ prob, float* __restrict__ logp, const int32_t* __restrict__ ids,
                                     int T, int Tk, long total) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long c = i / Tk, t = i - c * Tk;
    const long dst = (long)ids[c] * T + t;
    if (prob) prob[dst] = prob_k[i];
    if (logp) { logp[dst * 2] = logp_k[i * 2]; logp[dst * 2 + 1] = logp_k[i * 2 + 1]; }
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

cudaError_t launch_window_gather_ln(const float* proj, const float* pe, float* out_h, bf16* out_ln,
                                    const float* ln_g, const float* ln_b, long n_rows, int W, int half,
                                    int jump, int tiled, cudaStream_t s) {
  if (n_rows <= 0) return cudaSuccess;
  long blocks = (n_rows + 7) / 8;                 // 8 warps (rows) per block per pass
  if (blocks > 148 * 8) blocks = 148 * 8;         // a multiple of the SM count; warps then stride over rows
  return launch_k(window_gather_ln_kernel, (unsigned)blocks, 256, 0, s, proj, pe, out_h, out_ln, ln_g, ln_b, n_rows, W, half, jump, tiled);
}

cudaError_t launch_retile_ln(const float* h_rows, float* h_tiled, bf16* out_ln, const float* ln_g,
                             const float* ln_b, long n_rows, cudaStream_t s) {
  if (n_rows <= 0) return cudaSuccess;
  long blocks = (n_rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return launch_k(retile_ln_kernel, (unsigned)blocks, 256, 0, s, h_rows, h_tiled, out_ln, ln_g, ln_b, n_rows);
}

cudaError_t launch_boost(const float* prob_nW, int L, int half, int jump, int W, float* probs_LW,
                         float* mean_L, cudaStream_t s) {
  if (L <= 0) return cudaSuccess;
  int blocks = (L + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return launch_k(boost_kernel, blocks, 256, 0, s, prob_nW, L, half, jump, W, probs_LW, mean_L);
}

cudaError_t launch_gather_clips(const void* x, void* out, const int32_t* ids, int n, int T, int Tk, int row_bytes,
                                cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (((long)Tk * row_bytes) % 16 || ((long)T * row_bytes) % 16) return cudaErrorInvalidValue;
  const long vin = (long)T * row_bytes / 16, vout = (long)Tk * row_bytes / 16, total = vout * n;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  return launch_k(gather_clips_kernel, (unsigned)blocks, 256, 0, s, (const uint4*)x, (uint4*)out, ids, vin, vout, total);
}

cudaError_t launch_scatter_clips(const float* prob_k, const float* logp_k, float* prob, float* logp, const int32_t* ids,
                                 int n, int T, int Tk, cudaStream_t s) {
  if (n <= 0 || (!prob && !logp)) return cudaSuccess;
  const long total = (long)n * Tk;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  return launch_k(scatter_clips_kernel, (unsigned)blocks, 256, 0, s, prob_k, logp_k, prob, logp, ids, T, Tk, total);
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
