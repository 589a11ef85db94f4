// fp32 CUDA-core GEMM with fused prologue (LayerNorm, window gather) and epilogue
// (bias, ReLU, positional encoding, residual, output split / bf16 cast).
//
// Serves every per-frame Linear of the reference on the fp32 parity path
// (vad/models/self_attention.py:13 input Linear; vad/modeling/transformer.py:281-284 Q/K/V,
// :347 final projection, :370-375 feed-forward) and the front-end of both paths.
//
//   C[m, n] = epi( sum_k pro(A)[m, k] * W[n, k] + bias[n] )
//
// Tiling: 128x128 output tile per CTA, BK = 16, 256 threads, 8x8 register micro-tile split
// in 4+4 along both axes so that every shared-memory read is a conflict-free float4.
// Global loads of the next K chunk are issued before the FMAs of the current one.
#include "vadb_common.cuh"

namespace vadb {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

__device__ __forceinline__ int window_src_row(int m, int W, int half, int jump) {
  // vad/predictor.py:186-199: rel = [-half..0) step jump, 0, [1..half] step jump
  int i = m / W, k = m - i * W;
  int nl = (half + jump - 1) / jump;
  int rel = (k < nl) ? (-half + k * jump) : (k == nl ? 0 : 1 + (k - nl - 1) * jump);
  return half + i + rel;
}

__device__ __forceinline__ void load8(const float* p, int k, int K, bool row_ok, bool vec,
                                      float (&v)[8]) {
  if (row_ok && vec && k + 8 <= K) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p + k));
    float4 b = __ldg(reinterpret_cast<const float4*>(p + k + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (row_ok && k + i < K) ? __ldg(p + k + i) : 0.f;
  }
}

__device__ __forceinline__ void load8(const bf16* p, int k, int K, bool row_ok, bool vec,
                                      float (&v)[8]) {
  if (row_ok && vec && k + 8 <= K) {
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(p + k));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h2[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (row_ok && k + i < K) ? __bfloat162float(p[k + i]) : 0.f;
  }
}

template <typename AT, bool LN>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(GemmArgs g, int a_vec, int w_vec) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Ws[BK][BN];
  __shared__ float s_mean[BM], s_rstd[BM];

  const int t = threadIdx.x;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int M = g.M, N = g.N, K = g.K;
  const AT* A = reinterpret_cast<const AT*>(g.A);

  // loader coordinates
  const int lr = t & (BM - 1);
  const int kofs = (t >> 7) * 8;
  const int gm = m0 + lr;
  const bool a_ok = gm < M;
  long a_row = gm;
  if (g.win_W > 0 && a_ok) a_row = window_src_row(gm, g.win_W, g.win_half, g.win_jump);
  const AT* a_ptr = A + (a_ok ? a_row * (long)K : 0);
  const int gn = n0 + lr;
  const bool w_ok = gn < N;
  const float* w_ptr = g.W + (w_ok ? (long)gn * K : 0);

  if (LN) {
    // per-row LayerNorm statistics (biased variance, eps = 1e-5), K == 128: one float4 / lane
    const int warp = t >> 5, lane = t & 31;
    for (int r = warp * (BM / 8); r < (warp + 1) * (BM / 8); ++r) {
      int row = m0 + r;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < M) x = __ldg(reinterpret_cast<const float4*>(
                        reinterpret_cast<const float*>(g.A) + (long)row * K) + lane);
      float s = x.x + x.y + x.z + x.w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      float mean = s * (1.0f / 128.0f);
      float dx = x.x - mean, dy = x.y - mean, dz = x.z - mean, dw = x.w - mean;
      float q = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      if (lane == 0) {
        s_mean[r] = mean;
        s_rstd[r] = 1.0f / sqrtf(q * (1.0f / 128.0f) + LN_EPS);
      }
    }
    __syncthreads();
  }
  float my_mean = 0.f, my_rstd = 1.f;
  if (LN) { my_mean = s_mean[lr]; my_rstd = s_rstd[lr]; }

  float ra[8], rw[8];
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int ty = t >> 4, tx = t & 15;

  load8(a_ptr, kofs, K, a_ok, a_vec != 0, ra);
  load8(w_ptr, kofs, K, w_ok, w_vec != 0, rw);

  for (int k0 = 0; k0 < K; k0 += BK) {
    if (LN) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int k = k0 + kofs + i;
        float gam = (k < K) ? __ldg(g.ln_g + k) : 0.f;
        float bet = (k < K) ? __ldg(g.ln_b + k) : 0.f;
        ra[i] = a_ok ? ((ra[i] - my_mean) * my_rstd * gam + bet) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[kofs + i][lr] = ra[i];
      Ws[kofs + i][lr] = rw[i];
    }
    __syncthreads();
    if (k0 + BK < K) {
      load8(a_ptr, k0 + BK + kofs, K, a_ok, a_vec != 0, ra);
      load8(w_ptr, k0 + BK + kofs, K, w_ok, w_vec != 0, rw);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Ws[kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gr = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gr >= M) continue;
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      const int c0 = n0 + jb * 64 + tx * 4;
      if (c0 >= N) continue;
      float4 bi = __ldg(reinterpret_cast<const float4*>(g.bias + c0));
      float v0 = acc[i][jb * 4 + 0] + bi.x, v1 = acc[i][jb * 4 + 1] + bi.y;
      float v2 = acc[i][jb * 4 + 2] + bi.z, v3 = acc[i][jb * 4 + 3] + bi.w;
      if (g.relu) {
        v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
      }
      if (g.pe) {
        float4 p = __ldg(reinterpret_cast<const float4*>(g.pe + (long)(gr % g.pe_T) * N + c0));
        v0 += p.x; v1 += p.y; v2 += p.z; v3 += p.w;
      }
      if (g.residual) {
        float4 r = *reinterpret_cast<const float4*>(g.residual + (long)gr * N + c0);
        v0 += r.x; v1 += r.y; v2 += r.z; v3 += r.w;
      }
      const int ob = c0 / g.out_split, oc = c0 - ob * g.out_split;
      if (g.out_is_bf16) {
        bf16* o = reinterpret_cast<bf16*>(g.out[ob]) + (long)gr * g.out_split + oc;
        __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(o) = pk;
      } else {
        float* o = reinterpret_cast<float*>(g.out[ob]) + (long)gr * g.out_split + oc;
        *reinterpret_cast<float4*>(o) = make_float4(v0, v1, v2, v3);
      }
    }
  }
}

}  // namespace

cudaError_t launch_gemm_f32(const GemmArgs& a, cudaStream_t s) {
  if (a.M <= 0) return cudaSuccess;
  if (a.N % BN != 0 || a.out_split % BN != 0) return cudaErrorInvalidValue;
  if (a.ln_g && (a.K != 128 || a.a_is_bf16)) return cudaErrorInvalidValue;
  if (a.pe && a.N != 128) return cudaErrorInvalidValue;
  dim3 grid(a.N / BN, (a.M + BM - 1) / BM);
  const int esz = a.a_is_bf16 ? 2 : 4;
  int a_vec = (a.K % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.A) % 16) == 0) &&
              ((a.K * esz) % 16 == 0);
  int w_vec = (a.K % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.W) % 16) == 0);
  if (a.ln_g) {
    gemm_f32_kernel<float, true><<<grid, NT, 0, s>>>(a, a_vec, w_vec);
  } else if (a.a_is_bf16) {
    gemm_f32_kernel<bf16, false><<<grid, NT, 0, s>>>(a, a_vec, w_vec);
  } else {
    gemm_f32_kernel<float, false><<<grid, NT, 0, s>>>(a, a_vec, w_vec);
  }
  return cudaGetLastError();
}

}  // namespace vadb
