// fp32 fused scaled-dot-product attention over the frame axis on CUDA cores
// (the <=1e-3 parity path; the bf16 throughput path is k_attn_tc.cu).
//
// Computes, per clip b:  O = softmax(Q K^T / sqrt(128) + keymask) V
// following vad/modeling/transformer.py:351-363 (scaled_dot_product), :319-325 (key padding
// mask: keys j >= lengths[b] get -inf), :333 (softmax over keys), :338 (P V) -- with the
// three [B,1,T,T] score tensors the reference materialises kept on chip (online softmax).
//
// One CTA = 64 query rows of one clip; K/V stream through shared memory in 64-key tiles.
// 256 threads: thread (ty,tx) owns S rows ty*4..+3 x cols tx+16j, and O rows ty*4..+3 x
// cols {tx*4..+3, 64+tx*4..+3}; row statistics are reduced across the 16 tx lanes.
#include <math_constants.h>

#include "vadb_common.cuh"

namespace vadb {
namespace {

constexpr int BQ = 64, BKV = 64, NT = 256;
constexpr int KS = D + 4;      // padded K row stride (floats): conflict-free float4 reads
constexpr int PS = BKV + 4;    // padded P row stride

struct Smem {
  float q[BQ][D];
  float k[BKV][KS];
  float v[BKV][D];
  float p[BQ][PS];
};

__global__ void __launch_bounds__(NT) attn_f32_kernel(const float* __restrict__ Q,
                                                      const float* __restrict__ Kg,
                                                      const float* __restrict__ Vg,
                                                      float* __restrict__ O,
                                                      const int32_t* __restrict__ lengths,
                                                      int T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int nq = (T + BQ - 1) / BQ;
  const int b = blockIdx.x / nq, q0 = (blockIdx.x % nq) * BQ;
  const long base = (long)b * T * D;
  int len = lengths ? lengths[b] : T;
  len = min(max(len, 0), T);
  const float scale = 1.0f / sqrtf((float)D);   // scores / np.sqrt(d_head), transformer.py:362

  // load Q tile (rows beyond T -> 0)
  for (int i = t; i < BQ * (D / 4); i += NT) {
    int r = i / (D / 4), c = i % (D / 4);
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) val = __ldg(reinterpret_cast<const float4*>(Q + base + (long)(q0 + r) * D) + c);
    *reinterpret_cast<float4*>(&sm.q[r][c * 4]) = val;
  }

  float m_run[4], l_run[4], acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -CUDART_INF_F;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }

  const int n_tiles = (len + BKV - 1) / BKV;   // tiles entirely beyond len contribute nothing
  for (int kt = 0; kt < n_tiles; ++kt) {
    const int k0 = kt * BKV;
    __syncthreads();   // previous tile fully consumed (also orders the Q stores on kt == 0)
    for (int i = t; i < BKV * (D / 4); i += NT) {
      int r = i / (D / 4), c = i % (D / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < T) {
        kv = __ldg(reinterpret_cast<const float4*>(Kg + base + (long)(k0 + r) * D) + c);
        vv = __ldg(reinterpret_cast<const float4*>(Vg + base + (long)(k0 + r) * D) + c);
      }
      *reinterpret_cast<float4*>(&sm.k[r][c * 4]) = kv;
      *reinterpret_cast<float4*>(&sm.v[r][c * 4]) = vv;
    }
    __syncthreads();

    // S = Q K^T (4 rows x 4 cols per thread)
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d0 = 0; d0 < D; d0 += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&sm.q[ty * 4 + i][d0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(&sm.k[tx + 16 * j][d0]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }

    // online softmax update
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = k0 + tx + 16 * j;
        s[i][j] = (key < len) ? s[i][j] * scale : -CUDART_INF_F;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[i], mx);
      const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float corr = expf(m_run[i] - m_use);          // exp(-inf) = 0 on the first tile
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(s[i][j] - m_use);
        sm.p[ty * 4 + i][tx + 16 * j] = p;
        ps += p;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l_run[i] = l_run[i] * corr + ps;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] *= corr;
    }
    __syncthreads();

    // O += P V
#pragma unroll 4
    for (int kk = 0; kk < BKV; kk += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(&sm.p[ty * 4 + i][kk]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 v0 = *reinterpret_cast<const float4*>(&sm.v[kk + u][tx * 4]);
        float4 v1 = *reinterpret_cast<const float4*>(&sm.v[kk + u][64 + tx * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = (u == 0) ? pv[i].x : (u == 1) ? pv[i].y : (u == 2) ? pv[i].z : pv[i].w;
          acc[i][0] = fmaf(p, v0.x, acc[i][0]); acc[i][1] = fmaf(p, v0.y, acc[i][1]);
          acc[i][2] = fmaf(p, v0.z, acc[i][2]); acc[i][3] = fmaf(p, v0.w, acc[i][3]);
          acc[i][4] = fmaf(p, v1.x, acc[i][4]); acc[i][5] = fmaf(p, v1.y, acc[i][5]);
          acc[i][6] = fmaf(p, v1.z, acc[i][6]); acc[i][7] = fmaf(p, v1.w, acc[i][7]);
        }
      }
    }
  }

  // normalise and store (l == 0 only when every key is masked: NaN, like the reference)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + ty * 4 + i;
    if (r >= T) continue;
    const float inv = 1.0f / l_run[i];
    float* o = O + base + (long)r * D;
    const float nanv = (l_run[i] == 0.f) ? CUDART_NAN_F : 0.f;
    *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[i][0] * inv + nanv, acc[i][1] * inv + nanv,
                                                         acc[i][2] * inv + nanv, acc[i][3] * inv + nanv);
    *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[i][4] * inv + nanv, acc[i][5] * inv + nanv,
                                                              acc[i][6] * inv + nanv, acc[i][7] * inv + nanv);
  }
}

}  // namespace

cudaError_t launch_attn_f32(const float* q, const float* k, const float* v, float* o,
                            const int32_t* lengths, int B, int T, cudaStream_t s) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem));
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((long)B * ((T + BQ - 1) / BQ)));
  attn_f32_kernel<<<grid, NT, sizeof(Smem), s>>>(q, k, v, o, lengths, T);
  return cudaGetLastError();
}

}  // namespace vadb
