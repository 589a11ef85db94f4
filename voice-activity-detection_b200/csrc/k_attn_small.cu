// Attention for very short sequences (T <= 8 frames): the regime of the reference Predictor, whose model
// only ever sees W = 2*(half-1)//jump + 3 = 7-frame context windows (vad/predictor.py:57-59, :182-224;
// SURVEY.md section 0 "R1").
//
//   O[w] = softmax(Q[w] K[w]^T / sqrt(128) + keymask(lengths[w])) V[w]      bf16 in, bf16 out, fp32 math
//
// Reference semantics: vad/modeling/transformer.py:351-363, :319-346 (same as k_attn_tc.cu).
//
// A 7x7 score matrix wastes 94 % of a 128-row tcgen05 tile (and the big kernel spends a whole
// TMA/TMEM/mbarrier pipeline round trip per window), so this path is HBM-bound work on the warp-level
// tensor-core instruction instead: ONE WARP PER PAIR OF WINDOWS, no shared memory, no block-level
// synchronisation, every byte read once with 16-byte loads.
//
//   S_a, S_b  two m16n8k16 chains over d = 128: the 16 A rows are [window a | window b] (8 rows each,
//             rows >= T are clamped duplicates), B is the 8 keys of window a resp. b; the diagonal
//             blocks of the two results are the two 8x8 score matrices
//   softmax   each score row lives in the 4 lanes of a quad (2 keys per lane): two xor-shuffles for
//             the max and the sum; masked / padded keys are -inf
//   O         one m16n8k16 per 8 output dims with a block-diagonal A = diag(P_a, P_b) (K = 16 = the 8 keys
//             of a then the 8 keys of b), so both windows ride the same instruction and the
//             accumulator fragment IS the P fragment (no shuffles)
// The contraction index of Q K^T and the output-dim index of P V are permuted so that every lane
// reads/writes contiguous 64-byte (Q, K, O) and 32-byte (V) pieces of a row.
//
// Algorithmic traffic: 4 * n_windows * T * 128 * 2 bytes (Q, K, V read once, O written once).
#include <math_constants.h>

#include "vadb_common.cuh"

namespace vadb {
namespace {

constexpr int SMALL_T_MAX = 8;
constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Lane (g = lane / 4, t4 = lane % 4) owns, of a row, the 32 dims [32 t4, 32 t4 + 32): 16 bf16 pairs.
// Q K^T k-step s (0..7) contracts over pairs 2s (fragment columns k = 2 t4, 2 t4 + 1) and 2s + 1
// (k = 2 t4 + 8, 2 t4 + 9) of every lane -- A and B use the same permutation of d, so the sum is the
// plain dot product.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
attn_small_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V,
                  bf16* __restrict__ O, const int32_t* __restrict__ lengths, int n_win, int T) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int warp_global = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * WARPS_PER_BLOCK;
  const int n_pairs = (n_win + 1) >> 1;
  const float c = 1.4426950408889634f * 0.08838834764831845f;   // log2(e) / sqrt(d_head)  (transformer.py:362)
  const int gq = min(g, T - 1);                                  // clamped row of this lane's quad
  const int k0 = min(2 * t4, T - 1), k1 = min(2 * t4 + 1, T - 1); // clamped key rows of the P V B fragment

  for (int pair = warp_global; pair < n_pairs; pair += n_warps) {
    const int wa = 2 * pair;
    const bool has_b = wa + 1 < n_win;
    const int wb = has_b ? wa + 1 : wa;
    const size_t ra = (size_t)wa * T, rb = (size_t)wb * T;
    int len_a = T, len_b = T;
    if (lengths) {
      len_a = min(max(lengths[wa], 0), T);
      len_b = min(max(lengths[wb], 0), T);
    }
    // ---- loads: 64 B of the Q and K rows of this quad, 32 B of four V rows (all independent, issued up front)
    uint4 qa[4], qb[4], ka[4], kb[4], va0[2], va1[2], vb0[2], vb1[2];
    {
      const bf16* pq_a = Q + (ra + gq) * D + t4 * 32;
      const bf16* pq_b = Q + (rb + gq) * D + t4 * 32;
      const bf16* pk_a = K + (ra + gq) * D + t4 * 32;
      const bf16* pk_b = K + (rb + gq) * D + t4 * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qa[i] = ldg_nc16(pq_a + i * 8);
        qb[i] = ldg_nc16(pq_b + i * 8);
        ka[i] = ldg_nc16(pk_a + i * 8);
        kb[i] = ldg_nc16(pk_b + i * 8);
      }
      // P V: lane needs, for output dims [16 g, 16 g + 16), the V rows of keys 2 t4 and 2 t4 + 1
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        va0[i] = ldg_nc16(V + (ra + k0) * D + g * 16 + i * 8);
        va1[i] = ldg_nc16(V + (ra + k1) * D + g * 16 + i * 8);
        vb0[i] = ldg_nc16(V + (rb + k0) * D + g * 16 + i * 8);
        vb1[i] = ldg_nc16(V + (rb + k1) * D + g * 16 + i * 8);
      }
    }
    // ---- scores: rows 0-7 of the A tile = window a, rows 8-15 = window b
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const uint32_t* qaw = reinterpret_cast<const uint32_t*>(qa);
      const uint32_t* qbw = reinterpret_cast<const uint32_t*>(qb);
      const uint32_t* kaw = reinterpret_cast<const uint32_t*>(ka);
      const uint32_t* kbw = reinterpret_cast<const uint32_t*>(kb);
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        mma_bf16_16816(sa, qaw[2 * s], qbw[2 * s], qaw[2 * s + 1], qbw[2 * s + 1], kaw[2 * s], kaw[2 * s + 1]);
        mma_bf16_16816(sb, qaw[2 * s], qbw[2 * s], qaw[2 * s + 1], qbw[2 * s + 1], kbw[2 * s], kbw[2 * s + 1]);
      }
    }
    // sa[0], sa[1]: query g of window a x keys 2 t4, 2 t4 + 1 of window a;  sb[2], sb[3]: same for window b
    // (sa[2..3] / sb[0..1] are the off-diagonal blocks, unused)
    float pa0, pa1, pb0, pb1;
    {
      const float x0 = (2 * t4 < len_a) ? sa[0] * c : -CUDART_INF_F;
      const float x1 = (2 * t4 + 1 < len_a) ? sa[1] * c : -CUDART_INF_F;
      const float y0 = (2 * t4 < len_b) ? sb[2] * c : -CUDART_INF_F;
      const float y1 = (2 * t4 + 1 < len_b) ? sb[3] * c : -CUDART_INF_F;
      const float ma = quad_max(fmaxf(x0, x1)), mb = quad_max(fmaxf(y0, y1));
      // every key masked (length 0): -inf - -inf = NaN rows, as the reference's softmax produces
      pa0 = exp2f(x0 - ma); pa1 = exp2f(x1 - ma);
      pb0 = exp2f(y0 - mb); pb1 = exp2f(y1 - mb);
      const float ia = 1.0f / quad_sum(pa0 + pa1), ib = 1.0f / quad_sum(pb0 + pb1);
      pa0 *= ia; pa1 *= ia; pb0 *= ib; pb1 *= ib;
    }
    // ---- O = diag(P_a, P_b) [V_a ; V_b]: A fragment a0 = P_a (rows g, keys 2 t4..), a3 = P_b (rows g + 8,
    // k = 8 + 2 t4..), a1 = a2 = 0; B fragment of MMA m (output dim 16 n + m for fragment column n):
    // b0 = {V_a[2 t4][16 g + m], V_a[2 t4 + 1][16 g + m]}, b1 = same of window b
    const uint32_t pfa = pack2(pa0, pa1), pfb = pack2(pb0, pb1);
    float oacc[16][4];
    {
      const uint32_t* a0w = reinterpret_cast<const uint32_t*>(va0);
      const uint32_t* a1w = reinterpret_cast<const uint32_t*>(va1);
      const uint32_t* b0w = reinterpret_cast<const uint32_t*>(vb0);
      const uint32_t* b1w = reinterpret_cast<const uint32_t*>(vb1);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const uint32_t sel = (m & 1) ? 0x7632u : 0x5410u;        // high / low halves of the two rows
        const uint32_t b0 = __byte_perm(a0w[m >> 1], a1w[m >> 1], sel);
        const uint32_t b1 = __byte_perm(b0w[m >> 1], b1w[m >> 1], sel);
        oacc[m][0] = oacc[m][1] = oacc[m][2] = oacc[m][3] = 0.f;
        mma_bf16_16816(oacc[m], pfa, 0u, 0u, pfb, b0, b1);
      }
    }
    // oacc[m][0], [1]: window a, query g, dims 16 (2 t4) + m and 16 (2 t4 + 1) + m;  [2], [3]: window b.
    // Per lane that is dims [32 t4, 32 t4 + 32) of its row: 64 contiguous bytes.
    if (g < T) {
      uint4 oa[4], ob[4];
      uint32_t* oaw = reinterpret_cast<uint32_t*>(oa);
      uint32_t* obw = reinterpret_cast<uint32_t*>(ob);
#pragma unroll
      for (int m = 0; m < 16; m += 2) {
        oaw[m >> 1] = pack2(oacc[m][0], oacc[m + 1][0]);
        oaw[8 + (m >> 1)] = pack2(oacc[m][1], oacc[m + 1][1]);
        obw[m >> 1] = pack2(oacc[m][2], oacc[m + 1][2]);
        obw[8 + (m >> 1)] = pack2(oacc[m][3], oacc[m + 1][3]);
      }
      uint4* po_a = reinterpret_cast<uint4*>(O + (ra + g) * D + t4 * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) po_a[i] = oa[i];
      if (has_b) {
        uint4* po_b = reinterpret_cast<uint4*>(O + (rb + g) * D + t4 * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) po_b[i] = ob[i];
      }
    }
  }
}

}  // namespace

bool attn_small_supported(int T) { return T >= 1 && T <= SMALL_T_MAX; }

cudaError_t launch_attn_small(const bf16* q, const bf16* k, const bf16* v, bf16* o, const int32_t* lengths,
                              int B, int T, int num_sms, cudaStream_t s) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  if (!attn_small_supported(T)) return cudaErrorInvalidValue;
  const int n_pairs = (B + 1) / 2;
  const long want = ((long)n_pairs + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  const long cap = (long)num_sms * 16;            // 64 warps per SM; each warp then loops over its pairs
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  return launch_k(attn_small_kernel, grid, WARPS_PER_BLOCK * 32, 0, s, q, k, v, o, lengths, B, T);
}

}  // namespace vadb
