// Fused scaled-dot-product attention over the frame axis on the 5th-gen tensor cores (sm_100a).
//
//   O[b] = softmax(Q[b] K[b]^T / sqrt(128) + keymask(lengths[b])) V[b]        bf16 in, bf16 out
//
// Reference semantics: vad/modeling/transformer.py:351-363 (scaled_dot_product), :319-325 (key
// padding mask), :333 (softmax over keys), :338-346 (P V and the head merge) -- the three
// [B,1,T,T] fp32 score tensors the reference materialises never leave the SM.
//
// Work decomposition (one CTA = one clip x 256 query rows = two 128-row query tiles):
//   warp 8      TMA producer: Q (once) and a 3-stage ring of 64-key K/V tiles, 128B-swizzled
//   warp 9      MMA issuer (one elected thread): S_t = Q_t K^T and O_t += P_t V via tcgen05.mma,
//               accumulators in TMEM (S: 2 x 64 columns, O: 2 x 128 columns)
//   warps 0-3   softmax warpgroup of query tile 0: one thread per query row (TMEM lane),
//   warps 4-7   softmax warpgroup of query tile 1   tcgen05.ld S -> online softmax (fp32, exp2,
//               lazy rescale of O in TMEM) -> P (bf16) into swizzled shared memory
// The two query tiles share every K/V tile and ping-pong on the tensor pipe: while the softmax
// of one tile runs on the CUDA cores, the MMAs of the other tile run on the tensor cores.
//
// Algorithmic traffic per launch: read Q,K,V once + write O once = 4*B*T*128*2 bytes (SURVEY.md
// section 8d); K/V re-reads by the other query pairs of a clip are served from L2.
#include <math_constants.h>

#include "tc_common.cuh"
#include "vadb_common.cuh"

namespace vadb {
namespace {

using namespace tc;

constexpr int BM = 128;          // query rows per tile (UMMA M)
constexpr int BKV = 64;          // keys per K/V tile
constexpr int NSTAGE = 3;
constexpr int NTHREADS = 320;    // 8 softmax warps + producer + MMA

constexpr uint32_t Q_HALF_BYTES = BM * 128;          // [128 rows x 64 d] bf16 = 16 KB
constexpr uint32_t Q_TILE_BYTES = 2 * Q_HALF_BYTES;  // two d-halves
constexpr uint32_t KV_HALF_BYTES = BKV * 128;        // [64 keys x 64 d] = 8 KB
constexpr uint32_t K_TILE_BYTES = 2 * KV_HALF_BYTES;
constexpr uint32_t STAGE_BYTES = 2 * K_TILE_BYTES;   // K + V
constexpr uint32_t P_TILE_BYTES = BM * 128;          // [128 rows x 64 keys] bf16 = 16 KB

constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_KV = OFF_Q + 2 * Q_TILE_BYTES;
constexpr uint32_t OFF_P = OFF_KV + NSTAGE * STAGE_BYTES;
constexpr uint32_t OFF_BAR = OFF_P + 2 * P_TILE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256;
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES + 1024;   // slack for 1024-byte alignment

// barrier slots (8 bytes each) at OFF_BAR
enum { B_QFULL = 0, B_KFULL = 1, B_VFULL = 4, B_EMPTY = 7, B_SFULL = 10, B_PFULL = 12, B_OFULL = 14,
       B_COUNT = 16 };

constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_S = 0;      // S_t at columns [64 t, 64 t + 64)
constexpr uint32_t TM_O = 128;    // O_t at columns [128 + 128 t, +128)

constexpr uint32_t IDESC_QK = idesc_bf16(128, BKV, 0, 0);   // A = Q (K-major), B = K (K-major)
constexpr uint32_t IDESC_PV = idesc_bf16(128, 128, 0, 1);   // A = P (K-major), B = V (MN-major)

constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units: rescale O only when the max grew > 2^8

__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
               const __grid_constant__ CUtensorMap tm_v, bf16* __restrict__ O,
               const int32_t* __restrict__ lengths, int T, int npairs) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + OFF_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 8 * B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / npairs, pair = blockIdx.x % npairs;
  const int q0 = pair * 2 * BM;
  int len = lengths ? lengths[b] : T;
  len = min(max(len, 0), T);
  const int nkv = (len + BKV - 1) / BKV;
  const int ntile = (q0 + BM < T) ? 2 : 1;       // second query tile entirely past T: skip it

  if (nkv == 0) {
    // every key masked: softmax of all -inf -> NaN rows, as the reference produces
    for (int i = threadIdx.x; i < 2 * BM * (D / 2); i += NTHREADS) {
      const int r = q0 + i / (D / 2), c = (i % (D / 2)) * 2;
      if (r < T) *reinterpret_cast<uint32_t*>(O + ((long)b * T + r) * D + c) = 0x7FC07FC0u;
    }
    return;
  }

  if (threadIdx.x == 0) {
    mbar_init(BAR(B_QFULL), 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(BAR(B_KFULL + s), 1);
      mbar_init(BAR(B_VFULL + s), 1);
      mbar_init(BAR(B_EMPTY + s), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(BAR(B_SFULL + t), 1);
      mbar_init(BAR(B_PFULL + t), 128);
      mbar_init(BAR(B_OFULL + t), 1);
    }
    mbar_fence_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 9) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      mbar_arrive_expect_tx(BAR(B_QFULL), ntile * Q_TILE_BYTES);
      for (int t = 0; t < ntile; ++t)
        for (int hf = 0; hf < 2; ++hf)
          tma_load_3d(smem_base + OFF_Q + t * Q_TILE_BYTES + hf * Q_HALF_BYTES, &tm_q, BAR(B_QFULL),
                      hf * 64, q0 + t * BM, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % NSTAGE;
        mbar_wait(BAR(B_EMPTY + s), ((j / NSTAGE) & 1) ^ 1, 1);
        const uint32_t kdst = smem_base + OFF_KV + s * STAGE_BYTES;
        mbar_arrive_expect_tx(BAR(B_KFULL + s), K_TILE_BYTES);
        tma_load_3d(kdst, &tm_k, BAR(B_KFULL + s), 0, j * BKV, b);
        tma_load_3d(kdst + KV_HALF_BYTES, &tm_k, BAR(B_KFULL + s), 64, j * BKV, b);
        mbar_arrive_expect_tx(BAR(B_VFULL + s), K_TILE_BYTES);
        tma_load_3d(kdst + K_TILE_BYTES, &tm_v, BAR(B_VFULL + s), 0, j * BKV, b);
        tma_load_3d(kdst + K_TILE_BYTES + KV_HALF_BYTES, &tm_v, BAR(B_VFULL + s), 64, j * BKV, b);
      }
    }
  } else if (warp == 9) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      auto issue_qk = [&](int t, int s) {
        // S_t[128 x 64] = Q_t[128 x 128] K[64 x 128]^T : 8 MMAs of K = 16 (two 64-wide d halves)
        const uint32_t qa = smem_base + OFF_Q + t * Q_TILE_BYTES;
        const uint32_t kb = smem_base + OFF_KV + s * STAGE_BYTES;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem_base + TM_S + 64 * t, desc_kmajor_sw128(qa + hf * Q_HALF_BYTES + kk * 32),
                    desc_kmajor_sw128(kb + hf * KV_HALF_BYTES + kk * 32), IDESC_QK, (hf | kk) != 0);
      };
      auto issue_pv = [&](int t, int s, bool accumulate) {
        // O_t[128 x 128] += P_t[128 x 64] V[64 x 128] : 4 MMAs of K = 16 keys, N = 128 (two d halves)
        const uint32_t pa = smem_base + OFF_P + t * P_TILE_BYTES;
        const uint32_t vb = smem_base + OFF_KV + s * STAGE_BYTES + K_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem_base + TM_O + 128 * t, desc_kmajor_sw128(pa + kk * 32),
                  desc_mnmajor_sw128(vb + kk * 2048, KV_HALF_BYTES, 1024), IDESC_PV,
                  (accumulate || kk != 0) ? 1u : 0u);
      };
      mbar_wait(BAR(B_QFULL), 0, 2);
      mbar_wait(BAR(B_KFULL + 0), 0, 3);
      tc_fence_after();
      for (int t = 0; t < ntile; ++t) {
        issue_qk(t, 0);
        umma_commit(BAR(B_SFULL + t));
      }
      for (int j = 0; j < nkv; ++j) {
        const int s = j % NSTAGE, sn = (j + 1) % NSTAGE;
        for (int t = 0; t < ntile; ++t) {
          mbar_wait(BAR(B_PFULL + t), j & 1, 4);              // P_t(j) in smem, S_t consumed, O_t corrected
          if (t == 0) mbar_wait(BAR(B_VFULL + s), (j / NSTAGE) & 1, 5);
          tc_fence_after();
          issue_pv(t, s, j > 0);
          umma_commit(BAR(B_OFULL + t));
          if (t == ntile - 1) umma_commit(BAR(B_EMPTY + s)); // K/V stage s free once these MMAs retire
          if (j + 1 < nkv) {
            if (t == 0) { mbar_wait(BAR(B_KFULL + sn), ((j + 1) / NSTAGE) & 1, 6); tc_fence_after(); }
            issue_qk(t, sn);
            umma_commit(BAR(B_SFULL + t));
          }
        }
      }
    }
  } else if (warp < 4 * ntile) {
    // ======================= softmax warpgroups =======================
    const int t = warp >> 2;                       // query tile of this warpgroup
    const int row = (warp & 3) * 32 + lane;        // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t ts = tmem_base + lane_addr + TM_S + 64 * t;
    const uint32_t to = tmem_base + lane_addr + TM_O + 128 * t;
    unsigned char* p_row = smem_gen + OFF_P + t * P_TILE_BYTES;
    // softmax in the exp2 domain: p = 2^(s*c - m), c = log2(e)/sqrt(d_head)  (transformer.py:362)
    const float c = 1.4426950408889634f * 0.08838834764831845f;
    float m_ref = -CUDART_INF_F;     // reference max (log2 domain) the stored P/O are relative to
    float l_sum = 0.f;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(BAR(B_SFULL + t), j & 1, 7);
      tc_fence_after();
      uint32_t sv[2][32];
      tmem_ld32(ts, sv[0]);
      tmem_ld32(ts + 32, sv[1]);
      tmem_ld_wait();
      const int kbase = j * BKV;
      if (kbase + BKV > len) {                     // warp-uniform: only the last tile holds masked keys
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + h2 * 32 + i >= len) sv[h2][i] = 0xFF800000u;   // -inf
      }
      // row max with 4 independent chains (ILP: one thread owns the whole row)
      float mx4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(__uint_as_float(sv[0][i]), __uint_as_float(sv[0][i + 1])));
        mx4[((i >> 1) + 2) & 3] = fmaxf(mx4[((i >> 1) + 2) & 3], fmaxf(__uint_as_float(sv[1][i]), __uint_as_float(sv[1][i + 1])));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float mxl = mx * c;
      if (j == 0) {
        m_ref = mxl;
      } else {
        const bool grow = mxl > m_ref + RESCALE_THRESHOLD;
        // previous P V must have retired before O is touched or the P buffer is rewritten
        mbar_wait(BAR(B_OFULL + t), (j - 1) & 1, 8);
        if (__any_sync(0xffffffffu, grow)) {       // warp-uniform: tcgen05.ld/st are .sync.aligned
          tc_fence_after();
          const float m_new = fmaxf(m_ref, mxl);
          const float alpha = (m_new == -CUDART_INF_F) ? 1.f : fast_exp2(m_ref - m_new);
#pragma unroll 1
          for (int cb = 0; cb < 4; ++cb) {
            uint32_t ov[32];
            tmem_ld32(to + cb * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st32(to + cb * 32, ov);
          }
          tmem_st_wait();
          l_sum *= alpha;
          m_ref = m_new;
        }
      }
      const float m_use = (m_ref == -CUDART_INF_F) ? 0.f : m_ref;
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {             // 8 chunks of 8 keys = 16 bytes of bf16 each
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i0 = ch * 8 + e * 2;
          const float p0 = fast_exp2(fmaf(__uint_as_float(sv[i0 >> 5][i0 & 31]), c, -m_use));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sv[(i0 + 1) >> 5][(i0 + 1) & 31]), c, -m_use));
          ps4[e] += p0 + p1;
          pk[e] = pack_bf16(p0, p1);
        }
        *reinterpret_cast<uint4*>(p_row + sw128_offset(row, ch)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      const float psum = (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
      l_sum += psum;
      fence_proxy_async_smem();     // generic-proxy P writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(BAR(B_PFULL + t));
    }

    // epilogue: O_t / l -> bf16 -> global
    mbar_wait(BAR(B_OFULL + t), (nkv - 1) & 1, 9);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    const int qrow = q0 + t * BM + row;
    bf16* orow = O + ((long)b * T + qrow) * D;
#pragma unroll 1
    for (int cb = 0; cb < 4; ++cb) {
      uint32_t ov[32];
      tmem_ld32(to + cb * 32, ov);
      tmem_ld_wait();
      if (qrow < T) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
          pk.y = pack_bf16(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
          pk.z = pack_bf16(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
          pk.w = pack_bf16(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + cb * 32 + g * 8) = pk;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

CUresult make_tmap_bt128(CUtensorMap* map, const void* base, int B, int T, int box_rows) {
  cuuint64_t gdim[3] = {128, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstride[2] = {256, (cuuint64_t)T * 256};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim,
                                gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

cudaError_t launch_attn_tc(const bf16* q, const bf16* k, const bf16* v, bf16* o, const int32_t* lengths,
                           int B, int T, cudaStream_t s, std::string* err) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  CUtensorMap tq, tk, tv;
  CUresult r;
  if ((r = make_tmap_bt128(&tq, q, B, T, BM)) != CUDA_SUCCESS ||
      (r = make_tmap_bt128(&tk, k, B, T, BKV)) != CUDA_SUCCESS ||
      (r = make_tmap_bt128(&tv, v, B, T, BKV)) != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)SMEM_ALLOC);
  if (e != cudaSuccess) return e;
  const int npairs = (T + 2 * BM - 1) / (2 * BM);
  const long grid = (long)B * npairs;
  if (grid > 0x7fffffffL) return cudaErrorInvalidValue;
  attn_tc_kernel<<<(unsigned)grid, NTHREADS, SMEM_ALLOC, s>>>(tq, tk, tv, o, lengths, T, npairs);
  return cudaGetLastError();
}

}  // namespace vadb
