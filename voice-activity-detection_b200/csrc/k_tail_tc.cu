// Fused "layer tail" on the 5th-gen tensor cores (sm_100a): everything of an encoder layer that is per frame,
// from the attention output to the next layer's attention inputs, in ONE persistent kernel
//
//   h'  = h + O Wo^T + bo                                   (vad/modeling/transformer.py:347, :237)
//   h'' = h' + ReLU(LN2(h') W1^T + b1) W2^T + b2            (:234-238, :366-375)
//   not the last layer:  q,k,v = LN1_next(h'') Wqkv_next^T + b     (:235-236, :281-284 of layer l+1)
//   last layer:          p = sigmoid(z1 - z0), logp = log_softmax(z),  z = LN_f(h'') Wc^T + bc
//                                                            (:33, vad/models/self_attention.py:26-27, predictor.py:225)
//
// Per 128-row tile the only HBM traffic is: read O (bf16) and h (fp32), write h'' (fp32) and Q,K,V (bf16) =
// 2048 B/frame; the unfused sequence (out-proj GEMM, FFN kernel, Q/K/V GEMM) moved 4096 B/frame because h and
// its LayerNorm copies went out and came back between the kernels.  Neither LayerNorm output, nor the
// 512-wide hidden activation, nor h' ever leave the SM:
//   * the out-projection accumulates in TMEM; the epilogue adds bo + h, writes (h' + b2) BACK into the same
//     TMEM columns -- they become the initial value of the FFN output accumulator -- and LN2(h') as bf16 into
//     64 further TMEM columns, from where the FFN's first GEMM reads it as its A operand (tcgen05.mma, A in TMEM);
//   * hidden blocks: GEMM -> TMEM -> +b1, ReLU, bf16 -> TMEM -> A operand of the second GEMM (as k_ffn_tc.cu);
//   * h'' is read from TMEM once: stored to HBM, normalised (LN1 of the next layer) into the same 64 TMEM
//     columns, and the Q/K/V GEMMs read it from there.
// Weights (Wo, W1, W2, Wqkv_next: 384 KB in bf16) are pre-packed at load time as 32 KB blocks that are
// byte-for-byte the 128B-swizzled UMMA shared-memory image, in the order the MMA warp consumes them, and
// stream from L2 through a ring with plain bulk copies (cp.async.bulk, no tensor map): measured
// ~100 B/clk/SM with every SM pulling the same blocks (tools/ubench/l2stream.cu), three times what this
// kernel needs.
// The residual stream h lives in HBM in a TILED layout [tile][32 column-quads][128 rows][4 floats], so the
// epilogue threads (one thread = one row, fixed by the TMEM lane it may access) read and write it straight
// from registers with fully coalesced 16-byte accesses: no shared-memory staging for the 64 KB fp32 tile.
//
// CTA = 10 warps: warp 0 producer (weights, O tiles), warp 1 MMA issuer (one elected lane), warps 2-9 epilogue
// (two per TMEM lane quarter, 64 columns each).  TMEM: three 128-column regions that rotate roles from tile to
// tile (accumulator, hidden slot 0, hidden slot 1; then Q, K, V accumulators) + 64 columns for the LayerNorm
// operand.
#include "tc_common.cuh"
#include "vadb_common.cuh"

namespace vadb {
namespace {

using namespace tc;

constexpr int NTHREADS = 320;
constexpr int N_EPI_WARPS = 8;
constexpr int NW = 4;                              // weight ring stages
constexpr uint32_t BLK_BYTES = 128 * 128 * 2;      // [128 x 128] bf16 block = two SW128 halves of 16 KB
constexpr uint32_t HALF_BYTES = 128 * 128;
constexpr uint32_t STG_BYTES = 32 * 128;           // per-warp staging slab: 32 rows x 128 B
constexpr uint32_t IDESC = idesc_bf16(128, 128, 0, 0);

constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_O = OFF_W + NW * BLK_BYTES;
constexpr uint32_t OFF_STG = OFF_O + BLK_BYTES;
constexpr uint32_t OFF_BAR = OFF_STG + 2 * N_EPI_WARPS * STG_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 2048;      // barriers + LayerNorm exchange
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

enum { B_WFULL = 0, B_WEMPTY = 4, B_OFULL = 8, B_OEMPTY = 9, B_ACC1 = 10, B_ALN = 11, B_HIDFULL = 12, B_HIDBF = 14,
       B_OUTFULL = 16, B_QKVFULL = 17, B_REGFREE = 20, B_COUNT = 23 };
static_assert(NW <= 4 && 8 * B_COUNT + 4 <= 256, "barrier region");

constexpr uint32_t TM_ALN = 384;    // bf16 LayerNorm operand: 64 columns

struct TailParams {
  int M;
  int has_qkv;                 // 1: emit q,k,v of the next layer; 0: last layer (classifier epilogue)
  const unsigned char* wpack;  // packed weight blocks of this layer, consumption order (12 or 9 blocks of 32 KB)
  float* h;                    // tiled fp32 residual stream, updated in place (not written by the last layer)
  const float* bo;             // [128]
  const float* b1;             // [512]
  const float* b2;             // [128]
  const float* ln2_g;          // pre-LN of this layer's feed-forward sublayer
  const float* ln2_b;
  const float* ln1n_g;         // pre-LN of the NEXT layer's attention sublayer
  const float* ln1n_b;
  const float* bqkv;           // [384] next layer's q|k|v biases
  const float* cls_g;          // last layer: final LayerNorm, classifier
  const float* cls_b;
  const float* cls_w;          // [2,128]
  const float* cls_bias;       // [2]
  float* prob;                 // [M] or nullptr
  float* logp;                 // [M,2] or nullptr
  int logp_vec;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// FFN weight blocks in consumption order (GEMM1 runs one hidden block ahead of GEMM2 so the ReLU epilogue of
// block nb overlaps GEMM1 of block nb+1):  W1[0] W1[1] W2[0] W1[2] W2[1] W1[3] W2[2] W2[3]
__device__ __forceinline__ void ffn_seq(int q, int& is_w2, int& nb) {
  is_w2 = (0b11010100 >> q) & 1;
  nb = (0xED84 >> (2 * q)) & 3;
}

// LayerNorm statistics of a row whose 128 columns are split between two threads (same lane, the two warps of a
// TMEM lane quarter): partial sums exchanged through shared memory behind a 64-thread named barrier.
__device__ __forceinline__ void row_stats(const uint32_t (&v)[2][32], volatile float* xs, int row, int hsel, int q,
                                          float& mean, float& rstd, bool exact_div) {
  float s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s1 += __uint_as_float(v[i >> 5][i & 31]);
  xs[row * 2 + hsel] = s1;
  asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
  mean = (s1 + xs[row * 2 + (hsel ^ 1)]) * (1.0f / 128.0f);
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    const float d = __uint_as_float(v[i >> 5][i & 31]) - mean;
    s2 = fmaf(d, d, s2);
  }
  xs[256 + row * 2 + hsel] = s2;
  asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
  const float var = (s2 + xs[256 + row * 2 + (hsel ^ 1)]) * (1.0f / 128.0f) + LN_EPS;
  rstd = exact_div ? 1.0f / sqrtf(var) : rsqrtf(var);
}

// LayerNorm(row) * gamma + beta -> bf16 pairs -> this thread's 32 columns of the TMEM LayerNorm operand
__device__ __forceinline__ void emit_ln_tmem(const uint32_t (&v)[2][32], float mean, float rstd, const float* g,
                                             const float* b, int hsel, uint32_t taddr) {
  const float4* gp = reinterpret_cast<const float4*>(g + hsel * 64);
  const float4* bp = reinterpret_cast<const float4*>(b + hsel * 64);
  uint32_t pk[32];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float4 g4 = __ldg(gp + c), b4 = __ldg(bp + c);
    const int i0 = c * 4;
    const float y0 = (__uint_as_float(v[i0 >> 5][i0 & 31]) - mean) * rstd * g4.x + b4.x;
    const float y1 = (__uint_as_float(v[(i0 + 1) >> 5][(i0 + 1) & 31]) - mean) * rstd * g4.y + b4.y;
    const float y2 = (__uint_as_float(v[(i0 + 2) >> 5][(i0 + 2) & 31]) - mean) * rstd * g4.z + b4.z;
    const float y3 = (__uint_as_float(v[(i0 + 3) >> 5][(i0 + 3) & 31]) - mean) * rstd * g4.w + b4.w;
    pk[2 * c] = pack_bf16(y0, y1);
    pk[2 * c + 1] = pack_bf16(y2, y3);
  }
  tmem_st32(taddr, pk);
}

template <bool HAS_QKV>
__global__ void __launch_bounds__(NTHREADS, 1)
tail_tc_kernel(const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_q,
               const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const TailParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  unsigned char* smem_gen = smem_raw;
  if ((smem_base & 1023u) != 0) {   // 128B-swizzled tiles need 1024-byte alignment
    if (threadIdx.x == 0) printf("vadb: tail kernel shared memory window not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t bar0 = smem_base + OFF_BAR;
  volatile float* xs = reinterpret_cast<volatile float*>(smem_gen + OFF_BAR + 256);   // [2][128][2] LN exchange
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 8 * B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) >> 7;
  const int n_wblk = HAS_QKV ? 12 : 9;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NW; ++s) { mbar_init(BAR(B_WFULL + s), 1); mbar_init(BAR(B_WEMPTY + s), 1); }
    mbar_init(BAR(B_OFULL), 1); mbar_init(BAR(B_OEMPTY), 1);
    mbar_init(BAR(B_ACC1), 1);
    mbar_init(BAR(B_ALN), N_EPI_WARPS);
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_HIDFULL + s), 1); mbar_init(BAR(B_HIDBF + s), N_EPI_WARPS); }
    mbar_init(BAR(B_OUTFULL), 1);
    for (int j = 0; j < 3; ++j) { mbar_init(BAR(B_QKVFULL + j), 1); mbar_init(BAR(B_REGFREE + j), N_EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_o);
    if (HAS_QKV) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); }
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 0) {
    // ======================= producer: weight blocks (bulk copies), O tiles (TMA) =======================
    if (lane == 0) {
      int wc = 0;                                  // weight blocks issued so far
      auto load_w = [&](int b) {
        const int s = wc % NW;
        mbar_wait(BAR(B_WEMPTY + s), ((wc / NW) & 1) ^ 1, 40);
        mbar_arrive_expect_tx(BAR(B_WFULL + s), BLK_BYTES);
        bulk_load(smem_base + OFF_W + s * BLK_BYTES, p.wpack + (size_t)b * BLK_BYTES, BLK_BYTES, BAR(B_WFULL + s));
        ++wc;
      };
      auto load_o = [&](int tile, int n) {
        if (n > 0) mbar_wait(BAR(B_OEMPTY), (n - 1) & 1, 41);      // out-projection MMAs of the previous tile retired
        mbar_arrive_expect_tx(BAR(B_OFULL), BLK_BYTES);
        tma_load_2d(smem_base + OFF_O, &tm_o, BAR(B_OFULL), 0, tile * 128);
        tma_load_2d(smem_base + OFF_O + HALF_BYTES, &tm_o, BAR(B_OFULL), 64, tile * 128);
      };
      // the weights do not depend on the previous kernel: the first ring fill overlaps its tail (PDL)
      int pre = 0;
      if ((int)blockIdx.x < n_tiles)
        for (; pre < NW && pre < n_wblk; ++pre) load_w(pre);
      pdl_wait();
      int n = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
        if (n == 0) load_o(tile, 0);
        const int next = tile + (int)gridDim.x;
        if (next < n_tiles)      // next tile's residual rows: one contiguous 64 KB block in the tiled layout -> L2
          for (int c = 0; c < 4; ++c) l2_prefetch(p.h + (size_t)next * 16384 + (size_t)c * 4096, 16384);
        for (int b = (n == 0 ? pre : 0); b < n_wblk; ++b) {
          load_w(b);
          if (b == 5 && next < n_tiles) load_o(next, n + 1);       // O buffer was released early in this tile
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp walks, one elected lane issues) =======================
    pdl_wait();
    const uint32_t o_lo = desc_lo(smem_base + OFF_O, 16);
    const uint32_t w_lo0 = desc_lo(smem_base + OFF_W, 16);
    const uint32_t t_aln = tmem_base + TM_ALN;
    int wc = 0, n = 0, aln_use = 0;
    int hid_uses[2] = {0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const uint32_t r_acc = tmem_base + 128u * (uint32_t)(n % 3);
      const uint32_t r_hid[2] = {tmem_base + 128u * (uint32_t)((n + 1) % 3), tmem_base + 128u * (uint32_t)((n + 2) % 3)};
      // Regions of this tile = accumulators of the previous tile's q (-> acc), k (-> hidden slot 0), v or classifier
      // read (-> hidden slot 1).  With q/k/v every REGFREE barrier completes one phase per tile; in the last-layer
      // variant only the accumulator of a tile is released: one phase per barrier every three tiles.
      const uint32_t prev_par = HAS_QKV ? (uint32_t)((n - 1) & 1) : (uint32_t)(((n - 1) / 3) & 1);
      // ---- out-projection: acc = O Wo^T  (both operands from shared memory)
      mbar_wait(BAR(B_OFULL), n & 1, 42);
      if (n > 0 && HAS_QKV) mbar_wait(BAR(B_REGFREE + n % 3), prev_par, 43);           // q of the previous tile drained
      {
        const int ws = wc % NW;
        mbar_wait(BAR(B_WFULL + ws), (wc / NW) & 1, 44);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = w_lo0 + (uint32_t)ws * (BLK_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ss_lh(r_acc, o_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                       w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2, DESC_HI_SW128, IDESC, kk != 0 ? 1u : 0u);
          umma_commit(BAR(B_OEMPTY));
          umma_commit(BAR(B_WEMPTY + ws));
          umma_commit(BAR(B_ACC1));
        }
        __syncwarp();
        ++wc;
      }
      // ---- feed-forward: LN2 operand in TMEM (written by the epilogue warps), hidden blocks through TMEM
      mbar_wait(BAR(B_ALN), aln_use & 1, 45);
      ++aln_use;
      for (int q = 0; q < 8; ++q, ++wc) {
        int is_w2, nb;
        ffn_seq(q, is_w2, nb);
        const int ws = wc % NW, hs = nb & 1;
        mbar_wait(BAR(B_WFULL + ws), (wc / NW) & 1, 46);
        if (!is_w2 && n > 0) {
          // the hidden slots are the previous tile's K / V (or classifier) accumulators: drained?
          if (nb == 0 && HAS_QKV) mbar_wait(BAR(B_REGFREE + (n + 1) % 3), prev_par, 47);
          if (nb == 1) mbar_wait(BAR(B_REGFREE + (n + 2) % 3), prev_par, 48);
        }
        if (is_w2) mbar_wait(BAR(B_HIDBF + hs), (hid_uses[hs] - 1) & 1, 49);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = w_lo0 + (uint32_t)ws * (BLK_BYTES >> 4);
          if (!is_w2) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lh(r_hid[hs], t_aln + kk * 8, w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                         DESC_HI_SW128, IDESC, kk != 0 ? 1u : 0u);
            umma_commit(BAR(B_HIDFULL + hs));
          } else {
            // acc already holds h' + b2 (written by the epilogue warps): always accumulate
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lh(r_acc, r_hid[hs] + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8),
                         w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2, DESC_HI_SW128, IDESC, 1u);
            if (nb == 3) umma_commit(BAR(B_OUTFULL));
          }
          umma_commit(BAR(B_WEMPTY + ws));
        }
        __syncwarp();
        if (!is_w2) hid_uses[hs]++;
      }
      // ---- next layer's q, k, v = LN1_next(h'') Wqkv^T: accumulators in the two hidden slots and the old acc
      if (HAS_QKV) {
        mbar_wait(BAR(B_ALN), aln_use & 1, 50);
        ++aln_use;
        for (int j = 0; j < 3; ++j, ++wc) {
          const int ws = wc % NW;
          const uint32_t r_out = j == 0 ? r_hid[0] : j == 1 ? r_hid[1] : r_acc;
          mbar_wait(BAR(B_WFULL + ws), (wc / NW) & 1, 51);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t w_lo = w_lo0 + (uint32_t)ws * (BLK_BYTES >> 4);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lh(r_out, t_aln + kk * 8, w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                         DESC_HI_SW128, IDESC, kk != 0 ? 1u : 0u);
            umma_commit(BAR(B_QKVFULL + j));
            umma_commit(BAR(B_WEMPTY + ws));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================= epilogue warps: two per TMEM lane quarter, 64 columns each =======================
    pdl_wait();
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t stg_off0 = OFF_STG + (uint32_t)(warp - 2) * 2 * STG_BYTES;      // this warp's two staging slabs
    const uint32_t t_aln = tmem_base + lane_addr + TM_ALN + 32u * hsel;
    int n = 0, unit = 0;
    int hid_uses[2] = {0, 0};
    // this thread's 64 columns of residual row `row` of a tile: 16 coalesced float4 (tiled layout)
    float4 hreg[16];
    auto load_h = [&](int tile) {
      const float4* src = reinterpret_cast<const float4*>(p.h) + (size_t)tile * 4096 + (size_t)(hsel * 16) * 128 + row;
#pragma unroll
      for (int i = 0; i < 16; ++i) hreg[i] = __ldcg(src + i * 128);
    };
    if ((int)blockIdx.x < n_tiles) load_h(blockIdx.x);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const uint32_t r_acc = tmem_base + lane_addr + 128u * (uint32_t)(n % 3) + 64u * hsel;
      const uint32_t r_hid[2] = {tmem_base + lane_addr + 128u * (uint32_t)((n + 1) % 3) + 64u * hsel,
                                 tmem_base + lane_addr + 128u * (uint32_t)((n + 2) % 3) + 64u * hsel};
      uint32_t v[2][32];
      // ---- epilogue 1: h' = acc + bo + h;  acc <- h' + b2;  LN2(h') -> TMEM operand
      mbar_wait(BAR(B_ACC1), n & 1, 52);
      tc_fence_after();
      tmem_ld32(r_acc, v[0]);
      tmem_ld32(r_acc + 32, v[1]);
      tmem_ld_wait();
      {
        const float4* bp = reinterpret_cast<const float4*>(p.bo + hsel * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b4 = __ldg(bp + i);
          uint32_t* s4 = &v[i >> 3][(i & 7) * 4];
          s4[0] = __float_as_uint(__uint_as_float(s4[0]) + b4.x + hreg[i].x);
          s4[1] = __float_as_uint(__uint_as_float(s4[1]) + b4.y + hreg[i].y);
          s4[2] = __float_as_uint(__uint_as_float(s4[2]) + b4.z + hreg[i].z);
          s4[3] = __float_as_uint(__uint_as_float(s4[3]) + b4.w + hreg[i].w);
        }
      }
      {
        float mean, rstd;
        row_stats(v, xs, row, hsel, q, mean, rstd, false);
        emit_ln_tmem(v, mean, rstd, p.ln2_g, p.ln2_b, hsel, t_aln);
        const float4* bp = reinterpret_cast<const float4*>(p.b2 + hsel * 64);
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bp + cb * 8 + i);
            w[4 * i] = __float_as_uint(__uint_as_float(v[cb][4 * i]) + b4.x);
            w[4 * i + 1] = __float_as_uint(__uint_as_float(v[cb][4 * i + 1]) + b4.y);
            w[4 * i + 2] = __float_as_uint(__uint_as_float(v[cb][4 * i + 2]) + b4.z);
            w[4 * i + 3] = __float_as_uint(__uint_as_float(v[cb][4 * i + 3]) + b4.w);
          }
          tmem_st32(r_acc + cb * 32, w);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ALN));
      }
      // ---- hidden blocks: + b1, ReLU, bf16 pairs written back over the head of this thread's own 64 columns
      for (int nb = 0; nb < 4; ++nb) {
        const int hs = nb & 1;
        mbar_wait(BAR(B_HIDFULL + hs), hid_uses[hs] & 1, 53);
        hid_uses[hs]++;
        tc_fence_after();
        tmem_ld32(r_hid[hs], v[0]);
        tmem_ld32(r_hid[hs] + 32, v[1]);
        tmem_ld_wait();
        uint32_t pk[32];
        const float4* bp = reinterpret_cast<const float4*>(p.b1 + nb * 128 + hsel * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b4 = __ldg(bp + i);
          const uint32_t* s4 = &v[i >> 3][(i & 7) * 4];
          const float f0 = fmaxf(__uint_as_float(s4[0]) + b4.x, 0.f), f1 = fmaxf(__uint_as_float(s4[1]) + b4.y, 0.f);
          const float f2 = fmaxf(__uint_as_float(s4[2]) + b4.z, 0.f), f3 = fmaxf(__uint_as_float(s4[3]) + b4.w, 0.f);
          pk[2 * i] = pack_bf16(f0, f1);
          pk[2 * i + 1] = pack_bf16(f2, f3);
        }
        tmem_st32(r_hid[hs], pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HIDBF + hs));
      }
      // ---- final epilogue: acc = h'' (residual and b2 were in the accumulator from the start)
      mbar_wait(BAR(B_OUTFULL), n & 1, 54);
      tc_fence_after();
      tmem_ld32(r_acc, v[0]);
      tmem_ld32(r_acc + 32, v[1]);
      tmem_ld_wait();
      const int next = tile + (int)gridDim.x;
      if (HAS_QKV) {
        {
          float4* dst = reinterpret_cast<float4*>(p.h) + (size_t)tile * 4096 + (size_t)(hsel * 16) * 128 + row;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t* s4 = &v[i >> 3][(i & 7) * 4];
            dst[i * 128] = make_float4(__uint_as_float(s4[0]), __uint_as_float(s4[1]), __uint_as_float(s4[2]), __uint_as_float(s4[3]));
          }
        }
        float mean, rstd;
        row_stats(v, xs, row, hsel, q, mean, rstd, false);
        emit_ln_tmem(v, mean, rstd, p.ln1n_g, p.ln1n_b, hsel, t_aln);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ALN));
        if (next < n_tiles) load_h(next);          // lands while the q/k/v epilogues run
        // ---- q, k, v: + bias -> bf16 -> swizzled slab -> TMA store (rows past M clipped by the tensor map)
        for (int j = 0; j < 3; ++j) {
          const uint32_t r_out = j == 0 ? r_hid[0] : j == 1 ? r_hid[1] : r_acc;
          const int region = j == 0 ? (n + 1) % 3 : j == 1 ? (n + 2) % 3 : n % 3;
          mbar_wait(BAR(B_QKVFULL + j), n & 1, 55);
          tc_fence_after();
          const uint32_t so = stg_off0 + (unit & 1) * STG_BYTES;
          ++unit;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          const float4* bp = reinterpret_cast<const float4*>(p.bqkv + j * 128 + hsel * 64);
          // 32 columns at a time: the next tile's residual rows (64 registers) are in flight across this loop
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            uint32_t u[32];
            tmem_ld32(r_out + cb * 32, u);
            tmem_ld_wait();
            if (cb == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(BAR(B_REGFREE + region));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 ba = __ldg(bp + cb * 8 + 2 * c), bb = __ldg(bp + cb * 8 + 2 * c + 1);
              const uint32_t* f = &u[c * 8];
              *reinterpret_cast<uint4*>(smem_gen + so + sw128_offset(lane, cb * 4 + c)) =
                  make_uint4(pack_bf16(__uint_as_float(f[0]) + ba.x, __uint_as_float(f[1]) + ba.y),
                             pack_bf16(__uint_as_float(f[2]) + ba.z, __uint_as_float(f[3]) + ba.w),
                             pack_bf16(__uint_as_float(f[4]) + bb.x, __uint_as_float(f[5]) + bb.y),
                             pack_bf16(__uint_as_float(f[6]) + bb.z, __uint_as_float(f[7]) + bb.w));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(j == 0 ? &tm_q : j == 1 ? &tm_k : &tm_v, smem_base + so, hsel * 64, tile * 128 + q * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
        // Last layer: the encoder's final LayerNorm (transformer.py:33), the classifier and the log-softmax
        // (self_attention.py:26-27) and the caller's softmax(...)[...,1] = sigmoid(z1 - z0)
        // (predictor.py:225,257-258) on the row that is still in registers: h is never written.
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_REGFREE + n % 3));
        float mean, rstd;
        row_stats(v, xs, row, hsel, q, mean, rstd, true);
        const float4* gp = reinterpret_cast<const float4*>(p.cls_g + hsel * 64);
        const float4* bp2 = reinterpret_cast<const float4*>(p.cls_b + hsel * 64);
        const float4* w0p = reinterpret_cast<const float4*>(p.cls_w + hsel * 64);
        const float4* w1p = reinterpret_cast<const float4*>(p.cls_w + 128 + hsel * 64);
        float z0 = 0.f, z1 = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 g4 = __ldg(gp + c), b4 = __ldg(bp2 + c), u4 = __ldg(w0p + c), w4 = __ldg(w1p + c);
          const int i0 = c * 4;
          const float y0 = (__uint_as_float(v[i0 >> 5][i0 & 31]) - mean) * rstd * g4.x + b4.x;
          const float y1 = (__uint_as_float(v[(i0 + 1) >> 5][(i0 + 1) & 31]) - mean) * rstd * g4.y + b4.y;
          const float y2 = (__uint_as_float(v[(i0 + 2) >> 5][(i0 + 2) & 31]) - mean) * rstd * g4.z + b4.z;
          const float y3 = (__uint_as_float(v[(i0 + 3) >> 5][(i0 + 3) & 31]) - mean) * rstd * g4.w + b4.w;
          z0 += y0 * u4.x + y1 * u4.y + y2 * u4.z + y3 * u4.w;
          z1 += y0 * w4.x + y1 * w4.y + y2 * w4.z + y3 * w4.w;
        }
        // the upper-half thread hands its partial dots over in the lower-half thread's own (now dead)
        // exchange slots, so the next tile's statistics cannot race with this read
        if (hsel == 1) { xs[row * 2] = z0; xs[256 + row * 2] = z1; }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const long grow = (long)tile * 128 + row;
        if (hsel == 0 && grow < p.M) {
          const float a0 = z0 + xs[row * 2] + __ldg(p.cls_bias);
          const float a1 = z1 + xs[256 + row * 2] + __ldg(p.cls_bias + 1);
          const float mx = fmaxf(a0, a1);
          const float lse = mx + log1pf(expf(-fabsf(a1 - a0)));        // log_softmax([a0, a1]), stable
          if (p.logp) {
            if (p.logp_vec) *reinterpret_cast<float2*>(p.logp + grow * 2) = make_float2(a0 - lse, a1 - lse);
            else { p.logp[grow * 2] = a0 - lse; p.logp[grow * 2 + 1] = a1 - lse; }
          }
          if (p.prob) p.prob[grow] = 1.0f / (1.0f + expf(a0 - a1));     // softmax(logp)[1]
        }
        // the partner may only overwrite the exchange slots (next tile's statistics) after this read
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        if (next < n_tiles) load_h(next);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// fp32 [N_total, K_total] row-major weight -> one 32 KB block: rows n0..n0+127, columns k0..k0+127 as bf16 in the
// UMMA K-major 128B-swizzled shared-memory image (two 16 KB halves of 64 k each)
__global__ void pack_block_kernel(const float* __restrict__ w, int ld, int n0, int k0, unsigned char* __restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 8 consecutive k of one row
  if (t >= 128 * 16) return;
  const int row = t >> 4, k8 = t & 15;
  const float* src = w + (size_t)(n0 + row) * ld + k0 + k8 * 8;
  uint4 o;
  o.x = pack_bf16(src[0], src[1]); o.y = pack_bf16(src[2], src[3]);
  o.z = pack_bf16(src[4], src[5]); o.w = pack_bf16(src[6], src[7]);
  const int half = k8 >> 3, chunk = k8 & 7;
  *reinterpret_cast<uint4*>(dst + half * HALF_BYTES + sw128_offset(row, chunk)) = o;
}

CUresult tmap2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esz, long rows, long cols,
                int box_cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return encode_tiled(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

size_t tail_pack_bytes() { return (size_t)12 * BLK_BYTES; }

cudaError_t launch_tail_pack(const float* wo, const float* w1, const float* w2, const float* wqkv_next,
                             unsigned char* dst, cudaStream_t s) {
  int b = 0;
  auto blk = [&](const float* w, int ld, int n0, int k0) {
    pack_block_kernel<<<8, 256, 0, s>>>(w, ld, n0, k0, dst + (size_t)b * BLK_BYTES);
    ++b;
  };
  blk(wo, D, 0, 0);
  // W1[0] W1[1] W2[0] W1[2] W2[1] W1[3] W2[2] W2[3]  (ffn_seq)
  blk(w1, D, 0, 0); blk(w1, D, 128, 0); blk(w2, DFF, 0, 0); blk(w1, D, 256, 0);
  blk(w2, DFF, 0, 128); blk(w1, D, 384, 0); blk(w2, DFF, 0, 256); blk(w2, DFF, 0, 384);
  if (wqkv_next)
    for (int j = 0; j < 3; ++j) blk(wqkv_next, D, j * 128, 0);
  return cudaGetLastError();
}

cudaError_t launch_tail_tc(const TailTcArgs& a, int num_sms, cudaStream_t s, std::string* err) {
  if (a.M <= 0) return cudaSuccess;
  const bool has_qkv = a.q != nullptr;
  CUtensorMap to, tq, tk, tv;
  CUresult r = tmap2d(&to, a.o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 64, 128);
  if (has_qkv) {
    if (r == CUDA_SUCCESS) r = tmap2d(&tq, a.q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 64, 32);
    if (r == CUDA_SUCCESS) r = tmap2d(&tk, a.k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 64, 32);
    if (r == CUDA_SUCCESS) r = tmap2d(&tv, a.v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 64, 32);
  } else {
    tq = to; tk = to; tv = to;
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return cudaErrorInvalidValue;
  }
  TailParams p = {};
  p.M = a.M; p.has_qkv = has_qkv ? 1 : 0; p.wpack = a.wpack; p.h = a.h;
  p.bo = a.bo; p.b1 = a.b1; p.b2 = a.b2; p.ln2_g = a.ln2_g; p.ln2_b = a.ln2_b;
  p.ln1n_g = a.ln1n_g; p.ln1n_b = a.ln1n_b; p.bqkv = a.bqkv;
  p.cls_g = a.cls_ln_g; p.cls_b = a.cls_ln_b; p.cls_w = a.cls_w; p.cls_bias = a.cls_bias;
  p.prob = a.prob; p.logp = a.logp;
  p.logp_vec = (reinterpret_cast<uintptr_t>(a.logp) % 8) == 0;
  if (has_qkv ? (!p.ln1n_g || !p.ln1n_b || !p.bqkv) : (!p.cls_g || !p.cls_b || !p.cls_w || !p.cls_bias))
    return cudaErrorInvalidValue;
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(tail_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
  }
  const int n_tiles = (a.M + 127) / 128;
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  {
    cudaError_t e = has_qkv ? launch_k(tail_tc_kernel<true>, grid, NTHREADS, SMEM_BYTES, s, to, tq, tk, tv, p)
                            : launch_k(tail_tc_kernel<false>, grid, NTHREADS, SMEM_BYTES, s, to, tq, tk, tv, p);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace vadb
