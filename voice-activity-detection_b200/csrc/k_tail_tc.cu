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
//     TMEM columns -- they become the initial value of the FFN output accumulator -- and the normalised row as
//     bf16 into 64 further TMEM columns, from where the FFN's first GEMM reads it as its A operand
//     (tcgen05.mma with A in TMEM);
//   * hidden blocks: GEMM -> TMEM -> +b1, ReLU, bf16 -> TMEM -> A operand of the second GEMM;
//   * h'' is read from TMEM once: staged for the store, normalised into the same 64 TMEM columns, and the
//     Q/K/V GEMMs read it from there.
// LayerNorm's affine part is folded into the following Linear at load time (W' = W diag(gamma),
// b' = b + W beta), so the epilogue only emits (x - mean) * rstd; ReLU is computed as t + |t| = 2 ReLU(t) on the
// FMA pipe with W2 pre-scaled by 0.5 (exact: a power of two), and fp32 pairs are rounded to bf16 on the integer
// ALU (cvt.rn.bf16x2.f32 issues once per ~9 cycles per SM sub-partition on B200).
// Weights (Wo, W1', W2/2, Wqkv'_next: 384 KB in bf16) are pre-packed at load time as 32 KB blocks that are
// byte-for-byte the 128B-swizzled UMMA shared-memory image, in the order the MMA warp consumes them, and
// stream from L2 through a ring with plain bulk copies (cp.async.bulk, no tensor map): measured
// ~100 B/clk/SM with every SM pulling the same blocks (tools/ubench/l2stream.cu), three times what this
// kernel needs.
// The residual stream h lives in HBM in a TILED layout [tile][32 column-quads][128 rows][4 floats]: a tile is one
// contiguous 64 KB block that moves between HBM and shared memory with ONE bulk copy each way, and the epilogue
// threads (one thread = one row, fixed by the TMEM lane it may access) read/write it in shared memory without
// bank conflicts.
//
// CTA = 18 warps: warp 0 producer (weights, O tiles), warp 1 MMA issuer (one elected lane), warps 2-17 epilogue
// (four per TMEM lane quarter, 32 columns each: the epilogue arithmetic is latency-bound with fewer warps).
// TMEM: three 128-column regions that rotate roles from tile to tile (accumulator, hidden slot 0, hidden slot 1;
// then Q, K, V accumulators) + 64 columns for the LayerNorm operand + 32 columns through which the four threads
// of a row exchange their partial LayerNorm sums.
#include <stdlib.h>

#include "tc_common.cuh"
#include "vadb_common.cuh"

namespace vadb {
namespace {

using namespace tc;

constexpr int N_EPI_WARPS = 16;
constexpr int NTHREADS = 32 * (3 + N_EPI_WARPS);   // + producer, MMA issuer, residual-tile mover (last warp)
constexpr int NW = 3;                              // weight ring stages (2 stages + two staging slabs per warp: +4 % per forward)
constexpr int NSTG = 1;                            // staging slabs per epilogue warp
constexpr uint32_t BLK_BYTES = 128 * 128 * 2;      // [128 x 128] bf16 block = two SW128 halves of 16 KB
constexpr uint32_t HALF_BYTES = 128 * 128;
constexpr uint32_t H_BYTES = 128 * 128 * 4;        // one fp32 residual tile
constexpr uint32_t STG_BYTES = 32 * 64;            // per-warp staging slab: 32 rows x 64 B (32 bf16 columns), 64B swizzle
constexpr uint32_t IDESC = idesc_bf16(128, 128, 0, 0);

constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_O = OFF_W + NW * BLK_BYTES;
constexpr uint32_t OFF_H = OFF_O + BLK_BYTES;
constexpr uint32_t OFF_STG = OFF_H + H_BYTES;
constexpr uint32_t OFF_BAR = OFF_STG + NSTG * N_EPI_WARPS * STG_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

enum { B_WFULL = 0, B_WEMPTY = 4, B_OFULL = 8, B_OEMPTY = 9, B_ACC1 = 10, B_ALN = 11, B_HIDFULL = 12, B_HIDBF = 14,
       B_OUTFULL = 16, B_QKVFULL = 17, B_REGFREE = 20, B_HFULL = 23, B_HFREE = 24, B_HSTAGED = 25, B_COUNT = 26 };
static_assert(NW <= 4 && 8 * B_COUNT + 4 <= 256, "barrier region");

constexpr uint32_t TM_ALN = 384;    // bf16 LayerNorm operand: 64 columns
constexpr uint32_t TM_X = 448;      // partial-sum exchange: 2 buffers x 16 columns (4 threads x 4 floats per row)

enum { NB_QUARTER = 1 /* +q: the four warps of a TMEM lane quarter */, NB_EPI = 5 /* all epilogue warps */ };

struct TailParams {
  int M;
  const unsigned char* wpack;  // packed weight blocks of this layer, consumption order (12 or 9 blocks of 32 KB)
  float* h;                    // tiled fp32 residual stream, updated in place (not written by the last layer)
  const float* bo;             // [128]
  const float* b1p;            // [512] b1 + W1 beta2
  const float* b2;             // [128]
  const float* bqkvp;          // [384] next layer's q|k|v biases + Wqkv beta1_next
  const float* cls_gw;         // last layer: [2,128] gamma_f * Wc rows, then {sum gw0, sum gw1, beta_f.Wc0 + bc0, beta_f.Wc1 + bc1}
  float* prob;                 // [M] or nullptr
  float* logp;                 // [M,2] or nullptr
  int logp_vec;
  bf16* q; bf16* k; bf16* v;   // next layer's attention inputs [M,128]
  // head mode
  const float* pe_tiled;       // positional-encoding table / sqrt(d) in the tiled layout, pe_tiles tiles of 128 positions
  int pe_tiles;                // T / 128 (head mode needs T % 128 == 0): tile t of the batch uses PE tile t % pe_tiles
  int x_col1;                  // column coordinate of the second half of the feature tile (64: bf16, 32: fp32 as tf32)
  int x_tf32;                  // features and W_in are fp32 in shared memory, multiplied as tf32
  int stagger_cycles;          // odd CTAs start this many cycles late (de-synchronises the per-tile store bursts)
  long long* trace;            // developer instrumentation (VADB_TAIL_TRACE=1): clock64 stamps of CTA 0's third tile
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_st4v(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// FFN weight blocks in consumption order (GEMM1 runs one hidden block ahead of GEMM2 so the ReLU epilogue of
// block nb overlaps GEMM1 of block nb+1):  W1[0] W1[1] W2[0] W1[2] W2[1] W1[3] W2[2] W2[3]
__device__ __forceinline__ void ffn_seq(int q, int& is_w2, int& nb) {
  is_w2 = (0b11010100 >> q) & 1;
  nb = (0xED84 >> (2 * q)) & 3;
}

#define TT(slot) do { if (TRACE && blockIdx.x == 0 && n == 2 && lane == 0 && p.trace) p.trace[(slot)] = clock64(); } while (0)

// MODE 0: last layer (classifier epilogue), 1: layer tail + next layer's q/k/v, 2: HEAD of the model -- the same
// machinery for  h0 = x W_in^T + b_in + PE/sqrt(d)  (self_attention.py:13-14, transformer.py:401) followed by layer 0's
// q,k,v = LN1(h0) Wqkv^T + b: the "O tile" is the feature tile, the "residual tile" is the positional-encoding tile (kept
// in the same tiled layout), there is no feed-forward phase, and h0 is what gets stored.
template <int MODE, bool TRACE>
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
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 8 * B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) >> 7;
  constexpr bool HAS_QKV = MODE >= 1, HEAD = MODE == 2;
  // The last-layer variant stages no q|k|v: its 32 KB slab region is a fourth weight-ring stage (the FFN phase of that
  // variant otherwise stalls ~1.7 k cycles per tile on a block whose stage was released too late for the L2 latency)
  constexpr int NWK = HAS_QKV ? NW : 4;
  auto wslot_blk = [](int s) { return s == 3 ? (int)((OFF_STG - OFF_W) / BLK_BYTES) : s; };
  const int n_wblk = HEAD ? 4 : HAS_QKV ? 12 : 9;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(BAR(B_WFULL + s), 1); mbar_init(BAR(B_WEMPTY + s), 1); }
    mbar_init(BAR(B_OFULL), 1); mbar_init(BAR(B_OEMPTY), 1);
    mbar_init(BAR(B_ACC1), 1);
    mbar_init(BAR(B_ALN), N_EPI_WARPS);
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_HIDFULL + s), 1); mbar_init(BAR(B_HIDBF + s), N_EPI_WARPS); }
    mbar_init(BAR(B_OUTFULL), 1);
    for (int j = 0; j < 3; ++j) { mbar_init(BAR(B_QKVFULL + j), 1); mbar_init(BAR(B_REGFREE + j), N_EPI_WARPS); }
    mbar_init(BAR(B_HFULL), 1);
    mbar_init(BAR(B_HFREE), N_EPI_WARPS);
    mbar_init(BAR(B_HSTAGED), N_EPI_WARPS);
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
  if (p.stagger_cycles > 0 && (blockIdx.x & 1)) {
    const long long t0 = clock64();
    while (clock64() - t0 < p.stagger_cycles) __nanosleep(200);
  }

  if (warp == 0) {
    // ======================= producer: weight blocks (bulk copies), O tiles (TMA) =======================
    if (lane == 0) {
      int wc = 0;                                  // weight blocks issued so far
      auto load_w = [&](int b) {
        const int s = wc % NWK;
        mbar_wait(BAR(B_WEMPTY + s), ((wc / NWK) & 1) ^ 1, 40);
        mbar_arrive_expect_tx(BAR(B_WFULL + s), BLK_BYTES);
        bulk_load(smem_base + OFF_W + wslot_blk(s) * BLK_BYTES, p.wpack + (size_t)b * BLK_BYTES, BLK_BYTES, BAR(B_WFULL + s));
        ++wc;
      };
      int n = 0;
      auto load_h = [&](int tile, int nn) {
        if (nn > 0) mbar_wait(BAR(B_HFREE), (nn - 1) & 1, 57);     // epilogue 1 of the previous tile has read the buffer
        mbar_arrive_expect_tx(BAR(B_HFULL), H_BYTES);
        const float* src = HEAD ? p.pe_tiled + (size_t)(tile % p.pe_tiles) * 16384 : p.h + (size_t)tile * 16384;
        bulk_load(smem_base + OFF_H, src, H_BYTES, BAR(B_HFULL));
      };
      auto load_o = [&](int tile, int nn) {
        if (nn > 0) mbar_wait(BAR(B_OEMPTY), (nn - 1) & 1, 41);    // out-projection MMAs of the previous tile retired
        mbar_arrive_expect_tx(BAR(B_OFULL), BLK_BYTES);
        tma_load_2d(smem_base + OFF_O, &tm_o, BAR(B_OFULL), 0, tile * 128);
        tma_load_2d(smem_base + OFF_O + HALF_BYTES, &tm_o, BAR(B_OFULL), HEAD ? p.x_col1 : 64, tile * 128);
      };
      // the weights do not depend on the previous kernel: the first ring fill overlaps its tail (PDL)
      int pre = 0;
      if ((int)blockIdx.x < n_tiles)
        for (; pre < NWK && pre < n_wblk; ++pre) load_w(pre);
      pdl_wait();
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
        if (n == 0) {
          load_o(tile, 0);
          load_h(tile, 0);
        }
        const int next = tile + (int)gridDim.x;
        bool o_done = next >= n_tiles;
        for (int b = (n == 0 ? pre : 0); b < n_wblk; ++b) {
          load_w(b);
          TT(200 + b);
          // the O buffer and the residual buffer were released early in this tile: next tile's loads
          if (!o_done && b >= (HEAD ? 2 : 5)) { load_o(next, n + 1); if (!HAS_QKV) load_h(next, n + 1); TT(220); o_done = true; }
        }
        if (!o_done) { load_o(next, n + 1); if (!HAS_QKV) load_h(next, n + 1); }
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
      TT(0);
      mbar_wait(BAR(B_OFULL), n & 1, 42);
      TT(1);
      if (n > 0 && HAS_QKV) mbar_wait(BAR(B_REGFREE + n % 3), prev_par, 43);           // q of the previous tile drained
      {
        const int ws = wc % NWK;
        TT(2);
        mbar_wait(BAR(B_WFULL + ws), (wc / NWK) & 1, 44);
        TT(3);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = w_lo0 + (uint32_t)wslot_blk(ws) * (BLK_BYTES >> 4);
          if (HEAD && p.x_tf32) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss_lh_tf32(r_acc, o_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                              w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2, DESC_HI_SW128, idesc_tf32(128, 128), kk != 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss_lh(r_acc, o_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                         w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2, DESC_HI_SW128, IDESC, kk != 0 ? 1u : 0u);
          }
          umma_commit(BAR(B_OEMPTY));
          umma_commit(BAR(B_WEMPTY + ws));
          umma_commit(BAR(B_ACC1));
        }
        __syncwarp();
        TT(4);
        ++wc;
      }
      // ---- feed-forward: LN2 operand in TMEM (written by the epilogue warps), hidden blocks through TMEM
      if (!HEAD) {
        mbar_wait(BAR(B_ALN), aln_use & 1, 45);
        TT(5);
        ++aln_use;
      }
      for (int q = 0; q < (HEAD ? 0 : 8); ++q, ++wc) {
        int is_w2, nb;
        ffn_seq(q, is_w2, nb);
        const int ws = wc % NWK, hs = nb & 1;
        mbar_wait(BAR(B_WFULL + ws), (wc / NWK) & 1, 46);
        TT(10 + 4 * q);
        if (!is_w2 && n > 0) {
          // the hidden slots are the previous tile's K / V (or classifier) accumulators: drained?
          if (nb == 0 && HAS_QKV) mbar_wait(BAR(B_REGFREE + (n + 1) % 3), prev_par, 47);
          if (nb == 1) mbar_wait(BAR(B_REGFREE + (n + 2) % 3), prev_par, 48);
        }
        if (is_w2) mbar_wait(BAR(B_HIDBF + hs), (hid_uses[hs] - 1) & 1, 49);
        TT(11 + 4 * q);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = w_lo0 + (uint32_t)wslot_blk(ws) * (BLK_BYTES >> 4);
          if (!is_w2) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lh(r_hid[hs], t_aln + kk * 8, w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                         DESC_HI_SW128, IDESC, kk != 0 ? 1u : 0u);
            umma_commit(BAR(B_HIDFULL + hs));
          } else {
            // acc already holds h' + b2 (written by the epilogue warps): always accumulate.  The bf16 hidden
            // values of k = 32 c .. 32 c + 31 sit in the first 16 of the 32 columns their epilogue thread owns.
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lh(r_acc, r_hid[hs] + 32 * (kk >> 1) + (kk & 1) * 8,
                         w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2, DESC_HI_SW128, IDESC, 1u);
            if (nb == 3) umma_commit(BAR(B_OUTFULL));
          }
          umma_commit(BAR(B_WEMPTY + ws));
        }
        __syncwarp();
        TT(12 + 4 * q);
        if (!is_w2) hid_uses[hs]++;
      }
      // ---- next layer's q, k, v = LN1_next(h'') Wqkv^T: accumulators in the two hidden slots and the old acc
      if (HAS_QKV) {
        mbar_wait(BAR(B_ALN), aln_use & 1, 50);
        TT(50);
        ++aln_use;
        for (int j = 0; j < 3; ++j, ++wc) {
          const int ws = wc % NWK;
          const uint32_t r_out = j == 0 ? r_hid[0] : j == 1 ? r_hid[1] : r_acc;
          mbar_wait(BAR(B_WFULL + ws), (wc / NWK) & 1, 51);
          // head mode has no hidden blocks in between: q / k land on the previous tile's k / v accumulators
          if (HEAD && n > 0 && j < 2) mbar_wait(BAR(B_REGFREE + (n + 1 + j) % 3), prev_par, 59);
          TT(52 + 2 * j);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t w_lo = w_lo0 + (uint32_t)wslot_blk(ws) * (BLK_BYTES >> 4);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lh(r_out, t_aln + kk * 8, w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                         DESC_HI_SW128, IDESC, kk != 0 ? 1u : 0u);
            umma_commit(BAR(B_QKVFULL + j));
            umma_commit(BAR(B_WEMPTY + ws));
          }
          __syncwarp();
          TT(53 + 2 * j);
        }
      }
    }
  } else if (warp == 2 + N_EPI_WARPS) {
    // ======================= residual-tile mover (layers that write h back) =======================
    // The epilogue warps stage h'' in the shared-memory residual tile; this thread sends it to HBM with ONE 64 KB
    // bulk store and, once the store has read the buffer, fetches the next tile's rows into it -- no compute
    // warp ever waits for either copy.
    if (HAS_QKV && lane == 0) {
      pdl_wait();
      int n = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
        const int next = tile + (int)gridDim.x;
        mbar_wait(BAR(B_HSTAGED), n & 1, 58);
        bulk_store(p.h + (size_t)tile * 16384, smem_base + OFF_H, H_BYTES);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (next < n_tiles) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive_expect_tx(BAR(B_HFULL), H_BYTES);
          const float* src = HEAD ? p.pe_tiled + (size_t)(next % p.pe_tiles) * 16384 : p.h + (size_t)next * 16384;
          bulk_load(smem_base + OFF_H, src, H_BYTES, BAR(B_HFULL));
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // ======================= epilogue warps: four per TMEM lane quarter, 32 columns each =======================
    pdl_wait();
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int csel = (warp - 2) >> 2;              // which 32 of the 128 columns
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t stg_off0 = OFF_STG + (uint32_t)(warp - 2) * NSTG * STG_BYTES;
    int unit = 0;
    const uint32_t t_aln = tmem_base + lane_addr + TM_ALN + 16u * csel;
    const uint32_t t_x = tmem_base + lane_addr + TM_X;
    // this thread's 32 columns of row `row` of the fp32 residual tile in shared memory: 8 float4 at stride 2 KB,
    // a warp's 32 rows of one column quad are 512 contiguous bytes (tiled layout == the HBM layout)
    float4* hs_ptr = reinterpret_cast<float4*>(smem_gen + OFF_H) + (size_t)(csel * 8) * 128 + row;
    int n = 0, ln_count = 0;
    int hid_uses[2] = {0, 0};
#define TE(slot) do { if (warp == 2) TT(slot); } while (0)

    // Row sums across the four threads that share a row (same lane, the four warps of the quarter): each stores
    // its partials into its 4 of the 16 exchange columns of the row's TMEM lane, a 128-thread named barrier, and
    // everyone reads all 16.  Two buffers alternate, so a fast thread's next store cannot overtake a slow
    // thread's read.
    auto exchange4 = [&](float a, float b, float c, float d, float (&tot)[4]) {
      const uint32_t tx = t_x + 16u * (uint32_t)(ln_count & 1);
      ++ln_count;
      tmem_st4v(tx + 4u * csel, __float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
      tmem_st_wait();
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(NB_QUARTER + q) : "memory");
      tc_fence_after();
      uint32_t e[16];
      tmem_ld16(tx, e);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tot[k] = (__uint_as_float(e[k]) + __uint_as_float(e[4 + k])) + (__uint_as_float(e[8 + k]) + __uint_as_float(e[12 + k]));
    };
    // mean / rstd of the row from single-pass sums (fp32; |mean| of a residual row is of the order of its
    // standard deviation, so E[x^2] - mean^2 loses a bit or two, far below the bf16 rounding of the result)
    auto row_norm = [&](const uint32_t (&u)[32], float& rstd, float& nmr) {
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = __uint_as_float(u[i]);
        s1[i & 3] += x;
        s2[i & 3] = fmaf(x, x, s2[i & 3]);
      }
      float tot[4];
      exchange4((s1[0] + s1[1]) + (s1[2] + s1[3]), (s2[0] + s2[1]) + (s2[2] + s2[3]), 0.f, 0.f, tot);
      const float mean = tot[0] * (1.0f / 128.0f);
      const float var = fmaxf(fmaf(-mean, mean, tot[1] * (1.0f / 128.0f)), 0.f);
      rstd = rsqrtf(var + LN_EPS);
      nmr = -mean * rstd;
    };
    // (x - mean) * rstd -> bf16 pairs -> this thread's 16 columns of the TMEM LayerNorm operand (gamma / beta are
    // folded into the weights and biases of the GEMM that consumes it)
    auto emit_norm = [&](const uint32_t (&u)[32], float rstd, float nmr) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_bf16_alu(fmaf(__uint_as_float(u[2 * i]), rstd, nmr), fmaf(__uint_as_float(u[2 * i + 1]), rstd, nmr));
      tmem_st16(t_aln, pk);
    };

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const uint32_t r_acc = tmem_base + lane_addr + 128u * (uint32_t)(n % 3) + 32u * csel;
      const uint32_t r_hid[2] = {tmem_base + lane_addr + 128u * (uint32_t)((n + 1) % 3) + 32u * csel,
                                 tmem_base + lane_addr + 128u * (uint32_t)((n + 2) % 3) + 32u * csel};
      uint32_t u[32];
      // ---- epilogue 1: h' = acc + bo + h;  acc <- h' + b2;  normalised h' -> TMEM operand
      TE(100);
      mbar_wait(BAR(B_HFULL), n & 1, 56);
      mbar_wait(BAR(B_ACC1), n & 1, 52);
      TE(101);
      tc_fence_after();
      tmem_ld32(r_acc, u);
      tmem_ld_wait();
      TE(102);
      {
        const float4* bp = reinterpret_cast<const float4*>(p.bo + csel * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = __ldg(bp + i);
          const float4 h4 = hs_ptr[i * 128];
          u[4 * i] = __float_as_uint(__uint_as_float(u[4 * i]) + h4.x + b4.x);
          u[4 * i + 1] = __float_as_uint(__uint_as_float(u[4 * i + 1]) + h4.y + b4.y);
          u[4 * i + 2] = __float_as_uint(__uint_as_float(u[4 * i + 2]) + h4.z + b4.z);
          u[4 * i + 3] = __float_as_uint(__uint_as_float(u[4 * i + 3]) + h4.w + b4.w);
        }
      }
      if (HEAD) {
        // h0 = x W_in^T + b_in + PE/sqrt(d) (bo = b_in, residual tile = positional-encoding tile): staged in place for the
        // mover, normalised (LN1 of layer 0, affine part folded into Wqkv') for the q/k/v GEMMs
#pragma unroll
        for (int i = 0; i < 8; ++i)
          hs_ptr[i * 128] = make_float4(__uint_as_float(u[4 * i]), __uint_as_float(u[4 * i + 1]), __uint_as_float(u[4 * i + 2]),
                                        __uint_as_float(u[4 * i + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HSTAGED));
        float rstd, nmr;
        row_norm(u, rstd, nmr);
        emit_norm(u, rstd, nmr);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ALN));
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HFREE));      // this warp's part of the residual tile is in registers
        {
          float rstd, nmr;
          row_norm(u, rstd, nmr);
          TE(103);
          emit_norm(u, rstd, nmr);
          const float4* bp = reinterpret_cast<const float4*>(p.b2 + csel * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bp + i);
            u[4 * i] = __float_as_uint(__uint_as_float(u[4 * i]) + b4.x);
            u[4 * i + 1] = __float_as_uint(__uint_as_float(u[4 * i + 1]) + b4.y);
            u[4 * i + 2] = __float_as_uint(__uint_as_float(u[4 * i + 2]) + b4.z);
            u[4 * i + 3] = __float_as_uint(__uint_as_float(u[4 * i + 3]) + b4.w);
          }
          tmem_st32(r_acc, u);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ALN));
          TE(104);
        }
        // ---- hidden blocks: t = acc + b1';  2 ReLU(t) = t + |t|  -> bf16 pairs over the head of this thread's own columns
        for (int nb = 0; nb < 4; ++nb) {
          const int hs = nb & 1;
          TE(110 + 4 * nb);
          mbar_wait(BAR(B_HIDFULL + hs), hid_uses[hs] & 1, 53);
          TE(111 + 4 * nb);
          hid_uses[hs]++;
          tc_fence_after();
          tmem_ld32(r_hid[hs], u);
          tmem_ld_wait();
          TE(112 + 4 * nb);
          uint32_t pk[16];
          const float4* bp = reinterpret_cast<const float4*>(p.b1p + nb * 128 + csel * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bp + i);
            const float t0 = __uint_as_float(u[4 * i]) + b4.x, t1 = __uint_as_float(u[4 * i + 1]) + b4.y;
            const float t2 = __uint_as_float(u[4 * i + 2]) + b4.z, t3 = __uint_as_float(u[4 * i + 3]) + b4.w;
            pk[2 * i] = pack_bf16_alu(t0 + fabsf(t0), t1 + fabsf(t1));
            pk[2 * i + 1] = pack_bf16_alu(t2 + fabsf(t2), t3 + fabsf(t3));
          }
          tmem_st16(r_hid[hs], pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_HIDBF + hs));
          TE(113 + 4 * nb);
        }
      }
      // ---- final epilogue: acc = h'' (residual and b2 were in the accumulator from the start)
      if (!HEAD) {
        TE(130);
        mbar_wait(BAR(B_OUTFULL), n & 1, 54);
        TE(131);
        tc_fence_after();
        tmem_ld32(r_acc, u);
        tmem_ld_wait();
        TE(132);
      }
      if (HAS_QKV) {
        if (!HEAD) {
        // h'' -> the shared-memory residual tile (its old contents were consumed by epilogue 1 of every warp: all of
        // them have arrived on B_ALN since, which the MMAs behind B_OUTFULL waited for); the mover warp stores it
#pragma unroll
        for (int i = 0; i < 8; ++i)
          hs_ptr[i * 128] = make_float4(__uint_as_float(u[4 * i]), __uint_as_float(u[4 * i + 1]), __uint_as_float(u[4 * i + 2]),
                                        __uint_as_float(u[4 * i + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HSTAGED));
        TE(133);
        float rstd, nmr;
        row_norm(u, rstd, nmr);
        emit_norm(u, rstd, nmr);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ALN));
        TE(134);
        TE(135);
        }
        // ---- q, k, v: + bias' -> bf16 -> 64B-swizzled slab -> TMA store (rows past M clipped by the tensor map)
        for (int j = 0; j < 3; ++j) {
          const uint32_t r_out = j == 0 ? r_hid[0] : j == 1 ? r_hid[1] : r_acc;
          const int region = j == 0 ? (n + 1) % 3 : j == 1 ? (n + 2) % 3 : n % 3;
          TE(140 + 3 * j);
          mbar_wait(BAR(B_QKVFULL + j), n & 1, 55);
          TE(141 + 3 * j);
          tc_fence_after();
          tmem_ld32(r_out, u);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          const uint32_t stg_off = stg_off0 + (uint32_t)(unit % NSTG) * STG_BYTES;
          ++unit;
          if (lane == 0) {
            mbar_arrive(BAR(B_REGFREE + region));
            // the slab used NSTG stores ago is free again (the TMA engine also carries the 64 KB residual tile)
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSTG - 1) : "memory");
          }
          __syncwarp();
          const float4* bp = reinterpret_cast<const float4*>(p.bqkvp + j * 128 + csel * 32);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 ba = __ldg(bp + 2 * c), bb = __ldg(bp + 2 * c + 1);
            const uint32_t* f = &u[c * 8];
            // 64-byte swizzle: 16-byte chunk index XOR bits 1-2 of the row
            *reinterpret_cast<uint4*>(smem_gen + stg_off + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(pack_bf16_alu(__uint_as_float(f[0]) + ba.x, __uint_as_float(f[1]) + ba.y),
                           pack_bf16_alu(__uint_as_float(f[2]) + ba.z, __uint_as_float(f[3]) + ba.w),
                           pack_bf16_alu(__uint_as_float(f[4]) + bb.x, __uint_as_float(f[5]) + bb.y),
                           pack_bf16_alu(__uint_as_float(f[6]) + bb.z, __uint_as_float(f[7]) + bb.w));
          }
          fence_proxy_async_smem();
          __syncwarp();
          // (plain coalesced stores from the slab instead of a TMA store were measured slower: +10 us per launch --
          // at this point of a tile every SM writes at once and st.global back-pressure stalls the warps)
          if (lane == 0) {
            tma_store_2d(j == 0 ? &tm_q : j == 1 ? &tm_k : &tm_v, smem_base + stg_off, csel * 32, tile * 128 + q * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          TE(142 + 3 * j);
        }
      } else {
        // Last layer: the encoder's final LayerNorm (transformer.py:33), the classifier and the log-softmax
        // (self_attention.py:26-27) and the caller's softmax(...)[...,1] = sigmoid(z1 - z0)
        // (predictor.py:225,257-258) on the row that is still in registers: h is never written.  With
        // gw_c = gamma_f * Wc[c]:  z_c = rstd * (sum x gw_c - mean * sum gw_c) + (beta_f . Wc[c] + bc[c]).
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_REGFREE + n % 3));
        float s1 = 0.f, s2 = 0.f, d0 = 0.f, d1 = 0.f;
        const float4* g0 = reinterpret_cast<const float4*>(p.cls_gw + csel * 32);
        const float4* g1 = reinterpret_cast<const float4*>(p.cls_gw + 128 + csel * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a4 = __ldg(g0 + i), c4 = __ldg(g1 + i);
          const float x0 = __uint_as_float(u[4 * i]), x1 = __uint_as_float(u[4 * i + 1]);
          const float x2 = __uint_as_float(u[4 * i + 2]), x3 = __uint_as_float(u[4 * i + 3]);
          s1 += (x0 + x1) + (x2 + x3);
          s2 = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, s2))));
          d0 = fmaf(x0, a4.x, fmaf(x1, a4.y, fmaf(x2, a4.z, fmaf(x3, a4.w, d0))));
          d1 = fmaf(x0, c4.x, fmaf(x1, c4.y, fmaf(x2, c4.z, fmaf(x3, c4.w, d1))));
        }
        float tot[4];
        exchange4(s1, s2, d0, d1, tot);
        const long grow = (long)tile * 128 + row;
        if (csel == 0 && grow < p.M) {
          const float mean = tot[0] * (1.0f / 128.0f);
          const float var = fmaxf(fmaf(-mean, mean, tot[1] * (1.0f / 128.0f)), 0.f);
          const float rstd = 1.0f / sqrtf(var + LN_EPS);
          const float a0 = fmaf(rstd, fmaf(-mean, __ldg(p.cls_gw + 256), tot[2]), __ldg(p.cls_gw + 258));
          const float a1 = fmaf(rstd, fmaf(-mean, __ldg(p.cls_gw + 257), tot[3]), __ldg(p.cls_gw + 259));
          const float mx = fmaxf(a0, a1);
          const float lse = mx + log1pf(expf(-fabsf(a1 - a0)));        // log_softmax([a0, a1]), stable
          if (p.logp) {
            if (p.logp_vec) *reinterpret_cast<float2*>(p.logp + grow * 2) = make_float2(a0 - lse, a1 - lse);
            else { p.logp[grow * 2] = a0 - lse; p.logp[grow * 2 + 1] = a1 - lse; }
          }
          if (p.prob) p.prob[grow] = 1.0f / (1.0f + expf(a0 - a1));     // softmax(logp)[1]
        }
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

// fp32 [N_total, K_total] row-major weight -> one 32 KB block: rows n0..n0+127, columns k0..k0+127, scaled by
// `scale` and (optionally) per column by col_scale[k] (a folded LayerNorm gamma), as bf16 in the UMMA K-major
// 128B-swizzled shared-memory image (two 16 KB halves of 64 k each)
__global__ void pack_block_kernel(const float* __restrict__ w, int ld, int n0, int k0, const float* __restrict__ col_scale,
                                  float scale, int k_valid, unsigned char* __restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 8 consecutive k of one row
  if (t >= 128 * 16) return;
  const int row = t >> 4, k8 = t & 15;
  const float* src = w + (size_t)(n0 + row) * ld + k0 + k8 * 8;
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)      // columns >= k_valid: zero (front-end weight [128, F] padded to K = 128)
    x[j] = (k0 + k8 * 8 + j < k_valid) ? src[j] * scale * (col_scale ? col_scale[k0 + k8 * 8 + j] : 1.0f) : 0.f;
  uint4 o;
  o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]);
  o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
  const int half = k8 >> 3, chunk = k8 & 7;
  *reinterpret_cast<uint4*>(dst + half * HALF_BYTES + sw128_offset(row, chunk)) = o;
}

// fp32 [128, F <= 64] front-end weight -> one 32 KB block of fp32 words for kind::tf32: two halves of [128 rows x 32 k]
// (128-byte rows, 128B swizzle), columns >= F zero
__global__ void pack_tf32_kernel(const float* __restrict__ w, int F, unsigned char* __restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 consecutive k of one row
  if (t >= 128 * 16) return;
  const int row = t >> 4, k4 = t & 15;
  float4 v;
  v.x = k4 * 4 + 0 < F ? w[(size_t)row * F + k4 * 4 + 0] : 0.f; v.y = k4 * 4 + 1 < F ? w[(size_t)row * F + k4 * 4 + 1] : 0.f;
  v.z = k4 * 4 + 2 < F ? w[(size_t)row * F + k4 * 4 + 2] : 0.f; v.w = k4 * 4 + 3 < F ? w[(size_t)row * F + k4 * 4 + 3] : 0.f;
  *reinterpret_cast<float4*>(dst + (k4 >> 3) * HALF_BYTES + sw128_offset(row, k4 & 7)) = v;
}

// out[n] = bias[n] + sum_k W[n][k] * beta[k]   (LayerNorm beta folded into the bias of the following Linear)
__global__ void fold_bias_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ beta,
                                 int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < D; ++k) acc = fmaf(w[(size_t)n * D + k], beta[k], acc);
  out[n] = bias[n] + acc;
}

// last layer: gw[c][k] = gamma_f[k] * Wc[c][k];  consts = {sum gw0, sum gw1, beta_f.Wc0 + bc0, beta_f.Wc1 + bc1}
__global__ void fold_cls_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ wc,
                                const float* __restrict__ bc, float* __restrict__ out) {
  if (threadIdx.x < 2) {
    const int c = threadIdx.x;
    float sg = 0.f, sb = 0.f;
    for (int k = 0; k < D; ++k) {
      const float gw = g[k] * wc[c * D + k];
      out[c * D + k] = gw;
      sg += gw;
      sb = fmaf(b[k], wc[c * D + k], sb);
    }
    out[256 + c] = sg;
    out[258 + c] = sb + bc[c];
  }
}

CUresult tmap2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esz, long rows, long cols,
                int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return encode_tiled(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, swz);
}

}  // namespace

size_t tail_pack_bytes() { return (size_t)12 * BLK_BYTES; }
size_t tail_aux_floats() { return 512 + 384 + 260; }     // b1' | bqkv' | classifier fold

cudaError_t launch_tail_pack(const TailPackArgs& a, unsigned char* dst, float* aux, cudaStream_t s) {
  int b = 0;
  auto blk = [&](const float* w, int ld, int n0, int k0, const float* cs, float scale) {
    pack_block_kernel<<<8, 256, 0, s>>>(w, ld, n0, k0, cs, scale, 1 << 30, dst + (size_t)b * BLK_BYTES);
    ++b;
  };
  blk(a.wo, D, 0, 0, nullptr, 1.0f);
  // W1'[0] W1'[1] W2[0]/2 W1'[2] W2[1]/2 W1'[3] W2[2]/2 W2[3]/2  (ffn_seq); W1' = W1 diag(gamma2)
  blk(a.w1, D, 0, 0, a.ln2_g, 1.0f); blk(a.w1, D, 128, 0, a.ln2_g, 1.0f); blk(a.w2, DFF, 0, 0, nullptr, 0.5f);
  blk(a.w1, D, 256, 0, a.ln2_g, 1.0f); blk(a.w2, DFF, 0, 128, nullptr, 0.5f); blk(a.w1, D, 384, 0, a.ln2_g, 1.0f);
  blk(a.w2, DFF, 0, 256, nullptr, 0.5f); blk(a.w2, DFF, 0, 384, nullptr, 0.5f);
  fold_bias_kernel<<<2, 256, 0, s>>>(a.w1, a.b1, a.ln2_b, DFF, aux);
  if (a.wqkv_next) {
    for (int j = 0; j < 3; ++j) blk(a.wqkv_next, D, j * 128, 0, a.ln1n_g, 1.0f);
    fold_bias_kernel<<<2, 256, 0, s>>>(a.wqkv_next, a.bqkv_next, a.ln1n_b, 3 * D, aux + 512);
  } else {
    fold_cls_kernel<<<1, 32, 0, s>>>(a.lnf_g, a.lnf_b, a.wc, a.bc, aux + 512 + 384);
  }
  return cudaGetLastError();
}

// head of the model: [W_in | Wq' | Wk' | Wv'] for bf16 features (dst_bf16) and, when F <= 64, for fp32 features
// multiplied as tf32 (dst_tf32: W_in as fp32 words); aux384 = q|k|v biases of layer 0 + Wqkv beta1
cudaError_t launch_head_pack(const float* w_in, int F, const float* wqkv0, const float* bqkv0, const float* ln1_g,
                             const float* ln1_b, unsigned char* dst_bf16, unsigned char* dst_tf32, float* aux384,
                             cudaStream_t s) {
  if (F <= 128) pack_block_kernel<<<8, 256, 0, s>>>(w_in, F, 0, 0, nullptr, 1.0f, F, dst_bf16);
  if (F <= 64) pack_tf32_kernel<<<8, 256, 0, s>>>(w_in, F, dst_tf32);
  for (int j = 0; j < 3; ++j) {
    pack_block_kernel<<<8, 256, 0, s>>>(wqkv0, D, j * 128, 0, ln1_g, 1.0f, 1 << 30, dst_bf16 + (size_t)(1 + j) * BLK_BYTES);
    pack_block_kernel<<<8, 256, 0, s>>>(wqkv0, D, j * 128, 0, ln1_g, 1.0f, 1 << 30, dst_tf32 + (size_t)(1 + j) * BLK_BYTES);
  }
  fold_bias_kernel<<<2, 256, 0, s>>>(wqkv0, bqkv0, ln1_b, 3 * D, aux384);
  return cudaGetLastError();
}

cudaError_t launch_tail_tc(const TailTcArgs& a, int num_sms, cudaStream_t s, std::string* err) {
  if (a.M <= 0) return cudaSuccess;
  const bool has_qkv = a.q != nullptr;
  const bool head = a.x != nullptr;
  if (head && (!has_qkv || !a.pe_tiled || a.pe_tiles <= 0)) return cudaErrorInvalidValue;
  CUtensorMap to, tq, tk, tv;
  CUresult r;
  if (!head) r = tmap2d(&to, a.o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  else if (a.x_is_bf16) r = tmap2d(&to, a.x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, a.F, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  else r = tmap2d(&to, a.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, a.F, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (has_qkv) {
    if (r == CUDA_SUCCESS) r = tmap2d(&tq, a.q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (r == CUDA_SUCCESS) r = tmap2d(&tk, a.k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (r == CUDA_SUCCESS) r = tmap2d(&tv, a.v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  } else {
    tq = to; tk = to; tv = to;
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return cudaErrorInvalidValue;
  }
  TailParams p = {};
  p.M = a.M; p.wpack = a.wpack; p.h = a.h;
  p.bo = a.bo; p.b2 = a.b2; p.b1p = a.aux; p.bqkvp = a.aux + 512; p.cls_gw = a.aux + 512 + 384;
  p.prob = a.prob; p.logp = a.logp; p.q = a.q; p.k = a.k; p.v = a.v;
  if (head) {
    p.bqkvp = a.aux;                      // head aux: only the folded q|k|v biases
    p.pe_tiled = a.pe_tiled; p.pe_tiles = a.pe_tiles; p.x_col1 = a.x_is_bf16 ? 64 : 32; p.x_tf32 = a.x_is_bf16 ? 0 : 1;
  }
  p.logp_vec = (reinterpret_cast<uintptr_t>(a.logp) % 8) == 0;
  if (!a.aux || !a.bo || (!head && !a.b2)) return cudaErrorInvalidValue;
  static const int stagger = getenv("VADB_TAIL_STAGGER") ? atoi(getenv("VADB_TAIL_STAGGER")) : 0;
  p.stagger_cycles = stagger;
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(tail_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
  }
  const int n_tiles = (a.M + 127) / 128;
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  static const char* want_trace = getenv("VADB_TAIL_TRACE");          // "1": q|k|v variant, "0": last-layer variant
  if (want_trace && !head && (atoi(want_trace) != 0) == has_qkv) {
    long long* dtr = nullptr;
    cudaMalloc(&dtr, 256 * sizeof(long long));
    cudaMemsetAsync(dtr, 0, 256 * sizeof(long long), s);
    p.trace = dtr;
    if (has_qkv) tail_tc_kernel<1, true><<<grid, NTHREADS, SMEM_BYTES, s>>>(to, tq, tk, tv, p);
    else tail_tc_kernel<0, true><<<grid, NTHREADS, SMEM_BYTES, s>>>(to, tq, tk, tv, p);
    long long ht[256];
    cudaMemcpyAsync(ht, dtr, sizeof ht, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(dtr);
    const long long t0 = ht[0];
    fprintf(stderr, "[tail trace] CTA 0, third tile, cycles relative to the MMA warp's tile start\n");
    for (int i = 0; i < 256; ++i)
      if (ht[i]) fprintf(stderr, "[tail trace] %3d %8lld\n", i, ht[i] - t0);
    return cudaGetLastError();
  }
  {
    cudaError_t e = head ? launch_k(tail_tc_kernel<2, false>, grid, NTHREADS, SMEM_BYTES, s, to, tq, tk, tv, p)
                  : has_qkv ? launch_k(tail_tc_kernel<1, false>, grid, NTHREADS, SMEM_BYTES, s, to, tq, tk, tv, p)
                            : launch_k(tail_tc_kernel<0, false>, grid, NTHREADS, SMEM_BYTES, s, to, tq, tk, tv, p);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace vadb
