// Per-frame Linear layers on the 5th-gen tensor cores (sm_100a): persistent, warp-specialised
// tcgen05 GEMM with fused prologue (LayerNorm) and epilogue (bias, ReLU, residual, bf16/fp32).
//
//   C[m, n] = epi( pro(A)[m, :] . W[n, :] + bias[n] ),  A [M,K], W [N,K] (nn.Linear layout), bf16 operands,
//   fp32 accumulation in TMEM.
//
// Serves (bf16 compute mode) vad/modeling/transformer.py:281-284 (Q/K/V projections behind the
// pre-LayerNorm of :235-236), :347 (final projection + residual :237), :370-375 (feed-forward,
// ReLU, residual).  These layers are HBM-bound (intensity <= 100 FLOP/B): the design goal is to
// stream activations once, in full 128-byte lines, while the weights stay resident in shared memory.
//
// CTA (persistent over 128-row tiles, weights loaded once):
//   warp 0      TMA producer: W (once), A k-chunks when A is bf16 in global memory
//   warp 1      MMA issuer (one elected thread), TMEM allocator
//   warps 2-5   A producers (A = fp32 residual stream through LayerNorm, or raw feature rows for the
//               front end): coalesced loads prefetched one batch ahead, warp-shuffle row statistics,
//               bf16 into the 128B-swizzled UMMA operand layout
//   warps 6-13  epilogue: tcgen05.ld accumulators -> bias/ReLU/residual/positional encoding -> per-warp
//               swizzled staging slab -> per-warp TMA store (full-line writes, no cross-warp barrier)
// Accumulators: four 128-column TMEM slots used as a ring over (tile, n-block) jobs, so the MMAs of
// the next job overlap the epilogue of the previous one.
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"
#include "vadb_common.cuh"

namespace vadb {
namespace {

using namespace tc;

constexpr int NTHREADS = 448;          // TMA + MMA + 4 producer warps + 8 epilogue warps
constexpr int N_EPI_WARPS = 8;
constexpr uint32_t BLK_BYTES = 128 * 128 * 2;     // [128 x 128] bf16 block = two SW128 halves of 16 KB
constexpr uint32_t HALF_BYTES = 128 * 128;        // [128 rows x 64 k] bf16
constexpr uint32_t STG_BYTES = 32 * 128;          // per-warp staging: 32 rows x 128 B
constexpr uint32_t IDESC = idesc_bf16(128, 128, 0, 0);
constexpr uint32_t IDESC_TF32 = idesc_tf32(128, 128);

enum { BAR_WFULL = 0, BAR_AFULL = 1, BAR_AEMPTY = 4, BAR_ACCFULL = 7, BAR_ACCEMPTY = 11, BAR_RFULL = 15, BAR_REMPTY = 16,
       BAR_COUNT = 17 };
constexpr uint32_t RES_BYTES = 128 * 128 * 4;     // one fp32 residual tile = four SW128 boxes of 32 columns

struct GemmTcParams {
  int M, N, K;              // N multiple of 128, <= 512; K (padded) multiple of 128; N*K <= 65536
  int n_a_stages;           // 1..3
  int n_stg;                // staging slabs per epilogue warp (1 or 2)
  int prod;                 // 0: A is bf16 in global memory (TMA); 1: A = LayerNorm(fp32 rows, K == 128);
                            // 2: A = fp32/bf16 rows of a_cols <= 128 columns, converted and zero-padded
  const void* a_src;        // producer modes: source rows
  int a_cols;               // producer mode 2: valid columns (multiple of 4)
  int a_is_bf16;            // producer mode 2: source element type
  int win_W, win_half, win_jump;   // producer mode 2: window gather (vad/predictor.py:182-218) when win_W > 0
  const float* ln_g;
  const float* ln_b;
  const float* bias;        // [N]
  int relu;
  const float* residual;    // fp32 [*, 128] added in the epilogue (N == 128) or nullptr
  int res_mod;              // > 0: residual row = output row % res_mod (positional-encoding table)
  int res_tma;              // residual tiles arrive by TMA (tm_o2 doubles as their load map) one tile ahead of
                            // the epilogue instead of per-thread row loads whose DRAM latency nothing hides
  int tf32;                 // A and W are fp32 in global memory, loaded by TMA and multiplied as tf32 (front end):
                            // K counts 64-float stages (K/128 of them), element coordinates are halved
  int out_f32;              // 1: fp32 output [M, N]; 0: bf16
  int out_split;            // bf16 outputs: columns [j*128, j*128+128) -> out map j when split (q,k,v)
  int out_tiled;            // fp32 N == 128 output in the tiled residual layout: direct coalesced stores from registers
  // fp32 output with N == 128 only: additionally emit LayerNorm(out_row) in bf16 through tm_o1 -- the A
  // operand of the NEXT kernel (pre-LN of the following sublayer, transformer.py:235-236), so that kernel
  // is fed by TMA instead of register-staged producer warps
  const float* emit_g;
  const float* emit_b;
  void* out_ptr[3];         // raw output pointers (epilogue stores are coalesced st.global from a staged slab)
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ long window_src_row(long m, int W, int half, int jump) {
  // vad/predictor.py:186-199: rel = [-half..0) step jump, 0, [1..half] step jump
  const long i = m / W;
  const int k = (int)(m - i * W);
  const int nl = (half + jump - 1) / jump;
  const int rel = (k < nl) ? (-half + k * jump) : (k == nl ? 0 : 1 + (k - nl - 1) * jump);
  return half + i + rel;
}

// TRACE: developer instrumentation (VADB_GEMM_TRACE=1): CTA 0 stamps clock64() per role for its first tiles
#define GTR(slot) do { if (TRACE && blockIdx.x == 0 && trace && (slot) < 1024) trace[(slot)] = clock64(); } while (0)

// PROD = false: A arrives by TMA and the four producer warps do not exist (10 warps -> up to 168
// registers per thread, no spills: with 224 KB of the SM given to shared memory the L1 that would absorb
// spills is only a few KB); PROD = true: 14 warps, 128 registers.
template <bool TRACE, bool PROD>
__global__ void __launch_bounds__(PROD ? NTHREADS : NTHREADS - 128, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_a,
               const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1,
               const __grid_constant__ CUtensorMap tm_o2, const GemmTcParams p, long long* trace) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  unsigned char* smem_gen = smem_raw;
  if ((smem_base & 1023u) != 0) {   // 128B-swizzled tiles need 1024-byte alignment
    if (threadIdx.x == 0) printf("vadb: gemm kernel shared memory window not 1024-byte aligned\n");
    __trap();
  }

  const int NB = p.N >> 7, KC = p.K >> 7;
  const uint32_t off_w = 0;
  const uint32_t off_a = off_w + (uint32_t)(NB * KC) * BLK_BYTES;
  const uint32_t off_r = off_a + (uint32_t)p.n_a_stages * BLK_BYTES;
  const uint32_t off_stg = off_r + (p.res_tma ? RES_BYTES : 0u);
  const uint32_t off_bar = off_stg + (uint32_t)p.n_stg * N_EPI_WARPS * STG_BYTES;
  const uint32_t bar0 = smem_base + off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + off_bar + 8 * BAR_COUNT);
  volatile float* xs = reinterpret_cast<volatile float*>(smem_gen + off_bar + 256);   // [2][128 rows][2] LN exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) >> 7;
  const int NA = p.n_a_stages;
  constexpr int EPI0 = PROD ? 6 : 2;                // first epilogue warp

  if (threadIdx.x == 0) {
    mbar_init(BAR(BAR_WFULL), 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(BAR(BAR_AFULL + s), p.prod ? 4 : 1);      // one arrival per producer warp / TMA tx
      mbar_init(BAR(BAR_AEMPTY + s), 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(BAR(BAR_ACCFULL + s), 1);
      mbar_init(BAR(BAR_ACCEMPTY + s), N_EPI_WARPS);
    }
    mbar_init(BAR(BAR_RFULL), 1);
    mbar_init(BAR(BAR_REMPTY), N_EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    if (!p.prod) tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_o0);
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the set-up above and the weight loads below do not depend on the previous kernel; everything
  // that touches activations comes after pdl_wait()
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (!(warp == 0 && lane == 0)) pdl_wait();

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      mbar_arrive_expect_tx(BAR(BAR_WFULL), (uint32_t)(NB * KC) * BLK_BYTES);
      for (int nb = 0; nb < NB; ++nb)
        for (int kc = 0; kc < KC; ++kc)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_2d(smem_base + off_w + (uint32_t)(nb * KC + kc) * BLK_BYTES + hf * HALF_BYTES, &tm_w,
                        BAR(BAR_WFULL), (kc * 128 + hf * 64) >> p.tf32, nb * 128);
      pdl_wait();
      if (!p.prod) {
        int ac = 0, n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
          for (int kc = 0; kc < KC; ++kc, ++ac) {
            const int s = ac % NA;
            mbar_wait(BAR(BAR_AEMPTY + s), ((ac / NA) & 1) ^ 1, 11);
            mbar_arrive_expect_tx(BAR(BAR_AFULL + s), BLK_BYTES);
            for (int hf = 0; hf < 2; ++hf)
              tma_load_2d(smem_base + off_a + s * BLK_BYTES + hf * HALF_BYTES, &tm_a, BAR(BAR_AFULL + s),
                          (kc * 128 + hf * 64) >> p.tf32, tile * 128);
          }
          if (p.res_tma) {        // residual rows of this tile (rows past M read as zeros)
            mbar_wait(BAR(BAR_REMPTY), (n & 1) ^ 1, 18);
            mbar_arrive_expect_tx(BAR(BAR_RFULL), RES_BYTES);
            for (int c = 0; c < 4; ++c)
              tma_load_2d(smem_base + off_r + c * (RES_BYTES / 4), &tm_o2, BAR(BAR_RFULL), c * 32,
                          p.res_mod > 0 ? (tile * 128) % p.res_mod : tile * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp walks, one elected lane issues) =======================
    mbar_wait(BAR(BAR_WFULL), 0, 12);
    const uint32_t a_lo0 = desc_lo(smem_base + off_a, 16);
    const uint32_t w_lo0 = desc_lo(smem_base + off_w, 16);
    int job = 0, ac0 = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ac0 += KC) {
      for (int nb = 0; nb < NB; ++nb, ++job) {
        const int slot = job & 3;
        GTR(100 + job * 4 + 0);
        mbar_wait(BAR(BAR_ACCEMPTY + slot), ((job >> 2) & 1) ^ 1, 13);
        GTR(100 + job * 4 + 1);
        for (int kc = 0; kc < KC; ++kc) {
          const int ac = ac0 + kc, s = ac % NA;
          if (nb == 0) mbar_wait(BAR(BAR_AFULL + s), (ac / NA) & 1, 14);
          tc_fence_after();
          GTR(100 + job * 4 + 2);
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)s * (BLK_BYTES >> 4);
            const uint32_t w_lo = w_lo0 + (uint32_t)(nb * KC + kc) * (BLK_BYTES >> 4);
            if (p.tf32) {
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_ss_lh_tf32(tmem_base + slot * 128, a_lo + hf * (HALF_BYTES >> 4) + kk * 2,
                                  w_lo + hf * (HALF_BYTES >> 4) + kk * 2, DESC_HI_SW128, IDESC_TF32,
                                  (kc | hf | kk) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_ss_lh(tmem_base + slot * 128, a_lo + hf * (HALF_BYTES >> 4) + kk * 2,
                             w_lo + hf * (HALF_BYTES >> 4) + kk * 2, DESC_HI_SW128, IDESC,
                             (kc | hf | kk) != 0 ? 1u : 0u);
            }
            if (nb == NB - 1) umma_commit(BAR(BAR_AEMPTY + s));   // A chunk no longer needed
            if (kc == KC - 1) umma_commit(BAR(BAR_ACCFULL + slot));
          }
          __syncwarp();
          GTR(100 + job * 4 + 3);
        }
      }
    }
  } else if (PROD && warp < 6) {
    // ======================= A producers: fp32/bf16 rows -> (LayerNorm) -> bf16 UMMA operand ===========
    if (p.prod) {
      const int pw = warp - 2;
      float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.prod == 1) {
        gam = __ldg(reinterpret_cast<const float4*>(p.ln_g) + lane);
        bet = __ldg(reinterpret_cast<const float4*>(p.ln_b) + lane);
      }
      const int cols = (p.prod == 1) ? 128 : p.a_cols;
      const bool lane_ok = lane * 4 < cols;
      const bool src_bf16 = p.prod == 2 && p.a_is_bf16;
      // raw (unconverted) row fragment: 4 fp32, or 4 bf16 in .x/.y -- converted only when consumed, so
      // the loads issued one batch ahead really stay in flight
      auto load_row = [&](long row) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.M && lane_ok) {
          long src = row;
          if (p.win_W > 0) src = window_src_row(row, p.win_W, p.win_half, p.win_jump);
          if (src_bf16) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.a_src) + src * cols) + lane);
            v.x = __uint_as_float(raw.x); v.y = __uint_as_float(raw.y);
          } else {
            v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.a_src) + src * cols) + lane);
          }
        }
        return v;
      };
      int n = 0;
      const int chunk = (lane & 15) >> 1, sub = (lane & 1) * 8;
      float4 x[8], xn[8];
      {
        const long row0 = (long)blockIdx.x * 128 + pw * 32;
#pragma unroll
        for (int u = 0; u < 8; ++u) xn[u] = (blockIdx.x < n_tiles) ? load_row(row0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const long row_bytes = (long)cols * (src_bf16 ? 2 : 4);
      auto prefetch_tile = [&](long tl) {            // this warp's 32 contiguous rows of tile tl -> L2
        const long r0 = tl * 128 + pw * 32;
        if (lane == 0 && r0 < p.M && p.win_W == 0) {
          const long nrows = (p.M - r0 < 32) ? (p.M - r0) : 32;
          const long bytes = (nrows * row_bytes) & ~15L;
          if (bytes > 0) l2_prefetch(reinterpret_cast<const char*>(p.a_src) + r0 * row_bytes, (uint32_t)bytes);
        }
      };
      prefetch_tile((long)blockIdx.x + gridDim.x);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
        const int s = n % NA;
        prefetch_tile((long)tile + 2 * gridDim.x);     // two tiles ahead of the register-staged loads
        if (warp == 2 && lane == 0) GTR(10 + n * 4 + 0);
        mbar_wait(BAR(BAR_AEMPTY + s), ((n / NA) & 1) ^ 1, 15);
        if (warp == 2 && lane == 0) GTR(10 + n * 4 + 1);
        unsigned char* a_half = smem_gen + off_a + s * BLK_BYTES + (lane >> 4) * HALF_BYTES;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) x[u] = xn[u];
          {  // prefetch the next batch of 8 rows (next tile's first batch after the last one)
            const bool last = r0 == 24;
            const long nrow0 = last ? ((long)(tile + gridDim.x) * 128 + pw * 32) : ((long)tile * 128 + pw * 32 + r0 + 8);
            const bool any = !last || (tile + (int)gridDim.x < n_tiles);
#pragma unroll
            for (int u = 0; u < 8; ++u) xn[u] = any ? load_row(nrow0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          // row statistics for the 8 rows of the batch with the reduction STEPS outermost: 8 independent
          // shuffle chains per step instead of 8 serialised 10-step chains (was ~340 cycles per row)
          float y[8][4];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            y[u][0] = x[u].x; y[u][1] = x[u].y; y[u][2] = x[u].z; y[u][3] = x[u].w;
            if (src_bf16) {
              const uint32_t r0b = __float_as_uint(x[u].x), r1b = __float_as_uint(x[u].y);
              y[u][0] = __uint_as_float(r0b << 16); y[u][1] = __uint_as_float(r0b & 0xFFFF0000u);
              y[u][2] = __uint_as_float(r1b << 16); y[u][3] = __uint_as_float(r1b & 0xFFFF0000u);
            }
          }
          if (p.prod == 1) {
            float st[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) st[u] = (y[u][0] + y[u][1]) + (y[u][2] + y[u][3]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
              for (int u = 0; u < 8; ++u) st[u] += __shfl_xor_sync(0xffffffffu, st[u], o);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float mean = st[u] * (1.0f / 128.0f);
              y[u][0] -= mean; y[u][1] -= mean; y[u][2] -= mean; y[u][3] -= mean;
              st[u] = (y[u][0] * y[u][0] + y[u][1] * y[u][1]) + (y[u][2] * y[u][2] + y[u][3] * y[u][3]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
              for (int u = 0; u < 8; ++u) st[u] += __shfl_xor_sync(0xffffffffu, st[u], o);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float rstd = rsqrtf(st[u] * (1.0f / 128.0f) + LN_EPS);
              y[u][0] = y[u][0] * rstd * gam.x + bet.x; y[u][1] = y[u][1] * rstd * gam.y + bet.y;
              y[u][2] = y[u][2] * rstd * gam.z + bet.z; y[u][3] = y[u][3] * rstd * gam.w + bet.w;
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = pw * 32 + r0 + u;
            uint2 pk;
            pk.x = pack_bf16(y[u][0], y[u][1]);
            pk.y = pack_bf16(y[u][2], y[u][3]);
            *reinterpret_cast<uint2*>(a_half + sw128_offset(r, chunk) + sub) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_AFULL + s));
        if (warp == 2 && lane == 0) GTR(10 + n * 4 + 2);
      }
    }
  } else {
    // ======================= epilogue: 8 warps, two per TMEM lane quarter =======================
    // Warp (q, hsel) owns rows 32q..32q+31 and columns [64 hsel, 64 hsel + 64) of every 128-column
    // accumulator slot; it stages its 32-row slab in its own 4 KB buffer and issues its own TMA store,
    // so the epilogue needs no cross-warp barrier.
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int hsel = (warp - EPI0) >> 2;             // which column half of the slot
    const int row = q * 32 + lane;                   // row inside the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t stg_off0 = off_stg + (uint32_t)(warp - EPI0) * p.n_stg * STG_BYTES;
    int job = 0, unit = 0;
    const bool dbl = p.n_stg == 2;
    // with two slabs a store may still be reading the other one: wait only for the store before last
    auto next_slab = [&]() -> uint32_t {
      const uint32_t so = stg_off0 + (dbl ? (unit & 1) * STG_BYTES : 0);
      if (lane == 0) {
        if (dbl) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else tma_store_wait_read0();
      }
      __syncwarp();
      ++unit;
      return so;
    };
    auto issue_store = [&](const CUtensorMap* m, uint32_t so, int c0, int r0) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(m, smem_base + so, c0, r0);
        tma_store_commit();
      }
    };
    int tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const long grow = (long)tile * 128 + row;
      const bool row_ok = grow < p.M;
      const float* res_row = nullptr;
      if (p.residual) res_row = p.residual + (p.res_mod > 0 ? (grow % p.res_mod) : grow) * 128 + hsel * 64;
      for (int nb = 0; nb < NB; ++nb, ++job) {
        const int slot = job & 3;
        if (warp == EPI0 && lane == 0) GTR(400 + job * 4 + 0);
        mbar_wait(BAR(BAR_ACCFULL + slot), (job >> 2) & 1, 16);
        tc_fence_after();
        if (warp == EPI0 && lane == 0) GTR(400 + job * 4 + 1);
        const uint32_t tacc = tmem_base + lane_addr + slot * 128 + hsel * 64;
        const int ocol0 = (p.out_split ? 0 : nb * 128) + hsel * 64;
        uint32_t v[2][32];
        uint32_t bf_so = 0;
        tmem_ld32(tacc, v[0]);
        tmem_ld32(tacc + 32, v[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_ACCEMPTY + slot));   // this warp's share of the slot is drained
        if (warp == EPI0 && lane == 0) GTR(400 + job * 4 + 2);
        // phase 1: bias / ReLU / residual, in place in the accumulator registers
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + nb * 128 + hsel * 64 + cb * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bp + i);                  // warp-uniform address: one transaction
            v[cb][4 * i] = __float_as_uint(__uint_as_float(v[cb][4 * i]) + b4.x);
            v[cb][4 * i + 1] = __float_as_uint(__uint_as_float(v[cb][4 * i + 1]) + b4.y);
            v[cb][4 * i + 2] = __float_as_uint(__uint_as_float(v[cb][4 * i + 2]) + b4.z);
            v[cb][4 * i + 3] = __float_as_uint(__uint_as_float(v[cb][4 * i + 3]) + b4.w);
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[cb][i] = __float_as_uint(fmaxf(__uint_as_float(v[cb][i]), 0.f));
          }
          if (p.res_tma) {
            if (cb == 0) mbar_wait(BAR(BAR_RFULL), tcount & 1, 19);
            const unsigned char* rb = smem_gen + off_r + (uint32_t)(hsel * 2 + cb) * (RES_BYTES / 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 r4 = *reinterpret_cast<const float4*>(rb + sw128_offset(row, i));
              v[cb][4 * i] = __float_as_uint(__uint_as_float(v[cb][4 * i]) + r4.x);
              v[cb][4 * i + 1] = __float_as_uint(__uint_as_float(v[cb][4 * i + 1]) + r4.y);
              v[cb][4 * i + 2] = __float_as_uint(__uint_as_float(v[cb][4 * i + 2]) + r4.z);
              v[cb][4 * i + 3] = __float_as_uint(__uint_as_float(v[cb][4 * i + 3]) + r4.w);
            }
          } else if (p.residual && row_ok) {                   // positional table / non-TMA residual rows
            const float4* rp = reinterpret_cast<const float4*>(res_row + cb * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 r4 = rp[i];   // plain load: the buffer may also be this kernel's output
              v[cb][4 * i] = __float_as_uint(__uint_as_float(v[cb][4 * i]) + r4.x);
              v[cb][4 * i + 1] = __float_as_uint(__uint_as_float(v[cb][4 * i + 1]) + r4.y);
              v[cb][4 * i + 2] = __float_as_uint(__uint_as_float(v[cb][4 * i + 2]) + r4.z);
              v[cb][4 * i + 3] = __float_as_uint(__uint_as_float(v[cb][4 * i + 3]) + r4.w);
            }
          }
        }
        if (p.res_tma) {          // residual tile consumed: the TMA warp may fetch the next one
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(BAR_REMPTY));
        }
        // phase 2: stage and store
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          if (p.out_f32 && p.out_tiled) {
            // tiled residual layout [tile][32 column quads][128 rows][4 floats]: a warp's 32 rows of one column
            // quad are 512 contiguous bytes
            float4* dst = reinterpret_cast<float4*>(p.out_ptr[0]) + (size_t)tile * 4096 + (size_t)(hsel * 16 + cb * 8) * 128 + row;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              dst[c * 128] = make_float4(__uint_as_float(v[cb][4 * c]), __uint_as_float(v[cb][4 * c + 1]),
                                         __uint_as_float(v[cb][4 * c + 2]), __uint_as_float(v[cb][4 * c + 3]));
          } else if (p.out_f32) {
            // one store unit = 32 fp32 columns (128 B per row)
            const uint32_t so = next_slab();
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(smem_gen + so + sw128_offset(lane, c)) =
                  make_uint4(v[cb][4 * c], v[cb][4 * c + 1], v[cb][4 * c + 2], v[cb][4 * c + 3]);
            issue_store(&tm_o0, so, ocol0 + cb * 32, tile * 128 + q * 32);
          } else {
            // one store unit = 64 bf16 columns (128 B per row) = both 32-column chunks
            if (cb == 0) bf_so = next_slab();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t* f = &v[cb][8 * c];
              *reinterpret_cast<uint4*>(smem_gen + bf_so + sw128_offset(lane, cb * 4 + c)) =
                  make_uint4(pack_bf16(__uint_as_float(f[0]), __uint_as_float(f[1])),
                             pack_bf16(__uint_as_float(f[2]), __uint_as_float(f[3])),
                             pack_bf16(__uint_as_float(f[4]), __uint_as_float(f[5])),
                             pack_bf16(__uint_as_float(f[6]), __uint_as_float(f[7])));
            }
            if (cb == 1) {
              const CUtensorMap* om = (p.out_split && nb == 1) ? &tm_o1 : (p.out_split && nb == 2) ? &tm_o2 : &tm_o0;
              issue_store(om, bf_so, ocol0, tile * 128 + q * 32);
            }
          }
        }
        if (p.emit_g) {
          // LayerNorm of the finished row (biased variance, eps 1e-5): this thread holds 64 of the 128
          // columns, its partner (same lane, the other warp of this lane quarter) the rest
          float s1 = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) s1 += __uint_as_float(v[i >> 5][i & 31]);
          xs[row * 2 + hsel] = s1;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          const float mean = (s1 + xs[row * 2 + (hsel ^ 1)]) * (1.0f / 128.0f);
          float s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const float d = __uint_as_float(v[i >> 5][i & 31]) - mean;
            s2 = fmaf(d, d, s2);
          }
          xs[256 + row * 2 + hsel] = s2;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          const float rstd = rsqrtf((s2 + xs[256 + row * 2 + (hsel ^ 1)]) * (1.0f / 128.0f) + LN_EPS);
          const float4* gp = reinterpret_cast<const float4*>(p.emit_g + hsel * 64);
          const float4* bp2 = reinterpret_cast<const float4*>(p.emit_b + hsel * 64);
          const uint32_t so = next_slab();
          unsigned char* stg = smem_gen + so;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i0 = c * 8 + e * 4;
              const float4 g4 = __ldg(gp + (i0 >> 2)), b4 = __ldg(bp2 + (i0 >> 2));
              const float y0 = (__uint_as_float(v[i0 >> 5][i0 & 31]) - mean) * rstd * g4.x + b4.x;
              const float y1 = (__uint_as_float(v[(i0 + 1) >> 5][(i0 + 1) & 31]) - mean) * rstd * g4.y + b4.y;
              const float y2 = (__uint_as_float(v[(i0 + 2) >> 5][(i0 + 2) & 31]) - mean) * rstd * g4.z + b4.z;
              const float y3 = (__uint_as_float(v[(i0 + 3) >> 5][(i0 + 3) & 31]) - mean) * rstd * g4.w + b4.w;
              pk[2 * e] = pack_bf16(y0, y1);
              pk[2 * e + 1] = pack_bf16(y2, y3);
            }
            *reinterpret_cast<uint4*>(stg + sw128_offset(lane, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          issue_store(&tm_o1, so, hsel * 64, tile * 128 + q * 32);
        }
        if (warp == EPI0 && lane == 0) GTR(400 + job * 4 + 3);
      }
    }
  }

  if (warp >= EPI0 && lane == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

CUresult make_tmap_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esz, long rows,
                      long cols, int box_cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return encode_tiled(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

cudaError_t launch_gemm_tc(const GemmTcArgs& a, int num_sms, cudaStream_t s, std::string* err) {
  if (a.M <= 0) return cudaSuccess;
  auto bad = [&](const char* m) { if (err) *err = m; return cudaErrorInvalidValue; };
  if (a.N % 128 || a.K % 128 || a.N > 512 || (long)a.N * a.K > 65536) return bad("unsupported N/K");
  if ((a.N > 128) && (a.K > 128)) return bad("N > 128 requires K == 128");
  if (a.ln_g && a.K != 128) return bad("LayerNorm prologue needs K == 128");
  if (a.residual && a.N != 128) return bad("residual needs N == 128");
  if (a.a_rows && (a.K != 128 || a.a_cols <= 0 || a.a_cols > 128 || a.a_cols % 4)) return bad("bad a_cols");
  if (a.a_f32_tma && (!a.w_f32 || a.a_cols <= 0 || a.a_cols % 4 || a.K != 128 * ((a.a_cols + 63) / 64) || a.ln_g || a.a_rows))
    return bad("bad tf32 front-end arguments");
  GemmTcParams p = {};
  p.M = a.M; p.N = a.N; p.K = a.K; p.tf32 = a.a_f32_tma ? 1 : 0;
  p.prod = a.ln_g ? 1 : (a.a_rows ? 2 : 0);
  p.a_src = a.ln_g ? (const void*)a.a_f32 : a.a_rows;
  p.a_cols = a.a_cols; p.a_is_bf16 = a.a_rows_bf16;
  p.win_W = a.win_W; p.win_half = a.win_half; p.win_jump = a.win_jump;
  p.ln_g = a.ln_g; p.ln_b = a.ln_b;
  p.bias = a.bias; p.relu = a.relu; p.residual = a.residual; p.res_mod = a.res_mod;
  p.out_f32 = a.out_f32; p.out_split = a.out[1] != nullptr && !a.emit_ln_g;
  p.emit_g = a.emit_ln_g; p.emit_b = a.emit_ln_b;
  p.out_tiled = a.out_tiled;
  if (a.out_tiled && !(a.out_f32 && a.N == 128)) return bad("tiled output needs fp32 N == 128");
  // residual tiles by TMA: the residual stream itself, or the positional-encoding table when a 128-row tile
  // never wraps around it (res_mod % 128 == 0: row (tile * 128) % res_mod onwards is contiguous)
  p.res_tma = (a.residual && (a.res_mod == 0 || a.res_mod % 128 == 0) && !p.prod && a.out_f32) ? 1 : 0;
  for (int j = 0; j < 3; ++j) p.out_ptr[j] = a.out[j] ? a.out[j] : a.out[0];
  if (a.emit_ln_g && !(a.out_f32 && a.N == 128 && a.out[1])) return bad("LayerNorm emit needs fp32 N == 128 output + out[1]");
  const uint32_t w_bytes = (uint32_t)(a.N / 128) * (a.K / 128) * BLK_BYTES;
  // shared-memory plan: resident W + A ring + per-warp staging slabs.  Prefer two staging slabs per
  // warp (a TMA store takes ~1 us to drain a slab) with at least two A stages; fall back to one slab.
  const uint32_t LIMIT = 232448u, misc = 256 + 2048;
  const uint32_t stg1 = N_EPI_WARPS * STG_BYTES;
  const uint32_t fixed = w_bytes + (p.res_tma ? RES_BYTES : 0u) + misc;
  p.n_stg = (fixed + 2 * BLK_BYTES + 2 * stg1 <= LIMIT) ? 2 : 1;
  p.n_a_stages = 1;
  while (p.n_a_stages < 3 && fixed + (p.n_a_stages + 1) * BLK_BYTES + p.n_stg * stg1 <= LIMIT) ++p.n_a_stages;
  const uint32_t smem = fixed + p.n_a_stages * BLK_BYTES + p.n_stg * stg1;
  if (smem > LIMIT) return bad("shared memory budget exceeded");

  CUtensorMap tw, ta, to[3];
  CUresult r;
  if (p.tf32) {
    // fp32 operands, boxes of 32 floats (one 128-byte swizzled row chunk); columns >= a_cols read as zeros
    r = make_tmap_2d(&tw, a.w_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.N, a.a_cols, 32, 128);
    if (r == CUDA_SUCCESS) r = make_tmap_2d(&ta, a.a_f32_tma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, a.a_cols, 32, 128);
  } else {
    r = make_tmap_2d(&tw, a.w_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.N, a.K, 64, 128);
    if (r == CUDA_SUCCESS)
      // a_cols (front end, bf16 features): the A rows are only a_cols wide; TMA zero-fills up to K
      r = p.prod ? CUDA_SUCCESS
                 : make_tmap_2d(&ta, a.a_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M,
                                (a.a_cols > 0 && !a.a_rows) ? a.a_cols : a.K, 64, 128);
  }
  if (p.prod) ta = tw;
  const int ocols = p.out_split ? 128 : a.N;
  for (int j = 0; j < 3 && r == CUDA_SUCCESS; ++j) {
    void* optr = a.out[j] ? a.out[j] : a.out[0];
    if (a.emit_ln_g && j == 1)       // bf16 LayerNorm copy [M,128]
      r = make_tmap_2d(&to[j], optr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 64, 32);
    else if (p.res_tma && j == 2)    // residual load map: 128-row boxes of 32 fp32 columns
      r = make_tmap_2d(&to[j], a.residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.res_mod > 0 ? a.res_mod : a.M, 128, 32, 128);
    else
      r = a.out_f32 ? make_tmap_2d(&to[j], optr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, ocols, 32, 32)
                    : make_tmap_2d(&to[j], optr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, ocols, 64, 32);
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return cudaErrorInvalidValue;
  }
  const int n_tiles = (a.M + 127) / 128;
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  static const char* want_trace = getenv("VADB_GEMM_TRACE");     // e.g. "384" = trace launches with N == 384
  if (want_trace && atoi(want_trace) == a.N + (p.prod ? 0 : 1000)) {
    cudaError_t e = p.prod ? cudaFuncSetAttribute(gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long* dtr = nullptr;
    cudaMalloc(&dtr, 1024 * sizeof(long long));
    cudaMemsetAsync(dtr, 0, 1024 * sizeof(long long), s);
    if (p.prod) gemm_tc_kernel<true, true><<<grid, NTHREADS, smem, s>>>(tw, ta, to[0], to[1], to[2], p, dtr);
    else gemm_tc_kernel<true, false><<<grid, NTHREADS - 128, smem, s>>>(tw, ta, to[0], to[1], to[2], p, dtr);
    std::vector<long long> ht(1024);
    cudaMemcpyAsync(ht.data(), dtr, 1024 * sizeof(long long), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(dtr);
    long long t0 = 0;
    for (long long v : ht) if (v && (!t0 || v < t0)) t0 = v;
    auto rel = [&](int i) { return ht[i] ? (long long)(ht[i] - t0) : -1LL; };
    const int NBt = a.N / 128;
    fprintf(stderr, "[gemm trace] N=%d K=%d prod=%d stages A=%d stg=%d\n", a.N, a.K, p.prod, p.n_a_stages, p.n_stg);
    for (int n = 0; n < 7; ++n) {
      fprintf(stderr, "[gemm trace] tile#%d producer: wait_empty %lld->%lld filled %lld\n", n, rel(10 + n * 4), rel(10 + n * 4 + 1), rel(10 + n * 4 + 2));
      for (int nb = 0; nb < NBt; ++nb) {
        const int job = n * NBt + nb;
        fprintf(stderr, "[gemm trace]   job %d mma: wait_acc %lld->%lld a_ready %lld issued %lld | epi: wait %lld->%lld drained %lld stored %lld\n",
                job, rel(100 + job * 4), rel(100 + job * 4 + 1), rel(100 + job * 4 + 2), rel(100 + job * 4 + 3),
                rel(400 + job * 4), rel(400 + job * 4 + 1), rel(400 + job * 4 + 2), rel(400 + job * 4 + 3));
      }
    }
    return cudaGetLastError();
  }
  // opt in to the full dynamic shared-memory carve-out once per device and variant
  static thread_local int attr_dev[2] = {-1, -1};
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev[p.prod ? 1 : 0] != dev) {
    cudaError_t e = p.prod ? cudaFuncSetAttribute(gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)
                           : cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_dev[p.prod ? 1 : 0] = dev;
  }
  if (p.prod) return launch_k(gemm_tc_kernel<false, true>, grid, NTHREADS, smem, s, tw, ta, to[0], to[1], to[2], p, (long long*)nullptr);
  return launch_k(gemm_tc_kernel<false, false>, grid, NTHREADS - 128, smem, s, tw, ta, to[0], to[1], to[2], p, (long long*)nullptr);
  return cudaGetLastError();
}

}  // namespace vadb
