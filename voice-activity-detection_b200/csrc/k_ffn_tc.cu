// Fused position-wise feed-forward sublayer on the 5th-gen tensor cores (sm_100a):
//
//   h <- h + ReLU( LayerNorm(h) W1^T + b1 ) W2^T + b2          (vad/modeling/transformer.py:234-238, :366-375)
//
// in ONE persistent kernel: the [frames x 512] hidden activation never leaves the SM (the unfused path
// writes and re-reads 2 KB per frame for it).  Per 128-row tile:
//   TMA (warp 0)               a = LayerNorm(h) in bf16 (emitted by the previous kernel's epilogue) -> A operand
//   GEMM2 x4 (MMA warp)        hid_nb[128x128] = A . W1[nb]^T           -> TMEM slot nb & 1 (fp32)
//   epilogue warps (6-13)      + b1, ReLU, bf16 pairs written back over the head of the same TMEM columns
//   GEMM3 x4 (MMA warp)        acc[128x128] += hid_nb (A operand read from TMEM) . W2[:, nb]^T
//   epilogue warps             + b2 + residual h -> fp32 -> per-warp staging slab -> TMA store; optionally
//                              LayerNorm(new h) -> bf16 for the next layer's Q/K/V GEMM
// W1/W2 (256 KB in bf16) do not fit next to the operands, so their eight [128 x 128] blocks stream
// through a 4-stage ring per tile, in the exact order the MMA warp consumes them; they are L2-resident.
// TMEM: two hidden slots (columns 0-255) + double-buffered output accumulator (256-511).
#include "tc_common.cuh"
#include "vadb_common.cuh"

namespace vadb {
namespace {

using namespace tc;

constexpr int NTHREADS = 320;          // TMA + MMA + 8 epilogue warps (up to 168 registers per thread: no spills)
constexpr int N_EPI_WARPS = 8;
constexpr int NW = 3;                  // weight ring stages
constexpr int NA = 2;                  // A (LayerNorm output) stages
constexpr uint32_t BLK_BYTES = 128 * 128 * 2;     // [128 x 128] bf16 block = two SW128 halves of 16 KB
constexpr uint32_t HALF_BYTES = 128 * 128;
constexpr uint32_t STG_BYTES = 32 * 128;          // per-warp staging slab: 32 rows x 128 B
constexpr uint32_t IDESC = idesc_bf16(128, 128, 0, 0);

constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_A = OFF_W + NW * BLK_BYTES;
constexpr uint32_t OFF_STG = OFF_A + NA * BLK_BYTES;
constexpr uint32_t OFF_BAR = OFF_STG + 2 * N_EPI_WARPS * STG_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 2048;      // barriers + LayerNorm exchange
constexpr uint32_t SMEM_ALLOC = SMEM_BYTES;                // the dynamic smem window is declared 1024-aligned

enum { B_AFULL = 0, B_AEMPTY = 2, B_WFULL = 4 /* 4 slots reserved */, B_WEMPTY = 8, B_HIDFULL = 12, B_HIDBF = 14, B_OUTFULL = 16,
       B_OUTEMPTY = 18, B_RES = 20 /* one per epilogue warp */, B_COUNT = 28 };

constexpr uint32_t TM_HID = 0;      // hidden slot s at columns 128 s (fp32); its bf16 copy (A operand of GEMM3)
                                    // is written by each epilogue thread over the head of its own 64 columns
constexpr uint32_t TM_OUT = 256;    // output accumulator b at columns 256 + 128 b

struct FfnParams {
  int M;
  const float* h;        // [M,128] fp32 residual stream (input; the output map points at the same buffer)
  const float* emit_g;   // optional LayerNorm emit of the output rows (next layer's pre-LN)
  const float* emit_b;
  float* h_out;          // == h (in place)
  bf16* emit_out;
  const float* b1;       // [512]
  const float* b2;       // [128]
  // last layer: final LayerNorm + classifier + log-softmax fused into the epilogue (h is not stored)
  const float* cls_g;    // [128] final LayerNorm gamma, nullptr -> not the last layer
  const float* cls_b;    // [128]
  const float* cls_w;    // [2,128] classifier weight
  const float* cls_bias; // [2]
  float* prob;           // [M] or nullptr
  float* logp;           // [M,2] or nullptr
  int logp_vec;          // logp is 8-byte aligned: one float2 store per row
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}

// Order in which the MMA warp consumes the weight blocks of one tile (GEMM2 runs one n-block ahead of
// GEMM3 so the ReLU epilogue of block nb overlaps GEMM2 of block nb+1):
//   seq:  0      1      2      3      4      5      6      7
//         W1[0]  W1[1]  W2[0]  W1[2]  W2[1]  W1[3]  W2[2]  W2[3]
__device__ __forceinline__ void wblock_of_seq(int q, int& is_w2, int& nb) {
  is_w2 = (0b11010100 >> q) & 1;          // bit q set -> W2 block
  nb = (0xED84 >> (2 * q)) & 3;            // {0, 1, 0, 2, 1, 3, 2, 3}
}

__global__ void __launch_bounds__(NTHREADS, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
              const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_out,
              const __grid_constant__ CUtensorMap tm_emit, const FfnParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  unsigned char* smem_gen = smem_raw;
  if ((smem_base & 1023u) != 0) {   // 128B-swizzled tiles need 1024-byte alignment
    if (threadIdx.x == 0) printf("vadb: ffn kernel shared memory window not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t bar0 = smem_base + OFF_BAR;
  volatile float* xs = reinterpret_cast<volatile float*>(smem_gen + OFF_BAR + 256);   // [2][128][2] LN exchange
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 8 * B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) >> 7;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(BAR(B_AFULL + s), 1); mbar_init(BAR(B_AEMPTY + s), 1); }
    for (int s = 0; s < NW; ++s) { mbar_init(BAR(B_WFULL + s), 1); mbar_init(BAR(B_WEMPTY + s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_HIDFULL + s), 1);
      mbar_init(BAR(B_HIDBF + s), N_EPI_WARPS);
      mbar_init(BAR(B_OUTFULL + s), 1);
      mbar_init(BAR(B_OUTEMPTY + s), N_EPI_WARPS);
    }
    for (int w = 0; w < N_EPI_WARPS; ++w) mbar_init(BAR(B_RES + w), 1);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_out);
    tma_prefetch_desc(&tm_a);
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();                 // the set-up above overlaps the previous kernel's tail (vadb_common.cuh)

  if (warp == 0) {
    // ======================= weight streamer =======================
    if (lane == 0) {
      int wc = 0, n = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
        {  // A operand of this tile: LayerNorm(h) rows in bf16, rows past M read as zeros
          const int as = n % NA;
          mbar_wait(BAR(B_AEMPTY + as), ((n / NA) & 1) ^ 1, 20);
          mbar_arrive_expect_tx(BAR(B_AFULL + as), BLK_BYTES);
          tma_load_2d(smem_base + OFF_A + as * BLK_BYTES, &tm_a, BAR(B_AFULL + as), 0, tile * 128);
          tma_load_2d(smem_base + OFF_A + as * BLK_BYTES + HALF_BYTES, &tm_a, BAR(B_AFULL + as), 64, tile * 128);
        }
        for (int q = 0; q < 8; ++q, ++wc) {
          int is_w2, nb;
          wblock_of_seq(q, is_w2, nb);
          const int s = wc % NW;
          mbar_wait(BAR(B_WEMPTY + s), ((wc / NW) & 1) ^ 1, 21);
          mbar_arrive_expect_tx(BAR(B_WFULL + s), BLK_BYTES);
          const uint32_t dst = smem_base + OFF_W + s * BLK_BYTES;
          if (!is_w2) {          // W1 [512,128]: rows nb*128.., k halves
            tma_load_2d(dst, &tm_w1, BAR(B_WFULL + s), 0, nb * 128);
            tma_load_2d(dst + HALF_BYTES, &tm_w1, BAR(B_WFULL + s), 64, nb * 128);
          } else {               // W2 [128,512]: all 128 rows, k columns nb*128..
            tma_load_2d(dst, &tm_w2, BAR(B_WFULL + s), nb * 128, 0);
            tma_load_2d(dst + HALF_BYTES, &tm_w2, BAR(B_WFULL + s), nb * 128 + 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp walks, one elected lane issues) =======================
    const uint32_t a_lo0 = desc_lo(smem_base + OFF_A, 16);
    const uint32_t w_lo0 = desc_lo(smem_base + OFF_W, 16);
    int wc = 0, n = 0;
    int hid_uses[2] = {0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const int as = n % NA, ob = n & 1;
      mbar_wait(BAR(B_AFULL + as), (n / NA) & 1, 22);
      mbar_wait(BAR(B_OUTEMPTY + ob), ((n >> 1) & 1) ^ 1, 23);    // final epilogue of tile n-2 drained acc[ob]
      for (int q = 0; q < 8; ++q, ++wc) {
        int is_w2, nb;
        wblock_of_seq(q, is_w2, nb);
        const int ws = wc % NW, hs = nb & 1;
        mbar_wait(BAR(B_WFULL + ws), (wc / NW) & 1, 24);
        if (is_w2) mbar_wait(BAR(B_HIDBF + hs), (hid_uses[hs] - 1) & 1, 25);   // ReLU'd bf16 block in TMEM
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_lo = w_lo0 + (uint32_t)ws * (BLK_BYTES >> 4);
          if (!is_w2) {
            // GEMM2: hid slot hs = A(LN) . W1[nb]^T, K = 128
            const uint32_t a_lo = a_lo0 + (uint32_t)as * (BLK_BYTES >> 4);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss_lh(tmem_base + TM_HID + 128u * hs, a_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2,
                         w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2, DESC_HI_SW128, IDESC,
                         kk != 0 ? 1u : 0u);
            umma_commit(BAR(B_HIDFULL + hs));
            if (nb == 3) umma_commit(BAR(B_AEMPTY + as));          // LN operand no longer needed
          } else {
            // GEMM3: acc[ob] += hid_nb(bf16, TMEM) . W2[:, nb]^T, K = 128
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lh(tmem_base + TM_OUT + 128u * ob,
                         tmem_base + TM_HID + 128u * hs + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8),
                         w_lo + (kk >> 2) * (HALF_BYTES >> 4) + (kk & 3) * 2, DESC_HI_SW128, IDESC,
                         (nb | kk) != 0 ? 1u : 0u);
            if (nb == 3) umma_commit(BAR(B_OUTFULL + ob));
          }
          umma_commit(BAR(B_WEMPTY + ws));
        }
        __syncwarp();
        if (!is_w2) hid_uses[hs]++;
      }
    }
  } else {
    // ======================= epilogue warps: two per TMEM lane quarter, 64 columns each =======================
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t stg_off0 = OFF_STG + (uint32_t)(warp - 2) * 2 * STG_BYTES;      // this warp's two staging slabs
    int n = 0;
    const uint32_t res_bar = BAR(B_RES + (warp - 2));
    auto issue_store = [&](const CUtensorMap* m, uint32_t so, int c0, int r0) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(m, smem_base + so, c0, r0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    int hid_uses[2] = {0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      // This warp's share of the residual rows ([32 rows x 64 fp32] of h) lands by TMA in its own two staging
      // slabs while the hidden blocks are processed; the final epilogue updates the slabs in place and
      // stores them back.  (Per-thread row loads here cost two exposed DRAM round trips per tile.)
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // last tile's stores have left the slabs
        mbar_arrive_expect_tx(res_bar, 2 * STG_BYTES);
        tma_load_2d(smem_base + stg_off0, &tm_out, res_bar, hsel * 64, tile * 128 + q * 32);
        tma_load_2d(smem_base + stg_off0 + STG_BYTES, &tm_out, res_bar, hsel * 64 + 32, tile * 128 + q * 32);
      }
      __syncwarp();
      for (int nb = 0; nb < 4; ++nb) {
        const int hs = nb & 1;
        mbar_wait(BAR(B_HIDFULL + hs), hid_uses[hs] & 1, 27);
        hid_uses[hs]++;
        tc_fence_after();
        const uint32_t th = tmem_base + lane_addr + TM_HID + 128u * hs + 64u * hsel;
        uint32_t v[2][32];
        tmem_ld32(th, v[0]);
        tmem_ld32(th + 32, v[1]);
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
        tmem_st32(th, pk);               // bf16 pairs over the head of this thread's own 64 fp32 columns
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HIDBF + hs));
      }
      // final epilogue: + b2 + residual -> fp32 -> staging -> TMA store (two units of 32 columns)
      const int ob = n & 1;
      mbar_wait(BAR(B_OUTFULL + ob), (n >> 1) & 1, 28);
      tc_fence_after();
      const uint32_t tacc = tmem_base + lane_addr + TM_OUT + 128u * ob + 64u * hsel;
      uint32_t v[2][32];
      tmem_ld32(tacc, v[0]);
      tmem_ld32(tacc + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_OUTEMPTY + ob));
      mbar_wait(res_bar, n & 1, 29);
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        const float4* bp = reinterpret_cast<const float4*>(p.b2 + hsel * 64 + cb * 32);
        const uint32_t so = stg_off0 + cb * STG_BYTES;
        unsigned char* stg = smem_gen + so;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 b4 = __ldg(bp + c4);
          float4* cell = reinterpret_cast<float4*>(stg + sw128_offset(lane, c4));
          const float4 r4 = *cell;
          const uint32_t* s4 = &v[cb][c4 * 4];
          const float4 f4 = make_float4(__uint_as_float(s4[0]) + b4.x + r4.x, __uint_as_float(s4[1]) + b4.y + r4.y,
                                        __uint_as_float(s4[2]) + b4.z + r4.z, __uint_as_float(s4[3]) + b4.w + r4.w);
          if (!p.cls_g) *cell = f4;
          v[cb][c4 * 4] = __float_as_uint(f4.x); v[cb][c4 * 4 + 1] = __float_as_uint(f4.y);
          v[cb][c4 * 4 + 2] = __float_as_uint(f4.z); v[cb][c4 * 4 + 3] = __float_as_uint(f4.w);
        }
        if (!p.cls_g) issue_store(&tm_out, so, hsel * 64 + cb * 32, tile * 128 + q * 32);
      }
      if (p.cls_g) {
        // Last layer: the encoder's final LayerNorm (transformer.py:33), the classifier and the
        // log-softmax (self_attention.py:26-27) and the caller's softmax(...)[...,1] = sigmoid(z1 - z0)
        // (predictor.py:225,257-258) on the row that is still in registers: h is never written and the
        // separate classifier pass (512 B/frame re-read) disappears.  This thread holds 64 of the 128
        // columns, its partner (same lane, other warp of the quarter) the rest.
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
        const float rstd = 1.0f / sqrtf((s2 + xs[256 + row * 2 + (hsel ^ 1)]) * (1.0f / 128.0f) + LN_EPS);
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
      } else if (p.emit_g) {
        // LayerNorm of the new h row for the next layer's Q/K/V GEMM (transformer.py:235-236): this
        // thread holds 64 of the 128 columns, its partner (same lane, other warp of the quarter) the rest
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
        const uint32_t so = stg_off0;                 // slab 0 again, once its fp32 store has been read out
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
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
        issue_store(&tm_emit, so, hsel * 64, tile * 128 + q * 32);
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

CUresult tmap2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esz, long rows, long cols,
                int box_cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return encode_tiled(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

cudaError_t launch_ffn_tc(const FfnTcArgs& a, int num_sms, cudaStream_t s, std::string* err) {
  if (a.M <= 0) return cudaSuccess;
  CUtensorMap t1, t2, ta, to, te;
  CUresult r = tmap2d(&t1, a.w1_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 512, 128, 64, 128);
  if (r == CUDA_SUCCESS) r = tmap2d(&t2, a.w2_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 128, 512, 64, 128);
  if (r == CUDA_SUCCESS) r = tmap2d(&ta, a.a_ln, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, 128, 64, 128);
  if (r == CUDA_SUCCESS) r = tmap2d(&to, a.h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, 128, 32, 32);
  if (r == CUDA_SUCCESS)
    r = tmap2d(&te, a.emit_out ? (const void*)a.emit_out : (const void*)a.a_ln, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
               a.M, 128, 64, 32);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return cudaErrorInvalidValue;
  }
  FfnParams p = {};
  p.M = a.M; p.h = a.h; p.b1 = a.b1; p.b2 = a.b2;
  p.emit_g = a.emit_out ? a.emit_ln_g : nullptr; p.emit_b = a.emit_ln_b;
  p.h_out = a.h; p.emit_out = a.emit_out;
  p.cls_g = a.cls_ln_g; p.cls_b = a.cls_ln_b; p.cls_w = a.cls_w; p.cls_bias = a.cls_bias;
  p.prob = a.prob; p.logp = a.logp;
  p.logp_vec = (reinterpret_cast<uintptr_t>(a.logp) % 8) == 0;
  if (p.cls_g && (!p.cls_b || !p.cls_w || !p.cls_bias)) return cudaErrorInvalidValue;
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC);
    if (e != cudaSuccess) return e;
    attr_dev = dev;
  }
  const int n_tiles = (a.M + 127) / 128;
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  { cudaError_t e = launch_k(ffn_tc_kernel, grid, NTHREADS, SMEM_ALLOC, s, t1, t2, ta, to, te, p); if (e != cudaSuccess) return e; }
  return cudaGetLastError();
}

}  // namespace vadb
