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
//   warps 2-5   LayerNorm producers (A = fp32 residual stream): coalesced loads, warp-shuffle row
//               statistics, bf16 into the 128B-swizzled UMMA operand layout
//   warps 6-9   epilogue: tcgen05.ld accumulators -> bias/ReLU/residual -> swizzled staging tile ->
//               TMA store (full-line writes); one thread per row
// Accumulators: four 128-column TMEM slots used as a ring over (tile, n-block) jobs, so the MMAs of
// the next job overlap the epilogue of the previous one.
#include "tc_common.cuh"
#include "vadb_common.cuh"

namespace vadb {
namespace {

using namespace tc;

constexpr int NTHREADS = 320;
constexpr uint32_t BLK_BYTES = 128 * 128 * 2;     // [128 x 128] bf16 block = two SW128 halves of 16 KB
constexpr uint32_t HALF_BYTES = 128 * 128;        // [128 rows x 64 k] bf16
constexpr uint32_t STG_BYTES = 128 * 128;         // staging tile: 128 rows x 128 B
constexpr uint32_t IDESC = idesc_bf16(128, 128, 0, 0);

enum { BAR_WFULL = 0, BAR_AFULL = 1, BAR_AEMPTY = 4, BAR_ACCFULL = 7, BAR_ACCEMPTY = 11, BAR_COUNT = 15 };

struct GemmTcParams {
  int M, N, K;              // N, K multiples of 128; N <= 512; N*K <= 65536
  int n_a_stages;           // 2 or 3
  int ln;                   // 1: A = fp32 [M,128] through LayerNorm (K == 128)
  const float* a_f32;       // LN mode source
  const float* ln_g;
  const float* ln_b;
  const float* bias;        // [N]
  int relu;
  const float* residual;    // fp32 [M, 128] (N == 128) or nullptr
  int out_f32;              // 1: fp32 output [M, N]; 0: bf16
  int out_split;            // bf16 outputs: columns [j*128, j*128+128) -> out map j when split (q,k,v)
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_a,
               const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1,
               const __grid_constant__ CUtensorMap tm_o2, const GemmTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int NB = p.N >> 7, KC = p.K >> 7;
  const uint32_t off_w = 0;
  const uint32_t off_a = off_w + (uint32_t)(NB * KC) * BLK_BYTES;
  const uint32_t off_stg = off_a + (uint32_t)p.n_a_stages * BLK_BYTES;
  const uint32_t off_bar = off_stg + 2 * STG_BYTES;
  const uint32_t bar0 = smem_base + off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + off_bar + 8 * BAR_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) >> 7;
  const int NA = p.n_a_stages;

  if (threadIdx.x == 0) {
    mbar_init(BAR(BAR_WFULL), 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(BAR(BAR_AFULL + s), p.ln ? 128 : 1);
      mbar_init(BAR(BAR_AEMPTY + s), 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(BAR(BAR_ACCFULL + s), 1);
      mbar_init(BAR(BAR_ACCEMPTY + s), 128);
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    if (!p.ln) tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_o0);
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      mbar_arrive_expect_tx(BAR(BAR_WFULL), (uint32_t)(NB * KC) * BLK_BYTES);
      for (int nb = 0; nb < NB; ++nb)
        for (int kc = 0; kc < KC; ++kc)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_2d(smem_base + off_w + (uint32_t)(nb * KC + kc) * BLK_BYTES + hf * HALF_BYTES, &tm_w,
                        BAR(BAR_WFULL), kc * 128 + hf * 64, nb * 128);
      if (!p.ln) {
        int ac = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
          for (int kc = 0; kc < KC; ++kc, ++ac) {
            const int s = ac % NA;
            mbar_wait(BAR(BAR_AEMPTY + s), ((ac / NA) & 1) ^ 1, 11);
            mbar_arrive_expect_tx(BAR(BAR_AFULL + s), BLK_BYTES);
            for (int hf = 0; hf < 2; ++hf)
              tma_load_2d(smem_base + off_a + s * BLK_BYTES + hf * HALF_BYTES, &tm_a, BAR(BAR_AFULL + s),
                          kc * 128 + hf * 64, tile * 128);
          }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      mbar_wait(BAR(BAR_WFULL), 0, 12);
      int job = 0, ac0 = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ac0 += KC) {
        for (int nb = 0; nb < NB; ++nb, ++job) {
          const int slot = job & 3;
          mbar_wait(BAR(BAR_ACCEMPTY + slot), ((job >> 2) & 1) ^ 1, 13);
          tc_fence_after();
          for (int kc = 0; kc < KC; ++kc) {
            const int ac = ac0 + kc, s = ac % NA;
            if (nb == 0) { mbar_wait(BAR(BAR_AFULL + s), (ac / NA) & 1, 14); tc_fence_after(); }
            const uint32_t a_addr = smem_base + off_a + s * BLK_BYTES;
            const uint32_t w_addr = smem_base + off_w + (uint32_t)(nb * KC + kc) * BLK_BYTES;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ss(tmem_base + slot * 128, desc_kmajor_sw128(a_addr + hf * HALF_BYTES + kk * 32),
                        desc_kmajor_sw128(w_addr + hf * HALF_BYTES + kk * 32), IDESC,
                        (kc | hf | kk) != 0);
            if (nb == NB - 1) umma_commit(BAR(BAR_AEMPTY + s));   // A chunk no longer needed
          }
          umma_commit(BAR(BAR_ACCFULL + slot));
        }
      }
    }
  } else if (warp < 6) {
    // ======================= LayerNorm producers (A = LN(fp32 rows)) =======================
    if (p.ln) {
      const int pw = warp - 2;
      const float4 gam = __ldg(reinterpret_cast<const float4*>(p.ln_g) + lane);
      const float4 bet = __ldg(reinterpret_cast<const float4*>(p.ln_b) + lane);
      int n = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
        const int s = n % NA;
        mbar_wait(BAR(BAR_AEMPTY + s), ((n / NA) & 1) ^ 1, 15);
        unsigned char* a_half = smem_gen + off_a + s * BLK_BYTES + (lane >> 4) * HALF_BYTES;
        const int chunk = (lane & 15) >> 1, sub = (lane & 1) * 8;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
          float4 x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const long row = (long)tile * 128 + pw * 32 + r0 + u;
            x[u] = (row < p.M) ? __ldg(reinterpret_cast<const float4*>(p.a_f32 + row * 128) + lane)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float sum = (x[u].x + x[u].y) + (x[u].z + x[u].w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum * (1.0f / 128.0f);
            const float dx = x[u].x - mean, dy = x[u].y - mean, dz = x[u].z - mean, dw = x[u].w - mean;
            float sq = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = rsqrtf(sq * (1.0f / 128.0f) + LN_EPS);
            const int r = pw * 32 + r0 + u;
            uint2 pk;
            pk.x = pack_bf16(dx * rstd * gam.x + bet.x, dy * rstd * gam.y + bet.y);
            pk.y = pack_bf16(dz * rstd * gam.z + bet.z, dw * rstd * gam.w + bet.w);
            *reinterpret_cast<uint2*>(a_half + sw128_offset(r, chunk) + sub) = pk;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(BAR(BAR_AFULL + s));
      }
    }
  } else {
    // ======================= epilogue =======================
    const int ew = warp - 6;                         // 0..3
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                   // row inside the tile
    const int et = threadIdx.x - 6 * 32;             // 0..127
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    unsigned char* stg = smem_gen + off_stg;
    int job = 0, unit = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long grow = (long)tile * 128 + row;
      for (int nb = 0; nb < NB; ++nb, ++job) {
        const int slot = job & 3;
        mbar_wait(BAR(BAR_ACCFULL + slot), (job >> 2) & 1, 16);
        tc_fence_after();
        const uint32_t tacc = tmem_base + lane_addr + slot * 128;
        const CUtensorMap* om = (p.out_split && nb == 1) ? &tm_o1 : (p.out_split && nb == 2) ? &tm_o2 : &tm_o0;
        const int ocol0 = p.out_split ? 0 : nb * 128;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(tacc + cb * 32, v);
          tmem_ld_wait();
          if (cb == 3) { tc_fence_before(); mbar_arrive(BAR(BAR_ACCEMPTY + slot)); }   // slot drained
          float f[32];
          {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + nb * 128 + cb * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = __ldg(bp + i);               // warp-uniform address: one transaction
              f[4 * i] = __uint_as_float(v[4 * i]) + b4.x;
              f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
              f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z;
              f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (p.residual && grow < p.M) {
            const float4* rp = reinterpret_cast<const float4*>(p.residual + grow * 128 + cb * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 r4 = rp[i];   // plain load: the buffer is also this kernel's output
              f[4 * i] += r4.x; f[4 * i + 1] += r4.y; f[4 * i + 2] += r4.z; f[4 * i + 3] += r4.w;
            }
          }
          if (p.out_f32) {
            // one store unit = 32 fp32 columns (128 B per row)
            unsigned char* sb = stg + (unit & 1) * STG_BYTES;
            if (et == 0) tma_store_wait_read<1>();
            named_bar_sync(1, 128);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<float4*>(sb + sw128_offset(row, c)) =
                  make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (et == 0) {
              tma_store_2d(om, smem_u32(sb), ocol0 + cb * 32, tile * 128);
              tma_store_commit();
            }
            ++unit;
          } else {
            // one store unit = 64 bf16 columns (128 B per row) = two 32-column accumulator chunks
            unsigned char* sb = stg + (unit & 1) * STG_BYTES;
            if ((cb & 1) == 0) {
              if (et == 0) tma_store_wait_read<1>();
              named_bar_sync(1, 128);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(sb + sw128_offset(row, (cb & 1) * 4 + c)) =
                  make_uint4(pack_bf16(f[8 * c], f[8 * c + 1]), pack_bf16(f[8 * c + 2], f[8 * c + 3]),
                             pack_bf16(f[8 * c + 4], f[8 * c + 5]), pack_bf16(f[8 * c + 6], f[8 * c + 7]));
            if (cb & 1) {
              fence_proxy_async_smem();
              named_bar_sync(1, 128);
              if (et == 0) {
                tma_store_2d(om, smem_u32(sb), ocol0 + (cb >> 1) * 64, tile * 128);
                tma_store_commit();
              }
              ++unit;
            }
          }
        }
      }
    }
    if (et == 0) tma_store_wait_all();
    (void)ew;
  }

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
  return cuTensorMapEncodeTiled(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

cudaError_t launch_gemm_tc(const GemmTcArgs& a, int num_sms, cudaStream_t s, std::string* err) {
  if (a.M <= 0) return cudaSuccess;
  auto bad = [&](const char* m) { if (err) *err = m; return cudaErrorInvalidValue; };
  if (a.N % 128 || a.K % 128 || a.N > 512 || (long)a.N * a.K > 65536) return bad("unsupported N/K");
  if ((a.N > 128) && (a.K > 128)) return bad("N > 128 requires K == 128");
  if (a.ln_g && a.K != 128) return bad("LayerNorm prologue needs K == 128");
  if (a.residual && a.N != 128) return bad("residual needs N == 128");
  GemmTcParams p = {};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.ln = a.ln_g != nullptr; p.a_f32 = a.a_f32; p.ln_g = a.ln_g; p.ln_b = a.ln_b;
  p.bias = a.bias; p.relu = a.relu; p.residual = a.residual;
  p.out_f32 = a.out_f32; p.out_split = a.out[1] != nullptr;
  const uint32_t w_bytes = (uint32_t)(a.N / 128) * (a.K / 128) * BLK_BYTES;
  const uint32_t fixed = w_bytes + 2 * STG_BYTES + 256 + 1024;
  p.n_a_stages = (fixed + 3 * BLK_BYTES <= 232448u) ? 3 : 2;
  const uint32_t smem = fixed + p.n_a_stages * BLK_BYTES;
  if (smem > 232448u) return bad("shared memory budget exceeded");

  CUtensorMap tw, ta, to[3];
  CUresult r = make_tmap_2d(&tw, a.w_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.N, a.K, 64, 128);
  if (r == CUDA_SUCCESS)
    r = p.ln ? CUDA_SUCCESS
             : make_tmap_2d(&ta, a.a_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, a.K, 64, 128);
  if (p.ln) ta = tw;
  const int ocols = p.out_split ? 128 : a.N;
  for (int j = 0; j < 3 && r == CUDA_SUCCESS; ++j) {
    void* optr = a.out[j] ? a.out[j] : a.out[0];
    r = a.out_f32 ? make_tmap_2d(&to[j], optr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, ocols, 32, 128)
                  : make_tmap_2d(&to[j], optr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.M, ocols, 64, 128);
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int n_tiles = (a.M + 127) / 128;
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  gemm_tc_kernel<<<grid, NTHREADS, smem, s>>>(tw, ta, to[0], to[1], to[2], p);
  return cudaGetLastError();
}

}  // namespace vadb
