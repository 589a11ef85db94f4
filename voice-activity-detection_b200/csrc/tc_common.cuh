// sm_100a building blocks shared by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / st / fences) and UMMA shared-memory descriptors.
// Raw PTX on purpose: this library targets B200 only (no multi-arch dispatch, no CUTLASS dependency).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vadb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (mbarrier.try_wait may suspend the thread for a system-defined time).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
// The slow path (clock reads, report, trap) is kept out of line so that the hot loops that wait on
// barriers stay compact in the instruction cache.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  bool reported = false;
  while (!mbar_try_wait(bar, parity)) {
    const long long dt = clock64() - t0;
    if (dt > 2000000000LL && !reported) {   // ~1 s: report every blocked waiter ...
      printf("vadb: mbarrier wait timeout (tag %d, block %d, thread %d, parity %u)\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      reported = true;
    }
    if (dt > 5000000000LL) __trap();        // ... then fail the launch instead of hanging the GPU
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;   // a second bounded hardware wait before leaving the hot path
  mbar_wait_slow(bar, parity, tag);
}

// ------------------------------------------------------------------ proxy / tcgen05 fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// Asynchronous prefetch of a contiguous global range into L2 (no registers, no completion tracking):
// lets register-staged producer loops run against L2 latency instead of HBM latency.
__device__ __forceinline__ void l2_prefetch(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------ TMEM allocation (one full warp)
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(smem_result_addr), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ------------------------------------------------------------------ UMMA
// Shared-memory matrix descriptor (64-bit): start address [0,14) (>>4), leading byte offset
// [16,30) (>>4), stride byte offset [32,46) (>>4), version=1 [46,48), layout type [61,64)
// (2 = SWIZZLE_128B).  Tiles are 1024-byte aligned so base_offset [49,52) stays 0.
//
// K-major, 128B swizzle (rows of 64 bf16 = 128 B; 8-row groups of 1024 B): SBO = 1024, LBO unused.
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major, 128B swizzle: 64 MN-elements (128 B) contiguous per K row, 8 K rows per 1024 B atom;
// SBO = byte stride between 8-row K groups, LBO = byte stride between 64-element MN chunks.
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                       uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32 accumulate:
// c_format [4,6)=1 (F32), a_format [7,10)=1 (BF16), b_format [10,13)=1, a_major bit 15,
// b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the descriptors passed as 32-bit halves: lo = (addr >> 4) | (LBO >> 4) << 16 differs per
// operand, hi (SBO, version, swizzle mode) is shared.  Keeps the per-MMA issue cost to an add or two.
__device__ __forceinline__ void umma_ss_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32: both operands are fp32 words in shared memory (K-major, 128B swizzle: 32 elements per 128-byte
// row), of which the tensor core uses the 19 high bits; K = 8 per instruction (32 bytes, the same
// descriptor advance as 16 bf16).  Used by the front end so that fp32 log-mel rows go from HBM to the MMA by
// TMA alone, with no conversion pass.
__device__ __forceinline__ void umma_ss_lh_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor for kind::tf32 (tf32 x tf32 -> fp32): a_format = b_format = 2 (TF32), both K-major.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows on the 128 TMEM lanes, K = 16 bf16
// packed two per 32-bit column -> 8 columns) is read from tensor memory, so it costs no shared-memory
// bandwidth.  Used for P V with P written by the softmax threads straight into TMEM.
__device__ __forceinline__ void umma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// hi word of a 128B-swizzled descriptor with SBO = 1024 B: SBO>>4 at [0,14), version 1 at bit 14,
// layout SWIZZLE_128B (2) at [29,32)
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
// One lane of a converged warp; the compiler treats the guarded code as warp-uniform.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(pred));
  return pred != 0;
}

// Arrive on an mbarrier when all tcgen05 ops previously issued by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------ TMEM <-> registers
// 32x32b: lane i of the warp accesses TMEM lane (32*(warp%4) + i), 32 consecutive columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15])
      : "memory");
}
// Same from a pointer into a register array (indices are compile-time constants after inlining).
__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows x 128 B] tile stored
// with the 128-byte swizzle TMA / UMMA use (Swizzle<3,4,3>: chunk index XOR (row mod 8)).
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}

// 2^x on the SFU (MUFU.EX2); 2^-inf = 0, no range reduction needed for softmax arguments
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Same, but pinned in program order relative to other volatile asm (named barriers): used where the
// exponentials of two warpgroups are deliberately serialised on the SFU.
__device__ __forceinline__ float fast_exp2_pinned(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe (no SFU): round-to-nearest split x = n + f with the 1.5*2^23 magic constant,
// degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5), exponent inserted
// by an integer add.  Arguments below -125 are clamped (result ~2e-38, i.e. zero for softmax).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05517197400331497f, 0.2426111400127411f);
  p = fmaf(p, f, 0.693260908126831f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

// Write a warp's staged [32 rows x 128 B] slab (128B-swizzled, see sw128_offset) to global memory with
// fully coalesced 16-byte stores: each instruction covers 4 rows x 128 B = 4 whole lines.  Plain stores
// retire without a completion wait, so one slab per warp is enough.
__device__ __forceinline__ void store_slab(const unsigned char* stg, unsigned char* dst, long row0, long M,
                                           long pitch_bytes, int col_byte_off, int lane) {
  __syncwarp();                                   // every lane's row is in the slab
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3), c = lane & 7;
    const uint4 val = *reinterpret_cast<const uint4*>(stg + sw128_offset(r, c));
    if (row0 + r < M)
      *reinterpret_cast<uint4*>(dst + (row0 + r) * pitch_bytes + col_byte_off + c * 16) = val;
  }
  __syncwarp();                                   // slab reusable
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&h);
}

// Same packing on the integer ALU: cvt.rn.bf16x2.f32 (F2FP) issues once per ~9 cycles per SM sub-partition on
// B200 (tools/ubench/explab.cu), two IADDs and a PRMT issue at full rate.  Rounds to nearest with ties
// away from zero (cvt rounds ties to even): the results differ only on exact ties.
__device__ __forceinline__ uint32_t pack_bf16_alu(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

}  // namespace tc

// Host side.  cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint) so that
// libvadb200.so has no link-time dependency on libcuda.so.1 and still loads on a machine without a driver.
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base,
                      const cuuint64_t* gdim, const cuuint64_t* gstride, const cuuint32_t* box,
                      const cuuint32_t* estr, CUtensorMapSwizzle swz);

// bf16 [B, T, 128] tensor viewed as 3-D {128, T, B}; box {64 cols, box_rows, 1},
// 128-byte swizzle, out-of-bounds rows (t >= T) read as zeros.
CUresult make_tmap_bt128(CUtensorMap* map, const void* base, int B, int T, int box_rows);

}  // namespace vadb
