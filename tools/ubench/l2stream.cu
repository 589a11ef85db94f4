// micro-benchmark (developer tool): how fast can every SM of a B200 pull blocks into shared memory by TMA
// (cp.async.bulk) while nothing consumes them?  Sizes the weight streaming of the fused per-layer kernel.
//   mode 0  all CTAs stream the SAME `wbytes` region (weights: L2-resident, identical addresses everywhere)
//   mode 1  every CTA streams its OWN region of `wbytes` (L2-resident after the first pass, no sharing)
//   mode 2  every CTA streams distinct data from a buffer far larger than L2 (HBM-bound reference)
//   mode 3  mode 0 + mode 2 interleaved 3:2 (weights : activations, the fused kernel's mix)
//   mode 4  mode 0 in clusters of 2 with multicast: each CTA issues half of every block for both
// Output: us per launch, aggregate TB/s, bytes per clock per SM (at the clock the kernel measured).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}

constexpr int MAXST = 6;

// one producer thread (thread 0) and one consumer thread (thread 32) per CTA
template <int MODE>
__global__ void __launch_bounds__(64, 1) stream_kernel(const char* w, const char* big, size_t big_bytes, uint32_t wbytes, uint32_t blk, int stages, int iters, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar0 = base + (uint32_t)stages * blk;
  uint32_t crank = 0;
  if (MODE == 4) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 64 + 8 * s, MODE == 4 ? 2 : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE == 4) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const int nblk = wbytes / blk;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    size_t boff = ((size_t)blockIdx.x * 7919u * blk) % (big_bytes - (size_t)blk);
    boff &= ~(size_t)1023;
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mbar_wait(bar0 + 64 + 8 * s, ((i / stages) & 1) ^ 1);
      mbar_expect(bar0 + 8 * s, blk);
      const char* src;
      bool from_big = MODE == 2 || (MODE == 3 && (i % 5) >= 3);
      if (from_big) { src = big + boff; boff += (size_t)gridDim.x * blk; if (boff + blk > big_bytes) boff = ((size_t)blockIdx.x * blk) & ~(size_t)1023; }
      else if (MODE == 1) src = w + (size_t)blockIdx.x * wbytes + (size_t)(i % nblk) * blk;
      else src = w + (size_t)(i % nblk) * blk;
      if (MODE == 4) {
        // this CTA fetches its half of the block and multicasts it to both CTAs of the pair
        const uint32_t half = blk / 2;
        bulk_load_mc(base + s * blk + crank * half, src + crank * half, half, bar0 + 8 * s, (uint16_t)3);
      } else {
        bulk_load(base + s * blk, src, blk, bar0 + 8 * s);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mbar_wait(bar0 + 8 * s, (i / stages) & 1);
      if (MODE == 4) { mbar_arrive_cluster(bar0 + 64 + 8 * s, 0); mbar_arrive_cluster(bar0 + 64 + 8 * s, 1); }
      else mbar_arrive(bar0 + 64 + 8 * s);
    }
  }
  __syncthreads();
  if (MODE == 4) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
}

template <int MODE>
float run(const char* w, const char* big, size_t big_bytes, uint32_t wbytes, uint32_t blk, int stages, int iters, int grid, long long* cyc) {
  const size_t smem = (size_t)stages * blk + 256;
  cudaFuncSetAttribute(stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = MODE == 4 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < 2; ++r) cudaLaunchKernelEx(&cfg, stream_kernel<MODE>, w, big, big_bytes, wbytes, blk, stages, iters, cyc);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int r = 0; r < reps; ++r) cudaLaunchKernelEx(&cfg, stream_kernel<MODE>, w, big, big_bytes, wbytes, blk, stages, iters, cyc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t wbytes = 384 * 1024;
  const size_t big_bytes = (size_t)2 << 30;
  char *w, *big; long long* cyc;
  cudaMalloc(&w, (size_t)sms * wbytes); cudaMalloc(&big, big_bytes); cudaMallocManaged(&cyc, 8);
  cudaMemset(w, 1, (size_t)sms * wbytes); cudaMemset(big, 1, big_bytes);
  const int iters = 12 * 28;     // 28 passes over 12 blocks of 32 KB = 10.5 MB per CTA
  const char* names[5] = {"same region (weights)", "own region (L2)", "HBM stream", "weights:HBM 3:2", "weights, cluster-2 multicast"};
  for (uint32_t blk : {32768u, 16384u}) {
    for (int stages : {2, 3, 4, 6}) {
      for (int mode = 0; mode < 5; ++mode) {
        const int it = iters * (32768 / blk);
        float ms = 0;
        if (mode == 0) ms = run<0>(w, big, big_bytes, wbytes, blk, stages, it, sms, cyc);
        if (mode == 1) ms = run<1>(w, big, big_bytes, wbytes, blk, stages, it, sms, cyc);
        if (mode == 2) ms = run<2>(w, big, big_bytes, wbytes, blk, stages, it, sms, cyc);
        if (mode == 3) ms = run<3>(w, big, big_bytes, wbytes, blk, stages, it, sms, cyc);
        if (mode == 4) ms = run<4>(w, big, big_bytes, wbytes, blk, stages, it, sms & ~1, cyc);
        const double bytes = (double)it * blk * (mode == 4 ? (sms & ~1) : sms);
        printf("blk=%5u stages=%d %-30s: %8.1f us  %6.2f TB/s into smem  %6.1f B/clk/SM (cta0: %lld cycles -> %.0f MHz)\n", blk, stages, names[mode],
               ms * 1e3, bytes / (ms * 1e-3) / 1e12, (double)it * blk / (double)*cyc, *cyc, (double)*cyc / (ms * 1e3));
      }
    }
  }
  return 0;
}
