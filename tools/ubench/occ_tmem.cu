// does using tensor memory (tcgen05.alloc) limit a kernel to one CTA per SM?  (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void __launch_bounds__(128, 4) k_plain(float* o) { o[threadIdx.x] = 1.f; }
__global__ void __launch_bounds__(128, 4) k_tmem(float* o, int cols) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  o[blockIdx.x] = (float)slot;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(cols) : "memory");
}
__global__ void __launch_bounds__(128, 4) k_bulk(float* o, const float* in) {
  __shared__ __align__(128) float buf[256];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(b) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];" ::"r"((uint32_t)__cvta_generic_to_shared(buf)), "l"(in), "r"(b) : "memory");
  }
  __syncthreads();
  o[threadIdx.x] = buf[threadIdx.x];
}
int main() {
  int occ;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_plain, 128, 0); printf("plain kernel: %d CTAs/SM\n", occ);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tmem, 128, 0); printf("tcgen05.alloc kernel: %d CTAs/SM (occupancy API)\n", occ);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bulk, 128, 0); printf("cp.async.bulk kernel: %d CTAs/SM\n", occ);
  // actually run 296 CTAs of the tmem kernel, each holding 128 columns for ~1 ms: 2 waves if one CTA per SM
  float* o; cudaMalloc(&o, 4096 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {148, 296, 592}) {
    cudaEventRecord(e0); k_tmem<<<grid, 128>>>(o, 128); cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("tmem kernel grid %d: %.2f ms (%s)\n", grid, ms, cudaGetErrorString(e));
  }
  return 0;
}
