// micro-benchmark (developer tool): cp.async.bulk shared -> global store throughput per SM on B200, with all SMs
// storing at once vs a single SM, for the chunk sizes the fused layer-tail kernel uses.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) store_kernel(char* dst, uint32_t chunk, int iters, int mode, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0 && threadIdx.x == 0) {            // bulk stores from one thread, at most 2 groups in flight
    for (int i = 0; i < iters; ++i) {
      char* d = dst + ((size_t)blockIdx.x * iters + i) * chunk;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(smem_u32(smem)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (mode == 1) {                         // plain coalesced st.global.v4 from 128 threads
    for (int i = 0; i < iters; ++i) {
      float4* d = reinterpret_cast<float4*>(dst + ((size_t)blockIdx.x * iters + i) * chunk);
      for (uint32_t k = threadIdx.x; k < chunk / 16; k += blockDim.x) d[k] = reinterpret_cast<float4*>(smem)[k];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  char* dst; long long* cyc; cudaMalloc(&dst, (size_t)4 << 30); cudaMallocManaged(&cyc, 8);
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int mode = 0; mode < 2; ++mode)
    for (uint32_t chunk : {65536u, 16384u, 2048u})
      for (int grid : {sms, 1}) {
        const int iters = (int)(((size_t)16 << 20) / chunk);      // 16 MB per CTA
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        store_kernel<<<grid, 128, 65536>>>(dst, chunk, iters, mode, cyc);
        cudaEventRecord(e0);
        store_kernel<<<grid, 128, 65536>>>(dst, chunk, iters, mode, cyc);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s chunk=%6u grid=%3d: %8.1f us  %6.2f TB/s  %6.1f B/clk/SM\n", mode ? "st.global.v4" : "bulk store  ", chunk, grid, ms * 1e3,
               (double)grid * iters * chunk / (ms * 1e-3) / 1e12, (double)iters * chunk / (double)*cyc);
      }
  return 0;
}
