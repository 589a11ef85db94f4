// micro-benchmark (developer tool): per-SM throughput of 2-D TMA tensor stores (shared -> global) for the slab geometries
// of the fused layer-tail kernel: a [M,128] bf16 tensor written in boxes of {32 cols = 64 B, 32 rows} (one 2 KB slab per
// epilogue warp, 64B swizzle), {64 cols = 128 B, 32 rows} (4 KB, 128B swizzle) and {64 cols, 128 rows} (16 KB), each issuing
// thread keeping one store in flight (wait_group.read 0 before reusing its slab), as the kernel does.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmastore tmastore.cu && ./tmastore
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// nthr issuing threads (one per warp), each owns a slab of box_bytes and walks its own column group / row block
__global__ void __launch_bounds__(512, 1) store_kernel(const __grid_constant__ CUtensorMap tm, int box_cols, int box_rows, int n_issuers,
                                                       int tiles, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < n_issuers && lane == 0) {
    const uint32_t slab = smem_u32(smem) + (uint32_t)warp * (uint32_t)(box_cols * box_rows * 2);
    const int col_groups = 128 / box_cols, row_groups = 128 / box_rows;       // issuers cover one 128 x 128 tile per pass
    const int cg = warp % col_groups, rg = (warp / col_groups) % row_groups;
    for (int t = 0; t < tiles; ++t) {
      const int tile = blockIdx.x * tiles + t;
      for (int rep = 0; rep < (col_groups * row_groups) / n_issuers; ++rep) {
        const int rgg = rg + rep * (n_issuers / col_groups);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(slab), "r"(cg * box_cols), "r"(tile * 128 + rgg * box_rows) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encoder\n"); return 1; }
  EncodeFn enc = (EncodeFn)p;
  const int tiles = 64;                                        // 128-row tiles per CTA: 64 x 32 KB = 2 MB per CTA per tensor
  const long M = (long)sms * tiles * 128;
  char* dst; long long* cyc; cudaMalloc(&dst, (size_t)M * 256); cudaMallocManaged(&cyc, 8);
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  struct Geo { int cols, rows, issuers; CUtensorMapSwizzle swz; const char* name; };
  const Geo geos[] = {{32, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B, "box 64B x 32 rows (2 KB), 16 issuers"},
                      {64, 32, 8, CU_TENSOR_MAP_SWIZZLE_128B, "box 128B x 32 rows (4 KB), 8 issuers"},
                      {64, 128, 2, CU_TENSOR_MAP_SWIZZLE_128B, "box 128B x 128 rows (16 KB), 2 issuers"},
                      {64, 64, 4, CU_TENSOR_MAP_SWIZZLE_128B, "box 128B x 64 rows (8 KB), 4 issuers"}};
  for (const Geo& g : geos)
    for (int grid : {sms, 1}) {
      CUtensorMap tm;
      cuuint64_t gdim[2] = {128, (cuuint64_t)M}; cuuint64_t gstr[1] = {256};
      cuuint32_t box[2] = {(cuuint32_t)g.cols, (cuuint32_t)g.rows}; cuuint32_t es[2] = {1, 1};
      if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dst, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, g.swz,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      store_kernel<<<grid, 512, 65536>>>(tm, g.cols, g.rows, g.issuers, tiles, cyc);
      cudaEventRecord(e0);
      store_kernel<<<grid, 512, 65536>>>(tm, g.cols, g.rows, g.issuers, tiles, cyc);
      cudaEventRecord(e1);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)tiles * 128 * 256;
      printf("%-42s grid=%3d: %8.1f us  %6.2f TB/s  %6.1f B/clk/SM\n", g.name, grid, ms * 1e3, grid * bytes / (ms * 1e-3) / 1e12, bytes / (double)*cyc);
    }
  return 0;
}
