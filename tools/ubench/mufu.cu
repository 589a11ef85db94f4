// micro-benchmark: MUFU.EX2 / FFMA issue rates per SMSP on B200 (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i + 1);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[(i + 8) & 15])); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 256;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      double per = (double)*cyc / (iters * 16.0);
      printf("warps/SM=%2d (%d per SMSP) mode=%s: %.2f cycles per warp-instr per warp -> %.2f cycles per instr per SMSP\n", warps,
             warps / 4, mode == 0 ? "MUFU.EX2" : mode == 1 ? "FFMA" : "MUFU+FFMA", per, per / (warps / 4.0));
    }
  }
  return 0;
}
