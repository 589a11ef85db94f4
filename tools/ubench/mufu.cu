// micro-benchmark: MUFU.EX2 / FFMA issue rates per SMSP on B200 (developer tool)
//   mode 0  ex2.approx.ftz.f32            mode 3  ex2.approx.ftz.bf16x2 (two results per instruction?)
//   mode 1  fma.rn.f32                    mode 4  ex2.approx.f16x2
//   mode 2  MUFU + FFMA interleaved       mode 5  3 MUFU.EX2 : 1 polynomial exp2 (8 FMA-pipe instrs)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05517197400331497f, 0.2426111400127411f);
  p = fmaf(p, f, 0.693260908126831f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
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
      if (MODE == 3) { unsigned u = __float_as_uint(x[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      if (MODE == 4) { unsigned u = __float_as_uint(x[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      if (MODE == 5) { if ((i & 3) == 3) x[i] = poly_exp2(x[i]) - 1.5f; else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
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
  const char* names[6] = {"MUFU.EX2", "FFMA", "MUFU+FFMA", "EX2.bf16x2", "EX2.f16x2", "3 MUFU : 1 poly"};
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int mode = 0; mode < 6; ++mode) {
      if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 3) k<3><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 4) k<4><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 5) k<5><<<1, warps * 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      double per = (double)*cyc / (iters * 16.0);
      printf("warps/SM=%2d (%d per SMSP) mode=%s: %.2f cycles per warp-instr per warp -> %.2f cycles per instr per SMSP\n", warps,
             warps / 4, names[mode], per, per / (warps / 4.0));
    }
  }
  return 0;
}
