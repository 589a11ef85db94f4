// micro-benchmark: the softmax exponential phase of the attention kernel in isolation (developer tool)
//   mode 0  replica of one step: 64 x (FFMA, MUFU.EX2, FADD) + 32 x cvt.rn.bf16x2.f32 + row max
//   mode 1  same, pack by PRMT of the high halves (truncation) instead of cvt
//   mode 2  cvt.rn.bf16x2.f32 throughput alone
//   mode 3  mode 0 without the row max
//   mode 4  tcgen05-free "everything but MUFU": FFMA + FADD + cvt (no ex2)
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&h);
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(const float* in, float* out, long long* cyc, int iters) {
  float sv[64];
  for (int i = 0; i < 64; ++i) sv[i] = in[threadIdx.x * 64 + i];
  float m = in[threadIdx.x], acc = 0.f;
  unsigned accu = 0;
  const float c = 0.12751743f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 64; i += 2) { accu ^= pack_bf16(sv[i] + acc, sv[i + 1]); }
      acc += 1.0f;
      continue;
    }
    float ps4[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned pk[32];
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      const float a0 = fmaf(sv[i], c, -m), a1 = fmaf(sv[i + 1], c, -m);
      const float p0 = MODE == 4 ? a0 * a0 : ex2(a0), p1 = MODE == 4 ? a1 * a1 : ex2(a1);
      ps4[(i >> 1) & 3] += p0 + p1;
      if (MODE == 1) pk[i >> 1] = __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
      else pk[i >> 1] = pack_bf16(p0, p1);
    }
    float mx = -1e30f;
    if (MODE != 3) {
#pragma unroll
      for (int i = 0; i < 64; ++i) mx = fmaxf(mx, sv[i]);
    }
    acc += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]) + mx;
#pragma unroll
    for (int i = 0; i < 32; ++i) accu ^= pk[i];
    m += 1e-6f * acc;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(accu);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 1 << 22); cudaMemset(in, 0, 1 << 22); cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 64;
  const char* names[5] = {"step replica (cvt pack)", "step replica (PRMT pack)", "cvt.rn.bf16x2 alone (32/iter)", "replica w/o row max", "no MUFU"};
  for (int warps = 4; warps <= 8; warps *= 2)
    for (int mode = 0; mode < 5; ++mode) {
      if (mode == 0) k<0><<<1, warps * 32>>>(in, out, cyc, iters);
      if (mode == 1) k<1><<<1, warps * 32>>>(in, out, cyc, iters);
      if (mode == 2) k<2><<<1, warps * 32>>>(in, out, cyc, iters);
      if (mode == 3) k<3><<<1, warps * 32>>>(in, out, cyc, iters);
      if (mode == 4) k<4><<<1, warps * 32>>>(in, out, cyc, iters);
      cudaDeviceSynchronize();
      printf("warps/SM=%d mode=%-32s: %.0f cycles per 64-element step\n", warps, names[mode], (double)*cyc / iters);
    }
  return 0;
}
