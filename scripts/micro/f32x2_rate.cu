// Microbenchmark: issue rate of the packed fp32x2 instructions (FFMA2 / FADD2 / FMUL2) against scalar FFMA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  uint64_t a[8];
  float s[16];
  for (int i = 0; i < 8; ++i) a[i] = 0x3f8000003f800000ull + threadIdx.x + i;
  for (int i = 0; i < 16; ++i) s[i] = 1.f + 1e-3f * (threadIdx.x + i);
  const uint64_t c = 0x3f8000013f800001ull;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(a[i]) : "l"(c));
      if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c));
      if (MODE == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c));
      if (MODE == 3) { asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(s[2 * i]) : "f"(1.0001f)); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(s[2 * i + 1]) : "f"(1.0001f)); }
    }
  }
  long long t1 = clock64();
  float r = 0; for (int i = 0; i < 8; ++i) r += (float)(a[i] & 0xff) + s[2 * i] + s[2 * i + 1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    long long c[4];
    k<0><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&c[0], cyc, 8, cudaMemcpyDeviceToHost);
    k<1><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&c[1], cyc, 8, cudaMemcpyDeviceToHost);
    k<2><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&c[2], cyc, 8, cudaMemcpyDeviceToHost);
    k<3><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&c[3], cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * warps;  // packed warp-instructions (or scalar pairs) per SM
    printf("warps/SM %2d: SM clk per warp-instr: fma.f32x2 %.3f  add.f32x2 %.3f  mul.f32x2 %.3f  | 2 x fma.f32 %.3f\n", warps,
           c[0] / n, c[1] / n, c[2] / n, c[3] / n);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
