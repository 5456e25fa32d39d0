// Microbenchmark: MUFU.EX2 / FFMA2 / F2FP issue rates per SM (profiling aid, not product code).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
      if (MODE == 1) { asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); }
      if (MODE == 2) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(a[i])); a[i] = __uint_as_float(r); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    long long c[3];
    k<0><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&c[0], cyc, 8, cudaMemcpyDeviceToHost);
    k<1><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&c[1], cyc, 8, cudaMemcpyDeviceToHost);
    k<2><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&c[2], cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * warps;  // warp-instructions per SM
    printf("warps/SM %2d: clk per warp-instr per SM  ex2 %.3f  fma %.3f  cvt.bf16x2 %.3f\n", warps, c[0] / n, c[1] / n, c[2] / n);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
