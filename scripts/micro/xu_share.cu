// Microbenchmark: do MUFU.EX2 and F2FP.BF16.PACK_AB (cvt.rn.bf16x2.f32) share the XU pipe on sm_100a?
// Per iteration and warp: 8 ex2 (MODE bit 0), 4 cvt.bf16x2 (bit 1), 4 integer round-and-pack (2 add + prmt, bit 2).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE & 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(a[i]));
      else asm volatile("add.f32 %0, %1, 0f3F800000;" : "=f"(e[i]) : "f"(a[i]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t r = 0;
      if (MODE & 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e[2 * i + 1]), "f"(e[2 * i]));
      if (MODE & 4) {
        uint32_t x, y;
        asm volatile("add.u32 %0, %1, 0x8000;" : "=r"(x) : "r"(__float_as_uint(e[2 * i])));
        asm volatile("add.u32 %0, %1, 0x8000;" : "=r"(y) : "r"(__float_as_uint(e[2 * i + 1])));
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(x), "r"(y));
      }
      acc ^= r;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * 0.999f;
  }
  long long t1 = clock64();
  float s = __uint_as_float(acc);
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
long long run(int warps, float* out, long long* cyc, int iters) {
  long long c;
  k<MODE><<<148, warps * 32>>>(out, iters, cyc);
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  return c;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps : {4, 8, 16}) {
    const double n = (double)iters * warps;  // iterations per SM
    printf("warps/SM %2d: clk per warp-iteration per SM:  none %.1f  ex2x8 %.1f  cvtx4 %.1f  ex2x8+cvtx4 %.1f  intpackx4 %.1f  ex2x8+intpackx4 %.1f\n",
           warps, run<0>(warps, out, cyc, iters) / n, run<1>(warps, out, cyc, iters) / n, run<2>(warps, out, cyc, iters) / n,
           run<3>(warps, out, cyc, iters) / n, run<4>(warps, out, cyc, iters) / n, run<5>(warps, out, cyc, iters) / n);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
