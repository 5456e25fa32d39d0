// Microbenchmark: the pass-2 instruction mix of the attention kernel (rel LDS + scale/bias FMA + MUFU.EX2 + row sum +
// bf16 pack + swizzled STS.128 per key) with W warps per SM, no TMEM / barriers.  Variants (template bits):
//   1: scalar fp32 instead of packed fp32x2     2: four independent sum accumulators
//   4: rel read as aligned float4 (stands for 4 alignment-shifted copies of the table)   8: no MUFU
//   16: every second key pair takes exp2 from a degree-3 polynomial on the FMA pipe (Cody-Waite) instead of MUFU
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void un(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// exp2 of two non-positive arguments without the XU: n = round(x), 2^(x - n) by a degree-3 polynomial, exponent
// added as an integer.  max relative error 7.7e-5.
__device__ __forceinline__ void ex2_poly2(uint64_t t, float& e0, float& e1) {
  float a, b; un(t, a, b);
  t = pk(fmaxf(a, -125.f), fmaxf(b, -125.f));
  const uint64_t magic = pk(12582912.f, 12582912.f), nmagic = pk(-12582912.f, -12582912.f);
  const uint64_t xf = add2(t, magic);
  const uint64_t nf = add2(xf, nmagic);
  float n0, n1; un(nf, n0, n1);
  const uint64_t f = add2(t, pk(-n0, -n1));
  uint64_t p = fma2(f, pk(0.05508868f, 0.05508868f), pk(0.24260405f, 0.24260405f));
  p = fma2(p, f, pk(0.69327623f, 0.69327623f));
  p = fma2(p, f, pk(0.99992895f, 0.99992895f));
  float p0, p1, x0, x1; un(p, p0, p1); un(xf, x0, x1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(x0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(x1) << 23));
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  __shared__ __align__(16) float rel[2048];
  __shared__ uint4 pbuf[2][128 * 8];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) rel[i] = -0.001f * i;
  __syncthreads();
  const int r = threadIdx.x & 127;
  float v[64];
  for (int i = 0; i < 64; ++i) v[i] = -0.01f * (i + r);
  uint64_t sum2[4] = {0, 0, 0, 0};
  float sums[4] = {0, 0, 0, 0};
  const float kScale = 0.18f, nm = -1.f;
  const uint64_t scale2 = pk(kScale, kScale), nm2 = pk(nm, nm);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float* rel_c = (MODE & 4) ? rel + ((383 - r) & ~3) + (it & 3) * 64 : rel + 383 - r + (it & 3) * 64;
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      uint32_t o[4];
      float rl[8];
      if (MODE & 4) {
        const float4 a = *reinterpret_cast<const float4*>(rel_c + q4 * 8), b = *reinterpret_cast<const float4*>(rel_c + q4 * 8 + 4);
        rl[0] = a.x; rl[1] = a.y; rl[2] = a.z; rl[3] = a.w; rl[4] = b.x; rl[5] = b.y; rl[6] = b.z; rl[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) rl[j] = rel_c[q4 * 8 + j];
      }
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        const int j = q4 * 4 + j2;
        float t0f, t1f;
        if (MODE & 1) {
          t0f = fmaf(v[2 * j], kScale, rl[2 * j2] + nm);
          t1f = fmaf(v[2 * j + 1], kScale, rl[2 * j2 + 1] + nm);
        } else {
          uint64_t t = fma2(pk(v[2 * j], v[2 * j + 1]), scale2, add2(pk(rl[2 * j2], rl[2 * j2 + 1]), nm2));
          un(t, t0f, t1f);
        }
        float e0, e1;
        if (MODE & 8) { e0 = t0f * 0.5f; e1 = t1f * 0.5f; }
        else if ((MODE & 16) && (j2 & 1)) { ex2_poly2(pk(t0f, t1f), e0, e1); }
        else { e0 = ex2(t0f); e1 = ex2(t1f); }
        const int acc = (MODE & 2) ? (j2 & 3) : 0;
        if (MODE & 1) sums[acc] += e0 + e1; else sum2[acc] = add2(sum2[acc], pk(e0, e1));
        __nv_bfloat162 b = __floats2bfloat162_rn(e0, e1);
        o[j2] = *reinterpret_cast<uint32_t*>(&b);
      }
      pbuf[it & 1][r * 8 + (q4 ^ (r & 7))] = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __int_as_float(__float_as_int(v[i]) ^ (it & 1));   // "new scores" (stands for the TMEM load)
  }
  long long t1 = clock64();
  float a = 0, b;
  for (int i = 0; i < 4; ++i) { float x, y; un(sum2[i], x, y); a += x + y + sums[i]; }
  b = v[3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(float* out, long long* cyc, int iters) {
  printf("mode %2d (%s %s %s %s):", MODE, MODE & 1 ? "scalar" : "packed", MODE & 2 ? "4sum" : "1sum", MODE & 4 ? "lds128" : "lds32 ", MODE & 8 ? "noMUFU" : (MODE & 16 ? "halfPoly" : "MUFU  "));
  for (int warps : {4, 8, 16}) {
    long long c;
    k<MODE><<<148, warps * 32>>>(out, iters, cyc);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  %2d warps/SM: %5.0f", warps, (double)c / iters);
  }
  printf("   cycles per 64-key block per warp\n");
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  run<0>(out, cyc, iters); run<1>(out, cyc, iters); run<2>(out, cyc, iters); run<3>(out, cyc, iters);
  run<4>(out, cyc, iters); run<5>(out, cyc, iters); run<6>(out, cyc, iters); run<7>(out, cyc, iters);
  run<8>(out, cyc, iters); run<12>(out, cyc, iters); run<16>(out, cyc, iters); run<18>(out, cyc, iters); run<20>(out, cyc, iters); run<22>(out, cyc, iters);
  printf("XU bound: 512 cycles per block per warp on the same SM sub-partition (4 warps/SM: 512, 8: 1024, 16: 2048)\n%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
