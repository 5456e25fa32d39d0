// Microbenchmark: tcgen05.ld (32x32b.x32) throughput per SM with W warps (profiling aid).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* cyc, int depth) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[32];
    const uint32_t col = (uint32_t)(((it * 4 + (warp >> 2)) * 32) & 511) & ~31u;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(base + col) : "memory");
    if ((it % depth) == depth - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(v[it & 31] & 0x3f800000u);
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int depth : {1, 2}) for (int warps : {4, 8, 16}) {
    long long c;
    k<<<148, warps * 32>>>(out, iters, cyc, depth);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * warps * 4096;   // per SM
    printf("warps/SM %2d, %d loads per wait: %.1f B/clk per SM (%.0f clk per 32x32 load per warp)\n", warps, depth, bytes / c, (double)c / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
