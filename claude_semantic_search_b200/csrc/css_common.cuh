// css_common.cuh -- shared host/device helpers for libcss_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cfloat>
#include <climits>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>

#include "../../include/css_b200.h"

namespace css {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;  // kernels launched by this library

#define CSS_CUDA(expr)                                                         \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      css::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                     __FILE__, __LINE__);                                      \
      return (_e == cudaErrorMemoryAllocation) ? CSS_ERR_OOM : CSS_ERR_CUDA;   \
    }                                                                          \
  } while (0)

#define CSS_CHECK(expr)                \
  do {                                 \
    int _s = (expr);                   \
    if (_s != CSS_OK) return _s;       \
  } while (0)

#define CSS_REQUIRE(cond, ...)         \
  do {                                 \
    if (!(cond)) {                     \
      css::set_error(__VA_ARGS__);     \
      return CSS_ERR_INVALID;          \
    }                                  \
  } while (0)

// Launch-error check + launch accounting; used right after every <<< >>>.
#define CSS_LAUNCHED()                                                         \
  do {                                                                         \
    css::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess) {                                                   \
      css::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                      \
      return CSS_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

// Makes `device` current for the scope, restores the previous one after.
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = (cudaSetDevice(device) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Process-wide switches, initialised from the environment (CSS_SCAN_BF16, CSS_SCAN_INTERLEAVE, CSS_SCAN_LIST,
// CSS_SCAN_ADAPTIVE) and changeable at run time through css_set_option (benchmarks compare the paths in one run).
struct Options {
  std::atomic<int> scan_bf16{1};        // two-phase batch-1 scan (0: single fp32 sweep)
  std::atomic<int> scan_int8{1};        // int8 shadow rows as the first tier of the two-phase scan (0: bf16 shadow only)
  std::atomic<int> scan_interleave{1};  // dense bf16 sweep deals 8-row units block-cyclically
  std::atomic<int> scan_list{0};        // per-block list length 32 | 64 (0: by k)
  std::atomic<int> scan_adaptive{1};    // bypass phase 1 while most queries cannot be proven
  std::atomic<int> scan_mapped{1};      // single-query host calls: results + completion flag written to mapped host memory
  std::atomic<int> scan_pdl{1};         // fp32-fallback launch as a programmatic dependent of the bf16 sweep
};
Options& options();

int ensure_device(int device);  // CSS_OK iff `device` is an sm_100 GPU
int sm_count(int device);

// ---- device helpers ---------------------------------------------------------
// Streaming 128-bit load: read-only path, do not allocate in L1 (corpus rows are
// touched once per scan).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Total order used everywhere a top-k is formed: higher key first, ties by
// lower id.  (key = score for inner product, -distance for L2.)
__device__ __forceinline__ bool better(float ka, int ia, float kb, int ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

constexpr int kEmptyId = INT_MAX;  // id of an unfilled top-k slot inside kernels

}  // namespace css
