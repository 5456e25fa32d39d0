// runtime.cu -- error state, device probing and small shared utilities.
#include "css_common.cuh"

namespace css {

static thread_local char t_error[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

int ensure_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s); libcss_b200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    (void)cudaGetLastError();
    return CSS_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (have %d)", device, n);
    return CSS_ERR_INVALID;
  }
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) {
    set_error("cannot query compute capability of device %d", device);
    return CSS_ERR_CUDA;
  }
  if (major != 10) {
    set_error("device %d is sm_%d%d; libcss_b200 is built for sm_100a only", device, major, minor);
    return CSS_ERR_NO_DEVICE;
  }
  return CSS_OK;
}

int sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 148;
  return n > 0 ? n : 148;
}

}  // namespace css

extern "C" {

int css_abi_version(void) { return CSS_ABI_VERSION; }

const char* css_last_error(void) { return css::t_error; }

int css_device_count(int* n_out) {
  if (!n_out) {
    css::set_error("n_out is NULL");
    return CSS_ERR_INVALID;
  }
  *n_out = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    css::set_error("no CUDA device available (%s)",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    (void)cudaGetLastError();
    return CSS_ERR_NO_DEVICE;
  }
  int usable = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess &&
        major == 10)
      ++usable;
  }
  if (usable == 0) {
    css::set_error("%d CUDA device(s) present but none is sm_100", n);
    return CSS_ERR_NO_DEVICE;
  }
  *n_out = usable;
  return CSS_OK;
}

int css_device_info(int device, int64_t info_out[5]) {
  if (!info_out) {
    css::set_error("info_out is NULL");
    return CSS_ERR_INVALID;
  }
  CSS_CHECK(css::ensure_device(device));
  css::DeviceGuard g(device);
  size_t free_b = 0, total_b = 0;
  CSS_CUDA(cudaMemGetInfo(&free_b, &total_b));
  int major = 0, minor = 0;
  CSS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CSS_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  info_out[0] = css::sm_count(device);
  info_out[1] = (int64_t)total_b;
  info_out[2] = (int64_t)free_b;
  info_out[3] = major;
  info_out[4] = minor;
  return CSS_OK;
}

int64_t css_kernel_launch_count(void) { return css::g_launches.load(); }

}  // extern "C"
