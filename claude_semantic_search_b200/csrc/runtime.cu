// runtime.cu -- error state, device probing and small shared utilities.
#include "tc_common.cuh"
#include <cstdlib>

namespace css {

static thread_local char t_error[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

Options& options() {
  static Options o;
  static const bool init = [] {
    auto env = [](const char* n, std::atomic<int>& dst) {
      const char* v = getenv(n);
      if (v) dst.store(atoi(v));
    };
    env("CSS_SCAN_BF16", o.scan_bf16);
    env("CSS_SCAN_INT8", o.scan_int8);
    env("CSS_SCAN_INTERLEAVE", o.scan_interleave);
    env("CSS_SCAN_LIST", o.scan_list);
    env("CSS_SCAN_ADAPTIVE", o.scan_adaptive);
    env("CSS_SCAN_PDL", o.scan_pdl);
    env("CSS_SCAN_MAPPED", o.scan_mapped);
    return true;
  }();
  (void)init;
  return o;
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


// ---- TMA tensor maps ---------------------------------------------------------------
// libcuda is not linked (the .so must load on a box without a driver); the encoder is
// resolved through the runtime's driver entry point query on first use.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
  static std::atomic<void*> cached{nullptr};
  void* fn = cached.load(std::memory_order_acquire);
  if (fn) return reinterpret_cast<PFN_encodeTiled>(fn);
  cudaDriverEntryPointQueryResult qres;
  void* p = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    (void)cudaGetLastError();
    return nullptr;
  }
  cached.store(p, std::memory_order_release);
  return reinterpret_cast<PFN_encodeTiled>(p);
}

int encode_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                        uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return CSS_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld_elems * 2) % 16 != 0 || box_cols * 2 != 128 ||
      box_rows > 256) {
    set_error("tensor map: unsupported alignment / box (base %p, ld %llu, box %u x %u)", base,
              (unsigned long long)ld_elems, box_rows, box_cols);
    return CSS_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu ld %llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems);
    return CSS_ERR_CUDA;
  }
  return CSS_OK;
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

extern "C" int css_set_option(const char* name, int value) {
  if (!name) return CSS_ERR_INVALID;
  css::Options& o = css::options();
  std::atomic<int>* slot = nullptr;
  if (!strcmp(name, "scan_bf16")) slot = &o.scan_bf16;
  else if (!strcmp(name, "scan_int8")) slot = &o.scan_int8;
  else if (!strcmp(name, "scan_interleave")) slot = &o.scan_interleave;
  else if (!strcmp(name, "scan_list")) slot = &o.scan_list;
  else if (!strcmp(name, "scan_adaptive")) slot = &o.scan_adaptive;
  else if (!strcmp(name, "scan_pdl")) slot = &o.scan_pdl;
  else if (!strcmp(name, "scan_mapped")) slot = &o.scan_mapped;
  if (!slot) {
    css::set_error("unknown option %s", name);
    return CSS_ERR_INVALID;
  }
  slot->store(value);
  return CSS_OK;
}
