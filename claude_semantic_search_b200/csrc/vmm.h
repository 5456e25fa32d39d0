// vmm.h -- growable device arrays on CUDA virtual memory management.
//
// The flat index is append-only and can reach tens of GB per GPU (config 4: 38.4 GB of fp32 rows +
// 19.2 GB of bf16 shadow rows per shard).  Growing a cudaMalloc'ed array means a second allocation and
// a device-to-device copy, i.e. old + new resident at once (115 GB transient for that shard).  Here an
// array reserves a virtual range once (sized for the device's HBM) and growth maps more physical
// memory behind the same addresses: no copy, no transient, pointers into the array stay valid.
// libcuda is not linked (the .so must load without a driver): entry points are resolved through
// cudaGetDriverEntryPoint on first use, like the TMA descriptor encoder.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace css {

void set_error(const char* fmt, ...);

struct VmmApi {
  CUresult (*getGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*addressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  bool ok = false;
};

inline const VmmApi& vmm_api() {
  static const VmmApi api = [] {
    VmmApi a;
    auto get = [](const char* name) -> void* {
      cudaDriverEntryPointQueryResult q;
      void* p = nullptr;
      if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
      }
      return p;
    };
    a.getGranularity = reinterpret_cast<decltype(a.getGranularity)>(get("cuMemGetAllocationGranularity"));
    a.addressReserve = reinterpret_cast<decltype(a.addressReserve)>(get("cuMemAddressReserve"));
    a.addressFree = reinterpret_cast<decltype(a.addressFree)>(get("cuMemAddressFree"));
    a.create = reinterpret_cast<decltype(a.create)>(get("cuMemCreate"));
    a.release = reinterpret_cast<decltype(a.release)>(get("cuMemRelease"));
    a.map = reinterpret_cast<decltype(a.map)>(get("cuMemMap"));
    a.unmap = reinterpret_cast<decltype(a.unmap)>(get("cuMemUnmap"));
    a.setAccess = reinterpret_cast<decltype(a.setAccess)>(get("cuMemSetAccess"));
    a.ok = a.getGranularity && a.addressReserve && a.addressFree && a.create && a.release && a.map && a.unmap &&
           a.setAccess;
    return a;
  }();
  return api;
}

struct VmmArray {
  CUdeviceptr base = 0;
  size_t reserved = 0;   // bytes of virtual range
  size_t mapped = 0;     // bytes backed by physical memory (a prefix of the range)
  size_t gran = 0;
  int device = 0;
  struct Chunk {
    CUmemGenericAllocationHandle handle;
    size_t size;
  };
  std::vector<Chunk> chunks;
  void* ptr() const { return reinterpret_cast<void*>(base); }
};

// Reserve `max_bytes` of address space on `device` (nothing is allocated yet).
inline int vmm_reserve(VmmArray* a, int device, size_t max_bytes) {
  const VmmApi& api = vmm_api();
  if (!api.ok) {
    set_error("CUDA virtual memory management is not available from this driver");
    return -3;
  }
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  size_t gran = 0;
  if (api.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) {
    set_error("cuMemGetAllocationGranularity failed");
    return -3;
  }
  const size_t bytes = (std::max<size_t>(max_bytes, 1) + gran - 1) / gran * gran;
  CUdeviceptr p = 0;
  if (api.addressReserve(&p, bytes, 0, 0, 0) != CUDA_SUCCESS) {
    set_error("cuMemAddressReserve of %zu bytes failed", bytes);
    return -4;
  }
  a->base = p;
  a->reserved = bytes;
  a->mapped = 0;
  a->gran = gran;
  a->device = device;
  return 0;
}

// Make at least `need_bytes` of the range usable.  New memory is NOT initialised.
inline int vmm_grow(VmmArray* a, size_t need_bytes) {
  if (need_bytes <= a->mapped) return 0;
  const VmmApi& api = vmm_api();
  if (need_bytes > a->reserved) {
    set_error("array would exceed its reserved range (%zu > %zu bytes)", need_bytes, a->reserved);
    return -4;
  }
  const size_t target = (need_bytes + a->gran - 1) / a->gran * a->gran;
  const size_t add = target - a->mapped;
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = a->device;
  CUmemGenericAllocationHandle hnd;
  CUresult r = api.create(&hnd, add, &prop, 0);
  if (r != CUDA_SUCCESS) {
    set_error("cuMemCreate of %zu bytes failed (%d): device memory exhausted", add, (int)r);
    return -4;
  }
  r = api.map(a->base + a->mapped, add, 0, hnd, 0);
  if (r != CUDA_SUCCESS) {
    api.release(hnd);
    set_error("cuMemMap failed (%d)", (int)r);
    return -3;
  }
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = a->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = api.setAccess(a->base + a->mapped, add, &acc, 1);
  if (r != CUDA_SUCCESS) {
    api.unmap(a->base + a->mapped, add);
    api.release(hnd);
    set_error("cuMemSetAccess failed (%d)", (int)r);
    return -3;
  }
  a->chunks.push_back({hnd, add});
  a->mapped = target;
  return 0;
}

inline void vmm_release(VmmArray* a) {
  if (!a->base) return;
  const VmmApi& api = vmm_api();
  size_t off = 0;
  for (const auto& c : a->chunks) {
    api.unmap(a->base + off, c.size);
    api.release(c.handle);
    off += c.size;
  }
  a->chunks.clear();
  api.addressFree(a->base, a->reserved);
  *a = VmmArray();
}

}  // namespace css
