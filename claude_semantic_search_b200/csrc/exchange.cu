// exchange.cu -- result exchange between the shards of one search (the one exchange step
// of the row-sharded search), host side.  The device side is emit_topk() in index_kernels.cuh: the CTA
// that finishes a query's local top-k stores it into every peer's receive area over NVLink, raises a
// flag there, waits for the lists of all peers in its own memory and merges them -- inside the scan
// kernel, no NCCL call and no extra launch.
//
// One css_exchange per GPU.  Its receive area + flags are ONE cudaMalloc block, so one CUDA IPC
// handle (64 bytes) per rank is all the ranks of a torchrun job have to exchange (any all-gather of
// bytes does; claude_semantic_search_b200/sharded.py uses torch.distributed).  Inside one process
// (css_index_create_sharded) the peers are wired with plain peer pointers.
//
// Every rank must issue the same sequence of exchange searches (collective semantics): the epoch
// counter advances by one per call on every rank, and the flags carry it.
#include "index_internal.h"

using namespace css;

namespace {

constexpr int kExMaxNq = 64;

size_t flags_bytes(int n_ranks) { return (size_t)2 * kExMaxNq * n_ranks * sizeof(unsigned); }
size_t flags_bytes_padded(int n_ranks) { return (flags_bytes(n_ranks) + 255) / 256 * 256; }
size_t slots_bytes(int n_ranks) { return (size_t)2 * kExMaxNq * n_ranks * CSS_MAX_K * sizeof(ExEntry); }

}  // namespace

namespace css {

int exchange_next(css_exchange* ex, ExchangeDev* out) {
  std::lock_guard<std::mutex> lk(ex->mu);
  // a list that never arrived in an earlier search (a peer that failed before launching): loud, not silent
  // (the kernel writes the status word into mapped host memory; reading it costs nothing)
  if (ex->status_host && *ex->status_host != 0) {
    set_error("result exchange: a peer's list did not arrive within 10 s in an earlier search");
    return CSS_ERR_CUDA;
  }
  ex->epoch += 1;
  memset(out, 0, sizeof(*out));
  out->n_ranks = ex->n_ranks;
  out->rank = ex->rank;
  out->epoch = ex->epoch;
  out->max_nq = ex->max_nq;
  out->status = ex->status;
  out->no_pdl = ex->shared_device ? 1 : 0;
  for (int r = 0; r < ex->n_ranks; ++r) {
    out->slots[r] = ex->slots[r];
    out->flags[r] = ex->flags[r];
  }
  return CSS_OK;
}

int exchange_connect_local(css_exchange** exs, int n) {
  for (int a = 0; a < n; ++a) {
    DeviceGuard g(exs[a]->device);
    for (int b = 0; b < n; ++b) {
      if (a != b && exs[a]->device == exs[b]->device) exs[a]->shared_device = true;
      if (a == b || exs[a]->device == exs[b]->device) continue;
      int can = 0;
      CSS_CUDA(cudaDeviceCanAccessPeer(&can, exs[a]->device, exs[b]->device));
      if (!can) {
        set_error("device %d cannot access device %d (no peer path)", exs[a]->device, exs[b]->device);
        return CSS_ERR_UNSUPPORTED;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(exs[b]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", exs[a]->device, exs[b]->device,
                  cudaGetErrorString(e));
        return CSS_ERR_CUDA;
      }
      (void)cudaGetLastError();
    }
    for (int b = 0; b < n; ++b) {
      exs[a]->slots[b] = exs[b]->slots_local;
      exs[a]->flags[b] = exs[b]->flags_local;
    }
    exs[a]->connected = true;
  }
  return CSS_OK;
}

}  // namespace css

extern "C" {

int css_exchange_create(int device, int n_ranks, int rank, css_exchange** out, unsigned char* handle_out) {
  CSS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  CSS_REQUIRE(n_ranks >= 1 && n_ranks <= CSS_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank %d of %d (at most %d)",
              rank, n_ranks, CSS_MAX_RANKS);
  CSS_CHECK(ensure_device(device));
  DeviceGuard g(device);
  css_exchange* ex = new (std::nothrow) css_exchange();
  if (!ex) {
    set_error("out of host memory");
    return CSS_ERR_OOM;
  }
  ex->device = device;
  ex->n_ranks = n_ranks;
  ex->rank = rank;
  ex->max_nq = kExMaxNq;
  unsigned char* block = nullptr;
  const size_t fb = flags_bytes_padded(n_ranks), total = fb + slots_bytes(n_ranks);
  cudaError_t e = cudaMalloc(&block, total);
  if (e == cudaSuccess) e = cudaMemset(block, 0, total);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("exchange buffer allocation failed: %s", cudaGetErrorString(e));
    cudaFree(block);
    delete ex;
    return CSS_ERR_OOM;
  }
  ex->flags_local = reinterpret_cast<unsigned*>(block);
  void* sh = nullptr;
  void* sd = nullptr;
  if (cudaHostAlloc(&sh, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(&sd, sh, 0) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("exchange status word allocation failed");
    cudaFree(block);
    delete ex;
    return CSS_ERR_OOM;
  }
  ex->status_host = reinterpret_cast<volatile int*>(sh);
  *ex->status_host = 0;
  ex->status = reinterpret_cast<int*>(sd);
  ex->slots_local = reinterpret_cast<ExEntry*>(block + fb);
  ex->slots[rank] = ex->slots_local;
  ex->flags[rank] = ex->flags_local;
  if (handle_out) {
    cudaIpcMemHandle_t hnd;
    e = cudaIpcGetMemHandle(&hnd, block);
    if (e != cudaSuccess) {
      set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
      cudaFree(block);
      delete ex;
      return CSS_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == CSS_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &hnd, sizeof(hnd));
  }
  if (n_ranks == 1) ex->connected = true;
  *out = ex;
  return CSS_OK;
}

int css_exchange_connect(css_exchange* ex, const unsigned char* handles) {
  CSS_REQUIRE(ex != nullptr && handles != nullptr, "NULL argument");
  std::lock_guard<std::mutex> lk(ex->mu);
  CSS_REQUIRE(!ex->connected || ex->n_ranks == 1, "exchange already connected");
  DeviceGuard g(ex->device);
  const size_t fb = flags_bytes_padded(ex->n_ranks);
  for (int r = 0; r < ex->n_ranks; ++r) {
    if (r == ex->rank) continue;
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handles + (size_t)r * CSS_IPC_HANDLE_BYTES, sizeof(hnd));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
      return CSS_ERR_CUDA;
    }
    ex->opened[r] = true;
    ex->flags[r] = reinterpret_cast<unsigned*>(p);
    ex->slots[r] = reinterpret_cast<ExEntry*>(reinterpret_cast<unsigned char*>(p) + fb);
  }
  ex->connected = true;
  return CSS_OK;
}

int css_exchange_destroy(css_exchange* ex) {
  if (!ex) return CSS_OK;
  {
    DeviceGuard g(ex->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < ex->n_ranks; ++r)
      if (ex->opened[r]) cudaIpcCloseMemHandle(ex->flags[r]);
    cudaFree(ex->flags_local);
    if (ex->status_host) cudaFreeHost(const_cast<int*>(ex->status_host));
  }
  delete ex;
  return CSS_OK;
}

}  // extern "C"
