// index_io.cu -- faiss-compatible persistence of the flat index (src/storage.py:306,879-884:
// faiss.read_index / faiss.write_index), for a single-device index and for a multi-device one.
//
// faiss IndexFlat file format (faiss/impl/index_write.cpp, write_index_header + IndexFlat codes;
// restated from the published format, faiss >= 1.7):
//   u32  fourcc  "IxFI" (inner product) | "IxF2" (L2)
//   i32  d
//   i64  ntotal
//   i64  dummy (1 << 20), i64 dummy (1 << 20)
//   u8   is_trained (1)
//   i32  metric_type (0 = IP, 1 = L2)
//   u64  count = ntotal * d          (vector<float> xb / codes.size()/4)
//   f32  data[count]
//
// Both directions are double-buffered: two pinned chunks, the file I/O of chunk i+1 overlaps the
// PCIe copy of chunk i.  Loaded rows are copied straight into their place in x (no staging copy) and
// finished there (bf16 shadow, norm / rounding-error maxima).
#include "index_internal.h"

#include <algorithm>
#include <string>
#include <vector>

using namespace css;

namespace {

uint32_t fourcc(const char s[4]) {
  return (uint32_t)(unsigned char)s[0] | ((uint32_t)(unsigned char)s[1] << 8) |
         ((uint32_t)(unsigned char)s[2] << 16) | ((uint32_t)(unsigned char)s[3] << 24);
}

css_index* shard_of(css_index* h, const RowSpan& sp) { return sp.shard < 0 ? h : h->shards[sp.shard]; }

struct IoBuffers {
  void* buf[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev[2];   // one event per shard and buffer
  std::vector<char> used[2];
  size_t bytes = 0;
  int n_shards = 1;
  int init(css_index* h, size_t chunk_bytes) {
    n_shards = h->shards.empty() ? 1 : (int)h->shards.size();
    bytes = chunk_bytes;
    for (int b = 0; b < 2; ++b) {
      if (cudaMallocHost(&buf[b], chunk_bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaMallocHost(%zu) failed", chunk_bytes);
        return CSS_ERR_OOM;
      }
      ev[b].assign(n_shards, nullptr);
      used[b].assign(n_shards, 0);
    }
    return CSS_OK;
  }
  // all copies that used buffer b are complete
  int wait(css_index* h, int b) {
    for (int s = 0; s < n_shards; ++s) {
      if (!used[b][s]) continue;
      css_index* sh = h->shards.empty() ? h : h->shards[s];
      DeviceGuard g(sh->device);
      CSS_CUDA(cudaEventSynchronize(ev[b][s]));
      used[b][s] = 0;
    }
    return CSS_OK;
  }
  int mark(css_index* h, int b, int shard) {
    const int s = shard < 0 ? 0 : shard;
    css_index* sh = h->shards.empty() ? h : h->shards[s];
    if (!ev[b][s]) CSS_CUDA(cudaEventCreateWithFlags(&ev[b][s], cudaEventDisableTiming));
    CSS_CUDA(cudaEventRecord(ev[b][s], sh->stream));
    used[b][s] = 1;
    return CSS_OK;
  }
  void release(css_index* h) {
    for (int b = 0; b < 2; ++b) {
      for (int s = 0; s < (int)ev[b].size(); ++s) {
        if (!ev[b][s]) continue;
        css_index* sh = h->shards.empty() ? h : h->shards[s];
        DeviceGuard g(sh->device);
        cudaEventSynchronize(ev[b][s]);
        cudaEventDestroy(ev[b][s]);
      }
      if (buf[b]) cudaFreeHost(buf[b]);
      buf[b] = nullptr;
    }
  }
};

// rows per I/O chunk: whole shard blocks (so a chunk splits into at most a few spans), about 32 MB
int64_t chunk_rows_for(int d) {
  const int64_t rows = std::max<int64_t>(1, ((int64_t)32 << 20) / ((int64_t)d * 4));
  return std::max<int64_t>(kShardBlock, rows / kShardBlock * kShardBlock);
}

}  // namespace

extern "C" {

int css_index_save(css_index* h, const char* path) {
  CSS_REQUIRE(h != nullptr && path != nullptr, "NULL argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const bool composite = !h->shards.empty();
  const int64_t ntotal = composite ? h->composite_ntotal : h->ntotal;
  std::string tmp = std::string(path) + ".tmp";
  FILE* fp = fopen(tmp.c_str(), "wb");
  if (!fp) {
    set_error("cannot open %s for writing", tmp.c_str());
    return CSS_ERR_IO;
  }
  bool ok = true;
  auto W = [&](const void* p, size_t n) { ok = ok && (fwrite(p, 1, n, fp) == n); };
  uint32_t cc = fourcc(h->metric == CSS_METRIC_INNER_PRODUCT ? "IxFI" : "IxF2");
  int32_t d = h->dim;
  int64_t nt = ntotal, dummy = (int64_t)1 << 20;
  uint8_t trained = 1;
  int32_t metric = h->metric;
  uint64_t count = (uint64_t)ntotal * (uint64_t)h->dim;
  W(&cc, 4); W(&d, 4); W(&nt, 8); W(&dummy, 8); W(&dummy, 8); W(&trained, 1); W(&metric, 4);
  W(&count, 8);
  int rc = CSS_OK;
  if (ntotal > 0 && ok) {
    const int64_t chunk_rows = chunk_rows_for(d);
    IoBuffers io;
    rc = io.init(h, (size_t)chunk_rows * d * 4);
    std::vector<RowSpan> spans;
    // D2H of chunk i+1 is in flight while chunk i is written to the file
    auto fetch = [&](int64_t r0, int b) -> int {
      const int64_t nr = std::min(chunk_rows, ntotal - r0);
      spans_of(h, r0, nr, &spans);
      for (const RowSpan& sp : spans) {
        css_index* sh = shard_of(h, sp);
        DeviceGuard gs(sh->device);
        CSS_CUDA(cudaMemcpyAsync(reinterpret_cast<float*>(io.buf[b]) + (size_t)(sp.global - r0) * d,
                                 sh->x + (size_t)sp.local * d, (size_t)sp.n * d * 4, cudaMemcpyDeviceToHost, sh->stream));
        CSS_CHECK(io.mark(h, b, sp.shard));
      }
      return CSS_OK;
    };
    if (rc == CSS_OK) rc = fetch(0, 0);
    int b = 0;
    for (int64_t r0 = 0; r0 < ntotal && rc == CSS_OK && ok; r0 += chunk_rows, b ^= 1) {
      if (r0 + chunk_rows < ntotal) rc = fetch(r0 + chunk_rows, b ^ 1);
      if (rc == CSS_OK) rc = io.wait(h, b);
      if (rc == CSS_OK) W(io.buf[b], (size_t)std::min(chunk_rows, ntotal - r0) * d * 4);
    }
    io.release(h);
  }
  ok = ok && (fclose(fp) == 0);
  if (rc == CSS_OK && !ok) {
    set_error("write to %s failed", tmp.c_str());
    rc = CSS_ERR_IO;
  }
  if (rc == CSS_OK && rename(tmp.c_str(), path) != 0) {
    set_error("cannot rename %s to %s", tmp.c_str(), path);
    rc = CSS_ERR_IO;
  }
  if (rc != CSS_OK) remove(tmp.c_str());
  return rc;
}

int css_index_load(css_index* h, const char* path) {
  CSS_REQUIRE(h != nullptr && path != nullptr, "NULL argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const bool composite = !h->shards.empty();
  FILE* fp = fopen(path, "rb");
  if (!fp) {
    set_error("cannot open %s", path);
    return CSS_ERR_IO;
  }
  uint32_t cc = 0;
  int32_t d = 0, metric = 0;
  int64_t nt = 0, dummy = 0;
  uint8_t trained = 0;
  uint64_t count = 0;
  bool ok = true;
  auto R = [&](void* p, size_t n) { ok = ok && (fread(p, 1, n, fp) == n); };
  R(&cc, 4); R(&d, 4); R(&nt, 8); R(&dummy, 8); R(&dummy, 8); R(&trained, 1); R(&metric, 4);
  if (ok && metric > 1) {
    float metric_arg;
    R(&metric_arg, 4);
  }
  R(&count, 8);
  int rc = CSS_OK;
  const int64_t row_limit = composite ? (((int64_t)1 << 31) - 64) * (int64_t)h->shards.size() : ((int64_t)1 << 31) - 64;
  if (!ok || (cc != fourcc("IxFI") && cc != fourcc("IxF2"))) {
    set_error("%s is not a faiss IndexFlat file", path);
    rc = CSS_ERR_IO;
  } else if (d != h->dim) {
    set_error("%s has d=%d, index has d=%d", path, d, h->dim);
    rc = CSS_ERR_IO;
  } else if (nt < 0 || nt >= row_limit || count != (uint64_t)nt * (uint64_t)d) {
    set_error("%s: inconsistent header (ntotal=%lld count=%llu)", path, (long long)nt,
              (unsigned long long)count);
    rc = CSS_ERR_IO;
  }
  if (rc == CSS_OK) {
    const int new_metric = (cc == fourcc("IxF2")) ? CSS_METRIC_L2 : CSS_METRIC_INNER_PRODUCT;
    h->metric = new_metric;
    if (composite) {
      rc = sharded_reset(h);
      for (css_index* s : h->shards) s->metric = new_metric;
      if (rc == CSS_OK) rc = sharded_reserve(h, nt);
    } else {
      rc = single_reset(h);
      if (rc == CSS_OK) rc = single_grow(h, nt);
    }
  }
  if (rc == CSS_OK && nt > 0) {
    const int64_t chunk_rows = chunk_rows_for(d);
    IoBuffers io;
    rc = io.init(h, (size_t)chunk_rows * d * 4);
    std::vector<RowSpan> spans;
    int b = 0;
    for (int64_t r0 = 0; r0 < nt && rc == CSS_OK; r0 += chunk_rows, b ^= 1) {
      const int64_t nr = std::min(chunk_rows, nt - r0);
      const size_t nb = (size_t)nr * d * 4;
      rc = io.wait(h, b);   // the copies of two chunks ago have drained this buffer
      if (rc != CSS_OK) break;
      if (fread(io.buf[b], 1, nb, fp) != nb) {
        set_error("%s: truncated", path);
        rc = CSS_ERR_IO;
        break;
      }
      spans_of(h, r0, nr, &spans);
      for (const RowSpan& sp : spans) {
        css_index* sh = shard_of(h, sp);
        DeviceGuard gs(sh->device);
        rc = put_rows_host_async(sh, reinterpret_cast<const float*>(io.buf[b]) + (size_t)(sp.global - r0) * d, sp.local,
                                 sp.n, /*normalize=*/0);
        if (rc == CSS_OK) rc = single_mark_alive(sh, sp.local, sp.n);
        if (rc == CSS_OK) rc = io.mark(h, b, sp.shard);
        if (rc != CSS_OK) break;
        sh->ntotal = sp.local + sp.n;
        sh->version++;
      }
    }
    if (rc == CSS_OK) rc = io.wait(h, 0);
    if (rc == CSS_OK) rc = io.wait(h, 1);
    io.release(h);
    if (rc == CSS_OK && composite) h->composite_ntotal = nt;
  }
  fclose(fp);
  return rc;
}

}  // extern "C"
