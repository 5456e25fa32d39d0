// index.cu -- host side of the flat index C ABI (see include/css_b200.h).
#include "index_internal.h"

#include <algorithm>
#include <vector>

using namespace css;

namespace {

template <typename T>
int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) return CSS_OK;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return CSS_ERR_OOM;
  }
  return CSS_OK;
}

inline int64_t words_for(int64_t rows) { return (rows + 31) / 32; }

// Grow every per-row array to `new_cap` rows, preserving contents.
int grow(css_index* h, int64_t new_cap) {
  if (new_cap <= h->capacity) return CSS_OK;
  // keep row-bitmask words whole and rows a multiple of 32
  new_cap = (new_cap + 31) / 32 * 32;
  const size_t d = (size_t)h->dim;
  float* nx = nullptr;
  __nv_bfloat16* nxb = nullptr;
  uint32_t *nalive = nullptr, *nmask = nullptr;
  CSS_CHECK(dev_alloc(&nx, (size_t)new_cap * d));
  if (dev_alloc(&nxb, (size_t)new_cap * d) != CSS_OK) {
    cudaFree(nx);
    return CSS_ERR_OOM;
  }
  if (dev_alloc(&nalive, (size_t)words_for(new_cap)) != CSS_OK ||
      dev_alloc(&nmask, (size_t)words_for(new_cap)) != CSS_OK) {
    cudaFree(nx);
    cudaFree(nxb);
    cudaFree(nalive);
    return CSS_ERR_OOM;
  }
  cudaStream_t st = h->stream;
  CSS_CUDA(cudaMemsetAsync(nalive, 0, (size_t)words_for(new_cap) * 4, st));
  CSS_CUDA(cudaMemsetAsync(nmask, 0, (size_t)words_for(new_cap) * 4, st));
  if (h->ntotal > 0) {
    CSS_CUDA(cudaMemcpyAsync(nx, h->x, (size_t)h->ntotal * d * 4, cudaMemcpyDeviceToDevice, st));
    CSS_CUDA(cudaMemcpyAsync(nxb, h->xb, (size_t)h->ntotal * d * 2, cudaMemcpyDeviceToDevice, st));
    CSS_CUDA(cudaMemcpyAsync(nalive, h->alive, (size_t)words_for(h->ntotal) * 4,
                             cudaMemcpyDeviceToDevice, st));
  }
  for (int c = 0; c < CSS_MAX_COLUMNS; ++c) {
    if (!h->cols[c]) continue;
    int32_t* nc = nullptr;
    CSS_CHECK(dev_alloc(&nc, (size_t)new_cap));
    int64_t blocks = (new_cap + 255) / 256;
    fill_i32_kernel<<<(unsigned)blocks, 256, 0, st>>>(nc, new_cap, CSS_NULL_VALUE);
    CSS_LAUNCHED();
    if (h->ntotal > 0)
      CSS_CUDA(cudaMemcpyAsync(nc, h->cols[c], (size_t)h->ntotal * 4, cudaMemcpyDeviceToDevice, st));
    CSS_CUDA(cudaStreamSynchronize(st));
    cudaFree(h->cols[c]);
    h->cols[c] = nc;
  }
  CSS_CUDA(cudaStreamSynchronize(st));
  cudaFree(h->x);
  cudaFree(h->xb);
  cudaFree(h->alive);
  cudaFree(h->mask);
  h->x = nx;
  h->xb = nxb;
  h->alive = nalive;
  h->mask = nmask;
  h->capacity = new_cap;
  return CSS_OK;
}

int ensure_room(css_index* h, int64_t extra) {
  int64_t need = h->ntotal + extra;
  if (need <= h->capacity) return CSS_OK;
  int64_t cap = std::max<int64_t>(need, std::max<int64_t>(1024, h->capacity + h->capacity / 2));
  return grow(h, cap);
}

// Rows [ntotal, ntotal+n) were produced in `src_dev`; normalise/copy + bookkeeping.
int append_from_device(css_index* h, const float* src_dev, int64_t n, int normalize,
                       cudaStream_t st) {
  const int warps_per_block = 8;
  int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
  const size_t off = (size_t)h->ntotal * h->dim;
  append_rows_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, st>>>(
      src_dev, n, h->dim, normalize, h->x + off, h->xb + off, h->max_norm_dev);
  CSS_LAUNCHED();
  // new rows are alive
  int64_t w0 = h->ntotal >> 5, w1 = (h->ntotal + n - 1) >> 5;
  int64_t nw = w1 - w0 + 1;
  set_bits_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(h->alive, h->ntotal, n, 1);
  CSS_LAUNCHED();
  return CSS_OK;
}

template <int KPL, int METRIC>
int launch_scan_kd(css_index* h, const ScanParams& p, int nq, cudaStream_t st) {
  const bool d768 = (h->dim == 768);
  size_t smem = sizeof(KeyId) * kMergeCap + (d768 ? 0 : (size_t)h->dim * 4);
  dim3 grid((unsigned)h->scan_blocks, (unsigned)nq);
  if (d768) {
    auto kern = scan_topk_kernel<KPL, METRIC, true>;
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kScanThreads, smem, st>>>(p);
  } else {
    auto kern = scan_topk_kernel<KPL, METRIC, false>;
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kScanThreads, smem, st>>>(p);
  }
  CSS_LAUNCHED();
  return CSS_OK;
}

template <int METRIC>
int launch_scan_m(css_index* h, const ScanParams& p, int nq, cudaStream_t st) {
  if (p.k <= 32) return launch_scan_kd<1, METRIC>(h, p, nq, st);
  if (p.k <= 64) return launch_scan_kd<2, METRIC>(h, p, nq, st);
  return launch_scan_kd<4, METRIC>(h, p, nq, st);
}

// Compile a host css_filter into device FilterParams (uploads bitsets / row mask).
int build_filter_params(css_index* h, const css_filter* f, FilterParams* fp, cudaStream_t st) {
  memset(fp, 0, sizeof(*fp));
  fp->n = h->ntotal;
  fp->out = h->mask;
  fp->n_pass = h->n_pass_dev;
  fp->alive = (f && f->ignore_alive) ? nullptr : h->alive;
  if (!f) return CSS_OK;
  CSS_REQUIRE(f->n_clauses >= 0 && f->n_clauses <= CSS_MAX_CLAUSES, "n_clauses %d out of range",
              f->n_clauses);
  CSS_REQUIRE(f->n_clauses == 0 || f->clauses, "clauses is NULL");
  size_t words = 0;
  for (int i = 0; i < f->n_clauses; ++i) {
    const css_clause& c = f->clauses[i];
    CSS_REQUIRE(c.column >= 0 && c.column < CSS_MAX_COLUMNS, "clause %d: column %d out of range", i,
                c.column);
    CSS_REQUIRE(c.kind == CSS_CLAUSE_RANGE || c.kind == CSS_CLAUSE_SET, "clause %d: bad kind", i);
    if (c.kind == CSS_CLAUSE_SET) {
      CSS_REQUIRE(c.set_nbits >= 0 && (c.set_nbits == 0 || c.set_bits), "clause %d: bad set", i);
      words += (size_t)(c.set_nbits + 31) / 32;
    }
  }
  if (words > h->set_scratch_words) {
    if (h->set_scratch) cudaFree(h->set_scratch);
    h->set_scratch_words = 0;
    size_t want = std::max<size_t>(words * 2, 1024);
    CSS_CHECK(dev_alloc(&h->set_scratch, want));
    h->set_scratch_words = want;
  }
  // stage all bitsets contiguously in pinned memory, one H2D copy
  if (words) {
    CSS_CHECK(ensure_pinned(h, words * 4));
    uint32_t* stage = reinterpret_cast<uint32_t*>(h->pinned);
    size_t off = 0;
    for (int i = 0; i < f->n_clauses; ++i) {
      const css_clause& c = f->clauses[i];
      if (c.kind != CSS_CLAUSE_SET) continue;
      size_t w = (size_t)(c.set_nbits + 31) / 32;
      if (w) memcpy(stage + off, c.set_bits, w * 4);
      off += w;
    }
    CSS_CUDA(cudaMemcpyAsync(h->set_scratch, stage, words * 4, cudaMemcpyHostToDevice, st));
  }
  size_t off = 0;
  for (int i = 0; i < f->n_clauses; ++i) {
    const css_clause& c = f->clauses[i];
    DevClause& d = fp->c[i];
    d.col = h->cols[c.column];
    d.kind = c.kind;
    d.lo = c.lo;
    d.hi = c.hi;
    d.bits = nullptr;
    d.nbits = 0;
    if (c.kind == CSS_CLAUSE_SET) {
      d.bits = h->set_scratch + off;
      d.nbits = c.set_nbits;
      off += (size_t)(c.set_nbits + 31) / 32;
    }
  }
  fp->n_clauses = f->n_clauses;
  if (f->row_mask) {
    int64_t w = words_for(h->ntotal);
    if (w > h->rowmask_words) {
      if (h->rowmask_scratch) cudaFree(h->rowmask_scratch);
      h->rowmask_words = 0;
      CSS_CHECK(dev_alloc(&h->rowmask_scratch, (size_t)words_for(h->capacity)));
      h->rowmask_words = words_for(h->capacity);
    }
    // the staging buffer for set bits was already consumed by an async copy on the same
    // stream; pageable source is fine here (synchronous w.r.t. host)
    CSS_CUDA(cudaStreamSynchronize(st));
    CSS_CUDA(cudaMemcpyAsync(h->rowmask_scratch, f->row_mask, (size_t)w * 4, cudaMemcpyHostToDevice, st));
    fp->row_mask = h->rowmask_scratch;
  }
  return CSS_OK;
}

// Evaluate `f` into h->mask.  *mask_out = nullptr when every row passes trivially.
int eval_filter(css_index* h, const css_filter* f, const uint32_t** mask_out, int64_t* n_pass,
                bool need_count, cudaStream_t st) {
  const bool trivial = (!f || (f->n_clauses == 0 && !f->row_mask)) && (!h->any_dead || (f && f->ignore_alive));
  if (trivial && !need_count) {
    *mask_out = nullptr;
    if (n_pass) *n_pass = h->ntotal;
    return CSS_OK;
  }
  if (h->ntotal == 0) {
    *mask_out = h->mask;
    if (n_pass) *n_pass = 0;
    return CSS_OK;
  }
  FilterParams fp;
  CSS_CHECK(build_filter_params(h, f, &fp, st));
  CSS_CUDA(cudaMemsetAsync(h->n_pass_dev, 0, sizeof(unsigned long long), st));
  int64_t blocks = (h->ntotal + 255) / 256;
  filter_mask_kernel<<<(unsigned)blocks, 256, 0, st>>>(fp);
  CSS_LAUNCHED();
  *mask_out = h->mask;
  if (n_pass) {
    unsigned long long c = 0;
    CSS_CUDA(cudaMemcpyAsync(&c, h->n_pass_dev, sizeof(c), cudaMemcpyDeviceToHost, st));
    CSS_CUDA(cudaStreamSynchronize(st));
    *n_pass = (int64_t)c;
  }
  return CSS_OK;
}

}  // namespace

namespace css {

int ensure_pinned(css_index* h, size_t bytes) {
  if (bytes <= h->pinned_bytes) return CSS_OK;
  if (h->pinned) {
    cudaDeviceSynchronize();  // an async copy may still be reading the old block
    cudaFreeHost(h->pinned);
  }
  h->pinned = nullptr;
  h->pinned_bytes = 0;
  size_t want = std::max<size_t>(bytes, (size_t)1 << 20);
  cudaError_t e = cudaMallocHost(&h->pinned, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    return CSS_ERR_OOM;
  }
  h->pinned_bytes = want;
  return CSS_OK;
}

int ensure_query_scratch(css_index* h, int nq) {
  if (nq <= h->max_nq) return CSS_OK;
  int want = std::max(nq, std::max(16, h->max_nq * 2));
  cudaFree(h->q_dev);
  cudaFree(h->part);
  cudaFree(h->ticket);
  cudaFree(h->D_dev);
  cudaFree(h->ovf_list);
  cudaFree(h->ovf_count);
  h->q_dev = nullptr; h->part = nullptr; h->ticket = nullptr; h->D_dev = nullptr;
  h->ovf_list = nullptr; h->ovf_count = nullptr;
  h->max_nq = 0;
  CSS_CHECK(dev_alloc(&h->q_dev, (size_t)want * h->dim));
  CSS_CHECK(dev_alloc(&h->part, (size_t)want * h->scan_blocks * CSS_MAX_K));
  CSS_CHECK(dev_alloc(&h->ticket, (size_t)want));
  CSS_CHECK(dev_alloc(&h->D_dev, (size_t)want * CSS_MAX_K * 3 + 8));   // scores, then the ids of the same call (one D2H)
  CSS_CHECK(dev_alloc(&h->ovf_list, (size_t)want));
  CSS_CHECK(dev_alloc(&h->ovf_count, (size_t)1));
  CSS_CUDA(cudaMemsetAsync(h->ticket, 0, (size_t)want * sizeof(unsigned int), h->stream));
  CSS_CUDA(cudaStreamSynchronize(h->stream));
  h->max_nq = want;
  return CSS_OK;
}

// Two-phase exact scan (inner product, d = 768, k <= 32, with or without a row mask): phase 1 streams the bf16 shadow
// rows -- half the bytes of the fp32 corpus -- and leaves the 32 best of every scan block's slice by that
// score; phase 2 (rescore768_kernel) proves that the true top-k lies inside those lists, re-scores the
// candidates in fp32 with the arithmetic of the fp32 scan and emits the exact result; queries it cannot
// prove are re-run by the fp32 scan on the device (qlist), so the answer is always the exact one.
// CSS_SCAN_BF16=0 disables it.
int launch_phase1(css_index* h, const float* q_dev, int nq, const uint32_t* mask_dev, ScanParams* out, cudaStream_t st) {
  ScanParams p;
  p.x = h->x;
  p.xb = h->xb;
  p.n = h->ntotal;
  p.d = h->dim;
  p.q = q_dev;
  p.mask = mask_dev;   // nullable: filter / alive bits (the selected rows of a window are compacted first)
  p.k = kTwoPhaseMaxK;
  p.part = h->part;
  p.ticket = h->ticket;
  p.id_offset = 0;
  p.D = nullptr;
  p.I = nullptr;
  p.qlist = nullptr;
  p.qcount = nullptr;
  p.no_merge = 1;
  p.zero_on_entry = h->ovf_count;
  const size_t smem = sizeof(KeyId) * kMergeCap;
  auto kern = scan_topk_kernel<1, CSS_METRIC_INNER_PRODUCT, true, true>;
  CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)h->scan_blocks, (unsigned)nq), kScanThreads, smem, st>>>(p);
  CSS_LAUNCHED();
  if (out) *out = p;
  return CSS_OK;
}

int two_phase_scan(css_index* h, const float* q_dev, int nq, int k, const uint32_t* mask_dev, int64_t id_offset,
                   float* D_dev, int64_t* I_dev, cudaStream_t st) {
  const int kp = kTwoPhaseMaxK;
  ScanParams p;
  CSS_CHECK(launch_phase1(h, q_dev, nq, mask_dev, &p, st));
  RescoreParams r;
  r.x = h->x;
  r.q = q_dev;
  r.k = k;
  r.kp = kp;
  r.blocks = h->scan_blocks;
  r.eps_scale = 1.10f / 512.f;
  r.max_norm = h->max_norm_dev;
  r.part = h->part;
  r.id_offset = id_offset;
  r.D = D_dev;
  r.I = I_dev;
  r.ovf_list = h->ovf_list;
  r.ovf_count = h->ovf_count;
  {
    const size_t smem = sizeof(KeyId) * kRescoreSort;
    CSS_CUDA(cudaFuncSetAttribute(rescore768_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rescore768_kernel<<<(unsigned)nq, kRescoreThreads, smem, st>>>(r);
    CSS_LAUNCHED();
  }
  // unproven queries: fp32 scan, driven by the device-side list (an empty list costs one idle launch)
  ScanParams f = p;
  f.xb = nullptr;
  f.k = k;
  f.id_offset = id_offset;
  f.D = D_dev;
  f.I = I_dev;
  f.qlist = h->ovf_list;
  f.qcount = h->ovf_count;
  f.no_merge = 0;
  f.zero_on_entry = nullptr;
  CSS_CHECK((launch_scan_m<CSS_METRIC_INNER_PRODUCT>(h, f, nq, st)));
  return CSS_OK;
}

int scan_search(css_index* h, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                int64_t id_offset, float* D_dev, int64_t* I_dev, cudaStream_t st) {
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  CSS_CHECK(ensure_query_scratch(h, nq));
  static const bool bf16_phase = [] { const char* v = getenv("CSS_SCAN_BF16"); return v ? atoi(v) != 0 : true; }();
  if (bf16_phase && h->metric == CSS_METRIC_INNER_PRODUCT && h->dim == 768 && k <= kTwoPhaseMaxK &&
      nq <= 64 && h->xb != nullptr && h->ntotal > 0 && (int64_t)h->scan_blocks * kTwoPhaseMaxK <= kRescoreSort)
    return two_phase_scan(h, q_dev, nq, k, mask_dev, id_offset, D_dev, I_dev, st);
  // gridDim.y is limited to 65535; chunk the query batch
  const int chunk = 4096;
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    int nqc = std::min(chunk, nq - q0);
    ScanParams p;
    p.x = h->x;
    p.xb = nullptr;
    p.n = h->ntotal;
    p.d = h->dim;
    p.q = q_dev + (size_t)q0 * h->dim;
    p.mask = mask_dev;
    p.k = k;
    p.part = h->part;
    p.ticket = h->ticket;
    p.id_offset = id_offset;
    p.D = D_dev + (size_t)q0 * k;
    p.I = I_dev + (size_t)q0 * k;
    p.qlist = nullptr;
    p.qcount = nullptr;
    p.no_merge = 0;
    p.zero_on_entry = nullptr;
    if (h->metric == CSS_METRIC_INNER_PRODUCT)
      CSS_CHECK((launch_scan_m<CSS_METRIC_INNER_PRODUCT>(h, p, nqc, st)));
    else
      CSS_CHECK((launch_scan_m<CSS_METRIC_L2>(h, p, nqc, st)));
  }
  return CSS_OK;
}

}  // namespace css

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int css_index_create(int dim, int metric, int device, css_index** out) {
  CSS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  CSS_REQUIRE(dim >= 1 && dim <= 65536, "dim %d out of range", dim);
  CSS_REQUIRE(metric == CSS_METRIC_INNER_PRODUCT || metric == CSS_METRIC_L2, "unknown metric %d",
              metric);
  CSS_CHECK(ensure_device(device));
  DeviceGuard g(device);
  css_index* h = new (std::nothrow) css_index();
  if (!h) {
    set_error("out of host memory");
    return CSS_ERR_OOM;
  }
  h->dim = dim;
  h->metric = metric;
  h->device = device;
  h->n_sm = sm_count(device);
  h->scan_blocks = h->n_sm;  // one 512-thread CTA per SM (128 regs/thread)
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    delete h;
    return CSS_ERR_CUDA;
  }
  if (dev_alloc(&h->n_pass_dev, 1) != CSS_OK || dev_alloc(&h->max_norm_dev, 1) != CSS_OK ||
      cudaMemset(h->max_norm_dev, 0, sizeof(float)) != cudaSuccess) {
    cudaFree(h->n_pass_dev);
    cudaStreamDestroy(h->stream);
    delete h;
    return CSS_ERR_OOM;
  }
  *out = h;
  return CSS_OK;
}

int css_index_destroy(css_index* h) {
  if (!h) return CSS_OK;
  {
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    batched_release(h);
    cudaFree(h->x);
    cudaFree(h->xb);
    cudaFree(h->alive);
    cudaFree(h->mask);
    for (int c = 0; c < CSS_MAX_COLUMNS; ++c) cudaFree(h->cols[c]);
    cudaFree(h->q_dev);
    cudaFree(h->part);
    cudaFree(h->ticket);
    cudaFree(h->D_dev);
    cudaFree(h->ovf_list);
    cudaFree(h->ovf_count);
    cudaFree(h->set_scratch);
    cudaFree(h->rowmask_scratch);
    cudaFree(h->n_pass_dev);
    cudaFree(h->max_norm_dev);
    if (h->pinned) cudaFreeHost(h->pinned);
    cudaStreamDestroy(h->stream);
  }
  delete h;
  return CSS_OK;
}

int css_index_dim(const css_index* h) { return h ? h->dim : CSS_ERR_INVALID; }
int css_index_metric(const css_index* h) { return h ? h->metric : CSS_ERR_INVALID; }
int64_t css_index_ntotal(const css_index* h) { return h ? h->ntotal : (int64_t)CSS_ERR_INVALID; }
int64_t css_index_capacity(const css_index* h) { return h ? h->capacity : (int64_t)CSS_ERR_INVALID; }

int css_index_reserve(css_index* h, int64_t capacity) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(capacity >= 0 && capacity < ((int64_t)1 << 31), "capacity out of range");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  return grow(h, capacity);
}

int css_index_reset(css_index* h) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  h->ntotal = 0;
  h->any_dead = false;
  CSS_CUDA(cudaMemsetAsync(h->max_norm_dev, 0, sizeof(float), h->stream));
  if (h->capacity > 0) {
    CSS_CUDA(cudaMemsetAsync(h->alive, 0, (size_t)words_for(h->capacity) * 4, h->stream));
    for (int c = 0; c < CSS_MAX_COLUMNS; ++c) {
      if (!h->cols[c]) continue;
      int64_t blocks = (h->capacity + 255) / 256;
      fill_i32_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(h->cols[c], h->capacity, CSS_NULL_VALUE);
      CSS_LAUNCHED();
    }
    CSS_CUDA(cudaStreamSynchronize(h->stream));
  }
  return CSS_OK;
}

int css_index_compact(css_index* h, const int64_t* keep_ids_host, int64_t n_keep) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(n_keep >= 0 && (n_keep == 0 || keep_ids_host != nullptr), "bad keep list");
  std::lock_guard<std::mutex> lk(h->mu);
  CSS_REQUIRE(n_keep <= h->ntotal, "keep list longer than the index (%lld > %lld)", (long long)n_keep,
              (long long)h->ntotal);
  for (int64_t i = 0; i < n_keep; ++i) {
    const int64_t id = keep_ids_host[i];
    CSS_REQUIRE(id >= 0 && id < h->ntotal && (i == 0 || id > keep_ids_host[i - 1]),
                "keep ids must be strictly ascending row ids (entry %lld = %lld)", (long long)i, (long long)id);
  }
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const int d = h->dim;
  const int64_t win = 65536;   // rows per window: 200 MB of fp32 staging at d = 768
  int64_t* ids_dev = nullptr;
  float* sx = nullptr;
  __nv_bfloat16* sb = nullptr;
  int32_t* sc = nullptr;
  uint32_t* sw = nullptr;
  auto cleanup = [&]() {
    cudaFree(ids_dev); cudaFree(sx); cudaFree(sb); cudaFree(sc); cudaFree(sw);
  };
  if (n_keep > 0) {
    const int64_t w = std::min(win, n_keep);
    if (cudaMalloc(&ids_dev, (size_t)n_keep * 8) != cudaSuccess || cudaMalloc(&sx, (size_t)w * d * 4) != cudaSuccess ||
        cudaMalloc(&sb, (size_t)w * d * 2) != cudaSuccess || cudaMalloc(&sc, (size_t)w * 4) != cudaSuccess ||
        cudaMalloc(&sw, (size_t)words_for(n_keep) * 4) != cudaSuccess) {
      (void)cudaGetLastError();
      cleanup();
      set_error("compaction staging allocation failed");
      return CSS_ERR_OOM;
    }
    cudaError_t ce = cudaMemcpyAsync(ids_dev, keep_ids_host, (size_t)n_keep * 8, cudaMemcpyHostToDevice, st);
    // alive bits first (they are read by row id from the old layout)
    if (ce == cudaSuccess) {
      gather_bits_kernel<<<(unsigned)((n_keep + 255) / 256), 256, 0, st>>>(h->alive, ids_dev, n_keep, sw);
      css::g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    for (int64_t i0 = 0; i0 < n_keep && ce == cudaSuccess; i0 += win) {
      const int64_t n = std::min(win, n_keep - i0);
      gather_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(h->x, h->xb, ids_dev + i0, n, d, sx, sb);
      css::g_launches.fetch_add(1, std::memory_order_relaxed);
      ce = cudaMemcpyAsync(h->x + i0 * d, sx, (size_t)n * d * 4, cudaMemcpyDeviceToDevice, st);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->xb + i0 * d, sb, (size_t)n * d * 2, cudaMemcpyDeviceToDevice, st);
      for (int c = 0; c < CSS_MAX_COLUMNS && ce == cudaSuccess; ++c) {
        if (!h->cols[c]) continue;
        gather_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->cols[c], ids_dev + i0, n, sc);
        css::g_launches.fetch_add(1, std::memory_order_relaxed);
        ce = cudaMemcpyAsync(h->cols[c] + i0, sc, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
      }
    }
    if (ce == cudaSuccess) ce = cudaMemsetAsync(h->alive, 0, (size_t)words_for(h->capacity) * 4, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->alive, sw, (size_t)words_for(n_keep) * 4, cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    cleanup();
    if (ce != cudaSuccess) {
      set_error("compaction failed: %s", cudaGetErrorString(ce));
      return CSS_ERR_CUDA;
    }
  } else if (h->capacity > 0) {
    CSS_CUDA(cudaMemsetAsync(h->alive, 0, (size_t)words_for(h->capacity) * 4, st));
    CSS_CUDA(cudaStreamSynchronize(st));
  }
  // rows past the new end: NULL columns again (what a fresh row looks like)
  for (int c = 0; c < CSS_MAX_COLUMNS; ++c) {
    if (!h->cols[c] || h->ntotal <= n_keep) continue;
    const int64_t n = h->ntotal - n_keep;
    fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->cols[c] + n_keep, n, CSS_NULL_VALUE);
    CSS_LAUNCHED();
  }
  CSS_CUDA(cudaStreamSynchronize(st));
  h->ntotal = n_keep;
  h->any_dead = true;   // conservative: the scan keeps honouring the alive bits
  return CSS_OK;
}

int css_index_add(css_index* h, const float* x_host, int64_t n, int normalize, int64_t* first_id_out) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(n >= 0, "n < 0");
  CSS_REQUIRE(n == 0 || x_host != nullptr, "x_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CSS_REQUIRE(h->ntotal + n < ((int64_t)1 << 31), "index would exceed 2^31 rows per shard");
  if (first_id_out) *first_id_out = h->ntotal;
  if (n == 0) return CSS_OK;
  CSS_CHECK(ensure_room(h, n));
  // Stage through a bounded device buffer: rows are written in place by the
  // append kernel (it normalises), so upload into the tail of x's own storage
  // is not possible when normalising in a different layout; use chunks.
  const int64_t chunk_rows = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)h->dim * 4));
  float* stage = nullptr;
  CSS_CHECK(dev_alloc(&stage, (size_t)std::min(chunk_rows, n) * h->dim));
  int rc = CSS_OK;
  for (int64_t r0 = 0; r0 < n && rc == CSS_OK; r0 += chunk_rows) {
    int64_t nr = std::min(chunk_rows, n - r0);
    cudaError_t e = cudaMemcpyAsync(stage, x_host + (size_t)r0 * h->dim, (size_t)nr * h->dim * 4,
                                    cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) {
      set_error("H2D copy failed: %s", cudaGetErrorString(e));
      rc = CSS_ERR_CUDA;
      break;
    }
    rc = append_from_device(h, stage, nr, normalize, h->stream);
    if (rc != CSS_OK) break;
    e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
      set_error("append failed: %s", cudaGetErrorString(e));
      rc = CSS_ERR_CUDA;
      break;
    }
    h->ntotal += nr;
  }
  cudaFree(stage);
  return rc;
}

int css_index_add_device(css_index* h, const float* x_dev, int64_t n, int normalize,
                         int64_t* first_id_out, void* stream) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(n >= 0, "n < 0");
  CSS_REQUIRE(n == 0 || x_dev != nullptr, "x_dev is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CSS_REQUIRE(h->ntotal + n < ((int64_t)1 << 31), "index would exceed 2^31 rows per shard");
  if (first_id_out) *first_id_out = h->ntotal;
  if (n == 0) return CSS_OK;
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  if (h->ntotal + n > h->capacity) {
    // growth reallocates: order it after the caller's stream work and before ours
    CSS_CUDA(cudaStreamSynchronize(st));
    CSS_CHECK(ensure_room(h, n));
  }
  CSS_CHECK(append_from_device(h, x_dev, n, normalize, st));
  h->ntotal += n;
  return CSS_OK;
}

int css_index_get_rows(css_index* h, int64_t start, int64_t n, float* out_host) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(start >= 0 && n >= 0 && start + n <= h->ntotal, "row range out of bounds");
  if (n == 0) return CSS_OK;
  CSS_REQUIRE(out_host != nullptr, "out_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CSS_CUDA(cudaMemcpyAsync(out_host, h->x + (size_t)start * h->dim, (size_t)n * h->dim * 4,
                           cudaMemcpyDeviceToHost, h->stream));
  CSS_CUDA(cudaStreamSynchronize(h->stream));
  return CSS_OK;
}

int css_index_set_column(css_index* h, int column, const int32_t* values_host, int64_t start,
                         int64_t n) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(column >= 0 && column < CSS_MAX_COLUMNS, "column %d out of range", column);
  CSS_REQUIRE(start >= 0 && n >= 0 && start + n <= h->ntotal, "row range out of bounds");
  if (n == 0) return CSS_OK;
  CSS_REQUIRE(values_host != nullptr, "values_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  if (!h->cols[column]) {
    CSS_CHECK(dev_alloc(&h->cols[column], (size_t)h->capacity));
    int64_t blocks = (h->capacity + 255) / 256;
    fill_i32_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(h->cols[column], h->capacity,
                                                             CSS_NULL_VALUE);
    CSS_LAUNCHED();
  }
  CSS_CUDA(cudaMemcpyAsync(h->cols[column] + start, values_host, (size_t)n * 4,
                           cudaMemcpyHostToDevice, h->stream));
  CSS_CUDA(cudaStreamSynchronize(h->stream));
  return CSS_OK;
}

int css_index_set_alive(css_index* h, const uint8_t* alive_host, int64_t start, int64_t n) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(start >= 0 && n >= 0 && start + n <= h->ntotal, "row range out of bounds");
  if (n == 0) return CSS_OK;
  CSS_REQUIRE(alive_host != nullptr, "alive_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  uint8_t* tmp = nullptr;
  CSS_CHECK(dev_alloc(&tmp, (size_t)n));
  int rc = CSS_OK;
  cudaError_t e = cudaMemcpyAsync(tmp, alive_host, (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    alive_bytes_to_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(tmp, start, n, h->alive);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    set_error("set_alive failed: %s", cudaGetErrorString(e));
    rc = CSS_ERR_CUDA;
  }
  cudaFree(tmp);
  if (rc == CSS_OK) {
    for (int64_t i = 0; i < n; ++i)
      if (!alive_host[i]) {
        h->any_dead = true;
        break;
      }
  }
  return rc;
}

int css_index_filter_mask(css_index* h, const css_filter* f, uint32_t* mask_out_host,
                          int64_t* n_pass_out) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(mask_out_host != nullptr || h->ntotal == 0, "mask_out_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CSS_CHECK(ensure_device(h->device));
  const uint32_t* m = nullptr;
  int64_t n_pass = 0;
  // force evaluation even for the trivial filter so the mask is materialised
  css_filter empty;
  memset(&empty, 0, sizeof(empty));
  CSS_CHECK(eval_filter(h, f ? f : &empty, &m, &n_pass, /*need_count=*/true, h->stream));
  if (h->ntotal > 0) {
    CSS_CUDA(cudaMemcpyAsync(mask_out_host, h->mask, (size_t)words_for(h->ntotal) * 4,
                             cudaMemcpyDeviceToHost, h->stream));
    CSS_CUDA(cudaStreamSynchronize(h->stream));
    // bits beyond ntotal in the last word are zero by construction (ballot of row < n)
  }
  if (n_pass_out) *n_pass_out = n_pass;
  return CSS_OK;
}

int css_index_filter_mask_device(css_index* h, const css_filter* f, const uint32_t** mask_dev_out,
                                 int64_t* n_pass_out, void* stream) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(mask_dev_out != nullptr, "mask_dev_out is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  return eval_filter(h, f, mask_dev_out, n_pass_out, n_pass_out != nullptr, st);
}

int css_debug_scan_bf16(css_index* h, const float* q_dev, int nq, void* stream) {
  CSS_REQUIRE(h != nullptr && q_dev != nullptr, "NULL argument");
  CSS_REQUIRE(nq >= 1 && nq <= 64, "nq=%d outside [1, 64]", nq);
  CSS_REQUIRE(h->metric == CSS_METRIC_INNER_PRODUCT && h->dim == 768 && h->ntotal > 0 && h->xb != nullptr,
              "the two-phase scan needs a non-empty 768-d inner-product index");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CSS_CHECK(ensure_query_scratch(h, nq));
  return launch_phase1(h, q_dev, nq, h->any_dead ? h->alive : nullptr, nullptr, stream ? (cudaStream_t)stream : h->stream);
}

int css_index_search_device(css_index* h, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                            int64_t id_offset, float* D_dev, int64_t* I_dev, void* stream) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(nq >= 0, "nq < 0");
  if (nq == 0) return CSS_OK;
  CSS_REQUIRE(q_dev && D_dev && I_dev, "NULL device buffer");
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const uint32_t* m = mask_dev;
  if (!m && h->any_dead) m = h->alive;
  if (nq >= CSS_BATCH_MIN_NQ && h->metric == CSS_METRIC_INNER_PRODUCT && h->dim % 64 == 0 &&
      h->ntotal >= 65536)
    return batched_search(h, q_dev, nq, k, m, id_offset, D_dev, I_dev, st);
  return scan_search(h, q_dev, nq, k, m, id_offset, D_dev, I_dev, st);
}

int css_index_search(css_index* h, const float* q_host, int nq, int k, const css_filter* filter,
                     float* D_host, int64_t* I_host) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(nq >= 0, "nq < 0");
  if (nq == 0) return CSS_OK;
  CSS_REQUIRE(q_host && D_host && I_host, "NULL host buffer");
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  CSS_CHECK(ensure_device(h->device));
  std::unique_lock<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const size_t qbytes = (size_t)nq * h->dim * 4;
  const size_t dbytes = (size_t)nq * k * 4, ibytes = (size_t)nq * k * 8;
  CSS_CHECK(ensure_query_scratch(h, nq));
  const uint32_t* m = nullptr;
  CSS_CHECK(eval_filter(h, filter, &m, nullptr, false, st));
  if (!m && h->any_dead) m = h->alive;
  // pinned staging: [q | D | I]
  CSS_CHECK(ensure_pinned(h, qbytes + dbytes + ibytes + 64));
  unsigned char* pin = reinterpret_cast<unsigned char*>(h->pinned);
  // set bitsets staged in the same pinned block were consumed by an async copy: wait for it
  if (filter && filter->n_clauses) CSS_CUDA(cudaStreamSynchronize(st));
  memcpy(pin, q_host, qbytes);
  size_t d_off = (qbytes + 15) / 16 * 16;
  size_t i_off = (d_off + dbytes + 15) / 16 * 16;
  CSS_CUDA(cudaMemcpyAsync(h->q_dev, pin, qbytes, cudaMemcpyHostToDevice, st));
  // results of this call: scores at D_dev, ids right behind them (same spacing as in the pinned block)
  int64_t* I_dev = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(h->D_dev) + (i_off - d_off));
  int rc;
  if (nq >= CSS_BATCH_MIN_NQ && h->metric == CSS_METRIC_INNER_PRODUCT && h->dim % 64 == 0 &&
      h->ntotal >= 65536)
    rc = batched_search(h, h->q_dev, nq, k, m, 0, h->D_dev, I_dev, st);
  else
    rc = scan_search(h, h->q_dev, nq, k, m, 0, h->D_dev, I_dev, st);
  if (rc != CSS_OK) return rc;
  CSS_CUDA(cudaMemcpyAsync(pin + d_off, h->D_dev, (i_off - d_off) + ibytes, cudaMemcpyDeviceToHost, st));
  CSS_CUDA(cudaStreamSynchronize(st));
  memcpy(D_host, pin + d_off, dbytes);
  memcpy(I_host, pin + i_off, ibytes);
  return CSS_OK;
}

int css_topk_merge_strided_device(const float* D_in, int64_t d_list_stride, const int64_t* I_in, int64_t i_list_stride,
                                  int n_lists, int nq, int k, int metric, float* D_out, int64_t* I_out, void* stream) {
  CSS_REQUIRE(D_in && I_in && D_out && I_out, "NULL device buffer");
  CSS_REQUIRE(n_lists >= 1 && nq >= 0 && k >= 1 && k <= CSS_MAX_K, "bad merge shape");
  CSS_REQUIRE((int64_t)n_lists * k <= 8192, "n_lists*k too large for the merge kernel");
  CSS_REQUIRE(d_list_stride >= (int64_t)nq * k && i_list_stride >= (int64_t)nq * k, "list stride below nq*k");
  if (nq == 0) return CSS_OK;
  int n_sort = 32;
  while (n_sort < n_lists * k) n_sort <<= 1;
  size_t smem = (size_t)n_sort * sizeof(KeyId64);
  cudaStream_t st = (cudaStream_t)stream;
  int threads = std::min(512, std::max(32, n_sort / 2));
  if (metric == CSS_METRIC_INNER_PRODUCT) {
    auto kern = merge_lists_kernel<CSS_METRIC_INNER_PRODUCT>;
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nq, threads, smem, st>>>(D_in, I_in, d_list_stride, i_list_stride, n_lists, nq, k, D_out, I_out);
  } else {
    auto kern = merge_lists_kernel<CSS_METRIC_L2>;
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nq, threads, smem, st>>>(D_in, I_in, d_list_stride, i_list_stride, n_lists, nq, k, D_out, I_out);
  }
  CSS_LAUNCHED();
  return CSS_OK;
}

int css_topk_merge_device(const float* D_in, const int64_t* I_in, int n_lists, int nq, int k,
                          int metric, float* D_out, int64_t* I_out, void* stream) {
  return css_topk_merge_strided_device(D_in, (int64_t)nq * k, I_in, (int64_t)nq * k, n_lists, nq, k, metric, D_out,
                                       I_out, stream);
}

// ---------------------------------------------------------------------------
// faiss IndexFlat file format (faiss/impl/index_write.cpp, write_index_header +
// IndexFlat codes; restated from the published format, faiss >= 1.7):
//   u32  fourcc  "IxFI" (inner product) | "IxF2" (L2)
//   i32  d
//   i64  ntotal
//   i64  dummy (1 << 20), i64 dummy (1 << 20)
//   u8   is_trained (1)
//   i32  metric_type (0 = IP, 1 = L2)
//   u64  count = ntotal * d          (vector<float> xb / codes.size()/4)
//   f32  data[count]
// ---------------------------------------------------------------------------
static uint32_t fourcc(const char s[4]) {
  return (uint32_t)(unsigned char)s[0] | ((uint32_t)(unsigned char)s[1] << 8) |
         ((uint32_t)(unsigned char)s[2] << 16) | ((uint32_t)(unsigned char)s[3] << 24);
}

int css_index_save(css_index* h, const char* path) {
  CSS_REQUIRE(h != nullptr && path != nullptr, "NULL argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
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
  int64_t nt = h->ntotal, dummy = (int64_t)1 << 20;
  uint8_t trained = 1;
  int32_t metric = h->metric;
  uint64_t count = (uint64_t)h->ntotal * (uint64_t)h->dim;
  W(&cc, 4); W(&d, 4); W(&nt, 8); W(&dummy, 8); W(&dummy, 8); W(&trained, 1); W(&metric, 4);
  W(&count, 8);
  int rc = CSS_OK;
  if (h->ntotal > 0 && ok) {
    const size_t chunk_bytes = (size_t)32 << 20;
    rc = ensure_pinned(h, chunk_bytes);
    const size_t total = (size_t)count * 4;
    for (size_t off = 0; off < total && rc == CSS_OK && ok; off += chunk_bytes) {
      size_t nb = std::min(chunk_bytes, total - off);
      cudaError_t e = cudaMemcpyAsync(h->pinned, reinterpret_cast<const char*>(h->x) + off, nb,
                                      cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) {
        set_error("D2H copy failed: %s", cudaGetErrorString(e));
        rc = CSS_ERR_CUDA;
        break;
      }
      W(h->pinned, nb);
    }
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
  if (!ok || (cc != fourcc("IxFI") && cc != fourcc("IxF2"))) {
    set_error("%s is not a faiss IndexFlat file", path);
    rc = CSS_ERR_IO;
  } else if (d != h->dim) {
    set_error("%s has d=%d, index has d=%d", path, d, h->dim);
    rc = CSS_ERR_IO;
  } else if (nt < 0 || nt >= ((int64_t)1 << 31) || count != (uint64_t)nt * (uint64_t)d) {
    set_error("%s: inconsistent header (ntotal=%lld count=%llu)", path, (long long)nt,
              (unsigned long long)count);
    rc = CSS_ERR_IO;
  }
  if (rc == CSS_OK) {
    h->ntotal = 0;
    h->any_dead = false;
    cudaMemsetAsync(h->max_norm_dev, 0, sizeof(float), h->stream);
    h->metric = (cc == fourcc("IxF2")) ? CSS_METRIC_L2 : CSS_METRIC_INNER_PRODUCT;
    if (h->capacity > 0)
      cudaMemsetAsync(h->alive, 0, (size_t)words_for(h->capacity) * 4, h->stream);
    rc = ensure_room(h, nt);
  }
  if (rc == CSS_OK && nt > 0) {
    const int64_t chunk_rows = std::max<int64_t>(1, ((int64_t)32 << 20) / ((int64_t)d * 4));
    rc = ensure_pinned(h, (size_t)chunk_rows * d * 4);
    float* stage = nullptr;
    if (rc == CSS_OK) rc = dev_alloc(&stage, (size_t)std::min(chunk_rows, nt) * d);
    for (int64_t r0 = 0; r0 < nt && rc == CSS_OK; r0 += chunk_rows) {
      int64_t nr = std::min(chunk_rows, nt - r0);
      size_t nb = (size_t)nr * d * 4;
      if (fread(h->pinned, 1, nb, fp) != nb) {
        set_error("%s: truncated", path);
        rc = CSS_ERR_IO;
        break;
      }
      cudaError_t e = cudaMemcpyAsync(stage, h->pinned, nb, cudaMemcpyHostToDevice, h->stream);
      if (e != cudaSuccess) {
        set_error("H2D copy failed: %s", cudaGetErrorString(e));
        rc = CSS_ERR_CUDA;
        break;
      }
      rc = append_from_device(h, stage, nr, /*normalize=*/0, h->stream);
      if (rc != CSS_OK) break;
      e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) {
        set_error("load failed: %s", cudaGetErrorString(e));
        rc = CSS_ERR_CUDA;
        break;
      }
      h->ntotal += nr;
    }
    cudaFree(stage);
  }
  fclose(fp);
  return rc;
}

}  // extern "C"
