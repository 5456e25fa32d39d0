// index.cu -- host side of the flat index C ABI (see include/css_b200.h): one single-device shard.
// The multi-device composite lives in index_sharded.cu, the result exchange in exchange.cu.
#include "index_internal.h"

#include <algorithm>
#include <cstdlib>
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

// The int8 shadow copy (first tier of the two-phase scan) exists where the kernel that reads it does.
inline bool has_int8_shadow(const css_index* h) { return h->dim == 768 && h->metric == CSS_METRIC_INNER_PRODUCT; }

int vmm_status(int rc) { return rc == 0 ? CSS_OK : (rc == -4 ? CSS_ERR_OOM : CSS_ERR_CUDA); }

// Largest row count the virtual ranges are sized for: what fits the device's HBM as fp32 rows.
int64_t max_rows_for(const css_index* h) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = (size_t)192 << 30;
  int64_t rows = (int64_t)(total_b / ((size_t)h->dim * 4));
  rows = std::min<int64_t>(rows, ((int64_t)1 << 31) - 64);
  return std::max<int64_t>(rows, 1024) / 32 * 32;
}

int reserve_ranges(css_index* h) {
  if (h->vx.base) return CSS_OK;
  const int64_t mr = max_rows_for(h);
  const size_t d = (size_t)h->dim;
  CSS_CHECK(vmm_status(vmm_reserve(&h->vx, h->device, (size_t)mr * d * 4)));
  CSS_CHECK(vmm_status(vmm_reserve(&h->vxb, h->device, (size_t)mr * d * 2)));
  if (has_int8_shadow(h)) {
    CSS_CHECK(vmm_status(vmm_reserve(&h->vxq, h->device, (size_t)mr * d)));
    CSS_CHECK(vmm_status(vmm_reserve(&h->vxs, h->device, (size_t)mr * 4)));
    h->xq = reinterpret_cast<int8_t*>(h->vxq.ptr());
    h->xs = reinterpret_cast<float*>(h->vxs.ptr());
  }
  CSS_CHECK(vmm_status(vmm_reserve(&h->valive, h->device, (size_t)words_for(mr) * 4)));
  CSS_CHECK(vmm_status(vmm_reserve(&h->vmask, h->device, (size_t)words_for(mr) * 4)));
  h->x = reinterpret_cast<float*>(h->vx.ptr());
  h->xb = reinterpret_cast<__nv_bfloat16*>(h->vxb.ptr());
  h->alive = reinterpret_cast<uint32_t*>(h->valive.ptr());
  h->mask = reinterpret_cast<uint32_t*>(h->vmask.ptr());
  return CSS_OK;
}

// Grow every per-row array to `new_cap` rows: more physical memory is mapped behind the same
// addresses (vmm.h); existing rows are not touched, nothing is copied.
int grow(css_index* h, int64_t new_cap) {
  if (new_cap <= h->capacity) return CSS_OK;
  new_cap = (new_cap + 31) / 32 * 32;   // keep row-bitmask words whole
  CSS_CHECK(reserve_ranges(h));
  const size_t d = (size_t)h->dim;
  const int64_t old_cap = h->capacity;
  CSS_CHECK(vmm_status(vmm_grow(&h->vx, (size_t)new_cap * d * 4)));
  CSS_CHECK(vmm_status(vmm_grow(&h->vxb, (size_t)new_cap * d * 2)));
  if (h->xq) {
    CSS_CHECK(vmm_status(vmm_grow(&h->vxq, (size_t)new_cap * d)));
    CSS_CHECK(vmm_status(vmm_grow(&h->vxs, (size_t)new_cap * 4)));
  }
  CSS_CHECK(vmm_status(vmm_grow(&h->valive, (size_t)words_for(new_cap) * 4)));
  CSS_CHECK(vmm_status(vmm_grow(&h->vmask, (size_t)words_for(new_cap) * 4)));
  cudaStream_t st = h->stream;
  const int64_t w0 = words_for(old_cap), w1 = words_for(new_cap);
  CSS_CUDA(cudaMemsetAsync(h->alive + w0, 0, (size_t)(w1 - w0) * 4, st));
  CSS_CUDA(cudaMemsetAsync(h->mask + w0, 0, (size_t)(w1 - w0) * 4, st));
  for (int c = 0; c < CSS_MAX_COLUMNS; ++c) {
    if (!h->cols[c]) continue;
    CSS_CHECK(vmm_status(vmm_grow(&h->vcols[c], (size_t)new_cap * 4)));
    const int64_t n = new_cap - old_cap;
    fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->cols[c] + old_cap, n, CSS_NULL_VALUE);
    CSS_LAUNCHED();
  }
  CSS_CUDA(cudaStreamSynchronize(st));
  h->capacity = new_cap;
  return CSS_OK;
}

int ensure_room(css_index* h, int64_t extra) {
  int64_t need = h->ntotal + extra;
  if (need <= h->capacity) return CSS_OK;
  // geometric growth bounds the number of mappings; it costs no copy and no transient memory
  int64_t cap = std::max<int64_t>(need, std::max<int64_t>(1024, h->capacity + h->capacity / 4));
  return grow(h, cap);
}

int ensure_column(css_index* h, int column) {
  if (h->cols[column]) return CSS_OK;
  CSS_CHECK(reserve_ranges(h));
  CSS_CHECK(vmm_status(vmm_reserve(&h->vcols[column], h->device, (size_t)(h->vmask.reserved / 4 * 32) * 4)));
  if (h->capacity > 0) {
    CSS_CHECK(vmm_status(vmm_grow(&h->vcols[column], (size_t)h->capacity * 4)));
    int32_t* c = reinterpret_cast<int32_t*>(h->vcols[column].ptr());
    fill_i32_kernel<<<(unsigned)((h->capacity + 255) / 256), 256, 0, h->stream>>>(c, h->capacity, CSS_NULL_VALUE);
    CSS_LAUNCHED();
  }
  h->cols[column] = reinterpret_cast<int32_t*>(h->vcols[column].ptr());
  return CSS_OK;
}

// Rows [row0, row0+n) of x hold (or, with src != their place, receive) new data: normalise / copy,
// bf16 shadow, norm and rounding-error maxima.
int finish_rows(css_index* h, const float* src_dev, int64_t row0, int64_t n, int normalize, cudaStream_t st) {
  const int warps_per_block = 8;
  int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
  const size_t off = (size_t)row0 * h->dim;
  append_rows_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, st>>>(
      src_dev, n, h->dim, normalize, h->x + off, h->xb + off, h->max_norm_dev, h->max_err_dev,
      h->xq ? h->xq + off : nullptr, h->xs ? h->xs + row0 : nullptr, h->max_err8_dev);
  CSS_LAUNCHED();
  return CSS_OK;
}

int mark_alive(css_index* h, int64_t row0, int64_t n, cudaStream_t st) {
  int64_t w0 = row0 >> 5, w1 = (row0 + n - 1) >> 5;
  int64_t nw = w1 - w0 + 1;
  set_bits_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(h->alive, row0, n, 1);
  CSS_LAUNCHED();
  return CSS_OK;
}

// Rows [ntotal, ntotal+n) were produced in `src_dev`; normalise/copy + bookkeeping.
int append_from_device(css_index* h, const float* src_dev, int64_t n, int normalize, cudaStream_t st) {
  CSS_CHECK(finish_rows(h, src_dev, h->ntotal, n, normalize, st));
  return mark_alive(h, h->ntotal, n, st);
}

template <int KPL, int METRIC>
int launch_scan_kd(css_index* h, const ScanParams& p, int nq, cudaStream_t st) {
  const bool d768 = (h->dim == 768);
  size_t smem = sizeof(KeyId) * kMergeCap + (d768 ? 0 : (size_t)h->dim * 4);
  dim3 grid((unsigned)h->scan_blocks, (unsigned)nq);
  if (d768) {
    auto kern = scan_topk_kernel<KPL, METRIC, true>;
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (p.pdl_wait) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = grid;
      cfg.blockDim = dim3(kScanThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      CSS_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    } else {
      kern<<<grid, kScanThreads, smem, st>>>(p);
    }
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
int build_filter_params(css_index* h, const css_filter* f, FilterParams* fp, cudaStream_t st, size_t pinned_front) {
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
    // behind the caller's own staging area (query / results of css_index_search): no wait between the two uses
    CSS_CHECK(ensure_pinned(h, pinned_front + words * 4));
    uint32_t* stage = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(h->pinned) + pinned_front);
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
    // pageable source: the driver has staged the bytes when the call returns
    CSS_CUDA(cudaMemcpyAsync(h->rowmask_scratch, f->row_mask, (size_t)w * 4, cudaMemcpyHostToDevice, st));
    fp->row_mask = h->rowmask_scratch;
  }
  return CSS_OK;
}

}  // namespace

namespace css {

// Evaluate `f` into h->mask.  *mask_out = nullptr when every row passes trivially; *ignore_alive_out
// tells the caller that the alive bits must NOT be substituted for a null mask (css_filter.ignore_alive:
// HybridStorage's reference mode wants the orphans in its global top-100 window).
// Serialised form of a filter (clauses + set bits) for the "same filter as last time?" test.
static bool filter_key(const css_filter* f, std::vector<uint32_t>* key) {
  key->clear();
  if (f && f->row_mask) return false;   // an explicit row mask is not compared (ntotal / 8 bytes): always evaluated
  key->push_back(f ? (uint32_t)f->n_clauses : 0u);
  key->push_back(f ? (uint32_t)f->ignore_alive : 0u);
  if (!f) return true;
  for (int i = 0; i < f->n_clauses; ++i) {
    const css_clause& c = f->clauses[i];
    key->push_back((uint32_t)c.column);
    key->push_back((uint32_t)c.kind);
    key->push_back((uint32_t)c.lo);
    key->push_back((uint32_t)c.hi);
    key->push_back((uint32_t)c.set_nbits);
    if (c.kind == CSS_CLAUSE_SET && c.set_bits)
      key->insert(key->end(), c.set_bits, c.set_bits + (size_t)(c.set_nbits + 31) / 32);
  }
  return true;
}

int eval_filter(css_index* h, const css_filter* f, const uint32_t** mask_out, int64_t* n_pass,
                bool need_count, cudaStream_t st, bool* ignore_alive_out, size_t pinned_front) {
  const bool ignore_alive = f && f->ignore_alive;
  if (ignore_alive_out) *ignore_alive_out = ignore_alive;
  const bool trivial = (!f || (f->n_clauses == 0 && !f->row_mask)) && (!h->any_dead || ignore_alive);
  if (trivial && !need_count) {
    *mask_out = nullptr;
    if (n_pass) *n_pass = h->ntotal;
    return CSS_OK;
  }
  // no clauses, rows deleted: the alive bits ARE the mask (no evaluation: 27 us per query on a 1 M-row index)
  if ((!f || (f->n_clauses == 0 && !f->row_mask)) && !ignore_alive && !need_count && h->ntotal > 0) {
    *mask_out = h->alive;
    return CSS_OK;
  }
  if (h->ntotal == 0) {
    *mask_out = h->mask;
    if (n_pass) *n_pass = 0;
    return CSS_OK;
  }
  // the mask of an unchanged filter over an unchanged index is still in h->mask (evaluated on the handle's own
  // stream order; a different stream would have to wait for that evaluation, so the shortcut is for st == h->stream)
  std::vector<uint32_t> key;
  const bool cacheable = filter_key(f, &key) && st == h->stream && f != nullptr && f->n_clauses > 0;
  if (cacheable && h->mask_version == h->version && key == h->mask_key && (!n_pass || h->mask_n_pass >= 0)) {
    *mask_out = h->mask;
    if (n_pass) *n_pass = h->mask_n_pass;
    return CSS_OK;
  }
  h->mask_version = ~0ull;
  FilterParams fp;
  CSS_CHECK(build_filter_params(h, f, &fp, st, pinned_front));
  if (n_pass) CSS_CUDA(cudaMemsetAsync(h->n_pass_dev, 0, sizeof(unsigned long long), st));
  else fp.n_pass = nullptr;
  int64_t blocks = (h->ntotal + 1023) / 1024;   // four rows per thread
  filter_mask_kernel<<<(unsigned)blocks, 256, 0, st>>>(fp);
  CSS_LAUNCHED();
  *mask_out = h->mask;
  if (n_pass) {
    unsigned long long c = 0;
    CSS_CUDA(cudaMemcpyAsync(&c, h->n_pass_dev, sizeof(c), cudaMemcpyDeviceToHost, st));
    CSS_CUDA(cudaStreamSynchronize(st));
    *n_pass = (int64_t)c;
  }
  if (cacheable) {
    h->mask_key.swap(key);
    h->mask_version = h->version;
    h->mask_n_pass = n_pass ? *n_pass : -1;
  }
  return CSS_OK;
}

int ensure_pinned(css_index* h, size_t bytes) {
  if (bytes <= h->pinned_bytes) return CSS_OK;
  if (h->pinned) {
    cudaDeviceSynchronize();  // an async copy may still be reading the old block
    cudaFreeHost(h->pinned);
  }
  h->pinned = nullptr;
  h->pinned_bytes = 0;
  h->pinned_dev = nullptr;
  size_t want = std::max<size_t>(bytes, (size_t)1 << 20);
  cudaError_t e = cudaHostAlloc(&h->pinned, want, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
    return CSS_ERR_OOM;
  }
  if (cudaHostGetDevicePointer(&h->pinned_dev, h->pinned, 0) != cudaSuccess) {
    (void)cudaGetLastError();
    h->pinned_dev = nullptr;   // no mapped view: the copy-based path is used
  }
  h->pinned_bytes = want;
  return CSS_OK;
}

static void free_scratch(css_scan_scratch* sc) {
  batched_release(sc);
  cudaFree(sc->q_dev);
  cudaFree(sc->part);
  cudaFree(sc->part_exact);
  cudaFree(sc->trace);
  cudaFree(sc->ticket);
  cudaFree(sc->D_dev);
  cudaFree(sc->ovf_list);
  *sc = css_scan_scratch();
}

// Scratch of the searches issued on stream `st` (created on first use, grown geometrically).
int get_scratch(css_index* h, cudaStream_t st, int nq, css_scan_scratch** out) {
  css_scan_scratch& sc = h->scratch[st];
  *out = &sc;
  if (nq <= sc.max_nq) return CSS_OK;
  const int want = std::max(nq, std::max(16, sc.max_nq * 2));
  if (sc.max_nq > 0) CSS_CUDA(cudaStreamSynchronize(st));   // work in flight may still use the old buffers
  void* keep_batched = sc.batched;
  sc.batched = nullptr;
  free_scratch(&sc);
  sc.batched = keep_batched;
  CSS_CHECK(dev_alloc(&sc.q_dev, (size_t)want * h->dim));
  CSS_CHECK(dev_alloc(&sc.part, (size_t)want * h->scan_blocks * CSS_MAX_K));
  CSS_CHECK(dev_alloc(&sc.part_exact, (size_t)want * h->scan_blocks * CSS_MAX_K));
  CSS_CHECK(dev_alloc(&sc.ticket, (size_t)want * 2));   // tickets, then the unit cursors of the int8 sweep
  // scores, then the ids of the same call, then the overflow count: one D2H returns all three
  CSS_CHECK(dev_alloc(&sc.D_dev, (size_t)want * CSS_MAX_K * 3 + 16));
  sc.ovf_count = reinterpret_cast<int*>(sc.D_dev + (size_t)want * CSS_MAX_K * 3 + 8);
  CSS_CHECK(dev_alloc(&sc.ovf_list, (size_t)want));
  CSS_CUDA(cudaMemsetAsync(sc.ticket, 0, (size_t)want * 2 * sizeof(unsigned int), st));
  CSS_CUDA(cudaMemsetAsync(sc.ovf_count, 0, sizeof(int), st));
  sc.max_nq = want;
  return CSS_OK;
}

IdMap index_idmap(const css_index* h, int64_t id_offset) {
  IdMap m;
  m.offset = id_offset;
  m.shift = h->id_shift;
  m.ndev = h->id_ndev;
  m.shard = h->id_shard;
  return m;
}

static void fill_common(css_index* h, css_scan_scratch* sc, ScanParams* p, const float* q_dev, const uint32_t* mask_dev,
                        const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev) {
  memset(p, 0, sizeof(*p));
  p->x = h->x;
  p->n = h->ntotal;
  p->d = h->dim;
  p->q = q_dev;
  p->mask = mask_dev;
  p->part = sc->part;
  p->part_exact = sc->part_exact;
  p->ticket = sc->ticket;
  p->cursor = sc->ticket + sc->max_nq;
  p->idmap = idmap;
  p->D = D_dev;
  p->I = I_dev;
  p->max_norm = h->max_norm_dev;
  p->max_err = h->max_err_dev;
  p->max_err8 = h->max_err8_dev;
  p->ovf_list = sc->ovf_list;
  p->ovf_count = sc->ovf_count;
  p->stats_dev = h->stats_dev;
  p->stats_host = h->stats_host_devptr;
  p->done_flag = sc->done_flag;
  p->trace = sc->trace;
  p->done_seq = sc->done_seq;
  if (ex) p->ex = *ex;
}

// Two-phase exact scan (inner product, d = 768, k <= 32, with or without a row mask): phase 1 streams the bf16 shadow
// rows -- half the bytes of the fp32 corpus -- and leaves the best rows of every scan block by that score; the
// last CTA to finish (two_phase_finish) proves that the true top-k lies inside those lists, re-scores the
// candidates in fp32 with the arithmetic of the fp32 scan and emits the exact result -- one launch; queries it
// cannot prove are queued on the device and re-run by the fp32 scan (scan_fallback), so the answer is always
// the exact one.  CSS_SCAN_BF16=0 disables it; CSS_SCAN_LIST=32|64 fixes the per-block list length (default 32
// for k <= 16, 64 above); CSS_SCAN_INTERLEAVE=0 walks contiguous row ranges per warp instead of dealing 8-row
// units block-cyclically.
template <int KPL, int SH>
static int launch_phase1_as(const ScanParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = scan_topk_kernel<KPL, CSS_METRIC_INNER_PRODUCT, true, SH>;
  // (the dynamic shared-memory limit of every scan kernel was raised when the index was created: preload_index_kernels)
  kern<<<grid, kScanThreads, smem, st>>>(p);
  CSS_LAUNCHED();
  return CSS_OK;
}

// tier: 1 = bf16 shadow rows (1536 B per row), 2 = int8 shadow rows (768 + 4 B per row)
static int launch_phase1(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                         const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev, int no_merge,
                         cudaStream_t st, int tier) {
  const int list_env = options().scan_list.load();
  const int interleave = options().scan_interleave.load();
  const int kp = list_env == 32 || list_env == 64 ? list_env : (k <= 16 ? 32 : 64);
  ScanParams p;
  fill_common(h, sc, &p, q_dev, mask_dev, idmap, ex, D_dev, I_dev);
  p.xb = h->xb;
  p.xq = h->xq;
  p.xs = h->xs;
  p.k = kp;
  p.k_out = k;
  p.no_merge = no_merge;
  p.interleave = interleave;
  p.zero_on_entry = sc->ovf_count;
  p.mask_dense = (mask_dev != nullptr && mask_dev == h->alive) ? 1 : 0;   // deletions only: sweep every row, drop the dead
  const size_t smem = sizeof(KeyId) * kMergeCap;
  const dim3 grid((unsigned)h->scan_blocks, (unsigned)nq);
  if (tier == 2) {
    CSS_REQUIRE(h->xq != nullptr, "this index has no int8 shadow copy");
    const size_t smem8 = std::max(smem, (size_t)kI8SmemBytes);   // the sweep's ring of bulk copies; the lists alias it
    if (mask_dev != nullptr && !p.mask_dense)   // filtered: the gather instantiation
      return kp == 32 ? launch_phase1_as<1, 3>(p, grid, smem8, st) : launch_phase1_as<2, 3>(p, grid, smem8, st);
    return kp == 32 ? launch_phase1_as<1, 2>(p, grid, smem8, st) : launch_phase1_as<2, 2>(p, grid, smem8, st);
  }
  return kp == 32 ? launch_phase1_as<1, 1>(p, grid, smem, st) : launch_phase1_as<2, 1>(p, grid, smem, st);
}

int scan_fallback(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                  const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev, cudaStream_t st, bool pdl_ok) {
  ScanParams f;
  fill_common(h, sc, &f, q_dev, mask_dev, idmap, ex, D_dev, I_dev);
  f.k = k;
  f.k_out = k;
  f.qlist = sc->ovf_list;
  f.qcount = sc->ovf_count;
  f.pdl_wait = options().scan_pdl.load() != 0 && pdl_ok ? 1 : 0;
  // an empty list costs one idle launch; few slices: unproven queries are rare
  return launch_scan_m<CSS_METRIC_INNER_PRODUCT>(h, f, std::min(nq, 8), st);
}

// Which tier answers this call: 2 = int8 shadow sweep, 1 = bf16 shadow sweep, 0 = plain fp32 sweep.  The kernels
// mirror {queries, unproven} into mapped host memory; when more than half of the recent queries of a tier could not
// be proven (a corpus whose rows sit within that tier's error bound of each other), its phase 1 is wasted work and
// the next 4096 scan queries skip the tier, after which it is probed again.  Exactness never depends on this choice.
static int scan_tier(css_index* h, int nq) {
  if (!options().scan_bf16.load()) return h->last_tier = 0;
  const int top = (options().scan_int8.load() && h->xq) ? 2 : 1;
  if (!options().scan_adaptive.load() || !h->stats_host) return h->last_tier = top;
  const unsigned q = h->stats_host[0], u = h->stats_host[1];
  if (h->last_tier > 0 && q - h->seen_q >= 64u) {
    const bool bad = (u - h->seen_u) * 2u > (q - h->seen_q);
    h->seen_q = q;
    h->seen_u = u;
    if (bad) h->tier_ban[h->last_tier] = 4096;
  }
  int t = top;
  while (t > 0 && h->tier_ban[t] > 0) --t;
  for (int i = 1; i <= 2; ++i)
    if (h->tier_ban[i] > 0) h->tier_ban[i] -= nq;
  if (t != h->last_tier) {   // a tier is judged on its own queries only
    h->seen_q = q;
    h->seen_u = u;
    h->last_tier = t;
  }
  return t;
}

int scan_search(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev, cudaStream_t st,
                bool defer_fallback, bool* two_phase_used) {
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  if (two_phase_used) *two_phase_used = false;
  int tier = 0;
  if (h->metric == CSS_METRIC_INNER_PRODUCT && h->dim == 768 && k <= kTwoPhaseMaxK && nq <= 64 && h->ntotal > 0 &&
      h->scan_blocks <= kTwoPhaseMaxBlocks)
    tier = scan_tier(h, nq);
  if (tier > 0) {
    // Exchange + several queries in one launch: a CTA that waited for its peers inside phase 1 would hold up the
    // fallback launch behind it, which a peer's phase 1 may in turn be waiting for (query u unproven here, query v
    // unproven there).  So with nq > 1 the producers only publish and the fallback launch awaits + merges.
    ExchangeDev xd;
    if (ex) {
      xd = *ex;
      xd.deferred = nq > 1 ? 1 : 0;
      xd.nq = nq;
      ex = &xd;
    }
    CSS_CHECK(launch_phase1(h, sc, q_dev, nq, k, mask_dev, idmap, ex, D_dev, I_dev, 0, st, tier));
    if (two_phase_used) *two_phase_used = true;
    if (defer_fallback) return CSS_OK;   // the caller reads the overflow count with the result
    return scan_fallback(h, sc, q_dev, nq, k, mask_dev, idmap, ex, D_dev, I_dev, st, /*pdl_ok=*/!(ex && ex->no_pdl));
  }
  // gridDim.y is limited to 65535; chunk the query batch
  const int chunk = 4096;
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    int nqc = std::min(chunk, nq - q0);
    ScanParams p;
    fill_common(h, sc, &p, q_dev + (size_t)q0 * h->dim, mask_dev, idmap, ex, D_dev + (size_t)q0 * k, I_dev + (size_t)q0 * k);
    p.k = k;
    p.k_out = k;
    if (h->metric == CSS_METRIC_INNER_PRODUCT)
      CSS_CHECK((launch_scan_m<CSS_METRIC_INNER_PRODUCT>(h, p, nqc, st)));
    else
      CSS_CHECK((launch_scan_m<CSS_METRIC_L2>(h, p, nqc, st)));
  }
  return CSS_OK;
}

// One search on device buffers, any nq: tensor-core path for batches, streaming scan otherwise.
int search_on_device(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* m,
                     const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev, cudaStream_t st,
                     bool defer_fallback, bool* two_phase_used) {
  if (two_phase_used) *two_phase_used = false;
  if (!ex && nq >= CSS_BATCH_MIN_NQ && h->metric == CSS_METRIC_INNER_PRODUCT && h->dim % 64 == 0 && h->ntotal >= 65536)
    return batched_search(h, sc, q_dev, nq, k, m, idmap, D_dev, I_dev, st);
  return scan_search(h, sc, q_dev, nq, k, m, idmap, ex, D_dev, I_dev, st, defer_fallback, two_phase_used);
}

// Load every kernel a search can launch on the current device now.  CUDA loads kernels lazily, on first launch, and
// that load waits for the device to go idle: during a multi-device search a scan kernel may be spinning for a
// peer's list while the host -- which has yet to launch that peer -- sits in the lazy load of the next kernel.
template <typename K>
static int preload_kernel(K kern, size_t smem) {
  cudaFuncAttributes a;
  CSS_CUDA(cudaFuncGetAttributes(&a, kern));
  if (smem) {
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  return CSS_OK;
}
template <int KPL>
static int preload_scan_kpl(size_t smem, size_t smem_generic) {
  CSS_CHECK(preload_kernel(scan_topk_kernel<KPL, CSS_METRIC_INNER_PRODUCT, true>, smem));
  CSS_CHECK(preload_kernel(scan_topk_kernel<KPL, CSS_METRIC_L2, true>, smem));
  CSS_CHECK(preload_kernel(scan_topk_kernel<KPL, CSS_METRIC_INNER_PRODUCT, false>, smem_generic));
  CSS_CHECK(preload_kernel(scan_topk_kernel<KPL, CSS_METRIC_L2, false>, smem_generic));
  return CSS_OK;
}
static int preload_index_kernels(int dim) {
  const size_t smem = sizeof(KeyId) * kMergeCap, smem_generic = smem + (size_t)dim * 4;
  CSS_CHECK(preload_scan_kpl<1>(smem, smem_generic));
  CSS_CHECK(preload_scan_kpl<2>(smem, smem_generic));
  CSS_CHECK(preload_scan_kpl<4>(smem, smem_generic));
  CSS_CHECK(preload_kernel(scan_topk_kernel<1, CSS_METRIC_INNER_PRODUCT, true, 1>, smem));
  CSS_CHECK(preload_kernel(scan_topk_kernel<2, CSS_METRIC_INNER_PRODUCT, true, 1>, smem));
  CSS_CHECK(preload_kernel(scan_topk_kernel<1, CSS_METRIC_INNER_PRODUCT, true, 2>, std::max(smem, (size_t)kI8SmemBytes)));
  CSS_CHECK(preload_kernel(scan_topk_kernel<2, CSS_METRIC_INNER_PRODUCT, true, 2>, std::max(smem, (size_t)kI8SmemBytes)));
  CSS_CHECK(preload_kernel(scan_topk_kernel<1, CSS_METRIC_INNER_PRODUCT, true, 3>, std::max(smem, (size_t)kI8SmemBytes)));
  CSS_CHECK(preload_kernel(scan_topk_kernel<2, CSS_METRIC_INNER_PRODUCT, true, 3>, std::max(smem, (size_t)kI8SmemBytes)));
  CSS_CHECK(preload_kernel(filter_mask_kernel, 0));
  CSS_CHECK(preload_kernel(append_rows_kernel, 0));
  CSS_CHECK(preload_kernel(set_bits_kernel, 0));
  CSS_CHECK(preload_kernel(fill_i32_kernel, 0));
  return CSS_OK;
}

int index_create_single(int dim, int metric, int device, css_index** out) {
  CSS_CHECK(ensure_device(device));
  DeviceGuard g(device);
  CSS_CHECK(preload_index_kernels(dim));
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
  void* sh = nullptr;
  if (dev_alloc(&h->n_pass_dev, 1) != CSS_OK || dev_alloc(&h->max_norm_dev, 1) != CSS_OK ||
      dev_alloc(&h->max_err_dev, 1) != CSS_OK || dev_alloc(&h->stats_dev, 2) != CSS_OK ||
      dev_alloc(&h->max_err8_dev, 1) != CSS_OK || cudaMemset(h->max_err8_dev, 0, sizeof(float)) != cudaSuccess ||
      cudaMemset(h->max_norm_dev, 0, sizeof(float)) != cudaSuccess ||
      cudaMemset(h->max_err_dev, 0, sizeof(float)) != cudaSuccess ||
      cudaMemset(h->stats_dev, 0, 2 * sizeof(unsigned)) != cudaSuccess ||
      cudaHostAlloc(&sh, 2 * sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("index bookkeeping allocation failed");
    cudaFree(h->n_pass_dev);
    cudaFree(h->max_norm_dev);
    cudaFree(h->max_err_dev);
    cudaFree(h->max_err8_dev);
    cudaFree(h->stats_dev);
    cudaStreamDestroy(h->stream);
    delete h;
    return CSS_ERR_OOM;
  }
  h->stats_host = reinterpret_cast<volatile unsigned*>(sh);
  h->stats_host[0] = 0;
  h->stats_host[1] = 0;
  void* dp = nullptr;
  if (cudaHostGetDevicePointer(&dp, sh, 0) == cudaSuccess) h->stats_host_devptr = reinterpret_cast<unsigned*>(dp);
  else (void)cudaGetLastError();
  *out = h;
  return CSS_OK;
}

void index_destroy_single(css_index* h) {
  DeviceGuard g(h->device);
  cudaStreamSynchronize(h->stream);
  cudaDeviceSynchronize();
  for (auto& kv : h->scratch) free_scratch(&kv.second);
  h->scratch.clear();
  vmm_release(&h->vx);
  vmm_release(&h->vxb);
  vmm_release(&h->vxq);
  vmm_release(&h->vxs);
  vmm_release(&h->valive);
  vmm_release(&h->vmask);
  for (int c = 0; c < CSS_MAX_COLUMNS; ++c) vmm_release(&h->vcols[c]);
  cudaFree(h->set_scratch);
  cudaFree(h->rowmask_scratch);
  cudaFree(h->n_pass_dev);
  cudaFree(h->max_norm_dev);
  cudaFree(h->max_err_dev);
  cudaFree(h->max_err8_dev);
  cudaFree(h->stats_dev);
  cudaFree(h->ids_scratch);
  if (h->stats_host) cudaFreeHost(const_cast<unsigned*>(h->stats_host));
  if (h->pinned) cudaFreeHost(h->pinned);
  cudaStreamDestroy(h->stream);
  delete h;
}

// Overwrite / append rows [row0, row0+n) from HOST memory without a staging buffer: the bytes are
// copied straight into x and finished in place (normalisation, bf16 shadow, maxima).  Asynchronous
// on the handle's stream; the caller synchronises.  Rows beyond ntotal must fit the capacity.
int put_rows_host_async(css_index* h, const float* x_host, int64_t row0, int64_t n, int normalize) {
  h->version++;
  float* dst = h->x + (size_t)row0 * h->dim;
  CSS_CUDA(cudaMemcpyAsync(dst, x_host, (size_t)n * h->dim * 4, cudaMemcpyHostToDevice, h->stream));
  return finish_rows(h, dst, row0, n, normalize, h->stream);
}

int single_add(css_index* h, const float* x_host, int64_t n, int normalize, bool sync) {
  CSS_REQUIRE(h->ntotal + n < ((int64_t)1 << 31) - 64, "index would exceed 2^31 rows per shard");
  CSS_CHECK(ensure_room(h, n));
  CSS_CHECK(put_rows_host_async(h, x_host, h->ntotal, n, normalize));
  CSS_CHECK(mark_alive(h, h->ntotal, n, h->stream));
  h->ntotal += n;
  h->version++;
  if (sync) CSS_CUDA(cudaStreamSynchronize(h->stream));
  return CSS_OK;
}

int single_reset(css_index* h) {
  h->version++;
  h->ntotal = 0;
  h->any_dead = false;
  CSS_CUDA(cudaMemsetAsync(h->max_norm_dev, 0, sizeof(float), h->stream));
  CSS_CUDA(cudaMemsetAsync(h->max_err_dev, 0, sizeof(float), h->stream));
  CSS_CUDA(cudaMemsetAsync(h->max_err8_dev, 0, sizeof(float), h->stream));
  if (h->capacity > 0) {
    CSS_CUDA(cudaMemsetAsync(h->alive, 0, (size_t)words_for(h->capacity) * 4, h->stream));
    for (int c = 0; c < CSS_MAX_COLUMNS; ++c) {
      if (!h->cols[c]) continue;
      int64_t blocks = (h->capacity + 255) / 256;
      fill_i32_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(h->cols[c], h->capacity, CSS_NULL_VALUE);
      CSS_LAUNCHED();
    }
  }
  CSS_CUDA(cudaStreamSynchronize(h->stream));
  return CSS_OK;
}

int single_set_column(css_index* h, int column, const int32_t* values_host, int64_t start, int64_t n, bool sync) {
  h->version++;
  CSS_CHECK(ensure_column(h, column));
  CSS_CUDA(cudaMemcpyAsync(h->cols[column] + start, values_host, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  if (sync) CSS_CUDA(cudaStreamSynchronize(h->stream));
  return CSS_OK;
}

int single_set_alive_ids(css_index* h, const int64_t* ids_host, int64_t n, int alive) {
  if (n == 0) return CSS_OK;
  h->version++;
  if (n > h->ids_scratch_n) {
    cudaFree(h->ids_scratch);
    h->ids_scratch_n = 0;
    const int64_t want = std::max<int64_t>(n, 4096);
    CSS_CHECK(dev_alloc(&h->ids_scratch, (size_t)want));
    h->ids_scratch_n = want;
  }
  CSS_CUDA(cudaMemcpyAsync(h->ids_scratch, ids_host, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  set_bits_by_id_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->alive, h->ids_scratch, n, h->ntotal,
                                                                            alive ? 1 : 0);
  CSS_LAUNCHED();
  CSS_CUDA(cudaStreamSynchronize(h->stream));
  if (!alive) h->any_dead = true;
  return CSS_OK;
}

int single_grow(css_index* h, int64_t cap) { return grow(h, cap); }
int single_ensure_room(css_index* h, int64_t extra) { return ensure_room(h, extra); }
int single_mark_alive(css_index* h, int64_t row0, int64_t n) {
  h->version++;
  return mark_alive(h, row0, n, h->stream);
}

}  // namespace css

#define CSS_IS_COMPOSITE(h) (!(h)->shards.empty())

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
  return index_create_single(dim, metric, device, out);
}

int css_index_destroy(css_index* h) {
  if (!h) return CSS_OK;
  if (CSS_IS_COMPOSITE(h)) return sharded_destroy(h);
  index_destroy_single(h);
  return CSS_OK;
}

int css_index_dim(const css_index* h) { return h ? h->dim : CSS_ERR_INVALID; }
int css_index_metric(const css_index* h) { return h ? h->metric : CSS_ERR_INVALID; }
int64_t css_index_ntotal(const css_index* h) {
  if (!h) return (int64_t)CSS_ERR_INVALID;
  return CSS_IS_COMPOSITE(h) ? h->composite_ntotal : h->ntotal;
}
int64_t css_index_capacity(const css_index* h) {
  if (!h) return (int64_t)CSS_ERR_INVALID;
  if (!CSS_IS_COMPOSITE(h)) return h->capacity;
  int64_t c = 0;
  for (const css_index* s : h->shards) c += s->capacity;
  return c;
}
int css_index_n_devices(const css_index* h) {
  if (!h) return CSS_ERR_INVALID;
  return CSS_IS_COMPOSITE(h) ? (int)h->shards.size() : 1;
}

int css_index_reserve(css_index* h, int64_t capacity) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(capacity >= 0, "capacity out of range");
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_reserve(h, capacity);
  CSS_REQUIRE(capacity < ((int64_t)1 << 31) - 64, "capacity out of range (2^31 rows per shard)");
  DeviceGuard g(h->device);
  return grow(h, capacity);
}

int css_index_reset(css_index* h) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_reset(h);
  DeviceGuard g(h->device);
  return single_reset(h);
}

int css_index_compact(css_index* h, const int64_t* keep_ids_host, int64_t n_keep) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(n_keep >= 0 && (n_keep == 0 || keep_ids_host != nullptr), "bad keep list");
  std::lock_guard<std::mutex> lk(h->mu);
  const int64_t nt = CSS_IS_COMPOSITE(h) ? h->composite_ntotal : h->ntotal;
  CSS_REQUIRE(n_keep <= nt, "keep list longer than the index (%lld > %lld)", (long long)n_keep, (long long)nt);
  for (int64_t i = 0; i < n_keep; ++i) {
    const int64_t id = keep_ids_host[i];
    CSS_REQUIRE(id >= 0 && id < nt && (i == 0 || id > keep_ids_host[i - 1]),
                "keep ids must be strictly ascending row ids (entry %lld = %lld)", (long long)i, (long long)id);
  }
  if (CSS_IS_COMPOSITE(h)) return sharded_compact(h, keep_ids_host, n_keep);
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const int d = h->dim;
  const int64_t win = 65536;   // rows per window: 200 MB of fp32 staging at d = 768
  int64_t* ids_dev = nullptr;
  float* sx = nullptr;
  __nv_bfloat16* sb = nullptr;
  int8_t* sq = nullptr;
  float* ss = nullptr;
  int32_t* sc = nullptr;
  uint32_t* sw = nullptr;
  auto cleanup = [&]() {
    cudaFree(ids_dev); cudaFree(sx); cudaFree(sb); cudaFree(sq); cudaFree(ss); cudaFree(sc); cudaFree(sw);
  };
  if (n_keep > 0) {
    const int64_t w = std::min(win, n_keep);
    if (cudaMalloc(&ids_dev, (size_t)n_keep * 8) != cudaSuccess || cudaMalloc(&sx, (size_t)w * d * 4) != cudaSuccess ||
        cudaMalloc(&sb, (size_t)w * d * 2) != cudaSuccess || cudaMalloc(&sc, (size_t)w * 4) != cudaSuccess ||
        (h->xq && (cudaMalloc(&sq, (size_t)w * d) != cudaSuccess || cudaMalloc(&ss, (size_t)w * 4) != cudaSuccess)) ||
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
      gather_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(h->x, h->xb, ids_dev + i0, n, d, sx, sb, h->xq, h->xs, sq,
                                                                  ss);
      css::g_launches.fetch_add(1, std::memory_order_relaxed);
      ce = cudaMemcpyAsync(h->x + i0 * d, sx, (size_t)n * d * 4, cudaMemcpyDeviceToDevice, st);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->xb + i0 * d, sb, (size_t)n * d * 2, cudaMemcpyDeviceToDevice, st);
      if (ce == cudaSuccess && h->xq) ce = cudaMemcpyAsync(h->xq + i0 * d, sq, (size_t)n * d, cudaMemcpyDeviceToDevice, st);
      if (ce == cudaSuccess && h->xq) ce = cudaMemcpyAsync(h->xs + i0, ss, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
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
  h->version++;
  h->any_dead = true;   // conservative: the scan keeps honouring the alive bits
  return CSS_OK;
}

int css_index_add(css_index* h, const float* x_host, int64_t n, int normalize, int64_t* first_id_out) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(n >= 0, "n < 0");
  CSS_REQUIRE(n == 0 || x_host != nullptr, "x_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_add(h, x_host, n, normalize, first_id_out);
  DeviceGuard g(h->device);
  if (first_id_out) *first_id_out = h->ntotal;
  if (n == 0) return CSS_OK;
  return single_add(h, x_host, n, normalize, /*sync=*/true);
}

int css_index_add_device(css_index* h, const float* x_dev, int64_t n, int normalize,
                         int64_t* first_id_out, void* stream) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(n >= 0, "n < 0");
  CSS_REQUIRE(n == 0 || x_dev != nullptr, "x_dev is NULL");
  if (CSS_IS_COMPOSITE(h)) {
    set_error("css_index_add_device: a multi-device index takes host rows (css_index_add)");
    return CSS_ERR_UNSUPPORTED;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CSS_REQUIRE(h->ntotal + n < ((int64_t)1 << 31) - 64, "index would exceed 2^31 rows per shard");
  if (first_id_out) *first_id_out = h->ntotal;
  if (n == 0) return CSS_OK;
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  if (h->ntotal + n > h->capacity) {
    // growth maps new memory on the handle's stream: order it after the caller's stream work
    CSS_CUDA(cudaStreamSynchronize(st));
    CSS_CHECK(ensure_room(h, n));
  }
  CSS_CHECK(append_from_device(h, x_dev, n, normalize, st));
  h->ntotal += n;
  h->version++;
  return CSS_OK;
}

int css_index_get_rows(css_index* h, int64_t start, int64_t n, float* out_host) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  const int64_t nt = CSS_IS_COMPOSITE(h) ? h->composite_ntotal : h->ntotal;
  CSS_REQUIRE(start >= 0 && n >= 0 && start + n <= nt, "row range out of bounds");
  if (n == 0) return CSS_OK;
  CSS_REQUIRE(out_host != nullptr, "out_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_get_rows(h, start, n, out_host);
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
  const int64_t nt = CSS_IS_COMPOSITE(h) ? h->composite_ntotal : h->ntotal;
  CSS_REQUIRE(start >= 0 && n >= 0 && start + n <= nt, "row range out of bounds");
  if (n == 0) return CSS_OK;
  CSS_REQUIRE(values_host != nullptr, "values_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_set_column(h, column, values_host, start, n);
  DeviceGuard g(h->device);
  return single_set_column(h, column, values_host, start, n, /*sync=*/true);
}

int css_index_set_alive(css_index* h, const uint8_t* alive_host, int64_t start, int64_t n) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  const int64_t nt = CSS_IS_COMPOSITE(h) ? h->composite_ntotal : h->ntotal;
  CSS_REQUIRE(start >= 0 && n >= 0 && start + n <= nt, "row range out of bounds");
  if (n == 0) return CSS_OK;
  CSS_REQUIRE(alive_host != nullptr, "alive_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_set_alive(h, alive_host, start, n);
  DeviceGuard g(h->device);
  uint8_t* tmp = nullptr;
  CSS_CHECK(dev_alloc(&tmp, (size_t)n));
  h->version++;
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
  if (rc == CSS_OK && memchr(alive_host, 0, (size_t)n) != nullptr) h->any_dead = true;
  return rc;
}

int css_index_set_alive_ids(css_index* h, const int64_t* ids_host, int64_t n, int alive) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(n >= 0 && (n == 0 || ids_host != nullptr), "bad id list");
  if (n == 0) return CSS_OK;
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_set_alive_ids(h, ids_host, n, alive);
  DeviceGuard g(h->device);
  return single_set_alive_ids(h, ids_host, n, alive);
}

int css_index_filter_mask(css_index* h, const css_filter* f, uint32_t* mask_out_host,
                          int64_t* n_pass_out) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  const int64_t nt = CSS_IS_COMPOSITE(h) ? h->composite_ntotal : h->ntotal;
  CSS_REQUIRE(mask_out_host != nullptr || nt == 0, "mask_out_host is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_filter_mask(h, f, mask_out_host, n_pass_out);
  DeviceGuard g(h->device);
  CSS_CHECK(ensure_device(h->device));
  const uint32_t* m = nullptr;
  int64_t n_pass = 0;
  // force evaluation even for the trivial filter so the mask is materialised
  css_filter empty;
  memset(&empty, 0, sizeof(empty));
  CSS_CHECK(eval_filter(h, f ? f : &empty, &m, &n_pass, /*need_count=*/true, h->stream, nullptr, 0));
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
  CSS_REQUIRE(!CSS_IS_COMPOSITE(h), "css_index_filter_mask_device: single-device indexes only");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  return eval_filter(h, f, mask_dev_out, n_pass_out, n_pass_out != nullptr, st, nullptr, 0);
}

static int debug_scan_tier(css_index* h, const float* q_dev, int nq, void* stream, int tier);

int css_debug_scan_bf16(css_index* h, const float* q_dev, int nq, void* stream) {
  return debug_scan_tier(h, q_dev, nq, stream, 1);
}
int css_debug_scan_int8(css_index* h, const float* q_dev, int nq, void* stream) {
  return debug_scan_tier(h, q_dev, nq, stream, 2);
}

static int debug_scan_tier(css_index* h, const float* q_dev, int nq, void* stream, int tier) {
  CSS_REQUIRE(h != nullptr && q_dev != nullptr, "NULL argument");
  CSS_REQUIRE(nq >= 1 && nq <= 64, "nq=%d outside [1, 64]", nq);
  CSS_REQUIRE(!CSS_IS_COMPOSITE(h) && h->metric == CSS_METRIC_INNER_PRODUCT && h->dim == 768 && h->ntotal > 0,
              "the two-phase scan needs a non-empty single-device 768-d inner-product index");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  css_scan_scratch* sc = nullptr;
  CSS_CHECK(get_scratch(h, st, nq, &sc));
  return launch_phase1(h, sc, q_dev, nq, 10, h->any_dead ? h->alive : nullptr, index_idmap(h, 0), nullptr, nullptr,
                       nullptr, /*no_merge=*/1, st, tier);
}

int css_index_search_device(css_index* h, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                            int64_t id_offset, float* D_dev, int64_t* I_dev, void* stream) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(nq >= 0, "nq < 0");
  if (nq == 0) return CSS_OK;
  CSS_REQUIRE(q_dev && D_dev && I_dev, "NULL device buffer");
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  CSS_REQUIRE(!CSS_IS_COMPOSITE(h), "css_index_search_device: single-device indexes only (a multi-device index "
              "is searched through css_index_search)");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const uint32_t* m = mask_dev;
  if (!m && h->any_dead) m = h->alive;
  css_scan_scratch* sc = nullptr;
  CSS_CHECK(get_scratch(h, st, std::min(nq, 1024), &sc));
  return search_on_device(h, sc, q_dev, nq, k, m, index_idmap(h, id_offset), nullptr, D_dev, I_dev, st,
                          /*defer_fallback=*/false, nullptr);
}

int css_index_search_exchange_device(css_index* h, css_exchange* ex, const float* q_dev, int nq, int k,
                                     const uint32_t* mask_dev, int64_t id_offset, float* D_dev, int64_t* I_dev,
                                     void* stream) {
  CSS_REQUIRE(h != nullptr && ex != nullptr, "NULL handle");
  CSS_REQUIRE(!CSS_IS_COMPOSITE(h), "css_index_search_exchange_device: single-device indexes only");
  CSS_REQUIRE(nq >= 1 && nq <= ex->max_nq, "nq=%d outside [1, %d] (larger batches: gather the lists with NCCL)", nq,
              ex->max_nq);
  CSS_REQUIRE(q_dev && D_dev && I_dev, "NULL device buffer");
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  CSS_REQUIRE(ex->connected && ex->device == h->device, "exchange not connected / on another device");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  ExchangeDev xd;
  CSS_CHECK(exchange_next(ex, &xd));
  const uint32_t* m = mask_dev;
  if (!m && h->any_dead) m = h->alive;
  css_scan_scratch* sc = nullptr;
  CSS_CHECK(get_scratch(h, st, nq, &sc));
  return scan_search(h, sc, q_dev, nq, k, m, index_idmap(h, id_offset), &xd, D_dev, I_dev, st,
                     /*defer_fallback=*/false, nullptr);
}

int css_index_search(css_index* h, const float* q_host, int nq, int k, const css_filter* filter,
                     float* D_host, int64_t* I_host) {
  CSS_REQUIRE(h != nullptr, "index is NULL");
  CSS_REQUIRE(nq >= 0, "nq < 0");
  if (nq == 0) return CSS_OK;
  CSS_REQUIRE(q_host && D_host && I_host, "NULL host buffer");
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  std::unique_lock<std::mutex> lk(h->mu);
  if (CSS_IS_COMPOSITE(h)) return sharded_search(h, q_host, nq, k, filter, D_host, I_host);
  CSS_CHECK(ensure_device(h->device));
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const size_t qbytes = (size_t)nq * h->dim * 4;
  const size_t dbytes = (size_t)nq * k * 4, ibytes = (size_t)nq * k * 8;
  // queries are staged and answered in slices of at most 1024 (the scratch is sized for one slice)
  if (nq > 1024) {
    lk.unlock();
    for (int q0 = 0; q0 < nq; q0 += 1024) {
      const int n = std::min(1024, nq - q0);
      CSS_CHECK(css_index_search(h, q_host + (size_t)q0 * h->dim, n, k, filter, D_host + (size_t)q0 * k,
                                 I_host + (size_t)q0 * k));
    }
    return CSS_OK;
  }
  css_scan_scratch* sc = nullptr;
  CSS_CHECK(get_scratch(h, st, nq, &sc));
  // pinned staging: [q | D | I | overflow count | clause bitsets of the filter]
  const size_t front = (qbytes + dbytes + ibytes + 128 + 63) / 64 * 64;
  CSS_CHECK(ensure_pinned(h, front));
  const uint32_t* m = nullptr;
  bool ignore_alive = false;
  CSS_CHECK(eval_filter(h, filter, &m, nullptr, false, st, &ignore_alive, front));
  if (!m && h->any_dead && !ignore_alive) m = h->alive;
  unsigned char* pin = reinterpret_cast<unsigned char*>(h->pinned);
  memcpy(pin, q_host, qbytes);
  size_t d_off = (qbytes + 15) / 16 * 16;
  size_t i_off = (d_off + dbytes + 15) / 16 * 16;
  size_t c_off = (i_off + ibytes + 15) / 16 * 16;
  CSS_CUDA(cudaMemcpyAsync(sc->q_dev, pin, qbytes, cudaMemcpyHostToDevice, st));
  if (nq == 1 && h->pinned_dev != nullptr && options().scan_mapped.load() != 0) {
    // Single query: the scan kernel writes D / I straight into the mapped staging block and raises a flag there;
    // the host polls the flag instead of paying two D2H copies and a stream synchronisation (~8 us of a 150 us query).
    unsigned char* pdev = reinterpret_cast<unsigned char*>(h->pinned_dev);
    volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(pin + c_off);
    const unsigned seq = (++h->call_seq) & 0x3fffffffu;
    *flag = 0u;
    sc->done_flag = reinterpret_cast<unsigned*>(pdev + c_off);
    sc->done_seq = seq;
    bool tp = false;
    const int rc = search_on_device(h, sc, sc->q_dev, 1, k, m, index_idmap(h, 0), nullptr, reinterpret_cast<float*>(pdev + d_off),
                                    reinterpret_cast<int64_t*>(pdev + i_off), st, /*defer_fallback=*/true, &tp);
    sc->done_flag = nullptr;
    CSS_CHECK(rc);
    unsigned seen = 0;
    CSS_CHECK(await_done_flag(flag, seq, st, &seen, false));
    if (seen & 1u) {
      // not proven from the shadow lists: the fp32 scan answers (into the same mapped block)
      CSS_CHECK(scan_fallback(h, sc, sc->q_dev, 1, k, m, index_idmap(h, 0), nullptr, reinterpret_cast<float*>(pdev + d_off),
                              reinterpret_cast<int64_t*>(pdev + i_off), st, /*pdl_ok=*/false));
      CSS_CUDA(cudaStreamSynchronize(st));
    }
    memcpy(D_host, pin + d_off, dbytes);
    memcpy(I_host, pin + i_off, ibytes);
    return CSS_OK;
  }
  // results of this call: scores at D_dev, ids right behind them (same spacing as in the pinned block)
  int64_t* I_dev = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(sc->D_dev) + (i_off - d_off));
  bool two_phase = false;
  CSS_CHECK(search_on_device(h, sc, sc->q_dev, nq, k, m, index_idmap(h, 0), nullptr, sc->D_dev, I_dev, st,
                             /*defer_fallback=*/true, &two_phase));
  CSS_CUDA(cudaMemcpyAsync(pin + d_off, sc->D_dev, (i_off - d_off) + ibytes, cudaMemcpyDeviceToHost, st));
  if (two_phase) CSS_CUDA(cudaMemcpyAsync(pin + c_off, sc->ovf_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  CSS_CUDA(cudaStreamSynchronize(st));
  if (two_phase && *reinterpret_cast<const int*>(pin + c_off) > 0) {
    // some queries could not be proven from the bf16 lists: the fp32 scan answers exactly those
    CSS_CHECK(scan_fallback(h, sc, sc->q_dev, nq, k, m, index_idmap(h, 0), nullptr, sc->D_dev, I_dev, st, /*pdl_ok=*/false));
    CSS_CUDA(cudaMemcpyAsync(pin + d_off, sc->D_dev, (i_off - d_off) + ibytes, cudaMemcpyDeviceToHost, st));
    CSS_CUDA(cudaStreamSynchronize(st));
  }
  memcpy(D_host, pin + d_off, dbytes);
  memcpy(I_host, pin + i_off, ibytes);
  return CSS_OK;
}

}  // extern "C"

namespace css {
// Poll the completion flag of a single-query launch (see css_index_search).  *seen_out: the flag value.  With
// `final_only` a flag that says "unproven" is not the end: the fp32 re-run that is already in the stream raises
// the flag again.
int await_done_flag(volatile unsigned* flag, unsigned seq, cudaStream_t st, unsigned* seen_out, bool final_only) {
  unsigned seen = 0;
  for (unsigned spins = 1;; ++spins) {
    seen = *flag;
    if ((seen >> 1) == seq && !(final_only && (seen & 1u))) break;
    if ((spins & 1023u) == 0u) {
      // a failed kernel never raises the flag; a finished stream means the result (and the flag) are in place
      const cudaError_t qe = cudaStreamQuery(st);
      if (qe == cudaSuccess) {
        seen = *flag;
        break;
      }
      if (qe != cudaErrorNotReady) {
        set_error("search kernel failed: %s", cudaGetErrorString(qe));
        return CSS_ERR_CUDA;
      }
    }
  }
  if ((seen >> 1) != seq) {
    set_error("search kernel finished without signalling its result");
    return CSS_ERR_CUDA;
  }
  if (seen_out) *seen_out = seen;
  return CSS_OK;
}
}  // namespace css

extern "C" {

int css_index_search_exchange(css_index* h, css_exchange* ex, const float* q_host, int k, const uint32_t* mask_dev,
                              int64_t id_offset, float* D_host, int64_t* I_host) {
  CSS_REQUIRE(h != nullptr && ex != nullptr, "NULL handle");
  CSS_REQUIRE(!CSS_IS_COMPOSITE(h), "css_index_search_exchange: single-device indexes only");
  CSS_REQUIRE(q_host && D_host && I_host, "NULL host buffer");
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  CSS_REQUIRE(ex->connected && ex->device == h->device, "exchange not connected / on another device");
  std::lock_guard<std::mutex> lk(h->mu);
  CSS_CHECK(ensure_device(h->device));
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const size_t qbytes = (size_t)h->dim * 4, dbytes = (size_t)k * 4, ibytes = (size_t)k * 8;
  css_scan_scratch* sc = nullptr;
  CSS_CHECK(get_scratch(h, st, 1, &sc));
  const size_t front = (qbytes + dbytes + ibytes + 128 + 63) / 64 * 64;
  CSS_CHECK(ensure_pinned(h, front));
  CSS_REQUIRE(h->pinned_dev != nullptr, "no mapped view of the pinned staging block");
  unsigned char* pin = reinterpret_cast<unsigned char*>(h->pinned);
  unsigned char* pdev = reinterpret_cast<unsigned char*>(h->pinned_dev);
  const size_t d_off = (qbytes + 15) / 16 * 16, i_off = (d_off + dbytes + 15) / 16 * 16, c_off = (i_off + ibytes + 15) / 16 * 16;
  memcpy(pin, q_host, qbytes);
  CSS_CUDA(cudaMemcpyAsync(sc->q_dev, pin, qbytes, cudaMemcpyHostToDevice, st));
  ExchangeDev xd;
  CSS_CHECK(exchange_next(ex, &xd));
  xd.deferred = 0;
  xd.nq = 1;
  const uint32_t* m = mask_dev;
  if (!m && h->any_dead) m = h->alive;
  volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(pin + c_off);
  const unsigned seq = (++h->call_seq) & 0x3fffffffu;
  *flag = 0u;
  sc->done_flag = reinterpret_cast<unsigned*>(pdev + c_off);
  sc->done_seq = seq;
  float* Dm = reinterpret_cast<float*>(pdev + d_off);
  int64_t* Im = reinterpret_cast<int64_t*>(pdev + i_off);
  const int rc = scan_search(h, sc, sc->q_dev, 1, k, m, index_idmap(h, id_offset), &xd, Dm, Im, st, /*defer_fallback=*/true, nullptr);
  sc->done_flag = nullptr;
  CSS_CHECK(rc);
  unsigned seen = 0;
  CSS_CHECK(await_done_flag(flag, seq, st, &seen, false));
  if (seen & 1u) {
    // not proven from the shadow lists: the fp32 scan answers, publishes this rank's list and merges
    CSS_CHECK(scan_fallback(h, sc, sc->q_dev, 1, k, m, index_idmap(h, id_offset), &xd, Dm, Im, st, /*pdl_ok=*/false));
    CSS_CUDA(cudaStreamSynchronize(st));
  }
  memcpy(D_host, pin + d_off, dbytes);
  memcpy(I_host, pin + i_off, ibytes);
  return CSS_OK;
}

int css_debug_scan_trace(css_index* h, const float* q_dev, int k, int64_t* out_host, int n_out, void* stream) {
  CSS_REQUIRE(h != nullptr && q_dev != nullptr && out_host != nullptr, "NULL argument");
  CSS_REQUIRE(!CSS_IS_COMPOSITE(h) && h->ntotal > 0, "single-device, non-empty indexes only");
  CSS_REQUIRE(n_out >= (h->scan_blocks + 1) * 8, "out_host too small: (blocks + 1) * 8 entries");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  css_scan_scratch* sc = nullptr;
  CSS_CHECK(get_scratch(h, st, 1, &sc));
  const size_t n = (size_t)(h->scan_blocks + 1) * 8;
  if (!sc->trace) CSS_CHECK(dev_alloc(&sc->trace, n));
  CSS_CUDA(cudaMemsetAsync(sc->trace, 0, n * sizeof(long long), st));
  long long* tr = sc->trace;
  float* D = sc->D_dev;
  int64_t* I = reinterpret_cast<int64_t*>(sc->D_dev + CSS_MAX_K);
  const int rc = scan_search(h, sc, q_dev, 1, k, h->any_dead ? h->alive : nullptr, index_idmap(h, 0), nullptr, D, I, st,
                             /*defer_fallback=*/true, nullptr);
  // the stamps are only wanted for this one scan: the scratch keeps the buffer, later scans do not write it
  sc->trace = nullptr;
  cudaError_t e = cudaMemcpyAsync(out_host, tr, n * sizeof(long long), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  sc->trace = nullptr;
  cudaFree(tr);
  CSS_CHECK(rc);
  CSS_CUDA(e);
  return CSS_OK;
}

int css_index_scan_stats(css_index* h, int64_t out[6]) {
  CSS_REQUIRE(h != nullptr && out != nullptr, "NULL argument");
  std::lock_guard<std::mutex> lk(h->mu);
  out[0] = out[1] = out[2] = out[3] = out[4] = 0;
  out[5] = -1;
  auto one = [&](css_index* s) -> int {
    DeviceGuard g(s->device);
    unsigned v[2] = {0, 0};
    CSS_CUDA(cudaMemcpy(v, s->stats_dev, sizeof(v), cudaMemcpyDeviceToHost));
    out[0] += v[0];
    out[1] += v[1];
    out[2] += (s->tier_ban[1] > 0 || s->tier_ban[2] > 0) ? 1 : 0;
    float e = 0.f;
    CSS_CUDA(cudaMemcpy(&e, s->max_err_dev, sizeof(e), cudaMemcpyDeviceToHost));
    out[3] = std::max<int64_t>(out[3], (int64_t)(e * 1e9f));
    CSS_CUDA(cudaMemcpy(&e, s->max_err8_dev, sizeof(e), cudaMemcpyDeviceToHost));
    out[4] = std::max<int64_t>(out[4], e < 1e9f ? (int64_t)(e * 1e9f) : (int64_t)INT64_MAX);
    out[5] = std::max<int64_t>(out[5], s->last_tier);
    return CSS_OK;
  };
  if (CSS_IS_COMPOSITE(h)) {
    for (css_index* s : h->shards) CSS_CHECK(one(s));
    return CSS_OK;
  }
  return one(h);
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

}  // extern "C"
