// index_sharded.cu -- one flat index over several GPUs of one box, inside ONE process, behind the same
// css_index_* entry points (north_star: the drop-in API keeps its shape while the corpus is row-sharded
// over the 8 GPUs; SURVEY.md 8b proposed `n_dev` in the create call, 8e the layout).
//
// Layout: global row block b (4096 rows) lives on shard b % n_dev as that shard's local block
// b / n_dev.  Ids stay dense and append-only (faiss IndexFlat semantics) for any number of devices,
// every shard's rows are a dense prefix of its arrays, and the shards stay balanced to within one
// block without knowing the final size.  Each shard is an ordinary single-device css_index whose
// kernels translate local rows to global ids (IdMap).
//
// Search (batch-1 / small nq): the query is copied to every device, every device runs its scan, and
// the local top-k lists meet through the in-kernel result exchange over NVLink peer memory
// (exchange.cu / emit_topk): no NCCL, no merge launch; the host reads the merged list from device 0.
// Query batches (tensor-core path) gather the per-shard lists on the host and merge there.
#include "index_internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <thread>
#include <vector>

using namespace css;

namespace {

// One host thread per shard (shards 1..S-1; the caller's thread serves shard 0), each bound to its shard's device.
// A search launches ~3 operations per device; issued from one thread the last of 8 devices would start ~80 us after
// the first -- a third of the 240 us scan.  Workers spin briefly after a job (a burst of queries finds them hot) and
// then block on a condition variable.
struct ShardWorkers {
  std::vector<std::thread> threads;
  std::function<int(int)> job;          // shard index -> status
  std::atomic<uint64_t> seq{0};
  std::atomic<int> pending{0};
  std::atomic<int> failed{0};
  std::string error;                    // first failure's message (under mu)
  std::mutex mu;
  std::condition_variable cv;
  bool stop = false;

  void start(css_index* h) {
    const int S = (int)h->shards.size();
    for (int s = 1; s < S; ++s) threads.emplace_back([this, h, s] { run(h, s); });
  }
  void run(css_index* h, int s) {
    cudaSetDevice(h->shards[s]->device);
    uint64_t seen = 0;
    for (;;) {
      // spin ~200 us for the next job, then sleep
      const auto t0 = std::chrono::steady_clock::now();
      while (seq.load(std::memory_order_acquire) == seen) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return stop || seq.load(std::memory_order_acquire) != seen; });
          break;
        }
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (stop) return;
      }
      seen = seq.load(std::memory_order_acquire);
      const int rc = job(s);
      if (rc != CSS_OK) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed.exchange(rc)) error = css_last_error();
      }
      pending.fetch_sub(1, std::memory_order_release);
    }
  }
  // Run fn(s) for every shard: shards 1.. on the workers, shard 0 here.  Returns the first failure.
  int run_all(int S, std::function<int(int)> fn) {
    failed.store(0);
    if (threads.empty()) {
      for (int s = 0; s < S; ++s) CSS_CHECK(fn(s));
      return CSS_OK;
    }
    job = std::move(fn);
    pending.store(S - 1, std::memory_order_release);
    {
      std::lock_guard<std::mutex> lk(mu);   // pairs with the predicate check of a worker about to sleep
      seq.fetch_add(1, std::memory_order_release);
    }
    cv.notify_all();
    int rc0 = job(0);
    while (pending.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    if (rc0 != CSS_OK) return rc0;
    const int rc = failed.load();
    if (rc != CSS_OK) set_error("%s", error.c_str());
    return rc;
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
      seq.fetch_add(1, std::memory_order_release);
    }
    cv.notify_all();
    for (auto& t : threads) t.join();
    threads.clear();
  }
};

ShardWorkers* workers_of(css_index* h) { return reinterpret_cast<ShardWorkers*>(h->workers); }

inline int64_t words_for(int64_t rows) { return (rows + 31) / 32; }

// Rows of shard s when the index holds N global rows.
int64_t local_count(int64_t S, int64_t s, int64_t N) {
  const int64_t blocks = N >> kShardBlockShift, rem = N & (kShardBlock - 1);
  int64_t c = (blocks / S + (s < blocks % S ? 1 : 0)) << kShardBlockShift;
  if (s == blocks % S) c += rem;
  return c;
}

int sync_all(css_index* h) {
  for (css_index* s : h->shards) {
    DeviceGuard g(s->device);
    CSS_CUDA(cudaStreamSynchronize(s->stream));
  }
  return CSS_OK;
}

// Host merge of per-shard lists: [S][nq][k] -> [nq][k]; best first, ties by ascending id, -1 = hole.
void merge_host(const float* D_all, const int64_t* I_all, int S, int nq, int k, int metric, float* D, int64_t* I) {
  std::vector<std::pair<float, int64_t>> v;
  for (int q = 0; q < nq; ++q) {
    v.clear();
    for (int s = 0; s < S; ++s)
      for (int j = 0; j < k; ++j) {
        const size_t at = ((size_t)s * nq + q) * k + j;
        if (I_all[at] >= 0) v.push_back({metric == CSS_METRIC_INNER_PRODUCT ? D_all[at] : -D_all[at], I_all[at]});
      }
    const size_t take = std::min<size_t>(k, v.size());
    std::partial_sort(v.begin(), v.begin() + take, v.end(), [](const auto& a, const auto& b) {
      return a.first > b.first || (a.first == b.first && a.second < b.second);
    });
    for (int j = 0; j < k; ++j) {
      if ((size_t)j < take) {
        D[(size_t)q * k + j] = metric == CSS_METRIC_INNER_PRODUCT ? v[j].first : -v[j].first;
        I[(size_t)q * k + j] = v[j].second;
      } else {
        D[(size_t)q * k + j] = metric == CSS_METRIC_INNER_PRODUCT ? -FLT_MAX : FLT_MAX;
        I[(size_t)q * k + j] = -1;
      }
    }
  }
}

}  // namespace

namespace css {

void spans_of(const css_index* h, int64_t g0, int64_t n, std::vector<RowSpan>* out) {
  out->clear();
  if (h->shards.empty()) {
    if (n > 0) out->push_back({-1, g0, n, g0});
    return;
  }
  const int64_t S = (int64_t)h->shards.size();
  while (n > 0) {
    const int64_t blk = g0 >> kShardBlockShift, within = g0 & (kShardBlock - 1);
    const int64_t take = std::min(n, kShardBlock - within);
    out->push_back({(int)(blk % S), ((blk / S) << kShardBlockShift) + within, take, g0});
    g0 += take;
    n -= take;
  }
}

int sharded_destroy(css_index* h) {
  if (h->workers) {
    workers_of(h)->shutdown();
    delete workers_of(h);
    h->workers = nullptr;
  }
  for (css_exchange* ex : h->shard_ex) css_exchange_destroy(ex);
  for (css_index* s : h->shards) index_destroy_single(s);
  if (h->pinned) cudaFreeHost(h->pinned);
  delete h;
  return CSS_OK;
}

int sharded_reserve(css_index* h, int64_t capacity) {
  const int64_t S = (int64_t)h->shards.size();
  for (int64_t s = 0; s < S; ++s) {
    css_index* sh = h->shards[s];
    DeviceGuard g(sh->device);
    CSS_CHECK(single_grow(sh, local_count(S, s, capacity)));
  }
  return CSS_OK;
}

int sharded_reset(css_index* h) {
  for (css_index* s : h->shards) {
    DeviceGuard g(s->device);
    CSS_CHECK(single_reset(s));
  }
  h->composite_ntotal = 0;
  return CSS_OK;
}

int sharded_add(css_index* h, const float* x_host, int64_t n, int normalize, int64_t* first_id_out) {
  if (first_id_out) *first_id_out = h->composite_ntotal;
  if (n == 0) return CSS_OK;
  const int64_t S = (int64_t)h->shards.size();
  const int64_t N1 = h->composite_ntotal + n;
  for (int64_t s = 0; s < S; ++s) {
    css_index* sh = h->shards[s];
    DeviceGuard g(sh->device);
    CSS_CHECK(single_ensure_room(sh, local_count(S, s, N1) - sh->ntotal));
  }
  std::vector<RowSpan> spans;
  spans_of(h, h->composite_ntotal, n, &spans);
  for (const RowSpan& sp : spans) {
    css_index* sh = h->shards[sp.shard];
    DeviceGuard g(sh->device);
    if (sp.local != sh->ntotal) {
      set_error("shard %d out of step (local row %lld, shard holds %lld)", sp.shard, (long long)sp.local,
                (long long)sh->ntotal);
      return CSS_ERR_INVALID;
    }
    CSS_CHECK(single_add(sh, x_host + (size_t)(sp.global - h->composite_ntotal) * h->dim, sp.n, normalize, /*sync=*/false));
  }
  CSS_CHECK(sync_all(h));
  h->composite_ntotal = N1;
  return CSS_OK;
}

int sharded_get_rows(css_index* h, int64_t start, int64_t n, float* out_host) {
  std::vector<RowSpan> spans;
  spans_of(h, start, n, &spans);
  for (const RowSpan& sp : spans)
    CSS_CHECK(css_index_get_rows(h->shards[sp.shard], sp.local, sp.n, out_host + (size_t)(sp.global - start) * h->dim));
  return CSS_OK;
}

int sharded_set_column(css_index* h, int column, const int32_t* values_host, int64_t start, int64_t n) {
  std::vector<RowSpan> spans;
  spans_of(h, start, n, &spans);
  for (const RowSpan& sp : spans) {
    css_index* sh = h->shards[sp.shard];
    std::lock_guard<std::mutex> lk(sh->mu);
    DeviceGuard g(sh->device);
    // pageable source: the copy has consumed the host bytes when the call returns
    CSS_CHECK(single_set_column(sh, column, values_host + (sp.global - start), sp.local, sp.n, /*sync=*/false));
  }
  return sync_all(h);
}

int sharded_set_alive(css_index* h, const uint8_t* alive_host, int64_t start, int64_t n) {
  std::vector<RowSpan> spans;
  spans_of(h, start, n, &spans);
  for (const RowSpan& sp : spans)
    CSS_CHECK(css_index_set_alive(h->shards[sp.shard], alive_host + (sp.global - start), sp.local, sp.n));
  return CSS_OK;
}

int sharded_set_alive_ids(css_index* h, const int64_t* ids_host, int64_t n, int alive) {
  const int64_t S = (int64_t)h->shards.size();
  std::vector<std::vector<int64_t>> per(S);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t g = ids_host[i];
    if (g < 0 || g >= h->composite_ntotal) continue;
    const int64_t blk = g >> kShardBlockShift;
    per[blk % S].push_back(((blk / S) << kShardBlockShift) + (g & (kShardBlock - 1)));
  }
  for (int64_t s = 0; s < S; ++s)
    if (!per[s].empty()) CSS_CHECK(css_index_set_alive_ids(h->shards[s], per[s].data(), (int64_t)per[s].size(), alive));
  return CSS_OK;
}

// Bits of the global mask <-> the shards' local masks: whole 128-word blocks move.
static void scatter_mask_to_global(const uint32_t* local, int64_t local_rows, int64_t S, int64_t s, uint32_t* global,
                                   int64_t global_rows) {
  const int64_t wpb = kShardBlock / 32;
  const int64_t lw = words_for(local_rows), gw = words_for(global_rows);
  for (int64_t lb = 0; lb * wpb < lw; ++lb) {
    const int64_t gb = lb * S + s;
    const int64_t nw = std::min(wpb, std::min(lw - lb * wpb, gw - gb * wpb));
    if (nw > 0) memcpy(global + gb * wpb, local + lb * wpb, (size_t)nw * 4);
  }
}
static void gather_mask_from_global(const uint32_t* global, int64_t global_rows, int64_t S, int64_t s, uint32_t* local,
                                    int64_t local_rows) {
  const int64_t wpb = kShardBlock / 32;
  const int64_t lw = words_for(local_rows), gw = words_for(global_rows);
  for (int64_t lb = 0; lb * wpb < lw; ++lb) {
    const int64_t gb = lb * S + s;
    const int64_t nw = std::min(wpb, std::min(lw - lb * wpb, gw - gb * wpb));
    if (nw > 0) memcpy(local + lb * wpb, global + gb * wpb, (size_t)nw * 4);
  }
}

int sharded_filter_mask(css_index* h, const css_filter* f, uint32_t* mask_out_host, int64_t* n_pass_out) {
  const int64_t S = (int64_t)h->shards.size(), N = h->composite_ntotal;
  int64_t total = 0;
  if (N > 0) memset(mask_out_host, 0, (size_t)words_for(N) * 4);
  std::vector<uint32_t> local, rm;
  for (int64_t s = 0; s < S; ++s) {
    css_index* sh = h->shards[s];
    if (sh->ntotal == 0) continue;
    css_filter fs;
    memset(&fs, 0, sizeof(fs));
    if (f) fs = *f;
    if (f && f->row_mask) {
      rm.assign((size_t)words_for(sh->ntotal), 0u);
      gather_mask_from_global(f->row_mask, N, S, s, rm.data(), sh->ntotal);
      fs.row_mask = rm.data();
    }
    local.assign((size_t)words_for(sh->ntotal), 0u);
    int64_t np = 0;
    CSS_CHECK(css_index_filter_mask(sh, &fs, local.data(), &np));
    total += np;
    scatter_mask_to_global(local.data(), sh->ntotal, S, s, mask_out_host, N);
  }
  if (n_pass_out) *n_pass_out = total;
  return CSS_OK;
}

int sharded_search(css_index* h, const float* q_host, int nq, int k, const css_filter* filter, float* D_host,
                   int64_t* I_host) {
  const int S = (int)h->shards.size();
  const int64_t N = h->composite_ntotal;
  DeviceGuard g0(h->device);
  if (nq > 1024) {
    for (int q0 = 0; q0 < nq; q0 += 1024) {
      const int n = std::min(1024, nq - q0);
      CSS_CHECK(sharded_search(h, q_host + (size_t)q0 * h->dim, n, k, filter, D_host + (size_t)q0 * k,
                               I_host + (size_t)q0 * k));
    }
    return CSS_OK;
  }
  int64_t min_rows = N;
  for (css_index* s : h->shards) min_rows = std::min(min_rows, s->ntotal);
  const bool batched = nq >= CSS_BATCH_MIN_NQ && h->metric == CSS_METRIC_INNER_PRODUCT && h->dim % 64 == 0 &&
                       min_rows >= 65536;
  const bool exchange = !batched && nq <= h->shard_ex[0]->max_nq;
  const size_t qbytes = (size_t)nq * h->dim * 4;
  const size_t dbytes = (size_t)nq * k * 4, ibytes = (size_t)nq * k * 8;
  const size_t d_off = (qbytes + 15) / 16 * 16;
  const size_t res_stride = ((dbytes + 15) / 16 * 16) + ((ibytes + 15) / 16 * 16);
  const size_t i_rel = (dbytes + 15) / 16 * 16;
  CSS_CHECK(ensure_pinned(h, d_off + res_stride * S + 64));
  unsigned char* pin = reinterpret_cast<unsigned char*>(h->pinned);
  memcpy(pin, q_host, qbytes);
  // single query on the exchange path: device 0 writes the merged result straight into the mapped staging block and
  // raises a flag there (see css_index_search); no D2H copy, no stream synchronisation
  const size_t f_off = d_off + res_stride * S;
  const bool mapped = exchange && nq == 1 && h->pinned_dev != nullptr && options().scan_mapped.load() != 0;
  volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(pin + f_off);
  const unsigned seq = (++h->call_seq) & 0x3fffffffu;
  if (mapped) *flag = 0u;
  std::vector<std::vector<uint32_t>> rms(filter && filter->row_mask ? S : 0);
  // Pass 1, per shard: scratch, filter evaluation.  Anything that may free device memory (scratch or filter buffers
  // growing) happens here, sequentially: cudaFree waits for ALL work of its device, and once the scans of pass 2
  // are in flight that includes kernels waiting for the lists of shards that have not been launched yet.
  std::vector<css_scan_scratch*> scs(S, nullptr);
  std::vector<const uint32_t*> masks(S, nullptr);
  for (int s = 0; s < S; ++s) {
    css_index* sh = h->shards[s];
    std::lock_guard<std::mutex> lk(sh->mu);
    const auto it = sh->scratch.find(sh->stream);
    const bool ready = it != sh->scratch.end() && it->second.max_nq >= nq;
    if (ready && !filter && !sh->any_dead) {   // nothing to evaluate, nothing to grow: no CUDA call at all
      scs[s] = &it->second;
      continue;
    }
    DeviceGuard g(sh->device);
    cudaStream_t st = sh->stream;
    CSS_CHECK(get_scratch(sh, st, nq, &scs[s]));
    css_filter fs;
    const css_filter* fp = filter;
    if (filter && filter->row_mask) {
      fs = *filter;
      rms[s].assign((size_t)std::max<int64_t>(1, words_for(sh->ntotal)), 0u);
      gather_mask_from_global(filter->row_mask, N, S, s, rms[s].data(), sh->ntotal);
      fs.row_mask = rms[s].data();
      fp = &fs;
    }
    bool ignore_alive = false;
    CSS_CHECK(eval_filter(sh, fp, &masks[s], nullptr, false, st, &ignore_alive, 0));
    if (!masks[s] && sh->any_dead && !ignore_alive) masks[s] = sh->alive;
  }
  // Pass 2, one host thread per shard: query upload, the scans (and, on the exchange path, the in-kernel merge)
  std::vector<ExchangeDev> xds(S);
  if (exchange)
    for (int s = 0; s < S; ++s) CSS_CHECK(exchange_next(h->shard_ex[s], &xds[s]));
  auto launch = [&](int s) -> int {
    css_index* sh = h->shards[s];
    std::lock_guard<std::mutex> lk(sh->mu);
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != sh->device) CSS_CUDA(cudaSetDevice(sh->device));   // workers are bound already; the caller's thread is restored below
    cudaStream_t st = sh->stream;
    css_scan_scratch* sc = scs[s];
    const uint32_t* m = masks[s];
    CSS_CUDA(cudaMemcpyAsync(sc->q_dev, pin, qbytes, cudaMemcpyHostToDevice, st));
    int64_t* I_dev = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(sc->D_dev) + i_rel);
    if (exchange && mapped && s == 0) {
      unsigned char* pdev = reinterpret_cast<unsigned char*>(h->pinned_dev);
      sc->done_flag = reinterpret_cast<unsigned*>(pdev + f_off);
      sc->done_seq = seq;
      const int rc = scan_search(sh, sc, sc->q_dev, nq, k, m, index_idmap(sh, 0), &xds[s], reinterpret_cast<float*>(pdev + d_off),
                                 reinterpret_cast<int64_t*>(pdev + d_off + i_rel), st, /*defer_fallback=*/false, nullptr);
      sc->done_flag = nullptr;
      CSS_CHECK(rc);
    } else if (exchange) {
      CSS_CHECK(scan_search(sh, sc, sc->q_dev, nq, k, m, index_idmap(sh, 0), &xds[s], sc->D_dev, I_dev, st,
                            /*defer_fallback=*/false, nullptr));
      if (s == 0) CSS_CUDA(cudaMemcpyAsync(pin + d_off, sc->D_dev, i_rel + ibytes, cudaMemcpyDeviceToHost, st));
    } else {
      CSS_CHECK(search_on_device(sh, sc, sc->q_dev, nq, k, m, index_idmap(sh, 0), nullptr, sc->D_dev, I_dev, st,
                                 /*defer_fallback=*/false, nullptr));
      CSS_CUDA(cudaMemcpyAsync(pin + d_off + res_stride * s, sc->D_dev, i_rel + ibytes, cudaMemcpyDeviceToHost, st));
    }
    return CSS_OK;
  };
  // the tensor-core path may grow its own state (allocations): keep it on the caller's thread, one shard after the other
  if (!exchange || !h->workers) {
    for (int s = 0; s < S; ++s) {
      DeviceGuard g(h->shards[s]->device);
      CSS_CHECK(launch(s));
    }
  } else {
    CSS_CHECK(workers_of(h)->run_all(S, launch));
    cudaSetDevice(h->device);
  }
  if (exchange) {
    // every device merged the same lists; device 0's copy is the answer.  Its completion implies that
    // every shard has published, i.e. consumed the pinned query.
    css_index* s0 = h->shards[0];
    DeviceGuard g(s0->device);
    if (mapped) CSS_CHECK(await_done_flag(flag, seq, s0->stream, nullptr, /*final_only=*/true));
    else CSS_CUDA(cudaStreamSynchronize(s0->stream));
    memcpy(D_host, pin + d_off, dbytes);
    memcpy(I_host, pin + d_off + i_rel, ibytes);
    return CSS_OK;
  }
  CSS_CHECK(sync_all(h));
  std::vector<float> D_all((size_t)S * nq * k);
  std::vector<int64_t> I_all((size_t)S * nq * k);
  for (int s = 0; s < S; ++s) {
    memcpy(D_all.data() + (size_t)s * nq * k, pin + d_off + res_stride * s, dbytes);
    memcpy(I_all.data() + (size_t)s * nq * k, pin + d_off + res_stride * s + i_rel, ibytes);
  }
  merge_host(D_all.data(), I_all.data(), S, nq, k, h->metric, D_host, I_host);
  return CSS_OK;
}

// Compaction across devices: kept row j moves from global row keep[j] to global row j, which in general
// lives on another device.  The rows stream through a host buffer in ascending order (a written position
// j is below every source position still to be read, keep[j'] >= j' > j), then every shard is cut to its
// new length.  Alive bits are reset to "alive" and the metadata columns to NULL: the caller re-uploads
// them (HybridStorage._rebuild_columns does, right after).
int sharded_compact(css_index* h, const int64_t* keep, int64_t n_keep) {
  const int64_t S = (int64_t)h->shards.size();
  const int64_t chunk = 16384;
  std::vector<float> buf((size_t)std::min(chunk, std::max<int64_t>(n_keep, 1)) * h->dim);
  std::vector<RowSpan> spans;
  for (int64_t j0 = 0; j0 < n_keep; j0 += chunk) {
    const int64_t n = std::min(chunk, n_keep - j0);
    for (int64_t a = 0; a < n;) {   // runs of consecutive source rows
      int64_t b = a + 1;
      while (b < n && keep[j0 + b] == keep[j0 + b - 1] + 1) ++b;
      CSS_CHECK(sharded_get_rows(h, keep[j0 + a], b - a, buf.data() + (size_t)a * h->dim));
      a = b;
    }
    spans_of(h, j0, n, &spans);
    for (const RowSpan& sp : spans) {
      css_index* sh = h->shards[sp.shard];
      std::lock_guard<std::mutex> lk(sh->mu);
      DeviceGuard g(sh->device);
      CSS_CHECK(put_rows_host_async(sh, buf.data() + (size_t)(sp.global - j0) * h->dim, sp.local, sp.n, 0));
      CSS_CUDA(cudaStreamSynchronize(sh->stream));
    }
  }
  for (int64_t s = 0; s < S; ++s) {
    css_index* sh = h->shards[s];
    std::lock_guard<std::mutex> lk(sh->mu);
    DeviceGuard g(sh->device);
    const int64_t nl = local_count(S, s, n_keep);
    const int64_t old = sh->ntotal;
    if (sh->capacity > 0) CSS_CUDA(cudaMemsetAsync(sh->alive, 0, (size_t)words_for(sh->capacity) * 4, sh->stream));
    if (nl > 0) CSS_CHECK(single_mark_alive(sh, 0, nl));
    for (int c = 0; c < CSS_MAX_COLUMNS; ++c) {
      if (!sh->cols[c] || old == 0) continue;
      fill_i32_kernel<<<(unsigned)((old + 255) / 256), 256, 0, sh->stream>>>(sh->cols[c], old, CSS_NULL_VALUE);
      CSS_LAUNCHED();
    }
    CSS_CUDA(cudaStreamSynchronize(sh->stream));
    sh->ntotal = nl;
    sh->version++;
    sh->any_dead = false;
  }
  h->composite_ntotal = n_keep;
  return CSS_OK;
}

}  // namespace css

extern "C" {

int css_index_create_sharded(int dim, int metric, const int* devices, int n_dev, css_index** out) {
  CSS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  CSS_REQUIRE(dim >= 1 && dim <= 65536, "dim %d out of range", dim);
  CSS_REQUIRE(metric == CSS_METRIC_INNER_PRODUCT || metric == CSS_METRIC_L2, "unknown metric %d", metric);
  CSS_REQUIRE(devices != nullptr && n_dev >= 1 && n_dev <= CSS_MAX_RANKS, "n_dev %d outside [1, %d]", n_dev, CSS_MAX_RANKS);
  // a device may be listed more than once (several shards on one GPU): of no use in production, but it runs
  // the whole multi-shard path -- block-cyclic ids, in-kernel exchange -- on a single-GPU box (tests)
  css_index* h = new (std::nothrow) css_index();
  if (!h) {
    set_error("out of host memory");
    return CSS_ERR_OOM;
  }
  h->dim = dim;
  h->metric = metric;
  h->device = devices[0];
  int rc = CSS_OK;
  for (int s = 0; s < n_dev && rc == CSS_OK; ++s) {
    css_index* sh = nullptr;
    rc = index_create_single(dim, metric, devices[s], &sh);
    if (rc != CSS_OK) break;
    sh->id_shift = kShardBlockShift;
    sh->id_ndev = n_dev;
    sh->id_shard = s;
    h->shards.push_back(sh);
    css_exchange* ex = nullptr;
    rc = css_exchange_create(devices[s], n_dev, s, &ex, nullptr);
    if (rc == CSS_OK) h->shard_ex.push_back(ex);
  }
  if (rc == CSS_OK && n_dev > 1) rc = exchange_connect_local(h->shard_ex.data(), n_dev);
  if (rc != CSS_OK) {
    if (h->shards.empty()) delete h;
    else sharded_destroy(h);
    return rc;
  }
  h->n_sm = h->shards[0]->n_sm;
  // host threads for the per-device launches: only worth it (and only safe against oversubscription in tests that
  // put many shards on one GPU) with distinct devices; CSS_SHARD_THREADS=0 keeps everything on the caller's thread
  bool distinct = n_dev >= 2;
  for (int a = 0; a < n_dev; ++a)
    for (int b = 0; b < a; ++b) distinct = distinct && devices[a] != devices[b];
  const char* tv = getenv("CSS_SHARD_THREADS");
  if (distinct && !(tv && atoi(tv) == 0)) {
    ShardWorkers* w = new (std::nothrow) ShardWorkers();
    if (w) {
      w->start(h);
      h->workers = w;
    }
  }
  *out = h;
  return CSS_OK;
}

}  // extern "C"
