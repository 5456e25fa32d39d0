// index_internal.h -- layout of the opaque css_index handle (library-internal).
//
// Data layout in HBM (per index == per GPU shard):
//   x      float32 [capacity, dim] row-major  -- the source of truth, what faiss
//                                               IndexFlat stores (3072 B/row at d=768)
//   xb     bf16    [capacity, dim] row-major  -- shadow copy: stream of phase 1 of the two-phase
//                                               batch-1 scan (second tier) and K-major B operand of the
//                                               tcgen05 batched score GEMM
//   xq     int8    [capacity, 768] row-major  -- shadow copy, first tier of the two-phase scan (768 B/row),
//   xs     float32 [capacity]                    with one scale per row: x ~= xs[row] * xq[row]
//                                               (768-d inner-product indexes only)
//   cols   int32   [CSS_MAX_COLUMNS][capacity] SoA metadata columns (lazily allocated)
//   alive  uint32  [capacity/32]              -- 1 bit per row
//   mask   uint32  [capacity/32]              -- last evaluated filter
// x / xb / cols / alive / mask are growable arrays backed by CUDA virtual memory (vmm.h): growth maps
// more physical memory behind the same addresses, nothing is copied and nothing is held twice.
//
// A multi-device index (css_index_create_sharded) is a css_index whose `shards` vector holds one
// ordinary single-device css_index per device; global row block b (kShardBlock rows) lives on shard
// b % n_dev as local block b / n_dev, so ids stay dense and append-only.
#pragma once
#include <map>
#include <vector>

#include "css_common.cuh"
#include "index_kernels.cuh"
#include "vmm.h"

// Per-stream scratch of the search paths: two host threads searching one index on two streams
// never share a partial-list buffer, ticket or overflow list.
struct css_scan_scratch {
  int max_nq = 0;
  float* q_dev = nullptr;            // [max_nq, dim]
  css::KeyId* part = nullptr;        // [max_nq][scan_blocks][CSS_MAX_K]
  float* part_exact = nullptr;       // [max_nq][scan_blocks][CSS_MAX_K] two-phase scan: fp32 scores of the list entries
  unsigned int* ticket = nullptr;    // [max_nq]
  float* D_dev = nullptr;            // [max_nq, CSS_MAX_K] scores + ids of one host call (12 B per slot) + overflow count
  int* ovf_list = nullptr;           // two-phase scan: [max_nq] queries handed to the fp32 scan
  int* ovf_count = nullptr;          // [1] (inside D_dev's allocation so one D2H returns it with the result)
  void* batched = nullptr;           // BatchedState of search_batched.cu
  long long* trace = nullptr;        // css_debug_scan_trace: timeline stamps of the next scans on this stream
  unsigned* done_flag = nullptr;     // set around a single-query host call: mapped host word the scan kernel signals
  unsigned done_seq = 0;
};

struct css_exchange {
  int device = 0;
  int n_ranks = 1;
  int rank = 0;
  int max_nq = 0;
  unsigned epoch = 0;                // last epoch used
  css::ExEntry* slots_local = nullptr;
  unsigned* flags_local = nullptr;
  int* status = nullptr;             // device view of status_host
  volatile int* status_host = nullptr;  // mapped pinned: 1 = a peer's list timed out
  css::ExEntry* slots[css::kMaxRanks] = {};
  unsigned* flags[css::kMaxRanks] = {};
  bool opened[css::kMaxRanks] = {};  // peer mappings opened through CUDA IPC (to be closed)
  bool connected = false;
  bool shared_device = false;        // another shard of this process lives on the same GPU
  std::mutex mu;
};

struct css_index {
  int dim = 0;
  int metric = 0;
  int device = 0;
  int n_sm = 148;
  int64_t ntotal = 0;
  int64_t capacity = 0;
  float* x = nullptr;
  __nv_bfloat16* xb = nullptr;
  int8_t* xq = nullptr;              // int8 shadow rows + per-row scales (768-d inner-product indexes only)
  float* xs = nullptr;
  int32_t* cols[CSS_MAX_COLUMNS] = {};
  uint32_t* alive = nullptr;
  uint32_t* mask = nullptr;
  css::VmmArray vx, vxb, vxq, vxs, valive, vmask, vcols[CSS_MAX_COLUMNS];
  bool any_dead = false;
  float* max_norm_dev = nullptr;     // largest row norm stored (upper bound), 1 float
  float* max_err_dev = nullptr;      // largest ||x - bf16(x)|| stored (upper bound), 1 float
  float* max_err8_dev = nullptr;     // largest ||x - xs * xq|| stored (upper bound; +inf after a non-finite row), 1 float

  // scratch (device)
  int scan_blocks = 148;
  std::map<cudaStream_t, css_scan_scratch> scratch;
  uint32_t* set_scratch = nullptr;   // clause bitsets
  size_t set_scratch_words = 0;
  uint32_t* rowmask_scratch = nullptr;  // uploaded explicit row mask
  int64_t rowmask_words = 0;
  unsigned long long* n_pass_dev = nullptr;
  // the filter evaluated into `mask` last: an unchanged filter over an unchanged index is not evaluated again
  uint64_t version = 0;              // bumped by everything that changes rows, columns or alive bits
  uint64_t mask_version = ~0ull;
  std::vector<uint32_t> mask_key;
  int64_t mask_n_pass = -1;
  int64_t* ids_scratch = nullptr;    // css_index_set_alive_ids
  int64_t ids_scratch_n = 0;
  // scratch (pinned host)
  void* pinned = nullptr;            // mapped (the scan kernels write single-query results straight into it)
  void* pinned_dev = nullptr;        // device view of `pinned`
  size_t pinned_bytes = 0;
  unsigned call_seq = 0;             // sequence number of the single-query host calls (completion flag value)

  // two-phase scan statistics: device counters {queries, unproven} mirrored by the kernel into mapped
  // host memory, read without synchronisation to steer the adaptive path choice
  unsigned* stats_dev = nullptr;
  volatile unsigned* stats_host = nullptr;   // mapped pinned
  unsigned* stats_host_devptr = nullptr;
  unsigned seen_q = 0, seen_u = 0;   // counters at the last decision
  int64_t tier_ban[3] = {0, 0, 0};   // [t] > 0: the next that many scan queries skip tier t (1 = bf16, 2 = int8 shadow)
  int last_tier = -1;                // tier of the previous scan call (-1: none yet)

  // multi-device composite
  std::vector<css_index*> shards;
  std::vector<css_exchange*> shard_ex;
  int64_t composite_ntotal = 0;
  void* workers = nullptr;           // ShardWorkers (index_sharded.cu): one launch thread per shard

  // id mapping applied to returned rows (composite shards: block-cyclic)
  int id_shift = 0, id_ndev = 1, id_shard = 0;

  cudaStream_t stream = nullptr;
  std::mutex mu;
};

namespace css {
constexpr int kShardBlockShift = 12;             // 4096-row blocks dealt round-robin over the devices
constexpr int64_t kShardBlock = (int64_t)1 << kShardBlockShift;

// Batch-1 / small-nq streaming scan (index.cu).
int scan_search(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev, cudaStream_t st,
                bool defer_fallback, bool* two_phase_used);
// fp32 scan of the queries the two-phase scan queued (ovf_list of the scratch).
int scan_fallback(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                  const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev, cudaStream_t st, bool pdl_ok);
// tcgen05 batched search (search_batched.cu).
int batched_search(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                   const IdMap& idmap, float* D_dev, int64_t* I_dev, cudaStream_t st);
void batched_release(css_scan_scratch* sc);
int get_scratch(css_index* h, cudaStream_t st, int nq, css_scan_scratch** out);
int ensure_pinned(css_index* h, size_t bytes);
int await_done_flag(volatile unsigned* flag, unsigned seq, cudaStream_t st, unsigned* seen_out, bool final_only);
IdMap index_idmap(const css_index* h, int64_t id_offset);
int eval_filter(css_index* h, const css_filter* f, const uint32_t** mask_out, int64_t* n_pass, bool need_count,
                cudaStream_t st, bool* ignore_alive_out, size_t pinned_front);
int search_on_device(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* m,
                     const IdMap& idmap, const ExchangeDev* ex, float* D_dev, int64_t* I_dev, cudaStream_t st,
                     bool defer_fallback, bool* two_phase_used);

// single-device primitives the composite builds on (index.cu); the caller holds the lock and the device
int index_create_single(int dim, int metric, int device, css_index** out);
void index_destroy_single(css_index* h);
int single_add(css_index* h, const float* x_host, int64_t n, int normalize, bool sync);
int single_reset(css_index* h);
int single_set_column(css_index* h, int column, const int32_t* values_host, int64_t start, int64_t n, bool sync);
int single_set_alive_ids(css_index* h, const int64_t* ids_host, int64_t n, int alive);
int single_grow(css_index* h, int64_t cap);
int single_ensure_room(css_index* h, int64_t extra);
int single_mark_alive(css_index* h, int64_t row0, int64_t n);
int put_rows_host_async(css_index* h, const float* x_host, int64_t row0, int64_t n, int normalize);

// result exchange (exchange.cu)
int exchange_next(css_exchange* ex, ExchangeDev* out);   // advance the epoch, fill the device view
int exchange_connect_local(css_exchange** exs, int n);   // peers in this process: direct peer pointers

// A run of rows inside one shard (shard = -1: the index itself is a single shard).
struct RowSpan {
  int shard;
  int64_t local;   // first local row
  int64_t n;
  int64_t global;  // first global row
};
void spans_of(const css_index* h, int64_t g0, int64_t n, std::vector<RowSpan>* out);

// multi-device composite (index_sharded.cu); the caller holds the composite's lock
int sharded_destroy(css_index* h);
int sharded_reserve(css_index* h, int64_t capacity);
int sharded_reset(css_index* h);
int sharded_add(css_index* h, const float* x_host, int64_t n, int normalize, int64_t* first_id_out);
int sharded_get_rows(css_index* h, int64_t start, int64_t n, float* out_host);
int sharded_set_column(css_index* h, int column, const int32_t* values_host, int64_t start, int64_t n);
int sharded_set_alive(css_index* h, const uint8_t* alive_host, int64_t start, int64_t n);
int sharded_set_alive_ids(css_index* h, const int64_t* ids_host, int64_t n, int alive);
int sharded_filter_mask(css_index* h, const css_filter* f, uint32_t* mask_out_host, int64_t* n_pass_out);
int sharded_search(css_index* h, const float* q_host, int nq, int k, const css_filter* filter, float* D_host,
                   int64_t* I_host);
int sharded_compact(css_index* h, const int64_t* keep_ids_host, int64_t n_keep);
}  // namespace css

// nq at and above which css_index_search uses the tensor-core path.
#define CSS_BATCH_MIN_NQ 16
