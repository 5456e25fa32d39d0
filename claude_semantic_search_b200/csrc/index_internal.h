// index_internal.h -- layout of the opaque css_index handle (library-internal).
//
// Data layout in HBM (per index == per GPU shard):
//   x      float32 [capacity, dim] row-major  -- the source of truth, what faiss
//                                               IndexFlat stores (3072 B/row at d=768)
//   xb     bf16    [capacity, dim] row-major  -- shadow copy feeding the tcgen05
//                                               batched score GEMM (K-major operand)
//   cols   int32   [CSS_MAX_COLUMNS][capacity] SoA metadata columns (lazily allocated)
//   alive  uint32  [capacity/32]              -- 1 bit per row
//   mask   uint32  [capacity/32]              -- last evaluated filter
#pragma once
#include "css_common.cuh"
#include "index_kernels.cuh"

struct css_index {
  int dim = 0;
  int metric = 0;
  int device = 0;
  int n_sm = 148;
  int64_t ntotal = 0;
  int64_t capacity = 0;
  float* x = nullptr;
  __nv_bfloat16* xb = nullptr;
  int32_t* cols[CSS_MAX_COLUMNS] = {};
  uint32_t* alive = nullptr;
  uint32_t* mask = nullptr;
  bool any_dead = false;
  float* max_norm_dev = nullptr;     // largest row norm stored (upper bound), 1 float

  // scratch (device)
  int scan_blocks = 148;
  int max_nq = 0;                    // scratch sized for this many queries
  float* q_dev = nullptr;            // [max_nq, dim]
  css::KeyId* part = nullptr;        // [max_nq][scan_blocks][CSS_MAX_K]
  unsigned int* ticket = nullptr;    // [max_nq]
  float* D_dev = nullptr;            // [max_nq, CSS_MAX_K] scores + ids of one host call (12 B per slot)
  int* ovf_list = nullptr;           // two-phase scan: [max_nq] queries handed to the fp32 scan
  int* ovf_count = nullptr;          // [1]
  uint32_t* set_scratch = nullptr;   // clause bitsets
  size_t set_scratch_words = 0;
  uint32_t* rowmask_scratch = nullptr;  // uploaded explicit row mask
  int64_t rowmask_words = 0;
  unsigned long long* n_pass_dev = nullptr;
  // scratch (pinned host)
  void* pinned = nullptr;
  size_t pinned_bytes = 0;

  // batched-search state (search_batched.cu)
  void* batched = nullptr;

  cudaStream_t stream = nullptr;
  std::mutex mu;
};

namespace css {
// Batch-1 / small-nq streaming scan (index.cu).
int scan_search(css_index* h, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                int64_t id_offset, float* D_dev, int64_t* I_dev, cudaStream_t st);
// tcgen05 batched search (search_batched.cu).
int batched_search(css_index* h, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                   int64_t id_offset, float* D_dev, int64_t* I_dev, cudaStream_t st);
void batched_release(css_index* h);
int ensure_query_scratch(css_index* h, int nq);
int ensure_pinned(css_index* h, size_t bytes);
}  // namespace css

// nq at and above which css_index_search uses the tensor-core path.
#define CSS_BATCH_MIN_NQ 16
