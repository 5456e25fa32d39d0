/*
 * css_b200.h -- C ABI of the B200-native hot path for claude-semantic-search.
 *
 * This is the drop-in boundary: everything the reference computes through
 * `faiss` (src/storage.py) and `sentence_transformers` (src/embeddings.py) is
 * reached through the entry points below.  Plain pointers and sizes only; no
 * torch / C++ types.  The shared library is libcss_b200.so (built by
 * claude_semantic_search_b200/build.py with nvcc for sm_100a).
 *
 * Conventions
 *   - every function returns 0 on success, a negative css_status otherwise;
 *     css_last_error() returns a thread-local message for the last failure.
 *   - the caller owns every host buffer; the library owns all device memory.
 *   - "_device" variants take device pointers + a CUDA stream (as void*) and do
 *     not synchronise; the plain variants take HOST buffers, copy in/out and
 *     return when the result is in the host buffer.
 *   - handles are internally locked: add/save/reset exclude search.  "_device" searches issued on
 *     different streams use separate scratch buffers and may overlap on the device.
 *   - a shard (one device) holds fewer than 2^31 rows: row ids are 32-bit inside the kernels and
 *     widened (+ id_offset / the multi-device id map) on output.
 *   - there is no CPU fallback: without an sm_100 device every compute entry
 *     point fails with CSS_ERR_NO_DEVICE.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   faiss.IndexFlatIP(d) / IndexFlatL2(d)      src/storage.py:252-258  -> css_index_create
 *   faiss_index.add(x)                         src/storage.py:358-359  -> css_index_add
 *   faiss_index.ntotal                         src/storage.py:308,421  -> css_index_ntotal
 *   faiss_index.search(q, k)                   src/storage.py:436      -> css_index_search
 *   _matches_filters per candidate row         src/storage.py:508-543  -> css_index_filter_mask (+ filter arg of search)
 *   faiss.write_index / read_index             src/storage.py:306,879-884 -> css_index_save / css_index_load
 *   SentenceTransformer(...).encode(...)       src/embeddings.py:184-188,216-222 -> css_encoder_encode
 *   SentenceTransformer(name).to(device)       src/embeddings.py:86-97 -> css_encoder_create
 *   tokenisation inside .encode(...)           src/embeddings.py:216-222 -> css_tokenizer_encode_batch
 */
#ifndef CSS_B200_H
#define CSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSS_ABI_VERSION 1

typedef enum css_status {
  CSS_OK = 0,
  CSS_ERR_INVALID = -1,      /* bad argument */
  CSS_ERR_NO_DEVICE = -2,    /* no sm_100 CUDA device / driver */
  CSS_ERR_CUDA = -3,         /* a CUDA call failed; see css_last_error() */
  CSS_ERR_OOM = -4,          /* device or host allocation failed */
  CSS_ERR_IO = -5,           /* file I/O or file format */
  CSS_ERR_UNSUPPORTED = -6,  /* valid request this build cannot serve (e.g. k too large) */
  CSS_ERR_OVERFLOW = -7      /* internal candidate buffer overflow (never silently truncated) */
} css_status;

typedef enum css_metric {
  CSS_METRIC_INNER_PRODUCT = 0, /* faiss.METRIC_INNER_PRODUCT, IndexFlatIP */
  CSS_METRIC_L2 = 1             /* faiss.METRIC_L2, IndexFlatL2 (squared distances) */
} css_metric;

int css_abi_version(void);
const char* css_last_error(void);
/* Number of usable sm_100 devices (CSS_ERR_NO_DEVICE if none). */
int css_device_count(int* n_out);
/* {SM count, HBM bytes total, HBM bytes free, cc major, cc minor} of `device`. */
int css_device_info(int device, int64_t info_out[5]);

/* ------------------------------------------------------------------ */
/* Flat index (half B of the hot path)                                 */
/* ------------------------------------------------------------------ */
typedef struct css_index css_index;

/* Maximum k served by the fused scan (reference uses k' = min(100, ntotal),
 * src/storage.py:432). */
#define CSS_MAX_K 128
/* Number of int32 metadata columns an index can carry for filtering. */
#define CSS_MAX_COLUMNS 12
#define CSS_MAX_CLAUSES 16

int css_index_create(int dim, int metric, int device, css_index** out);
/* The same index row-sharded over n_dev GPUs of this box inside ONE process (north_star: the corpus
 * sharded over the 8 GPUs behind the unchanged HybridStorage API; the reference hard-codes device 0,
 * src/storage.py:269-284).  Every css_index_* entry point below works on the returned handle except the
 * "_device" variants: rows are dealt to the devices in 4096-row blocks (ids stay dense and append-only),
 * a search runs on all devices at once and the per-device top-k lists are merged inside the scan
 * kernels over NVLink peer memory (no NCCL, no extra launch); css_index_compact resets alive bits and
 * metadata columns (re-upload them afterwards). */
#define CSS_MAX_RANKS 8
int css_index_create_sharded(int dim, int metric, const int* devices, int n_dev, css_index** out);
int css_index_n_devices(const css_index* h);
int css_index_destroy(css_index* h);
int css_index_dim(const css_index* h);
int css_index_metric(const css_index* h);
int64_t css_index_ntotal(const css_index* h);
int64_t css_index_capacity(const css_index* h);
/* Pre-size device storage (rows).  Growth is otherwise geometric. */
int css_index_reserve(css_index* h, int64_t capacity);
/* Drop all rows (faiss Index.reset()). */
int css_index_reset(css_index* h);

/* Append n rows (row-major float32 [n, dim], HOST memory).  If normalize != 0
 * each row is scaled by 1 / (||x||_2 + 1e-8) on the device, the arithmetic of
 * src/storage.py:347-350.  Row i receives id first_id + i; ids are dense,
 * append-only and never reused (faiss IndexFlat semantics). */
int css_index_add(css_index* h, const float* x_host, int64_t n, int normalize,
                  int64_t* first_id_out);
/* Same with x in DEVICE memory of the index's device (e.g. encoder output). */
int css_index_add_device(css_index* h, const float* x_dev, int64_t n, int normalize,
                         int64_t* first_id_out, void* stream);
/* Compaction on the device (HybridStorage.optimize / _rebuild_faiss_index, src/storage.py:930-969,
 * which the reference leaves as a stub): the index keeps exactly the rows keep_ids[0..n_keep)
 * (strictly ascending row ids), renumbered 0..n_keep-1 in that order, together with their
 * metadata columns and alive bits.  HBM-bound gather, no host round trip of the vectors. */
int css_index_compact(css_index* h, const int64_t* keep_ids_host, int64_t n_keep);
/* faiss reconstruct_n: copy rows [start, start+n) back to host. */
int css_index_get_rows(css_index* h, int64_t start, int64_t n, float* out_host);

/* Metadata columns (int32, one value per row; dictionary ids / ranks / raw
 * integers chosen by the host, CSS_NULL_VALUE for SQL NULL). */
#define CSS_NULL_VALUE INT32_MIN
int css_index_set_column(css_index* h, int column, const int32_t* values_host,
                         int64_t start, int64_t n);
/* alive[i] != 0 -> row start+i may be returned.  New rows are alive.  Rows
 * whose chunk was deleted/re-indexed are orphans in the reference
 * (src/storage.py:449-451,836-846) and are cleared here. */
int css_index_set_alive(css_index* h, const uint8_t* alive_host, int64_t start, int64_t n);
/* The same for a list of row ids in ONE device call (remove_chunks_for_file / delete_chunks_by_session,
 * src/storage.py:817-846: hundreds of orphaned rows per file); ids outside [0, ntotal) are ignored. */
int css_index_set_alive_ids(css_index* h, const int64_t* ids_host, int64_t n, int alive);

typedef enum css_clause_kind {
  CSS_CLAUSE_RANGE = 0, /* lo <= v <= hi (inclusive, on the int32 column value) */
  CSS_CLAUSE_SET = 1    /* 0 <= v < set_nbits and bit v of set_bits is 1 */
} css_clause_kind;

typedef struct css_clause {
  int32_t column;
  int32_t kind;             /* css_clause_kind */
  int32_t lo, hi;           /* RANGE */
  const uint32_t* set_bits; /* SET: HOST bitset, ceil(set_nbits/32) words */
  int32_t set_nbits;
  int32_t reserved;
} css_clause;

/* Conjunction of clauses AND alive AND (optional) explicit row bitmask. */
typedef struct css_filter {
  int32_t n_clauses;
  int32_t ignore_alive;       /* 0: dead rows never match (default) */
  const css_clause* clauses;  /* n_clauses entries */
  const uint32_t* row_mask;   /* optional HOST bitmask, ceil(ntotal/32) words, bit i = row i */
} css_filter;

/* Evaluate the filter over all rows on the device and copy the bitmask
 * (ceil(ntotal/32) words; bit i%32 of word i/32 = row i) to host.  This is the
 * bit-exact counterpart of evaluating src/storage.py:508-543 row by row. */
int css_index_filter_mask(css_index* h, const css_filter* f, uint32_t* mask_out_host,
                          int64_t* n_pass_out);

/* Exact top-k of nq queries (HOST float32 [nq, dim]) over the rows passing
 * `filter` (NULL = every alive row).  D: float32 [nq, k], I: int64 [nq, k],
 * both HOST.  Rows are ordered best first (IP: score descending; L2: squared
 * distance ascending), ties broken by ascending id; unfilled slots carry
 * id -1 and score -FLT_MAX (IP) / FLT_MAX (L2) like faiss.
 * nq < CSS_BATCH_MIN_NQ runs the HBM-bound streaming scan (inner product, d = 768, k <= 32: the two-phase
 * exact scan -- bf16 shadow sweep, proof, fp32 re-score -- otherwise one fp32 sweep); larger batches
 * run the tcgen05 score GEMM with in-epilogue candidate selection followed by
 * an exact fp32 re-score of the candidates. */
int css_index_search(css_index* h, const float* q_host, int nq, int k,
                     const css_filter* filter, float* D_host, int64_t* I_host);

/* Device-pointer variant: q_dev float32 [nq, dim], D_dev float32 [nq,k], I_dev
 * int64 [nq,k]; mask_dev is an optional DEVICE bitmask as produced by
 * css_index_filter_mask_device (NULL = all alive rows).  id_offset is added
 * to every returned id (global id of a row shard).  Asynchronous on `stream`. */
int css_index_search_device(css_index* h, const float* q_dev, int nq, int k,
                            const uint32_t* mask_dev, int64_t id_offset,
                            float* D_dev, int64_t* I_dev, void* stream);
/* Result exchange between the row shards of one search, fused into the scan kernel (SURVEY 8e's one
 * exchange step without NCCL): every rank creates one css_exchange on its GPU, the ranks all-gather the
 * CSS_IPC_HANDLE_BYTES-byte handles (rank order) and connect.  css_index_search_exchange_device is then
 * a COLLECTIVE: every rank calls it with the same queries / nq / k in the same order; the CTA that
 * finishes a query's local top-k stores it into every peer's memory over NVLink, waits for the peers'
 * lists and merges, so D_dev / I_dev hold the merged GLOBAL top-k on every rank when the kernel ends.
 * nq <= 64 (streaming-scan sizes; larger batches gather their lists with NCCL + css_topk_merge_device). */
typedef struct css_exchange css_exchange;
#define CSS_IPC_HANDLE_BYTES 64
int css_exchange_create(int device, int n_ranks, int rank, css_exchange** out,
                        unsigned char* handle_out /* CSS_IPC_HANDLE_BYTES, nullable */);
int css_exchange_connect(css_exchange* ex, const unsigned char* handles /* n_ranks x CSS_IPC_HANDLE_BYTES */);
int css_exchange_destroy(css_exchange* ex);
int css_index_search_exchange_device(css_index* h, css_exchange* ex, const float* q_dev, int nq, int k,
                                     const uint32_t* mask_dev, int64_t id_offset, float* D_dev,
                                     int64_t* I_dev, void* stream);

/* The same for ONE query in host memory (collective over the ranks like the call above): the query is staged
 * through the handle's pinned block, the merged top-k comes back through mapped host memory with a completion
 * flag the call polls -- no D2H copy, no stream synchronisation.  mask_dev: optional device row mask of this
 * shard.  Runs on the handle's own stream. */
int css_index_search_exchange(css_index* h, css_exchange* ex, const float* q_host, int k,
                              const uint32_t* mask_dev, int64_t id_offset, float* D_host, int64_t* I_host);

/* Counters of the two-phase batch-1 scan since the index was created: out = {queries answered by it,
 * queries it could not prove from the shadow-row lists (re-run by the fp32 sweep), 1 if the adaptive switch
 * currently bypasses a tier, largest ||x - bf16(x)|| of a stored row x 1e9, largest ||x - scale * int8(x)||
 * of a stored row x 1e9 (INT64_MAX once a non-finite row was stored), tier of the last scan call (2 = int8
 * shadow sweep, 1 = bf16 shadow sweep, 0 = fp32 sweep, -1 = none yet)}. */
int css_index_scan_stats(css_index* h, int64_t out[6]);

/* Evaluate a filter into the index's internal device mask and return its
 * device address (valid until the next call that changes the index/mask). */
int css_index_filter_mask_device(css_index* h, const css_filter* f,
                                 const uint32_t** mask_dev_out, int64_t* n_pass_out,
                                 void* stream);

/* Merge `n_lists` sorted top-k lists per query (e.g. one per GPU after the
 * NCCL all-gather): D_in float32 [n_lists, nq, k], I_in int64 [n_lists, nq, k],
 * all DEVICE; writes the merged best-k per query.  metric as in css_metric. */
int css_topk_merge_device(const float* D_in, const int64_t* I_in, int n_lists, int nq,
                          int k, int metric, float* D_out, int64_t* I_out, void* stream);
/* The same with list l at D_in + l * d_list_stride / I_in + l * i_list_stride (elements): lets every rank
 * ship its scores and ids in ONE packed buffer, i.e. one all-gather per search instead of two. */
int css_topk_merge_strided_device(const float* D_in, int64_t d_list_stride, const int64_t* I_in,
                                  int64_t i_list_stride, int n_lists, int nq, int k, int metric, float* D_out,
                                  int64_t* I_out, void* stream);

/* faiss-compatible persistence: IndexFlatIP ("IxFI") / IndexFlatL2 ("IxF2")
 * files as written by faiss.write_index (src/storage.py:879-884). */
int css_index_save(css_index* h, const char* path);
int css_index_load(css_index* h, const char* path);


/* ------------------------------------------------------------------ */
/* MPNet chunk encoder (half A of the hot path)                        */
/* ------------------------------------------------------------------ */
/* Replaces SentenceTransformer("all-mpnet-base-v2").encode(...) as called at
 * src/embeddings.py:184-188 (single text) and :216-222 (batch): MPNetModel forward
 * (transformers modeling_mpnet.py), masked mean pooling and L2 normalisation, on
 * already-tokenised input.  bf16 tensor-core GEMMs with fp32 accumulation, fp32
 * LayerNorm / softmax / pooling. */
typedef struct css_encoder css_encoder;

typedef struct css_mpnet_config {
  int32_t vocab_size;          /* 30527 */
  int32_t hidden_size;         /* 768  (only value this build serves) */
  int32_t num_layers;          /* 12 */
  int32_t num_heads;           /* 12   (head dim must be 64) */
  int32_t intermediate_size;   /* 3072 */
  int32_t max_position;        /* 514  rows of the position table */
  int32_t rel_buckets;         /* 32 */
  int32_t rel_max_distance;    /* 128 */
  int32_t pad_token_id;        /* 1: position ids are pad_token_id + 1 + index */
  float layer_norm_eps;        /* 1e-5 */
} css_mpnet_config;

/* One transformer layer; every pointer is HOST float32, nn.Linear layout [out, in]. */
typedef struct css_mpnet_layer {
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b; /* attention.attn.{q,k,v,o} */
  const float *ln1_w, *ln1_b;                                 /* attention.LayerNorm */
  const float *ffn1_w, *ffn1_b;                               /* intermediate.dense [3072,768] */
  const float *ffn2_w, *ffn2_b;                               /* output.dense [768,3072] */
  const float *ln2_w, *ln2_b;                                 /* output.LayerNorm */
} css_mpnet_layer;

typedef struct css_mpnet_weights {
  const float* word_emb;   /* [vocab_size, hidden] */
  const float* pos_emb;    /* [max_position, hidden] */
  const float *emb_ln_w, *emb_ln_b;
  const float* rel_bias;   /* encoder.relative_attention_bias.weight [rel_buckets, num_heads] */
  const css_mpnet_layer* layers; /* num_layers entries */
} css_mpnet_weights;

/* Bucket of a relative position (memory - context), the arithmetic of
 * MPNetEncoder.relative_position_bucket.  Pure host function (no device needed). */
int css_mpnet_relative_bucket(int relative_position, int num_buckets, int max_distance);

/* Uploads the weights (linear weights converted to bf16, Q/K/V fused) and sizes the
 * activation workspace for `max_tokens` tokens per forward pass (0 = default). */
int css_encoder_create(const css_mpnet_config* cfg, const css_mpnet_weights* w, int device,
                       int64_t max_tokens, css_encoder** out);
int css_encoder_destroy(css_encoder* h);
int css_encoder_dim(const css_encoder* h);
int64_t css_encoder_max_tokens(const css_encoder* h);
/* Longest sequence one call accepts (max_position - pad_token_id - 1, at most 512). */
int css_encoder_max_seq_len(const css_encoder* h);

/* Encode n_seq token sequences packed back to back: ids[cu_seqlens[i] .. cu_seqlens[i+1])
 * is sequence i (already truncated, with its <s> ... </s> framing, no padding).
 * out: HOST float32 [n_seq, hidden]: mean over the sequence's tokens of the last
 * hidden state (sentence-transformers Pooling, mean mode), then if normalize != 0
 * divided by max(||.||_2, 1e-12) (sentence-transformers Normalize).  Any number of
 * tokens: the call splits into passes of at most max_tokens. */
int css_encoder_encode(css_encoder* h, const int32_t* ids_host, const int32_t* cu_seqlens_host,
                       int32_t n_seq, int normalize, float* out_host);
/* Device variant: ids_dev / cu_seqlens_dev / out_dev on the encoder's device;
 * cu_seqlens_host is still needed to size the launch.  total tokens <= max_tokens.
 * Asynchronous on `stream` (NULL = the handle's stream). */
int css_encoder_encode_device(css_encoder* h, const int32_t* ids_dev, const int32_t* cu_seqlens_dev,
                              const int32_t* cu_seqlens_host, int32_t n_seq, int normalize,
                              float* out_dev, void* stream);

/* ------------------------------------------------------------------ */
/* Batch WordPiece tokenizer (host side of half A)                     */
/* ------------------------------------------------------------------ */
/* Replaces the tokenisation inside SentenceTransformer.encode (src/embeddings.py:216-222), i.e. the
 * fast MPNetTokenizer pipeline BertNormalizer -> BertPreTokenizer -> WordPiece over vocab.txt with
 * <s> ... </s> framing and truncation to max_len tokens; multi-threaded, producing directly the packed
 * ids / cu_seqlens that css_encoder_encode consumes.  Any valid UTF-8 is tokenised natively (Unicode
 * tables generated from that pipeline, scripts/gen_unicode_tables.py).  css_tokenizer_add_special
 * registers a literal (e.g. "<s>", "<mask>") that is cut out of the RAW text and mapped straight to
 * `id`, leftmost-longest, like the reference's added special tokens.  Malformed UTF-8 is NOT
 * tokenised: needs_fallback[i] = 1 and sequence i is empty.  ids_out holds n * max_len entries. */
typedef struct css_tokenizer css_tokenizer;
int css_tokenizer_create(const char* vocab_path, int do_lower_case, css_tokenizer** out);
int css_tokenizer_destroy(css_tokenizer* h);
int css_tokenizer_vocab_size(const css_tokenizer* h);
int css_tokenizer_add_special(css_tokenizer* h, const char* literal, int32_t id);
int css_tokenizer_encode_batch(css_tokenizer* h, const char* const* texts, const int64_t* lens, int32_t n,
                               int32_t max_len, int32_t* ids_out, int32_t* cu_seqlens_out,
                               uint8_t* needs_fallback, int32_t n_threads);

/* Diagnostic entry points used by the kernel-level parity tests (HOST float32 buffers,
 * rounded to bf16 on the device exactly as the encoder does):
 *   gemm:      out[M,N] = bf16(A[M,K]) * bf16(B[N,K])^T + bias[N], optionally GELU(erf),
 *              through the tcgen05 kernel of the encoder; out is the bf16 result widened.
 *   attention: ctx[T,768] from packed qkv[T,2304] (q | k | v), per-head relative bias
 *              rel_table[12][2*rel_half+1] indexed by (key - query) + rel_half. */
int css_debug_gemm(const float* A, const float* B, const float* bias, int M, int N, int K, int gelu,
                   int device, float* out);
/*   gemm_resid_ln: out[M,768] = LayerNorm(bf16(A[M,K]) * bf16(B[768,K])^T + bias + bf16(resid[M,768])) * gamma + beta
 *              through the fused epilogue of the attention-output / FFN-down projections
 *              (modeling_mpnet.py:183-185,239-243); out is the bf16 result widened.  mode bit 0: 2-CTA
 *              tiles, bit 1: LayerNorm inside the GEMM epilogue (else epilogue statistics + apply kernel). */
int css_debug_gemm_resid_ln(const float* A, const float* B, const float* bias, const float* resid,
                            const float* gamma, const float* beta, int M, int K, float eps, int mode,
                            int device, float* out);
int css_debug_attention(const float* qkv, const int32_t* cu_seqlens, int n_seq, const float* rel_table,
                        int rel_half, int device, float* ctx);

/*   scan_bf16: phase 1 alone of the two-phase batch-1 scan (the bf16 shadow sweep that leaves the per-block
 *              candidate lists), on `stream`, for the roofline measurement of bench.py; no result is produced. */
int css_debug_scan_bf16(css_index* h, const float* q_dev, int nq, void* stream);
/*   scan_int8: the same for the first tier, the int8 shadow sweep (768 + 4 bytes per 768-d row). */
int css_debug_scan_int8(css_index* h, const float* q_dev, int nq, void* stream);
/*   scan_trace: one batch-1 scan (current tier, top-k) with a timeline: out_host receives (scan blocks + 1) x 8
 *              %globaltimer stamps (ns) -- row b = block b {start, sweep done, block list merged, list re-scored,
 *              ticket drawn}, last row = the finishing block {slot 0: lists requested, slot 4: result emitted};
 *              unused slots are 0.  Profiling aid (scripts/scan_trace.py). */
int css_debug_scan_trace(css_index* h, const float* q_dev, int k, int64_t* out_host, int n_out, void* stream);

/* Process-wide switches (also read from the environment at load: CSS_SCAN_BF16, CSS_SCAN_INT8, CSS_SCAN_INTERLEAVE,
 * CSS_SCAN_LIST, CSS_SCAN_ADAPTIVE): "scan_bf16" 1 = two-phase batch-1 scan, 0 = single fp32 sweep;
 * "scan_int8" 1 = the two-phase scan sweeps the int8 shadow rows first (0: the bf16 shadow rows);
 * "scan_interleave" 1 = block-cyclic 8-row units in the bf16 sweep; "scan_list" 32 | 64 = per-block list length
 * (0: chosen by k); "scan_adaptive" 1 = bypass phase 1 while most queries cannot be proven.  Results are exact
 * under every setting; benchmarks use this to time the paths side by side in one process. */
int css_set_option(const char* name, int value);

/* Timing hook for benchmarks: number of kernels this library has launched in
 * this process (all handles). */
int64_t css_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CSS_B200_H */
