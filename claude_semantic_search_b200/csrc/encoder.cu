// encoder.cu -- host side of the MPNet chunk encoder C ABI (include/css_b200.h).
//
// Forward pass of one packed batch of T tokens (n_seq sequences):
//   E1  embed + LayerNorm                      -> x    bf16 [T, 768]
//   per layer (12x):
//     E2  x * Wqkv^T + b                       -> qkv  bf16 [T, 2304]   tcgen05 GEMM
//     E3  softmax(q k^T / 8 + rel_bias) v      -> ctx  bf16 [T, 768]
//     E4  LayerNorm(ctx * Wo^T + b + x)        -> x1   bf16 [T, 768]    tcgen05 GEMM, LayerNorm in the epilogue
//     E6a gelu(x1 * W1^T + b)                  -> h    bf16 [T, 3072]   tcgen05 GEMM
//     E6b LayerNorm(h * W2^T + b + x1)         -> x    bf16 [T, 768]    tcgen05 GEMM, LayerNorm in the epilogue
//   E7  mean over tokens, L2 normalise         -> out  f32  [n_seq, 768]
#include "encoder_kernels.cuh"
#include "attention_tc.cuh"
#include "gemm_tc.cuh"
#include "query_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

using namespace css;
using namespace css::enc;

struct EncLayer {
  __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
  // LayerNorm-folded pipeline (EpiBiasBf16<., true>): weights scaled by the gamma of the LayerNorm that
  // feeds them, their column sums c and the constants d = b + W beta.  wqkv_f folds the PREVIOUS layer's
  // output LayerNorm (layer 0 has none: its input is the embedding LayerNorm's output).
  __nv_bfloat16 *wqkv_f = nullptr, *w1_f = nullptr;
  float *cqkv = nullptr, *dqkv = nullptr, *c1 = nullptr, *d1 = nullptr;
};

struct css_encoder {
  css_mpnet_config cfg{};
  int device = 0;
  int n_sm = 148;
  int max_seq = 512;
  int64_t max_tokens = 0;
  cudaStream_t stream = nullptr;
  // weights
  float *word_emb = nullptr, *pos_emb = nullptr, *emb_ln_w = nullptr, *emb_ln_b = nullptr;
  float* rel_table = nullptr;  // [heads][2*rel_half+1]
  float* rel_max = nullptr;    // [heads] max of each head's table (softmax upper bound)
  int rel_half = 0;
  std::vector<EncLayer> layers;
  std::vector<void*> owned;  // every device allocation, freed on destroy
  // workspace
  __nv_bfloat16 *x = nullptr, *x1 = nullptr, *qkv = nullptr, *ctx = nullptr, *h = nullptr;
  float2* ln_stats = nullptr;   // [max_tokens][3][4] partial row statistics written by the EpiResidLN epilogues (query path: [64][96])
  float2 *mr1 = nullptr, *mr2 = nullptr;   // [max_tokens] (mean, rstd) of the attention / output LayerNorm inputs
  int32_t *ids_dev = nullptr, *cu_dev = nullptr;
  float* out_dev = nullptr;
  int64_t max_seqs = 0;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // Single-query latency path: the 62 launches of one forward pass over one short
  // sequence are captured once per (token bucket, normalize) into a CUDA graph and replayed.
  struct QueryGraph {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  std::map<int, QueryGraph> query_graphs;
  std::mutex mu;
};

namespace {

template <typename T>
int enc_alloc(css_encoder* e, T** p, size_t count) {
  *p = nullptr;
  if (count == 0) return CSS_OK;
  cudaError_t err = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(err));
    return CSS_ERR_OOM;
  }
  e->owned.push_back(*p);
  return CSS_OK;
}

int upload_f32(css_encoder* e, float** dst, const float* src, size_t n) {
  CSS_REQUIRE(src != nullptr, "a weight pointer is NULL");
  CSS_CHECK(enc_alloc(e, dst, n));
  CSS_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  return CSS_OK;
}

// Upload fp32 [n] into a bf16 buffer slice through a device staging buffer.
int upload_bf16_into(css_encoder* e, __nv_bfloat16* dst, const float* src, size_t n, float* stage) {
  CSS_REQUIRE(src != nullptr, "a weight pointer is NULL");
  CSS_CUDA(cudaMemcpyAsync(stage, src, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(stage, dst, (int64_t)n);
  CSS_LAUNCHED();
  // `stage` is reused by the next upload: same stream, so ordering is preserved, but the
  // pageable source must be consumed before the caller may free it -> sync here.
  CSS_CUDA(cudaStreamSynchronize(e->stream));
  return CSS_OK;
}

float bf16_round_host(float x) {   // round-to-nearest-even to bf16, as __float2bfloat16_rn
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

// W [N, K] (nn.Linear layout), LayerNorm (gamma, beta) [K] feeding it, bias b [N]:
//   Wf = W * gamma (per input column), c[n] = sum_k bf16(Wf[n, k]), d[n] = b[n] + sum_k beta[k] W[n, k]
void fold_layernorm(const float* W, const float* b, const float* gamma, const float* beta, size_t N, size_t K,
                    float* Wf, float* c, float* d) {
  for (size_t n = 0; n < N; ++n) {
    double cs = 0.0, ds = b[n];
    for (size_t k = 0; k < K; ++k) {
      const float wf = W[n * K + k] * gamma[k];
      Wf[n * K + k] = wf;
      cs += (double)bf16_round_host(wf);
      ds += (double)beta[k] * (double)W[n * K + k];
    }
    c[n] = (float)cs;
    d[n] = (float)ds;
  }
}

int ensure_pinned(css_encoder* e, size_t bytes) {
  if (bytes <= e->pinned_bytes) return CSS_OK;
  if (e->pinned) {
    cudaStreamSynchronize(e->stream);
    cudaFreeHost(e->pinned);
  }
  e->pinned = nullptr;
  e->pinned_bytes = 0;
  size_t want = std::max<size_t>(bytes, (size_t)1 << 20);
  cudaError_t err = cudaMallocHost(&e->pinned, want);
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(err));
    return CSS_ERR_OOM;
  }
  e->pinned_bytes = want;
  return CSS_OK;
}

// One forward pass over T packed tokens; everything on `st`.
template <int MODE>
int skinny_gemm(int T, int N, int K, const SkinnyParams& p, cudaStream_t st) {
  cudaError_t ce;
  // 64-row bucket: the wide (N >= 2304) projections run two 8-column groups per CTA on the same activation
  // fragments (same box, p50: 0.534 -> 0.471 ms; the 32-row bucket loses 4 % with it and keeps one group).
  // CSS_QUERY_NT=1 / 2 forces one / two groups for both buckets.
  static const int nt_env = [] { const char* v = getenv("CSS_QUERY_NT"); return v ? atoi(v) : 0; }();
  if constexpr (MODE != kSkResidLN) {
    const bool two = nt_env ? nt_env == 2 : T == 64;
    if (K == kHidden && N >= 3 * kHidden && two) {
      CSS_CUDA(T == 32 ? (skinny_launch<2, 4, 6, MODE, 2>(p, N, st)) : (skinny_launch<4, 4, 6, MODE, 2>(p, N, st)));
      CSS_LAUNCHED();
      return CSS_OK;
    }
  }
  if (K == kHidden) ce = T == 32 ? skinny_launch<2, 4, 6, MODE>(p, N, st) : skinny_launch<4, 4, 6, MODE>(p, N, st);
  else ce = T == 32 ? skinny_launch<2, 8, 12, MODE>(p, N, st) : skinny_launch<4, 8, 12, MODE>(p, N, st);
  CSS_CUDA(ce);
  CSS_LAUNCHED();
  return CSS_OK;
}

// The interactive query path: T = 32 or 64 token rows, sequences of at most 64 tokens (query_kernels.cuh).
// Same LayerNorm-folded pipeline as forward(), weight-streaming kernels instead of 128-row tiles; the row
// statistics stay as per-CTA partials (two buffers: attention-output / layer-output LayerNorm) that the
// consuming kernels reduce themselves, and every kernel is a programmatic dependent launch whose
// weight loads overlap its predecessor's tail.
int forward_query(css_encoder* e, const int32_t* cu_dev, int n_seq, int T, int normalize, float* out_dev,
                  cudaStream_t st) {
  const css_mpnet_config& c = e->cfg;
  float2* parts1 = e->ln_stats;                                       // statistics of x1 (attention-output LayerNorm input)
  float2* parts2 = e->ln_stats + (size_t)kQueryMaxRows * kQueryParts;  // statistics of x (layer-output LayerNorm input)
  for (int l = 0; l < c.num_layers; ++l) {
    const EncLayer& w = e->layers[l];
    const EncLayer* prev = l > 0 ? &e->layers[l - 1] : nullptr;
    {
      SkinnyParams p{};
      p.A = e->x;
      p.out = e->qkv;
      p.ldo = 3 * kHidden;
      p.eps = c.layer_norm_eps;
      if (l == 0) {
        p.W = w.wqkv;
        p.bias = w.bqkv;
        CSS_CHECK(skinny_gemm<kSkBias>(T, 3 * kHidden, kHidden, p, st));
      } else {
        p.W = w.wqkv_f;
        p.bias = w.dqkv;
        p.colsum = w.cqkv;
        p.stat_parts = parts2;
        CSS_CHECK(skinny_gemm<kSkFold>(T, 3 * kHidden, kHidden, p, st));
      }
    }
    CSS_CUDA(pdl_launch(query_attention_kernel, dim3(kHeads, (unsigned)n_seq, kQueryAttnRowGroups), dim3(kQueryAttnThreads), st,
                        (const __nv_bfloat16*)e->qkv, cu_dev, (const float*)e->rel_table, e->rel_half, e->ctx));
    CSS_LAUNCHED();
    {
      SkinnyParams p{};
      p.A = e->ctx;
      p.W = w.wo;
      p.out = e->x1;
      p.ldo = kHidden;
      p.bias = w.bo;
      p.eps = c.layer_norm_eps;
      p.resid = e->x;
      p.stat_parts = prev ? parts2 : nullptr;
      p.rgamma = prev ? prev->ln2_w : nullptr;
      p.rbeta = prev ? prev->ln2_b : nullptr;
      p.parts = parts1;
      CSS_CHECK(skinny_gemm<kSkResidLN>(T, kHidden, kHidden, p, st));
    }
    {
      SkinnyParams p{};
      p.A = e->x1;
      p.W = w.w1_f;
      p.out = e->h;
      p.ldo = kFfn;
      p.bias = w.d1;
      p.colsum = w.c1;
      p.eps = c.layer_norm_eps;
      p.stat_parts = parts1;
      CSS_CHECK(skinny_gemm<kSkFoldGelu>(T, kFfn, kHidden, p, st));
    }
    {
      SkinnyParams p{};
      p.A = e->h;
      p.W = w.w2;
      p.out = e->x;
      p.ldo = kHidden;
      p.bias = w.b2;
      p.eps = c.layer_norm_eps;
      p.resid = e->x1;
      p.stat_parts = parts1;
      p.rgamma = w.ln1_w;
      p.rbeta = w.ln1_b;
      p.parts = parts2;
      CSS_CHECK(skinny_gemm<kSkResidLN>(T, kHidden, kFfn, p, st));
    }
  }
  CSS_CUDA(pdl_launch(ln_stats_finalize_parts_kernel, dim3((unsigned)((T + 3) / 4)), dim3(128), st,
                      (const float2*)parts2, T, c.layer_norm_eps, e->mr2));
  CSS_LAUNCHED();
  const EncLayer& last = e->layers[c.num_layers - 1];
  pool_normalize_ln_kernel<<<(unsigned)n_seq, 256, 0, st>>>(e->x, e->mr2, last.ln2_w, last.ln2_b, cu_dev, normalize,
                                                            out_dev);
  CSS_LAUNCHED();
  return CSS_OK;
}

int forward(css_encoder* e, const int32_t* ids_dev, const int32_t* cu_dev, int n_seq, int T, int max_len,
            int normalize, float* out_dev, cudaStream_t st, bool query_path = false) {
  const css_mpnet_config& c = e->cfg;
  const int warps_per_block = 8;
  const unsigned row_blocks = (unsigned)((T + warps_per_block - 1) / warps_per_block);
  const unsigned ln_blocks = (unsigned)((T + 8 * kLnApplyRows - 1) / (8 * kLnApplyRows));
  embed_ln_kernel<<<row_blocks, warps_per_block * 32, 0, st>>>(ids_dev, cu_dev, n_seq, T, e->word_emb, e->pos_emb,
                                                               c.vocab_size, c.max_position, c.pad_token_id,
                                                               e->emb_ln_w, e->emb_ln_b, c.layer_norm_eps, e->x);
  CSS_LAUNCHED();
  const int Lp = (max_len + 63) & ~63;
  const size_t attn_smem = (size_t)Lp * 256 + (size_t)2 * Lp * sizeof(float);
  CSS_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
  const dim3 attn_grid((unsigned)((max_len + kAttnQRows - 1) / kAttnQRows), kHeads, (unsigned)n_seq);
  // tcgen05 attention (CSS_ATTN_TC=0 selects the mma.sync kernel); longer sequences always take mma.sync
  static const bool attn_tc_env = [] { const char* v = getenv("CSS_ATTN_TC"); return v ? atoi(v) != 0 : true; }();
  const bool attn_tc = attn_tc_env && max_len <= kAttnTcMaxLen;
  // CSS_LN_FUSED=1: LayerNorm inside the attention-output GEMM epilogue (panel order); 0 (default, measured
  // 127 + 61 us against 196 us): epilogue statistics + apply kernel
  static const bool ln_fused = [] { const char* v = getenv("CSS_LN_FUSED"); return v ? atoi(v) != 0 : false; }();
  // CSS_LN_FOLD=1 (default): no normalised activation is ever materialised -- the projections emit the
  // pre-LayerNorm value + row statistics, the consuming GEMMs fold the LayerNorm into their weights /
  // epilogue, the residual adds rebuild LN(.) in fp32.  0: ln_apply_kernel after each projection.
  static const bool ln_fold = [] { const char* v = getenv("CSS_LN_FOLD"); return v ? atoi(v) != 0 : true; }();
  // CSS_QUERY_SKINNY=1 (default): the single-query graph path over a bucket of 32 or 64 token rows runs the
  // weight-streaming kernels of query_kernels.cuh (batch passes never do: their result must not depend on
  // how sequences happen to be grouped)
  static const bool query_skinny = [] { const char* v = getenv("CSS_QUERY_SKINNY"); return v ? atoi(v) != 0 : true; }();
  if (query_path && ln_fold && query_skinny && (T == 32 || T == 64) && max_len <= kQueryMaxRows)
    return forward_query(e, cu_dev, n_seq, T, normalize, out_dev, st);
  if (ln_fold) {
    const unsigned fin_blocks = (unsigned)((T + 255) / 256);
    for (int l = 0; l < c.num_layers; ++l) {
      const EncLayer& w = e->layers[l];
      const EncLayer* prev = l > 0 ? &e->layers[l - 1] : nullptr;
      if (l == 0) {   // e->x holds the embedding LayerNorm's output
        EpiBiasBf16<false>::Params p{e->qkv, w.bqkv, 3 * kHidden, nullptr, nullptr};
        CSS_CHECK((gemm::run<256, EpiBiasBf16<false>>(e->x, kHidden, w.wqkv, kHidden, T, 3 * kHidden, kHidden, 0, p,
                                                        e->n_sm, st)));
      } else {        // e->x holds the previous layer's pre-LayerNorm output, mr2 its row statistics
        EpiBiasBf16<false, true>::Params p{e->qkv, w.dqkv, 3 * kHidden, w.cqkv, e->mr2};
        CSS_CHECK((gemm::run<256, EpiBiasBf16<false, true>>(e->x, kHidden, w.wqkv_f, kHidden, T, 3 * kHidden, kHidden,
                                                              0, p, e->n_sm, st)));
      }
      if (attn_tc) {
        CSS_CHECK(attention_tc_launch(e->qkv, T, cu_dev, n_seq, max_len, e->rel_table, e->rel_max, e->rel_half,
                                      e->ctx, e->n_sm, st));
      } else {
        attention_kernel<<<attn_grid, kAttnThreads, attn_smem, st>>>(e->qkv, cu_dev, e->rel_table, e->rel_half,
                                                                     e->ctx);
        CSS_LAUNCHED();
      }
      {
        EpiResidLN<false>::Params p{e->x1, w.bo, e->x, w.ln1_w, w.ln1_b, c.layer_norm_eps, e->ln_stats,
                                    prev ? e->mr2 : nullptr, prev ? prev->ln2_w : nullptr, prev ? prev->ln2_b : nullptr};
        CSS_CHECK((gemm::run<256, EpiResidLN<false>>(e->ctx, kHidden, w.wo, kHidden, T, kHidden, kHidden, 0, p,
                                                       e->n_sm, st)));
      }
      ln_stats_finalize_kernel<<<fin_blocks, 256, 0, st>>>(e->ln_stats, T, c.layer_norm_eps, e->mr1);
      CSS_LAUNCHED();
      {
        EpiBiasBf16<true, true>::Params p{e->h, w.d1, kFfn, w.c1, e->mr1};
        CSS_CHECK((gemm::run<256, EpiBiasBf16<true, true>>(e->x1, kHidden, w.w1_f, kHidden, T, kFfn, kHidden, 0, p,
                                                             e->n_sm, st)));
      }
      {
        EpiResidLN<false>::Params p{e->x, w.b2, e->x1, w.ln2_w, w.ln2_b, c.layer_norm_eps, e->ln_stats,
                                    e->mr1, w.ln1_w, w.ln1_b};
        CSS_CHECK((gemm::run<256, EpiResidLN<false>>(e->h, kFfn, w.w2, kFfn, T, kHidden, kFfn, 0, p, e->n_sm, st)));
      }
      ln_stats_finalize_kernel<<<fin_blocks, 256, 0, st>>>(e->ln_stats, T, c.layer_norm_eps, e->mr2);
      CSS_LAUNCHED();
    }
    const EncLayer& last = e->layers[c.num_layers - 1];
    pool_normalize_ln_kernel<<<(unsigned)n_seq, 256, 0, st>>>(e->x, e->mr2, last.ln2_w, last.ln2_b, cu_dev, normalize,
                                                              out_dev);
    CSS_LAUNCHED();
    return CSS_OK;
  }

  for (int l = 0; l < c.num_layers; ++l) {
    const EncLayer& w = e->layers[l];
    {
      EpiBiasBf16<false>::Params p{e->qkv, w.bqkv, 3 * kHidden, nullptr, nullptr};
      CSS_CHECK((gemm::run<256, EpiBiasBf16<false>>(e->x, kHidden, w.wqkv, kHidden, T, 3 * kHidden, kHidden, 0, p,
                                                      e->n_sm, st)));
    }
    if (attn_tc) {
      CSS_CHECK(attention_tc_launch(e->qkv, T, cu_dev, n_seq, max_len, e->rel_table, e->rel_max, e->rel_half, e->ctx,
                                    e->n_sm, st));
    } else {
      attention_kernel<<<attn_grid, kAttnThreads, attn_smem, st>>>(e->qkv, cu_dev, e->rel_table, e->rel_half, e->ctx);
      CSS_LAUNCHED();
    }
    if (ln_fused) {
      EpiResidLN<true>::Params p{e->x1, w.bo, e->x, w.ln1_w, w.ln1_b, c.layer_norm_eps, nullptr, nullptr, nullptr, nullptr};
      CSS_CHECK((gemm::run<256, EpiResidLN<true>>(e->ctx, kHidden, w.wo, kHidden, T, kHidden, kHidden, 0, p, e->n_sm,
                                                    st)));
    } else {
      EpiResidLN<false>::Params p{e->x1, w.bo, e->x, w.ln1_w, w.ln1_b, c.layer_norm_eps, e->ln_stats, nullptr, nullptr, nullptr};
      CSS_CHECK((gemm::run<256, EpiResidLN<false>>(e->ctx, kHidden, w.wo, kHidden, T, kHidden, kHidden, 0, p, e->n_sm,
                                                     st)));
      ln_apply_kernel<<<ln_blocks, 256, 0, st>>>(e->x1, e->ln_stats, T, w.ln1_w, w.ln1_b,
                                                                   c.layer_norm_eps);
      CSS_LAUNCHED();
    }
    {
      EpiBiasBf16<true>::Params p{e->h, w.b1, kFfn, nullptr, nullptr};
      CSS_CHECK((gemm::run<256, EpiBiasBf16<true>>(e->x1, kHidden, w.w1, kHidden, T, kFfn, kHidden, 0, p, e->n_sm,
                                                     st)));
    }
    {
      EpiResidLN<false>::Params p{e->x, w.b2, e->x1, w.ln2_w, w.ln2_b, c.layer_norm_eps, e->ln_stats, nullptr, nullptr, nullptr};
      CSS_CHECK((gemm::run<256, EpiResidLN<false>>(e->h, kFfn, w.w2, kFfn, T, kHidden, kFfn, 0, p, e->n_sm, st)));
    }
    ln_apply_kernel<<<ln_blocks, 256, 0, st>>>(e->x, e->ln_stats, T, w.ln2_w, w.ln2_b,
                                                                 c.layer_norm_eps);
    CSS_LAUNCHED();
  }
  pool_normalize_kernel<<<(unsigned)n_seq, 256, 0, st>>>(e->x, cu_dev, normalize, out_dev);
  CSS_LAUNCHED();
  return CSS_OK;
}

// Token bucket of the single-query graph path (0: not served by a graph).
int query_bucket(int T) {
  static const bool enabled = [] { const char* v = getenv("CSS_QUERY_GRAPH"); return v ? atoi(v) != 0 : true; }();
  if (!enabled || T > kAttnTcMaxLen) return 0;
  for (int b : {32, 64, 128, 256, 384})
    if (T <= b) return b;
  return 0;
}

// One sequence of <= 384 tokens already in e->ids_dev / e->cu_dev: replay (or first capture) the graph
// of a forward pass over `bucket` token rows.  Rows past the sequence end hold stale finite values;
// every kernel is row-wise (GEMM, LayerNorm) or bounded by cu_seqlens (attention, pooling), so they
// never reach the result.
int forward_query_graph(css_encoder* e, int bucket, int normalize, cudaStream_t st) {
  const int key = bucket * 2 + (normalize ? 1 : 0);
  auto it = e->query_graphs.find(key);
  if (it == e->query_graphs.end()) {
    const int64_t l0 = g_launches.load(std::memory_order_relaxed);
    CSS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    const int rc = forward(e, e->ids_dev, e->cu_dev, 1, bucket, bucket, normalize, e->out_dev, st, true);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != CSS_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (ce != cudaSuccess || !graph) {
      (void)cudaGetLastError();
      set_error("graph capture of the query path failed: %s", cudaGetErrorString(ce));
      return CSS_ERR_CUDA;
    }
    css_encoder::QueryGraph qg;
    const cudaError_t ie = cudaGraphInstantiate(&qg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
      return CSS_ERR_CUDA;
    }
    qg.launches = g_launches.load(std::memory_order_relaxed) - l0;
    g_launches.store(l0, std::memory_order_relaxed);   // counted per replay below
    it = e->query_graphs.emplace(key, qg).first;
  }
  CSS_CUDA(cudaGraphLaunch(it->second.exec, st));
  g_launches.fetch_add(it->second.launches, std::memory_order_relaxed);
  return CSS_OK;
}

int validate_cu(const css_encoder* e, const int32_t* cu, int n_seq, int* max_len_out) {
  CSS_REQUIRE(cu[0] == 0, "cu_seqlens[0] must be 0");
  int mx = 0;
  for (int i = 0; i < n_seq; ++i) {
    const int len = cu[i + 1] - cu[i];
    CSS_REQUIRE(len >= 1, "sequence %d is empty (length %d)", i, len);
    CSS_REQUIRE(len <= e->max_seq, "sequence %d has %d tokens, limit is %d", i, len, e->max_seq);
    mx = std::max(mx, len);
  }
  *max_len_out = mx;
  return CSS_OK;
}

}  // namespace

extern "C" {

int css_mpnet_relative_bucket(int relative_position, int num_buckets, int max_distance) {
  // MPNetEncoder.relative_position_bucket (modeling_mpnet.py:338-357), float32 arithmetic
  // in the order torch evaluates it.
  int ret = 0;
  int n = -relative_position;
  num_buckets /= 2;
  if (n < 0) {
    ret += num_buckets;
    n = -n;
  }
  const int max_exact = num_buckets / 2;
  if (n < max_exact) return ret + n;
  float v = logf((float)n / (float)max_exact);
  v = v / (float)log((double)max_distance / (double)max_exact);
  v = v * (float)(num_buckets - max_exact);
  int large = max_exact + (int)v;
  if (large > num_buckets - 1) large = num_buckets - 1;
  return ret + large;
}

int css_encoder_create(const css_mpnet_config* cfg, const css_mpnet_weights* w, int device, int64_t max_tokens,
                       css_encoder** out) {
  CSS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  CSS_REQUIRE(cfg != nullptr && w != nullptr, "cfg / weights is NULL");
  if (cfg->hidden_size != kHidden || cfg->num_heads != kHeads || cfg->intermediate_size != kFfn) {
    set_error("this build serves hidden=768, heads=12, ffn=3072 (got %d, %d, %d)", cfg->hidden_size, cfg->num_heads,
              cfg->intermediate_size);
    return CSS_ERR_UNSUPPORTED;
  }
  CSS_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= 64, "num_layers %d out of range", cfg->num_layers);
  CSS_REQUIRE(cfg->vocab_size >= 2 && cfg->max_position >= 4, "bad vocab_size / max_position");
  CSS_REQUIRE(cfg->rel_buckets >= 4 && cfg->rel_buckets % 2 == 0 && cfg->rel_max_distance > cfg->rel_buckets / 4,
              "bad relative-position bucket configuration");
  CSS_REQUIRE(w->layers != nullptr, "weights->layers is NULL");
  CSS_CHECK(ensure_device(device));
  DeviceGuard g(device);
  css_encoder* e = new (std::nothrow) css_encoder();
  if (!e) {
    set_error("out of host memory");
    return CSS_ERR_OOM;
  }
  e->cfg = *cfg;
  e->device = device;
  e->n_sm = sm_count(device);
  e->max_seq = std::min(kMaxSeq, cfg->max_position - cfg->pad_token_id - 1);
  // default pass size: 148 SMs x 768 tokens = 296 chunks of 384 tokens -> 444 row panels of 256, 3996 / 5328 / 1332
  // GEMM tiles and 3552 attention units, all multiples of the 74 CTA pairs / 148 CTAs (no tail wave)
  if (max_tokens <= 0) max_tokens = (int64_t)e->n_sm * 768;
  max_tokens = std::max<int64_t>(max_tokens, e->max_seq);
  max_tokens = (max_tokens + 127) / 128 * 128;
  e->max_tokens = max_tokens;
  e->max_seqs = max_tokens;  // every sequence has >= 1 token
  int rc = CSS_OK;
  auto fail = [&](int code) {
    css_encoder_destroy(e);
    return code;
  };
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    return fail(CSS_ERR_CUDA);
  }
  const size_t H = kHidden;
  if ((rc = upload_f32(e, &e->word_emb, w->word_emb, (size_t)cfg->vocab_size * H)) != CSS_OK) return fail(rc);
  if ((rc = upload_f32(e, &e->pos_emb, w->pos_emb, (size_t)cfg->max_position * H)) != CSS_OK) return fail(rc);
  if ((rc = upload_f32(e, &e->emb_ln_w, w->emb_ln_w, H)) != CSS_OK) return fail(rc);
  if ((rc = upload_f32(e, &e->emb_ln_b, w->emb_ln_b, H)) != CSS_OK) return fail(rc);
  // dense relative-position bias table: rel_table[h][d + rel_half] = rel_bias[bucket(d)][h]
  {
    if (!w->rel_bias) {
      set_error("weights->rel_bias is NULL");
      return fail(CSS_ERR_INVALID);
    }
    e->rel_half = e->max_seq - 1;
    const int width = 2 * e->rel_half + 1;
    std::vector<float> table((size_t)kHeads * width);
    for (int d = -e->rel_half; d <= e->rel_half; ++d) {
      const int b = css_mpnet_relative_bucket(d, cfg->rel_buckets, cfg->rel_max_distance);
      for (int hh = 0; hh < kHeads; ++hh) table[(size_t)hh * width + d + e->rel_half] = w->rel_bias[(size_t)b * kHeads + hh];
    }
    std::vector<float> tmax(kHeads, -INFINITY);
    for (int hh = 0; hh < kHeads; ++hh)
      for (int d = 0; d < width; ++d) tmax[hh] = std::max(tmax[hh], table[(size_t)hh * width + d]);
    if ((rc = enc_alloc(e, &e->rel_max, (size_t)kHeads)) != CSS_OK) return fail(rc);
    if (cudaMemcpy(e->rel_max, tmax.data(), kHeads * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("rel_max upload failed");
      return fail(CSS_ERR_CUDA);
    }
    if ((rc = enc_alloc(e, &e->rel_table, table.size())) != CSS_OK) return fail(rc);
    if (cudaMemcpy(e->rel_table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("rel_table upload failed");
      return fail(CSS_ERR_CUDA);
    }
  }
  float* stage = nullptr;
  if ((rc = enc_alloc(e, &stage, (size_t)kFfn * H)) != CSS_OK) return fail(rc);
  e->layers.resize(cfg->num_layers);
  for (int l = 0; l < cfg->num_layers; ++l) {
    const css_mpnet_layer& s = w->layers[l];
    EncLayer& d = e->layers[l];
    if ((rc = enc_alloc(e, &d.wqkv, 3 * H * H)) != CSS_OK) return fail(rc);
    if ((rc = enc_alloc(e, &d.wo, H * H)) != CSS_OK) return fail(rc);
    if ((rc = enc_alloc(e, &d.w1, (size_t)kFfn * H)) != CSS_OK) return fail(rc);
    if ((rc = enc_alloc(e, &d.w2, (size_t)kFfn * H)) != CSS_OK) return fail(rc);
    if ((rc = enc_alloc(e, &d.bqkv, 3 * H)) != CSS_OK) return fail(rc);
    if ((rc = upload_bf16_into(e, d.wqkv, s.q_w, H * H, stage)) != CSS_OK) return fail(rc);
    if ((rc = upload_bf16_into(e, d.wqkv + H * H, s.k_w, H * H, stage)) != CSS_OK) return fail(rc);
    if ((rc = upload_bf16_into(e, d.wqkv + 2 * H * H, s.v_w, H * H, stage)) != CSS_OK) return fail(rc);
    if ((rc = upload_bf16_into(e, d.wo, s.o_w, H * H, stage)) != CSS_OK) return fail(rc);
    if ((rc = upload_bf16_into(e, d.w1, s.ffn1_w, (size_t)kFfn * H, stage)) != CSS_OK) return fail(rc);
    if ((rc = upload_bf16_into(e, d.w2, s.ffn2_w, (size_t)kFfn * H, stage)) != CSS_OK) return fail(rc);
    if (!s.q_b || !s.k_b || !s.v_b) {
      set_error("layer %d: a q/k/v bias pointer is NULL", l);
      return fail(CSS_ERR_INVALID);
    }
    if (cudaMemcpyAsync(d.bqkv, s.q_b, H * 4, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
        cudaMemcpyAsync(d.bqkv + H, s.k_b, H * 4, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
        cudaMemcpyAsync(d.bqkv + 2 * H, s.v_b, H * 4, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) {
      set_error("bias upload failed");
      return fail(CSS_ERR_CUDA);
    }
    if ((rc = upload_f32(e, &d.bo, s.o_b, H)) != CSS_OK) return fail(rc);
    if ((rc = upload_f32(e, &d.b1, s.ffn1_b, kFfn)) != CSS_OK) return fail(rc);
    if ((rc = upload_f32(e, &d.b2, s.ffn2_b, H)) != CSS_OK) return fail(rc);
    if ((rc = upload_f32(e, &d.ln1_w, s.ln1_w, H)) != CSS_OK) return fail(rc);
    if ((rc = upload_f32(e, &d.ln1_b, s.ln1_b, H)) != CSS_OK) return fail(rc);
    if ((rc = upload_f32(e, &d.ln2_w, s.ln2_w, H)) != CSS_OK) return fail(rc);
    if ((rc = upload_f32(e, &d.ln2_b, s.ln2_b, H)) != CSS_OK) return fail(rc);
    // LayerNorm-folded copies: FFN up-projection with this layer's attention LayerNorm, QKV projection
    // with the previous layer's output LayerNorm
    {
      std::vector<float> wf((size_t)kFfn * H), cc(kFfn), dd(kFfn);
      fold_layernorm(s.ffn1_w, s.ffn1_b, s.ln1_w, s.ln1_b, kFfn, H, wf.data(), cc.data(), dd.data());
      if ((rc = enc_alloc(e, &d.w1_f, (size_t)kFfn * H)) != CSS_OK) return fail(rc);
      if ((rc = upload_bf16_into(e, d.w1_f, wf.data(), (size_t)kFfn * H, stage)) != CSS_OK) return fail(rc);
      if ((rc = upload_f32(e, &d.c1, cc.data(), kFfn)) != CSS_OK) return fail(rc);
      if ((rc = upload_f32(e, &d.d1, dd.data(), kFfn)) != CSS_OK) return fail(rc);
      if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
        set_error("weight upload failed");
        return fail(CSS_ERR_CUDA);
      }
      if (l > 0) {
        const css_mpnet_layer& pl = w->layers[l - 1];
        std::vector<float> cq(3 * H), dq(3 * H);
        if ((rc = enc_alloc(e, &d.wqkv_f, 3 * H * H)) != CSS_OK) return fail(rc);
        const float* ws[3] = {s.q_w, s.k_w, s.v_w};
        const float* bs[3] = {s.q_b, s.k_b, s.v_b};
        for (int j = 0; j < 3; ++j) {
          fold_layernorm(ws[j], bs[j], pl.ln2_w, pl.ln2_b, H, H, wf.data(), cq.data() + j * H, dq.data() + j * H);
          if ((rc = upload_bf16_into(e, d.wqkv_f + (size_t)j * H * H, wf.data(), H * H, stage)) != CSS_OK) return fail(rc);
        }
        if ((rc = upload_f32(e, &d.cqkv, cq.data(), 3 * H)) != CSS_OK) return fail(rc);
        if ((rc = upload_f32(e, &d.dqkv, dq.data(), 3 * H)) != CSS_OK) return fail(rc);
        if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
          set_error("weight upload failed");
          return fail(CSS_ERR_CUDA);
        }
      }
    }
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
      set_error("weight upload failed");
      return fail(CSS_ERR_CUDA);
    }
  }
  // workspace
  const size_t T = (size_t)max_tokens;
  if ((rc = enc_alloc(e, &e->x, T * H)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->x1, T * H)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->qkv, T * 3 * H)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->ctx, T * H)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->h, T * kFfn)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->ln_stats, std::max(T * 3 * kGemmEpiColSplit, (size_t)2 * kQueryMaxRows * kQueryParts))) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->mr1, T)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->mr2, T)) != CSS_OK) return fail(rc);
  cudaMemsetAsync(e->mr1, 0, T * sizeof(float2), e->stream);
  cudaMemsetAsync(e->mr2, 0, T * sizeof(float2), e->stream);
  if ((rc = enc_alloc(e, &e->ids_dev, T)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->cu_dev, (size_t)e->max_seqs + 1)) != CSS_OK) return fail(rc);
  if ((rc = enc_alloc(e, &e->out_dev, (size_t)e->max_seqs * H)) != CSS_OK) return fail(rc);
  // the graph path runs whole token buckets: rows past a query's end must hold finite values
  cudaMemsetAsync(e->x, 0, T * H * 2, e->stream);
  cudaMemsetAsync(e->x1, 0, T * H * 2, e->stream);
  cudaMemsetAsync(e->qkv, 0, T * 3 * H * 2, e->stream);
  cudaMemsetAsync(e->ctx, 0, T * H * 2, e->stream);
  cudaMemsetAsync(e->h, 0, T * kFfn * 2, e->stream);
  cudaMemsetAsync(e->ids_dev, 0, T * 4, e->stream);
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
    set_error("encoder initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(CSS_ERR_CUDA);
  }
  *out = e;
  return CSS_OK;
}

int css_encoder_destroy(css_encoder* e) {
  if (!e) return CSS_OK;
  {
    DeviceGuard g(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto& kv : e->query_graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (void* p : e->owned) cudaFree(p);
    if (e->pinned) cudaFreeHost(e->pinned);
    if (e->stream) cudaStreamDestroy(e->stream);
  }
  delete e;
  return CSS_OK;
}

int css_encoder_dim(const css_encoder* e) { return e ? e->cfg.hidden_size : CSS_ERR_INVALID; }
int64_t css_encoder_max_tokens(const css_encoder* e) { return e ? e->max_tokens : (int64_t)CSS_ERR_INVALID; }
int css_encoder_max_seq_len(const css_encoder* e) { return e ? e->max_seq : CSS_ERR_INVALID; }

int css_encoder_encode_device(css_encoder* e, const int32_t* ids_dev, const int32_t* cu_seqlens_dev,
                              const int32_t* cu_seqlens_host, int32_t n_seq, int normalize, float* out_dev,
                              void* stream) {
  CSS_REQUIRE(e != nullptr, "encoder is NULL");
  CSS_REQUIRE(n_seq >= 0, "n_seq < 0");
  if (n_seq == 0) return CSS_OK;
  CSS_REQUIRE(ids_dev && cu_seqlens_dev && cu_seqlens_host && out_dev, "NULL buffer");
  int max_len = 0;
  CSS_CHECK(validate_cu(e, cu_seqlens_host, n_seq, &max_len));
  const int T = cu_seqlens_host[n_seq];
  CSS_REQUIRE(T <= e->max_tokens, "%d tokens exceed the workspace (%lld)", T, (long long)e->max_tokens);
  std::lock_guard<std::mutex> lk(e->mu);
  DeviceGuard g(e->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  return forward(e, ids_dev, cu_seqlens_dev, n_seq, T, max_len, normalize, out_dev, st);
}

int css_encoder_encode(css_encoder* e, const int32_t* ids_host, const int32_t* cu_seqlens_host, int32_t n_seq,
                       int normalize, float* out_host) {
  CSS_REQUIRE(e != nullptr, "encoder is NULL");
  CSS_REQUIRE(n_seq >= 0, "n_seq < 0");
  if (n_seq == 0) return CSS_OK;
  CSS_REQUIRE(ids_host && cu_seqlens_host && out_host, "NULL buffer");
  int max_len_all = 0;
  CSS_CHECK(validate_cu(e, cu_seqlens_host, n_seq, &max_len_all));
  CSS_CHECK(ensure_device(e->device));
  std::lock_guard<std::mutex> lk(e->mu);
  DeviceGuard g(e->device);
  cudaStream_t st = e->stream;
  const size_t H = kHidden;
  int s0 = 0;
  while (s0 < n_seq) {
    // greedy pass: as many whole sequences as fit max_tokens
    int s1 = s0;
    int max_len = 0;
    const int base = cu_seqlens_host[s0];
    while (s1 < n_seq && cu_seqlens_host[s1 + 1] - base <= e->max_tokens && s1 - s0 < e->max_seqs) {
      max_len = std::max(max_len, cu_seqlens_host[s1 + 1] - cu_seqlens_host[s1]);
      ++s1;
    }
    const int ns = s1 - s0;
    const int T = cu_seqlens_host[s1] - base;
    const size_t ids_bytes = (size_t)T * 4, cu_bytes = (size_t)(ns + 1) * 4, out_bytes = (size_t)ns * H * 4;
    const size_t cu_off = (ids_bytes + 15) / 16 * 16, out_off = (cu_off + cu_bytes + 15) / 16 * 16;
    CSS_CHECK(ensure_pinned(e, out_off + out_bytes));
    unsigned char* pin = reinterpret_cast<unsigned char*>(e->pinned);
    memcpy(pin, ids_host + base, ids_bytes);
    int32_t* cu_pin = reinterpret_cast<int32_t*>(pin + cu_off);
    for (int i = 0; i <= ns; ++i) cu_pin[i] = cu_seqlens_host[s0 + i] - base;
    CSS_CUDA(cudaMemcpyAsync(e->ids_dev, pin, ids_bytes, cudaMemcpyHostToDevice, st));
    CSS_CUDA(cudaMemcpyAsync(e->cu_dev, cu_pin, cu_bytes, cudaMemcpyHostToDevice, st));
    const int bucket = (n_seq == 1) ? query_bucket(T) : 0;
    if (bucket) CSS_CHECK(forward_query_graph(e, bucket, normalize, st));
    else CSS_CHECK(forward(e, e->ids_dev, e->cu_dev, ns, T, max_len, normalize, e->out_dev, st));
    CSS_CUDA(cudaMemcpyAsync(pin + out_off, e->out_dev, out_bytes, cudaMemcpyDeviceToHost, st));
    CSS_CUDA(cudaStreamSynchronize(st));
    memcpy(out_host + (size_t)s0 * H, pin + out_off, out_bytes);
    s0 = s1;
  }
  return CSS_OK;
}

// ---- diagnostics ---------------------------------------------------------------------
namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cudaMalloc(%zu) failed", bytes);
      return CSS_ERR_OOM;
    }
    return CSS_OK;
  }
};
static __global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __bfloat162float(s[i]);
}
int to_bf16_dev(const float* host, size_t n, DevBuf& f32, DevBuf& b16, cudaStream_t st) {
  CSS_CHECK(f32.alloc(n * 4));
  CSS_CHECK(b16.alloc(n * 2));
  CSS_CUDA(cudaMemcpyAsync(f32.p, host, n * 4, cudaMemcpyHostToDevice, st));
  f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)f32.p, (__nv_bfloat16*)b16.p, (int64_t)n);
  CSS_LAUNCHED();
  return CSS_OK;
}
}  // namespace

int css_debug_gemm(const float* A, const float* B, const float* bias, int M, int N, int K, int gelu, int device,
                   float* out) {
  CSS_REQUIRE(A && B && bias && out, "NULL buffer");
  CSS_CHECK(ensure_device(device));
  DeviceGuard g(device);
  cudaStream_t st = nullptr;
  DevBuf a32, a16, b32, b16, biasd, o16, o32;
  CSS_CHECK(to_bf16_dev(A, (size_t)M * K, a32, a16, st));
  CSS_CHECK(to_bf16_dev(B, (size_t)N * K, b32, b16, st));
  CSS_CHECK(biasd.alloc((size_t)N * 4));
  CSS_CUDA(cudaMemcpyAsync(biasd.p, bias, (size_t)N * 4, cudaMemcpyHostToDevice, st));
  CSS_CHECK(o16.alloc((size_t)M * N * 2));
  CSS_CHECK(o32.alloc((size_t)M * N * 4));
  CSS_CUDA(cudaMemsetAsync(o16.p, 0xff, (size_t)M * N * 2, st));
  int rc;
  // gelu bit 0: GELU epilogue; bits 1-2 force a kernel: 2 = single-CTA, 4 = 2-CTA pair
  const int force = gelu & 6;
  const bool two = force == 4 || (force == 0 && gemm::use_2cta());
  if (gelu & 1) {
    EpiBiasBf16<true>::Params p{(__nv_bfloat16*)o16.p, (const float*)biasd.p, N, nullptr, nullptr};
    rc = two ? gemm::launch2<256, EpiBiasBf16<true>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st)
             : gemm::launch<256, EpiBiasBf16<true>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st);
  } else {
    EpiBiasBf16<false>::Params p{(__nv_bfloat16*)o16.p, (const float*)biasd.p, N, nullptr, nullptr};
    rc = two ? gemm::launch2<256, EpiBiasBf16<false>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st)
             : gemm::launch<256, EpiBiasBf16<false>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st);
  }
  CSS_CHECK(rc);
  const int64_t n = (int64_t)M * N;
  bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)o16.p, (float*)o32.p, n);
  CSS_LAUNCHED();
  CSS_CUDA(cudaMemcpyAsync(out, o32.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CSS_CUDA(cudaStreamSynchronize(st));
  return CSS_OK;
}

int css_debug_gemm_resid_ln(const float* A, const float* B, const float* bias, const float* resid, const float* gamma,
                            const float* beta, int M, int K, float eps, int mode, int device, float* out) {
  // mode bit 0: 2-CTA tiles; bit 1: LayerNorm fused into the epilogue (panel order), else epilogue
  // statistics + ln_apply_kernel
  CSS_REQUIRE(A && B && bias && resid && gamma && beta && out, "NULL buffer");
  CSS_CHECK(ensure_device(device));
  DeviceGuard g(device);
  cudaStream_t st = nullptr;
  const int N = kHidden;
  const bool two_cta = (mode & 1) != 0, fused = (mode & 2) != 0;
  DevBuf a32, a16, b32, b16, r32, r16, biasd, gd, bd, o16, o32, statd;
  CSS_CHECK(statd.alloc((size_t)M * 3 * kGemmEpiColSplit * sizeof(float2)));
  CSS_CHECK(to_bf16_dev(A, (size_t)M * K, a32, a16, st));
  CSS_CHECK(to_bf16_dev(B, (size_t)N * K, b32, b16, st));
  CSS_CHECK(to_bf16_dev(resid, (size_t)M * N, r32, r16, st));
  CSS_CHECK(biasd.alloc((size_t)N * 4));
  CSS_CHECK(gd.alloc((size_t)N * 4));
  CSS_CHECK(bd.alloc((size_t)N * 4));
  CSS_CUDA(cudaMemcpyAsync(biasd.p, bias, (size_t)N * 4, cudaMemcpyHostToDevice, st));
  CSS_CUDA(cudaMemcpyAsync(gd.p, gamma, (size_t)N * 4, cudaMemcpyHostToDevice, st));
  CSS_CUDA(cudaMemcpyAsync(bd.p, beta, (size_t)N * 4, cudaMemcpyHostToDevice, st));
  CSS_CHECK(o16.alloc((size_t)M * N * 2));
  CSS_CHECK(o32.alloc((size_t)M * N * 4));
  CSS_CUDA(cudaMemsetAsync(o16.p, 0xff, (size_t)M * N * 2, st));
  int rc;
  if (fused) {
    EpiResidLN<true>::Params p{(__nv_bfloat16*)o16.p, (const float*)biasd.p, (const __nv_bfloat16*)r16.p,
                               (const float*)gd.p, (const float*)bd.p, eps, nullptr, nullptr, nullptr, nullptr};
    rc = two_cta ? gemm::launch2<256, EpiResidLN<true>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st)
                 : gemm::launch<256, EpiResidLN<true>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st);
  } else {
    EpiResidLN<false>::Params p{(__nv_bfloat16*)o16.p, (const float*)biasd.p, (const __nv_bfloat16*)r16.p,
                                (const float*)gd.p, (const float*)bd.p, eps, (float2*)statd.p, nullptr, nullptr, nullptr};
    rc = two_cta ? gemm::launch2<256, EpiResidLN<false>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st)
                 : gemm::launch<256, EpiResidLN<false>>(a16.p, K, b16.p, K, M, N, K, 0, p, sm_count(device), st);
    if (rc == CSS_OK) {
      ln_apply_kernel<<<(unsigned)((M + 8 * kLnApplyRows - 1) / (8 * kLnApplyRows)), 256, 0, st>>>((__nv_bfloat16*)o16.p, (const float2*)statd.p, M,
                                                               (const float*)gd.p, (const float*)bd.p, eps);
      CSS_LAUNCHED();
    }
  }
  CSS_CHECK(rc);
  const int64_t n = (int64_t)M * N;
  bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)o16.p, (float*)o32.p, n);
  CSS_LAUNCHED();
  CSS_CUDA(cudaMemcpyAsync(out, o32.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CSS_CUDA(cudaStreamSynchronize(st));
  return CSS_OK;
}

int css_debug_attention(const float* qkv, const int32_t* cu_seqlens, int n_seq, const float* rel_table, int rel_half,
                        int device, float* ctx) {
  CSS_REQUIRE(qkv && cu_seqlens && rel_table && ctx && n_seq >= 1, "bad arguments");
  CSS_CHECK(ensure_device(device));
  DeviceGuard g(device);
  cudaStream_t st = nullptr;
  const int T = cu_seqlens[n_seq];
  int max_len = 0;
  for (int i = 0; i < n_seq; ++i) max_len = std::max(max_len, cu_seqlens[i + 1] - cu_seqlens[i]);
  CSS_REQUIRE(max_len >= 1 && max_len <= kMaxSeq && rel_half >= max_len - 1, "bad sequence lengths");
  DevBuf q32, q16, cud, reld, c16, c32;
  CSS_CHECK(to_bf16_dev(qkv, (size_t)T * 3 * kHidden, q32, q16, st));
  CSS_CHECK(cud.alloc((size_t)(n_seq + 1) * 4));
  CSS_CUDA(cudaMemcpyAsync(cud.p, cu_seqlens, (size_t)(n_seq + 1) * 4, cudaMemcpyHostToDevice, st));
  const size_t rel_n = (size_t)kHeads * (2 * rel_half + 1);
  CSS_CHECK(reld.alloc(rel_n * 4));
  CSS_CUDA(cudaMemcpyAsync(reld.p, rel_table, rel_n * 4, cudaMemcpyHostToDevice, st));
  CSS_CHECK(c16.alloc((size_t)T * kHidden * 2));
  CSS_CHECK(c32.alloc((size_t)T * kHidden * 4));
  const int Lp = (max_len + 63) & ~63;
  const size_t smem = (size_t)Lp * 256 + (size_t)2 * Lp * sizeof(float);
  CSS_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((max_len + kAttnQRows - 1) / kAttnQRows), kHeads, (unsigned)n_seq);
  const char* tc_env = getenv("CSS_ATTN_TC");
  if ((tc_env ? atoi(tc_env) != 0 : true) && max_len <= kAttnTcMaxLen) {
    std::vector<float> tmax(kHeads, -INFINITY);
    for (int hh = 0; hh < kHeads; ++hh)
      for (int d = 0; d < 2 * rel_half + 1; ++d) tmax[hh] = std::max(tmax[hh], rel_table[(size_t)hh * (2 * rel_half + 1) + d]);
    DevBuf maxd;
    CSS_CHECK(maxd.alloc(kHeads * 4));
    CSS_CUDA(cudaMemcpy(maxd.p, tmax.data(), kHeads * 4, cudaMemcpyHostToDevice));
    // CSS_ATTN_TRACE=<file>: dump CTA 0's clock64 stamps [6 roles][128 items][8 tags] (profiling aid)
    const char* trace_path = getenv("CSS_ATTN_TRACE");
    DevBuf traced;
    const size_t trace_n = 6 * 128 * 8;
    if (trace_path) {
      CSS_CHECK(traced.alloc(trace_n * 8));
      CSS_CUDA(cudaMemset(traced.p, 0, trace_n * 8));
    }
    for (int rep = 0; rep < (trace_path ? 2 : 1); ++rep)
      CSS_CHECK(attention_tc_launch((const __nv_bfloat16*)q16.p, T, (const int32_t*)cud.p, n_seq, max_len,
                                    (const float*)reld.p, (const float*)maxd.p, rel_half, (__nv_bfloat16*)c16.p,
                                    sm_count(device), st, (long long*)traced.p));
    CSS_CUDA(cudaStreamSynchronize(st));
    // CSS_ATTN_TIME=<reps>: CUDA-event time of the kernel alone (profiling aid, printed to stderr)
    if (const char* reps_env = getenv("CSS_ATTN_TIME")) {
      const int reps = std::max(1, atoi(reps_env));
      cudaEvent_t e0, e1;
      CSS_CUDA(cudaEventCreate(&e0));
      CSS_CUDA(cudaEventCreate(&e1));
      CSS_CUDA(cudaEventRecord(e0, st));
      for (int rep = 0; rep < reps; ++rep)
        CSS_CHECK(attention_tc_launch((const __nv_bfloat16*)q16.p, T, (const int32_t*)cud.p, n_seq, max_len,
                                      (const float*)reld.p, (const float*)maxd.p, rel_half, (__nv_bfloat16*)c16.p,
                                      sm_count(device), st, (long long*)traced.p));
      CSS_CUDA(cudaEventRecord(e1, st));
      CSS_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      CSS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      fprintf(stderr, "css_debug_attention: %d sequences, %d tokens, %.2f us per launch (%d launches)\n", n_seq, T,
              1e3f * ms / reps, reps);
      if (trace_path) {   // CTA 0 of the last launch: cycles and nanoseconds from its first to its last instruction
        long long t[4];
        CSS_CUDA(cudaMemcpy(t, (long long*)traced.p + 127 * 8 + 6, 16, cudaMemcpyDeviceToHost));
        CSS_CUDA(cudaMemcpy(t + 2, (long long*)traced.p + (128 + 127) * 8 + 6, 16, cudaMemcpyDeviceToHost));
        fprintf(stderr, "css_debug_attention: CTA 0 of the last launch: %lld cycles in %.2f us (%.0f MHz)\n", t[2] - t[0],
                1e-3 * (double)(t[3] - t[1]), 1e3 * (double)(t[2] - t[0]) / (double)(t[3] - t[1]));
        std::vector<long long> se(2 * 148);
        CSS_CUDA(cudaMemcpy(se.data(), (long long*)traced.p + 5 * 1024 + 600, se.size() * 8, cudaMemcpyDeviceToHost));
        long long s0 = se[0], s1 = se[0], e0 = se[1], e1 = se[1], dmin = se[1] - se[0], dmax = dmin;
        for (int b = 0; b < 148; ++b) {
          s0 = std::min(s0, se[2 * b]); s1 = std::max(s1, se[2 * b]);
          e0 = std::min(e0, se[2 * b + 1]); e1 = std::max(e1, se[2 * b + 1]);
          dmin = std::min(dmin, se[2 * b + 1] - se[2 * b]); dmax = std::max(dmax, se[2 * b + 1] - se[2 * b]);
        }
        for (int b = 0; b < 148; ++b) fprintf(stderr, "%d%c", (int)((se[2 * b + 1] - se[2 * b]) / 1000), b % 37 == 36 ? '\n' : ' ');
        fprintf(stderr, "css_debug_attention: CTA starts spread %.1f us, ends spread %.1f us, CTA lifetime %.1f .. %.1f us, "
                "first start to last end %.1f us\n", 1e-3 * (s1 - s0), 1e-3 * (e1 - e0), 1e-3 * dmin, 1e-3 * dmax,
                1e-3 * (e1 - s0));
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (trace_path) {
      std::vector<long long> host(trace_n);
      CSS_CUDA(cudaMemcpy(host.data(), traced.p, trace_n * 8, cudaMemcpyDeviceToHost));
      if (FILE* f = fopen(trace_path, "wb")) {
        fwrite(host.data(), 8, trace_n, f);
        fclose(f);
      }
    }
  } else {
    attention_kernel<<<grid, kAttnThreads, smem, st>>>((const __nv_bfloat16*)q16.p, (const int32_t*)cud.p,
                                                       (const float*)reld.p, rel_half, (__nv_bfloat16*)c16.p);
    CSS_LAUNCHED();
  }
  const int64_t n = (int64_t)T * kHidden;
  bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)c16.p, (float*)c32.p, n);
  CSS_LAUNCHED();
  CSS_CUDA(cudaMemcpyAsync(ctx, c32.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CSS_CUDA(cudaStreamSynchronize(st));
  return CSS_OK;
}

}  // extern "C"
