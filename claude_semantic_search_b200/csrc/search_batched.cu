// search_batched.cu -- S3: exact top-k for query batches (nq >= CSS_BATCH_MIN_NQ) on the
// tensor cores.  Replaces faiss IndexFlatIP.search's sgemm path (src/storage.py:436 with a
// batch of queries) without materialising the nq x N score matrix.
//
//   1. queries -> bf16 A operand; the corpus' bf16 shadow copy (xb) is the B operand.
//   2. seed pass over the first 16K rows: the GEMM epilogue keeps only the maximum of every
//      32-row chunk; the k-th largest of these (k distinct rows) is a valid lower bound of the
//      k-th best score and becomes the first threshold.  The lists are then cleared.
//   3. the corpus (from row 0 again) is swept in passes of growing size (128K, 1M, 8M ... rows).
//      Each pass is one tcgen05 GEMM (gemm_tc.cuh) whose epilogue compares every score with
//      the per-query threshold and appends (score, row) to the query's candidate list; after
//      each pass a small kernel sorts every list, sets the next threshold to
//      (k-th best bf16 score) - 2*eps_q and drops entries below it.
//   4. the final kernel re-scores the surviving candidates in fp32 with the arithmetic of the
//      streaming scan (bit-identical scores to the batch-1 path) and writes the top-k.
//
// Exactness.  With xb = bf16(x), qb = bf16(q):  x.q - xb.qb = (x - xb).qb + x.(q - qb), so
//   |x.q - xb.qb| <= max_err ||qb|| + max_norm ||q - qb||  (+ fp32 summation, d 1.3e-7 max_norm ||q||) =: eps_q
// with max_err the largest ||x - bf16(x)|| of any stored row (tracked exactly at add time; the worst
// case of round-to-nearest with bf16's 8-bit significand is 2^-8 ||x||) and ||q - qb|| computed per
// query.  Every row of the true top-k has a bf16 score >= (true k-th bf16 score) - 2 eps_q, which
// is never below the running threshold: no true neighbour is ever filtered.  A candidate
// list that would exceed its capacity flags the query, and flagged queries are re-run by
// the exact streaming scan (never truncated silently).
#include "index_internal.h"
#include "gemm_tc.cuh"

#include <algorithm>
#include <cstdlib>

namespace css {
namespace {

constexpr int kCap = 4096;          // candidate slots per query
constexpr int kQChunk = 1024;       // queries per sweep
constexpr int kSeedRows = 16384;    // pass 0 ("seed"): only the maximum of every 32-row chunk is kept
constexpr int kFirstPassRows = 131072;
constexpr int kPassGrowth = 8;
constexpr int kSelThreads = 512;
constexpr int kFallbackSlices = 8;

struct Cand {
  float score;
  int row;
};

struct BatchedState {
  int max_nq = 0;
  __nv_bfloat16* qb = nullptr;   // [max_nq, dim] bf16 queries
  float* qnorm = nullptr;        // [3][nq]: ||q||, ||bf16(q)||, ||q - bf16(q)|| (stride = the chunk's nq)
  float* thr = nullptr;          // [max_nq] threshold of the current pass
  unsigned* count = nullptr;     // [max_nq]
  Cand* cand = nullptr;          // [max_nq][kCap]
  int* ovf_flag = nullptr;       // [max_nq] 1 = list overflowed
  int* ovf_list = nullptr;       // [max_nq] compacted overflowed query indices
  int* ovf_count = nullptr;      // [1]
};

// ---- 1. query preparation -------------------------------------------------------------
__global__ void prep_queries_kernel(const float* __restrict__ q, int nq, int d, __nv_bfloat16* __restrict__ qb,
                                    float* __restrict__ qnorm, float* __restrict__ thr, unsigned* __restrict__ count,
                                    int* __restrict__ ovf_flag, int* __restrict__ ovf_count) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) *ovf_count = 0;
  if (qi >= nq) return;
  const float* src = q + (size_t)qi * d;
  float ss = 0.f, sb = 0.f, se = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v = src[j];
    ss = fmaf(v, v, ss);
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    qb[(size_t)qi * d + j] = b;
    const float bv = __bfloat162float(b), e = v - bv;
    sb = fmaf(bv, bv, sb);
    se = fmaf(e, e, se);
  }
  ss = warp_sum(ss);
  sb = warp_sum(sb);
  se = warp_sum(se);
  if (lane == 0) {
    qnorm[qi] = sqrtf(ss);
    qnorm[nq + qi] = sqrtf(sb) * 1.0001f;   // ||bf16(q)||
    qnorm[2 * nq + qi] = sqrtf(se) * 1.0001f;   // ||q - bf16(q)||
    thr[qi] = -INFINITY;
    count[qi] = 0;
    ovf_flag[qi] = 0;
  }
}

// ---- 2. GEMM epilogue: threshold filter -------------------------------------------------
struct EpiSearch {
  static constexpr bool kMasksColumns = true;
  static constexpr bool kPanel = false;
  static constexpr bool kResidPrefetch = false;
  static constexpr int kStageBytes = 0;
  struct Params {
    const float* thr;       // [nq]
    unsigned* count;        // [nq]
    Cand* cand;             // [nq][kCap]
    int* ovf_flag;          // [nq]
    const uint32_t* mask;   // nullable row bitmask (global rows)
    int row0;               // global row of column 0 of this pass (multiple of 32)
    int n_rows;             // rows in this pass
    int seed;               // 1: keep only the maximum of each 32-row chunk (threshold seeding)
  };
  const Params& p;
  __device__ EpiSearch(const Params& p_, int, uint8_t*) : p(p_) {}
  __device__ __forceinline__ void chunk(int, int m_warp, int lane, int M, int n0, const uint32_t (&v)[32]) {
    const int m = m_warp + lane;
    if (m >= M) return;
    if (p.seed) {
      // chunk-aligned mask word (row0 and n0 are multiples of 32) restricted to real rows
      uint32_t mw = p.mask ? __ldg(p.mask + ((p.row0 + n0) >> 5)) : 0xffffffffu;
      const int left = p.n_rows - n0;
      if (left < 32) mw &= left > 0 ? ((1u << left) - 1u) : 0u;
      float best = -INFINITY;
      int bj = -1;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = __uint_as_float(v[j]);
        if (((mw >> j) & 1u) && s > best) {
          best = s;
          bj = j;
        }
      }
      if (bj >= 0) offer(m, n0 + bj, best);
      return;
    }
    const float t = __ldg(p.thr + m);
    // two-level reject: maxima of the four 8-column groups, then their maximum.  A warp takes
    // the slow path when ANY of its 32 query rows has a hit, so the slow path itself only
    // descends into groups that hold one (round-1 profile: flat 32-way checks made the first
    // sweep pass epilogue-bound).
    float gm[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float a = __uint_as_float(v[g * 8]);
#pragma unroll
      for (int j = 1; j < 8; ++j) a = fmaxf(a, __uint_as_float(v[g * 8 + j]));
      gm[g] = a;
    }
    const float mx = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
    if (!(mx >= t)) return;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (gm[g] >= t) {
        // one out-of-line call per 8-column group with a hit: ONE slot reservation for all of its hits.  (A clustered
        // corpus in session order puts a query's whole cluster -- hundreds of rows above the running threshold -- into
        // adjacent columns; one call + one atomic per hit made those tiles epilogue-bound: 2.06 vs 1.20 ms per batch.)
        unsigned hm = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) hm |= (__uint_as_float(v[g * 8 + j]) >= t) ? (1u << j) : 0u;
        append_group(p.count, p.cand, p.ovf_flag, p.mask, p.row0, p.n_rows, m, n0 + g * 8, hm, v[g * 8], v[g * 8 + 1],
                     v[g * 8 + 2], v[g * 8 + 3], v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
      }
    }
  }
  __device__ __noinline__ static void append_group(unsigned* count, Cand* cand, int* ovf_flag, const uint32_t* mask,
                                                  int row0, int n_rows, int m, int r0, unsigned hm, uint32_t s0,
                                                  uint32_t s1, uint32_t s2, uint32_t s3, uint32_t s4, uint32_t s5,
                                                  uint32_t s6, uint32_t s7) {
    const int left = n_rows - r0;                       // columns of this group inside the pass
    if (left < 8) hm &= left > 0 ? ((1u << left) - 1u) : 0u;
    const int g = row0 + r0;                            // multiple of 8
    if (mask) hm &= (__ldg(mask + (g >> 5)) >> (g & 31)) & 0xffu;
    if (!hm) return;
    const unsigned n = __popc(hm);
    unsigned pos = atomicAdd(count + m, n);
    if (pos + n > (unsigned)kCap) {
      ovf_flag[m] = 1;
      return;
    }
    Cand* out = cand + (size_t)m * kCap;
    const uint32_t sv[8] = {s0, s1, s2, s3, s4, s5, s6, s7};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((hm >> j) & 1u) {
        Cand c;
        c.score = __uint_as_float(sv[j]);
        c.row = g + j;
        out[pos++] = c;
      }
    }
  }
  // Rare path, kept out of line (and free of `this`, so the functor stays in registers): the
  // hot loop above stays compact; inlining 32 copies of this body slowed the whole epilogue.
  __device__ __forceinline__ void offer(int m, int r, float s) {
    append_hit(p.count, p.cand, p.ovf_flag, p.mask, p.row0, p.n_rows, m, r, s);
  }
  __device__ __noinline__ static void append_hit(unsigned* count, Cand* cand, int* ovf_flag, const uint32_t* mask,
                                                 int row0, int n_rows, int m, int r, float s) {
    if (r >= n_rows) return;
    const int g = row0 + r;
    if (mask && !((__ldg(mask + (g >> 5)) >> (g & 31)) & 1u)) return;
    const unsigned pos = atomicAdd(count + m, 1u);
    if (pos < (unsigned)kCap) {
      Cand c;
      c.score = s;
      c.row = g;
      cand[(size_t)m * kCap + pos] = c;
    } else {
      ovf_flag[m] = 1;
    }
  }
  __device__ __forceinline__ void prefetch(int, int, int, int, int) {}
  __device__ __forceinline__ void prefetch_none() {}
  __device__ __forceinline__ void tile_begin(int, int, int) {}
  __device__ __forceinline__ void tile_end(int, int, int, int) {}
  __device__ __forceinline__ void finish() {}
};

__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
  return (a.score > b.score) || (a.score == b.score && a.row < b.row);
}

__device__ __forceinline__ void sort_cands(Cand* s, int n_sort, int tid) {
  for (int size = 2; size <= n_sort; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (n_sort >> 1); t += kSelThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const Cand a = s[lo], b = s[hi];
        if (cand_better(a, b) != desc) {
          s[lo] = b;
          s[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// ---- 3./4. per-query list maintenance -------------------------------------------------
// One CTA per query.  kFinal = false: sort, set thr = kth - 2 eps, compact.
// kFinal = true: additionally re-score the survivors in fp32 and emit D/I.
struct SelectParams {
  const float* x;       // fp32 corpus
  const float* q;       // fp32 queries [nq, d]
  int d;
  int k;
  int nq;               // stride of the qnorm planes
  const float* max_norm;
  const float* max_err;
  const float* qnorm;
  float* thr;
  unsigned* count;
  Cand* cand;
  int* ovf_flag;
  int* ovf_list;
  int* ovf_count;
  IdMap idmap;
  float* D;
  int64_t* I;
  int seed;             // 1: this call follows the seed pass -> only the threshold survives
};

template <bool kFinal>
__global__ void __launch_bounds__(kSelThreads) select_kernel(SelectParams p) {
  __shared__ Cand s[kCap];
  const int qi = blockIdx.x;
  const int tid = threadIdx.x;
  unsigned cnt = p.count[qi];
  const bool overflow = cnt > (unsigned)kCap || p.ovf_flag[qi] != 0;
  if (overflow) {
    // the exact scan re-runs this query; nothing here may be trusted
    if (kFinal && tid == 0) {
      const int slot = atomicAdd(p.ovf_count, 1);
      p.ovf_list[slot] = qi;
    }
    if (!kFinal && tid == 0) p.ovf_flag[qi] = 1;
    return;
  }
  int n_sort = 32;
  while (n_sort < (int)cnt) n_sort <<= 1;
  Cand* list = p.cand + (size_t)qi * kCap;
  for (int i = tid; i < n_sort; i += kSelThreads) {
    Cand c;
    if (i < (int)cnt) {
      c = list[i];
    } else {
      c.score = -INFINITY;
      c.row = INT_MAX;
    }
    s[i] = c;
  }
  sort_cands(s, n_sort, tid);
  const float mn = *p.max_norm;
  // last term: fp32 accumulation of d products on the tensor cores (truncating adds: <= d 2^-23 sum|x_i q_i|) and in the re-score
  const float eps = 1.001f * (*p.max_err) * p.qnorm[p.nq + qi] + mn * p.qnorm[2 * p.nq + qi] +
                    (float)p.d * 1.3e-7f * mn * p.qnorm[qi];
  float thr = -INFINITY;
  if ((int)cnt >= p.k) thr = s[p.k - 1].score - 2.f * eps;
  // survivors: sorted prefix with score >= thr
  __shared__ int s_keep;
  if (tid == 0) s_keep = (int)cnt;
  __syncthreads();
  for (int i = tid; i < (int)cnt; i += kSelThreads) {
    const bool here = s[i].score >= thr;
    const bool next = (i + 1 < (int)cnt) ? (s[i + 1].score >= thr) : false;
    if (here && !next) s_keep = i + 1;
    if (i == 0 && !here) s_keep = 0;
  }
  __syncthreads();
  const int keep = s_keep;
  if (!kFinal) {
    if (p.seed) {
      // the seed candidates are chunk maxima only: the sweep restarts at row 0 and finds them again
      if (tid == 0) {
        p.count[qi] = 0;
        p.thr[qi] = thr;
      }
      return;
    }
    for (int i = tid; i < keep; i += kSelThreads) list[i] = s[i];
    if (tid == 0) {
      p.count[qi] = (unsigned)keep;
      p.thr[qi] = thr;
    }
    return;
  }
  // ---- final: exact fp32 re-score of the survivors (same arithmetic as scan_topk_kernel) ----
  const int lane = tid & 31, warp = tid >> 5;
  const float* q = p.q + (size_t)qi * p.d;
  for (int i = warp; i < keep; i += kSelThreads / 32) {
    const float* row = p.x + (size_t)s[i].row * p.d;
    float a = 0.f;
    if (p.d == 768) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float4 xv = ld_stream_f4(reinterpret_cast<const float4*>(row) + j * 32 + lane);
        const float4 qv = __ldg(reinterpret_cast<const float4*>(q) + j * 32 + lane);
        a = fmaf(xv.x, qv.x, a);
        a = fmaf(xv.y, qv.y, a);
        a = fmaf(xv.z, qv.z, a);
        a = fmaf(xv.w, qv.w, a);
      }
    } else {
      for (int j = lane; j < p.d; j += 32) a = fmaf(__ldg(row + j), q[j], a);
    }
    a = warp_sum(a);
    __syncwarp();
    if (lane == 0) s[i].score = a;
  }
  __syncthreads();
  int n2 = 32;
  while (n2 < keep) n2 <<= 1;
  for (int i = keep + tid; i < n2; i += kSelThreads) {
    s[i].score = -INFINITY;
    s[i].row = INT_MAX;
  }
  sort_cands(s, n2, tid);
  for (int i = tid; i < p.k; i += kSelThreads) {
    const bool filled = i < keep && s[i].row != INT_MAX;
    p.D[(size_t)qi * p.k + i] = filled ? s[i].score : -FLT_MAX;
    p.I[(size_t)qi * p.k + i] = filled ? (int64_t)map_id(p.idmap, s[i].row) : (int64_t)-1;
  }
}

// Tuning knobs (rows): CSS_BATCH_SEED_ROWS, CSS_BATCH_FIRST_ROWS override the defaults.
int64_t env_rows(const char* name, int64_t dflt) {
  const char* v = getenv(name);
  if (!v) return dflt;
  const long long x = atoll(v);
  return x >= 2048 ? (int64_t)(x / 256 * 256) : dflt;
}

int ensure_state(css_index* h, css_scan_scratch* sc, int nq) {
  BatchedState* st = reinterpret_cast<BatchedState*>(sc->batched);
  if (!st) {
    st = new (std::nothrow) BatchedState();
    if (!st) {
      set_error("out of host memory");
      return CSS_ERR_OOM;
    }
    sc->batched = st;
  }
  if (nq <= st->max_nq) return CSS_OK;
  cudaFree(st->qb); cudaFree(st->qnorm); cudaFree(st->thr); cudaFree(st->count); cudaFree(st->cand);
  cudaFree(st->ovf_flag); cudaFree(st->ovf_list); cudaFree(st->ovf_count);
  *st = BatchedState();
  const size_t n = (size_t)nq;
  auto A = [&](void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
      return (int)CSS_ERR_OOM;
    }
    return (int)CSS_OK;
  };
  CSS_CHECK(A((void**)&st->qb, n * h->dim * 2));
  CSS_CHECK(A((void**)&st->qnorm, n * 12));
  CSS_CHECK(A((void**)&st->thr, n * 4));
  CSS_CHECK(A((void**)&st->count, n * 4));
  CSS_CHECK(A((void**)&st->cand, n * kCap * sizeof(Cand)));
  CSS_CHECK(A((void**)&st->ovf_flag, n * 4));
  CSS_CHECK(A((void**)&st->ovf_list, n * 4));
  CSS_CHECK(A((void**)&st->ovf_count, 4));
  st->max_nq = nq;
  return CSS_OK;
}

template <int KPL>
int launch_fallback(css_index* h, const ScanParams& p, cudaStream_t st) {
  const bool d768 = (h->dim == 768);
  size_t smem = sizeof(KeyId) * kMergeCap + (d768 ? 0 : (size_t)h->dim * 4);
  dim3 grid((unsigned)h->scan_blocks, kFallbackSlices);
  if (d768) {
    auto kern = scan_topk_kernel<KPL, CSS_METRIC_INNER_PRODUCT, true>;
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kScanThreads, smem, st>>>(p);
  } else {
    auto kern = scan_topk_kernel<KPL, CSS_METRIC_INNER_PRODUCT, false>;
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kScanThreads, smem, st>>>(p);
  }
  CSS_LAUNCHED();
  return CSS_OK;
}

}  // namespace

int batched_search(css_index* h, css_scan_scratch* sc, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                   const IdMap& idmap, float* D_dev, int64_t* I_dev, cudaStream_t stream) {
  CSS_REQUIRE(k >= 1 && k <= CSS_MAX_K, "k=%d outside [1, %d]", k, CSS_MAX_K);
  CSS_REQUIRE(h->dim % gemm::BK == 0, "batched search needs dim %% 64 == 0");
  CSS_REQUIRE(sc->max_nq >= std::min(nq, kQChunk), "scratch too small");
  CSS_CHECK(ensure_state(h, sc, std::min(nq, kQChunk)));
  BatchedState* st = reinterpret_cast<BatchedState*>(sc->batched);
  const int d = h->dim;
  const int64_t N = h->ntotal;

  for (int q0 = 0; q0 < nq; q0 += kQChunk) {
    const int nqc = std::min(kQChunk, nq - q0);
    const float* qc = q_dev + (size_t)q0 * d;
    prep_queries_kernel<<<(unsigned)((nqc + 7) / 8), 256, 0, stream>>>(qc, nqc, d, st->qb, st->qnorm, st->thr,
                                                                       st->count, st->ovf_flag, st->ovf_count);
    CSS_LAUNCHED();
    SelectParams sp;
    sp.x = h->x;
    sp.q = qc;
    sp.d = d;
    sp.k = k;
    sp.nq = nqc;
    sp.max_norm = h->max_norm_dev;
    sp.max_err = h->max_err_dev;
    sp.qnorm = st->qnorm;
    sp.thr = st->thr;
    sp.count = st->count;
    sp.cand = st->cand;
    sp.ovf_flag = st->ovf_flag;
    sp.ovf_list = st->ovf_list;
    sp.ovf_count = st->ovf_count;
    sp.idmap = idmap;
    sp.D = D_dev + (size_t)q0 * k;
    sp.I = I_dev + (size_t)q0 * k;
    sp.seed = 0;

    EpiSearch::Params ep;
    ep.thr = st->thr;
    ep.count = st->count;
    ep.cand = st->cand;
    ep.ovf_flag = st->ovf_flag;
    ep.mask = mask_dev;
    // ---- seed pass: chunk maxima of the first rows -> first threshold ----
    {
      static const int64_t seed_dflt = env_rows("CSS_BATCH_SEED_ROWS", kSeedRows);
      const int64_t seed_rows = std::min<int64_t>(N, std::max<int64_t>(seed_dflt, (int64_t)64 * k));
      ep.row0 = 0;
      ep.n_rows = (int)seed_rows;
      ep.seed = 1;
      CSS_CHECK((gemm::run<256, EpiSearch>(st->qb, d, h->xb, d, nqc, (int)seed_rows, d, /*m_fastest=*/1, ep, h->n_sm,
                                          stream)));
      sp.seed = 1;
      select_kernel<false><<<(unsigned)nqc, kSelThreads, 0, stream>>>(sp);
      CSS_LAUNCHED();
      sp.seed = 0;
      ep.seed = 0;
    }
    // ---- sweep ----
    int64_t r0 = 0;
    static const int64_t first_dflt = env_rows("CSS_BATCH_FIRST_ROWS", kFirstPassRows);
    int64_t pass_rows = first_dflt;
    while (r0 < N) {
      int64_t nr = std::min<int64_t>(pass_rows, N - r0);
      // fold a short tail into this pass
      if (N - (r0 + nr) < nr / 2) nr = N - r0;
      ep.row0 = (int)r0;
      ep.n_rows = (int)nr;
      CSS_CHECK((gemm::run<256, EpiSearch>(st->qb, d, h->xb + (size_t)r0 * d, d, nqc, (int)nr, d,
                                          /*m_fastest=*/1, ep, h->n_sm, stream)));
      r0 += nr;
      if (r0 < N) {
        select_kernel<false><<<(unsigned)nqc, kSelThreads, 0, stream>>>(sp);
        CSS_LAUNCHED();
      }
      pass_rows *= kPassGrowth;
    }
    select_kernel<true><<<(unsigned)nqc, kSelThreads, 0, stream>>>(sp);
    CSS_LAUNCHED();

    // overflowed queries (if any): exact streaming scan, driven by the device-side list
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.x = h->x;
    p.n = N;
    p.d = d;
    p.q = qc;
    p.mask = mask_dev;
    p.k = k;
    p.k_out = k;
    p.part = sc->part;
    p.ticket = sc->ticket;
    p.cursor = sc->ticket + sc->max_nq;
    p.idmap = idmap;
    p.D = sp.D;
    p.I = sp.I;
    p.qlist = st->ovf_list;
    p.qcount = st->ovf_count;
    if (k <= 32) CSS_CHECK(launch_fallback<1>(h, p, stream));
    else if (k <= 64) CSS_CHECK(launch_fallback<2>(h, p, stream));
    else CSS_CHECK(launch_fallback<4>(h, p, stream));
  }
  return CSS_OK;
}

void batched_release(css_scan_scratch* sc) {
  BatchedState* st = reinterpret_cast<BatchedState*>(sc->batched);
  if (!st) return;
  cudaFree(st->qb); cudaFree(st->qnorm); cudaFree(st->thr); cudaFree(st->count); cudaFree(st->cand);
  cudaFree(st->ovf_flag); cudaFree(st->ovf_list); cudaFree(st->ovf_count);
  delete st;
  sc->batched = nullptr;
}

}  // namespace css
