/*
 * oracle/flat_ip.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * CPU restatement of the flat exact top-k the reference reaches through
 * faiss-cpu (a third-party dependency, `faiss-cpu>=1.11.0` in the reference's
 * pyproject.toml:9, NOT vendored under /root/reference):
 *
 *   faiss.IndexFlatIP(d).search(q, k)      reference call site src/storage.py:436
 *   faiss.IndexFlatL2(d).search(q, k)      reference call site src/storage.py:258
 *
 * Published algorithm restated here: every query is scored against every row
 * (float32 inner product, or squared L2 distance), and the k best are kept in
 * a binary heap, then sorted best-first.  faiss leaves the order of equal
 * scores unspecified; this oracle fixes it to (score desc, id asc) -- the
 * order the CUDA path is required to produce -- so comparisons are exact.
 * Unfilled slots: id -1, score -FLT_MAX (IP) / FLT_MAX (L2), as faiss does.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  float key; /* larger is better (IP: score, L2: -distance) */
  int64_t id;
} ent_t;

/* a is better than b */
static inline int better(float ka, int64_t ia, float kb, int64_t ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

/* min-heap on `better` (root = worst of the kept k) */
static void heap_sift_down(ent_t* h, int n, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, w = i;
    if (l < n && better(h[w].key, h[w].id, h[l].key, h[l].id)) w = l;
    if (r < n && better(h[w].key, h[w].id, h[r].key, h[r].id)) w = r;
    if (w == i) return;
    ent_t t = h[i];
    h[i] = h[w];
    h[w] = t;
    i = w;
  }
}

static inline void heap_offer(ent_t* h, int* n, int k, float key, int64_t id) {
  if (*n < k) {
    /* sift up */
    int i = (*n)++;
    h[i].key = key;
    h[i].id = id;
    while (i > 0) {
      int p = (i - 1) / 2;
      if (better(h[p].key, h[p].id, h[i].key, h[i].id)) {
        ent_t t = h[i];
        h[i] = h[p];
        h[p] = t;
        i = p;
      } else
        break;
    }
  } else if (better(key, id, h[0].key, h[0].id)) {
    h[0].key = key;
    h[0].id = id;
    heap_sift_down(h, k, 0);
  }
}

static int cmp_best_first(const void* a, const void* b) {
  const ent_t* x = (const ent_t*)a;
  const ent_t* y = (const ent_t*)b;
  if (better(x->key, x->id, y->key, y->id)) return -1;
  if (better(y->key, y->id, x->key, x->id)) return 1;
  return 0;
}

static inline float dot_f32(const float* a, const float* b, int d) {
  float acc[16];
  int j, t;
  for (t = 0; t < 16; ++t) acc[t] = 0.f;
  for (j = 0; j + 16 <= d; j += 16)
    for (t = 0; t < 16; ++t) acc[t] += a[j + t] * b[j + t];
  float s = 0.f;
  for (t = 0; t < 16; ++t) s += acc[t];
  for (; j < d; ++j) s += a[j] * b[j];
  return s;
}

static inline float l2sq_f32(const float* a, const float* b, int d) {
  float acc[16];
  int j, t;
  for (t = 0; t < 16; ++t) acc[t] = 0.f;
  for (j = 0; j + 16 <= d; j += 16)
    for (t = 0; t < 16; ++t) {
      float v = a[j + t] - b[j + t];
      acc[t] += v * v;
    }
  float s = 0.f;
  for (t = 0; t < 16; ++t) s += acc[t];
  for (; j < d; ++j) {
    float v = a[j] - b[j];
    s += v * v;
  }
  return s;
}

/*
 * x: [n, d] row-major; q: [nq, d]; mask: optional bitmask (bit i%32 of word
 * i/32 set = row i may be returned), NULL = all rows.  metric: 0 IP, 1 L2.
 * D: [nq, k], I: [nq, k].  nthreads <= 0: all cores.
 * Rows are split across threads (the batch-1 case is the one that matters:
 * faiss itself runs nq == 1 on one thread; giving the baseline every core is
 * the generous reading).
 */
int oracle_flat_search(const float* x, int64_t n, int d, const float* q, int nq, int k, int metric,
                       const uint32_t* mask, float* D, int64_t* I, int nthreads) {
  if (k <= 0 || nq < 0 || d <= 0 || n < 0) return -1;
#ifdef _OPENMP
  int T = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  int T = 1;
#endif
  if (T < 1) T = 1;
  ent_t* heaps = (ent_t*)malloc(sizeof(ent_t) * (size_t)T * (size_t)k);
  int* sizes = (int*)malloc(sizeof(int) * (size_t)T);
  ent_t* all = (ent_t*)malloc(sizeof(ent_t) * (size_t)T * (size_t)k);
  if (!heaps || !sizes || !all) {
    free(heaps);
    free(sizes);
    free(all);
    return -2;
  }
  for (int qi = 0; qi < nq; ++qi) {
    const float* qv = q + (size_t)qi * d;
    for (int t = 0; t < T; ++t) sizes[t] = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(T)
#endif
    {
#ifdef _OPENMP
      int t = omp_get_thread_num();
#else
      int t = 0;
#endif
      int64_t r0 = n * t / T, r1 = n * (t + 1) / T;
      ent_t* h = heaps + (size_t)t * k;
      int hs = 0;
      for (int64_t r = r0; r < r1; ++r) {
        if (mask && !((mask[r >> 5] >> (r & 31)) & 1u)) continue;
        float key = metric == 0 ? dot_f32(x + (size_t)r * d, qv, d) : -l2sq_f32(x + (size_t)r * d, qv, d);
        heap_offer(h, &hs, k, key, r);
      }
      sizes[t] = hs;
    }
    int m = 0;
    for (int t = 0; t < T; ++t) {
      memcpy(all + m, heaps + (size_t)t * k, sizeof(ent_t) * (size_t)sizes[t]);
      m += sizes[t];
    }
    qsort(all, (size_t)m, sizeof(ent_t), cmp_best_first);
    for (int j = 0; j < k; ++j) {
      if (j < m) {
        D[(size_t)qi * k + j] = metric == 0 ? all[j].key : -all[j].key;
        I[(size_t)qi * k + j] = all[j].id;
      } else {
        D[(size_t)qi * k + j] = metric == 0 ? -FLT_MAX : FLT_MAX;
        I[(size_t)qi * k + j] = -1;
      }
    }
  }
  free(heaps);
  free(sizes);
  free(all);
  return 0;
}

/* Row normalisation of src/storage.py:347-350: x / (||x||_2 + 1e-8), float32. */
void oracle_normalize_rows(float* x, int64_t n, int d) {
  for (int64_t r = 0; r < n; ++r) {
    float* row = x + (size_t)r * d;
    float ss = 0.f;
    for (int j = 0; j < d; ++j) ss += row[j] * row[j];
    float den = sqrtf(ss) + 1e-8f;
    for (int j = 0; j < d; ++j) row[j] = row[j] / den;
  }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
