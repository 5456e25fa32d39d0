// search_batched.cu -- batched (nq >= CSS_BATCH_MIN_NQ) exact top-k.
// Placeholder body until the tcgen05 score-GEMM lands: routes through the
// streaming scan (still CUDA, still exact), one grid row per query.
#include "index_internal.h"

namespace css {

int batched_search(css_index* h, const float* q_dev, int nq, int k, const uint32_t* mask_dev,
                   int64_t id_offset, float* D_dev, int64_t* I_dev, cudaStream_t st) {
  return scan_search(h, q_dev, nq, k, mask_dev, id_offset, D_dev, I_dev, st);
}

void batched_release(css_index* h) { (void)h; }

}  // namespace css
