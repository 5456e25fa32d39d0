// gemm_tc.cuh -- persistent, warp-specialised tcgen05 GEMM:  C[M,N] = A[M,K] * B[N,K]^T
//
//   A, B : bf16, K-major (row-major [rows, K]); tiles arrive by TMA with the 128-byte
//          swizzle into a kStages-deep shared-memory ring
//   C    : fp32 accumulators in TMEM (2 x BN columns, double-buffered so the epilogue of
//          tile i overlaps the main loop of tile i+1)
//   roles: warp 0 = TMA producer (1 thread), warp 1 = MMA issuer (1 thread) + TMEM owner,
//          warps 2..9 = epilogue (TMEM -> registers -> Epi functor -> global)
//
// One CTA per SM, tiles of 128 x BN visited round-robin.  The epilogue is a template
// parameter: the encoder instantiates bias / bias+GELU / bias+residual writers, the
// batched search instantiates a per-query threshold filter (search_batched.cu).
#pragma once
#include "tc_common.cuh"

namespace css {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;  // 320

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = 2 * BN;  // 512 (BN=256) or 256 (BN=128)
};

struct Shape {
  int M, N, K;
  int m_fastest;  // 1: consecutive tiles share the B (N) block; 0: share the A (M) block
};

// Epi concept:
//   struct Epi { struct Params {...};
//     __device__ Epi(const Params&, int epi_thread /*0..255*/);
//     __device__ void chunk(int m /*global row*/, bool row_ok, int n0 /*global col of v[0]*/, const uint32_t (&v)[32]);
//     __device__ void tile_end(int m_blk, int n_blk);
//     __device__ void finish(); };
template <int BN, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               Shape shape, typename Epi::Params ep) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty = bars + C::kStages;         // [kStages]  MMA -> TMA
  uint64_t* tfull = bars + 2 * C::kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty = tfull + 2;                // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
    for (int s = 0; s < C::kStages; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tfull + s, 1);
      tc::mbar_init(tempty + s, kEpiWarps);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, C::kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (shape.M + BM - 1) / BM;
  const int num_n = (shape.N + BN - 1) / BN;  // a ragged last block is zero-filled by TMA; Epi masks it
  const int num_tiles = num_m * num_n;
  const int num_kb = shape.K / BK;

  auto tile_coords = [&](int t, int& mb, int& nb) {
    if (shape.m_fastest) {
      mb = t % num_m;
      nb = t / num_m;
    } else {
      nb = t % num_n;
      mb = t / num_n;
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int mb, nb;
        tile_coords(t, mb, nb);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty + stage, phase ^ 1);
          tc::mbar_expect_tx(full + stage, C::kStageBytes);
          tc::tma_load_2d(sA + stage * C::kABytes, &tmap_a, full + stage, kb * BK, mb * BM, tc::kEvictNormal);
          tc::tma_load_2d(sB + stage * C::kBBytes, &tmap_b, full + stage, kb * BK, nb * BN, tc::kEvictNormal);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        tc::mbar_wait(tempty + as, aphase ^ 1);  // epilogue has drained this accumulator
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full + stage, phase);
          tc::tc_fence_after();
          const uint64_t da = tc::make_kmajor_sw128_desc(tc::smem_u32(sA + stage * C::kABytes));
          const uint64_t db = tc::make_kmajor_sw128_desc(tc::smem_u32(sB + stage * C::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_bf16(tmem_d, da + k * tc::kDescKStep, db + k * tc::kDescKStep, idesc, (kb | k) != 0);
          tc::umma_commit(empty + stage);  // smem slot reusable once these MMAs retire
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc::umma_commit(tfull + as);  // accumulator complete
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;          // 0..7
    const int quarter = warp & 3;     // TMEM lane quarter this warp may read
    const int half = ew >> 2;         // column half of the tile
    const int row_in_tile = quarter * 32 + lane;
    Epi epi(ep, ew * 32 + lane);
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int mb, nb;
      tile_coords(t, mb, nb);
      tc::mbar_wait(tfull + as, aphase);
      tc::tc_fence_after();
      const int m = mb * BM + row_in_tile;
      const bool row_ok = m < shape.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(as * BN + half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(taddr + c, v);
        tc::tmem_ld_wait();
        epi.chunk(m, row_ok, nb * BN + half * (BN / 2) + c, v);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty + as);
      epi.tile_end(mb, nb);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    epi.finish();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, C::kTmemCols);
}

// Host launcher.  A: [M, K] row-major bf16 (ld = lda elements), B: [N, K] row-major bf16.
template <int BN, class Epi>
int launch(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int m_fastest,
           const typename Epi::Params& ep, int n_sm, cudaStream_t st) {
  using C = Cfg<BN>;
  CSS_REQUIRE(M >= 1 && N >= 1 && (N % BN == 0 || Epi::kMasksColumns) && K % BK == 0 && K >= BK,
              "gemm shape M=%d N=%d K=%d unsupported (BN=%d)", M, N, K, BN);
  CUtensorMap ta, tb;
  CSS_CHECK(encode_tmap_bf16_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK));
  CSS_CHECK(encode_tmap_bf16_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, BK));
  auto kern = gemm_tc_kernel<BN, Epi>;
  static std::atomic<uint64_t> attr_set{0};  // per instantiation, one bit per device
  int dev = 0;
  CSS_CUDA(cudaGetDevice(&dev));
  if (!(attr_set.load(std::memory_order_relaxed) >> (dev & 63) & 1)) {
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set.fetch_or(uint64_t(1) << (dev & 63), std::memory_order_relaxed);
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < n_sm ? tiles : n_sm;
  Shape shape{M, N, K, m_fastest};
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(ta, tb, shape, ep);
  CSS_LAUNCHED();
  return CSS_OK;
}

}  // namespace gemm
}  // namespace css
