// gemm_tc.cuh -- persistent, warp-specialised tcgen05 GEMM:  C[M,N] = A[M,K] * B[N,K]^T
//
//   A, B : bf16, K-major (row-major [rows, K]); tiles arrive by TMA with the 128-byte
//          swizzle into a kStages-deep shared-memory ring
//   C    : fp32 accumulators in TMEM (2 x BN columns, double-buffered so the epilogue of
//          tile i overlaps the main loop of tile i+1)
//   roles: warp 0 = TMA producer (1 thread), warp 1 = MMA issuer (1 thread) + TMEM owner,
//          warps 2..17 = epilogue (TMEM -> registers -> Epi functor -> global)
//
// One CTA per SM, tiles of 128 x BN visited round-robin.  The epilogue is a template
// parameter: the encoder instantiates bias / bias+GELU / bias+residual writers, the
// batched search instantiates a per-query threshold filter (search_batched.cu).
#pragma once
#include "tc_common.cuh"
#include <cstdlib>

namespace css {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = kGemmEpiWarps;          // 16
constexpr int kColSplit = kGemmEpiColSplit;      // 4 column slices of BN / 4
constexpr int kThreads = (2 + kEpiWarps) * 32;   // 576

constexpr int kSmemBudget = 232448;  // 227 KB: the most one CTA may opt in to

template <int BN, int kEpiStageBytes>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kEpiBytes = kEpiWarps * kEpiStageBytes;  // per-warp output staging
  static constexpr int kStagesFit = (kSmemBudget - kBarBytes - 1024 - kEpiBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = 2 * BN;  // 512 (BN=256) or 256 (BN=128)
  static_assert(kStages >= 2, "pipeline too shallow");  // fallback kernel (CSS_GEMM_2CTA=0); the 2-CTA kernel keeps >= 4
};

struct Shape {
  int M, N, K;
  int m_fastest;  // 1: consecutive tiles share the B (N) block; 0: share the A (M) block
  int panel;      // 1: a CTA (pair) visits every N block of one M block back to back, so its epilogue
                  //    sees whole output rows (row LayerNorm fused into the epilogue)
};

// Epi concept:
//   struct Epi { struct Params {...};
//     static constexpr int kStageBytes;   // shared memory per epilogue warp (output staging), may be 0
//     void tile_begin(m_warp, lane, M);   // before the wait for the tile's accumulator (per-row loads go here)
//     __device__ Epi(const Params&, int epi_thread /*0..511*/, uint8_t* warp_stage);
//     // lane i holds row m_warp + i, columns n0 .. n0+31, of the accumulator
//     __device__ void chunk(int slot, int m_warp, int lane, int M, int n0, const uint32_t (&v)[32]);
//     // prefetch(slot, ...) is called one whole TILE ahead: right after chunk(slot, ...) of the
//     // current tile, with the coordinates of the same chunk of this CTA's next tile (and for
//     // every chunk of the first tile before the loop), so that operands the epilogue reads
//     // from global memory (residual rows) have a tile's worth of time to arrive
//     __device__ void prefetch(int slot, int m_warp, int lane, int M, int n0);
//     __device__ void prefetch_none();   // called instead when this CTA has no next tile
//   static constexpr bool kResidPrefetch: the TMA producer pulls the tile of Params::resid ([M, N] bf16)
//     that the epilogue of the NEXT tile will read into L2 (cp.async.bulk.prefetch.tensor)
//     // m_cta: first row of this CTA's 128 rows of the tile; called by all 8 epilogue warps
//     __device__ void tile_end(int m_cta, int n_blk, int num_n, int M);
//     __device__ void finish(); };
template <int BN, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_r, Shape shape, typename Epi::Params ep) {
  using C = Cfg<BN, Epi::kStageBytes>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic: a round trip through an integer makes the compiler forget that this
  // is shared memory, and the epilogue's staging traffic turns into generic LD / ST (ncu: long-scoreboard stalls)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kStages * C::kABytes;
  uint8_t* sEpi = smem + C::kStages * C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + C::kEpiBytes);
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
  __syncwarp();
  if (warp == 1) tc::tmem_alloc(tmem_slot, C::kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (shape.M + BM - 1) / BM;
  const int num_n = (shape.N + BN - 1) / BN;  // a ragged last block is zero-filled by TMA; Epi masks it
  const int num_tiles = num_m * num_n;
  const int num_kb = shape.K / BK;

  // u-th tile of this CTA
  auto tile_at = [&](int u, int& mb, int& nb) -> bool {
    if (shape.panel) {
      mb = (int)blockIdx.x + (u / num_n) * (int)gridDim.x;
      nb = u % num_n;
      return mb < num_m;
    }
    const int t = (int)blockIdx.x + u * (int)gridDim.x;
    if (t >= num_tiles) return false;
    if (shape.m_fastest) {
      mb = t % num_m;
      nb = t / num_m;
    } else {
      nb = t % num_n;
      mb = t / num_n;
    }
    return true;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      int mb, nb;
      if constexpr (Epi::kResidPrefetch) {
        if (tile_at(0, mb, nb))
          for (int j = 0; j < BN / 64; ++j) tc::tma_prefetch_l2_2d(&tmap_r, nb * BN + j * 64, mb * BM);
      }
      for (int u = 0; tile_at(u, mb, nb); ++u) {
        if constexpr (Epi::kResidPrefetch) {   // residual rows of the next tile -> L2, two mainloops before they are read
          int pmb, pnb;
          if (tile_at(u + 1, pmb, pnb))
            for (int j = 0; j < BN / 64; ++j) tc::tma_prefetch_l2_2d(&tmap_r, pnb * BN + j * 64, pmb * BM);
        }
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
      int mb, nb;
      for (int u = 0; tile_at(u, mb, nb); ++u) {
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
    const int ew = warp - 2;          // 0..15
    const int quarter = warp & 3;     // TMEM lane quarter this warp may read
    const int colq = ew >> 2;         // column slice of the tile
    constexpr int kCols = BN / kColSplit;
    constexpr int kChunks = kCols / 32;
    Epi epi(ep, ew * 32 + lane, sEpi + ew * Epi::kStageBytes);
    int as = 0;
    uint32_t aphase = 0;
    int mb, nb;
    if (tile_at(0, mb, nb)) {  // operands of the first tile (e.g. residual rows) start loading now
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        epi.prefetch(c, mb * BM + quarter * 32, lane, shape.M, nb * BN + colq * kCols + c * 32);
    }
    for (int u = 0; tile_at(u, mb, nb); ++u) {
      // same chunk of the next tile of this CTA (requested a tile ahead)
      int nmb = 0, nnb = 0;
      const bool has_next_tile = tile_at(u + 1, nmb, nnb);
      const int next_m_warp = nmb * BM + quarter * 32;
      const int next_n_base = nnb * BN + colq * kCols;
      const int m_warp = mb * BM + quarter * 32;  // first row of this warp's 32 rows
      epi.tile_begin(m_warp, lane, shape.M);
      tc::mbar_wait(tfull + as, aphase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(as * BN + colq * kCols);
      const int n_base = nb * BN + colq * kCols;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[32];   // one chunk at a time: with four warps per sub-partition the others cover the TMEM latency
        tc::tmem_ld_32x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
        if (c == kChunks - 1) {
          // every TMEM read of this accumulator has landed in registers: hand it back early
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_relaxed(tempty + as);
        }
        epi.chunk(c, m_warp, lane, shape.M, n_base + c * 32, v);
        if (has_next_tile) epi.prefetch(c, next_m_warp, lane, shape.M, next_n_base + c * 32);
        else epi.prefetch_none();
      }
      epi.tile_end(mb * BM, nb, num_n, shape.M);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    epi.finish();
  }
  __syncwarp();  // the single-lane roles rejoin their warps before the block-wide barriers

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, C::kTmemCols);
}

// Host launcher.  A: [M, K] row-major bf16 (ld = lda elements), B: [N, K] row-major bf16.
template <int BN, class Epi>
int launch(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int m_fastest,
           const typename Epi::Params& ep, int n_sm, cudaStream_t st) {
  using C = Cfg<BN, Epi::kStageBytes>;
  CSS_REQUIRE(M >= 1 && N >= 1 && (N % BN == 0 || Epi::kMasksColumns) && K % BK == 0 && K >= BK,
              "gemm shape M=%d N=%d K=%d unsupported (BN=%d)", M, N, K, BN);
  CUtensorMap ta, tb, tr;
  CSS_CHECK(encode_tmap_bf16_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK));
  CSS_CHECK(encode_tmap_bf16_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, BK));
  tr = ta;
  if constexpr (Epi::kResidPrefetch)
    CSS_CHECK(encode_tmap_bf16_2d(&tr, Epi::resid_ptr(ep), (uint64_t)M, (uint64_t)N, (uint64_t)N, BM, 64));
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
  Shape shape{M, N, K, m_fastest, Epi::kPanel ? 1 : 0};
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(ta, tb, tr, shape, ep);
  CSS_LAUNCHED();
  return CSS_OK;
}


// =====================================================================================
// 2-CTA variant: a cluster of two CTAs (one SM pair) computes a 256 x 256 tile with
// tcgen05.mma.cta_group::2.  Each CTA stages its own 128 A rows and HALF of the B tile
// (128 of the 256 N rows), so a pair reads (256 + 256) x K operand rows per 256 x 256
// outputs -- 1.5x less L2->SM traffic per FLOP than the 128 x 256 single-CTA tile, which is
// L2-bandwidth bound (round-1 profile).  The leader CTA (cluster rank 0) issues every MMA;
// its `full` barriers collect the TMA bytes of both CTAs; MMA completion is multicast to
// the `empty` / `tfull` barriers of both; both CTAs' epilogue warps arrive on the leader's
// `tempty`.  Accumulators: 128 rows x 256 columns x 2 buffers in each CTA's TMEM.
// =====================================================================================
template <int BN, int kEpiStageBytes>
struct Cfg2 {
  static_assert(BN == 256, "the 2-CTA kernel is built for 256 x 256 tiles");
  static constexpr int kABytes = BM * BK * 2;          // 128 rows of A per CTA
  static constexpr int kBBytes = (BN / 2) * BK * 2;    // 128 of the 256 B rows per CTA
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kEpiBytes = kEpiWarps * kEpiStageBytes;
  static constexpr int kStagesFit = (kSmemBudget - kBarBytes - 1024 - kEpiBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;
  static constexpr int kTmemCols = 2 * BN;
  static_assert(kStages >= 4, "pipeline too shallow");
};

template <int BN, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_r, Shape shape, typename Epi::Params ep) {
  using C = Cfg2<BN, Epi::kStageBytes>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic: a round trip through an integer makes the compiler forget that this
  // is shared memory, and the epilogue's staging traffic turns into generic LD / ST (ncu: long-scoreboard stalls)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kStages * C::kABytes;
  uint8_t* sEpi = smem + C::kStages * C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + C::kEpiBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tfull = bars + 2 * C::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = tc::cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
    for (int s = 0; s < C::kStages; ++s) {
      tc::mbar_init(full + s, 1);    // leader's own arrive.expect_tx; bytes of both CTAs
      tc::mbar_init(empty + s, 1);   // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tfull + s, 1);
      tc::mbar_init(tempty + s, 2 * kEpiWarps);  // epilogue warps of both CTAs
    }
    tc::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tc::tmem_alloc_2sm(tmem_slot, C::kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();  // peer barriers are initialised before anyone signals them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr int TM = 2 * BM;  // 256 rows per cluster tile
  const int num_m = (shape.M + TM - 1) / TM;
  const int num_n = (shape.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = shape.K / BK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  // u-th tile of this CTA pair
  auto tile_at = [&](int u, int& mb, int& nb) -> bool {
    if (shape.panel) {
      mb = cluster_id + (u / num_n) * num_clusters;
      nb = u % num_n;
      return mb < num_m;
    }
    const int t = cluster_id + u * num_clusters;
    if (t >= num_tiles) return false;
    if (shape.m_fastest) {
      mb = t % num_m;
      nb = t / num_m;
    } else {
      nb = t % num_n;
      mb = t / num_n;
    }
    return true;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      int mb, nb;
      if constexpr (Epi::kResidPrefetch) {
        if (tile_at(0, mb, nb))
          for (int j = 0; j < BN / 64; ++j) tc::tma_prefetch_l2_2d(&tmap_r, nb * BN + j * 64, mb * TM + (int)cta_rank * BM);
      }
      for (int u = 0; tile_at(u, mb, nb); ++u) {
        if constexpr (Epi::kResidPrefetch) {   // this CTA's residual rows of the next tile -> L2
          int pmb, pnb;
          if (tile_at(u + 1, pmb, pnb))
            for (int j = 0; j < BN / 64; ++j)
              tc::tma_prefetch_l2_2d(&tmap_r, pnb * BN + j * 64, pmb * TM + (int)cta_rank * BM);
        }
        const int a_row = mb * TM + (int)cta_rank * BM;
        const int b_row = nb * BN + (int)cta_rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty + stage, phase ^ 1);
          if (leader) tc::mbar_expect_tx(full + stage, 2 * C::kStageBytes);
          tc::tma_load_2d_2sm(sA + stage * C::kABytes, &tmap_a, full + stage, kb * BK, a_row, tc::kEvictNormal);
          tc::tma_load_2d_2sm(sB + stage * C::kBBytes, &tmap_b, full + stage, kb * BK, b_row, tc::kEvictNormal);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc = tc::make_idesc_bf16_f32(TM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int mb, nb;
      for (int u = 0; tile_at(u, mb, nb); ++u) {
        tc::mbar_wait(tempty + as, aphase ^ 1);
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(full + stage, phase);
          tc::tc_fence_after();
          const uint64_t da = tc::make_kmajor_sw128_desc(tc::smem_u32(sA + stage * C::kABytes));
          const uint64_t db = tc::make_kmajor_sw128_desc(tc::smem_u32(sB + stage * C::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_bf16_2sm(tmem_d, da + k * tc::kDescKStep, db + k * tc::kDescKStep, idesc, (kb | k) != 0);
          tc::umma_commit_2sm(empty + stage);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc::umma_commit_2sm(tfull + as);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = warp - 2;          // 0..15
    const int quarter = warp & 3;
    const int colq = ew >> 2;
    constexpr int kCols = BN / kColSplit;
    constexpr int kChunks = kCols / 32;
    Epi epi(ep, ew * 32 + lane, sEpi + ew * Epi::kStageBytes);
    int as = 0;
    uint32_t aphase = 0;
    const int row_off = (int)cta_rank * BM + quarter * 32;
    int mb, nb;
    if (tile_at(0, mb, nb)) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        epi.prefetch(c, mb * TM + row_off, lane, shape.M, nb * BN + colq * kCols + c * 32);
    }
    for (int u = 0; tile_at(u, mb, nb); ++u) {
      int nmb = 0, nnb = 0;
      const bool has_next_tile = tile_at(u + 1, nmb, nnb);
      const int next_m_warp = nmb * TM + row_off;
      const int next_n_base = nnb * BN + colq * kCols;
      const int m_warp = mb * TM + row_off;
      epi.tile_begin(m_warp, lane, shape.M);
      tc::mbar_wait(tfull + as, aphase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(as * BN + colq * kCols);
      const int n_base = nb * BN + colq * kCols;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[32];
        tc::tmem_ld_32x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
        if (c == kChunks - 1) {
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_leader_relaxed(tempty + as);
        }
        epi.chunk(c, m_warp, lane, shape.M, n_base + c * 32, v);
        if (has_next_tile) epi.prefetch(c, next_m_warp, lane, shape.M, next_n_base + c * 32);
        else epi.prefetch_none();
      }
      epi.tile_end(mb * TM + (int)cta_rank * BM, nb, num_n, shape.M);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    epi.finish();
  }
  __syncwarp();  // the single-lane roles rejoin their warps before the block-wide barriers

  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still signal it
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, C::kTmemCols);
}

// 1 = use the 2-CTA kernel (default), 0 = single-CTA kernel; CSS_GEMM_2CTA overrides.
inline bool use_2cta() {
  static const int v = [] {
    const char* e = getenv("CSS_GEMM_2CTA");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

template <int BN, class Epi>
int launch2(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int m_fastest,
            const typename Epi::Params& ep, int n_sm, cudaStream_t st) {
  using C = Cfg2<BN, Epi::kStageBytes>;
  CSS_REQUIRE(M >= 1 && N >= 1 && (N % BN == 0 || Epi::kMasksColumns) && K % BK == 0 && K >= BK,
              "gemm shape M=%d N=%d K=%d unsupported (BN=%d)", M, N, K, BN);
  CUtensorMap ta, tb, tr;
  CSS_CHECK(encode_tmap_bf16_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK));
  CSS_CHECK(encode_tmap_bf16_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN / 2, BK));
  tr = ta;
  if constexpr (Epi::kResidPrefetch)
    CSS_CHECK(encode_tmap_bf16_2d(&tr, Epi::resid_ptr(ep), (uint64_t)M, (uint64_t)N, (uint64_t)N, BM, 64));
  auto kern = gemm_tc2_kernel<BN, Epi>;
  static std::atomic<uint64_t> attr_set{0};
  int dev = 0;
  CSS_CUDA(cudaGetDevice(&dev));
  if (!(attr_set.load(std::memory_order_relaxed) >> (dev & 63) & 1)) {
    CSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set.fetch_or(uint64_t(1) << (dev & 63), std::memory_order_relaxed);
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  int clusters = n_sm / 2;
  if (tiles < clusters) clusters = tiles;
  Shape shape{M, N, K, m_fastest, Epi::kPanel ? 1 : 0};
  kern<<<2 * clusters, kThreads, C::kSmemBytes, st>>>(ta, tb, tr, shape, ep);
  CSS_LAUNCHED();
  return CSS_OK;
}

// Dispatch used by the encoder and the batched search.
template <int BN, class Epi>
int run(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int m_fastest,
        const typename Epi::Params& ep, int n_sm, cudaStream_t st) {
  if (use_2cta()) return launch2<BN, Epi>(A, lda, B, ldb, M, N, K, m_fastest, ep, n_sm, st);
  return launch<BN, Epi>(A, lda, B, ldb, M, N, K, m_fastest, ep, n_sm, st);
}

}  // namespace gemm
}  // namespace css
