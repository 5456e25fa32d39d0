// query_kernels.cuh -- kernels of the interactive query path (one short query is
// encoded per search, src/embeddings.py:216-222 called from the MCP server / CLI with a single text).
//
// A query is a bucket of 32 or 64 token rows.  The tcgen05 GEMMs of the indexing path put a
// 128-row tile on 6 (N = 768) to 24 (N = 3072) CTAs, each walking K serially: 9-30 us per
// projection, 0.89 ms per query, all of it latency.  At this size the forward pass is a weight
// stream (170 MB of bf16 per query) and the right shape is the opposite one: every SM pulls a
// slice of the weight matrix with all of its loads in flight at once.
//
//   skinny_gemm_kernel   out[M, N] = A[M, K] W[N, K]^T, M = 32 / 64.  CTA = 8 output columns,
//       KSPLIT warps each own K / KSPLIT of the reduction: a warp issues ALL of its weight loads
//       (16 B per lane, 8 rows x 64 B per instruction) before the first mma.sync, so the whole matrix
//       is in flight across the grid; activations come from L2 / L1.  The k index inside a 32-wide
//       block is permuted identically for both operands (lane q holds k = 8q .. 8q+7), which lets
//       both fragments be plain 16-byte loads.  Partial sums meet in shared memory (fixed order);
//       the epilogues are the ones of the tcgen05 path (bias / folded LayerNorm / GELU / residual +
//       row statistics), one thread per row.
//   query_attention_kernel   one CTA per (head, sequence), L <= 64: scores, softmax and P V in fp32
//       from shared memory, a warp per query row.
//   ln_stats_finalize_parts_kernel   (mean, rstd) per row from the N / 8 per-CTA partials.
#pragma once
#include "encoder_kernels.cuh"

namespace css {
namespace enc {

enum : int { kSkBias = 0, kSkFold = 1, kSkFoldGelu = 2, kSkResidLN = 3 };

struct SkinnyParams {
  const __nv_bfloat16* A;   // [M, K], row stride K
  const __nv_bfloat16* W;   // [N, K]
  __nv_bfloat16* out;       // [M, ldo]
  int ldo;
  const float* bias;        // b, or d when folded
  const float* colsum;      // folded: c [N]
  const float2* stat_parts; // folded: partial statistics of the rows of A; kSkResidLN: of the rows of `resid`
                            // when it is a pre-LayerNorm value (else nullptr).  [M][kQueryParts] (sum, sum of squares)
  float eps;
  const __nv_bfloat16* resid;   // kSkResidLN: [M, 768]
  const float* rgamma;
  const float* rbeta;
  float2* parts;                // kSkResidLN: [M][N / 8] (sum, sum of squares) of the output row slice
};

// Programmatic dependent launch: the next kernel of the stream may start (and issue its weight loads)
// while this one is still running; it must not touch anything a predecessor writes or reads before
// pdl_wait() returns (all predecessors complete, their memory visible).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kQueryMaxRows = 64;
constexpr int kQueryParts = kHidden / 8;   // partial statistics per row

// NT: 8-column groups per CTA (each warp runs NT n8 tiles on the same activation fragments: the activation
// re-read, which is what the SMs mostly ingest here, drops by NT; the wide projections use 2).
template <int MT, int KSPLIT, int KITERS, int MODE, int NT = 1>
static __global__ void __launch_bounds__(KSPLIT * 32) skinny_gemm_kernel(const SkinnyParams p) {
  constexpr int K = KSPLIT * KITERS * 32;
  constexpr int M = MT * 16;
  static_assert(M * NT <= KSPLIT * 32, "one epilogue thread per row and column group");
  static_assert(MODE != kSkResidLN || NT == 1, "the statistics partials are per 8-column CTA");
  __shared__ __align__(16) float red[KSPLIT][M][8 * NT];
  __shared__ float2 smr[M];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int kbase = warp * (KITERS * 32) + q * 8;

  // the warp's whole weight slice: NT * KITERS independent 16-byte loads per lane, issued back to back
  uint4 b[NT][KITERS];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const uint4* wp = reinterpret_cast<const uint4*>(p.W + (size_t)(blockIdx.x * (8 * NT) + nt * 8 + g) * K + kbase);
#pragma unroll
    for (int it = 0; it < KITERS; ++it)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b[nt][it].x), "=r"(b[nt][it].y), "=r"(b[nt][it].z), "=r"(b[nt][it].w)
                   : "l"(wp + it * 4));
  }
  // the epilogue thread of (row m, column group): its per-column constants are requested now, with the weights
  const int m = threadIdx.x % M, cg = threadIdx.x / M;
  const bool epi = threadIdx.x < M * NT;
  const int n0 = blockIdx.x * (8 * NT) + (epi ? cg : 0) * 8;
  float bias[8], cs[8], be[8];   // cs: column sums c (folded) or the residual LayerNorm's gamma; be: its beta
#pragma unroll
  for (int i = 0; i < 8; ++i) bias[i] = cs[i] = be[i] = 0.f;
  if (epi) {
    auto ld8 = [&](const float* src, float (&dst)[8]) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(src + n0)), y = __ldg(reinterpret_cast<const float4*>(src + n0) + 1);
      dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w; dst[4] = y.x; dst[5] = y.y; dst[6] = y.z; dst[7] = y.w;
    };
    ld8(p.bias, bias);
    if constexpr (MODE == kSkFold || MODE == kSkFoldGelu) ld8(p.colsum, cs);
    if constexpr (MODE == kSkResidLN) {
      if (p.stat_parts != nullptr) {
        ld8(p.rgamma, cs);
        ld8(p.rbeta, be);
      }
    }
  }
  pdl_wait();   // the weights above are constants; everything below reads what the previous kernels wrote
  uint4 r4 = make_uint4(0, 0, 0, 0);
  if constexpr (MODE == kSkResidLN) {
    if (epi) r4 = *reinterpret_cast<const uint4*>(p.resid + (size_t)m * kHidden + n0);
  }
  // partial row statistics: TPR adjacent lanes per row, all loads of a thread independent.  Where they are
  // reduced was measured on one box (p50 per query): 64-row bucket, after the main loop (loads overlap
  // the activation loads) 0.526 against 0.570 ms; 32-row bucket, before it 0.311 against 0.324-0.353 ms
  constexpr int TPR = KSPLIT * 32 / M;
  constexpr bool kLateStats = MT >= 4;
  static_assert(TPR >= 1 && TPR <= 32 && (TPR & (TPR - 1)) == 0 && kQueryParts % TPR == 0, "row groups");
  const bool has_stats = MODE != kSkBias && p.stat_parts != nullptr;
  float2 sp[kQueryParts / TPR];
  auto reduce_stats = [&]() {
    float s = 0.f, sq = 0.f;
#pragma unroll
    for (int j = 0; j < kQueryParts / TPR; ++j) {
      s += sp[j].x;
      sq += sp[j].y;
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float mean = s * (1.f / kHidden);
    if (threadIdx.x % TPR == 0)
      smr[threadIdx.x / TPR] = make_float2(mean, rsqrtf(fmaxf(sq * (1.f / kHidden) - mean * mean, 0.f) + p.eps));
  };
  if (has_stats) {
    const float2* src = p.stat_parts + (size_t)(threadIdx.x / TPR) * kQueryParts + threadIdx.x % TPR;
#pragma unroll
    for (int j = 0; j < kQueryParts / TPR; ++j) sp[j] = src[j * TPR];
    if constexpr (!kLateStats) reduce_stats();
  }
  float acc[NT][MT][4], acc2[NT][MT][4];   // 64-row bucket: two accumulator sets, half the dependent mma chain (0.526 against 0.537 ms)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][mt][i] = acc2[nt][mt][i] = 0.f;
#pragma unroll
  for (int it = 0; it < KITERS; ++it) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const __nv_bfloat16* ap = p.A + (size_t)(mt * 16 + g) * K + kbase + it * 32;
      const uint4 lo = *reinterpret_cast<const uint4*>(ap);
      const uint4 hi = *reinterpret_cast<const uint4*>(ap + (size_t)8 * K);
      const uint32_t a0[4] = {lo.x, hi.x, lo.y, hi.y};
      const uint32_t a1[4] = {lo.z, hi.z, lo.w, hi.w};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mma_bf16_16816(acc[nt][mt], a0, b[nt][it].x, b[nt][it].y);
        mma_bf16_16816(kLateStats ? acc2[nt][mt] : acc[nt][mt], a1, b[nt][it].z, b[nt][it].w);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][mt][i] += acc2[nt][mt][i];
  if constexpr (kLateStats) {
    if (has_stats) reduce_stats();
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      *reinterpret_cast<float2*>(&red[warp][mt * 16 + g][nt * 8 + 2 * q]) = make_float2(acc[nt][mt][0], acc[nt][mt][1]);
      *reinterpret_cast<float2*>(&red[warp][mt * 16 + g + 8][nt * 8 + 2 * q]) = make_float2(acc[nt][mt][2], acc[nt][mt][3]);
    }
  __syncthreads();
  // only now may the next kernel start prefetching its weights: one kernel ahead, never a cascade of
  // waiting grids (measured: triggering at kernel entry made the whole path 1.3-2x slower)
  pdl_launch_dependents();
  if (!epi) return;
  float v[8];
  {
    const float4 x = *reinterpret_cast<const float4*>(&red[0][m][cg * 8]), y = *reinterpret_cast<const float4*>(&red[0][m][cg * 8 + 4]);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
  }
#pragma unroll
  for (int w = 1; w < KSPLIT; ++w) {
    const float4 x = *reinterpret_cast<const float4*>(&red[w][m][cg * 8]), y = *reinterpret_cast<const float4*>(&red[w][m][cg * 8 + 4]);
    v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
  }
  float f[8];
  if constexpr (MODE == kSkBias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = v[i] + bias[i];
  } else if constexpr (MODE == kSkFold || MODE == kSkFoldGelu) {
    // LN(a) W^T + b = rstd (a (W gamma)^T - mu c) + d   (EpiBiasBf16<., true>)
    const float2 mr = smr[m];
    const float rs = mr.y, nb = -mr.x * mr.y;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f[i] = fmaf(rs, v[i], fmaf(nb, cs[i], bias[i]));
      if constexpr (MODE == kSkFoldGelu) f[i] = gelu_erf(f[i]);
    }
  } else {
    // v + bias + residual, the residual rebuilt as LN(resid) when `resid` is a pre-LayerNorm value (EpiResidLN<false>)
    float r[8] = {bf16_lo(r4.x), bf16_hi(r4.x), bf16_lo(r4.y), bf16_hi(r4.y), bf16_lo(r4.z), bf16_hi(r4.z), bf16_lo(r4.w), bf16_hi(r4.w)};
    if (p.stat_parts != nullptr) {
      const float2 mr = smr[m];
      const float rs = mr.y, nb = -mr.x * mr.y;
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fmaf(fmaf(r[i], rs, nb), cs[i], be[i]);
    }
    float s = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f[i] = (v[i] + bias[i]) + r[i];
      s += f[i];
      sq = fmaf(f[i], f[i], sq);
    }
    p.parts[(size_t)m * gridDim.x + blockIdx.x] = make_float2(s, sq);
  }
  *reinterpret_cast<uint4*>(p.out + (size_t)m * p.ldo + n0) =
      make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// (mean, rstd) per row from the kQueryParts per-CTA partials of skinny_gemm_kernel<kSkResidLN>: a warp per row
static __global__ void ln_stats_finalize_parts_kernel(const float2* __restrict__ parts, int T, float eps,
                                                      float2* __restrict__ mr) {
  pdl_wait();
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= T) return;
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < kQueryParts / 32; ++i) {
    const float2 a = parts[(size_t)t * kQueryParts + i * 32 + lane];
    s += a.x;
    q += a.y;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  const float mean = s * (1.f / kHidden);
  if (lane == 0) mr[t] = make_float2(mean, rsqrtf(fmaxf(q * (1.f / kHidden) - mean * mean, 0.f) + eps));
}

// Attention of one (head, sequence) with L <= 64 keys: scores = q.k / 8 + rel_table[h][clamp(j - i)],
// softmax and P V in fp32.  8 warps, a warp per query row; lane = key for the scores, lane = two
// output dimensions for P V.
constexpr int kQueryAttnThreads = 256;
constexpr int kQueryAttnRowGroups = 4;   // grid.z: CTAs sharing one (head, sequence), 8 query rows per pass each
static __global__ void __launch_bounds__(kQueryAttnThreads)
query_attention_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ cu,
                       const float* __restrict__ rel_table, int rel_half, __nv_bfloat16* __restrict__ ctx) {
  constexpr int kLd = 3 * kHidden;
  constexpr int kKs = 33;   // words per K row: 32 + 1 pad, conflict-free with lane = key
  __shared__ uint32_t sQ[kQueryMaxRows][32], sK[kQueryMaxRows][kKs], sV[kQueryMaxRows][32];
  __shared__ __align__(16) float sP[kQueryAttnThreads / 32][kQueryMaxRows];
  const int h = blockIdx.x, s = blockIdx.y;
  pdl_wait();
  const int t0 = cu[s];
  const int L = min(cu[s + 1] - t0, kQueryMaxRows);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kQueryMaxRows * 8; i += kQueryAttnThreads) {
    const int r = i >> 3, c = i & 7;
    uint4 qv = make_uint4(0, 0, 0, 0), kv = qv, vv = qv;
    if (r < L) {
      const __nv_bfloat16* gp = qkv + (size_t)(t0 + r) * kLd + h * kHeadDim + c * 8;
      qv = *reinterpret_cast<const uint4*>(gp);
      kv = *reinterpret_cast<const uint4*>(gp + kHidden);
      vv = *reinterpret_cast<const uint4*>(gp + 2 * kHidden);
    }
    *reinterpret_cast<uint4*>(&sQ[r][c * 4]) = qv;
    *reinterpret_cast<uint4*>(&sV[r][c * 4]) = vv;
    sK[r][c * 4 + 0] = kv.x; sK[r][c * 4 + 1] = kv.y; sK[r][c * 4 + 2] = kv.z; sK[r][c * 4 + 3] = kv.w;
  }
  __syncthreads();
  pdl_launch_dependents();
  const float* rt = rel_table + (size_t)h * (2 * rel_half + 1) + rel_half;
  constexpr int kWarps = kQueryAttnThreads / 32;
  for (int i = blockIdx.z * kWarps + warp; i < L; i += kWarps * gridDim.z) {
    float sa0 = 0.f, sa1 = 0.f, sb0 = 0.f, sb1 = 0.f;   // split accumulators: the chains are latency-bound
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const uint32_t qw = sQ[i][d], ka = sK[lane][d], kb = sK[lane + 32][d];
      const float q0 = bf16_lo(qw), q1 = bf16_hi(qw);
      sa0 = fmaf(q0, bf16_lo(ka), sa0);
      sa1 = fmaf(q1, bf16_hi(ka), sa1);
      sb0 = fmaf(q0, bf16_lo(kb), sb0);
      sb1 = fmaf(q1, bf16_hi(kb), sb1);
    }
    float sa = sa0 + sa1, sb = sb0 + sb1;
    const int ja = lane, jb = lane + 32;
    sa = ja < L ? fmaf(sa, 0.125f, __ldg(rt + max(-rel_half, min(rel_half, ja - i)))) : -INFINITY;
    sb = jb < L ? fmaf(sb, 0.125f, __ldg(rt + max(-rel_half, min(rel_half, jb - i)))) : -INFINITY;
    float mx = fmaxf(sa, sb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float pa = __expf(sa - mx), pb = __expf(sb - mx);   // exp(-inf) = 0 for the keys that do not exist
    const float inv = 1.f / warp_sum(pa + pb);
    sP[warp][ja] = pa * inv;
    sP[warp][jb] = pb * inv;
    __syncwarp();
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    const int L4 = (L + 3) & ~3;   // rows L .. L4-1 of V are zero and their weights are 0
    for (int j = 0; j < L4; j += 4) {
      const float4 pj = *reinterpret_cast<const float4*>(&sP[warp][j]);
      const float pv[4] = {pj.x, pj.y, pj.z, pj.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t vw = sV[j + u][lane];
        o0[u] = fmaf(pv[u], bf16_lo(vw), o0[u]);
        o1[u] = fmaf(pv[u], bf16_hi(vw), o1[u]);
      }
    }
    *reinterpret_cast<uint32_t*>(ctx + (size_t)(t0 + i) * kHidden + h * kHeadDim + 2 * lane) =
        pack_bf16((o0[0] + o0[1]) + (o0[2] + o0[3]), (o1[0] + o1[1]) + (o1[2] + o1[3]));
    __syncwarp();
  }
}

// launch with the programmatic-stream-serialization attribute (the kernel calls pdl_wait() itself)
template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  // CSS_QUERY_PDL=0: plain stream order (griddepcontrol.* are no-ops then)
  static const bool pdl = [] { const char* v = getenv("CSS_QUERY_PDL"); return v ? atoi(v) != 0 : true; }();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <int MT, int KSPLIT, int KITERS, int MODE, int NT = 1>
inline cudaError_t skinny_launch(const SkinnyParams& p, int N, cudaStream_t st) {
  return pdl_launch(skinny_gemm_kernel<MT, KSPLIT, KITERS, MODE, NT>, dim3((unsigned)(N / (8 * NT))), dim3(KSPLIT * 32), st, p);
}

}  // namespace enc
}  // namespace css
