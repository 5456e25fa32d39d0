// attention_tc.cuh -- E3 on the 5th-generation tensor cores.
//
// One persistent CTA per SM walks work items (sequence, head, block of 128 query rows):
//   warp 0      TMA producer: Q [128 x 64], K [Lp x 64], V [Lp x 64] bf16 tiles straight out of the
//               packed qkv activation [T, 2304] (128-byte swizzle)
//   warp 1      MMA issuer:   S = Q K^T   (tcgen05.mma M128 x N<=256 x K16, fp32 in TMEM cols [0, Lp))
//                             O = P V     (M128 x N64, V is the MN-major B operand, TMEM cols [448, 512))
//   warps 2-9   softmax + epilogue: two groups of 4 warps, one query row per thread in each group
//               (= one TMEM lane); group 0 owns the even 64-key blocks, group 1 the odd ones:
//               pass 1  row maximum of the raw scores           (tcgen05.ld)
//               pass 2  p = exp2(s*c + rel[j-i] - m), 64 keys at a time, written as the bf16 A operand of
//                       the PV product into a double-buffered shared tile -> the PV MMAs of block b run
//                       while block b+1 is exponentiated
//               epilogue O / sum -> ctx (bf16)
// Keys beyond the sequence end do not exist in the packed layout (their tile rows hold the next
// sequence's tokens or TMA zero fill): their probabilities are forced to 0, which is what the
// reference's additive finfo.min mask produces.  Sequences longer than kAttnTcMaxLen take the
// mma.sync kernel (encoder_kernels.cuh).
#pragma once
#include "encoder_kernels.cuh"

namespace css {
namespace enc {

constexpr int kAttnTcMaxLen = 448;                 // S occupies TMEM columns [0, Lp), O [448, 512)
constexpr int kAttnTcThreads = 320;                // 10 warps: TMA, MMA, 2 x 4 softmax
constexpr int kAtQ = 128 * 128;                    // Q tile bytes
constexpr int kAtKV = kAttnTcMaxLen * 128;         // K / V tile bytes (max)
constexpr int kAtP = 128 * 128;                    // one P block (128 rows x 64 keys bf16)
constexpr int kAtRelStride = 2 * kAttnTcMaxLen;     // floats per head: bias(d) * log2e for d in [-447, 448)
constexpr int kAtRel = kHeads * kAtRelStride * 4;  // every head's window, staged once per CTA
constexpr int kAtCuMax = 1024;                     // cu_seqlens entries cached in shared memory
constexpr int kAtXch = 4 * 128 * 4;                // per-row partial max / sum of the two softmax groups
constexpr int kAtSmem = kAtQ + 2 * kAtKV + 2 * kAtP + kAtRel + kAtXch + (kAtCuMax + 1) * 4 + 256 + 1024;
constexpr uint32_t kOCol = 448;

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// V tile as the MN-major (N = head dim contiguous) B operand: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;               // LBO: one 64-element MN atom only, unused
  d |= static_cast<uint64_t>(1024 >> 4) << 32;        // SBO: next group of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct AttnTcParams {
  const int32_t* cu;
  const float* rel_table;   // [heads][2*rel_half+1]
  const float* rel_max;     // [heads] max over the head's table
  int rel_half;
  int n_seq;
  int nqb;                  // query blocks per sequence in the item space = ceil(max_len / 128)
  __nv_bfloat16* ctx;       // [T, 768]
};

static __global__ void __launch_bounds__(kAttnTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAtQ;
  uint8_t* sV = sK + kAtKV;
  uint8_t* sP = sV + kAtKV;                       // 2 blocks
  float* sRel = reinterpret_cast<float*>(sP + 2 * kAtP);
  float* sXch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sRel) + kAtRel);   // [2][128] max, [2][128] sum
  int32_t* sCu = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(sXch) + kAtXch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sCu) + ((kAtCuMax + 1) * 4 + 7) / 8 * 8);
  uint64_t* qk_full = bars + 0;
  uint64_t* qk_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* s_full = bars + 4;
  uint64_t* s_free = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* p_full = bars + 7;    // [2]
  uint64_t* p_empty = bars + 9;   // [2]
  uint64_t* o_free = bars + 11;   // all 8 softmax warps have read O of the previous item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_q);
    tc::prefetch_tmap(&tmap_kv);
    tc::mbar_init(qk_full, 1);
    tc::mbar_init(qk_empty, 1);
    tc::mbar_init(v_full, 1);
    tc::mbar_init(v_empty, 1);
    tc::mbar_init(s_full, 1);
    tc::mbar_init(s_free, 8);
    tc::mbar_init(o_full, 1);
    tc::mbar_init(o_free, 8);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(p_full + i, 128);
      tc::mbar_init(p_empty + i, 1);
    }
    tc::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  // per-CTA constants: every head's bias window (x log2e) and the sequence offsets, so that no
  // item starts with a chain of dependent global loads (round-1 profile: ~2K cycles per item)
  for (int x = threadIdx.x; x < kHeads * kAtRelStride; x += kAttnTcThreads) {
    const int hh = x / kAtRelStride, d = x % kAtRelStride - (kAttnTcMaxLen - 1);
    const int dc = max(-p.rel_half, min(p.rel_half, d));
    sRel[x] = p.rel_table[(size_t)hh * (2 * p.rel_half + 1) + p.rel_half + dc] * 1.4426950408889634f;
  }
  const bool cu_in_smem = p.n_seq <= kAtCuMax;
  if (cu_in_smem)
    for (int x = threadIdx.x; x <= p.n_seq; x += kAttnTcThreads) sCu[x] = p.cu[x];
  const int32_t* cu = cu_in_smem ? sCu : p.cu;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_items = p.n_seq * kHeads * p.nqb;

  if (warp == 0) {
    if (lane == 0) {
      // ============================ TMA producer ============================
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int qb = item % p.nqb, h = (item / p.nqb) % kHeads, s = item / (p.nqb * kHeads);
        const int t0 = cu[s], L = cu[s + 1] - t0;
        if (qb * 128 >= L) continue;
        const int Lp = (L + 63) & ~63;
        const int nkb = Lp >> 6;
        tc::mbar_wait(qk_empty, (it & 1) ^ 1);
        tc::mbar_expect_tx(qk_full, kAtQ + Lp * 128);
        tc::tma_load_2d(sQ, &tmap_q, qk_full, h * kHeadDim, t0 + qb * 128, tc::kEvictNormal);
        for (int b = 0; b < nkb; ++b)
          tc::tma_load_2d(sK + b * 8192, &tmap_kv, qk_full, kHidden + h * kHeadDim, t0 + b * 64, tc::kEvictNormal);
        tc::mbar_wait(v_empty, (it & 1) ^ 1);
        tc::mbar_expect_tx(v_full, Lp * 128);
        for (int b = 0; b < nkb; ++b)
          tc::tma_load_2d(sV + b * 8192, &tmap_kv, v_full, 2 * kHidden + h * kHeadDim, t0 + b * 64, tc::kEvictNormal);
        ++it;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ============================ MMA issuer ============================
      uint32_t it = 0;
      uint32_t uses[2] = {0, 0};  // how often each P buffer has been consumed (block b uses buffer b & 1)
      const uint32_t idesc_pv = tc::make_idesc_bf16_f32(128, kHeadDim) | (1u << 16);  // B is MN-major
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int qb = item % p.nqb, s = item / (p.nqb * kHeads);
        const int t0 = cu[s], L = cu[s + 1] - t0;
        if (qb * 128 >= L) continue;
        const int Lp = (L + 63) & ~63;
        const int nkb = Lp >> 6;
        tc::mbar_wait(qk_full, it & 1);
        tc::mbar_wait(s_free, (it & 1) ^ 1);   // softmax has finished reading the previous S
        tc::tc_fence_after();
        const uint64_t dq = tc::make_kmajor_sw128_desc(tc::smem_u32(sQ));
        for (int n0 = 0; n0 < Lp; n0 += 256) {
          const int n = (Lp - n0) < 256 ? (Lp - n0) : 256;
          const uint32_t idesc = tc::make_idesc_bf16_f32(128, n);
          const uint64_t dk = tc::make_kmajor_sw128_desc(tc::smem_u32(sK + n0 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(tmem_base + (uint32_t)n0, dq + k * tc::kDescKStep, dk + k * tc::kDescKStep, idesc, k != 0);
        }
        tc::umma_commit(qk_empty);
        tc::umma_commit(s_full);
        tc::mbar_wait(v_full, it & 1);
        tc::mbar_wait(o_free, (it & 1) ^ 1);   // the previous item's O has been read out
        for (int b = 0; b < nkb; ++b) {
          const uint32_t buf = b & 1;
          tc::mbar_wait(p_full + buf, uses[buf] & 1);
          ++uses[buf];
          tc::tc_fence_after();
          const uint64_t dp = tc::make_kmajor_sw128_desc(tc::smem_u32(sP + buf * kAtP));
          const uint64_t dv = make_mnmajor_sw128_desc(tc::smem_u32(sV + b * 8192));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 16 keys per MMA: +32 B in P's rows, +16 rows (2048 B) in V
            tc::umma_bf16(tmem_base + kOCol, dp + k * tc::kDescKStep, dv + k * (2048 >> 4), idesc_pv, (b | k) != 0);
          tc::umma_commit(p_empty + buf);
        }
        tc::umma_commit(v_empty);
        tc::umma_commit(o_full);
        ++it;
      }
    }
  } else {
    // ============================ softmax + epilogue ============================
    const int grp = (warp - 2) >> 2;           // 0: even key blocks, 1: odd key blocks
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;         // row of the query block == TMEM lane
    const int st = threadIdx.x - 64;           // 0..255 among the softmax threads
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kScale = 0.125f * kLog2e;  // 1/sqrt(64) * log2(e)
    float* sMax = sXch;                        // [2][128]
    float* sSum = sXch + 256;                  // [2][128]
    uint32_t it = 0, u = 0;                    // u: uses of this group's P buffer
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int qb = item % p.nqb, h = (item / p.nqb) % kHeads, s = item / (p.nqb * kHeads);
      const int t0 = cu[s], L = cu[s + 1] - t0;
      if (qb * 128 >= L) continue;
      const int Lp = (L + 63) & ~63;
      const int nkb = Lp >> 6;
      const int q0 = qb * 128;
      const int i = min(q0 + r, L - 1);        // rows past the end mirror the last row, never stored
      const float* rel_i = sRel + h * kAtRelStride + (kAttnTcMaxLen - 1) - i;   // rel_i[j] = bias(j - i) * log2e
      tc::mbar_wait(s_full, it & 1);
      tc::tc_fence_after();
      // ---- pass 1: maximum of the raw scores over this group's real keys ----
      float mx = -INFINITY;
      for (int b = grp; b < nkb; b += 2) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c = b * 64 + hf * 32;
          if (c >= L) break;
          uint32_t v[32];
          tc::tmem_ld_32x32(lane_addr + (uint32_t)c, v);
          tc::tmem_ld_wait();
          if (c + 32 <= L) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < L) mx = fmaxf(mx, __uint_as_float(v[j]));
          }
        }
      }
      sMax[grp * 128 + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, sMax[(grp ^ 1) * 128 + r]);
      // upper bound of the row maximum of (s/8 + bias) in the log2 domain
      const float m_hat = mx * kScale + __ldg(p.rel_max + h) * kLog2e;
      // ---- pass 2: probabilities of this group's blocks as the A operand of P V ----
      float sum = 0.f;
      for (int b = grp; b < nkb; b += 2, ++u) {
        tc::mbar_wait(p_empty + grp, (u & 1) ^ 1);
        uint8_t* prow = sP + grp * kAtP + r * 128;
        uint32_t vv[2][32];
        tc::tmem_ld_32x32(lane_addr + (uint32_t)(b * 64), vv[0]);
        tc::tmem_ld_32x32(lane_addr + (uint32_t)(b * 64 + 32), vv[1]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c = b * 64 + hf * 32;
          const uint32_t (&v)[32] = vv[hf];
          float pr[32];
          if (c + 32 <= L) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pr[j] = ex2_approx(fmaf(__uint_as_float(v[j]), kScale, rel_i[c + j] - m_hat));
              sum += pr[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float e = ex2_approx(fmaf(__uint_as_float(v[j]), kScale, rel_i[c + j] - m_hat));
              pr[j] = (c + j < L) ? e : 0.f;
              sum += pr[j];
            }
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 o;
            o.x = pack_bf16(pr[q4 * 8 + 0], pr[q4 * 8 + 1]);
            o.y = pack_bf16(pr[q4 * 8 + 2], pr[q4 * 8 + 3]);
            o.z = pack_bf16(pr[q4 * 8 + 4], pr[q4 * 8 + 5]);
            o.w = pack_bf16(pr[q4 * 8 + 6], pr[q4 * 8 + 7]);
            const int chunk = hf * 4 + q4;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = o;
          }
        }
        fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core
        tc::tc_fence_before();
        tc::mbar_arrive(p_full + grp);
      }
      // S of this item is dead for this warp: the next item's Q K^T may overwrite it once all 8 agree
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(s_free);
      sSum[grp * 128 + r] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float inv = 1.f / (sum + sSum[(grp ^ 1) * 128 + r]);
      // ---- epilogue: group g writes head-dim columns [32 g, 32 g + 32) ----
      tc::mbar_wait(o_full, it & 1);
      tc::tc_fence_after();
      {
        uint32_t v[32];
        tc::tmem_ld_32x32(lane_addr + kOCol + grp * 32, v);
        tc::tmem_ld_wait();
        if (q0 + r < L) {
          __nv_bfloat16* dst = p.ctx + (size_t)(t0 + q0 + r) * kHidden + h * kHeadDim + grp * 32;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[q4 * 8 + 0]) * inv, __uint_as_float(v[q4 * 8 + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(v[q4 * 8 + 2]) * inv, __uint_as_float(v[q4 * 8 + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(v[q4 * 8 + 4]) * inv, __uint_as_float(v[q4 * 8 + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(v[q4 * 8 + 6]) * inv, __uint_as_float(v[q4 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + q4 * 8) = o;
          }
        }
      }
      tc::tc_fence_before();   // O reads are ordered before the next item's first P V
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(o_free);
      ++it;
    }
  }
  __syncwarp();
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// Host launcher.  qkv: [T, 2304] bf16 packed (q | k | v).
static int attention_tc_launch(const __nv_bfloat16* qkv, int T, const int32_t* cu_dev, int n_seq, int max_len,
                               const float* rel_table, const float* rel_max, int rel_half, __nv_bfloat16* ctx,
                               int n_sm, cudaStream_t st) {
  CUtensorMap tq, tkv;
  CSS_CHECK(encode_tmap_bf16_2d(&tq, qkv, (uint64_t)T, 3 * kHidden, 3 * kHidden, 128, 64));
  CSS_CHECK(encode_tmap_bf16_2d(&tkv, qkv, (uint64_t)T, 3 * kHidden, 3 * kHidden, 64, 64));
  static std::atomic<uint64_t> attr_set{0};
  int dev = 0;
  CSS_CUDA(cudaGetDevice(&dev));
  if (!(attr_set.load(std::memory_order_relaxed) >> (dev & 63) & 1)) {
    CSS_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr_set.fetch_or(uint64_t(1) << (dev & 63), std::memory_order_relaxed);
  }
  AttnTcParams p;
  p.cu = cu_dev;
  p.rel_table = rel_table;
  p.rel_max = rel_max;
  p.rel_half = rel_half;
  p.n_seq = n_seq;
  p.nqb = (max_len + 127) / 128;
  p.ctx = ctx;
  const int items = n_seq * kHeads * p.nqb;
  const int grid = items < n_sm ? items : n_sm;
  attention_tc_kernel<<<grid, kAttnTcThreads, kAtSmem, st>>>(tq, tkv, p);
  CSS_LAUNCHED();
  return CSS_OK;
}

}  // namespace enc
}  // namespace css
