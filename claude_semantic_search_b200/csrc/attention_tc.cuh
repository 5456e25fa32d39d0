// attention_tc.cuh -- E3 on the 5th-generation tensor cores.
//
// One persistent CTA per SM walks work units (sequence, head); a unit's K and V tiles are staged
// once and reused by its query blocks of 128 rows ("items").  The keys of an item are split into
// two halves of up to 192 keys, each with its own score tile and softmax group (8 warps); the
// groups agree on one row maximum, so both halves accumulate into ONE output tile, which is
// double-buffered across items:
//
//   TMEM columns   S0 [0,192)   S1 [192,384)   O[item & 1] [384 + 64 (item & 1), +64)
//
//   warp 0        TMA producer: Q [128 x 64] per item, K / V [Lp x 64] per unit, bf16 tiles straight
//                 out of the packed qkv activation [T, 2304] (128-byte swizzle)
//   warp 1        MMA issuer:   S_g = Q K_g^T  (tcgen05.mma M128 x N<=192 x K16)
//                               O  += P_g[b] V_g[b]  (M128 x N64, V is the MN-major B operand);
//                               the next item's Q K^T is issued in front of the last round of P V
//   warps 2-5     epilogue: O / sum -> bf16 tile in shared memory -> one TMA store into ctx, overlapped with
//                 the next item's softmax
//   warps 6...     softmax group 0 then group 1, kAtGW warps each (4: one per TMEM lane quarter, a thread
//                 owns a row's 64 columns of every key block; 8: two per quarter, 32 columns each):
//                 pass 1  row maximum of the raw scores (tcgen05.ld), exchanged over all 16 warps
//                 pass 2  p = exp2(s*c + rel[j-i] - m), one 64-key block at a time, written as the
//                         bf16 A operand of P V into a double-buffered shared tile per group
// Keys beyond the sequence end do not exist in the packed layout (their tile rows hold the next
// sequence's tokens or TMA zero fill): their probabilities are forced to 0, which is what the
// reference's additive finfo.min mask produces.  Sequences longer than kAttnTcMaxLen take the
// mma.sync kernel (encoder_kernels.cuh).
#pragma once
#include "encoder_kernels.cuh"

namespace css {
namespace enc {

constexpr int kAttnTcMaxLen = 384;                 // two halves of 192 keys
constexpr int kAtGW = 4;                           // softmax warps per key half: 4 (64 columns per thread and block) or 8 (32)
constexpr int kAtHf = 8 / kAtGW;                   // 32-column pieces of a key block per thread
constexpr int kAttnTcThreads = (6 + 2 * kAtGW) * 32;  // warps: TMA, MMA, 4 epilogue, 2 x kAtGW softmax
constexpr int kAtHalfCols = 192;
constexpr int kAtQ = 128 * 128;                    // Q tile bytes
constexpr int kAtKV = kAttnTcMaxLen * 128;         // K / V tile bytes (max)
constexpr int kAtP = 128 * 128;                    // one P block (128 rows x 64 keys bf16)
constexpr int kAtRelStride = 2 * kAttnTcMaxLen;    // floats per buffer: bias(d) * log2e for d in [-383, 383] of ONE head
constexpr int kAtRel = 2 * kAtRelStride * 4;       // double-buffered by unit parity, refilled by the softmax threads
constexpr int kAtO = 128 * 128;                    // output tile staging (128 rows x 64 bf16, 128-byte swizzle) for the TMA store
constexpr int kAtCuMax = 768;                      // cu_seqlens entries cached in shared memory
constexpr int kAtXmax = 2 * 4 * 128 * 4;           // [item parity][group * 2 + column half][row] partial maxima
constexpr int kAtStat = 2 * 4 * 128 * 4;           // [item parity][group * 2 + column half][row] partial sums
constexpr int kAtSmem = kAtQ + 2 * kAtKV + 4 * kAtP + kAtO + kAtRel + 64 + kAtXmax + kAtStat + (kAtCuMax + 1) * 4 + 4 + 256 + 1024;
static_assert(kAtSmem <= 232448, "attention shared memory exceeds the 227 KB opt-in limit");
constexpr uint32_t kOCol = 2 * kAtHalfCols;        // O[0] at 384, O[1] at 448

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
  __nv_bfloat16* ctx;       // [T, 768]
  int trace_cta;            // the CTA whose timeline is recorded (CSS_ATTN_TRACE_CTA)
  long long* trace;         // optional [6 roles][128 items][8 tags] clock64 stamps of one CTA + every CTA's start / end time (profiling aid)
};

// Timeline stamps of ONE CTA (AttnTcParams::trace_cta): every other CTA carries a null pointer.
__device__ __forceinline__ void attn_trace(long long* trace, int role, uint32_t item, int tag) {
  if (trace != nullptr && item < 128u) trace[(role * 128 + item) * 8 + tag] = clock64();
}

// Every role walks the same item sequence: units u = blockIdx.x, +gridDim.x, ...; unit = (sequence,
// head); items = the unit's query blocks.
struct AttnWalk {
  const int32_t* cu;
  int n_units, stride;
  int u, qb, nqb;
  int h, t0, L;
  int nb0, nb1;             // 64-key blocks of the two halves (nb0 >= nb1, nb1 may be 0)
  __device__ __forceinline__ void init(const int32_t* cu_, int n_units_) {
    cu = cu_;
    n_units = n_units_;
    stride = gridDim.x;
    u = (int)blockIdx.x - stride;
    qb = 0;
    nqb = 0;
  }
  __device__ __forceinline__ bool advance() {
    if (++qb < nqb) return true;
    u += stride;
    if (u >= n_units) return false;
    const int s = u / kHeads;
    h = u - s * kHeads;
    t0 = cu[s];
    L = cu[s + 1] - t0;
    nqb = (L + 127) >> 7;
    const int nkb = (L + 63) >> 6;
    nb0 = (nkb + 1) >> 1;
    nb1 = nkb - nb0;
    qb = 0;
    return true;
  }
  __device__ __forceinline__ bool first_of_unit() const { return qb == 0; }
  __device__ __forceinline__ bool last_of_unit() const { return qb == nqb - 1; }
};

// State of the MMA-issuing thread.
struct AttnMma {
  uint64_t *q_full, *q_empty, *k_full, *k_empty, *v_full, *v_empty;
  uint64_t *s_full, *s_free, *o_full, *o_free, *p_full, *p_empty;
  uint32_t sQ, sK, sV, sP, tmem_base;
  long long* trace = nullptr;
  uint32_t n_qk = 0, n_pv = 0;          // items whose Q K^T / P V have been issued
  uint32_t n_unit_k = 0, n_unit_v = 0;  // K / V loads consumed
  uint32_t n_s0 = 0, n_s1 = 0;          // items whose S_g has been produced
  uint32_t n_p0 = 0, n_p1 = 0;          // P blocks of group g consumed

  // S_0 = Q K_0^T and S_1 = Q K_1^T of item x
  __device__ __forceinline__ void qk(const AttnWalk& x) {
    attn_trace(trace, 0, n_qk, 0);
    tc::mbar_wait(q_full, n_qk & 1);
    if (x.first_of_unit()) tc::mbar_wait(k_full, n_unit_k & 1);
    const uint64_t dq = tc::make_kmajor_sw128_desc(sQ);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int nb = g ? x.nb1 : x.nb0;
      uint32_t& n_s = g ? n_s1 : n_s0;
      if (nb > 0) {
        tc::mbar_wait(s_free + g, (n_s & 1) ^ 1);   // group g has finished reading the previous S_g
        ++n_s;
        tc::tc_fence_after();
        const uint64_t dk = tc::make_kmajor_sw128_desc(sK + (g ? x.nb0 : 0) * 8192);
        const uint32_t idesc = tc::make_idesc_bf16_f32(128, nb * 64);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_bf16(tmem_base + (uint32_t)(g * kAtHalfCols), dq + k * tc::kDescKStep, dk + k * tc::kDescKStep, idesc,
                        k != 0);
        tc::umma_commit(s_full + g);
        attn_trace(trace, 0, n_qk, 1 + g);
      }
    }
    tc::umma_commit(q_empty);   // Q (and after the unit's last item K) may be overwritten
    if (x.last_of_unit()) {
      tc::umma_commit(k_empty);
      ++n_unit_k;
    }
    ++n_qk;
  }

  // O[item & 1] (+)= P_g[lb] V_g[lb]
  template <int G>
  __device__ __forceinline__ void pv(const AttnWalk& x, int lb) {
    uint32_t& n_p = G ? n_p1 : n_p0;
    const uint32_t buf = n_p & 1;
    tc::mbar_wait(p_full + G * 2 + buf, (n_p >> 1) & 1);
    ++n_p;
    const uint32_t ob = n_pv & 1;
    if (G == 0 && lb == 0) {
      if (x.first_of_unit()) tc::mbar_wait(v_full, n_unit_v & 1);
      tc::mbar_wait(o_free + ob, ((n_pv >> 1) & 1) ^ 1);   // the epilogue has read the item that used O[ob] before
    }
    tc::tc_fence_after();
    attn_trace(trace, 1, n_pv, G * 4 + lb);
    const uint32_t idesc_pv = tc::make_idesc_bf16_f32(128, kHeadDim) | (1u << 16);  // B is MN-major
    const uint64_t dp = tc::make_kmajor_sw128_desc(sP + (G * 2 + buf) * kAtP);
    const uint64_t dv = make_mnmajor_sw128_desc(sV + ((G ? x.nb0 : 0) + lb) * 8192);
#pragma unroll
    for (int k = 0; k < 4; ++k)   // 16 keys per MMA: +32 B in P's rows, +16 rows (2048 B) in V
      tc::umma_bf16(tmem_base + kOCol + ob * kHeadDim, dp + k * tc::kDescKStep, dv + k * (2048 >> 4), idesc_pv,
                    (G | lb | k) != 0);
    tc::umma_commit(p_empty + G * 2 + buf);
    attn_trace(trace, 3, n_pv, G * 4 + lb);
  }

  __device__ __forceinline__ void run(AttnWalk cur) {
    bool have = cur.advance();
    if (have) qk(cur);
    while (have) {
      AttnWalk nxt = cur;
      const bool have_next = nxt.advance();
      // S_g is free as soon as group g has loaded its last block (before that block's
      // exponentials), so the next item's Q K^T goes in front of the last round of P V
      for (int lb = 0; lb < cur.nb0; ++lb) {
        if (lb == cur.nb0 - 1 && have_next) qk(nxt);
        pv<0>(cur, lb);
        if (lb < cur.nb1) pv<1>(cur, lb);
      }
      tc::umma_commit(o_full + (n_pv & 1));
      if (cur.last_of_unit()) {
        tc::umma_commit(v_empty);   // every P V of the unit has been issued
        ++n_unit_v;
      }
      ++n_pv;
      cur = nxt;
      have = have_next;
    }
  }
};

static __global__ void __launch_bounds__(kAttnTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_o, AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic, so the compiler keeps the shared address space (LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAtQ;
  uint8_t* sV = sK + kAtKV;
  uint8_t* sP = sV + kAtKV;                       // [group][buffer] blocks
  uint8_t* sO = sP + 4 * kAtP;                    // 1024-byte aligned (176 KB into the aligned window)
  float* sRel = reinterpret_cast<float*>(sO + kAtO);
  float* sRelMax = sRel + 2 * kAtRelStride;       // [16] per-head table maximum * log2e
  float* sXmax = sRelMax + 16;
  float* sStat = sXmax + kAtXmax / 4;
  int32_t* sCu = reinterpret_cast<int32_t*>(sStat + kAtStat / 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sCu) + ((kAtCuMax + 1) * 4 + 7) / 8 * 8);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* s_full = bars + 6;     // [2]  MMA -> softmax group
  uint64_t* s_free = bars + 8;     // [2]  softmax group (8 warps) -> MMA
  uint64_t* o_full = bars + 10;    // [2]  MMA -> epilogue, by item parity
  uint64_t* o_free = bars + 12;    // [2]  epilogue (4 warps) -> MMA
  uint64_t* p_full = bars + 14;    // [2][2] softmax group (8 warps) -> MMA
  uint64_t* p_empty = bars + 18;   // [2][2] MMA -> softmax group
  uint64_t* st_full = bars + 22;   // [2]  all 16 softmax warps -> epilogue: row sums of an item, by item parity
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_q);
    tc::prefetch_tmap(&tmap_kv);
    tc::prefetch_tmap(&tmap_o);
    tc::mbar_init(q_full, 1);
    tc::mbar_init(q_empty, 1);
    tc::mbar_init(k_full, 1);
    tc::mbar_init(k_empty, 1);
    tc::mbar_init(v_full, 1);
    tc::mbar_init(v_empty, 1);
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(s_full + g, 1);
      tc::mbar_init(s_free + g, kAtGW);
      tc::mbar_init(o_full + g, 1);
      tc::mbar_init(o_free + g, 4);
      tc::mbar_init(st_full + g, 2 * kAtGW);
      for (int b = 0; b < 2; ++b) {
        tc::mbar_init(p_full + g * 2 + b, kAtGW);
        tc::mbar_init(p_empty + g * 2 + b, 1);
      }
    }
    tc::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  // per-CTA constants: every head's bias window (x log2e) and the sequence offsets, so that no
  // item starts with a chain of dependent global loads
  if (threadIdx.x < kHeads) sRelMax[threadIdx.x] = p.rel_max[threadIdx.x] * 1.4426950408889634f;
  const bool cu_in_smem = p.n_seq <= kAtCuMax;
  if (cu_in_smem)
    for (int x = threadIdx.x; x <= p.n_seq; x += kAttnTcThreads) sCu[x] = p.cu[x];
  const int32_t* cu = cu_in_smem ? sCu : p.cu;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  AttnWalk w;
  w.init(cu, p.n_seq * kHeads);
  long long* const trace = (p.trace != nullptr && (int)blockIdx.x == p.trace_cta) ? p.trace : nullptr;
  if (p.trace != nullptr && threadIdx.x == 0) {   // start: wall clock (ns) of every CTA, SM clock of the traced CTA
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    if (blockIdx.x < 200) p.trace[5 * 1024 + 600 + 2 * blockIdx.x] = (long long)ns;
    if (trace != nullptr) {
      trace[127 * 8 + 6] = clock64();
      trace[127 * 8 + 7] = (long long)ns;
    }
  }

  if (warp == 0) {
    if (lane == 0) {
      // ============================ TMA producer ============================
      uint32_t it = 0, un = 0;
      while (w.advance()) {
        if (w.first_of_unit()) {
          const int nkb = w.nb0 + w.nb1;
          tc::mbar_wait_sleep(k_empty, (un & 1) ^ 1, 200);
          tc::mbar_expect_tx(k_full, nkb * 8192);
          for (int b = 0; b < nkb; ++b)
            tc::tma_load_2d(sK + b * 8192, &tmap_kv, k_full, kHidden + w.h * kHeadDim, w.t0 + b * 64, tc::kEvictNormal);
        }
        tc::mbar_wait_sleep(q_empty, (it & 1) ^ 1, 200);
        tc::mbar_expect_tx(q_full, kAtQ);
        tc::tma_load_2d(sQ, &tmap_q, q_full, w.h * kHeadDim, w.t0 + w.qb * 128, tc::kEvictNormal);
        if (w.first_of_unit()) {
          const int nkb = w.nb0 + w.nb1;
          tc::mbar_wait_sleep(v_empty, (un & 1) ^ 1, 200);
          tc::mbar_expect_tx(v_full, nkb * 8192);
          for (int b = 0; b < nkb; ++b)
            tc::tma_load_2d(sV + b * 8192, &tmap_kv, v_full, 2 * kHidden + w.h * kHeadDim, w.t0 + b * 64, tc::kEvictNormal);
          ++un;
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ============================ MMA issuer ============================
      AttnMma m;
      m.q_full = q_full; m.q_empty = q_empty; m.k_full = k_full; m.k_empty = k_empty;
      m.v_full = v_full; m.v_empty = v_empty; m.s_full = s_full; m.s_free = s_free;
      m.o_full = o_full; m.o_free = o_free; m.p_full = p_full; m.p_empty = p_empty;
      m.sQ = tc::smem_u32(sQ); m.sK = tc::smem_u32(sK); m.sV = tc::smem_u32(sV); m.sP = tc::smem_u32(sP);
      m.tmem_base = tmem_base;
      m.trace = trace;
      m.run(w);
    }
  } else if (warp < 6) {
    // ============================ epilogue ============================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;         // row of the query block == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kOCol;
    uint32_t e = 0;                            // items done
    while (w.advance()) {
      const uint32_t par = e & 1, ph = (e >> 1) & 1;
      const bool tr = lane == 0 && quarter == 0;
      if (tr) attn_trace(trace, 4, e, 0);
      tc::mbar_wait_sleep(st_full + par, ph, 300);
      const float* st = sStat + par * 4 * 128;
      const float inv = 1.f / ((st[r] + st[128 + r]) + (st[256 + r] + st[384 + r]));
      if (tr) attn_trace(trace, 4, e, 1);
      tc::mbar_wait_sleep(o_full + par, ph, 100);
      tc::tc_fence_after();
      if (tr) attn_trace(trace, 4, e, 2);
      uint32_t a[32], b[32];
      tc::tmem_ld_32x32(lane_addr + par * kHeadDim, a);
      tc::tmem_ld_32x32(lane_addr + par * kHeadDim + 32, b);
      tc::tmem_ld_wait();
      // O[par] is in registers (the sums were read above): the item after next may reuse both
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(o_free + par);
      // The tile leaves through shared memory and ONE bulk tensor store.  128 threads each writing
      // their own 128-byte row straight to global memory (32 scattered 16-byte requests per store
      // instruction) kept the LSU busy exactly while the softmax warps ran the first key block of the
      // next item: that block took 4200 cycles instead of 2600 (timeline with the stores removed).
      if (threadIdx.x == 64) tc::tma_store_wait_read();       // the previous item's store has read sO
      asm volatile("bar.sync 2, 128;" ::: "memory");
      uint8_t* orow = sO + r * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(a[q4 * 8 + 0]) * inv, __uint_as_float(a[q4 * 8 + 1]) * inv);
        o.y = pack_bf16(__uint_as_float(a[q4 * 8 + 2]) * inv, __uint_as_float(a[q4 * 8 + 3]) * inv);
        o.z = pack_bf16(__uint_as_float(a[q4 * 8 + 4]) * inv, __uint_as_float(a[q4 * 8 + 5]) * inv);
        o.w = pack_bf16(__uint_as_float(a[q4 * 8 + 6]) * inv, __uint_as_float(a[q4 * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + ((q4 ^ (r & 7)) << 4)) = o;
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(b[q4 * 8 + 0]) * inv, __uint_as_float(b[q4 * 8 + 1]) * inv);
        o.y = pack_bf16(__uint_as_float(b[q4 * 8 + 2]) * inv, __uint_as_float(b[q4 * 8 + 3]) * inv);
        o.z = pack_bf16(__uint_as_float(b[q4 * 8 + 4]) * inv, __uint_as_float(b[q4 * 8 + 5]) * inv);
        o.w = pack_bf16(__uint_as_float(b[q4 * 8 + 6]) * inv, __uint_as_float(b[q4 * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + (((4 + q4) ^ (r & 7)) << 4)) = o;
      }
      fence_proxy_async_smem();                                 // generic-proxy writes -> visible to the TMA engine
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const int q0 = w.qb * 128;
      if (q0 + 128 <= w.L) {
        if (threadIdx.x == 64) {
          tc::tma_store_2d(&tmap_o, sO, w.h * kHeadDim, w.t0 + q0);
          tc::tma_store_commit();
        }
      } else {
        // last, partial query block of a sequence: the rows past its end belong to the next sequence,
        // so only the real rows are written (8 lanes x 16 B = one row, 4 rows per instruction)
        const int et = threadIdx.x - 64;   // 0..127
        for (int rr = et >> 3; rr < w.L - q0; rr += 16) {
          const int c = et & 7;
          const uint4 o = *reinterpret_cast<const uint4*>(sO + rr * 128 + ((c ^ (rr & 7)) << 4));
          *reinterpret_cast<uint4*>(p.ctx + (size_t)(w.t0 + q0 + rr) * kHidden + w.h * kHeadDim + c * 8) = o;
        }
      }
      if (tr) attn_trace(trace, 4, e, 3);
      ++e;
    }
    if (threadIdx.x == 64) tc::tma_store_wait_all();   // the last tile has left shared memory
  } else {
    // ============================ softmax groups ============================
    const int g = (warp - 6) / kAtGW;                          // key half
    const int ch = kAtHf == 1 ? ((warp - 6) >> 2) & 1 : 0;     // first 32-column piece of a key block owned by this warp
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                         // row of the query block == TMEM lane
    const int slot = g * 2 + ch;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(g * kAtHalfCols + ch * 32);
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kScale = 0.125f * kLog2e;  // 1/sqrt(64) * log2(e)
    uint32_t n = 0, ng = 0, np = 0;            // items / items with keys in this half / P blocks of this group
    uint32_t nu = 0, nu_next = 0;              // units started
    while (w.advance()) {
      const int nb = g ? w.nb1 : w.nb0;
      const int L = w.L;
      const int key0 = (g ? w.nb0 : 0) * 64 + ch * 32;   // first key of this thread's columns in block 0
      const int i = min(w.qb * 128 + r, L - 1);          // rows past the end mirror the last row, never stored
      // bias window of this unit's head (x log2e): refilled by the softmax threads at the unit's first item;
      // the barrier between pass 1 and pass 2 publishes it, and the buffer was last read two units ago
      float* relbuf = sRel + (nu & 1) * kAtRelStride;
      if (w.first_of_unit()) {
        const float* src = p.rel_table + (size_t)w.h * (2 * p.rel_half + 1) + p.rel_half;
        for (int x = threadIdx.x - 6 * 32; x < 2 * kAttnTcMaxLen - 1; x += 2 * kAtGW * 32) {
          const int d = max(-p.rel_half, min(p.rel_half, x - (kAttnTcMaxLen - 1)));   // key - query
          relbuf[x] = __ldg(src + d) * 1.4426950408889634f;
        }
      }
      if (w.last_of_unit()) ++nu_next;
      const float* rel_i = relbuf + (kAttnTcMaxLen - 1) - i;   // rel_i[j] = bias(j - i) * log2e
      const uint32_t par = n & 1;
      const bool tr = lane == 0 && ch == 0 && quarter == 0 && g == 0;
      if (tr) attn_trace(trace, 2 + g, n, 0);
      uint32_t v[kAtHf][32];
      // ---- pass 1: maximum of the raw scores over this thread's real keys ----
      float mx = -INFINITY;
      if (nb > 0) {
        tc::mbar_wait_spin(s_full + g, ng & 1);
        tc::tc_fence_after();
        if (tr) attn_trace(trace, 2 + g, n, 1);
        // blocks in descending order: block 0's scores are still in registers when pass 2 starts
        for (int lb = nb - 1; lb >= 0; --lb) {
          const int c0 = key0 + lb * 64;
          if (c0 >= L) continue;
#pragma unroll
          for (int hf = 0; hf < kAtHf; ++hf)
            if (c0 + hf * 32 < L) tc::tmem_ld_32x32(lane_addr + (uint32_t)(lb * 64 + hf * 32), v[hf]);
          tc::tmem_ld_wait();
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int hf = 0; hf < kAtHf; ++hf) {
            const int c = c0 + hf * 32;
            if (c + 32 <= L) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[hf][j]));
            } else if (c < L) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c + j < L) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[hf][j]));
            }
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        }
      }
      // one row maximum for the whole item: both halves then accumulate into the same O
      float* xm = sXmax + par * 4 * 128;
      xm[slot * 128 + r] = mx;
      if (kAtHf == 2) xm[(slot + 1) * 128 + r] = mx;
      if (tr) attn_trace(trace, 2 + g, n, 2);
      asm volatile("bar.sync 1, %0;" ::"n"(2 * kAtGW * 32) : "memory");
      if (tr) attn_trace(trace, 2 + g, n, 3);
      mx = fmaxf(fmaxf(xm[r], xm[128 + r]), fmaxf(xm[256 + r], xm[384 + r]));
      // upper bound of the row maximum of (s/8 + bias) in the log2 domain
      const float m_hat = fmaf(mx, kScale, sRelMax[w.h]);
      // ---- pass 2: probabilities, one 64-key block at a time, as the A operand of P V ----
      // Packed fp32x2 arithmetic (two keys per instruction).  Block 0's scores are left over from
      // pass 1; the scores of block lb+1 are requested as soon as block lb's exponentials are done,
      // so TMEM latency stays off the critical path.
      uint64_t sum2 = 0ull;
      const uint64_t scale2 = f32x2_pack(kScale, kScale);
      const uint64_t nm2 = f32x2_pack(-m_hat, -m_hat);
      for (int lb = 0; lb < nb; ++lb, ++np) {
        const uint32_t buf = np & 1;
        const int c0 = key0 + lb * 64;
        tc::mbar_wait_spin(p_empty + g * 2 + buf, ((np >> 1) & 1) ^ 1);
        uint8_t* prow = sP + (g * 2 + buf) * kAtP + r * 128;
        tc::tmem_ld_wait();
#pragma unroll
        for (int hf = 0; hf < kAtHf; ++hf) {
          const int c = c0 + hf * 32;
          const int piece = ch + hf;
          if (c + 32 <= L) {   // warp-uniform
            const float* rel_c = rel_i + c;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t pk[4];
#pragma unroll
              for (int j2 = 0; j2 < 4; ++j2) {
                const int j = q4 * 4 + j2;
                const uint64_t t = f32x2_fma(f32x2_pack(__uint_as_float(v[hf][2 * j]), __uint_as_float(v[hf][2 * j + 1])),
                                             scale2, f32x2_add(f32x2_pack(rel_c[2 * j], rel_c[2 * j + 1]), nm2));
                float t0, t1;
                f32x2_unpack(t, t0, t1);
                const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
                sum2 = f32x2_add(sum2, f32x2_pack(e0, e1));
                pk[j2] = pack_bf16(e0, e1);
              }
              *reinterpret_cast<uint4*>(prow + (((piece * 4 + q4) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else if (c < L) {
            const float* rel_c = rel_i + c;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t pk[4];
#pragma unroll
              for (int j2 = 0; j2 < 4; ++j2) {
                const int j = q4 * 4 + j2;
                const uint64_t t = f32x2_fma(f32x2_pack(__uint_as_float(v[hf][2 * j]), __uint_as_float(v[hf][2 * j + 1])),
                                             scale2, f32x2_add(f32x2_pack(rel_c[2 * j], rel_c[2 * j + 1]), nm2));
                float t0, t1;
                f32x2_unpack(t, t0, t1);
                const float e0 = (c + 2 * j < L) ? ex2_approx(t0) : 0.f;
                const float e1 = (c + 2 * j + 1 < L) ? ex2_approx(t1) : 0.f;
                sum2 = f32x2_add(sum2, f32x2_pack(e0, e1));
                pk[j2] = pack_bf16(e0, e1);
              }
              *reinterpret_cast<uint4*>(prow + (((piece * 4 + q4) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<uint4*>(prow + (((piece * 4 + q4) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        if (lb == nb - 1) {   // the last read of S_g has landed: the next item's Q K_g^T may overwrite it
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(s_free + g);
        } else {
#pragma unroll
          for (int hf = 0; hf < kAtHf; ++hf)
            if (c0 + 64 + hf * 32 < L) tc::tmem_ld_32x32(lane_addr + (uint32_t)((lb + 1) * 64 + hf * 32), v[hf]);
        }
        fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(p_full + g * 2 + buf);
        if (tr) attn_trace(trace, 2 + g, n, 4 + lb);
      }
      float sum_lo, sum_hi;
      f32x2_unpack(sum2, sum_lo, sum_hi);
      // row sums for the epilogue warps (a half without keys contributes 0); the slot was last used
      // by item n - 2, which the epilogue has finished once it has released that item's O buffer
      tc::mbar_wait_spin(o_free + par, ((n >> 1) & 1) ^ 1);
      sStat[(par * 4 + slot) * 128 + r] = sum_lo + sum_hi;
      if (kAtHf == 2) sStat[(par * 4 + slot + 1) * 128 + r] = 0.f;
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(st_full + par);
      if (nb > 0) ++ng;
      ++n;
      nu = nu_next;
    }
  }
  __syncwarp();
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
  if (p.trace != nullptr && threadIdx.x == 0) {   // end
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    if (blockIdx.x < 200) p.trace[5 * 1024 + 600 + 2 * blockIdx.x + 1] = (long long)ns;
    if (trace != nullptr) {
      trace[(128 + 127) * 8 + 6] = clock64();
      trace[(128 + 127) * 8 + 7] = (long long)ns;
    }
  }
}

// Host launcher.  qkv: [T, 2304] bf16 packed (q | k | v).
static int attention_tc_launch(const __nv_bfloat16* qkv, int T, const int32_t* cu_dev, int n_seq, int max_len,
                               const float* rel_table, const float* rel_max, int rel_half, __nv_bfloat16* ctx,
                               int n_sm, cudaStream_t st, long long* trace = nullptr) {
  CSS_REQUIRE(max_len <= kAttnTcMaxLen, "attention_tc: sequence of %d tokens exceeds %d", max_len, kAttnTcMaxLen);
  CUtensorMap tq, tkv, to;
  CSS_CHECK(encode_tmap_bf16_2d(&tq, qkv, (uint64_t)T, 3 * kHidden, 3 * kHidden, 128, 64));
  CSS_CHECK(encode_tmap_bf16_2d(&tkv, qkv, (uint64_t)T, 3 * kHidden, 3 * kHidden, 64, 64));
  CSS_CHECK(encode_tmap_bf16_2d(&to, ctx, (uint64_t)T, kHidden, kHidden, 128, 64));
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
  p.ctx = ctx;
  p.trace = trace;
  const char* cta_env = trace ? getenv("CSS_ATTN_TRACE_CTA") : nullptr;
  p.trace_cta = cta_env ? atoi(cta_env) : 0;
  const int units = n_seq * kHeads;
  const int grid = units < n_sm ? units : n_sm;
  attention_tc_kernel<<<grid, kAttnTcThreads, kAtSmem, st>>>(tq, tkv, to, p);
  CSS_LAUNCHED();
  return CSS_OK;
}

}  // namespace enc
}  // namespace css
