// index_kernels.cuh -- device kernels of the flat index:
//   S1  row normalise on add (+ bf16 shadow copy)        src/storage.py:347-350
//   S2  streaming fp32 scan with fused warp/block/grid top-k   src/storage.py:436 (faiss IndexFlat.search, small nq)
//   S4  filter predicate -> row bitmask                  src/storage.py:508-543
//   S5  merge of per-shard top-k lists
//
// Roofline: S2 is HBM-bound.  Algorithmic bytes = ntotal * dim * 4 per query
// (3072 B per 768-d row); the kernel reads every passing
// row exactly once with 128-bit streaming loads and keeps the top-k in
// registers, so nothing but k results per block is ever written.
#pragma once
#include <type_traits>

#include "css_common.cuh"

namespace css {

constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kRowsPerUnit = 8;      // one mask byte
constexpr int kRowsPerGroup = 4;     // rows in flight per warp iteration
constexpr int kMergeCap = 4096;      // entries sorted per round by the last block

struct KeyId {
  float key;
  int id;
};

// ------------------------------------------------------------------------
// Bitonic sort (descending by `better`) of n = power-of-two entries in smem.
// ------------------------------------------------------------------------
struct KeyId64 {
  float key;
  int pad;
  long long id;
};
__device__ __forceinline__ bool better(const KeyId& a, const KeyId& b) {
  return better(a.key, a.id, b.key, b.id);
}
__device__ __forceinline__ bool better(const KeyId64& a, const KeyId64& b) {
  return (a.key > b.key) || (a.key == b.key && a.id < b.id);
}

template <typename E>
__device__ __forceinline__ void bitonic_sort_desc(E* s, int n, int tid, int nthreads) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (n >> 1); t += nthreads) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        E a = s[lo], b = s[hi];
        bool a_better = better(a, b);
        if (a_better != desc) {
          s[lo] = b;
          s[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------
// Per-warp top-k held in registers: lane l owns KPL slots.
// ------------------------------------------------------------------------
template <int KPL>
struct WarpTopK {
  float key[KPL];
  int id[KPL];
  float lw_key;  // this lane's worst slot
  int lw_id;
  int lw_slot;
  float thr_key;  // warp-wide worst entry (uniform)
  int thr_id;

  __device__ __forceinline__ void init(int lane, int k) {
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      // Slots beyond k are permanently "full of +inf" so they are never the worst
      // and never selected: the list then holds exactly k live entries.
      bool live = (lane * KPL + s) < k;
      key[s] = live ? -INFINITY : INFINITY;
      id[s] = live ? kEmptyId : -1;
    }
    recompute();
  }

  __device__ __forceinline__ void recompute() {
    lw_key = key[0];
    lw_id = id[0];
    lw_slot = 0;
#pragma unroll
    for (int s = 1; s < KPL; ++s) {
      if (better(lw_key, lw_id, key[s], id[s])) {
        lw_key = key[s];
        lw_id = id[s];
        lw_slot = s;
      }
    }
    float wk = lw_key;
    int wi = lw_id;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ok = __shfl_xor_sync(0xffffffffu, wk, o);
      int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (better(wk, wi, ok, oi)) {
        wk = ok;
        wi = oi;
      }
    }
    thr_key = wk;
    thr_id = wi;
  }

  __device__ __forceinline__ void maybe_compact(int) {}
  // Warp-uniform call: (k_, i_) identical in all lanes.
  __device__ __forceinline__ void consider(float k_, int i_, int lane) {
    if (!better(k_, i_, thr_key, thr_id)) return;
    unsigned owners = __ballot_sync(0xffffffffu, lw_key == thr_key && lw_id == thr_id);
    int owner = __ffs(owners) - 1;
    if (lane == owner) {
#pragma unroll
      for (int s = 0; s < KPL; ++s) {
        if (s == lw_slot) {
          key[s] = k_;
          id[s] = i_;
        }
      }
    }
    recompute();
  }
};

// ------------------------------------------------------------------------
// Per-warp top-32 of the shadow sweeps (two-phase scan, 32-entry block lists): a sorted list, one entry per
// lane, plus a 32-entry buffer of pending candidates, one per lane.  A row that beats the 32nd key so far costs a
// predicated register write; every 32 such rows the buffer is sorted and merged into the list across the lanes.
// WarpTopK::consider pays a ballot and a 10-shuffle re-scan of the list per accepted row -- a quarter of all rows
// while a warp has seen only a few hundred (1 M rows over 2368 warps), 130 instructions per 8-row unit.  Ties at
// the 32nd key are dropped: the proof of two_phase_finish only needs every row outside a list to score no higher
// than the list's last entry.  Same members as WarpTopK<1> for the block merge (after flush()).
// ------------------------------------------------------------------------
struct WarpBufTop32 {
  float key[1];
  int id[1];
  float bkey;
  int bid;
  int nb;       // pending entries: lanes [0, nb)
  float thr;    // 32nd best key so far (-inf until 32 rows were kept)

  __device__ __forceinline__ void init(int, int) {
    key[0] = -INFINITY;
    id[0] = kEmptyId;
    bkey = -INFINITY;
    bid = kEmptyId;
    nb = 0;
    thr = -INFINITY;
  }
  __device__ __forceinline__ static void exchange(float& k, int& i, const int lane, const int o, const bool keep_better) {
    const float ok = __shfl_xor_sync(0xffffffffu, k, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    const bool other_better = better(ok, oi, k, i);
    if (other_better == keep_better) {
      k = ok;
      i = oi;
    }
  }
  __device__ __forceinline__ void compact(const int lane) {
    if (lane >= nb) {
      bkey = -INFINITY;
      bid = kEmptyId;
    }
    // bitonic sort of the buffer, best entry in lane 0
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int o = size >> 1; o > 0; o >>= 1) {
        const bool desc = (lane & size) == 0 || size == 32;
        const bool lower = (lane & o) == 0;
        exchange(bkey, bid, lane, o, lower == desc);
      }
    }
    // list[i] = better(list[i], buffer[31 - i]): the 32 best of both, as a bitonic sequence; then a bitonic merge
    const float rk = __shfl_sync(0xffffffffu, bkey, 31 - lane);
    const int ri = __shfl_sync(0xffffffffu, bid, 31 - lane);
    if (better(rk, ri, key[0], id[0])) {
      key[0] = rk;
      id[0] = ri;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) exchange(key[0], id[0], lane, o, (lane & o) == 0);
    thr = __shfl_sync(0xffffffffu, key[0], 31);
    nb = 0;
  }
  // Warp-uniform (k_, i_).  The caller runs maybe_compact() at least once per 8 calls (the buffer keeps 8 free slots).
  __device__ __forceinline__ void consider(const float k_, const int i_, const int lane) {
    if (k_ > thr) {
      if (lane == nb) {
        bkey = k_;
        bid = i_;
      }
      ++nb;
    }
  }
  __device__ __forceinline__ void maybe_compact(const int lane) {
    if (nb > 24) compact(lane);
  }
  __device__ __forceinline__ void flush(const int lane) {
    if (nb > 0) compact(lane);
  }
};

// Local row -> returned id.  One shard of a single-process multi-device index holds the row blocks
// b with b % ndev == shard (block = 2^shift rows, block-cyclic over the devices, so that ids stay dense
// and append-only like faiss's whatever the number of devices); ndev <= 1: id = offset + local row.
struct IdMap {
  long long offset;
  int shift;
  int ndev;
  int shard;
};
__device__ __forceinline__ long long map_id(const IdMap& m, int local) {
  if (m.ndev <= 1) return m.offset + local;
  const long long l = local;
  return m.offset + ((((l >> m.shift) * m.ndev + m.shard) << m.shift) | (l & ((1ll << m.shift) - 1)));
}

// ------------------------------------------------------------------------
// Result exchange between the shards of one search (one exchange step), fused into the
// kernel that produces the local top-k: the CTA that finishes a query's local list stores it into
// every peer's receive area over NVLink (peer-mapped memory: cudaDeviceEnablePeerAccess in one
// process, CUDA IPC between the ranks of a torchrun job), raises a per-(query, source) flag there,
// waits for the flags of all sources in its OWN memory and merges the n_ranks lists -- no NCCL
// call, no extra launch.  Receive areas and flags are double-buffered by the parity of `epoch`
// (a rank cannot be two searches ahead of a peer: it needs that peer's list to finish a search).
// ------------------------------------------------------------------------
constexpr int kMaxRanks = CSS_MAX_RANKS;
struct ExEntry {
  float key;
  int pad;
  long long id;   // global id, -1 = unfilled slot
};
struct ExchangeDev {
  int n_ranks;   // <= 1: no exchange
  int rank;
  unsigned epoch;
  int max_nq;
  ExEntry* slots[kMaxRanks];    // slots[r]: receive area of rank r, [2][max_nq][n_ranks][CSS_MAX_K]
  unsigned* flags[kMaxRanks];   // flags[r]: flags of rank r, [2][max_nq][n_ranks]
  int* status;                  // local; 1 = a peer's list did not arrive (timeout)
  int deferred;                 // 1: the producing CTA only publishes; the lists are awaited and merged by the
                                //    fp32-fallback launch that follows in the stream (two-phase scan, nq > 1)
  int nq;                       // queries of this search (deferred merge loop)
  int no_pdl;                   // 1: several shards of this process share a GPU (tests): a dependent launch parked on the
                                //    SMs would starve the peer shard's kernel this one is waiting for
};

struct ScanParams {
  const float* x;        // [n, d]
  const __nv_bfloat16* xb;  // [n, d] bf16 shadow rows (two-phase scan only)
  const int8_t* xq;      // [n, 768] int8 shadow rows (two-phase scan, first tier; nullable)
  const float* xs;       // [n] scale of each int8 row: x ~= xs[row] * xq[row]
  int64_t n;
  int d;
  const float* q;        // [nq, d]
  const uint32_t* mask;  // nullable bitmask over rows
  int mask_dense;        // 1: the mask is the alive bits alone (few cleared): the int8 tier sweeps every row and drops the dead
  int k;                 // entries kept per warp / per block list
  int k_out;             // entries of the result (== k except in the two-phase scan, where k is the list length)
  KeyId* part;           // [nq][gridDim.x][k]
  float* part_exact;     // [nq][gridDim.x][k] two-phase scan: fp32 scores of the list entries (written by the block that owns the list)
  unsigned int* ticket;  // [nq], zero on entry, zero on exit
  unsigned int* cursor;  // [nq], zero on entry, zero on exit: int8 sweep, next pair of dynamically dealt units
  IdMap idmap;
  float* D;              // [nq, k_out]
  int64_t* I;            // [nq, k_out]
  const int* qlist;      // nullable: explicit list of query indices to scan (device)
  const int* qcount;     // number of entries of qlist (device)
  int no_merge;          // 1: stop after the per-block lists (phase 1 timed alone)
  int interleave;        // 1: dense bf16 sweep walks 8-row units block-cyclically over the grid
  int* zero_on_entry;    // nullable: one int cleared by the first thread of the grid (the overflow count)
  int pdl_wait;          // 1: launched as a programmatic dependent of the previous kernel (griddepcontrol.wait first)
  // two-phase scan
  const float* max_norm; // largest stored row norm
  const float* max_err;  // largest ||x - bf16(x)|| over the stored rows
  const float* max_err8; // largest ||x - xs * xq|| over the stored rows (+inf once a non-finite row was stored)
  int* ovf_list;         // queries handed to the fp32 scan
  int* ovf_count;
  unsigned* stats_dev;   // [2] two-phase queries, unproven queries (device counters)
  volatile unsigned* stats_host;  // nullable: the same two counters mirrored into mapped host memory
  long long* trace;      // nullable (css_debug_scan_trace): [gridDim.x + 1][8] globaltimer stamps, see scan_stamp
  unsigned* done_flag;   // nullable (single-query host calls): mapped host word that receives done_seq << 1 | unproven
  unsigned done_seq;     //   once the query's result (D / I may be mapped host memory too) has been written
  ExchangeDev ex;
};

// Dot products (or negated squared distances) of up to 4 rows against the query.
template <int METRIC, bool D768>
__device__ __forceinline__ void score_rows(const ScanParams& p, const float4* qreg,
                                           const float* q_s, const int64_t (&r)[kRowsPerGroup],
                                           int lane, float (&acc)[kRowsPerGroup]) {
  if constexpr (D768) {
    float4 v[kRowsPerGroup][6];
#pragma unroll
    for (int i = 0; i < kRowsPerGroup; ++i) {
      const float4* row = reinterpret_cast<const float4*>(p.x + r[i] * 768);
#pragma unroll
      for (int j = 0; j < 6; ++j) v[i][j] = ld_stream_f4(row + j * 32 + lane);
    }
#pragma unroll
    for (int i = 0; i < kRowsPerGroup; ++i) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if constexpr (METRIC == CSS_METRIC_INNER_PRODUCT) {
          a = fmaf(v[i][j].x, qreg[j].x, a);
          a = fmaf(v[i][j].y, qreg[j].y, a);
          a = fmaf(v[i][j].z, qreg[j].z, a);
          a = fmaf(v[i][j].w, qreg[j].w, a);
        } else {
          float t;
          t = v[i][j].x - qreg[j].x; a = fmaf(t, t, a);
          t = v[i][j].y - qreg[j].y; a = fmaf(t, t, a);
          t = v[i][j].z - qreg[j].z; a = fmaf(t, t, a);
          t = v[i][j].w - qreg[j].w; a = fmaf(t, t, a);
        }
      }
      acc[i] = a;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kRowsPerGroup; ++i) acc[i] = 0.f;
    for (int j = lane; j < p.d; j += 32) {
      float qj = q_s[j];
#pragma unroll
      for (int i = 0; i < kRowsPerGroup; ++i) {
        float xv = __ldg(p.x + r[i] * (int64_t)p.d + j);
        if constexpr (METRIC == CSS_METRIC_INNER_PRODUCT) {
          acc[i] = fmaf(xv, qj, acc[i]);
        } else {
          float t = xv - qj;
          acc[i] = fmaf(t, t, acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kRowsPerGroup; ++i) {
    acc[i] = warp_sum(acc[i]);
    if constexpr (METRIC == CSS_METRIC_L2) acc[i] = -acc[i];
  }
}

// Two-phase scan, phase 1: inner products of 8 rows against the fp32 query, the rows read
// from the bf16 shadow copy (1536 B per 768-d row: 3 x 16 B per lane, lane l holds elements
// 256 j + 8 l .. + 8 of the row; qreg is loaded in the same layout).
__device__ __forceinline__ void score_rows_bf16(const ScanParams& p, const float4* qreg, const int64_t (&r)[kRowsPerUnit],
                                                int lane, float (&acc)[kRowsPerUnit]) {
  uint4 v[kRowsPerUnit][3];
#pragma unroll
  for (int i = 0; i < kRowsPerUnit; ++i) {
    const uint4* row = reinterpret_cast<const uint4*>(p.xb + r[i] * 768);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v[i][j].x), "=r"(v[i][j].y), "=r"(v[i][j].z), "=r"(v[i][j].w)
                   : "l"(row + j * 32 + lane));
    }
  }
#pragma unroll
  for (int i = 0; i < kRowsPerUnit; ++i) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 qa = qreg[2 * j], qb = qreg[2 * j + 1];
      a = fmaf(__uint_as_float(v[i][j].x << 16), qa.x, a);
      a = fmaf(__uint_as_float(v[i][j].x & 0xffff0000u), qa.y, a);
      a = fmaf(__uint_as_float(v[i][j].y << 16), qa.z, a);
      a = fmaf(__uint_as_float(v[i][j].y & 0xffff0000u), qa.w, a);
      a = fmaf(__uint_as_float(v[i][j].z << 16), qb.x, a);
      a = fmaf(__uint_as_float(v[i][j].z & 0xffff0000u), qb.y, a);
      a = fmaf(__uint_as_float(v[i][j].w << 16), qb.z, a);
      a = fmaf(__uint_as_float(v[i][j].w & 0xffff0000u), qb.w, a);
    }
    acc[i] = a;
  }
#pragma unroll
  for (int i = 0; i < kRowsPerUnit; ++i) acc[i] = warp_sum(acc[i]);
}

// ------------------------------------------------------------------------
// Two-phase scan, first tier: the rows are read from the int8 shadow copy (768 B per 768-d row + a 4-byte
// scale: x ~= xs[row] * xq[row], codes in [-127, 127], scale = max|x| / 127 per row), a QUARTER of the fp32
// bytes.  The query is split into two int8 vectors, q ~= a1 q1 + a2 q2 with a2 = a1 / 254 (q2 quantises the
// residual of q1), so its own quantisation error is ~2^-15 |q|_inf per element and the integer dot products
// I1 = q1.xq, I2 = q2.xq (dp4a, exact) give   x^.q^ = xs a1 (I1 + I2 / 254).
//   |x.q - x^.q^| <= ||x - x^|| ||q|| + ||x^|| ||q - q^||  <=  max_err8 ||q|| + (max_norm + max_err8) ||q - q^||
// (max_err8: largest quantisation-error norm of any stored row, tracked exactly at add time; ||q - q^|| is
// computed exactly below), plus a few float roundings covered by the 4e-6 ||q|| max_norm of two_phase_finish.
// Layout: a HALF-warp owns a row -- lane (hl = lane & 15) holds elements 256 j + 16 hl .. + 16, j = 0..2, of
// the row of its half (3 x 16 B loads per lane and row), the query codes live in registers in the same layout.
// ------------------------------------------------------------------------
constexpr float kInv254 = 1.f / 254.f;
constexpr int kI8UnitBytes = kRowsPerUnit * 768;            // one 8-row unit of int8 rows: 6 KB, contiguous
constexpr int kI8Stages = 2;                                // per warp: two units in flight (TMA bulk copies)
constexpr int kI8RingBytes = kScanWarps * kI8Stages * kI8UnitBytes;                  // 192 KB per CTA
constexpr int kI8SmemBytes = kI8RingBytes + kScanWarps * kI8Stages * 8;              // + one mbarrier per stage

// ---- mbarrier / bulk-copy primitives of the int8 sweep (one producer lane and 32 consumers per warp) ----
__device__ __forceinline__ uint32_t sb_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sb_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool sb_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(sb_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded: a pipeline bug must surface as a trapped kernel, never as a GPU that spins for ever.
__device__ __forceinline__ void sb_mbar_wait(uint64_t* bar, uint32_t parity) {
  if (sb_mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!sb_mbar_try_wait(bar, parity)) {
    if ((++spins & 4095u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ll) {   // ~2 s
        printf("css: int8 sweep: bulk copy did not arrive (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}
// 1-D bulk copy global -> shared through the TMA unit; completion (bytes) is signalled on `bar`.
__device__ __forceinline__ void sb_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sb_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(sb_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ int pack_i8x4(float c0, float c1, float c2, float c3) {
  return (__float2int_rn(c0) & 0xff) | ((__float2int_rn(c1) & 0xff) << 8) | ((__float2int_rn(c2) & 0xff) << 16) |
         ((__float2int_rn(c3) & 0xff) << 24);
}

// Quantise the query (every warp for itself).  Out: codes, a1, qn >= ||q||, dq >= ||q - (a1 q1 + a2 q2)||.
__device__ __forceinline__ void quantize_query_i8(const float* q, const int lane, int (&q1)[12], int (&q2)[12], float& a1,
                                                  float& qn, float& dq) {
  const int hl = lane & 15;
  float4 f[12];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int w = 0; w < 4; ++w) f[j * 4 + w] = __ldg(reinterpret_cast<const float4*>(q + j * 256 + hl * 16) + w);
  float amax = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(f[i].x), fabsf(f[i].y)), fmaxf(fabsf(f[i].z), fabsf(f[i].w))));
    ss = fmaf(f[i].x, f[i].x, ss);
    ss = fmaf(f[i].y, f[i].y, ss);
    ss = fmaf(f[i].z, f[i].z, ss);
    ss = fmaf(f[i].w, f[i].w, ss);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {   // both halves of the warp hold the same values
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  qn = sqrtf(ss) * 1.00001f;
  a1 = amax / 127.f;
  const float a2 = a1 * kInv254;
  const float inv1 = amax > 0.f ? 127.f / amax : 0.f;
  const float inv2 = a2 > 0.f ? 1.f / a2 : 0.f;
  float d2 = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float v[4] = {f[i].x, f[i].y, f[i].z, f[i].w};
    float c1[4], c2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      c1[e] = fminf(fmaxf(rintf(v[e] * inv1), -127.f), 127.f);
      const float r = fmaf(-a1, c1[e], v[e]);
      c2[e] = fminf(fmaxf(rintf(r * inv2), -127.f), 127.f);
      const float r2 = fmaf(-a2, c2[e], r);
      d2 = fmaf(r2, r2, d2);
    }
    q1[i] = pack_i8x4(c1[0], c1[1], c1[2], c1[3]);
    q2[i] = pack_i8x4(c2[0], c2[1], c2[2], c2[3]);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  dq = sqrtf(d2) * 1.001f;
}

// This lane's share of x^.q^ for one row (48 elements), already multiplied by the row's scale and a1.
__device__ __forceinline__ float dot_i8(const uint4 (&v)[3], const int (&q1)[12], const int (&q2)[12], const float scale) {
  int i1 = 0, i2 = 0, j1 = 0, j2 = 0;   // two chains per code vector: six dependent dp4a instead of twelve
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    i1 = __dp4a((int)v[j].x, q1[j * 4 + 0], i1);
    i2 = __dp4a((int)v[j].x, q2[j * 4 + 0], i2);
    j1 = __dp4a((int)v[j].y, q1[j * 4 + 1], j1);
    j2 = __dp4a((int)v[j].y, q2[j * 4 + 1], j2);
    i1 = __dp4a((int)v[j].z, q1[j * 4 + 2], i1);
    i2 = __dp4a((int)v[j].z, q2[j * 4 + 2], i2);
    j1 = __dp4a((int)v[j].w, q1[j * 4 + 3], j1);
    j2 = __dp4a((int)v[j].w, q2[j * 4 + 3], j2);
  }
  return fmaf((float)(i2 + j2), kInv254, (float)(i1 + j1)) * scale;
}

// Sum over the 16 lanes of each half; acc0 = the first half's row, acc1 = the second half's (warp-uniform).
__device__ __forceinline__ void half_sums(float f, float& acc0, float& acc1) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
  acc0 = __shfl_sync(0xffffffffu, f, 0);
  acc1 = __shfl_sync(0xffffffffu, f, 16);
}

// ------------------------------------------------------------------------
// Result of one query: s[0..k_out) is the local list, best first (unfilled slots carry kEmptyId).
// Without an exchange it is written to D/I; with one it is published to every rank, the lists of
// all ranks are awaited and merged (see ExchangeDev), and the merged list is written -- identical
// on every rank.  Called by all threads of one CTA; s must have room for n_ranks * k_out 16-byte
// entries (8 x 128 x 16 B = 16 KB of the 32 KB list buffer).
// ------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int METRIC>
__device__ __forceinline__ void write_result(const ScanParams& p, const int qi, const int i, const float key, const long long gid,
                                             const bool empty) {
  float dval;
  if constexpr (METRIC == CSS_METRIC_INNER_PRODUCT) dval = empty ? -FLT_MAX : key;
  else dval = empty ? FLT_MAX : -key;
  p.D[(int64_t)qi * p.k_out + i] = dval;
  p.I[(int64_t)qi * p.k_out + i] = empty ? (int64_t)-1 : (int64_t)gid;
}

// Wait for the lists of all ranks for query qi (own memory; bounded) and merge them into D/I.
// `m`: shared memory for n_ranks * k_out 16-byte entries.  All threads of the CTA.
template <int METRIC>
__device__ __forceinline__ void await_and_merge(const ScanParams& p, const int qi, KeyId64* m, const int tid,
                                                const int nthreads) {
  const int k = p.k_out, R = p.ex.n_ranks;
  const size_t cell = ((size_t)(p.ex.epoch & 1u) * p.ex.max_nq + qi) * R;   // [parity][query][source]
  if (tid < R) {
    const unsigned* f = p.ex.flags[p.ex.rank] + cell + tid;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(f) != p.ex.epoch) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > 10000000000ull) {   // 10 s: a missing peer is an error, not a hang
        *p.ex.status = 1;
        break;
      }
    }
  }
  __syncthreads();
  const ExEntry* mine = p.ex.slots[p.ex.rank] + cell * CSS_MAX_K;
  const int total = R * k;
  if (total <= 1024) {
    // rank counting instead of a sorting network: two barriers, and entry i goes straight to its output slot
    for (int i = tid; i < total; i += nthreads) {
      const int r = i / k, j = i - r * k;
      const int4 v = __ldcg(reinterpret_cast<const int4*>(mine + (size_t)r * CSS_MAX_K + j));
      KeyId64 e;
      e.key = __int_as_float(v.x);
      e.pad = 0;
      e.id = ((long long)v.w << 32) | (unsigned)v.z;
      if (e.id < 0) {
        e.key = -INFINITY;
        e.id = LLONG_MAX - i;     // holes: distinct ids, after every real entry
      }
      m[i] = e;
    }
    __syncthreads();
    for (int i = tid; i < total; i += nthreads) {
      const KeyId64 e = m[i];
      int rank = 0;
      for (int j = 0; j < total; ++j) rank += better(m[j], e) ? 1 : 0;
      if (rank < k) write_result<METRIC>(p, qi, rank, e.key, e.id, e.key == -INFINITY && e.id > LLONG_MAX - 2048);
    }
    return;
  }
  int n_sort = 32;
  while (n_sort < total) n_sort <<= 1;
  for (int i = tid; i < n_sort; i += nthreads) {
    KeyId64 e;
    e.key = -INFINITY;
    e.pad = 0;
    e.id = LLONG_MAX;
    if (i < total) {
      const int r = i / k, j = i - r * k;
      const int4 v = __ldcg(reinterpret_cast<const int4*>(mine + (size_t)r * CSS_MAX_K + j));
      const long long gid = ((long long)v.w << 32) | (unsigned)v.z;
      if (gid >= 0) {
        e.key = __int_as_float(v.x);
        e.id = gid;
      }
    }
    m[i] = e;
  }
  bitonic_sort_desc(m, n_sort, tid, nthreads);
  for (int i = tid; i < k; i += nthreads) {
    const KeyId64 e = m[i];
    write_result<METRIC>(p, qi, i, e.key, e.id, e.id == LLONG_MAX);
  }
}

// Timeline of one query for css_debug_scan_trace: row b = block b {0 start, 1 sweep done, 2 block list merged,
// 3 list re-scored, 4 ticket drawn}; row gridDim.x = the last block {0 lists requested, 4 result emitted}.
// Thread 0 only; ns of %globaltimer.
__device__ __forceinline__ void scan_stamp(const ScanParams& p, const int row, const int slot) {
  if (p.trace != nullptr && threadIdx.x == 0) p.trace[row * 8 + slot] = (long long)global_timer_ns();
}

// Single-query host calls: tell the polling host thread that the result is in place (all threads of the CTA that
// wrote it call this; the result and the flag may both live in mapped host memory).
__device__ __forceinline__ void signal_done(const ScanParams& p, const int tid, const unsigned unproven) {
  if (p.done_flag == nullptr) return;
  __threadfence_system();
  __syncthreads();
  if (tid == 0) st_release_sys(p.done_flag, (p.done_seq << 1) | unproven);
}

template <int METRIC>
__device__ __forceinline__ void emit_topk(const ScanParams& p, const int qi, KeyId* s, const int tid, const int nthreads) {
  const int k = p.k_out;
  if (p.ex.n_ranks <= 1) {
    for (int i = tid; i < k; i += nthreads) {
      const KeyId e = s[i];
      write_result<METRIC>(p, qi, i, e.key, e.id == kEmptyId ? 0ll : map_id(p.idmap, e.id), e.id == kEmptyId);
    }
    return;
  }
  const int R = p.ex.n_ranks;
  const size_t cell = ((size_t)(p.ex.epoch & 1u) * p.ex.max_nq + qi) * R;
  // publish: my list into slot (parity, qi, my rank) of every rank (mine included), then the flags
  for (int idx = tid; idx < R * k; idx += nthreads) {
    const int r = idx / k, i = idx - r * k;
    const KeyId e = s[i];
    const long long gid = (e.id == kEmptyId) ? -1ll : map_id(p.idmap, e.id);
    int4 v;
    v.x = __float_as_int(e.key);
    v.y = 0;
    v.z = (int)(gid & 0xffffffffll);
    v.w = (int)(gid >> 32);
    *reinterpret_cast<int4*>(p.ex.slots[r] + (cell + p.ex.rank) * CSS_MAX_K + i) = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < R) st_release_sys(p.ex.flags[tid] + cell + p.ex.rank, p.ex.epoch);
  if (p.ex.deferred) return;   // awaited + merged by the launch that follows (see ExchangeDev)
  __syncthreads();
  await_and_merge<METRIC>(p, qi, reinterpret_cast<KeyId64*>(s), tid, nthreads);
}

// ------------------------------------------------------------------------
// Two-phase scan, phase 2, run by the last CTA of a query's phase-1 grid (no extra launch).  Phase 1 (the
// shadow sweep: bf16 or int8 rows) left, per scan block, the kp best rows of that block by shadow score (sorted;
// `part`) and -- computed by the block itself right after its sweep -- their exact fp32 scores (`part_exact`,
// the arithmetic of the fp32 scan: bit-identical scores).  `eps` bounds |shadow score - exact score| of any
// stored row (scan_one_query; bf16: ||q|| (1.001 max_err + 4e-6 max_norm) from |x.q - bf16(x).q| <=
// ||x - bf16(x)|| ||q|| (Cauchy-Schwarz; max_err is the largest rounding-error norm of any stored row, tracked
// exactly at add time: about 0.4 * 2^-8 ||x|| for dense rows, against the worst case 2^-8 ||x|| of
// round-to-nearest with an 8-bit significand) plus the fp32 summations; int8: see quantize_query_i8).
// With T = the k-th best EXACT score among k or more distinct list entries (a lower bound of the corpus' k-th
// best exact score), every row of the true top-k has a shadow score >= T - eps =: thr.  A block list whose last
// entry is still >= thr may have dropped such a row: the query is queued for the fp32 scan (ovf_list), as it
// is when more than kRescoreCap rows pass -- never answered approximately.  Otherwise every row with shadow
// score >= thr is in the lists and the best k of them by exact score are the exact result.
// Selection is by rank counting (no sorting network: one barrier per step); s: kMergeCap entries.
// Returns true when the result was emitted.
// ------------------------------------------------------------------------
// Exact fp32 scores of two rows with the arithmetic of the fp32 scan (score_rows: lane l holds float4 32 j + l,
// FMAs in element order, butterfly sum) -- bit-identical scores.  Whole warp; q: the query in global memory.
__device__ __forceinline__ void exact_score_pair(const float* __restrict__ x, const float* __restrict__ q, const int id0,
                                                 const int id1, const int lane, float& a0, float& a1) {
  const float4* r0 = reinterpret_cast<const float4*>(x + (size_t)id0 * 768);
  const float4* r1 = reinterpret_cast<const float4*>(x + (size_t)id1 * 768);
  float4 v0[6], v1[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    v0[j] = ld_stream_f4(r0 + j * 32 + lane);
    v1[j] = ld_stream_f4(r1 + j * 32 + lane);
  }
  a0 = 0.f;
  a1 = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 qv = __ldg(reinterpret_cast<const float4*>(q) + j * 32 + lane);
    a0 = fmaf(v0[j].x, qv.x, a0); a0 = fmaf(v0[j].y, qv.y, a0); a0 = fmaf(v0[j].z, qv.z, a0); a0 = fmaf(v0[j].w, qv.w, a0);
    a1 = fmaf(v1[j].x, qv.x, a1); a1 = fmaf(v1[j].y, qv.y, a1); a1 = fmaf(v1[j].z, qv.z, a1); a1 = fmaf(v1[j].w, qv.w, a1);
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
}

constexpr int kRescoreCap = 2048;    // candidates kept per query at most (second half of s: rank-sort output)
constexpr int kTwoPhaseMaxK = 32;    // largest k the two-phase scan serves
constexpr int kRankSortMax = 1024;   // above this many candidates: bitonic sort instead of rank counting
constexpr int kTwoPhaseMaxBlocks = 160;  // scan blocks (= SMs) the register-held list entries of two_phase_finish cover

__device__ __forceinline__ KeyId ldcg_keyid(const KeyId* p) {
  const unsigned long long raw = __ldcg(reinterpret_cast<const unsigned long long*>(p));
  KeyId e;
  e.key = __uint_as_float((unsigned)(raw & 0xffffffffull));
  e.id = (int)(raw >> 32);
  return e;
}

// Number of entries of s[0..n) that come before e in the result order.  Eight loads in flight: the plain loop pays
// one shared-memory latency per entry (3.5 us for the 220 candidates of an int8-tier query).
__device__ __forceinline__ int rank_among(const KeyId* s, const int n, const KeyId e) {
  int rank = 0;
  int j = 0;
  for (; j + 8 <= n; j += 8) {
    KeyId t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = s[j + u];
#pragma unroll
    for (int u = 0; u < 8; ++u) rank += better(t[u], e) ? 1 : 0;
  }
  for (; j < n; ++j) rank += better(s[j], e) ? 1 : 0;
  return rank;
}

template <int KPL>
__device__ __forceinline__ bool two_phase_finish(const ScanParams& p, const int qi, KeyId* s, const int tid, const float eps) {
  __shared__ float s_t;
  __shared__ int s_cnt, s_unproven, s_anyfull, s_maxlast;   // s_maxlast: float bits, largest last entry of a full list
  constexpr int kp = 32 * KPL;   // == p.k: the list length is the warp list's capacity
  const int k = p.k_out, blocks = gridDim.x;
  const KeyId* lists = p.part + (size_t)qi * blocks * kp;
  // every list entry is read ONCE, all loads in flight together (one L2 round trip), and kept in registers for the
  // three selection steps below: thread t holds entries t, t + 512, ...  (kTwoPhaseMaxBlocks bounds the grid)
  constexpr int kPer = (kTwoPhaseMaxBlocks * 32 * KPL + kScanThreads - 1) / kScanThreads;
  const int total = blocks * kp;
  KeyId ent[kPer];
  float exact[kPer];   // fp32 score of the entry, computed by the block that owns the list (scan_one_query)
  const float* lists_exact = p.part_exact + (size_t)qi * blocks * kp;
#pragma unroll
  for (int c = 0; c < kPer; ++c) {
    const int i = tid + c * kScanThreads;
    ent[c].key = -INFINITY;
    ent[c].id = kEmptyId;
    exact[c] = -INFINITY;
    if (i < total) {
      ent[c] = ldcg_keyid(lists + i);
      exact[c] = __ldcg(lists_exact + i);
      if (ent[c].id == kEmptyId) exact[c] = -INFINITY;   // the owning block writes the scores of real entries only
    }
  }
  if (tid == 0) {
    s_t = -INFINITY;
    s_cnt = 0;
    s_unproven = 0;
    s_anyfull = 0;
    s_maxlast = __float_as_int(-INFINITY);
  }
  scan_stamp(p, gridDim.x, 0);
  // (A) T = the k-th best EXACT score among the list heads (k distinct rows score at least that; with fewer than k
  //     lists: among their first k entries), a lower bound of the k-th best exact score of the corpus.  Every row of
  //     the true top-k therefore has an exact score >= T and a shadow score >= T - eps =: thr -- ONE eps: the exact
  //     scores of the list entries are at hand (the owning blocks computed them), so the threshold need not be
  //     derived from shadow scores (k-th best shadow score - 2 eps: twice the slack, three times the candidates, and a
  //     second selection pass to tighten it).
  const int per = blocks >= k ? 1 : k;
  const int nsel = blocks * per;
#pragma unroll
  for (int c = 0; c < kPer; ++c) {
    const int i = tid + c * kScanThreads;
    const int within = i % kp;
    if (i < total && within < per) {
      KeyId e;
      e.key = exact[c];
      e.id = ent[c].id;
      s[(i / kp) * per + within] = e;
    }
  }
  __syncthreads();
  for (int i = tid; i < nsel; i += kScanThreads) {
    const KeyId e = s[i];
    if (e.id == kEmptyId) continue;
    if (rank_among(s, nsel, e) == k - 1) s_t = e.key;
  }
  __syncthreads();
  // eps: the tier's bound on |shadow score - exact score| (see scan_one_query); a bound that is not finite
  // (non-finite query or stored row) proves nothing
  const float thr = (s_t > -INFINITY) ? s_t - eps : -INFINITY;
  if (tid == 0 && !(eps < INFINITY)) s_unproven = 1;
  // (C) candidates over the full lists (shadow score >= thr), and the largest last entry of a FULL list: every row
  //     that is in no list has a shadow score at or below that
#pragma unroll
  for (int c = 0; c < kPer; ++c) {
    if (ent[c].id == kEmptyId) continue;
    const int i = tid + c * kScanThreads;
    if (i % kp == kp - 1) {
      s_anyfull = 1;
      const float f = ent[c].key;   // float maximum on the bit pattern (never NaN: NaN scores enter no list)
      if (f >= 0.f) atomicMax(&s_maxlast, __float_as_int(f));
      else atomicMin(reinterpret_cast<unsigned*>(&s_maxlast), __float_as_uint(f));
    }
    if (ent[c].key >= thr) {
      const int pos = atomicAdd(&s_cnt, 1);
      if (pos < kRescoreCap) {
        s[pos].key = exact[c];   // from here on the exact fp32 score
        s[pos].id = ent[c].id;
      }
    }
  }
  __syncthreads();
  const int keep = s_cnt;
  // queue the query for the fp32 scan (never answered approximately)
  auto give_up = [&]() {
    if (tid == 0) {
      p.ovf_list[atomicAdd(p.ovf_count, 1)] = qi;
      const unsigned nq_seen = atomicAdd(p.stats_dev, 1u) + 1u;
      const unsigned nu = atomicAdd(p.stats_dev + 1, 1u) + 1u;
      if (p.stats_host) {
        p.stats_host[0] = nq_seen;
        p.stats_host[1] = nu;
      }
    }
  };
  if (s_unproven || keep > kRescoreCap) {
    give_up();
    return false;
  }
  // (D) the candidates' exact fp32 scores were computed by the blocks that own them, in parallel over the grid
  //     (scan_one_query), so the tail of the query holds no DRAM round trip.
  // (E) order by (score desc, id asc)
  KeyId* out = s;
  if (keep <= kRankSortMax) {
    out = s + kRescoreCap;
    if (keep <= kScanThreads) {
      // Every warp sorts 32 candidates across its lanes (bitonic network of shuffles) and hands its best k to a
      // final rank count over at most 16 k entries: ~1 us.  Rank counting over all ~250 candidates of an int8-tier
      // query is 60 k comparisons on one SM: 4.5 us.
      const int lane = tid & 31, warp = tid >> 5;
      KeyId e;
      e.key = -INFINITY;
      e.id = kEmptyId;
      if (tid < keep) e = s[tid];
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int o = size >> 1; o > 0; o >>= 1) {
          const bool desc = (lane & size) == 0 || size == 32;
          WarpBufTop32::exchange(e.key, e.id, lane, o, ((lane & o) == 0) == desc);
        }
      }
      KeyId* best = s + kRescoreCap + 1024;   // [warps with candidates][k]
      const int nw = (keep + 31) >> 5;
      if (warp < nw && lane < k) best[warp * k + lane] = e;
      __syncthreads();
      const int nbest = nw * k;
      for (int i = tid; i < nbest; i += kScanThreads) {
        const KeyId c = best[i];
        if (c.id == kEmptyId) continue;
        const int rank = rank_among(best, nbest, c);
        if (rank < k) out[rank] = c;
      }
    } else {
      for (int i = tid; i < keep; i += kScanThreads) {
        const KeyId e = s[i];
        const int rank = rank_among(s, keep, e);
        if (rank < k) out[rank] = e;
      }
    }
    for (int i = keep + tid; i < k; i += kScanThreads) {
      out[i].key = -INFINITY;
      out[i].id = kEmptyId;
    }
    __syncthreads();
  } else {
    int n2 = 32;
    while (n2 < keep) n2 <<= 1;
    for (int i = keep + tid; i < n2; i += kScanThreads) {
      s[i].key = -INFINITY;
      s[i].id = kEmptyId;
    }
    bitonic_sort_desc(s, n2, tid, kScanThreads);
  }
  if (out != s) {   // emit_topk reuses s as its merge buffer: hand it the list at the front
    KeyId e;
    if (tid < k) e = out[tid];
    __syncthreads();
    if (tid < k) s[tid] = e;
    __syncthreads();
  }
  // (F) the proof, with the best threshold there is: T1 = the k-th best exact score just found (k distinct rows
  //     score at least that, T1 >= T), so every row of the true top-k has a shadow score >= T1 - eps -- at or above
  //     thr, i.e. among the candidates if it is in a list at all; and no row outside the lists reaches it when the
  //     largest last entry of a full list is below it.  (Fewer than k candidates: T1 = -inf, provable only when no
  //     list is full, i.e. nothing was dropped anywhere.)
  {
    const float t1 = s[k - 1].id != kEmptyId ? s[k - 1].key : -INFINITY;
    const bool proven = !s_anyfull || __int_as_float(s_maxlast) < t1 - eps;
    if (!proven) {
      __syncthreads();
      give_up();
      return false;
    }
  }
  emit_topk<CSS_METRIC_INNER_PRODUCT>(p, qi, s, tid, kScanThreads);
  scan_stamp(p, gridDim.x, 4);
  return true;
}

// Units of the dense int8 sweep for one warp.  Most of the warp's share is dealt statically, block-cyclically
// (round r: unit r * nwarps + gw) -- neighbouring rows land in different blocks, and the addresses are known two
// units ahead without any traffic.  The last eighth of the corpus (+ two rounds) is handed out from a per-query
// cursor: blocks do not stream at the same rate (sweep ends spread over 88 .. 110 us of a 1 M-row query when every
// warp gets the same share; SMs further from the memory partitions they read see more latency per bulk copy), so
// the fast warps take what the slow ones have not reached.  A ticket is worth four consecutive units, the tickets
// of the last two rounds one unit each (coarse tickets keep the single-address atomic rate at a quarter of the unit
// rate -- pairs of units on every ticket made a 4 M-row sweep 15 % slower --, fine ones let the warps finish
// within a unit of each other).  The cursor value is fetched one ticket ahead (lane 0's `raw`), its latency never
// waited for.
struct UnitFeed {
  int units, nwarps, gw, stat_rounds, stat_units, a_tickets, a_units;
  int r;           // next static round
  int cur;         // next unit of the ticket in hand
  int left;            // units left of the ticket in hand
  unsigned raw;        // lane 0: the cursor value fetched ahead
  unsigned* cursor;
  bool done;

  __device__ __forceinline__ void init(const ScanParams& p, const int qi, const int units_, const int nwarps_,
                                       const int gw_, const int lane) {
    units = units_;
    nwarps = nwarps_;
    gw = gw_;
    const int per_warp = units / nwarps;
    const int dyn_rounds = per_warp / 8 + 2;
    stat_rounds = per_warp > dyn_rounds ? per_warp - dyn_rounds : 0;
    stat_units = stat_rounds * nwarps;
    const int dyn_units = units - stat_units;
    const int b_units = dyn_units < 2 * nwarps ? dyn_units : 2 * nwarps;
    a_tickets = (dyn_units - b_units) / 4;
    a_units = a_tickets * 4;
    r = 0;
    cur = 0;
    left = 0;
    cursor = p.cursor + qi;
    done = false;
    raw = 0;
    if (lane == 0) raw = atomicAdd(cursor, 1u);
  }
  // Next unit of this warp, -1 when the corpus is exhausted.  Warp-uniform.
  __device__ __forceinline__ int next(const int lane) {
    if (r < stat_rounds) return (r++) * nwarps + gw;
    if (left > 0) {
      --left;
      return cur++;
    }
    if (done) return -1;
    const int t = (int)__shfl_sync(0xffffffffu, raw, 0);
    const int base = t < a_tickets ? stat_units + 4 * t : stat_units + a_units + (t - a_tickets);
    if (base >= units) {
      done = true;
      return -1;
    }
    if (lane == 0) raw = atomicAdd(cursor, 1u);
    const int cnt = t < a_tickets ? 4 : 1;
    left = (cnt < units - base ? cnt : units - base) - 1;
    cur = base + 1;
    return base;
  }
};

// grid = (blocks, nq).  Each warp owns a contiguous range of 8-row units, keeps
// a register top-k, the block merges its 16 warps in shared memory, and the last
// block to finish (atomic ticket) merges the per-block lists into D/I.
// SH = 1 / 2 (inner product, d = 768): phase 1 of the two-phase scan -- scores from the bf16 / int8 shadow rows.
template <int KPL, int METRIC, bool D768, int SH = 0>
__device__ __forceinline__ void scan_one_query(const ScanParams& p, const int qi) {
  static_assert(SH == 0 || (D768 && METRIC == CSS_METRIC_INNER_PRODUCT), "shadow phase: inner product, d = 768");
  static_assert(SH >= 0 && SH <= 3, "tier");
  // SH = 2: int8 tier, dense sweep (no mask, or the alive bits alone); SH = 3: int8 tier, filtered sweep (gather).
  // Two instantiations: the gather code in the dense kernel cost the dense sweep 2 % (register allocation).
  constexpr bool BF16 = SH == 1;
  constexpr bool I8 = SH == 2 || SH == 3;
  constexpr bool I8_GATHER = SH == 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyId* s_list = reinterpret_cast<KeyId*>(smem_raw);  // kMergeCap entries
  float* q_s = reinterpret_cast<float*>(smem_raw + sizeof(KeyId) * kMergeCap);  // d floats (generic path)
  __shared__ int s_is_last;
  // filtered scan: selected rows of a window (int8 tier: of a round of windows, see the masked sweep below)
  __shared__ unsigned short s_rows[kScanWarps][(SH == 3 ? 64 : 32) * kRowsPerUnit];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const float* q = p.q + (int64_t)qi * p.d;

  float4 qreg[6];
  float qn = 0.f;   // ||q|| (two-phase scan: the error bound of two_phase_finish)
  float eps = 0.f;  // two-phase scan: bound on |shadow score - exact score| of any stored row
  int q1c[12], q2c[12];   // int8 tier: the query's two code vectors
  float a1 = 0.f;
  UnitFeed feed;          // int8 tier, dense sweep: the warp's units
  int64_t u_cur = -1, u_nxt = -1;
  if constexpr (I8) {
    if constexpr (!I8_GATHER) {
      // dense sweep (below): the warp's first two units are requested before anything else, the query is
      // quantised while they are on their way
      // (a shard holds fewer than 2^31 rows: unit indices fit 32 bits)
      const int units0 = (int)((p.n + kRowsPerUnit - 1) / kRowsPerUnit);
      feed.init(p, qi, units0, (int)gridDim.x * kScanWarps, (int)blockIdx.x + (int)gridDim.x * warp, lane);
      u_cur = feed.next(lane);
      u_nxt = feed.next(lane);
      if (lane == 0) {
        unsigned char* ring = smem_raw + warp * (kI8Stages * kI8UnitBytes);
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kI8RingBytes) + warp * kI8Stages;
#pragma unroll
        for (int st = 0; st < kI8Stages; ++st) sb_mbar_init(bars + st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int st = 0; st < kI8Stages; ++st) {
          const int64_t us = st ? u_nxt : u_cur;
          if (us >= 0) {
            sb_mbar_expect_tx(bars + st, kI8UnitBytes);
            sb_bulk_load(ring + st * kI8UnitBytes, p.xq + us * kI8UnitBytes, kI8UnitBytes, bars + st);
          }
        }
      }
    }
    float dq;
    quantize_query_i8(q, lane, q1c, q2c, a1, qn, dq);
    const float e8 = *p.max_err8, mn = *p.max_norm;
    eps = 1.001f * (qn * e8 + dq * (mn + e8)) + 4e-6f * qn * mn;
  } else if constexpr (BF16) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      qreg[2 * j] = __ldg(reinterpret_cast<const float4*>(q + j * 256 + lane * 8));
      qreg[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(q + j * 256 + lane * 8) + 1);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      qn = fmaf(qreg[j].x, qreg[j].x, qn);
      qn = fmaf(qreg[j].y, qreg[j].y, qn);
      qn = fmaf(qreg[j].z, qreg[j].z, qn);
      qn = fmaf(qreg[j].w, qreg[j].w, qn);
    }
    qn = sqrtf(warp_sum(qn)) * 1.00001f;
    eps = qn * (1.001f * (*p.max_err) + 4e-6f * (*p.max_norm));
  } else if constexpr (D768) {
#pragma unroll
    for (int j = 0; j < 6; ++j) qreg[j] = __ldg(reinterpret_cast<const float4*>(q) + j * 32 + lane);
  } else {
    for (int j = tid; j < p.d; j += kScanThreads) q_s[j] = q[j];
    __syncthreads();
  }

  scan_stamp(p, blockIdx.x, 0);
  // shadow sweeps with 32-entry lists: buffered selection (see WarpBufTop32); p.k == 32 there
  constexpr bool kBuffered = SH != 0 && KPL == 1;
  typename std::conditional<kBuffered, WarpBufTop32, WarpTopK<KPL>>::type top;
  top.init(lane, p.k);

  const int64_t units = (p.n + kRowsPerUnit - 1) / kRowsPerUnit;
  const int64_t nwarps = (int64_t)gridDim.x * kScanWarps;
  const int64_t gw = (int64_t)blockIdx.x * kScanWarps + warp;
  const int64_t u_begin = (units * gw) / nwarps;
  const int64_t u_end = (units * (gw + 1)) / nwarps;
  const unsigned char* mask8 = reinterpret_cast<const unsigned char*>(p.mask);

  bool swept = false;
  if constexpr (I8 && !I8_GATHER) {
    {
      // Dense sweep, 8-row units dealt block-cyclically (see the bf16 sweep below).  (With the alive bits as the
      // only mask -- an index rows were deleted from -- every row is still streamed and the dead ones are dropped
      // after scoring: gathering the 99.99 % live rows instead costs 15 % more.)  A unit is 6 KB of contiguous
      // int8 rows: one lane per warp fetches it with a TMA bulk copy into the warp's two-stage ring in shared
      // memory (192 KB per CTA in flight, no registers held by data in flight); the warp copies an arrived unit
      // into registers, re-arms the stage with the unit after next at once and only then does the arithmetic, so
      // both stages stay in flight during the dot products.  A unit is four row pairs, one row per half-warp.
      const int hl = lane & 15, half = lane >> 4;
      const int64_t last = p.n - 1;
      unsigned char* ring = smem_raw + warp * (kI8Stages * kI8UnitBytes);
      uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kI8RingBytes) + warp * kI8Stages;
      // (u_cur and u_nxt were requested at the top; see UnitFeed for the order of the units)
      // scales of the rows in flight: lane i < 8 holds row i's of the current (sc_a) and the next unit (sc_b)
      float sc_a = 0.f, sc_b = 0.f;
      if (lane < kRowsPerUnit) {
        if (u_cur >= 0) sc_a = __ldg(p.xs + min(u_cur * kRowsPerUnit + lane, last));
        if (u_nxt >= 0) sc_b = __ldg(p.xs + min(u_nxt * kRowsPerUnit + lane, last));
      }
      // mask byte of the units in flight (all lanes the same)
      unsigned mk_a = 0xFFu, mk_b = 0xFFu;
      if (mask8 != nullptr) {
        if (u_cur >= 0) mk_a = __ldg(mask8 + u_cur);
        if (u_nxt >= 0) mk_b = __ldg(mask8 + u_nxt);
      }
      __syncwarp();
      for (uint32_t it = 0; u_cur >= 0; ++it) {
        const int64_t u = u_cur;
        const uint32_t st = it & 1u;
        sb_mbar_wait(bars + st, (it >> 1) & 1u);
        const unsigned char* src = ring + st * kI8UnitBytes + half * 768 + hl * 16;
        uint4 v[4][3];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
          for (int j = 0; j < 3; ++j) v[s][j] = *reinterpret_cast<const uint4*>(src + s * 1536 + j * 256);
        const float sc_cur = sc_a * a1;
        sc_a = sc_b;
        const unsigned mk_cur = mk_a;
        mk_a = mk_b;
        __syncwarp();   // every lane has read the stage: it may be overwritten
        const int64_t u2 = feed.next(lane);
        u_cur = u_nxt;
        u_nxt = u2;
        if (u2 >= 0) {
          if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            sb_mbar_expect_tx(bars + st, kI8UnitBytes);
            sb_bulk_load(ring + st * kI8UnitBytes, p.xq + u2 * kI8UnitBytes, kI8UnitBytes, bars + st);
          }
          if (lane < kRowsPerUnit) sc_b = __ldg(p.xs + min(u2 * kRowsPerUnit + lane, last));
          if (mask8 != nullptr) mk_b = __ldg(mask8 + u2);
        }
        const int64_t row0 = u * kRowsPerUnit;
        const bool tail = row0 + kRowsPerUnit > p.n || mk_cur != 0xFFu;   // the corpus' last unit, or a unit with dead rows
        float f[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) f[s] = dot_i8(v[s], q1c, q2c, __shfl_sync(0xffffffffu, sc_cur, 2 * s + half));
        if constexpr (kBuffered) {
          // the four butterflies interleave; afterwards the lanes of a half all hold their row's score and ONE vote
          // decides for the unit's eight rows
#pragma unroll
          for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int s = 0; s < 4; ++s) f[s] += __shfl_xor_sync(0xffffffffu, f[s], o);
          if (tail) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
              if (row0 + 2 * s + half > last || !((mk_cur >> (2 * s + half)) & 1u)) f[s] = -INFINITY;
          }
          const float fm = fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3]));
          if (__any_sync(0xffffffffu, fm > top.thr)) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const float a_lo = __shfl_sync(0xffffffffu, f[s], 0), a_hi = __shfl_sync(0xffffffffu, f[s], 16);
              top.consider(a_lo, (int)(row0 + 2 * s), lane);
              top.consider(a_hi, (int)(row0 + 2 * s + 1), lane);
            }
          }
        } else {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            float a_lo, a_hi;
            half_sums(f[s], a_lo, a_hi);
            if (row0 + 2 * s <= last && ((mk_cur >> (2 * s)) & 1u)) top.consider(a_lo, (int)(row0 + 2 * s), lane);
            if (row0 + 2 * s + 1 <= last && ((mk_cur >> (2 * s + 1)) & 1u)) top.consider(a_hi, (int)(row0 + 2 * s + 1), lane);
          }
        }
        top.maybe_compact(lane);
      }
      swept = true;
    }
  }
  if constexpr (I8_GATHER) {
    {
      // Filtered sweep.  The warp walks its contiguous range of units in rounds: it compacts the selected rows of
      // as many 256-row windows as it takes to collect ~256 rows (their offsets from the round's first row, 16 bits
      // each), then scores them eight at a time through the warp's two-stage ring in shared memory: a stage is
      // eight rows, every lane copying (cp.async, 16 B) exactly the twelve pieces it will read itself, two groups
      // ahead of the arithmetic -- no registers held by data in flight, no cross-lane hand-over.  (Eight 768-byte
      // TMA bulk copies per group ran at the TMA unit's request rate, ~30 ns each: no faster than plain register
      // loads, one group at a time.)
      const int hl = lane & 15, half = lane >> 4;
      unsigned char* ring = smem_raw + warp * (kI8Stages * kI8UnitBytes);
      unsigned short* pend = s_rows[warp];
      uint32_t it = 0;   // stages consumed so far: ring position
      int64_t ub = u_begin;
      while (ub < u_end) {
        const int64_t round_base = ub * kRowsPerUnit;
        int n_pend = 0;
        for (int windows = 0; ub < u_end && n_pend <= 256 && windows < 255; ++windows, ub += 32) {
          unsigned mb = 0;
          const int64_t u = ub + lane;
          if (u < u_end) {
            mb = (unsigned)mask8[u];
            const int64_t rows_left = p.n - u * kRowsPerUnit;
            if (rows_left < kRowsPerUnit) mb &= (1u << rows_left) - 1u;
          }
          const int cnt = __popc(mb);
          int pos = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pos, o);
            if (lane >= o) pos += t;
          }
          const int total = __shfl_sync(0xffffffffu, pos, 31);
          pos += n_pend - cnt;
          while (mb) {
            const int b = __ffs(mb) - 1;
            mb &= mb - 1;
            pend[pos++] = (unsigned short)(windows * 256 + lane * kRowsPerUnit + b);
          }
          n_pend += total;
        }
        __syncwarp();
        if (n_pend == 0) continue;
        const int groups = (n_pend + 7) >> 3;
        // group g into stage st_: this lane's twelve 16-byte pieces (rows 2 s + half, s = 0..3); lanes 0..7 keep
        // their row's scale
        auto issue = [&](const int g, const uint32_t st_, float& sc_out) {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const int64_t row = round_base + pend[min(8 * g + 2 * s + half, n_pend - 1)];
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.xq) + row * 768 + hl * 16;
            const uint32_t dst = sb_smem_u32(ring + st_ * kI8UnitBytes + (2 * s + half) * 768 + hl * 16);
#pragma unroll
            for (int j = 0; j < 3; ++j)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 256), "l"(src + j * 256) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (lane < kRowsPerUnit) sc_out = __ldg(p.xs + round_base + pend[min(8 * g + lane, n_pend - 1)]);
        };
        float sc_a = 0.f, sc_b = 0.f;
        issue(0, it & 1u, sc_a);
        if (groups > 1) issue(1, (it + 1u) & 1u, sc_b);
        for (int g = 0; g < groups; ++g, ++it) {
          const uint32_t st = it & 1u;
          if (g + 1 < groups) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
          const unsigned char* src = ring + st * kI8UnitBytes + half * 768 + hl * 16;
          uint4 v[4][3];
#pragma unroll
          for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[s][j] = *reinterpret_cast<const uint4*>(src + s * 1536 + j * 256);
          const float sc_cur = sc_a * a1;
          sc_a = sc_b;
          if (g + 2 < groups) issue(g + 2, st, sc_b);   // a lane overwrites only the pieces it has just read
          const int valid = min(kRowsPerUnit, n_pend - 8 * g);
          float f[4];
#pragma unroll
          for (int s = 0; s < 4; ++s) f[s] = dot_i8(v[s], q1c, q2c, __shfl_sync(0xffffffffu, sc_cur, 2 * s + half));
#pragma unroll
          for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int s = 0; s < 4; ++s) f[s] += __shfl_xor_sync(0xffffffffu, f[s], o);
#pragma unroll
          for (int s = 0; s < 4; ++s)
            if (2 * s + half >= valid) f[s] = -INFINITY;   // padding of the round's last group
          bool any = true;
          if constexpr (kBuffered) {
            const float fm = fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3]));
            any = __any_sync(0xffffffffu, fm > top.thr);
          }
          if (any) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const float a_lo = __shfl_sync(0xffffffffu, f[s], 0), a_hi = __shfl_sync(0xffffffffu, f[s], 16);
              if (2 * s < valid) top.consider(a_lo, (int)(round_base + pend[8 * g + 2 * s]), lane);
              if (2 * s + 1 < valid) top.consider(a_hi, (int)(round_base + pend[8 * g + 2 * s + 1]), lane);
            }
          }
          top.maybe_compact(lane);
        }
        __syncwarp();   // the round's offsets are dead: the next round may overwrite them
      }
      swept = true;
    }
  }
  if constexpr (BF16) {
    if (mask8 == nullptr && p.interleave) {
      // Dense sweep, 8-row units dealt block-cyclically: unit u belongs to block u % gridDim.x, warp
      // (u / gridDim.x) % 16.  Neighbouring rows -- the chunks of one session, which are each other's
      // nearest neighbours -- land in different blocks, so no single block list fills up with the rows
      // around the k-th score (the proof of two_phase_finish fails when one does).
      const int64_t stride = (int64_t)gridDim.x * kScanWarps;
      for (int64_t u = blockIdx.x + (int64_t)gridDim.x * warp; u < units; u += stride) {
        const int64_t row0 = u * kRowsPerUnit;
        const int valid = (int)min((int64_t)kRowsPerUnit, p.n - row0);
        int64_t r[kRowsPerUnit];
#pragma unroll
        for (int i = 0; i < kRowsPerUnit; ++i) r[i] = row0 + (i < valid ? i : 0);
        float acc[kRowsPerUnit];
        score_rows_bf16(p, qreg, r, lane, acc);
#pragma unroll
        for (int i = 0; i < kRowsPerUnit; ++i)
          if (i < valid) top.consider(acc[i], (int)(row0 + i), lane);
        top.maybe_compact(lane);
      }
      swept = true;
    }
  }
  for (int64_t ub = swept ? u_end : u_begin; ub < u_end; ub += 32) {
    // one coalesced mask fetch for the next 32 units (256 rows)
    unsigned mb = 0;
    {
      int64_t u = ub + lane;
      if (u < u_end) {
        mb = mask8 ? (unsigned)mask8[u] : 0xFFu;
        int64_t rows_left = p.n - u * kRowsPerUnit;
        if (rows_left < kRowsPerUnit) mb &= (1u << rows_left) - 1u;
      }
    }
    const int nu = (int)min((int64_t)32, u_end - ub);
    if constexpr (BF16) {
      if (mask8 == nullptr) {   // dense: one 8-row unit per step (only the corpus tail has cleared bits)
        for (int j = 0; j < nu; ++j) {
          const unsigned m = __shfl_sync(0xffffffffu, mb, j);
          const int64_t row0 = (ub + j) * kRowsPerUnit;
          int64_t r[kRowsPerUnit];
#pragma unroll
          for (int i = 0; i < kRowsPerUnit; ++i) r[i] = row0 + (((m >> i) & 1u) ? i : 0);   // rows past the end re-read the first
          float acc[kRowsPerUnit];
          score_rows_bf16(p, qreg, r, lane, acc);
#pragma unroll
          for (int i = 0; i < kRowsPerUnit; ++i)
            if ((m >> i) & 1u) top.consider(acc[i], (int)(row0 + i), lane);
          top.maybe_compact(lane);
        }
        continue;
      }
    }
    if (mask8 != nullptr) {
      // Filtered scan: compact the selected rows of the 256-row window into a per-warp list first,
      // so that every score_rows call carries kRowsPerGroup distinct rows whatever the selectivity
      // (unit by unit, a 5 % filter left one useful row -- and three duplicate loads -- per call).
      const int cnt = __popc(mb);
      int pos = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pos, o);
        if (lane >= o) pos += t;
      }
      const int total = __shfl_sync(0xffffffffu, pos, 31);
      pos -= cnt;
      unsigned mm = mb;
      while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        s_rows[warp][pos++] = (unsigned short)(lane * kRowsPerUnit + b);
      }
      __syncwarp();
      const int64_t base = ub * kRowsPerUnit;
      if constexpr (BF16) {
        for (int g0 = 0; g0 < total; g0 += kRowsPerUnit) {
          int64_t r[kRowsPerUnit];
#pragma unroll
          for (int i = 0; i < kRowsPerUnit; ++i) r[i] = base + s_rows[warp][min(g0 + i, total - 1)];
          float acc[kRowsPerUnit];
          score_rows_bf16(p, qreg, r, lane, acc);
#pragma unroll
          for (int i = 0; i < kRowsPerUnit; ++i)
            if (g0 + i < total) top.consider(acc[i], (int)r[i], lane);
          top.maybe_compact(lane);
        }
        __syncwarp();
        continue;
      }
      for (int g0 = 0; g0 < total; g0 += kRowsPerGroup) {
        int64_t r[kRowsPerGroup];
#pragma unroll
        for (int i = 0; i < kRowsPerGroup; ++i) r[i] = base + s_rows[warp][min(g0 + i, total - 1)];
        float acc[kRowsPerGroup];
        score_rows<METRIC, D768>(p, qreg, q_s, r, lane, acc);
#pragma unroll
        for (int i = 0; i < kRowsPerGroup; ++i)
          if (g0 + i < total) top.consider(acc[i], (int)r[i], lane);
      }
      __syncwarp();
      continue;
    }
    for (int j = 0; j < nu; ++j) {
      unsigned m = __shfl_sync(0xffffffffu, mb, j);
      const int64_t row0 = (ub + j) * kRowsPerUnit;
      while (m) {
        int64_t r[kRowsPerGroup];
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < kRowsPerGroup; ++i) {
          if (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            r[i] = row0 + b;
            cnt = i + 1;
          } else {
            r[i] = r[0];
          }
        }
        float acc[kRowsPerGroup];
        score_rows<METRIC, D768>(p, qreg, q_s, r, lane, acc);
#pragma unroll
        for (int i = 0; i < kRowsPerGroup; ++i)
          if (i < cnt) top.consider(acc[i], (int)r[i], lane);
      }
    }
  }

  scan_stamp(p, blockIdx.x, 1);
  // ---- block merge: 16 warps x 32 lanes x KPL slots -> sorted smem list ----
  constexpr int kBlockEntries = kScanThreads * KPL;  // power of two
  if constexpr (kBuffered) {
    // the warps' lists are sorted across the lanes: a tournament of pairwise merges (better(A[i], B[31 - i]) is a
    // bitonic sequence holding the 32 best of both; five exchange steps sort it) -- four barriers instead of the
    // 45 of a 512-entry sorting network
    top.flush(lane);
    __syncthreads();   // every warp has left the sweep: the lists may overwrite the ring of bulk copies
    float mk = top.key[0];
    int mi = top.id[0];
#pragma unroll
    for (int stride = 1; stride < kScanWarps; stride <<= 1) {
      const int role = warp & (2 * stride - 1);
      if (role == stride) {
        KeyId e;
        e.key = mk;
        e.id = mi;
        s_list[warp * 32 + lane] = e;
      }
      __syncthreads();
      if (role == 0) {
        const KeyId o = s_list[(warp + stride) * 32 + (31 - lane)];
        if (better(o.key, o.id, mk, mi)) {
          mk = o.key;
          mi = o.id;
        }
#pragma unroll
        for (int x = 16; x > 0; x >>= 1) WarpBufTop32::exchange(mk, mi, lane, x, (lane & x) == 0);
      }
    }
    if (warp == 0) {
      KeyId e;
      e.key = mk;
      e.id = mi;
      s_list[lane] = e;   // slot 0 was never a sender's
    }
    __syncthreads();
  } else {
    __syncthreads();
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      KeyId e;
      bool live = (lane * KPL + s) < p.k;
      e.key = live ? top.key[s] : -INFINITY;
      e.id = live ? top.id[s] : kEmptyId;
      s_list[tid * KPL + s] = e;
    }
    bitonic_sort_desc(s_list, kBlockEntries, tid, kScanThreads);
  }
  KeyId* my_part = p.part + ((int64_t)qi * gridDim.x + blockIdx.x) * p.k;
  for (int i = tid; i < p.k; i += kScanThreads) my_part[i] = s_list[i];
  scan_stamp(p, blockIdx.x, 2);
  if constexpr (SH != 0) if (!p.no_merge) {
    // Two-phase scan: every block re-scores ITS list in fp32 right away -- 148 blocks x one DRAM round trip in
    // parallel (and overlapping the blocks still sweeping) instead of the last block walking all candidates alone
    // (int8 tier: ~220 rows within 2 eps of the k-th score, seven dependent round trips = 10 us of a 135 us query).
    // 32 rows x 3 KB per block: 2 % of the sweep's bytes.
    float* my_exact = p.part_exact + ((int64_t)qi * gridDim.x + blockIdx.x) * p.k;
    for (int i0 = warp * 2; i0 < p.k; i0 += kScanWarps * 2) {
      const int id0 = s_list[i0].id, id1 = s_list[i0 + 1].id;
      if (id0 == kEmptyId) break;   // sorted: nothing but empty slots from here on
      float a0, a1;
      exact_score_pair(p.x, q, id0, id1 == kEmptyId ? id0 : id1, lane, a0, a1);
      if (lane == 0) {
        my_exact[i0] = a0;
        my_exact[i0 + 1] = id1 == kEmptyId ? -INFINITY : a1;
      }
    }
  }

  // ---- grid merge by the last block -----------------------------------------
  scan_stamp(p, blockIdx.x, 3);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned t = atomicAdd(p.ticket + qi, 1u);
    s_is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  scan_stamp(p, blockIdx.x, 4);
  if (!s_is_last) return;
  __threadfence();
  if (tid == 0) {   // ready for the next launch
    p.ticket[qi] = 0;
    if (p.cursor != nullptr) p.cursor[qi] = 0;
  }
  if (p.no_merge) return;   // phase 1 timed alone: the lists are all that is wanted

  if constexpr (SH != 0) {
    // two-phase scan: prove + re-score in fp32 (or queue the query for the fp32 scan)
    const bool proven = two_phase_finish<KPL>(p, qi, s_list, tid, eps);
    signal_done(p, tid, proven ? 0u : 1u);
    if (proven && tid == 0) {   // counters last: their round trip to L2 is not part of the query
      const unsigned nq_seen = atomicAdd(p.stats_dev, 1u) + 1u;
      if (p.stats_host) p.stats_host[0] = nq_seen;
    }
    return;
  }

  const KeyId* all = p.part + (int64_t)qi * gridDim.x * p.k;
  const int total = gridDim.x * p.k;
  // s_list[0..k) always holds the best-so-far; each round appends up to
  // kMergeCap - k new entries, pads with sentinels and sorts.
  int consumed = 0;
  bool first = true;
  while (consumed < total) {
    int keep = first ? 0 : p.k;
    int take = min(kMergeCap - keep, total - consumed);
    __syncthreads();
    for (int i = tid; i < kMergeCap - keep; i += kScanThreads) {
      KeyId e;
      if (i < take) {
        e = ldcg_keyid(all + consumed + i);
      } else {
        e.key = -INFINITY;
        e.id = kEmptyId;
      }
      s_list[keep + i] = e;
    }
    // sort only as many entries as needed (next power of two >= keep + take)
    int n_sort = 32;
    while (n_sort < keep + take) n_sort <<= 1;
    bitonic_sort_desc(s_list, n_sort, tid, kScanThreads);
    consumed += take;
    first = false;
  }
  emit_topk<METRIC>(p, qi, s_list, tid, kScanThreads);
  signal_done(p, tid, 0u);
}

// grid = (blocks, nq) scans query blockIdx.y; with p.qlist set, grid = (blocks, F) and
// slice y walks the listed queries y, y+F, ... (device-side fallback of the batched path).
template <int KPL, int METRIC, bool D768, int SH = 0>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(ScanParams p) {
  // programmatic dependent launch: the fp32-fallback launch behind a two-phase sweep is set up while the sweep
  // runs and only waits here (an empty overflow list -- the normal case -- then costs ~1 us instead of a launch gap)
  if (p.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if constexpr (SH != 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.zero_on_entry != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.zero_on_entry = 0;
  if (p.qlist == nullptr) {
    scan_one_query<KPL, METRIC, D768, SH>(p, blockIdx.y);
    return;
  }
  const int cnt = *p.qcount;
  for (int slot = blockIdx.y; slot < cnt; slot += gridDim.y) {
    scan_one_query<KPL, METRIC, D768, SH>(p, p.qlist[slot]);
    __syncthreads();
  }
  if (p.ex.n_ranks > 1 && p.ex.deferred && blockIdx.x == 0) {
    // deferred exchange: every query of the search has been published by now-or-soon (by phase 1, by the scans
    // above, or by a peer's); no publisher ever waits, so these waits cannot deadlock across ranks
    extern __shared__ __align__(16) unsigned char smem_raw[];
    for (int qi = blockIdx.y; qi < p.ex.nq; qi += gridDim.y) {
      await_and_merge<METRIC>(p, qi, reinterpret_cast<KeyId64*>(smem_raw), threadIdx.x, kScanThreads);
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------
// S1: append rows.  One warp per row: optional L2 normalisation with the
// reference's epsilon (x / (||x|| + 1e-8)), fp32 store + bf16 shadow store + (dst_q8 != nullptr) int8 shadow
// store: codes rint(v / scale) with scale = max|v| / 127 per row, the scale, and the exact quantisation-error
// norm ||v - scale * code|| folded into max_err8.
// ------------------------------------------------------------------------
static __global__ void append_rows_kernel(const float* __restrict__ src, int64_t n, int d, int normalize,
                                   float* __restrict__ dst, __nv_bfloat16* __restrict__ dst_bf16,
                                   float* __restrict__ max_norm, float* __restrict__ max_err,
                                   int8_t* __restrict__ dst_q8, float* __restrict__ dst_scale,
                                   float* __restrict__ max_err8) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* s = src + row * d;
  float denom = 1.f;
  float ss = 0.f, amax = 0.f;
  for (int j = lane; j < d; j += 32) {
    float v = s[j];
    ss = fmaf(v, v, ss);
    amax = fmaxf(amax, fabsf(v));
  }
  ss = warp_sum(ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  float nrm = sqrtf(ss);
  if (normalize) {
    denom = nrm + 1e-8f;
    nrm = nrm / denom;
    amax = amax / denom;   // division is monotone: the largest |stored value|
  }
  // a row with a non-finite element gets zero codes and makes max_err8 infinite: the int8 tier then proves nothing
  const bool finite = ss < INFINITY;   // false for NaN too
  const float scale = finite ? amax / 127.f : 0.f;
  const float inv_scale = (finite && amax > 0.f) ? 127.f / amax : 0.f;
  float es = 0.f;   // ||v - bf16(v)||^2 of the stored row
  float e8 = 0.f;   // ||v - scale * code||^2
  for (int j = lane; j < d; j += 32) {
    // numpy computes x / (norm + 1e-8); a true division keeps the last bit identical
    float v = normalize ? s[j] / denom : s[j];
    if (dst != src || normalize) dst[row * d + j] = v;
    if (dst_bf16) {
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      dst_bf16[row * d + j] = b;
      const float e = v - __bfloat162float(b);   // exact in fp32
      es = fmaf(e, e, es);
    }
    if (dst_q8) {
      const float c = finite ? fminf(fmaxf(rintf(v * inv_scale), -127.f), 127.f) : 0.f;
      dst_q8[row * d + j] = (int8_t)__float2int_rn(c);
      const float e = fmaf(-scale, c, v);
      e8 = fmaf(e, e, e8);
    }
  }
  es = warp_sum(es);
  if (dst_q8) {
    e8 = warp_sum(e8);
    if (lane == 0) {
      dst_scale[row] = scale;
      const float bound = (finite && e8 < INFINITY) ? sqrtf(e8) * 1.001f : INFINITY;
      atomicMax(reinterpret_cast<int*>(max_err8), __float_as_int(bound));
    }
  }
  // largest stored row norm and largest bf16 rounding-error norm: the error bounds of the two-phase scan
  // and of the batched search derive from them (non-negative floats order like their bit patterns;
  // the 1.0001 covers the rounding of the fp32 sums, non-finite rows poison neither)
  if (lane == 0 && max_norm && nrm == nrm) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(nrm * 1.0001f));
  if (lane == 0 && max_err && es == es && es < INFINITY)
    atomicMax(reinterpret_cast<int*>(max_err), __float_as_int(sqrtf(es) * 1.0001f));
}

// ------------------------------------------------------------------------
// Compaction (HybridStorage.optimize): row i of the window takes row ids[i] of the index.
// One warp per row, float4 copies of the fp32 row and its bf16 shadow into a staging
// buffer; the caller then copies the staging window to rows [i0, i0 + n) (ids ascending,
// so no source row of a later window has been overwritten).
// ------------------------------------------------------------------------
static __global__ void gather_rows_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ xb,
                                          const int64_t* __restrict__ ids, int64_t n, int d,
                                          float* __restrict__ out, __nv_bfloat16* __restrict__ outb,
                                          const int8_t* __restrict__ xq, const float* __restrict__ xs,
                                          int8_t* __restrict__ outq, float* __restrict__ outs) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int64_t src = ids[i];
  for (int j = lane; j < d; j += 32) {
    out[i * d + j] = x[src * d + j];
    outb[i * d + j] = xb[src * d + j];
    if (xq) outq[i * d + j] = xq[src * d + j];
  }
  if (xq && lane == 0) outs[i] = xs[src];
}
static __global__ void gather_i32_kernel(const int32_t* __restrict__ col, const int64_t* __restrict__ ids, int64_t n,
                                         int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = col[ids[i]];
}
// alive bits of the gathered rows, one output word per 32 rows (ballot)
static __global__ void gather_bits_kernel(const uint32_t* __restrict__ bits, const int64_t* __restrict__ ids, int64_t n,
                                          uint32_t* __restrict__ out_words) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool b = false;
  if (i < n) {
    const int64_t s = ids[i];
    b = (bits[s >> 5] >> (s & 31)) & 1u;
  }
  const unsigned w = __ballot_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0 && i < n) out_words[i >> 5] = w;
}

static __global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// Set bits [start, start+n) of a row bitmask to `value`.
static __global__ void set_bits_kernel(uint32_t* bits, int64_t start, int64_t n, int value) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // word index relative to start/32
  int64_t w0 = start >> 5;
  int64_t w1 = (start + n - 1) >> 5;
  int64_t word = w0 + w;
  if (word > w1) return;
  int64_t lo = max(start, word << 5) - (word << 5);
  int64_t hi = min(start + n, (word + 1) << 5) - (word << 5);  // exclusive
  uint32_t m = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((lo == 0) ? 0u : ((1u << lo) - 1u));
  if (value) atomicOr(bits + word, m);
  else atomicAnd(bits + word, ~m);
}

// Set / clear the bits of n listed rows (HybridStorage._kill_rows: one call per deletion batch).
static __global__ void set_bits_by_id_kernel(uint32_t* bits, const int64_t* __restrict__ ids, int64_t n, int64_t limit,
                                             int value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = ids[i];
  if (row < 0 || row >= limit) return;
  const uint32_t bit = 1u << (row & 31);
  if (value) atomicOr(bits + (row >> 5), bit);
  else atomicAnd(bits + (row >> 5), ~bit);
}

// alive bytes (0/1 per row) -> bits
static __global__ void alive_bytes_to_bits_kernel(const uint8_t* alive, int64_t start, int64_t n,
                                           uint32_t* bits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t row = start + i;
  uint32_t bit = 1u << (row & 31);
  if (alive[i]) atomicOr(bits + (row >> 5), bit);
  else atomicAnd(bits + (row >> 5), ~bit);
}

// ------------------------------------------------------------------------
// S4: filter predicate -> bitmask.  Four rows per thread, one mask word per 8 lanes.
// ------------------------------------------------------------------------
struct DevClause {
  const int32_t* col;     // nullptr: column never set -> all NULL -> nothing matches
  int32_t kind;
  int32_t lo, hi;
  const uint32_t* bits;   // device bitset
  int32_t nbits;
};
struct FilterParams {
  DevClause c[CSS_MAX_CLAUSES];
  int n_clauses;
  const uint32_t* alive;     // nullable
  const uint32_t* row_mask;  // nullable (device)
  int64_t n;
  uint32_t* out;             // ceil(n/32) words
  unsigned long long* n_pass;   // nullable: count of passing rows wanted
};

// One thread evaluates four consecutive rows: every clause column is read with one 128-bit load per thread whether
// or not an earlier clause already failed (independent loads, all in flight together), the four verdicts of a lane
// are a nibble of the mask word its group of 8 lanes assembles with three shuffles.  (One row per thread with
// short-circuit evaluation chained up to three dependent 4-byte loads per row: 80 us for the three-clause filter of
// config 5 over 10 M rows, 1.5 TB/s.)  Columns, alive bits and masks are allocated for a capacity that is a multiple
// of 32 rows: the loads of the last, partial group stay inside them.
static __global__ void __launch_bounds__(256, 6) filter_mask_kernel(FilterParams p) {
  const int64_t row0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int lane = threadIdx.x & 31;
  unsigned pass = 0;   // bit i: row0 + i passes
  if (row0 < p.n) {
    pass = 0xFu;
    const int sh = (int)(row0 & 31);
    if (p.alive) pass &= p.alive[row0 >> 5] >> sh;
    if (p.row_mask) pass &= p.row_mask[row0 >> 5] >> sh;
    pass &= 0xFu;
    const int64_t left = p.n - row0;
    if (left < 4) pass &= (1u << left) - 1u;
    // clauses four at a time: their columns' loads are issued together, then evaluated
    for (int i0 = 0; i0 < p.n_clauses; i0 += 4) {
      int4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = make_int4(0, 0, 0, 0);
        if (i0 + j < p.n_clauses && p.c[i0 + j].col != nullptr)
          v[j] = __ldg(reinterpret_cast<const int4*>(p.c[i0 + j].col + row0));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i0 + j < p.n_clauses) {
          const DevClause& c = p.c[i0 + j];
          if (c.col == nullptr) {
            pass = 0;   // column never set: all NULL, nothing matches
          } else {
            const int32_t x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            unsigned ok = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              bool m;
              if (x[r] == CSS_NULL_VALUE) m = false;
              else if (c.kind == CSS_CLAUSE_RANGE) m = (x[r] >= c.lo) && (x[r] <= c.hi);
              else m = (x[r] >= 0) && (x[r] < c.nbits) && ((c.bits[x[r] >> 5] >> (x[r] & 31)) & 1u);
              ok |= (m ? 1u : 0u) << r;
            }
            pass &= ok;
          }
        }
      }
    }
  }
  // lanes 8 g .. 8 g + 7 hold the eight nibbles of one mask word
  unsigned w = pass << (4 * (lane & 7));
  w |= __shfl_xor_sync(0xffffffffu, w, 1);
  w |= __shfl_xor_sync(0xffffffffu, w, 2);
  w |= __shfl_xor_sync(0xffffffffu, w, 4);
  if ((lane & 7) == 0 && row0 < p.n) p.out[row0 >> 5] = w;
  // the pass count is wanted by css_index_filter_mask only: one atomic per BLOCK there, none on the search path
  // (one per warp on a single address cost more than the rest of the kernel: 0.19 ms at 10 M rows)
  if (p.n_pass == nullptr) return;
  __shared__ unsigned s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if ((lane & 7) == 0 && w) atomicAdd(&s_cnt, (unsigned)__popc(w));
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(p.n_pass, (unsigned long long)s_cnt);
}

// ------------------------------------------------------------------------
// S5: merge n_lists top-k lists per query (one block per query).
// ------------------------------------------------------------------------
template <int METRIC>
__global__ void merge_lists_kernel(const float* __restrict__ D_in, const int64_t* __restrict__ I_in,
                                   int64_t d_stride, int64_t i_stride,   // elements between consecutive lists
                                   int n_lists, int nq, int k, float* __restrict__ D_out,
                                   int64_t* __restrict__ I_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyId64* s = reinterpret_cast<KeyId64*>(smem_raw);
  const int qi = blockIdx.x;
  const int tid = threadIdx.x;
  const int total = n_lists * k;
  int n_sort = 32;
  while (n_sort < total) n_sort <<= 1;
  for (int i = tid; i < n_sort; i += blockDim.x) {
    KeyId64 e;
    e.pad = 0;
    e.key = -INFINITY;
    e.id = LLONG_MAX;
    if (i < total) {
      int l = i / k, j = i % k;
      const int64_t within = (int64_t)qi * k + j;
      int64_t id = I_in[l * i_stride + within];
      if (id >= 0) {
        float dv = D_in[l * d_stride + within];
        e.key = (METRIC == CSS_METRIC_INNER_PRODUCT) ? dv : -dv;
        e.id = id;
      }
    }
    s[i] = e;
  }
  bitonic_sort_desc(s, n_sort, tid, blockDim.x);
  for (int i = tid; i < k; i += blockDim.x) {
    KeyId64 e = s[i];  // n_sort >= 32 >= ... k <= total <= n_sort
    bool empty = (e.id == LLONG_MAX);
    float dv;
    if constexpr (METRIC == CSS_METRIC_INNER_PRODUCT) dv = empty ? -FLT_MAX : e.key;
    else dv = empty ? FLT_MAX : -e.key;
    D_out[(int64_t)qi * k + i] = dv;
    I_out[(int64_t)qi * k + i] = empty ? (int64_t)-1 : (int64_t)e.id;
  }
}

}  // namespace css
