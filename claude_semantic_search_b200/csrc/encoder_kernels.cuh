// encoder_kernels.cuh -- device kernels of the MPNet chunk encoder other than the GEMM
// main loop (gemm_tc.cuh):
//   E1  token + position embedding gather, LayerNorm            modeling_mpnet.py:72-96
//   E3  multi-head attention with relative-position bias        modeling_mpnet.py:144-185,322-360
//   E5  LayerNorm over (GEMM + bias + residual): fused GEMM epilogue   modeling_mpnet.py:206,242
//   E7  masked mean pooling + L2 normalisation                  sentence-transformers Pooling / Normalize
//   GEMM epilogue functors: +bias, +bias+GELU(erf), +bias+residual
//
// Token layout: sequences are packed back to back (no padding rows); cu_seqlens[i] is the
// first token of sequence i.  Activations are [T, C] row-major.
#pragma once
#include "tc_common.cuh"

namespace css {
namespace enc {

constexpr int kHidden = 768;
constexpr int kHeads = 12;
constexpr int kHeadDim = 64;
constexpr int kFfn = 3072;
constexpr int kMaxSeq = 512;

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Streaming stores of the GEMM epilogues: the activations (hundreds of MB per launch) must not allocate in L1.  With
// 227 KB of shared memory the L1 is ~28 KB and has to keep what the epilogue re-reads for every tile -- the bias /
// gamma / beta vectors and the few spilled registers; with default stores their hit rate was 25 % / 16 % (ncu), i.e.
// every such load paid the L2 latency and the K = 768 projection ran at 42 % tensor-pipe activity.
__device__ __forceinline__ void st_global_stream(void* ptr, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_global_stream(void* ptr, const float2& v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(ptr), "f"(v.x), "f"(v.y) : "memory");
}

// Packed fp32x2 arithmetic of sm_100 (two independent fp32 lanes in one 64-bit register).
__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f32x2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// ------------------------------------------------------------------------
// fp32 -> bf16 (weights at load time)
// ------------------------------------------------------------------------
static __global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// ------------------------------------------------------------------------
// Row LayerNorm helper: one warp per row of 768, 24 values per lane held as 6 float4
// (lane owns columns j*128 + lane*4 .. +3).  Two-pass (mean, then centred variance).
// ------------------------------------------------------------------------
__device__ __forceinline__ void warp_layernorm_768(float4 (&v)[6], const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.f / kHidden);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.f / kHidden) + eps);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + j * 32 + lane);
    v[j].x = (v[j].x - mean) * rstd * g.x + b.x;
    v[j].y = (v[j].y - mean) * rstd * g.y + b.y;
    v[j].z = (v[j].z - mean) * rstd * g.z + b.z;
    v[j].w = (v[j].w - mean) * rstd * g.w + b.w;
  }
}

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* __restrict__ row, const float4 (&v)[6], int lane) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    uint2 p;
    p.x = pack_bf16(v[j].x, v[j].y);
    p.y = pack_bf16(v[j].z, v[j].w);
    *reinterpret_cast<uint2*>(row + j * 128 + lane * 4) = p;
  }
}

// E1: x[t] = LayerNorm(word_emb[ids[t]] + pos_emb[pad_id + 1 + index_in_sequence]).
static __global__ void embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ cu, int n_seq,
                                       int T, const float* __restrict__ word_emb, const float* __restrict__ pos_emb,
                                       int vocab, int max_pos, int pad_id, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  // sequence of token t: last i with cu[i] <= t
  int lo = 0, hi = n_seq;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(cu + mid) <= t) lo = mid; else hi = mid;
  }
  int id = __ldg(ids + t);
  id = min(max(id, 0), vocab - 1);
  // create_position_ids_from_input_ids: cumsum(ids != pad) * (ids != pad) + pad; packed
  // input holds no pad tokens, so the position is index + 1 + pad.
  int pos = (t - __ldg(cu + lo)) + 1 + pad_id;
  if (id == pad_id) pos = pad_id;
  pos = min(pos, max_pos - 1);
  const float4* w = reinterpret_cast<const float4*>(word_emb + (size_t)id * kHidden);
  const float4* p = reinterpret_cast<const float4*>(pos_emb + (size_t)pos * kHidden);
  float4 v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float4 a = __ldg(w + j * 32 + lane), b = __ldg(p + j * 32 + lane);
    v[j] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
  warp_layernorm_768(v, gamma, beta, eps, lane);
  store_row_bf16(x + (size_t)t * kHidden, v, lane);
}

// E7 for the LayerNorm-folded pipeline: x holds the pre-LayerNorm activation of the last projection;
// mean_t LN(x_t) = gamma * mean_t((x_t - mu_t) rstd_t) + beta.
static __global__ void pool_normalize_ln_kernel(const __nv_bfloat16* __restrict__ x, const float2* __restrict__ mr,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const int32_t* __restrict__ cu, int normalize, float* __restrict__ out) {
  const int s = blockIdx.x;
  const int t0 = cu[s], t1 = cu[s + 1];
  const int c = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int t = t0; t < t1; ++t) {
    const __nv_bfloat16* r = x + (size_t)t * kHidden;
    const float2 m = __ldg(mr + t);
    a0 += (__bfloat162float(r[c]) - m.x) * m.y;
    a1 += (__bfloat162float(r[c + 256]) - m.x) * m.y;
    a2 += (__bfloat162float(r[c + 512]) - m.x) * m.y;
  }
  const float denom = fmaxf((float)(t1 - t0), 1e-9f);  // Pooling: clamp(sum(mask), min=1e-9)
  a0 = a0 / denom * gamma[c] + beta[c];
  a1 = a1 / denom * gamma[c + 256] + beta[c + 256];
  a2 = a2 / denom * gamma[c + 512] + beta[c + 512];
  if (normalize) {
    __shared__ float red[8];
    float ss = a0 * a0 + a1 * a1 + a2 * a2;
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float nrm = fmaxf(sqrtf(tot), 1e-12f);  // F.normalize(p=2, eps=1e-12)
    a0 /= nrm; a1 /= nrm; a2 /= nrm;
  }
  float* o = out + (size_t)s * kHidden;
  o[c] = a0; o[c + 256] = a1; o[c + 512] = a2;
}

// E7: out[s] = normalise(mean over tokens of x).  One CTA (256 threads) per sequence,
// thread c owns columns c, c+256, c+512.
static __global__ void pool_normalize_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ cu,
                                             int normalize, float* __restrict__ out) {
  const int s = blockIdx.x;
  const int t0 = cu[s], t1 = cu[s + 1];
  const int c = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int t = t0; t < t1; ++t) {
    const __nv_bfloat16* r = x + (size_t)t * kHidden;
    a0 += __bfloat162float(r[c]);
    a1 += __bfloat162float(r[c + 256]);
    a2 += __bfloat162float(r[c + 512]);
  }
  const float denom = fmaxf((float)(t1 - t0), 1e-9f);  // Pooling: clamp(sum(mask), min=1e-9)
  a0 /= denom; a1 /= denom; a2 /= denom;
  if (normalize) {
    __shared__ float red[8];
    float ss = a0 * a0 + a1 * a1 + a2 * a2;
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float nrm = fmaxf(sqrtf(tot), 1e-12f);  // F.normalize(p=2, eps=1e-12)
    a0 /= nrm; a1 /= nrm; a2 /= nrm;
  }
  float* o = out + (size_t)s * kHidden;
  o[c] = a0; o[c + 256] = a1; o[c + 512] = a2;
}

// ------------------------------------------------------------------------
// E3: attention.  CTA = 4 warps = 64 query rows of one (sequence, head); the whole K and
// V of that (sequence, head) live in shared memory (L <= 512 keys x 64 dims x bf16), each
// warp runs an online-softmax sweep over 64-key blocks with mma.sync m16n8k16 (bf16 in,
// fp32 accumulate).  scores = q.k / 8 + rel_table[h][j - i]; keys >= L do not exist in
// the packed layout (the reference masks its padded keys with finfo.min -> weight 0).
// ------------------------------------------------------------------------
constexpr int kAttnThreads = 256;
constexpr int kAttnQRows = 128;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void cp_async_16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}

// smem: K [Lp][64] bf16 (128 B rows, 16-byte chunks XOR-swizzled by row & 7), V likewise,
// rel [2*Lp] fp32 (rel[d + Lp - 1] = bias of relative position d = j - i).
static __global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ cu,
                 const float* __restrict__ rel_table, int rel_half /* table holds d in [-rel_half, rel_half] */,
                 __nv_bfloat16* __restrict__ ctx) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int s = blockIdx.z, h = blockIdx.y;
  const int t0 = cu[s];
  const int L = cu[s + 1] - t0;
  const int q0 = blockIdx.x * kAttnQRows;
  if (q0 >= L) return;
  const int Lp = (L + 63) & ~63;
  uint8_t* sK = smem;
  uint8_t* sV = smem + (size_t)Lp * 128;
  float* sRel = reinterpret_cast<float*>(smem + (size_t)Lp * 256);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kLd = 3 * kHidden;

  // ---- stage K, V (cp.async) and the bias window ----
  {
    const int chunk = tid & 7;
    for (int r = tid >> 3; r < Lp; r += kAttnThreads / 8) {
      const uint32_t off = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
      if (r < L) {
        const __nv_bfloat16* g = qkv + (size_t)(t0 + r) * kLd + h * kHeadDim + chunk * 8;
        cp_async_16(tc::smem_u32(sK) + off, g + kHidden);
        cp_async_16(tc::smem_u32(sV) + off, g + 2 * kHidden);
      } else {
        *reinterpret_cast<uint4*>(sK + off) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sV + off) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* rt = rel_table + (size_t)h * (2 * rel_half + 1);
    for (int i = tid; i < 2 * Lp - 1; i += kAttnThreads) {
      int d = i - (Lp - 1);
      d = max(-rel_half, min(rel_half, d));
      sRel[i] = rt[d + rel_half];
    }
  }

  // ---- Q fragments straight from global (rows of this warp) ----
  const int qrow = q0 + warp * 16 + (lane >> 2);  // and qrow + 8
  uint32_t qa[4][4];
  {
    const __nv_bfloat16* g0 = qkv + (size_t)(t0 + min(qrow, L - 1)) * kLd + h * kHeadDim + (lane & 3) * 2;
    const __nv_bfloat16* g1 = qkv + (size_t)(t0 + min(qrow + 8, L - 1)) * kLd + h * kHeadDim + (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(g0 + ks * 16));
      qa[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(g1 + ks * 16));
      qa[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(g0 + ks * 16 + 8));
      qa[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(g1 + ks * 16 + 8));
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (q0 + warp * 16 >= L) return;  // this warp's 16 query rows are all past the sequence end

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;
  const uint32_t sK_u = tc::smem_u32(sK), sV_u = tc::smem_u32(sV);

  for (int kb = 0; kb < Lp; kb += 64) {
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      const int key = kb + nt * 8 + (lane & 7);
#pragma unroll
      for (int ks2 = 0; ks2 < 2; ++ks2) {
        uint32_t b[4];
        const int c = ks2 * 4 + (lane >> 3);
        ldmatrix_x4(b, sK_u + (uint32_t)key * 128u + (uint32_t)((c ^ (key & 7)) << 4));
        mma_bf16_16816(sc[nt], qa[ks2 * 2], b[0], b[1]);
        mma_bf16_16816(sc[nt], qa[ks2 * 2 + 1], b[2], b[3]);
      }
    }
    // scale, bias, mask; running max
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = kb + nt * 8 + (lane & 3) * 2 + (e & 1);
        const int i = min(qrow + ((e >> 1) << 3), L - 1);  // rows >= L are never stored
        float v = sc[nt][e] * 0.125f + sRel[j - i + Lp - 1];
        v = (j < L) ? v : -INFINITY;
        sc[nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: key kb is always < L
    const float corr0 = exp2f((m0 - mn0) * kLog2e), corr1 = exp2f((m1 - mn1) * kLog2e);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f((sc[nt][0] - mn0) * kLog2e), p1 = exp2f((sc[nt][1] - mn0) * kLog2e);
      const float p2 = exp2f((sc[nt][2] - mn1) * kLog2e), p3 = exp2f((sc[nt][3] - mn1) * kLog2e);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l0 = l0 * corr0 + rs0;
    l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      o[nd][0] *= corr0; o[nd][1] *= corr0; o[nd][2] *= corr1; o[nd][3] *= corr1;
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int key = kb + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
      for (int nd2 = 0; nd2 < 4; ++nd2) {
        uint32_t b[4];
        const int c = nd2 * 2 + (lane >> 4);
        ldmatrix_x4_trans(b, sV_u + (uint32_t)key * 128u + (uint32_t)((c ^ (key & 7)) << 4));
        mma_bf16_16816(o[nd2 * 2], pa[kk], b[0], b[1]);
        mma_bf16_16816(o[nd2 * 2 + 1], pa[kk], b[2], b[3]);
      }
    }
  }
  // row sums across the quad
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  __nv_bfloat16* c0 = ctx + (size_t)(t0 + qrow) * kHidden + h * kHeadDim + (lane & 3) * 2;
  __nv_bfloat16* c1 = c0 + (size_t)8 * kHidden;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    if (qrow < L) *reinterpret_cast<uint32_t*>(c0 + nd * 8) = pack_bf16(o[nd][0] * inv0, o[nd][1] * inv0);
    if (qrow + 8 < L) *reinterpret_cast<uint32_t*>(c1 + nd * 8) = pack_bf16(o[nd][2] * inv1, o[nd][3] * inv1);
  }
}

// ------------------------------------------------------------------------
// GEMM epilogue functors (see gemm_tc.cuh for the concept).
// ------------------------------------------------------------------------
// GELU with the exact-erf definition (transformers ACT2FN["gelu"]).  erf through Abramowitz &
// Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of the result): two MUFU
// (rcp, ex2) + ~10 FMA instead of the ~27-instruction erff() -- the FFN up-projection
// epilogue was ALU-bound on erff in the round-1 profile.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float erf_abs = fmaf(-p, e, 1.f);          // erf(|x| / sqrt 2)
  return 0.5f * x + 0.5f * fabsf(x) * erf_abs;     // 0.5 x (1 + sign(x) erf_abs)
}

// The same GELU for two values at once with the packed fp32x2 instructions of sm_100 (FFMA2 / FMUL2 /
// FADD2): ~10 issue slots per value instead of ~19 -- the FFN up-projection epilogue is bound by
// instruction issue, not by the tensor core.  gelu(x) = max(x, 0) - 0.5 |x| p(t) exp(-x^2 / 2).
__device__ __forceinline__ uint64_t gelu_erf_x2(uint64_t x2) {
  const uint64_t ax2 = x2 & 0x7fffffff7fffffffull;
  const uint64_t d2 = f32x2_fma(ax2, f32x2_pack(0.23164188f, 0.23164188f), f32x2_pack(1.f, 1.f));   // 1 + 0.3275911 |x| / sqrt 2
  float d0, d1, t0, t1;
  f32x2_unpack(d2, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t2 = f32x2_pack(t0, t1);
  // -0.5 * (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5)
  uint64_t q2 = f32x2_fma(t2, f32x2_pack(-0.5307027145f, -0.5307027145f), f32x2_pack(0.7265760135f, 0.7265760135f));
  q2 = f32x2_fma(q2, t2, f32x2_pack(-0.7107068705f, -0.7107068705f));
  q2 = f32x2_fma(q2, t2, f32x2_pack(0.142248368f, 0.142248368f));
  q2 = f32x2_fma(q2, t2, f32x2_pack(-0.127414796f, -0.127414796f));
  q2 = f32x2_mul(q2, t2);
  const uint64_t w2 = f32x2_mul(ax2, f32x2_pack(0.8493218f, 0.8493218f));   // |x| sqrt(log2(e) / 2)
  const uint64_t m2 = f32x2_mul(w2, w2);
  float m0, m1, e0, e1;
  f32x2_unpack(m2, m0, m1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-m0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-m1));
  float x0, x1;
  f32x2_unpack(x2, x0, x1);
  return f32x2_fma(f32x2_mul(ax2, q2), f32x2_pack(e0, e1), f32x2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

// Both functors stage the warp's 32 x 32 block through shared memory so that global
// stores are row-contiguous full sectors (round-1 profile: per-thread row stores cost 32
// half-written sectors per request and doubled the L2 write traffic).
//
// out_bf16[m, n] = acc + bias[n]            (QKV projection)
// out_bf16[m, n] = gelu(acc + bias[n])      (FFN up-projection)
//
// kFold: the A operand is the PRE-LayerNorm activation v (bf16) of the producing projection and the
// LayerNorm is folded into this GEMM:  LN(v) W^T + b = rstd_m (v (W gamma)^T - mu_m c_n) + d_n  with
// c_n = sum_k (W gamma)[n, k] and d_n = b_n + sum_k beta_k W[n, k] prepared at load time (the weight
// operand is W gamma, `bias` carries d, `colsum` carries c) and (mu_m, rstd_m) from
// ln_stats_finalize_kernel.  The normalised activation is never written to or read from memory.
template <bool kGelu, bool kFold = false>
struct EpiBiasBf16 {
  static constexpr bool kMasksColumns = false;
  static constexpr bool kPanel = false;
  static constexpr bool kResidPrefetch = false;
  static constexpr int kRowBytes = 80;                 // 64 B of payload + 16 B pad: conflict-free
  static constexpr int kStageBytes = 32 * kRowBytes;   // per epilogue warp
  struct Params {
    __nv_bfloat16* out;
    const float* bias;     // b, or d when kFold
    int ldo;
    const float* colsum;   // kFold: c [N]
    const float2* mr;      // kFold: (mean, rstd) per row [M]
  };
  const Params& p;
  uint8_t* stage;
  float2 mr_cur = make_float2(0.f, 1.f);
  __device__ EpiBiasBf16(const Params& p_, int, uint8_t* stage_) : p(p_), stage(stage_) {}
  // Called before the wait for the tile's accumulator: the row statistics arrive while the tensor core still works
  // on the tile.  (Requested a tile ahead they were spilled by ptxas -- a store that waits for the load on the spot.)
  __device__ __forceinline__ void tile_begin(int m_warp, int lane, int M) {
    if constexpr (kFold) mr_cur = __ldg(p.mr + min(m_warp + lane, M - 1));
  }
  __device__ __forceinline__ void chunk(int slot, int m_warp, int lane, int M, int n0, const uint32_t (&v)[32]) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
    uint4* srow = reinterpret_cast<uint4*>(stage + lane * kRowBytes);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 ba = __ldg(b4 + g * 2), bb = __ldg(b4 + g * 2 + 1);
      float f[8];
      if constexpr (kFold) {
        const float4* c4 = reinterpret_cast<const float4*>(p.colsum + n0);
        const float4 ca = __ldg(c4 + g * 2), cb = __ldg(c4 + g * 2 + 1);
        // y = rstd acc + (d - rstd mu c), two columns per FFMA2
        const uint64_t rs2 = f32x2_pack(mr_cur.y, mr_cur.y);
        const uint64_t nb2 = f32x2_pack(-mr_cur.x * mr_cur.y, -mr_cur.x * mr_cur.y);
        const uint64_t c2[4] = {f32x2_pack(ca.x, ca.y), f32x2_pack(ca.z, ca.w), f32x2_pack(cb.x, cb.y), f32x2_pack(cb.z, cb.w)};
        const uint64_t d2[4] = {f32x2_pack(ba.x, ba.y), f32x2_pack(ba.z, ba.w), f32x2_pack(bb.x, bb.y), f32x2_pack(bb.z, bb.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t acc2 = f32x2_pack(__uint_as_float(v[g * 8 + 2 * k]), __uint_as_float(v[g * 8 + 2 * k + 1]));
          f32x2_unpack(f32x2_fma(rs2, acc2, f32x2_fma(nb2, c2[k], d2[k])), f[2 * k], f[2 * k + 1]);
        }
      } else {
      f[0] = __uint_as_float(v[g * 8 + 0]) + ba.x; f[1] = __uint_as_float(v[g * 8 + 1]) + ba.y;
      f[2] = __uint_as_float(v[g * 8 + 2]) + ba.z; f[3] = __uint_as_float(v[g * 8 + 3]) + ba.w;
      f[4] = __uint_as_float(v[g * 8 + 4]) + bb.x; f[5] = __uint_as_float(v[g * 8 + 5]) + bb.y;
      f[6] = __uint_as_float(v[g * 8 + 6]) + bb.z; f[7] = __uint_as_float(v[g * 8 + 7]) + bb.w;
      }
      if constexpr (kGelu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) f32x2_unpack(gelu_erf_x2(f32x2_pack(f[2 * i], f[2 * i + 1])), f[2 * i], f[2 * i + 1]);
      }
      uint4 o;
      o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
      o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
      srow[g] = o;
    }
    __syncwarp();
    // 4 lanes per row (4 x 16 B = the row's 64 B), 8 rows per instruction
    const int piece = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2);
      const int m = m_warp + r;
      if (m < M) {
        const uint4 o = *reinterpret_cast<const uint4*>(stage + r * kRowBytes + piece * 16);
        st_global_stream(p.out + (size_t)m * p.ldo + n0 + piece * 8, o);
      }
    }
    __syncwarp();
  }
  __device__ __forceinline__ void prefetch(int, int, int, int, int) {}
  __device__ __forceinline__ void prefetch_none() {}
  __device__ __forceinline__ void tile_end(int, int, int, int) {}
  __device__ __forceinline__ void finish() {}
};

// (mean, rstd) of every row from the partial statistics the EpiResidLN epilogue leaves behind
// (fixed summation order: results are run-to-run identical).
static __global__ void ln_stats_finalize_kernel(const float2* __restrict__ parts, int T, float eps,
                                                float2* __restrict__ mr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  constexpr int kPart4 = 3 * kGemmEpiColSplit / 2;
  const float4* s4 = reinterpret_cast<const float4*>(parts + (size_t)t * (2 * kPart4));
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < kPart4; ++i) {
    const float4 a = __ldg(s4 + i);
    s += a.x + a.z;
    q += a.y + a.w;
  }
  const float mean = s * (1.f / kHidden);
  mr[t] = make_float2(mean, rsqrtf(fmaxf(q * (1.f / kHidden) - mean * mean, 0.f) + eps));
}

// y[m, :] = LayerNorm(acc[m, :] + bias + resid[m, :]) * gamma + beta  (attention output / FFN down
// projections, N = 768).  Per chunk, v = acc + bias + resid in fp32; the thread (= one row) adds v
// and v^2 to its row statistics; v goes out as bf16, staged through shared memory for full-sector
// stores.  The fp32 pre-LayerNorm activation never exists in memory.
//
// kFused = true (attention output projection, K = 768): the GEMM runs in panel order (a CTA
//   visits the three 256-column blocks of its rows back to back); after the third block the two
//   column halves of a row exchange their statistics through shared memory, then each warp
//   re-reads 16 of the CTA's 128 rows (L2 hits: written microseconds ago), normalises them and
//   rewrites them in place.  No LayerNorm kernel, no device-scope fence.
// kFused = false (FFN down projection, K = 3072): panel order would make every CTA pair stream
//   its own 1.5 MB A panel three times with 116 MB of panels live -- more than the L2 holds
//   (measured: 478 us against 381 us) -- so the tiles keep the L2-friendly order, the partial
//   statistics go to stats[row][tile][column half] and ln_apply_kernel finishes the rows
//   (1.5 KB + 48 B read per token instead of the 3 KB fp32 row of a plain LayerNorm kernel).
template <bool kFused>
struct EpiResidLN {
  static constexpr bool kMasksColumns = false;
  static constexpr bool kPanel = kFused;
  static constexpr bool kResidPrefetch = true;   // the TMA producer pulls the residual tile into L2 ahead of the epilogue
  // one 32 x 32 bf16 block (residual in, result out): rows of 64 B, no padding; the 16-byte piece g of
  // row r sits at position g ^ ((r >> 1) & 3), which makes both access patterns conflict-free (lane = row
  // for the arithmetic, 4 lanes per row for the copies) and keeps the epilogue at 4 KB per warp, i.e.
  // five pipeline stages for the main loop instead of four
  static constexpr int kRowBytes = 64;
  static constexpr int kSlotBytes = 32 * kRowBytes;
  static constexpr int kWarpStage = 2 * kSlotBytes;              // per epilogue warp: one slot per chunk of its column slice
  static constexpr int kSplit = kGemmEpiColSplit;              // column slices of a tile, one epilogue warp each
  static constexpr int kWarps = kGemmEpiWarps;
  static constexpr int kStatBytes = kFused ? 2 * kSplit * 128 * 2 * 4 : 0;  // [panel parity][column slice][row][sum, sumsq]
  static constexpr int kStageBytes = kWarpStage + kStatBytes / kWarps;  // per epilogue warp (the warps share the statistics)
  struct Params {
    __nv_bfloat16* out;            // [M, 768]
    const float* bias;             // [768]
    const __nv_bfloat16* resid;    // [M, 768]
    const float* gamma;
    const float* beta;
    float eps;
    float2* stats;                 // !kFused: [M][3 tiles][4 column slices] (sum, sum of squares)
    // LayerNorm-folded pipeline: `resid` holds the PRE-LayerNorm activation of the previous
    // projection and the residual is LN(resid) = (resid - mean) * rstd * rgamma + rbeta, rebuilt here
    // in fp32 (rmr == nullptr: `resid` is the residual itself)
    const float2* rmr;             // (mean, rstd) per row of `resid`
    const float* rgamma;
    const float* rbeta;
  };
  static const void* resid_ptr(const Params& q) { return q.resid; }
  const Params& p;
  uint8_t* stage;
  float* sstats;
  int ew, lane_;
  int my_row = 0;                  // absolute row of this thread's accumulator lane in the current tile
  uint32_t parity = 0;
  uint64_t sum2 = 0ull, sq2 = 0ull;   // row statistics, two interleaved partial sums each (fp32x2)
  float2 mr_cur = make_float2(0.f, 1.f), mr_next = make_float2(0.f, 1.f);
  static __device__ __forceinline__ int slot_off(int r, int g) { return r * kRowBytes + ((g ^ ((r >> 1) & 3)) << 4); }
  __device__ EpiResidLN(const Params& p_, int epi_thread, uint8_t* stage_) : p(p_) {
    ew = epi_thread >> 5;
    lane_ = epi_thread & 31;
    uint8_t* base = stage_ - ew * kStageBytes;
    stage = base + ew * kWarpStage;
    sstats = reinterpret_cast<float*>(base + kWarps * kWarpStage);
  }
  // The residual block of chunk `slot` of this CTA's NEXT tile is copied asynchronously
  // (cp.async, 16 B per lane, 4 lanes x 16 B = the 64 B a row contributes, 8 rows per instruction)
  // straight into the slot that the current tile's chunk has just finished with: a tile ahead, with
  // no registers held.  Holding the prefetch in registers made ptxas spill in-flight loads, i.e.
  // wait for them on the spot (ncu: the STL stalls on the long scoreboard) -- the K = 768 projection
  // then ran at 180 us against 98 us without the residual.
  __device__ __forceinline__ void prefetch(int slot, int m_warp, int lane, int M, int n0) {
    const int piece = lane & 3;
    const uint32_t dst0 = tc::smem_u32(stage + slot * kSlotBytes);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2);
      const int m = m_warp + r;
      const __nv_bfloat16* src = p.resid + (size_t)min(m, M - 1) * kHidden + n0 + piece * 8;   // rows >= M are never stored
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + slot_off(r, piece)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // (mean, rstd) of the next tile's residual rows, a tile ahead (requested at the start of the tile instead, before
    // the wait for the accumulator, the K = 768 projection was 7 % slower: 230k against 214k cycles)
    if (p.rmr != nullptr && slot == 0) mr_next = __ldg(p.rmr + min(m_warp + lane, M - 1));
  }
  __device__ __forceinline__ void prefetch_none() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void tile_begin(int, int, int) {}
  __device__ __forceinline__ void chunk(int slot, int m_warp, int lane, int M, int n0, const uint32_t (&v)[32]) {
    uint8_t* blk = stage + slot * kSlotBytes;
    // the copy into this slot was committed one tile ago; the only younger group is the other chunk's
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    uint4 rrow[4];   // this thread's row: 32 residual values
#pragma unroll
    for (int g = 0; g < 4; ++g) rrow[g] = *reinterpret_cast<const uint4*>(blk + slot_off(lane, g));
    __syncwarp();
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
    my_row = m_warp + lane;
    const bool rebuild = p.rmr != nullptr;   // kernel-uniform
    if (rebuild && slot == 0) mr_cur = mr_next;
    const uint64_t rs2 = f32x2_pack(mr_cur.y, mr_cur.y), nb2 = f32x2_pack(-mr_cur.x * mr_cur.y, -mr_cur.x * mr_cur.y);
    // packed fp32x2 arithmetic: two columns per FADD2 / FFMA2
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 ba = __ldg(b4 + g * 2), bb = __ldg(b4 + g * 2 + 1);
      const uint4 r4 = rrow[g];
      const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
      const uint64_t bias2[4] = {f32x2_pack(ba.x, ba.y), f32x2_pack(ba.z, ba.w), f32x2_pack(bb.x, bb.y), f32x2_pack(bb.z, bb.w)};
      uint64_t g2[4], be2[4];
      if (rebuild) {
        const float4* g4 = reinterpret_cast<const float4*>(p.rgamma + n0);
        const float4* e4 = reinterpret_cast<const float4*>(p.rbeta + n0);
        const float4 ga = __ldg(g4 + g * 2), gb = __ldg(g4 + g * 2 + 1), ea = __ldg(e4 + g * 2), eb = __ldg(e4 + g * 2 + 1);
        g2[0] = f32x2_pack(ga.x, ga.y); g2[1] = f32x2_pack(ga.z, ga.w); g2[2] = f32x2_pack(gb.x, gb.y); g2[3] = f32x2_pack(gb.z, gb.w);
        be2[0] = f32x2_pack(ea.x, ea.y); be2[1] = f32x2_pack(ea.z, ea.w); be2[2] = f32x2_pack(eb.x, eb.y); be2[3] = f32x2_pack(eb.z, eb.w);
      }
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t acc2 = f32x2_pack(__uint_as_float(v[g * 8 + 2 * k]), __uint_as_float(v[g * 8 + 2 * k + 1]));
        uint64_t res2 = f32x2_pack(bf16_lo(rw[k]), bf16_hi(rw[k]));
        if (rebuild) res2 = f32x2_fma(f32x2_fma(res2, rs2, nb2), g2[k], be2[k]);   // LN(resid) = ((v - mu) rstd) gamma + beta
        const uint64_t f2 = f32x2_add(f32x2_add(acc2, bias2[k]), res2);
        sum2 = f32x2_add(sum2, f2);
        sq2 = f32x2_fma(f2, f2, sq2);
        float f0, f1;
        f32x2_unpack(f2, f0, f1);
        o[k] = pack_bf16(f0, f1);
      }
      *reinterpret_cast<uint4*>(blk + slot_off(lane, g)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncwarp();
    // 4 lanes per row (4 x 16 B = the row's 64 B), 8 rows per instruction
    const int piece = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2);
      const int m = m_warp + r;
      if (m < M) {
        const uint4 o = *reinterpret_cast<const uint4*>(blk + slot_off(r, piece));
        st_global_stream(p.out + (size_t)m * kHidden + n0 + piece * 8, o);
      }
    }
    __syncwarp();
  }
  __device__ __forceinline__ void tile_end(int m_cta, int nb, int num_n, int M) {
    const int colq = ew >> 2;
    if constexpr (!kFused) {
      float s0, s1, q0, q1;
      f32x2_unpack(sum2, s0, s1);
      f32x2_unpack(sq2, q0, q1);
      if (my_row < M) st_global_stream(p.stats + ((size_t)my_row * num_n + nb) * kSplit + colq, make_float2(s0 + s1, q0 + q1));
      sum2 = 0ull;
      sq2 = 0ull;
    } else {
      if (nb != num_n - 1) return;
      // ---- the CTA's 128 rows are complete: exchange the row statistics of the two column halves ----
      float* st = sstats + parity * (kSplit * 128 * 2);
      const int row = my_row - m_cta;   // 0..127
      float s0, s1, q0, q1;
      f32x2_unpack(sum2, s0, s1);
      f32x2_unpack(sq2, q0, q1);
      st[(colq * 128 + row) * 2 + 0] = s0 + s1;
      st[(colq * 128 + row) * 2 + 1] = q0 + q1;
      sum2 = 0ull;
      sq2 = 0ull;
      // all epilogue warps: statistics visible, and every bf16 row of the panel is in global memory
      asm volatile("bar.sync 1, %0;" ::"n"(kWarps * 32) : "memory");
      // ---- normalise in place: warp ew owns 128 / kWarps rows, a lane 3 x 8 columns of a row ----
      // (columns 256 i + 8 lane .. + 8 for i = 0, 1, 2: every 16-byte access of the warp is contiguous)
      float ga[24], be[24];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int c = (i >> 1) * 256 + lane_ * 8 + (i & 1) * 4;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + c));
        ga[i * 4 + 0] = g4.x; ga[i * 4 + 1] = g4.y; ga[i * 4 + 2] = g4.z; ga[i * 4 + 3] = g4.w;
        be[i * 4 + 0] = b4.x; be[i * 4 + 1] = b4.y; be[i * 4 + 2] = b4.z; be[i * 4 + 3] = b4.w;
      }
      constexpr int kRowsPerWarp = 128 / kWarps;
#pragma unroll 4
      for (int rr_ = 0; rr_ < kRowsPerWarp; ++rr_) {
        const int r = ew * kRowsPerWarp + rr_;
        const int m = m_cta + r;
        if (m >= M) break;
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int c = 0; c < kSplit; ++c) {
          s += st[(c * 128 + r) * 2];
          q += st[(c * 128 + r) * 2 + 1];
        }
        const float mean = s * (1.f / kHidden);
        const float var = fmaxf(q * (1.f / kHidden) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.eps);
        uint4* ptr = reinterpret_cast<uint4*>(p.out + (size_t)m * kHidden + lane_ * 8);   // + 32 uint4 per i
        uint4 x[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = __ldcg(ptr + i * 32);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const uint32_t wds[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = i * 8 + k * 2;
            const float y0 = (bf16_lo(wds[k]) - mean) * rstd * ga[c] + be[c];
            const float y1 = (bf16_hi(wds[k]) - mean) * rstd * ga[c + 1] + be[c + 1];
            o[k] = pack_bf16(y0, y1);
          }
          ptr[i * 32] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      parity ^= 1;
    }
  }
  __device__ __forceinline__ void finish() {}
};

// Second half of the unfused variant: y[t] = (v[t] - mean) * rstd * gamma + beta in place, with the
// row statistics summed from the GEMM epilogue's partials in a fixed order.  A warp owns
// kLnApplyRows consecutive rows and requests all of them before touching any (HBM-bound:
// 1.5 KB read + 1.5 KB written per row, so bytes in flight are what matters).
constexpr int kLnApplyRows = 4;
static __global__ void __launch_bounds__(256)
ln_apply_kernel(__nv_bfloat16* __restrict__ x, const float2* __restrict__ stats, int T,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  const int lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kLnApplyRows;
  if (t0 >= T) return;
  // lane owns columns 256 i + 8 lane .. + 8, i = 0, 1, 2: every access of the warp is contiguous
  uint4 v[kLnApplyRows][3];
  constexpr int kPart4 = 3 * kGemmEpiColSplit / 2;   // float4s holding a row's 3 x 4 (sum, sumsq) partials
  float4 sp[kLnApplyRows][kPart4];
#pragma unroll
  for (int j = 0; j < kLnApplyRows; ++j) {
    const int t = min(t0 + j, T - 1);
    const uint4* ptr = reinterpret_cast<const uint4*>(x + (size_t)t * kHidden + lane * 8);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[j][i] = ptr[i * 32];
    // 12 partials (3 tiles x 4 column slices) = 96 B per row
    const float4* s4 = reinterpret_cast<const float4*>(stats + (size_t)t * (2 * kPart4));
#pragma unroll
    for (int i = 0; i < kPart4; ++i) sp[j][i] = __ldg(s4 + i);
  }
  float g[24], b[24];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (i >> 1) * 256 + lane * 8 + (i & 1) * 4;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
    g[i * 4 + 0] = g4.x; g[i * 4 + 1] = g4.y; g[i * 4 + 2] = g4.z; g[i * 4 + 3] = g4.w;
    b[i * 4 + 0] = b4.x; b[i * 4 + 1] = b4.y; b[i * 4 + 2] = b4.z; b[i * 4 + 3] = b4.w;
  }
#pragma unroll
  for (int j = 0; j < kLnApplyRows; ++j) {
    const int t = t0 + j;
    if (t >= T) break;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < kPart4; ++i) {   // fixed order: results are run-to-run identical
      s += sp[j][i].x + sp[j][i].z;
      q += sp[j][i].y + sp[j][i].w;
    }
    const float mean = s * (1.f / kHidden);
    const float rstd = rsqrtf(fmaxf(q * (1.f / kHidden) - mean * mean, 0.f) + eps);
    uint4* ptr = reinterpret_cast<uint4*>(x + (size_t)t * kHidden + lane * 8);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint32_t wds[4] = {v[j][i].x, v[j][i].y, v[j][i].z, v[j][i].w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = i * 8 + k * 2;
        o[k] = pack_bf16((bf16_lo(wds[k]) - mean) * rstd * g[c] + b[c], (bf16_hi(wds[k]) - mean) * rstd * g[c + 1] + b[c + 1]);
      }
      ptr[i * 32] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace enc
}  // namespace css
