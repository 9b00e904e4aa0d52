// dl_attn_fwd / dl_attn_bwd: fused scaled-dot-product attention on tcgen05 (bf16 operands, fp32
// accumulation in tensor memory) -- scores, softmax and p.v in ONE kernel, the backward recomputes
// the probabilities from the saved row log-sum-exp.  No score / probability map ever reaches HBM
// (except the raw scaled logits GuidedCrossAttention has to return, written once).
//
// Shapes served (all with head dim d in {64, 128} and Lk <= 512, i.e. the whole key range of one
// query tile fits the 512 tensor-memory columns, so the softmax needs no running rescale):
//   PGCA          1 head,  d = 128, Lq x Lk = 256 x 512 (model) or 1200 x 290 (long-sequence config)
//   paired PMMA   4 heads, d = 64,  two stacked query sets x 256 against one K/V (attention.py:44-88)
//   plain PMMA    4 heads, d = 128, 256 x 256 (attention.py:109-122)
//
// Forward: one CTA per (pair, head, query set, 128-query tile); 10 warps:
//   warp 0  TMA: Q tile + all K chunks, later V into the same shared-memory chunks
//   warp 1  tcgen05.mma issuer: S[:, chunk] = Q K_chunk^T for every 128-key chunk, then
//           O (+)= P_chunk V_chunk as the softmax warps hand over P chunks; O aliases S's first columns
//   warps 2-9  softmax: two threads per query row (64 of each chunk's 128 columns each): row max over
//           tensor memory, p = exp2(s*c - m*c) written as bf16 into a SWIZZLE_128B shared-memory A tile,
//           row sums; finally O / sum -> global, and the row's log2-sum-exp.
// Backward: one CTA per (pair, head); keys on the accumulator rows ("transposed" formulation):
//   for each 128-key tile:  for each (query set, 128-query tile):
//       S^T = K Q^T, dP^T = V dO^T                      (tensor memory, 128 columns each)
//       P^T = exp2(S^T*c - lse), dS^T = P^T (dP^T - D) * scale   -> bf16 shared-memory tiles
//       dV += P^T dO,  dK += dS^T Q                      (accumulate over the query loop)
//       dQ_tile (+)= dS K                                (dS^T tile read as an MN-major A operand)
//   dQ needs the sum over key tiles: every tile's contribution leaves as a bulk tensor operation from a
//   shared-memory staging tile -- a plain store for the first key tile, a bf16 reduce-add for the rest.
#include <mutex>

#include "../../include/druglamp_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace dl {
void count_launch(int n = 1);

namespace {

constexpr int kSoftWarps = 8;
constexpr int kAttnThreads = 64 + 32 * kSoftWarps;   // 320
constexpr int kBlk = 16384;                          // one [128 rows x 128 B] SWIZZLE_128B block

struct AttnParams {
  // forward outputs / backward inputs
  __nv_bfloat16* O;
  float* lse;
  __nv_bfloat16* raw;
  // backward
  __nv_bfloat16 *dQ, *dK, *dV;
  const float* dvec;
  long long o_ld, o_sb, o_ss;
  long long dq_ld, dq_sb, dq_ss, dk_ld, dk_sb, dv_ld, dv_sb;
  long long raw_ld;
  int B, H, S2, Lq, Lk;
  int nkc;                 // 128-key chunks
  int nqt;                 // 128-query tiles per set
  int raw_vec;             // raw rows are 4-byte aligned (even row stride): two logits per store
  int dq_acc;
  float scale, c1;         // c1 = scale * log2(e)
  uint32_t idesc_a, idesc_b, idesc_c;
};

__device__ __forceinline__ uint64_t desc_k(uint32_t blk_base, int kk) {
  // K-major operand: [rows x 128 B] blocks, 16 elements (32 B) per MMA inside a 64-element block
  return ptx::make_smem_desc(blk_base + (uint32_t)(kk >> 2) * kBlk + (uint32_t)(kk & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t blk_base, int kk) {
  // MN-major operand: the smem rows are the contraction index (16 rows = 2048 B per MMA), 64-wide
  // MN blocks kBlk apart
  return ptx::make_smem_desc(blk_base + (uint32_t)kk * 2048, kBlk, 1024);
}

// 32 fp32 -> 32 bf16 = four 16-byte chunks of a SWIZZLE_128B row (row base address `rowaddr`,
// first chunk index c0 in {0, 4}, xr = row & 7)
__device__ __forceinline__ void sts_row32(uint32_t rowaddr, int c0, int xr, const uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    ptx::sts128(rowaddr + (uint32_t)(((c0 + i) ^ xr) << 4), w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

__device__ __forceinline__ void stg_bf16x32(__nv_bfloat16* dst, const uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// =============================================================================== forward
template <int D, int NKC_MAX>
struct FwdCfg {
  static constexpr int NB = D / 64;                    // 64-column blocks per operand row
  static constexpr int CH = NB * kBlk;                 // one 128-key chunk of K (later V)
  static constexpr int KV_BYTES = NKC_MAX * CH;
  static constexpr int P_BYTES = 2 * 2 * kBlk;         // ring of two 128 x 128 bf16 P chunks (Q aliases #0)
  static constexpr int AUX_BYTES = 2048 + 256;         // row max / row sum exchange + barriers
  static constexpr int SMEM = KV_BYTES + P_BYTES + AUX_BYTES + 1024;
  static constexpr int TMEM_COLS = NKC_MAX <= 2 ? 256 : 512;
  static constexpr int MIN_CTAS = (SMEM <= 110 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int D, int NKC_MAX>
__global__ void __launch_bounds__(kAttnThreads, FwdCfg<D, NKC_MAX>::MIN_CTAS)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = FwdCfg<D, NKC_MAX>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sKV = base, sP = base + C::KV_BYTES, sQ = sP;
  float* xmax = reinterpret_cast<float*>(smem + C::KV_BYTES + C::P_BYTES);   // [2][128]
  float* xsum = xmax + 256;                                                   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + 256);
  const uint32_t bar_qk = ptx::smem_u32(bars), bar_s = bar_qk + 8, bar_v = bar_qk + 16, bar_o = bar_qk + 24;
  const uint32_t bar_p = bar_qk + 32;     // [4] P chunk written
  const uint32_t bar_pf = bar_qk + 64;    // [2] P ring buffer consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y % p.H, set = blockIdx.y / p.H, b = blockIdx.z;
  pdl_trigger();

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_o, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(bar_p + 8 * i, kSoftWarps);
    for (int i = 0; i < 2; ++i) ptx::mbar_init(bar_pf + 8 * i, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<C::TMEM_COLS>(ptx::smem_u32(tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  // Producer / issuer warps run warp-uniformly; one elected lane issues the TMA / tcgen05 instructions (as
  // single-lane divergent regions every tcgen05.mma paid an R2UR + ELECT sequence).
  if (warp == 0) {
    const bool leader = ptx::elect_one();
    // ---------------------------------------------------------------- TMA producer
    if (leader) {
      ptx::mbar_arrive_expect_tx(bar_qk, (uint32_t)(C::CH + p.nkc * C::CH));
#pragma unroll
      for (int j = 0; j < C::NB; ++j) ptx::tma_load_4d(sQ + j * kBlk, &tmQ, bar_qk, h * D + 64 * j, q0, b, set);
      for (int c = 0; c < p.nkc; ++c)
#pragma unroll
        for (int j = 0; j < C::NB; ++j)
          ptx::tma_load_4d(sKV + c * C::CH + j * kBlk, &tmK, bar_qk, h * D + 64 * j, c * 128, b, 0);
    }
    __syncwarp();
    // V replaces K once every score MMA has read it
    ptx::mbar_wait(bar_s, 0);
    if (leader) {
      ptx::mbar_arrive_expect_tx(bar_v, (uint32_t)(p.nkc * C::CH));
      for (int c = 0; c < p.nkc; ++c)
#pragma unroll
        for (int j = 0; j < C::NB; ++j)
          ptx::tma_load_4d(sKV + c * C::CH + j * kBlk, &tmV, bar_v, h * D + 64 * j, c * 128, b, 0);
    }
    __syncwarp();
  } else if (warp == 1) {
    const bool leader = ptx::elect_one();
    // ---------------------------------------------------------------- MMA issuer
    ptx::mbar_wait(bar_qk, 0);
    ptx::tc_fence_after();
    if (leader) {
      for (int c = 0; c < p.nkc; ++c)
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          ptx::mma_ss<false>(tmem + c * 128, desc_k(sQ, kk), desc_k(sKV + c * C::CH, kk), p.idesc_a,
                             (uint32_t)(kk != 0));
      ptx::mma_commit(bar_s);
    }
    __syncwarp();
    ptx::mbar_wait(bar_v, 0);
    for (int c = 0; c < p.nkc; ++c) {
      ptx::mbar_wait(bar_p + 8 * c, 0);
      ptx::tc_fence_after();
      const uint32_t sPc = sP + (c & 1) * 2 * kBlk;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          ptx::mma_ss<false>(tmem, desc_k(sPc, kk), desc_mn(sKV + c * C::CH, kk), p.idesc_b,
                             (uint32_t)((c | kk) != 0));
        ptx::mma_commit(bar_pf + 8 * (c & 1));
      }
      __syncwarp();
    }
    if (leader) ptx::mma_commit(bar_o);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int quad = warp & 3, hh = (warp - 2) >> 2;
    const int r = quad * 32 + lane, xr = r & 7;
    const int qrow = q0 + r;
    const bool row_ok = qrow < p.Lq;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    // pass A: row maximum (and the raw scaled logits, when asked for)
    float m = -INFINITY;
    for (int c = 0; c < p.nkc; ++c) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = c * 128 + hh * 64 + g * 32;
        uint32_t v[32];
        ptx::tmem_ld_32x32(tl + (uint32_t)col, v);
        ptx::tmem_ld_wait();
        const int nvalid = p.Lk - col;
        if (nvalid >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) m = fmaxf(m, __uint_as_float(v[j]));
        }
        if (p.raw != nullptr && nvalid > 0) {       // warp-uniform
          // raw scaled logits: transpose the warp's 32 x 32 block through shared memory (the second P
          // buffer is idle until pass B) so that 16 lanes write 64 contiguous bytes of ONE row --
          // thread = row would touch 32 different rows per store instruction
          uint32_t* stg = reinterpret_cast<uint32_t*>(smem + C::KV_BYTES + 2 * kBlk) + (warp - 2) * (16 * 33);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            stg[j * 33 + lane] = ptx::pack_bf16(__uint_as_float(v[2 * j]) * p.scale, __uint_as_float(v[2 * j + 1]) * p.scale);
          __syncwarp();
          const int hl = lane >> 4, wl = lane & 15;
          const int colw = col + 2 * wl;
#pragma unroll 4
          for (int it = 0; it < 16; ++it) {
            const int rr = 2 * it + hl;
            const int grow = q0 + quad * 32 + rr;
            const uint32_t word = stg[wl * 33 + rr];
            if (grow < p.Lq && colw < p.Lk) {
              __nv_bfloat16* dst = p.raw + (((long long)b * p.H + h) * p.Lq + grow) * p.raw_ld + colw;
              if (p.raw_vec && colw + 1 < p.Lk) {
                *reinterpret_cast<uint32_t*>(dst) = word;
              } else {
                reinterpret_cast<unsigned short*>(dst)[0] = (unsigned short)(word & 0xffffu);
                if (colw + 1 < p.Lk) reinterpret_cast<unsigned short*>(dst)[1] = (unsigned short)(word >> 16);
              }
            }
          }
          __syncwarp();
        }
      }
    }
    xmax[hh * 128 + r] = m;
    ptx::bar_sync(1, 32 * kSoftWarps);
    m = fmaxf(xmax[r], xmax[128 + r]);
    const float mc = m * p.c1;
    // pass B: probabilities -> shared memory (the A operand of P V), row sums
    float sum = 0.f;
    for (int c = 0; c < p.nkc; ++c) {
      if (c >= 2) ptx::mbar_wait(bar_pf + 8 * (c & 1), 0);     // chunk c-2 has been consumed
      const uint32_t prow = sP + (uint32_t)((c & 1) * 2 * kBlk + hh * kBlk + r * 128);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = c * 128 + hh * 64 + g * 32;
        uint32_t v[32];
        ptx::tmem_ld_32x32(tl + (uint32_t)col, v);
        ptx::tmem_ld_wait();
        const int nvalid = p.Lk - col;
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float p0 = ptx::ex2(fmaf(__uint_as_float(v[2 * j]), p.c1, -mc));
          float p1 = ptx::ex2(fmaf(__uint_as_float(v[2 * j + 1]), p.c1, -mc));
          if (2 * j >= nvalid) p0 = 0.f;
          if (2 * j + 1 >= nvalid) p1 = 0.f;
          sum += p0 + p1;
          w[j] = ptx::pack_bf16(p0, p1);
        }
        sts_row32(prow, g * 4, xr, w);
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_p + 8 * c);
    }
    xsum[hh * 128 + r] = sum;
    ptx::bar_sync(1, 32 * kSoftWarps);
    const float total = xsum[r] + xsum[128 + r];
    const float inv = 1.f / total;
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();
    __nv_bfloat16* orow = p.O + (long long)b * p.o_sb + (long long)qrow * p.o_ld + (long long)set * p.o_ss +
                          h * D + hh * (D / 2);
#pragma unroll
    for (int g = 0; g < D / 64; ++g) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tl + (uint32_t)(hh * (D / 2) + g * 32), v);
      ptx::tmem_ld_wait();
      if (row_ok) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          w[j] = ptx::pack_bf16(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
        stg_bf16x32(orow + g * 32, w);
      }
    }
    if (hh == 0 && row_ok)
      p.lse[(((long long)set * p.B + b) * p.H + h) * p.Lq + qrow] = mc + log2f(total);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::TMEM_COLS>(tmem);
  }
}

// =============================================================================== backward
// dvec[set, b, h, row] = sum_d dO * O over the head's columns: D/8 lanes per (row, head), one
// 16-byte load of each operand per lane, shuffle reduction inside the lane group
template <int D>
__global__ void __launch_bounds__(256)
attn_dvec_kernel(const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ dO,
                 float* __restrict__ dvec, long long o_ld, long long o_sb, long long o_ss,
                 int B, int H, int S2, int Lq) {
  pdl_trigger();
  pdl_wait();
  constexpr int G = D / 8;                           // lanes per (row, head): 8 or 16
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S2 * B * Lq * H;
  long long w = t / G;                               // (set, b, row, h) with h fastest: a row's heads are
  const int part = (int)(t % G);                     // adjacent in memory, so a warp reads contiguous bytes
  const bool ok = w < total;
  if (!ok) w = total - 1;
  const int h = (int)(w % H);
  long long u = w / H;
  const int row = (int)(u % Lq);
  u /= Lq;
  const int b = (int)(u % B), set = (int)(u / B);
  const long long off = (long long)b * o_sb + (long long)row * o_ld + (long long)set * o_ss + h * D + part * 8;
  float x[8], y[8];
  ldv(O + off, x);
  ldv(dO + off, y);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc = fmaf(x[i], y[i], acc);
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (ok && part == 0) dvec[(((long long)set * B + b) * H + h) * Lq + row] = acc;
}

template <int D>
struct BwdCfg {
  static constexpr int NB = D / 64;
  static constexpr int CH = NB * kBlk;                 // one 128-row operand tile (K, V, Q or dO)
  static constexpr int NST = D == 64 ? 2 : 1;          // Q/dO stages
  static constexpr int OFF_K = 0, OFF_V = CH, OFF_Q = 2 * CH;       // stage s: Q at OFF_Q + s*2*CH, dO after it
  static constexpr int OFF_PT = OFF_Q + NST * 2 * CH;
  static constexpr int OFF_DST = OFF_PT + 2 * kBlk;
  static constexpr int OFF_AUX = OFF_DST + 2 * kBlk;                // lse_s[2][128], dv_s[2][128], barriers
  static constexpr int SMEM = OFF_AUX + 2048 + 256 + 1024;
  static constexpr int T_ST = 0, T_DPT = 128, T_DV = 256, T_DK = 256 + D, T_DQ = 0;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQ, const AttnParams p) {
  using C = BwdCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sK = base + C::OFF_K, sV = base + C::OFF_V, sPT = base + C::OFF_PT, sDST = base + C::OFF_DST;
  float* lse_s = reinterpret_cast<float*>(smem + C::OFF_AUX);     // [2][128]
  float* dv_s = lse_s + 256;                                      // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dv_s + 256);
  const uint32_t bar_kv = ptx::smem_u32(bars);
  const uint32_t bar_qfull = bar_kv + 8;        // [2]
  const uint32_t bar_qempty = bar_kv + 24;      // [2]
  const uint32_t bar_st = bar_kv + 40, bar_pds = bar_kv + 48, bar_dq = bar_kv + 56, bar_tfree = bar_kv + 64;
  const uint32_t bar_dkv = bar_kv + 72, bar_dkvfree = bar_kv + 80;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nit = p.S2 * p.nqt;       // (set, query tile) iterations per key tile
  pdl_trigger();

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::prefetch_tmap(&tmdO);
    ptx::prefetch_tmap(&tmdQ);
    ptx::mbar_init(bar_kv, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_qfull + 8 * i, 1);
      ptx::mbar_init(bar_qempty + 8 * i, 1);
    }
    ptx::mbar_init(bar_st, 1);
    ptx::mbar_init(bar_pds, kSoftWarps);
    ptx::mbar_init(bar_dq, 1);
    ptx::mbar_init(bar_tfree, kSoftWarps);
    ptx::mbar_init(bar_dkv, 1);
    ptx::mbar_init(bar_dkvfree, kSoftWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(ptx::smem_u32(tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    {
      // ---------------------------------------------------------------- TMA producer (warp-uniform, elected lane)
      const bool leader = ptx::elect_one();
      uint32_t n = 0;
      for (int t = 0; t < p.nkc; ++t) {
        if (t > 0) ptx::mbar_wait(bar_dkv, (uint32_t)((t - 1) & 1));   // previous tile's MMAs are done with K/V
        if (leader) {
          ptx::mbar_arrive_expect_tx(bar_kv, 2 * C::CH);
#pragma unroll
          for (int j = 0; j < C::NB; ++j) {
            ptx::tma_load_4d(sK + j * kBlk, &tmK, bar_kv, h * D + 64 * j, t * 128, b, 0);
            ptx::tma_load_4d(sV + j * kBlk, &tmV, bar_kv, h * D + 64 * j, t * 128, b, 0);
          }
        }
        __syncwarp();
        for (int it = 0; it < nit; ++it, ++n) {
          const int st = (int)(n % C::NST);
          const uint32_t ph = (n / C::NST) & 1;
          ptx::mbar_wait(bar_qempty + 8 * st, ph ^ 1u);
          const uint32_t full = bar_qfull + 8 * st;
          const int set = it / p.nqt, q0 = (it % p.nqt) * 128;
          const uint32_t sQ = base + C::OFF_Q + st * 2 * C::CH, sdO = sQ + C::CH;
          if (leader) {
            ptx::mbar_arrive_expect_tx(full, 2 * C::CH);
#pragma unroll
            for (int j = 0; j < C::NB; ++j) {
              ptx::tma_load_4d(sQ + j * kBlk, &tmQ, full, h * D + 64 * j, q0, b, set);
              ptx::tma_load_4d(sdO + j * kBlk, &tmdO, full, h * D + 64 * j, q0, b, set);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------------------------------------------------------- MMA issuer (warp-uniform, elected lane)
      const bool leader = ptx::elect_one();
      uint32_t n = 0;
      for (int t = 0; t < p.nkc; ++t) {
        ptx::mbar_wait(bar_kv, (uint32_t)(t & 1));
        if (t > 0) ptx::mbar_wait(bar_dkvfree, (uint32_t)((t - 1) & 1));   // dK / dV accumulators drained
        for (int it = 0; it < nit; ++it, ++n) {
          const int st = (int)(n % C::NST);
          const uint32_t ph = (n / C::NST) & 1;
          ptx::mbar_wait(bar_qfull + 8 * st, ph);
          if (n > 0) ptx::mbar_wait(bar_tfree, (n - 1) & 1);    // S^T / dP^T / dQ columns drained
          ptx::tc_fence_after();
          const uint32_t sQ = base + C::OFF_Q + st * 2 * C::CH, sdO = sQ + C::CH;
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)          // S^T = K Q^T
              ptx::mma_ss<false>(tmem + C::T_ST, desc_k(sK, kk), desc_k(sQ, kk), p.idesc_a, (uint32_t)(kk != 0));
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)          // dP^T = V dO^T
              ptx::mma_ss<false>(tmem + C::T_DPT, desc_k(sV, kk), desc_k(sdO, kk), p.idesc_a, (uint32_t)(kk != 0));
            ptx::mma_commit(bar_st);
          }
          __syncwarp();
          ptx::mbar_wait(bar_pds, n & 1);
          ptx::tc_fence_after();
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)               // dV += P^T dO
              ptx::mma_ss<false>(tmem + C::T_DV, desc_k(sPT, kk), desc_mn(sdO, kk), p.idesc_b,
                                 (uint32_t)((it | kk) != 0));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)               // dK += dS^T Q
              ptx::mma_ss<false>(tmem + C::T_DK, desc_k(sDST, kk), desc_mn(sQ, kk), p.idesc_b,
                                 (uint32_t)((it | kk) != 0));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)               // dQ tile = dS K
              ptx::mma_ss<false>(tmem + C::T_DQ, desc_mn(sDST, kk), desc_mn(sK, kk), p.idesc_c, (uint32_t)(kk != 0));
            ptx::mma_commit(bar_qempty + 8 * st);
            ptx::mma_commit(bar_dq);
          }
          __syncwarp();
        }
        if (leader) ptx::mma_commit(bar_dkv);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int quad = warp & 3, hh = (warp - 2) >> 2;
    const int r = quad * 32 + lane, xr = r & 7;
    const int ts = threadIdx.x - 64;                      // 0..255
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t n = 0;
    for (int t = 0; t < p.nkc; ++t) {
      const int key = t * 128 + r;
      const bool key_ok = key < p.Lk;
      for (int it = 0; it < nit; ++it, ++n) {
        const int set = it / p.nqt, q0 = (it % p.nqt) * 128;
        const long long vrow = (((long long)set * p.B + b) * p.H + h) * p.Lq;   // + query row
        // per-column softmax statistics of this query tile
        float* ls = lse_s + (n & 1) * 128;
        float* ds_ = dv_s + (n & 1) * 128;
        if (ts < 128) {
          const bool ok = q0 + ts < p.Lq;
          ls[ts] = ok ? p.lse[vrow + q0 + ts] : INFINITY;
          ds_[ts] = ok ? p.dvec[vrow + q0 + ts] : 0.f;
        }
        if (ts == 0) ptx::bulk_wait_read0();     // the previous dQ store has read the P^T buffer it staged in
        ptx::bar_sync(1, 32 * kSoftWarps);
        ptx::mbar_wait(bar_st, n & 1);
        ptx::tc_fence_after();
        const uint32_t prow = sPT + (uint32_t)(hh * kBlk + r * 128);
        const uint32_t drow = sDST + (uint32_t)(hh * kBlk + r * 128);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = hh * 64 + g * 32;
          uint32_t s[32], dp[32];
          ptx::tmem_ld_32x32(tl + (uint32_t)(C::T_ST + col), s);
          ptx::tmem_ld_32x32(tl + (uint32_t)(C::T_DPT + col), dp);
          ptx::tmem_ld_wait();
          uint32_t wp[16], wd[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 l2 = *reinterpret_cast<const float2*>(ls + col + 2 * j);
            const float2 d2 = *reinterpret_cast<const float2*>(ds_ + col + 2 * j);
            float p0 = ptx::ex2(fmaf(__uint_as_float(s[2 * j]), p.c1, -l2.x));
            float p1 = ptx::ex2(fmaf(__uint_as_float(s[2 * j + 1]), p.c1, -l2.y));
            if (!key_ok) { p0 = 0.f; p1 = 0.f; }
            const float g0 = p0 * (__uint_as_float(dp[2 * j]) - d2.x) * p.scale;
            const float g1 = p1 * (__uint_as_float(dp[2 * j + 1]) - d2.y) * p.scale;
            wp[j] = ptx::pack_bf16(p0, p1);
            wd[j] = ptx::pack_bf16(g0, g1);
          }
          sts_row32(prow, g * 4, xr, wp);
          sts_row32(drow, g * 4, xr, wd);
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_pds);
        // dQ tile: rows are queries now.  TMEM -> bf16 -> the (free again) P^T buffer in the TMA layout
        // -> one bulk tensor store for the first key tile, a bf16 reduce-add (in L2) for the others and
        // for a shared query gradient: coalesced, asynchronous, no scratch round trip
        ptx::mbar_wait(bar_dq, n & 1);
        ptx::tc_fence_after();
        {
          constexpr int W = D / 2;                        // columns per thread: 32 (one block half) or 64 (a block)
          const uint32_t srow = sPT + (uint32_t)((hh * W / 64) * kBlk + r * 128);
#pragma unroll
          for (int g = 0; g < W / 32; ++g) {
            uint32_t v[32], w[16];
            ptx::tmem_ld_32x32(tl + (uint32_t)(C::T_DQ + hh * W + g * 32), v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = ptx::pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            sts_row32(srow, ((hh * W) % 64) / 8 + g * 4, xr, w);
          }
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ptx::bar_sync(2, 32 * kSoftWarps);
        if (ts == 0) {
          const bool add = t > 0 || p.dq_acc != 0;
#pragma unroll
          for (int j = 0; j < C::NB; ++j) {
            if (add) ptx::tma_reduce_add_4d(&tmdQ, sPT + j * kBlk, h * D + 64 * j, q0, b, set);
            else ptx::tma_store_4d(&tmdQ, sPT + j * kBlk, h * D + 64 * j, q0, b, set);
          }
          ptx::bulk_commit();
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_tfree);
      }
      // dK / dV of this key tile: rows are keys
      ptx::mbar_wait(bar_dkv, (uint32_t)(t & 1));
      ptx::tc_fence_after();
      __nv_bfloat16* kdst = p.dK + (long long)b * p.dk_sb + (long long)key * p.dk_ld + h * D + hh * (D / 2);
      __nv_bfloat16* vdst = p.dV + (long long)b * p.dv_sb + (long long)key * p.dv_ld + h * D + hh * (D / 2);
#pragma unroll
      for (int g = 0; g < D / 64; ++g) {
        uint32_t v[32], w[16];
        ptx::tmem_ld_32x32(tl + (uint32_t)(C::T_DV + hh * (D / 2) + g * 32), v);
        ptx::tmem_ld_wait();
        if (key_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = ptx::pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          stg_bf16x32(vdst + g * 32, w);
        }
        ptx::tmem_ld_32x32(tl + (uint32_t)(C::T_DK + hh * (D / 2) + g * 32), v);
        ptx::tmem_ld_wait();
        if (key_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = ptx::pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          stg_bf16x32(kdst + g * 32, w);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_dkvfree);
    }
    if (ts == 0) ptx::bulk_wait0();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

// =============================================================================== host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 tensor (sets, pairs, rows, cols) with unit column stride; box = 64 columns x 128 rows
int make_map(CUtensorMap* m, const void* ptr, long long cols, long long rows, long long pairs, long long sets,
             long long ld, long long sb, long long ss, const char* name) {
  DL_REQUIRE(((uintptr_t)ptr & 15) == 0, "dl_attn: %s must be 16-byte aligned", name);
  DL_REQUIRE(ld >= cols && ld % 8 == 0, "dl_attn: %s row stride %lld must be >= %lld and a multiple of 8 elements", name, ld, cols);
  DL_REQUIRE(pairs == 1 || (sb > 0 && sb % 8 == 0), "dl_attn: %s pair stride must be a positive multiple of 8 elements", name);
  DL_REQUIRE(sets == 1 || (ss > 0 && ss % 8 == 0), "dl_attn: %s set stride must be a positive multiple of 8 elements", name);
  const long long dummy = ld * rows;
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)pairs, (cuuint64_t)sets};
  cuuint64_t strides[3] = {(cuuint64_t)(ld * 2), (cuuint64_t)((pairs > 1 ? sb : dummy) * 2),
                           (cuuint64_t)((sets > 1 ? ss : dummy) * 2)};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  DL_REQUIRE(fn != nullptr, "dl_attn: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-2, "dl_attn: cuTensorMapEncodeTiled(%s) failed with CUresult %d (cols %lld rows %lld pairs %lld sets %lld ld %lld)",
                     name, (int)r, cols, rows, pairs, sets, ld);
  return 0;
}

int check_common(const dl_attn_args* a) {
  DL_REQUIRE(a != nullptr, "dl_attn: null args");
  DL_REQUIRE(a->q && a->k && a->v && a->o && a->lse, "dl_attn: q, k, v, o and lse must be non-null");
  DL_REQUIRE(a->d == 64 || a->d == 128, "dl_attn: head dim must be 64 or 128 (got %lld)", (long long)a->d);
  DL_REQUIRE(a->B >= 1 && a->H >= 1 && a->S2 >= 1 && a->Lq >= 1 && a->Lk >= 1, "dl_attn: bad extents");
  DL_REQUIRE(a->Lk <= 512, "dl_attn: at most 512 keys (got %lld)", (long long)a->Lk);
  DL_REQUIRE(a->B <= 65535 && a->H * a->S2 <= 65535, "dl_attn: grid too large");
  DL_REQUIRE(a->o_ld % 8 == 0 && a->o_sb % 8 == 0 && a->o_ss % 8 == 0 && ((uintptr_t)a->o & 15) == 0,
             "dl_attn: o needs 16-byte aligned rows");
  return 0;
}

void fill_common(AttnParams& p, const dl_attn_args* a) {
  p.O = (__nv_bfloat16*)a->o;
  p.lse = a->lse;
  p.raw = (__nv_bfloat16*)a->raw;
  p.dQ = (__nv_bfloat16*)a->dq; p.dK = (__nv_bfloat16*)a->dk; p.dV = (__nv_bfloat16*)a->dv;
  p.dvec = a->dvec;
  p.o_ld = a->o_ld; p.o_sb = a->o_sb; p.o_ss = a->o_ss;
  p.dq_ld = a->dq_ld; p.dq_sb = a->dq_sb; p.dq_ss = a->dq_ss;
  p.dk_ld = a->dk_ld; p.dk_sb = a->dk_sb; p.dv_ld = a->dv_ld; p.dv_sb = a->dv_sb;
  p.raw_ld = a->raw_ld;
  p.B = (int)a->B; p.H = (int)a->H; p.S2 = (int)a->S2; p.Lq = (int)a->Lq; p.Lk = (int)a->Lk;
  p.nkc = (int)((a->Lk + 127) / 128);
  p.nqt = (int)((a->Lq + 127) / 128);
  p.raw_vec = a->raw != nullptr && a->raw_ld % 2 == 0 && ((uintptr_t)a->raw & 3) == 0;
  p.dq_acc = a->dq_accumulate;
  p.scale = a->scale;
  p.c1 = a->scale * 1.4426950408889634f;
}

template <int D, int NKC_MAX>
int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, AttnParams& p,
               cudaStream_t stream) {
  using C = FwdCfg<D, NKC_MAX>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_fwd_kernel<D, NKC_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  });
  if (attr_err != cudaSuccess)
    return set_error((int)attr_err, "dl_attn_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  p.idesc_a = ptx::make_idesc(false, false, false, 128, 128);   // S = Q K^T
  p.idesc_b = ptx::make_idesc(false, false, true, 128, D);      // O = P V
  p.idesc_c = 0;
  dim3 grid((unsigned)p.nqt, (unsigned)(p.H * p.S2), (unsigned)p.B);
  DL_LAUNCH((attn_fwd_kernel<D, NKC_MAX>), grid, kAttnThreads, C::SMEM, stream, tq, tk, tv, p);
  DL_LAUNCH_CHECK("attn_fwd_kernel");
  count_launch();
  return 0;
}

template <int D>
int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
               const CUtensorMap& tdq, AttnParams& p, const dl_attn_args* a, cudaStream_t stream) {
  using C = BwdCfg<D>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  });
  if (attr_err != cudaSuccess)
    return set_error((int)attr_err, "dl_attn_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  const long long rows = (long long)p.S2 * p.B * p.H * p.Lq;
  const int threads = 256;
  DL_LAUNCH((attn_dvec_kernel<D>), ceil_div(rows * (D / 8), threads), threads, 0, stream,
            (const __nv_bfloat16*)a->o, (const __nv_bfloat16*)a->d_o, a->dvec, p.o_ld, p.o_sb, p.o_ss, p.B, p.H,
            p.S2, p.Lq);
  DL_LAUNCH_CHECK("attn_dvec_kernel");
  p.idesc_a = ptx::make_idesc(false, false, false, 128, 128);   // S^T = K Q^T, dP^T = V dO^T
  p.idesc_b = ptx::make_idesc(false, false, true, 128, D);      // dV += P^T dO, dK += dS^T Q
  p.idesc_c = ptx::make_idesc(false, true, true, 128, D);       // dQ = dS K
  dim3 grid((unsigned)p.H, (unsigned)p.B, 1);
  DL_LAUNCH((attn_bwd_kernel<D>), grid, kAttnThreads, C::SMEM, stream, tq, tk, tv, tdo, tdq, p);
  DL_LAUNCH_CHECK("attn_bwd_kernel");
  count_launch(2);
  return 0;
}

}  // namespace
}  // namespace dl

extern "C" int dl_attn_fwd(const dl_attn_args* a, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_common(a);
  if (rc) return rc;
  DL_REQUIRE(a->raw == nullptr || a->S2 == 1, "dl_attn_fwd: raw logits need a single query set");
  DL_REQUIRE(a->raw == nullptr || a->raw_ld >= a->Lk, "dl_attn_fwd: raw_ld < Lk");
  const long long cols = a->H * a->d;
  CUtensorMap tq, tk, tv;
  if ((rc = make_map(&tq, a->q, cols, a->Lq, a->B, a->S2, a->q_ld, a->q_sb, a->q_ss, "q"))) return rc;
  if ((rc = make_map(&tk, a->k, cols, a->Lk, a->B, 1, a->k_ld, a->k_sb, 0, "k"))) return rc;
  if ((rc = make_map(&tv, a->v, cols, a->Lk, a->B, 1, a->v_ld, a->v_sb, 0, "v"))) return rc;
  AttnParams p;
  fill_common(p, a);
  if (a->d == 64) return p.nkc <= 2 ? launch_fwd<64, 2>(tq, tk, tv, p, stream) : launch_fwd<64, 4>(tq, tk, tv, p, stream);
  return p.nkc <= 2 ? launch_fwd<128, 2>(tq, tk, tv, p, stream) : launch_fwd<128, 4>(tq, tk, tv, p, stream);
}

extern "C" int dl_attn_bwd(const dl_attn_args* a, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_common(a);
  if (rc) return rc;
  DL_REQUIRE(a->d_o && a->dq && a->dk && a->dv && a->dvec, "dl_attn_bwd: d_o, dq, dk, dv and dvec must be non-null");
  auto ok16 = [](const void* q, long long ld, long long sb, long long ss) {
    return ((uintptr_t)q & 15) == 0 && ld % 8 == 0 && sb % 8 == 0 && ss % 8 == 0;
  };
  DL_REQUIRE(ok16(a->dq, a->dq_ld, a->dq_sb, a->dq_ss) && ok16(a->dk, a->dk_ld, a->dk_sb, 0) &&
             ok16(a->dv, a->dv_ld, a->dv_sb, 0), "dl_attn_bwd: dq / dk / dv need 16-byte aligned rows");
  const long long cols = a->H * a->d;
  CUtensorMap tq, tk, tv, tdo, tdq;
  if ((rc = make_map(&tq, a->q, cols, a->Lq, a->B, a->S2, a->q_ld, a->q_sb, a->q_ss, "q"))) return rc;
  if ((rc = make_map(&tk, a->k, cols, a->Lk, a->B, 1, a->k_ld, a->k_sb, 0, "k"))) return rc;
  if ((rc = make_map(&tv, a->v, cols, a->Lk, a->B, 1, a->v_ld, a->v_sb, 0, "v"))) return rc;
  if ((rc = make_map(&tdo, a->d_o, cols, a->Lq, a->B, a->S2, a->o_ld, a->o_sb, a->o_ss, "d_o"))) return rc;
  if ((rc = make_map(&tdq, a->dq, cols, a->Lq, a->B, a->S2, a->dq_ld, a->dq_sb, a->dq_ss, "dq"))) return rc;
  AttnParams p;
  fill_common(p, a);
  if (a->d == 64) return launch_bwd<64>(tq, tk, tv, tdo, tdq, p, a, stream);
  return launch_bwd<128>(tq, tk, tv, tdo, tdq, p, a, stream);
}
