// dl_ffn_fwd / dl_ffn_bwd: the position-wise feed-forward network of a PMMA block (model/PMMA/mlp.py:44-50
// with the block's residual add, model/PMMA/block.py:45-47) as ONE tcgen05 kernel per direction: two chained
// GEMMs with the 4x-wide hidden activation kept on chip between them.
//
//   forward    H = dropout(gelu(X W1^T + b1))            [M, Dh]   (also stored: the backward's dW2 operand)
//              Y = dropout(H W2^T + b2) + residual        [M, 256]
//   backward   dP = (G W2) * dact                         [M, Dh]   (stored: the backward's dW1 operand)
//              dX = dP W1                                 [M, 256]
//
// Both are  E_c = epi1(A B1_c),  Y += E_c B2_c  over 128-column chunks c of the hidden dimension, for one
// 128-row tile of A per CTA:
//   tensor memory   two 128-column buffers for the chunk accumulator (the epilogue of chunk c overlaps the
//                   first GEMM of chunk c+1) + 256 columns for Y  = all 512 columns
//   shared memory   A tile (64 KB, resident for the whole row tile), a ring of three 32 KB weight units
//                   (half a chunk of B1 or B2 each), two 32 KB E tiles (bf16, SWIZZLE_128B K-major: the A
//                   operand of the second GEMM AND the source of the bulk tensor store that writes E to HBM)
//   warps           0: TMA producer   1: tcgen05.mma issuer   2..17: epilogue (four per TMEM lane quadrant,
//                   32 of a chunk's 128 columns each)   18: bulk tensor stores of the E tiles
// The hidden activation therefore never returns from HBM for the second GEMM, its store costs no LSU
// instructions, and one launch replaces two (forward) / two (backward) dl_gemm launches.
#include <mutex>
#include <stdlib.h>

#include "../../include/druglamp_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace dl {
void count_launch(int n = 1);

namespace {

constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps + 32;   // 608
constexpr int kD = 256;                               // model width: K of the first GEMM, N of the second
constexpr int kCh = 128;                              // hidden columns per chunk
constexpr int kBlk = 16384;                           // one [128 rows x 128 B] SWIZZLE_128B block
constexpr int kUnit = 32768;                          // ring unit: half a chunk of B1 or of B2
constexpr int kStages = 3;
constexpr int kABytes = 4 * kBlk;                     // 128 x 256 bf16
constexpr int kEBytes = 2 * kBlk;                     // 128 x 128 bf16
constexpr int kBarBytes = 256;
constexpr int kSmem = kABytes + kStages * kUnit + 2 * kEBytes + kBarBytes + 1024;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

struct FfnParams {
  const float* bias1;            // forward: [Dh]
  const float* bias2;            // forward: [256]
  __nv_bfloat16* dact;           // forward: d E / d pre-activation out (or NULL); backward: multiplier in
  const __nv_bfloat16* res;      // forward: residual (or NULL)
  __nv_bfloat16* Y;
  long long ldh, ldy, ldr;
  unsigned long long seed1, seed2;
  const long long* drop_step;
  float drop_p;
  int M, Dh, NC, tiles;
  int store_e;                   // E is written to HBM (training); 0 = forward-only scoring
  uint32_t idesc1, idesc2;
};

// MODE 0: forward (K-major weights, bias + GELU (+ derivative) + dropout, then bias + dropout + residual)
// MODE 1: backward (MN-major weights, multiply by the stored derivative, plain dX)
// PAIR: two CTAs (a cluster of two) own two adjacent 128-row tiles as ONE 256-row tcgen05.mma.cta_group::2
// problem.  Each CTA keeps its own A and E tiles but loads only HALF of every weight chunk (64 of the 128
// hidden columns of B1, 128 of the 256 output columns of B2); the leader CTA's warp issues every MMA for both.
// The weight traffic per CTA -- what bounds this kernel: 1 MB for a 128-row tile through a 96 KB ring -- halves,
// and a 32 KB ring unit now holds a whole chunk operand (two units per chunk instead of four).
template <int MODE, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
ffn_chain_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmE,
                 const FfnParams p) {
  constexpr bool BMN = MODE == 1;
  constexpr int UPC = PAIR ? 1 : 2;                      // ring units per chunk operand
  constexpr int NM1 = 16 / UPC, NM2 = 8 / UPC;           // MMAs per unit: first / second GEMM
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sA = base, sRing = base + kABytes, sE = sRing + kStages * kUnit;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kABytes + kStages * kUnit + 2 * kEBytes);
  const uint32_t bar_afull = ptx::smem_u32(bars), bar_aempty = bar_afull + 8;
  const uint32_t bar_full = bar_afull + 16;               // [3] weight unit landed
  const uint32_t bar_empty = bar_full + 8 * kStages;      // [3] weight unit consumed
  const uint32_t bar_sfull = bar_empty + 8 * kStages;     // [2] chunk accumulator ready
  const uint32_t bar_sempty = bar_sfull + 16;             // [2] chunk accumulator drained
  const uint32_t bar_efull = bar_sempty + 16;             // [2] E tile written
  const uint32_t bar_eempty = bar_efull + 16;             // [2] E tile consumed (second GEMM + bulk store)
  const uint32_t bar_yfull = bar_eempty + 16, bar_yempty = bar_yfull + 8;
  const uint32_t bar_efullm = bar_yempty + 8;             // [2] pair: E tiles of BOTH CTAs written (leader's issuer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * kStages + 8 + 2 + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)ptx::cluster_ctarank() : 0;
  // pair: both CTAs walk the same sequence of tile pairs; this CTA's rows are tile 2 * pair + rank
  const int tile_first = PAIR ? 2 * (int)(blockIdx.x >> 1) + rank : (int)blockIdx.x;
  const int tile_step = (int)gridDim.x;
  const int tile_end = PAIR ? 2 * ((p.tiles + 1) / 2) : p.tiles;     // a phantom last tile keeps the pair in step
  constexpr uint32_t kArr = PAIR ? 2 : 1;                 // arrivals per barrier that both CTAs feed
  pdl_trigger();

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB1);
    ptx::prefetch_tmap(&tmB2);
    if (p.store_e) ptx::prefetch_tmap(&tmE);
    ptx::mbar_init(bar_afull, kArr);                      // pair: leader's expect_tx + the peer producer's arrival
    ptx::mbar_init(bar_aempty, 1);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, kArr);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_sfull + 8 * i, 1);
      ptx::mbar_init(bar_sempty + 8 * i, kArr * kEpiWarps);   // pair: the epilogue warps of both CTAs, at the leader
      ptx::mbar_init(bar_efull + 8 * i, kEpiWarps);           // this CTA's E tile (its bulk-store warp)
      ptx::mbar_init(bar_efullm + 8 * i, kArr * kEpiWarps);
      ptx::mbar_init(bar_eempty + 8 * i, p.store_e ? 2 : 1);
    }
    ptx::mbar_init(bar_yfull, 1);
    ptx::mbar_init(bar_yempty, kArr * kEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) ptx::tmem_alloc_2sm<512>(ptx::smem_u32(tmem_slot));
    else ptx::tmem_alloc<512>(ptx::smem_u32(tmem_slot));
  }
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync_all();
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const int NC = p.NC;

  // Producer and issuer loops run warp-uniformly (all lanes wait and keep the loop state in uniform registers,
  // one elected lane issues): as single-lane divergent regions every tcgen05.mma paid an R2UR + ELECT sequence.
  if (warp == 0) {
    {
      // ------------------------------------------------------------------ TMA producer
      const bool leader = ptx::elect_one();
      uint32_t s = 0, ph = 0, ti = 0;
      auto slot_next = [&]() { if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; } };
      // pair: every load of either CTA is counted on the LEADER CTA's barrier
      auto ld = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int x0, int x1) {
        if constexpr (PAIR) ptx::tma_load_2d_2sm(dst, m, bar, x0, x1);
        else ptx::tma_load_2d(dst, m, bar, x0, x1);
      };
      auto begin_fill = [&](uint32_t bar, uint32_t bytes) -> uint32_t {       // -> the barrier the loads signal
        if constexpr (PAIR) {
          const uint32_t lbar = ptx::mapa(bar, 0);
          if (rank == 0) ptx::mbar_arrive_expect_tx(bar, 2 * bytes);
          else ptx::mbar_arrive_remote(lbar);
          return lbar;
        } else {
          ptx::mbar_arrive_expect_tx(bar, bytes);
          return bar;
        }
      };
      auto load_b1 = [&](int c) {          // chunk c of the first GEMM's weights
        for (int u = 0; u < UPC; ++u) {
          ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t dst = sRing + s * kUnit;
          if (leader) {
            const uint32_t full = begin_fill(bar_full + 8 * s, kUnit);
            if constexpr (!BMN && !PAIR) {        // [Dh, 256] K-major: K half u as two [128 rows x 64 k] blocks
#pragma unroll
              for (int j = 0; j < 2; ++j) ld(dst + j * kBlk, &tmB1, full, (2 * u + j) * 64, c * kCh);
            } else if constexpr (!BMN) {          // pair: this CTA's 64 hidden columns, four [64 rows x 64 k] blocks
#pragma unroll
              for (int j = 0; j < 4; ++j) ld(dst + j * 8192, &tmB1, full, j * 64, c * kCh + rank * 64);
            } else if constexpr (!PAIR) {         // [256 (k), Dh] MN-major: per 64-wide MN block, 128 k rows
#pragma unroll
              for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
                  ld(dst + blk * kBlk + kq * 8192, &tmB1, full, c * kCh + blk * 64, u * 128 + kq * 64);
            } else {                              // pair: this CTA's 64-wide MN block, all 256 k rows
#pragma unroll
              for (int kq = 0; kq < 4; ++kq) ld(dst + kq * 8192, &tmB1, full, c * kCh + rank * 64, kq * 64);
            }
          }
          __syncwarp();
          slot_next();
        }
      };
      auto load_b2 = [&](int c) {          // chunk c of the second GEMM's weights
        for (int u = 0; u < UPC; ++u) {
          ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t dst = sRing + s * kUnit;
          if (leader) {
            const uint32_t full = begin_fill(bar_full + 8 * s, kUnit);
            if constexpr (!BMN && !PAIR) {        // [256, Dh] K-major: one [256 rows x 64 k] block
              ld(dst, &tmB2, full, c * kCh + u * 64, 0);
            } else if constexpr (!BMN) {          // pair: this CTA's 128 output rows, two [128 rows x 64 k] blocks
#pragma unroll
              for (int j = 0; j < 2; ++j) ld(dst + j * kBlk, &tmB2, full, c * kCh + j * 64, rank * 128);
            } else if constexpr (!PAIR) {         // [Dh (k), 256] MN-major: four 64-wide MN blocks of 64 k rows
#pragma unroll
              for (int blk = 0; blk < 4; ++blk) ld(dst + blk * 8192, &tmB2, full, blk * 64, c * kCh + u * 64);
            } else {                              // pair: this CTA's two 64-wide MN blocks of 128 k rows
#pragma unroll
              for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                for (int kq = 0; kq < 2; ++kq)
                  ld(dst + blk * kBlk + kq * 8192, &tmB2, full, rank * 128 + blk * 64, c * kCh + kq * 64);
            }
          }
          __syncwarp();
          slot_next();
        }
      };
      for (int t = tile_first; t < tile_end; t += tile_step, ++ti) {
        ptx::mbar_wait(bar_aempty, (ti & 1) ^ 1u);
        if (leader) {
          const uint32_t afull = begin_fill(bar_afull, kABytes);
#pragma unroll
          for (int j = 0; j < 4; ++j) ld(sA + j * kBlk, &tmA, afull, j * 64, t * 128);
        }
        __syncwarp();
        load_b1(0);
        if (NC > 1) load_b1(1);
        for (int c = 0; c < NC; ++c) {          // the order the issuer consumes them in
          if (c + 2 < NC) load_b1(c + 2);
          load_b2(c);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ------------------------------------------------------------------ MMA issuer (pair: the leader, for both)
      // descriptor templates: only the 14-bit start-address field (16-byte units) changes per unit / K step
      const uint64_t k_tmpl = ptx::make_smem_desc(0, 16, 1024);                                      // K-major operands
      const uint64_t b1_tmpl = BMN ? ptx::make_smem_desc(0, kBlk, 1024) : k_tmpl;                    // first GEMM's weights
      const uint64_t b2_tmpl = BMN ? ptx::make_smem_desc(0, PAIR ? kBlk : 8192, 1024) : k_tmpl;      // second GEMM's weights
      constexpr uint32_t kStepMn = 2048 >> 4, kStepK = 32 >> 4, kBlkU = kBlk >> 4;
      constexpr uint32_t kB1BlkU = (PAIR ? 8192 : kBlk) >> 4;       // K-major B1: [64 | 128 rows x 128 B] blocks
      auto addr = [](uint32_t a) { return (uint64_t)((a & 0x3FFFF) >> 4); };
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) ptx::mma_ss_2sm(d, ad, bd, idesc, acc);
        else ptx::mma_ss<false>(d, ad, bd, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {                              // pair: arrives in both CTAs
        if constexpr (PAIR) ptx::mma_commit_2sm(bar);
        else ptx::mma_commit(bar);
      };
      const bool leader = ptx::elect_one();
      uint32_t s = 0, ph = 0, g1 = 0, g2 = 0, ti = 0;
      auto slot_next = [&]() { if (++s == (uint32_t)kStages) { s = 0; ph ^= 1u; } };
      auto mma1 = [&]() {                  // chunk accumulator (g1 & 1) = A B1_chunk, K = 256
        const uint32_t sb = g1 & 1, sph = (g1 >> 1) & 1;
        ptx::mbar_wait(bar_sempty + 8 * sb, sph ^ 1u);
        ptx::tc_fence_after();
        const uint32_t acc = tmem + sb * kCh;
        for (int u = 0; u < UPC; ++u) {
          ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::tc_fence_after();
          const uint64_t a0 = k_tmpl | addr(sA + (uint32_t)u * (NM1 / 4) * kBlk), b0 = b1_tmpl | addr(sRing + s * kUnit);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < NM1; ++kk) {
              const uint64_t ad = a0 + (uint64_t)((kk >> 2) * kBlkU + (kk & 3) * kStepK);
              const uint64_t bd = b0 + (uint64_t)(BMN ? kk * kStepMn : (kk >> 2) * kB1BlkU + (kk & 3) * kStepK);
              mma(acc, ad, bd, p.idesc1, (uint32_t)((u | kk) != 0));
            }
            commit(bar_empty + 8 * s);
          }
          __syncwarp();
          slot_next();
        }
        if (leader) commit(bar_sfull + 8 * sb);
        __syncwarp();
        ++g1;
      };
      for (int t = tile_first; t < tile_end; t += tile_step, ++ti) {
        ptx::mbar_wait(bar_afull, ti & 1);
        ptx::tc_fence_after();
        int issued = 0;
        mma1(); ++issued;
        if (NC > 1) { mma1(); ++issued; }
        if (issued == NC && leader) commit(bar_aempty);
        for (int c = 0; c < NC; ++c) {
          // chunk c + 2's first GEMM goes first: its accumulator buffer is free as soon as the epilogue of
          // chunk c has READ tensor memory, long before that epilogue hands over its E tile -- the weight
          // ring keeps flowing instead of holding B2(c) while nothing can be issued
          if (c + 2 < NC) {
            mma1();
            if (++issued == NC && leader) commit(bar_aempty);
          }
          const uint32_t sb = g2 & 1, eph = (g2 >> 1) & 1;
          ptx::mbar_wait((PAIR ? bar_efullm : bar_efull) + 8 * sb, eph);     // pair: the E tiles of both CTAs
          if (c == 0) ptx::mbar_wait(bar_yempty, (ti & 1) ^ 1u);      // the previous tile's Y has been read
          ptx::tc_fence_after();
          const uint32_t eb = sE + sb * kEBytes;
          for (int u = 0; u < UPC; ++u) {
            ptx::mbar_wait(bar_full + 8 * s, ph);
            ptx::tc_fence_after();
            const uint64_t a0 = k_tmpl | addr(eb + (uint32_t)u * (PAIR ? 0u : (uint32_t)kBlk)), b0 = b2_tmpl | addr(sRing + s * kUnit);
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < NM2; ++kk) {
                // A: the E tile, two [128 rows x 128 B] blocks of 64 hidden columns each
                const uint64_t ad = a0 + (uint64_t)((kk >> 2) * kBlkU + (kk & 3) * kStepK);
                // B: K-major [rows x 128 B] blocks (kBlk apart in pair mode) or MN-major 16-row K steps
                const uint64_t bd = b0 + (uint64_t)(BMN ? kk * kStepMn : (kk >> 2) * kBlkU + (kk & 3) * kStepK);
                mma(tmem + 2 * kCh, ad, bd, p.idesc2, (uint32_t)((c | u | kk) != 0));
              }
              commit(bar_empty + 8 * s);
            }
            __syncwarp();
            slot_next();
          }
          if (leader) commit(bar_eempty + 8 * sb);
          __syncwarp();
          ++g2;
        }
        if (leader) commit(bar_yfull);
        __syncwarp();
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // -------------------------------------------------------------------- epilogue warps
    const int we = warp - 2;
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = we >> 2;                     // which 32-column slice of a chunk / 64-column slice of Y
    const int r = q * 32 + lane, xr = r & 7;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t drop_thr = drop_threshold(p.drop_p);
    const float drop_inv = p.drop_p > 0.f ? drop_scale(drop_thr) : 1.f;
    const unsigned long long seed1 = drop_seed_at(p.seed1, p.drop_p > 0.f ? p.drop_step : nullptr);
    const unsigned long long seed2 = drop_seed_at(p.seed2, p.drop_p > 0.f ? p.drop_step : nullptr);
    const bool deriv = MODE == 0 && p.dact != nullptr;
    // barriers the issuer waits on live in the leader CTA (pair: the peer's warps arrive remotely)
    auto arrive_issuer = [&](uint32_t bar) {
      if constexpr (PAIR) ptx::mbar_arrive_remote(ptx::mapa(bar, 0));
      else ptx::mbar_arrive(bar);
    };
    uint32_t g = 0, ti = 0;
    for (int t = tile_first; t < tile_end; t += tile_step, ++ti) {
      const int row = t * 128 + r;
      const bool row_ok = row < p.M;
      for (int c = 0; c < NC; ++c, ++g) {
        const uint32_t sb = g & 1, ph = (g >> 1) & 1;
        const int hcol = c * kCh + half * 32;     // first hidden column of this warp's slice
        uint32_t aw[16];                          // backward: the slice of the stored derivative (32 bf16)
        if constexpr (MODE == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) aw[i] = 0u;
          if (row_ok) {
            const __nv_bfloat16* ap = p.dact + (long long)row * p.ldh + hcol;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
              asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(aw[8 * h2]), "=r"(aw[8 * h2 + 1]), "=r"(aw[8 * h2 + 2]), "=r"(aw[8 * h2 + 3]),
                             "=r"(aw[8 * h2 + 4]), "=r"(aw[8 * h2 + 5]), "=r"(aw[8 * h2 + 6]), "=r"(aw[8 * h2 + 7])
                           : "l"(ap + 16 * h2));
          }
        }
        if constexpr (MODE == 0) {
          // the next chunk's 32 bias values (one 128-byte line per warp) on their way into L1 while this chunk is
          // processed: their L2 latency sat exposed in front of every chunk's GELU
          if (lane == 0 && c + 1 < NC) asm volatile("prefetch.global.L1 [%0];" :: "l"(p.bias1 + hcol + kCh));
        }
        ptx::mbar_wait(bar_sfull + 8 * sb, ph);
        ptx::tc_fence_after();
        uint32_t vv[32];
        ptx::tmem_ld_32x32(tl + sb * kCh + (uint32_t)(half * 32), vv);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();                   // the accumulator buffer is free: chunk c + 2 may overwrite it
        __syncwarp();
        if (lane == 0) arrive_issuer(bar_sempty + 8 * sb);
        ptx::mbar_wait(bar_eempty + 8 * sb, ph ^ 1u);
        const uint32_t erow = sE + sb * kEBytes + (uint32_t)(half >> 1) * kBlk + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t* v = vv + 16 * j;
          float x[16];
          if constexpr (MODE == 0) {
            float b[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 f = reinterpret_cast<const float4*>(p.bias1 + hcol + 16 * j)[i];
              b[4 * i] = f.x; b[4 * i + 1] = f.y; b[4 * i + 2] = f.z; b[4 * i + 3] = f.w;
            }
            float dv[16];
            if (deriv) {
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                float2 y2, d2;
                gelu_fwd_grad2<__nv_bfloat16>(make_float2(__uint_as_float(v[i]) + b[i], __uint_as_float(v[i + 1]) + b[i + 1]), y2, d2);
                x[i] = y2.x; x[i + 1] = y2.y; dv[i] = d2.x; dv[i + 1] = d2.y;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                const float2 y2 = gelu_fwd2<__nv_bfloat16>(make_float2(__uint_as_float(v[i]) + b[i], __uint_as_float(v[i + 1]) + b[i + 1]));
                x[i] = y2.x; x[i + 1] = y2.y;
              }
            }
            if (p.drop_p > 0.f) {
              const unsigned long long e = (unsigned long long)row * p.Dh + (hcol + 16 * j);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t h = drop_hash(seed1, (e >> 1) + k);
                const float2 m = make_float2((h & 0xffffu) >= drop_thr ? drop_inv : 0.f, (h >> 16) >= drop_thr ? drop_inv : 0.f);
                const float2 xm = f2_mul(make_float2(x[2 * k], x[2 * k + 1]), m);
                x[2 * k] = xm.x; x[2 * k + 1] = xm.y;
                if (deriv) {
                  const float2 dm = f2_mul(make_float2(dv[2 * k], dv[2 * k + 1]), m);
                  dv[2 * k] = dm.x; dv[2 * k + 1] = dm.y;
                }
              }
            }
            if (deriv && row_ok) {
              uint32_t w[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) w[k] = ptx::pack_bf16(dv[2 * k], dv[2 * k + 1]);
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                           :: "l"(p.dact + (long long)row * p.ldh + hcol + 16 * j), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                              "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t a2 = aw[8 * j + k];
              x[2 * k] = __uint_as_float(v[2 * k]) * __uint_as_float(a2 << 16);
              x[2 * k + 1] = __uint_as_float(v[2 * k + 1]) * __uint_as_float(a2 & 0xffff0000u);
            }
          }
          uint32_t w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) w[k] = ptx::pack_bf16(x[2 * k], x[2 * k + 1]);
          const int ci = (half & 1) * 4 + 2 * j;       // 16-byte chunk of the row's 128-byte block
          ptx::sts128(erow + (uint32_t)((ci ^ xr) << 4), w[0], w[1], w[2], w[3]);
          ptx::sts128(erow + (uint32_t)(((ci + 1) ^ xr) << 4), w[4], w[5], w[6], w[7]);
        }
        ptx::fence_proxy_async();                 // generic-proxy writes -> visible to the MMA and the bulk store
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(bar_efull + 8 * sb);                       // this CTA's bulk-store warp (and the single-CTA issuer)
          if constexpr (PAIR) arrive_issuer(bar_efullm + 8 * sb);    // the leader's issuer reads both CTAs' E tiles
        }
      }
      // ---- Y: 32 rows x 64 columns per warp
      // the four residual segments of this thread's row (one 128-byte line) are requested before the wait for
      // the last GEMM: their HBM latency hides behind it instead of sitting in front of every 16 columns
      uint32_t rw4[4][8];
      if constexpr (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k) rw4[j][k] = 0u;
        if (p.res != nullptr && row_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(rw4[j][0]), "=r"(rw4[j][1]), "=r"(rw4[j][2]), "=r"(rw4[j][3]), "=r"(rw4[j][4]), "=r"(rw4[j][5]),
                           "=r"(rw4[j][6]), "=r"(rw4[j][7])
                         : "l"(p.res + (long long)row * p.ldr + half * 64 + 16 * j));
        }
      }
      ptx::mbar_wait(bar_yfull, ti & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = half * 64 + 16 * j;
        uint32_t v[16];
        ptx::tmem_ld_32x16(tl + (uint32_t)(2 * kCh + col), v);
        ptx::tmem_ld_wait();
        if (!row_ok) continue;
        float x[16];
        if constexpr (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = reinterpret_cast<const float4*>(p.bias2 + col)[i];
            x[4 * i] = __uint_as_float(v[4 * i]) + f.x; x[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + f.y;
            x[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + f.z; x[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + f.w;
          }
          if (p.drop_p > 0.f) {
            const unsigned long long e = (unsigned long long)row * kD + col;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t h = drop_hash(seed2, (e >> 1) + k);
              x[2 * k] *= (h & 0xffffu) >= drop_thr ? drop_inv : 0.f;
              x[2 * k + 1] *= (h >> 16) >= drop_thr ? drop_inv : 0.f;
            }
          }
          if (p.res != nullptr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              x[2 * k] += __uint_as_float(rw4[j][k] << 16);
              x[2 * k + 1] += __uint_as_float(rw4[j][k] & 0xffff0000u);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
        }
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = ptx::pack_bf16(x[2 * k], x[2 * k + 1]);
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p.Y + (long long)row * p.ldy + col), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                        "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_issuer(bar_yempty);
    }
  } else {
    // -------------------------------------------------------------------- E tiles -> HBM (bulk tensor stores)
    if (lane == 0 && p.store_e) {
      uint32_t g = 0;
      for (int t = tile_first; t < tile_end; t += tile_step) {
        for (int c = 0; c < NC; ++c, ++g) {
          const uint32_t sb = g & 1, ph = (g >> 1) & 1;
          ptx::mbar_wait(bar_efull + 8 * sb, ph);
          const uint32_t eb = sE + sb * kEBytes;
          ptx::tma_store_2d(&tmE, eb, c * kCh, t * 128);
          ptx::tma_store_2d(&tmE, eb + kBlk, c * kCh + 64, t * 128);
          ptx::bulk_commit();
          ptx::bulk_wait_read0();                 // the tile may be overwritten once it has been read
          ptx::mbar_arrive(bar_eempty + 8 * sb);
        }
      }
      ptx::bulk_wait0();
    }
  }
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync_all();      // neither CTA's shared / tensor memory may go while the other uses it
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_2sm<512>(tmem);
    else ptx::tmem_dealloc<512>(tmem);
  }
}

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

// bf16 matrix [rows, cols] with unit column stride; box = box_c columns x box_r rows, SWIZZLE_128B
int make_map(CUtensorMap* m, const void* ptr, long long cols, long long rows, long long ld, int box_c, int box_r,
             const char* name) {
  DL_REQUIRE(((uintptr_t)ptr & 15) == 0, "dl_ffn: %s must be 16-byte aligned", name);
  DL_REQUIRE(ld >= cols && ld % 8 == 0, "dl_ffn: %s row stride %lld must be >= %lld and a multiple of 8 elements", name, ld, cols);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn fn = encode_fn();
  DL_REQUIRE(fn != nullptr, "dl_ffn: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-2, "dl_ffn: cuTensorMapEncodeTiled(%s) failed with CUresult %d (cols %lld rows %lld ld %lld)",
                     name, (int)r, cols, rows, ld);
  return 0;
}

int check_args(const dl_ffn_args* a, const char* who) {
  DL_REQUIRE(a != nullptr, "%s: null args", who);
  DL_REQUIRE(a->x && a->w1 && a->w2 && a->y, "%s: x, w1, w2 and y must be non-null", who);
  DL_REQUIRE(a->D == kD, "%s: model width must be %d (got %lld)", who, kD, (long long)a->D);
  DL_REQUIRE(a->Dh >= kCh && a->Dh % kCh == 0 && a->Dh <= (1 << 20), "%s: hidden width %lld must be a multiple of %d", who, (long long)a->Dh, kCh);
  DL_REQUIRE(a->M >= 0 && a->M < (1ll << 31) - 128, "%s: bad row count %lld", who, (long long)a->M);
  DL_REQUIRE(a->ldy >= kD && a->ldy % 16 == 0 && ((uintptr_t)a->y & 31) == 0, "%s: y needs 32-byte aligned rows", who);
  return 0;
}

// CTA pairs are opt-in (DL_FFN_PAIR=1; they need at least two row tiles).  Measured on B200 at 16384 x 256 ->
// 1024 -> 256 (tools/ffn_probe.py): forward 45.1 us single-CTA / 47.1 us as pairs, backward 31.7 / 32.8 --
// halving the weight traffic per CTA buys nothing because the kernel is bound by the serial chain of its
// epilogue warps (tensor-memory load -> bias / GELU -> shared-memory store -> barrier, eight chunks per tile),
// not by the weight feed; the pair variant stays for the shapes where that changes (wider hidden layers).
bool use_pairs(long long M) {
  static const bool on = [] { const char* e = getenv("DL_FFN_PAIR"); return e && atoi(e) != 0; }();
  return on && M > 128;
}

template <int MODE, bool PAIR>
int launch(const CUtensorMap& ta, const CUtensorMap& tb1, const CUtensorMap& tb2, const CUtensorMap& te, FfnParams& p,
           cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(ffn_chain_kernel<MODE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  });
  if (attr_err != cudaSuccess)
    return set_error((int)attr_err, "dl_ffn: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  p.idesc1 = ptx::make_idesc(false, false, MODE == 1, PAIR ? 256 : 128, kCh);
  p.idesc2 = ptx::make_idesc(false, false, MODE == 1, PAIR ? 256 : 128, kD);
  if constexpr (PAIR) {
    const int pairs = (p.tiles + 1) / 2, max_pairs = sm_count() / 2;       // one CTA per SM: sm_count / 2 co-resident pairs
    const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
    (void)launch_cluster_k(ffn_chain_kernel<MODE, PAIR>, dim3(grid), dim3(kThreads), (size_t)kSmem, 2u, stream, ta, tb1, tb2, te, p);
  } else {
    const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
    DL_LAUNCH((ffn_chain_kernel<MODE, PAIR>), grid, kThreads, kSmem, stream, ta, tb1, tb2, te, p);
  }
  DL_LAUNCH_CHECK("ffn_chain_kernel");
  count_launch();
  return 0;
}

}  // namespace
}  // namespace dl

extern "C" int dl_ffn_fwd(const dl_ffn_args* a, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_args(a, "dl_ffn_fwd");
  if (rc) return rc;
  DL_REQUIRE(a->b1 && a->b2, "dl_ffn_fwd: b1 and b2 must be non-null");
  DL_REQUIRE(((uintptr_t)a->b1 & 15) == 0 && ((uintptr_t)a->b2 & 15) == 0, "dl_ffn_fwd: biases must be 16-byte aligned");
  DL_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "dl_ffn_fwd: drop_p must be in [0, 1)");
  DL_REQUIRE(!a->dact || a->hidden, "dl_ffn_fwd: dact needs hidden (both are the backward's inputs)");
  if (a->dact) DL_REQUIRE(a->ldh >= a->Dh && a->ldh % 16 == 0 && ((uintptr_t)a->dact & 31) == 0, "dl_ffn_fwd: dact needs 32-byte aligned rows");
  if (a->residual) DL_REQUIRE(a->ldr >= kD && a->ldr % 16 == 0 && ((uintptr_t)a->residual & 31) == 0, "dl_ffn_fwd: residual needs 32-byte aligned rows");
  if (a->M == 0) return 0;
  CUtensorMap ta, tb1, tb2, te;
  if ((rc = make_map(&ta, a->x, kD, a->M, a->ldx, 64, 128, "x"))) return rc;
  const bool pair = use_pairs(a->M);
  // [Dh, 256]: rows = hidden (pair: this CTA's 64 of a chunk's 128);  [256, Dh]: output rows (pair: 128 of the 256)
  if ((rc = make_map(&tb1, a->w1, kD, a->Dh, kD, 64, pair ? 64 : 128, "w1"))) return rc;
  if ((rc = make_map(&tb2, a->w2, a->Dh, kD, a->Dh, 64, pair ? 128 : 256, "w2"))) return rc;
  if (a->hidden) {
    if ((rc = make_map(&te, a->hidden, a->Dh, a->M, a->ldh, 64, 128, "hidden"))) return rc;
  } else {
    te = ta;
  }
  FfnParams p = {};
  p.bias1 = a->b1; p.bias2 = a->b2;
  p.dact = (__nv_bfloat16*)a->dact;
  p.res = (const __nv_bfloat16*)a->residual;
  p.Y = (__nv_bfloat16*)a->y;
  p.ldh = a->ldh; p.ldy = a->ldy; p.ldr = a->ldr;
  p.seed1 = a->seed1; p.seed2 = a->seed2; p.drop_step = (const long long*)a->drop_seed_step; p.drop_p = a->drop_p;
  p.M = (int)a->M; p.Dh = (int)a->Dh; p.NC = (int)(a->Dh / kCh); p.tiles = ceil_div(a->M, 128);
  p.store_e = a->hidden != nullptr;
  if (pair) return launch<0, true>(ta, tb1, tb2, te, p, stream);
  return launch<0, false>(ta, tb1, tb2, te, p, stream);
}

extern "C" int dl_ffn_bwd(const dl_ffn_args* a, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_args(a, "dl_ffn_bwd");
  if (rc) return rc;
  DL_REQUIRE(a->hidden && a->dact, "dl_ffn_bwd: hidden (out) and dact (in) must be non-null");
  DL_REQUIRE(a->ldh >= a->Dh && a->ldh % 16 == 0 && ((uintptr_t)a->dact & 31) == 0, "dl_ffn_bwd: dact needs 32-byte aligned rows");
  if (a->M == 0) return 0;
  CUtensorMap ta, tb1, tb2, te;
  if ((rc = make_map(&ta, a->x, kD, a->M, a->ldx, 64, 128, "g"))) return rc;
  if ((rc = make_map(&tb1, a->w2, a->Dh, kD, a->Dh, 64, 64, "w2"))) return rc;     // [256 (k), Dh]: MN-major
  if ((rc = make_map(&tb2, a->w1, kD, a->Dh, kD, 64, 64, "w1"))) return rc;        // [Dh (k), 256]: MN-major
  if ((rc = make_map(&te, a->hidden, a->Dh, a->M, a->ldh, 64, 128, "hidden"))) return rc;
  FfnParams p = {};
  p.dact = (__nv_bfloat16*)a->dact;
  p.Y = (__nv_bfloat16*)a->y;
  p.ldh = a->ldh; p.ldy = a->ldy;
  p.M = (int)a->M; p.Dh = (int)a->Dh; p.NC = (int)(a->Dh / kCh); p.tiles = ceil_div(a->M, 128);
  p.store_e = 1;
  if (use_pairs(a->M)) return launch<1, true>(ta, tb1, tb2, te, p, stream);
  return launch<1, false>(ta, tb1, tb2, te, p, stream);
}
