// dl_gemm: persistent, TMA-fed tcgen05 GEMM with double-buffered fp32 accumulators in tensor
// memory and a fused epilogue.
//
// One CTA per SM loops over 128 x BN output tiles (optionally K-slices of tiles: split-K).
// Warp roles (576 threads):
//   warp 0      TMA producer: fills a ring of kStages {A tile, B tile} buffers (SWIZZLE_128B),
//               running ahead across tiles
//   warp 1      allocates 2*BN TMEM columns, issues tcgen05.mma (one thread) into accumulator
//               buffer (tile & 1), commits to mbarriers
//   warps 2..17 epilogue (four warps per TMEM lane quadrant, a quarter of the columns each): tcgen05.ld
//               16 columns of the thread's row at a time, apply bias / activation / auxiliary
//               multiply / dropout / residual in registers and store whole 32-byte sectors with
//               16-byte vector stores (or atomically accumulate, split-K).  The epilogue of tile i
//               overlaps the main loop of tile i+1.
// A K-block is 128 bytes of K (64 bf16 or 32 tf32 values) = four tcgen05.mma instructions.
// Both operands may be K-major or MN-major (transposed storage); the shared-memory tile is
// always "rows x 128 B" so only the descriptors and TMA boxes differ (see ptx.cuh).
// Up to three batch dimensions (e.g. head, query-set, pair) map onto a 5-D tensor map.
#include <mutex>
#include <stdlib.h>

#include "../../include/druglamp_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace dl {
void count_launch(int n = 1);

namespace {

constexpr int BM = 128;
constexpr int kEpiWarps = 16;                 // four warps per TMEM lane quadrant
constexpr int kColSplit = kEpiWarps / 4;      // column slices per tile
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
// Weight-gradient GEMMs dW = dY^T X stream dY through shared memory as the MN-major A operand;
// while the tensor core consumes a stage, the first four (otherwise idle) epilogue warps also sum
// its columns: db[m] += sum_k A[k, m], the bias gradient, without a second pass over dY.
constexpr int kColsumWarps = 4;

// n / d for n < 2^31 with a multiply-high and a shift (every role decodes a tile index per tile;
// hardware integer division costs ~100 dependent cycles each)
struct FastDiv {
  uint32_t mul, shr, d;
  void init(uint32_t dv) {
    d = dv;
    shr = 0;
    while ((1u << shr) < dv) ++shr;
    mul = (uint32_t)(((1ull << 32) * ((1ull << shr) - dv)) / dv + 1);
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(n, mul) + n) >> shr; }
};

struct GemmParams {
  FastDiv fd_perz, fd_nt, fd_splits, fd_nb0, fd_nb1, fd_conv, fd_kred;
  int nkb_all;                  // K-blocks of the whole reduction
  float* colsum_a;              // bias gradient fused into a weight-gradient GEMM (see kColsumWarps)
  void* C;
  const float* bias;
  void* preact;
  const void* aux;
  const void* res;
  long long ldc, sc[3];
  long long ldr, sr[3];
  unsigned long long drop_seed;
  const long long* drop_step;   // device step counter mixed into the seed (or NULL)
  float drop_p;
  int M, N, K;
  int nb0, nb1;                 // extents of the two fastest batch dims
  int splits;                   // split-K factor (C is accumulated atomically when > 1)
  int mt, nt;                   // tiles along M and N
  int total_tiles;              // mt * nt * batch * splits
  int a_on[3], b_on[3];         // 0 when that batch stride is a broadcast
  int a_mn, b_mn;
  int c_bf16;
  int act, mul_mode;
  int in_on[3];                 // TMA epilogue: 0 when the elementwise input broadcasts over that batch dim
  int tma_in;                   // TMA epilogue: 0 none, 1 mul_aux, 2 residual arrives through the slab
  int pre_mode;                 // 0: preact = pre-activation; 1: preact = d(dropout(act(v)))/dv (backward's multiplier)
  int dbg;                      // bring-up only (DL_GEMM_DEBUG env): 1 no C stores, 2 no epilogue work,
                                // 3 plain stores without the lane transpose (A/B measurements)
  int conv_cin, conv_left;      // implicit-GEMM conv1d on A (0 = off): channels per tap, left padding
  int kred_kpb, kred_shift;     // K-reduction over batch[2] (0 = off): K-blocks per batch, B row shift
  int kred_tap;                 // 1: batch dim 0 is a convolution tap that also shifts B's rows
  float alpha;
  uint32_t idesc;
};

// SPLIT (fp32 operands only): "3xTF32".  kind::tf32 reads just the top 19 bits of each fp32
// operand, which costs ~1e-3 relative accuracy.  In SPLIT mode the epilogue warps rewrite every
// landed stage in place as hi = x & 0xffffe000 and lo = x - hi (exact) before the issuer runs
// three MMAs per K step (hi*hi + lo*hi + hi*lo): fp32-grade products on the tensor cores.  The
// conversion is elementwise, so it is independent of the swizzled layout.
// TMAEPI (bf16 output, BN = 256): the epilogue leaves through shared memory.  Every epilogue warp owns a
// 4 KB slab [32 rows x 128 B, SWIZZLE_128B]; an elementwise input (mul_aux or residual) arrives in it by
// TMA while the tile's MMAs still run, the warp reads its rows from it, writes the finished bf16 rows
// back into it and one lane issues a bulk tensor store.  Thread = row addressing of global memory makes
// every load / store instruction touch 32 different 128-byte lines (the L1 retires about one per
// cycle: 2048 cycles per 128 x 256 tile and operand, as long as the tile's MMAs at K = 256); through
// shared memory the same bytes cost a few hundred cycles and no LSU address traffic at all.
// CTA2 (bf16 operands, BN >= 128): a pair of CTAs (a cluster of two, one TPC) works on two vertically
// adjacent 128-row tiles as ONE 256 x BN tcgen05.mma.cta_group::2 tile.  Each CTA loads its own 128 rows of A
// but only HALF of the B tile (BN/2 columns); the leader CTA's thread issues every MMA for both, reading the
// operands out of both shared memories and writing both tensor memories.  The L2 -> SM operand traffic per
// output tile drops by a third (256-wide tiles: 48 -> 32 KB per K-block and CTA) and the smaller stages make
// the ring deeper -- the K <= 512 forward shapes and the split-K weight gradients are bound by exactly that feed.
template <int BN, bool TF32, bool SPLIT = false, bool TMAEPI = false, bool CTA2 = false>
struct Cfg {
  static_assert(!TMAEPI || (BN == 256 && !TF32), "TMA epilogue: bf16 operands, 256-wide tiles");
  static_assert(!CTA2 || (!TF32 && !SPLIT && BN >= 128), "CTA pairs: bf16 operands, tiles at least 128 wide");
  static constexpr int ELEM = TF32 ? 4 : 2;
  static constexpr int KE = 128 / ELEM;   // K elements per K-block
  static constexpr int UK = 32 / ELEM;    // K elements per tcgen05.mma
  static constexpr int MNB = 128 / ELEM;  // MN elements per 128-byte block (MN-major operands)
  static constexpr int BNL = CTA2 ? BN / 2 : BN;                   // B columns this CTA holds
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BNL * 128;
  static constexpr int LOAD_BYTES = A_BYTES + B_BYTES;             // what TMA delivers per stage (and CTA)
  static constexpr int STAGE_BYTES = SPLIT ? 2 * LOAD_BYTES : LOAD_BYTES;
  static constexpr int STAGES = CTA2 ? (TMAEPI ? 4 : (BN == 256 ? 6 : 8))
                                : TMAEPI ? 3 : SPLIT ? ((BN == 256) ? 2 : (BN == 128 ? 3 : 4))
                                               : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int EPI_BYTES = TMAEPI ? kEpiWarps * 4096 : 0;   // per-warp staging slabs
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert((3 * STAGES + 4 + kEpiWarps) * 8 + 8 <= BAR_BYTES, "barrier area");
};

struct Tile {
  int m0, n0, b0, b1, b2, kb_begin, nkb;
  unsigned zb;
};

// mrank / mstep: CTA pairs decode a tile PAIR index; CTA `mrank` of the pair owns its mrank-th 128 rows
template <int BN, int KE>
__device__ __forceinline__ Tile decode_tile(const GemmParams& p, int t, int mrank = 0, int mstep = 1) {
  Tile T;
  const uint32_t z = p.fd_perz.div((uint32_t)t), r = (uint32_t)t - z * p.fd_perz.d;
  const uint32_t mi = p.fd_nt.div(r), ni = r - mi * p.fd_nt.d;
  T.m0 = ((int)mi * mstep + mrank) * BM;
  T.n0 = (int)ni * BN;
  T.zb = p.fd_splits.div(z);
  const uint32_t ks = z - T.zb * p.fd_splits.d;
  const uint32_t q0 = p.fd_nb0.div(T.zb);
  T.b0 = (int)(T.zb - q0 * p.fd_nb0.d);
  T.b2 = (int)p.fd_nb1.div(q0);
  T.b1 = (int)(q0 - (uint32_t)T.b2 * p.fd_nb1.d);
  if (p.splits == 1) {
    T.kb_begin = 0;
    T.nkb = p.nkb_all;
  } else {                       // host guarantees nkb_all * splits < 2^31
    T.kb_begin = (int)p.fd_splits.div((uint32_t)p.nkb_all * ks);
    T.nkb = (int)p.fd_splits.div((uint32_t)p.nkb_all * (ks + 1)) - T.kb_begin;
  }
  return T;
}

// ---- element-vector helpers: 16 consecutive elements of one row -------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static __device__ __forceinline__ void load(const float* p, float (&x)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 f = reinterpret_cast<const float4*>(p)[i];
      x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
    }
  }
  static __device__ __forceinline__ void store(float* p, const float (&x)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
  }
};
// 16 bf16 = one 32-byte sector.  `wide` (32-byte aligned address): a single 256-bit access
// (LDG/STG.256, sm_100+) instead of two 128-bit ones -- thread = row, so every access is its own
// sector and the L1 processes one sector access per cycle: half the accesses, half the time.
template <> struct Vec16<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&x)[16], bool wide) {
    uint32_t w[8];
    if (wide) {
      asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                   : "l"(p));
    } else {
      const uint4 u0 = reinterpret_cast<const uint4*>(p)[0], u1 = reinterpret_cast<const uint4*>(p)[1];
      w[0] = u0.x; w[1] = u0.y; w[2] = u0.z; w[3] = u0.w;
      w[4] = u1.x; w[5] = u1.y; w[6] = u1.z; w[7] = u1.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {          // bf16 -> fp32 is a 16-bit shift
      x[2 * k] = __uint_as_float(w[k] << 16);
      x[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&x)[16], bool wide) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
      w[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    if (wide) {
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
    } else {
      reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
      reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
};

// vec: 16-byte vector access is legal; wide: 32-byte access is legal (bf16 only)
template <typename TC>
__device__ __forceinline__ void load16(const TC* p, float (&x)[16], bool vec, int nvalid, bool wide = false) {
  if (vec) {
    if constexpr (sizeof(TC) == 2) Vec16<TC>::load(p, x, wide);
    else Vec16<TC>::load(p, x);
    return;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = j < nvalid ? Cvt<TC>::to_f(p[j]) : 0.f;
}
template <typename TC>
__device__ __forceinline__ void store16(TC* p, const float (&x)[16], bool vec, int nvalid, bool wide = false) {
  if (vec) {
    if constexpr (sizeof(TC) == 2) Vec16<TC>::store(p, x, wide);
    else Vec16<TC>::store(p, x);
    return;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < nvalid) p[j] = Cvt<TC>::from_f(x[j]);
}

template <typename TC>
__device__ __forceinline__ float apply_mul(float x, float a, int mode) {
  return x * (mode == DL_MUL_GELU_GRAD ? gelu_grad<TC>(a)
              : mode == DL_MUL_RELU_MASK ? (a > 0.f ? 1.f : 0.f) : a);
}

// Launch-invariant epilogue facts, evaluated once per epilogue warp (not per tile).
struct EpiFlags {
  bool vec_c, vec_r, vec_p, vec_b;   // 16-byte vector access legal: C / residual / preact+aux / bias
  bool wide_c, wide_r, wide_p;       // one 32-byte access per 16 bf16 legal
  bool atomic, simple;
  float drop_inv;
  uint32_t drop_thr;
  unsigned long long drop_seed;      // host seed advanced by the device step counter
};

template <typename TC>
__device__ __forceinline__ EpiFlags make_epi_flags(const GemmParams& p) {
  EpiFlags f;
  constexpr long long kEl = 16 / sizeof(TC);
  auto strides_ok = [](long long ld, const long long* s, long long el) {
    return (ld % el) == 0 && (s[0] % el) == 0 && (s[1] % el) == 0 && (s[2] % el) == 0;
  };
  auto addr_ok = [](const void* q, uintptr_t bytes) { return (reinterpret_cast<uintptr_t>(q) & (bytes - 1)) == 0; };
  f.vec_c = strides_ok(p.ldc, p.sc, kEl) && addr_ok(p.C, 16);
  f.vec_r = strides_ok(p.ldr, p.sr, kEl) && addr_ok(p.res, 16);
  f.vec_p = f.vec_c && addr_ok(p.preact, 16) && addr_ok(p.aux, 16);
  f.vec_b = addr_ok(p.bias, 16);
  const bool w = sizeof(TC) == 2;
  f.wide_c = w && strides_ok(p.ldc, p.sc, 16) && addr_ok(p.C, 32);
  f.wide_r = w && strides_ok(p.ldr, p.sr, 16) && addr_ok(p.res, 32);
  f.wide_p = f.wide_c && addr_ok(p.preact, 32) && addr_ok(p.aux, 32);
  f.atomic = p.splits > 1;
  f.simple = !p.bias && !p.preact && !p.aux && !p.res && p.act == DL_ACT_NONE &&
             p.mul_mode == DL_MUL_NONE && p.drop_p == 0.f && !f.atomic && (p.dbg == 0 || p.dbg == 3);
  f.drop_thr = drop_threshold(p.drop_p);
  f.drop_seed = drop_seed_at(p.drop_seed, p.drop_p > 0.f ? p.drop_step : nullptr);
  f.drop_inv = p.drop_p > 0.f ? drop_scale(f.drop_thr) : 1.f;
  return f;
}

// 16 accumulator columns of one row -> global memory, with the fused element-wise work.
template <typename TC, bool SIMPLE, bool DERIV = false>
__device__ __forceinline__ void finish16(const GemmParams& p, const EpiFlags& f, const uint32_t* v,
                                         int row, int col, long long crow, long long rrow,
                                         unsigned zidx) {
  TC* Cp = reinterpret_cast<TC*>(p.C);
  const int nvalid = min(16, p.N - col);
  const bool full = nvalid == 16;
  const long long off = crow + col;
  float x[16];
  if constexpr (SIMPLE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]) * p.alpha;
    store16<TC>(Cp + off, x, full && f.vec_c, nvalid, f.wide_c);
    return;
  } else {
    TC* Pre = reinterpret_cast<TC*>(p.preact);
    const TC* Aux = reinterpret_cast<const TC*>(p.aux);
    const TC* Res = reinterpret_cast<const TC*>(p.res);
    const float2 al2 = f2(p.alpha);
    if (p.bias) {
      float b[16];
      load16<float>(p.bias + col, b, full && f.vec_b, nvalid);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 r = f2_fma(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), al2, make_float2(b[j], b[j + 1]));
        x[j] = r.x; x[j + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 r = f2_mul(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), al2);
        x[j] = r.x; x[j + 1] = r.y;
      }
    }
    constexpr bool deriv = DERIV;                    // preact_out = d dropout(act(v)) / dv (pre_mode 1): its own
    float dact[DERIV ? 16 : 1];                      // instantiation, so the common epilogues keep their registers
    if (!deriv && Pre) store16<TC>(Pre + off, x, full && f.vec_p, nvalid, f.wide_p);
    if (p.act == DL_ACT_GELU) {
      if constexpr (deriv) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float2 y2, d2;
          gelu_fwd_grad2<TC>(make_float2(x[j], x[j + 1]), y2, d2);
          x[j] = y2.x; x[j + 1] = y2.y; dact[j] = d2.x; dact[j + 1] = d2.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float2 y2 = gelu_fwd2<TC>(make_float2(x[j], x[j + 1]));
          x[j] = y2.x; x[j + 1] = y2.y;
        }
      }
    } else if (p.act == DL_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if constexpr (deriv) dact[j] = x[j] > 0.f ? 1.f : 0.f;
        x[j] = fmaxf(x[j], 0.f);
      }
    } else if constexpr (deriv) {
#pragma unroll
      for (int j = 0; j < 16; ++j) dact[j] = 1.f;
    }
    if (p.mul_mode != DL_MUL_NONE) {
      float m[16];
      load16<TC>(Aux + off, m, full && f.vec_p, nvalid, f.wide_p);
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = apply_mul<TC>(x[j], m[j], p.mul_mode);
    }
    if (p.drop_p > 0.f) {
      const unsigned long long e = ((unsigned long long)zidx * p.M + row) * p.N + col;
      if ((e & 1ull) == 0) {                     // aligned pairs: eight hashes for sixteen elements
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t h = drop_hash(f.drop_seed, (e >> 1) + k);
          const float2 m = make_float2((h & 0xffffu) >= f.drop_thr ? f.drop_inv : 0.f,
                                       (h >> 16) >= f.drop_thr ? f.drop_inv : 0.f);
          const float2 xm = f2_mul(make_float2(x[2 * k], x[2 * k + 1]), m);
          x[2 * k] = xm.x; x[2 * k + 1] = xm.y;
          if constexpr (deriv) {
            const float2 dm = f2_mul(make_float2(dact[2 * k], dact[2 * k + 1]), m);
            dact[2 * k] = dm.x; dact[2 * k + 1] = dm.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float m = drop_keep(f.drop_seed, e + j, f.drop_thr) ? f.drop_inv : 0.f;
          x[j] *= m;
          if constexpr (deriv) dact[j] *= m;
        }
      }
    }
    if constexpr (deriv) store16<TC>(Pre + off, dact, full && f.vec_p, nvalid, f.wide_p);
    if (Res) {
      float r[16];
      load16<TC>(Res + rrow + col, r, full && f.vec_r, nvalid, f.wide_r);
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] += r[j];
    }
    if (p.dbg == 1 && x[0] != 12345.678f) return;
    if (f.atomic) {
      float* dst = reinterpret_cast<float*>(p.C) + off;
      if (full && f.vec_c) {                           // 4 x 16-byte vector reductions (REDG.F32x4)
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                       :: "l"(dst + j), "f"(x[j]), "f"(x[j + 1]), "f"(x[j + 2]), "f"(x[j + 3]) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nvalid) atomicAdd(dst + j, x[j]);
      }
    } else {
      store16<TC>(Cp + off, x, full && f.vec_c, nvalid, f.wide_c);
    }
  }
}

// One warp's 32-row x (BN/4)-column slice of a tile, straight from tensor memory to global memory:
// thread = row (the tcgen05.ld 32x32b layout).  Every 32-byte sector a thread touches is written
// (or read) in full.  Plain stores (SIMPLE) pull 32 columns per tensor-memory round trip.
template <typename TC, int BN, bool SIMPLE, bool DERIV = false>
__device__ __forceinline__ void epilogue_slice(const GemmParams& p, const EpiFlags& f, uint32_t tmem_q,
                                               int lane, int row0, int n0, int col_begin,
                                               long long cbase, long long rbase, unsigned zidx) {
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const long long crow = cbase + (long long)row * p.ldc;
  const long long rrow = rbase + (long long)row * p.ldr;
  constexpr int W = BN / kColSplit;
  constexpr int STEP = (SIMPLE && W >= 32) ? 32 : 16;
  if constexpr (SIMPLE && sizeof(TC) == 2 && W == 64) {
    // Plain bf16 stores of a full 64-column slice: thread = row would make every store
    // instruction touch 32 different 128-byte lines (32 bytes of each).  A 4x4 register transpose
    // inside each group of four lanes (two shuffle stages) gives lane 4g+j chunk j of rows
    // 4g..4g+3, so one instruction writes 8 rows x 128 contiguous bytes: a quarter of the
    // line accesses the L1 has to process.
    if (f.wide_c && n0 + col_begin + W <= p.N && p.dbg != 3) {       // warp-uniform
      uint32_t P[4][8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_q + (uint32_t)(col_begin + 32 * h), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[16 * c + 2 * k]) * p.alpha,
                                                            __uint_as_float(v[16 * c + 2 * k + 1]) * p.alpha);
            P[2 * h + c][k] = *reinterpret_cast<const uint32_t*>(&b2);
          }
        }
      }
      const bool b0 = lane & 1, b1 = lane & 2;
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t y = __shfl_xor_sync(0xffffffffu, b0 ? P[c][k] : P[c + 1][k], 1);
          if (b0) P[c][k] = y; else P[c + 1][k] = y;
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t y = __shfl_xor_sync(0xffffffffu, b1 ? P[c][k] : P[c + 2][k], 2);
          if (b1) P[c][k] = y; else P[c + 2][k] = y;
        }
      }
      // P[i] = chunk (lane & 3) of row (lane & ~3) + i
      TC* Cp = reinterpret_cast<TC*>(p.C);
      const int rg = row0 + (lane & ~3);
      const long long cofs = cbase + n0 + col_begin + 16 * (lane & 3);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (rg + i < p.M) {
          TC* dst = Cp + cofs + (long long)(rg + i) * p.ldc;
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                       :: "l"(dst), "r"(P[i][0]), "r"(P[i][1]), "r"(P[i][2]), "r"(P[i][3]),
                          "r"(P[i][4]), "r"(P[i][5]), "r"(P[i][6]), "r"(P[i][7]) : "memory");
        }
      }
      return;
    }
  }
#pragma unroll 1
  for (int c0 = col_begin; c0 < col_begin + W; c0 += STEP) {
    const int col = n0 + c0;
    if (col >= p.N) break;                           // warp-uniform
    uint32_t v[STEP];
    if constexpr (STEP == 32) ptx::tmem_ld_32x32(tmem_q + (uint32_t)c0, v);   // warp-collective:
    else ptx::tmem_ld_32x16(tmem_q + (uint32_t)c0, v);                        // before any divergence
    ptx::tmem_ld_wait();
    if (!row_ok) continue;
    finish16<TC, SIMPLE, DERIV>(p, f, v, row, col, crow, rrow, zidx);
    if constexpr (STEP == 32) {
      if (col + 16 < p.N) finish16<TC, SIMPLE, DERIV>(p, f, v + 16, row, col + 16, crow, rrow, zidx);
    }
  }
}

// TMA epilogue: 16 accumulator columns of one row -> this thread's row of the warp's shared-memory
// slab (two 16-byte chunks, SWIZZLE_128B), with the same fused element-wise work as finish16.  An
// elementwise input (p.tma_in: 1 = mul_aux, 2 = residual) is read from the very chunks it overwrites.
template <bool DERIV>
__device__ __forceinline__ void tma_chunk16(const GemmParams& p, const EpiFlags& f, const uint32_t* v, int row,
                                            bool row_ok, int col, long long crow, unsigned zidx,
                                            uint32_t slab_row, int chunk0, int xr) {
  using TC = __nv_bfloat16;
  const int nvalid = min(16, p.N - col);
  if (nvalid <= 0) return;                           // warp-uniform: the store clips these columns
  const bool full = nvalid == 16;
  TC* Pre = reinterpret_cast<TC*>(p.preact);
  float x[16];
  const float2 al2 = f2(p.alpha);
  if (p.bias) {
    float b[16];
    load16<float>(p.bias + col, b, full && f.vec_b, nvalid);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float2 r = f2_fma(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), al2, make_float2(b[j], b[j + 1]));
      x[j] = r.x; x[j + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float2 r = f2_mul(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), al2);
      x[j] = r.x; x[j + 1] = r.y;
    }
  }
  float dact[DERIV ? 16 : 1];
  if (!DERIV && Pre && row_ok) store16<TC>(Pre + crow + col, x, full && f.vec_p, nvalid, f.wide_p);
  if (p.act == DL_ACT_GELU) {
    if constexpr (DERIV) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        float2 y2, d2;
        gelu_fwd_grad2<TC>(make_float2(x[j], x[j + 1]), y2, d2);
        x[j] = y2.x; x[j + 1] = y2.y; dact[j] = d2.x; dact[j + 1] = d2.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 y2 = gelu_fwd2<TC>(make_float2(x[j], x[j + 1]));
        x[j] = y2.x; x[j + 1] = y2.y;
      }
    }
  } else if (p.act == DL_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if constexpr (DERIV) dact[j] = x[j] > 0.f ? 1.f : 0.f;
      x[j] = fmaxf(x[j], 0.f);
    }
  } else if constexpr (DERIV) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dact[j] = 1.f;
  }
  const uint32_t a0 = slab_row + (uint32_t)((chunk0 ^ xr) << 4), a1 = slab_row + (uint32_t)(((chunk0 + 1) ^ xr) << 4);
  float in[16];
  if (p.tma_in != 0) {
    uint32_t w0[4], w1[4];
    ptx::lds128(a0, w0);
    ptx::lds128(a1, w1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      in[2 * k] = __uint_as_float(w0[k] << 16);
      in[2 * k + 1] = __uint_as_float(w0[k] & 0xffff0000u);
      in[8 + 2 * k] = __uint_as_float(w1[k] << 16);
      in[8 + 2 * k + 1] = __uint_as_float(w1[k] & 0xffff0000u);
    }
  }
  if (p.tma_in == 1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = apply_mul<TC>(x[j], in[j], p.mul_mode);
  }
  if (p.drop_p > 0.f) {
    const unsigned long long e = ((unsigned long long)zidx * p.M + row) * p.N + col;
    if ((e & 1ull) == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t h = drop_hash(f.drop_seed, (e >> 1) + k);
        const float m0 = (h & 0xffffu) >= f.drop_thr ? f.drop_inv : 0.f;
        const float m1 = (h >> 16) >= f.drop_thr ? f.drop_inv : 0.f;
        x[2 * k] *= m0;
        x[2 * k + 1] *= m1;
        if constexpr (DERIV) { dact[2 * k] *= m0; dact[2 * k + 1] *= m1; }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float m = drop_keep(f.drop_seed, e + j, f.drop_thr) ? f.drop_inv : 0.f;
        x[j] *= m;
        if constexpr (DERIV) dact[j] *= m;
      }
    }
  }
  if constexpr (DERIV) {
    if (row_ok) store16<TC>(Pre + crow + col, dact, full && f.vec_p, nvalid, f.wide_p);
  }
  if (p.tma_in == 2) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] += in[j];
  }
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = ptx::pack_bf16(x[2 * k], x[2 * k + 1]);
  ptx::sts128(a0, w[0], w[1], w[2], w[3]);
  ptx::sts128(a1, w[4], w[5], w[6], w[7]);
}

template <int BN, bool TF32, bool SPLIT, bool TMAEPI, bool CTA2>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmIn,
               const GemmParams p) {
  using C = Cfg<BN, TF32, SPLIT, TMAEPI, CTA2>;
  static_assert(!SPLIT || TF32, "SPLIT applies to fp32 operands");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_addr - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  const uint32_t bar_full = ptx::smem_u32(bars);                    // [STAGES] TMA landed
  const uint32_t bar_empty = bar_full + 8 * C::STAGES;              // [STAGES] MMAs consumed
  const uint32_t bar_conv = bar_empty + 8 * C::STAGES;              // [STAGES] hi/lo split done
  const uint32_t bar_tfull = bar_conv + 8 * C::STAGES;              // [2] accumulator ready
  const uint32_t bar_tempty = bar_tfull + 16;                       // [2] accumulator drained
  const uint32_t bar_in = bar_tempty + 16;                          // [kEpiWarps] TMA epilogue: input slab landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::STAGES + 4 + kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA pairs: both CTAs walk the same sequence of tile pairs; `rank` picks the 128 rows (and the half of B)
  const int rank = CTA2 ? (int)ptx::cluster_ctarank() : 0;
  const int tile_first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kMStep = CTA2 ? 2 : 1;
  pdl_trigger();

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      // pair: the leader's barrier takes its own expect_tx arrival and the peer producer's remote arrival
      ptx::mbar_init(bar_full + 8 * s, CTA2 ? 2 : 1);
      ptx::mbar_init(bar_empty + 8 * s, p.colsum_a ? 1 + kColsumWarps : 1);
      // pair: "stage landed" forwarded by the leader's issuer to the peer's column-sum warps (colsum_a)
      ptx::mbar_init(bar_conv + 8 * s, CTA2 ? 1 : 32 * kEpiWarps);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_tfull + 8 * i, 1);
      // pair: the leader issues for both, so it waits for the epilogue warps of both CTAs
      ptx::mbar_init(bar_tempty + 8 * i, CTA2 ? 2 * kEpiWarps : kEpiWarps);
    }
    for (int i = 0; i < kEpiWarps; ++i) ptx::mbar_init(bar_in + 8 * i, 1);
    if constexpr (TMAEPI) {
      ptx::prefetch_tmap(&tmC);
      if (p.tma_in) ptx::prefetch_tmap(&tmIn);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) ptx::tmem_alloc_2sm<C::TMEM_COLS>(ptx::smem_u32(tmem_slot));
    else ptx::tmem_alloc<C::TMEM_COLS>(ptx::smem_u32(tmem_slot));
  }
  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync_all();      // the peer's barriers exist before anything arrives on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // everything above overlaps the tail of the previous kernel; nothing below (TMA loads, epilogue
  // reads and writes) may start before that kernel's results are visible
  pdl_wait();

  // Producer and issuer run their loops WARP-UNIFORMLY (all 32 lanes wait on the barriers and keep the loop
  // state; one elected lane issues the TMA / tcgen05 instructions).  As single-lane divergent regions the
  // compiler kept every descriptor in per-thread registers and paid R2UR + ELECT for each tcgen05.mma: ~130
  // dependent instructions = ~790 cycles per K-block of four 128-cycle MMAs, which bounded every long-K GEMM.
  if (warp == 0) {
    {
      // ------------------------------------------------------------ TMA producer
      // pair: every load of either CTA is counted on the LEADER's barrier (shared::cluster address)
      auto load = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int x0, int x1, int x2, int x3, int x4) {
        if constexpr (CTA2) ptx::tma_load_5d_2sm(dst, m, bar, x0, x1, x2, x3, x4);
        else ptx::tma_load_5d(dst, m, bar, x0, x1, x2, x3, x4);
      };
      const int nb_off = rank * C::BNL;             // this CTA's columns of the B tile
      const bool leader_lane = ptx::elect_one();
      uint32_t s = 0, ph = 0;
      for (int t = tile_first; t < p.total_tiles; t += tile_step) {
        const Tile T = decode_tile<BN, C::KE>(p, t, rank, kMStep);
        const int a2 = p.a_on[0] ? T.b0 : 0, a3 = p.a_on[1] ? T.b1 : 0, a4 = p.a_on[2] ? T.b2 : 0;
        const int c2 = p.b_on[0] ? T.b0 : 0, c3 = p.b_on[1] ? T.b1 : 0, c4 = p.b_on[2] ? T.b2 : 0;
        for (int kb = 0; kb < T.nkb; ++kb) {
          ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          uint32_t full = bar_full + 8 * s;
          const uint32_t sA = base_addr + s * C::STAGE_BYTES, sB = sA + C::A_BYTES;
          if (++s == (uint32_t)C::STAGES) { s = 0; ph ^= 1u; }
          if (p.dbg == 5) {                     // bring-up: barrier protocol only, no loads (results are garbage)
            if (leader_lane) {
              if (!CTA2 || rank == 0) ptx::mbar_arrive_expect_tx(full, 0);
              else ptx::mbar_arrive_remote(ptx::mapa(full, 0));
            }
            continue;
          }
          const int k0 = (T.kb_begin + kb) * C::KE;
          int ka = k0, kbb = k0, a_row = T.m0, a4r = a4, c4r = c4;
          if (p.conv_cin) {
            // implicit-GEMM conv1d: K-block -> (tap, channel block); the A tile is the same rows
            // shifted by the tap, out-of-range rows are zero-filled by TMA ('same' padding)
            const int tap = (int)p.fd_conv.div((uint32_t)k0);
            ka = k0 - tap * p.conv_cin;
            a_row = T.m0 + tap - p.conv_left;
          }
          if (p.kred_kpb) {
            // K runs over (reduction batch, rows): conv weight gradient; batch dim 0 is the tap and
            // shifts the rows of B
            const int kbi = T.kb_begin + kb;
            const int rb = (int)p.fd_kred.div((uint32_t)kbi);
            ka = (kbi - rb * p.kred_kpb) * C::KE;
            kbb = ka + (p.kred_tap ? T.b0 : 0) + p.kred_shift;
            a4r = c4r = rb;
          }
          if (leader_lane) {
            if constexpr (CTA2) {
              const uint32_t lfull = ptx::mapa(full, 0);
              if (rank == 0) ptx::mbar_arrive_expect_tx(full, 2 * C::LOAD_BYTES);
              else ptx::mbar_arrive_remote(lfull);
              full = lfull;
            } else {
              ptx::mbar_arrive_expect_tx(full, C::LOAD_BYTES);
            }
            if (!p.a_mn) {
              load(sA, &tmA, full, ka, a_row, a2, a3, a4r);
            } else {
#pragma unroll
              for (int blk = 0; blk < BM / C::MNB; ++blk)
                load(sA + blk * C::KE * 128, &tmA, full, T.m0 + blk * C::MNB, ka, a2, a3, a4r);
            }
            if (!p.b_mn) {
              load(sB, &tmB, full, kbb, T.n0 + nb_off, c2, c3, c4r);
            } else {
#pragma unroll
              for (int blk = 0; blk < C::BNL / C::MNB; ++blk)
                load(sB + blk * C::KE * 128, &tmB, full, T.n0 + nb_off + blk * C::MNB, kbb, c2, c3, c4r);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ------------------------------------------------------------ MMA issuer (pair: the leader, for both)
      // Launch-invariant halves of the shared-memory descriptors: only the 14-bit start-address field
      // changes per stage and per K step.  MN-major: bf16 uses SWIZZLE_128B (8-row K atom), tf32 must use
      // 128B_BASE32B (4-row).
      constexpr uint32_t kMnSbo = TF32 ? 512 : 1024, kMnType = TF32 ? 1 : 2;
      const uint64_t a_tmpl = p.a_mn ? ptx::make_smem_desc(0, C::KE * 128, kMnSbo, kMnType) : ptx::make_smem_desc(0, 16, 1024);
      const uint64_t b_tmpl = p.b_mn ? ptx::make_smem_desc(0, C::KE * 128, kMnSbo, kMnType) : ptx::make_smem_desc(0, 16, 1024);
      const uint32_t a_kstep = (p.a_mn ? C::UK * 128 : 32) >> 4, b_kstep = (p.b_mn ? C::UK * 128 : 32) >> 4;   // 16-byte units
      const bool leader_lane = ptx::elect_one();
      uint32_t s = 0, ph = 0, ti = 0;
      for (int t = tile_first; t < p.total_tiles; t += tile_step, ++ti) {
        const Tile T = decode_tile<BN, C::KE>(p, t, rank, kMStep);
        const uint32_t ab = ti & 1, aph = (ti >> 1) & 1;
        ptx::mbar_wait(bar_tempty + 8 * ab, aph ^ 1u);     // epilogue has drained this buffer
        ptx::tc_fence_after();
        const uint32_t acc = tmem + ab * BN;
        for (int kb = 0; kb < T.nkb; ++kb) {
          ptx::mbar_wait((SPLIT ? bar_conv : bar_full) + 8 * s, ph);
          ptx::tc_fence_after();
          const uint32_t sA = base_addr + s * C::STAGE_BYTES, sB = sA + C::A_BYTES;
          const uint64_t ad0 = a_tmpl | (uint64_t)((sA & 0x3FFFF) >> 4), bd0 = b_tmpl | (uint64_t)((sB & 0x3FFFF) >> 4);
          if (leader_lane) {
            if constexpr (CTA2) {
              // only the leader's barrier sees the loads complete: tell the peer's column-sum warps.  (Plain
              // remote arrive: a .release.cluster one compiles to MEMBAR.ALL.GPU, which -- even predicated off for
              // launches without colsum_a -- cost ~250 cycles per K-block in this loop: 8192^3 1.23 -> 1.46 PFLOP/s.)
              if (p.colsum_a != nullptr) ptx::mbar_arrive_remote(ptx::mapa(bar_conv + 8 * s, 1));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = ad0 + (uint64_t)(k * a_kstep), bd = bd0 + (uint64_t)(k * b_kstep);
              if constexpr (CTA2) ptx::mma_ss_2sm(acc, ad, bd, p.idesc, (uint32_t)((kb | k) != 0));
              else ptx::mma_ss<TF32>(acc, ad, bd, p.idesc, (uint32_t)((kb | k) != 0));
              if constexpr (SPLIT) {
                // descriptors address 16-byte units: the lo copies sit LOAD_BYTES above the hi ones
                constexpr uint64_t kLo = (uint64_t)(C::LOAD_BYTES >> 4);
                ptx::mma_ss<TF32>(acc, ad + kLo, bd, p.idesc, 1u);
                ptx::mma_ss<TF32>(acc, ad, bd + kLo, p.idesc, 1u);
              }
            }
            // frees the stage once these MMAs have read it (pair: in both CTAs)
            if constexpr (CTA2) ptx::mma_commit_2sm(bar_empty + 8 * s);
            else ptx::mma_commit(bar_empty + 8 * s);
          }
          __syncwarp();
          if (++s == (uint32_t)C::STAGES) { s = 0; ph ^= 1u; }
        }
        if (leader_lane) {
          if constexpr (CTA2) ptx::mma_commit_2sm(bar_tfull + 8 * ab);
          else ptx::mma_commit(bar_tfull + 8 * ab);
        }
        __syncwarp();
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..9)
    const int we = warp - 2;
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = we >> 2;                     // which column slice of the tile
    const EpiFlags ef = p.c_bf16 ? make_epi_flags<__nv_bfloat16>(p) : make_epi_flags<float>(p);
    uint32_t it = 0, ti = 0, in_phase = 0;
    // the accumulator buffer is handed back to the issuing CTA (pair: the leader's barrier, from both CTAs)
    auto release_acc = [&](uint32_t ab_) {
      if constexpr (CTA2) ptx::mbar_arrive_remote(ptx::mapa(bar_tempty + 8 * ab_, 0));
      else ptx::mbar_arrive(bar_tempty + 8 * ab_);
    };
    for (int tl = tile_first; tl < p.total_tiles; tl += tile_step, ++ti) {
      const Tile T = decode_tile<BN, C::KE>(p, tl, rank, kMStep);
      if constexpr (SPLIT) {
        const int tid = threadIdx.x - 64;         // 0 .. 32*kEpiWarps-1
        for (int kb = 0; kb < T.nkb; ++kb, ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          uint4* hi = reinterpret_cast<uint4*>(smem + s * C::STAGE_BYTES);
          float4* lo = reinterpret_cast<float4*>(smem + s * C::STAGE_BYTES + C::LOAD_BYTES);
#pragma unroll 4
          for (int i = tid; i < C::LOAD_BYTES / 16; i += 32 * kEpiWarps) {
            const uint4 u = hi[i];
            const uint4 h = make_uint4(u.x & 0xffffe000u, u.y & 0xffffe000u, u.z & 0xffffe000u,
                                       u.w & 0xffffe000u);
            lo[i] = make_float4(__uint_as_float(u.x) - __uint_as_float(h.x),
                                __uint_as_float(u.y) - __uint_as_float(h.y),
                                __uint_as_float(u.z) - __uint_as_float(h.z),
                                __uint_as_float(u.w) - __uint_as_float(h.w));
            hi[i] = h;
          }
          ptx::fence_proxy_async();               // generic-proxy writes -> visible to the MMA
          ptx::mbar_arrive(bar_conv + 8 * s);
        }
      }
      if constexpr (!TF32) {
        if (p.colsum_a != nullptr && we < kColsumWarps) {
          // Column sums of the MN-major bf16 A tile (two blocks of [64 k-rows][64 mn = 128 B],
          // SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7)).  Thread = (block,
          // chunk, row class r & 7): its eight rows share the XOR, so its chunk never moves, and
          // the eight threads of a quarter-warp read eight different chunk positions (no bank
          // conflicts).  Only the first n-tile of every (m-tile, K-slice) adds to the result; the
          // warps still take part in the stage hand-shake for the other tiles.
          const int t = threadIdx.x - 64;
          const int blk = t >> 6, ch = (t >> 3) & 7, rcls = t & 7;
          const bool active = T.n0 == 0;
          const uint32_t bar_landed = (CTA2 && rank != 0) ? bar_conv : bar_full;
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          for (int kb = 0; kb < T.nkb; ++kb, ++it) {
            const int s = it % C::STAGES;
            const uint32_t ph = (it / C::STAGES) & 1;
            ptx::mbar_wait(bar_landed + 8 * s, ph);
            if (active) {
              const uint8_t* src = smem + s * C::STAGE_BYTES + blk * (C::KE * 128) + ((ch ^ rcls) << 4) + rcls * 128;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint4 u = *reinterpret_cast<const uint4*>(src + i * 1024);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  acc[2 * k] += __uint_as_float(w[k] << 16);
                  acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
                }
              }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_empty + 8 * s);
          }
          if (active) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
              acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
              acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
            }
            if (rcls == 0) {
              const int m = T.m0 + blk * 64 + ch * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (m + j < p.M) atomicAdd(p.colsum_a + m + j, acc[j]);
            }
          }
        }
      }
      const uint32_t ab = ti & 1, aph = (ti >> 1) & 1;
      if constexpr (TMAEPI) {
        // the warp's slab: wait until the previous tile's store has read it, then (optionally) start
        // the elementwise input's load -- it lands while this tile's MMAs still run
        const uint32_t slab = base_addr + C::STAGES * C::STAGE_BYTES + (uint32_t)we * 4096u;
        const int cb = half * (BN / kColSplit), row0 = T.m0 + q * 32;
        const bool live = T.n0 + cb < p.N;                               // warp-uniform
        if (lane == 0) ptx::bulk_wait_read0();
        __syncwarp();
        if (p.tma_in && live && lane == 0) {
          ptx::mbar_arrive_expect_tx(bar_in + 8 * we, 4096);
          ptx::tma_load_5d(slab, &tmIn, bar_in + 8 * we, T.n0 + cb, row0, p.in_on[0] ? T.b0 : 0,
                           p.in_on[1] ? T.b1 : 0, p.in_on[2] ? T.b2 : 0);
        }
        ptx::mbar_wait(bar_tfull + 8 * ab, aph);
        ptx::tc_fence_after();
        if (p.tma_in && live) ptx::mbar_wait(bar_in + 8 * we, in_phase), in_phase ^= 1u;
        const long long cbase = (long long)T.b0 * p.sc[0] + (long long)T.b1 * p.sc[1] + (long long)T.b2 * p.sc[2];
        const uint32_t tmem_q = tmem + ab * BN + ((uint32_t)(q * 32) << 16);
        const int row = row0 + lane;
        const bool row_ok = row < p.M;
        const long long crow = cbase + (long long)row * p.ldc;
        const uint32_t slab_row = slab + (uint32_t)lane * 128u;
        const bool deriv = p.pre_mode == 1 && p.preact != nullptr;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(tmem_q + (uint32_t)(cb + 16 * j), v);
          ptx::tmem_ld_wait();
          if (deriv) tma_chunk16<true>(p, ef, v, row, row_ok, T.n0 + cb + 16 * j, crow, T.zb, slab_row, 2 * j, lane & 7);
          else tma_chunk16<false>(p, ef, v, row, row_ok, T.n0 + cb + 16 * j, crow, T.zb, slab_row, 2 * j, lane & 7);
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (live) {
            ptx::tma_store_5d(&tmC, slab, T.n0 + cb, row0, T.b0, T.b1, T.b2);
            ptx::bulk_commit();
          }
          release_acc(ab);
        }
        continue;
      }
      ptx::mbar_wait(bar_tfull + 8 * ab, aph);
      ptx::tc_fence_after();
      const long long cbase = (long long)T.b0 * p.sc[0] + (long long)T.b1 * p.sc[1] + (long long)T.b2 * p.sc[2];
      const long long rbase = (long long)T.b0 * p.sr[0] + (long long)T.b1 * p.sr[1] + (long long)T.b2 * p.sr[2];
      const uint32_t tmem_q = tmem + ab * BN + ((uint32_t)(q * 32) << 16);
      const int row0 = T.m0 + q * 32, cb = half * (BN / kColSplit);
      if (p.dbg == 2) {
      } else if (p.c_bf16) {
        if (ef.simple) epilogue_slice<__nv_bfloat16, BN, true>(p, ef, tmem_q, lane, row0, T.n0, cb, cbase, rbase, T.zb);
        else if (p.pre_mode == 1 && p.preact) epilogue_slice<__nv_bfloat16, BN, false, true>(p, ef, tmem_q, lane, row0, T.n0, cb, cbase, rbase, T.zb);
        else epilogue_slice<__nv_bfloat16, BN, false>(p, ef, tmem_q, lane, row0, T.n0, cb, cbase, rbase, T.zb);
      } else {
        if (ef.simple) epilogue_slice<float, BN, true>(p, ef, tmem_q, lane, row0, T.n0, cb, cbase, rbase, T.zb);
        else if (p.pre_mode == 1 && p.preact) epilogue_slice<float, BN, false, true>(p, ef, tmem_q, lane, row0, T.n0, cb, cbase, rbase, T.zb);
        else epilogue_slice<float, BN, false>(p, ef, tmem_q, lane, row0, T.n0, cb, cbase, rbase, T.zb);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(ab);
    }
    if constexpr (TMAEPI) {
      if (lane == 0) ptx::bulk_wait0();          // the slabs must outlive the stores that read them
    }
  }
  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync_all();      // neither CTA's shared / tensor memory may go while the other uses it
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CTA2) ptx::tmem_dealloc_2sm<C::TMEM_COLS>(tmem);
    else ptx::tmem_dealloc<C::TMEM_COLS>(tmem);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// One operand: `mn` = logical MN extent, K = contraction extent, stored row-major with leading
// dimension ld, either [mn, K] (K-major) or [K, mn] (MN-major); nb/s = batch extents / strides.
int make_operand_map(CUtensorMap* m, const void* ptr, bool f32, bool mn_major, long long mn,
                     long long K, long long ld, const int64_t* nb, const int64_t* s, int tile_mn,
                     const char* name) {
  const int es = f32 ? 4 : 2;
  const int KE = 128 / es;
  DL_REQUIRE(((uintptr_t)ptr & 15) == 0, "dl_gemm: %s base pointer must be 16-byte aligned", name);
  DL_REQUIRE((ld * es) % 16 == 0, "dl_gemm: %s row stride (%lld elements) must be a multiple of 16 bytes", name, ld);
  const long long inner = mn_major ? mn : K, outer = mn_major ? K : mn;
  DL_REQUIRE(ld >= inner, "dl_gemm: %s leading dimension %lld < row length %lld", name, ld, inner);
  cuuint64_t dims[5] = {(cuuint64_t)inner, (cuuint64_t)outer, 1, 1, 1};
  cuuint64_t strides[4] = {(cuuint64_t)(ld * es), 0, 0, 0};
  const long long dummy = outer * ld;   // stride for broadcast (extent-1) dims: any legal value
  for (int i = 0; i < 3; ++i) {
    DL_REQUIRE(s[i] >= 0 && (s[i] * es) % 16 == 0, "dl_gemm: %s batch stride %d must be a non-negative multiple of 16 bytes", name, i);
    dims[2 + i] = (cuuint64_t)(s[i] ? nb[i] : 1);
    strides[1 + i] = (cuuint64_t)((s[i] ? s[i] : dummy) * es);
  }
  cuuint32_t box[5] = {(cuuint32_t)(mn_major ? 128 / es : KE), (cuuint32_t)(mn_major ? KE : tile_mn), 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  DL_REQUIRE(fn != nullptr, "dl_gemm: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  (f32 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-2, "dl_gemm: cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu,%llu,%llu ld %lld)",
                     name, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                     (unsigned long long)dims[2], (unsigned long long)dims[3],
                     (unsigned long long)dims[4], ld);
  return 0;
}

// bf16 [.., rows, cols] tensor as the epilogue's slabs see it: box = 64 columns x 32 rows, SWIZZLE_128B
int make_epi_map(CUtensorMap* m, const void* ptr, long long cols, long long rows, long long ld,
                 const int64_t* nb, const long long* s, const char* name) {
  cuuint64_t dims[5] = {(cuuint64_t)cols, (cuuint64_t)rows, 1, 1, 1};
  cuuint64_t strides[4] = {(cuuint64_t)(ld * 2), 0, 0, 0};
  const long long dummy = rows * ld;
  for (int i = 0; i < 3; ++i) {
    dims[2 + i] = (cuuint64_t)(s[i] ? nb[i] : 1);
    strides[1 + i] = (cuuint64_t)((s[i] ? s[i] : dummy) * 2);
  }
  cuuint32_t box[5] = {64, 32, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  DL_REQUIRE(fn != nullptr, "dl_gemm: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-2, "dl_gemm: cuTensorMapEncodeTiled(%s) failed with CUresult %d", name, (int)r);
  return 0;
}

// CTA pairs: the persistent loop assumes every launched pair is resident at once.  How many pairs the
// device co-schedules (GPC / TPC floor-sweeping can leave fewer than sm_count / 2) is asked once, for the
// largest pair configuration (every variant needs one whole SM per CTA).
int max_cta_pairs() {
  static int max_pairs = [] {
    using K = Cfg<256, false, false, false, true>;
    auto kern = gemm_tc_kernel<256, false, false, false, true>;
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (sm_count() / 2));
      cfg.blockDim = dim3(kGemmThreads);
      cfg.dynamicSmemBytes = K::SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    if (n <= 0 || n > sm_count() / 2) n = sm_count() / 2;
    if (getenv("DL_GEMM_VERBOSE")) fprintf(stderr, "dl_gemm: %d co-resident CTA pairs on %d SMs\n", n, sm_count());
    return n;
  }();
  return max_pairs;
}

template <int BN, bool TF32, bool SPLIT = false, bool TMAEPI = false, bool CTA2 = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, long long batch,
           cudaStream_t stream, const CUtensorMap* tmC = nullptr, const CUtensorMap* tmIn = nullptr) {
  using C = Cfg<BN, TF32, SPLIT, TMAEPI, CTA2>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BN, TF32, SPLIT, TMAEPI, CTA2>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  });
  if (attr_err != cudaSuccess)
    return set_error((int)attr_err, "dl_gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  p.idesc = ptx::make_idesc(TF32, p.a_mn != 0, p.b_mn != 0, CTA2 ? 2 * BM : BM, BN);
  p.mt = ceil_div(ceil_div(p.M, BM), CTA2 ? 2 : 1);        // CTA pairs: tile PAIRS along M
  p.nt = ceil_div(p.N, BN);
  const long long total = (long long)p.mt * p.nt * batch * p.splits;
  DL_REQUIRE(total < (1ll << 31), "dl_gemm: too many tiles");
  p.nkb_all = ceil_div(p.K, C::KE);
  DL_REQUIRE((long long)p.nkb_all * p.splits < (1ll << 31), "dl_gemm: K x split_k too large");
  p.fd_perz.init((uint32_t)(p.mt * p.nt));
  p.fd_nt.init((uint32_t)p.nt);
  p.fd_splits.init((uint32_t)p.splits);
  p.fd_nb0.init((uint32_t)p.nb0);
  p.fd_nb1.init((uint32_t)p.nb1);
  p.fd_conv.init((uint32_t)(p.conv_cin > 0 ? p.conv_cin : 1));
  p.fd_kred.init((uint32_t)(p.kred_kpb > 0 ? p.kred_kpb : 1));
  p.total_tiles = (int)total;
  if constexpr (CTA2) {
    const int pairs = max_cta_pairs();
    const int grid = 2 * (int)(total < pairs ? total : pairs);
    (void)launch_cluster_k(gemm_tc_kernel<BN, TF32, SPLIT, TMAEPI, CTA2>, dim3(grid), dim3(kGemmThreads),
                           (size_t)C::SMEM, 2u, stream, tmA, tmB, tmC ? *tmC : tmA, tmIn ? *tmIn : tmA, p);
  } else {
    const int grid = (int)(total < sm_count() ? total : sm_count());
    DL_LAUNCH((gemm_tc_kernel<BN, TF32, SPLIT, TMAEPI, CTA2>), grid, kGemmThreads, C::SMEM, stream, tmA, tmB,
              tmC ? *tmC : tmA, tmIn ? *tmIn : tmA, p);
  }
  DL_LAUNCH_CHECK("gemm_tc_kernel");
  count_launch();
  return 0;
}

}  // namespace
}  // namespace dl

extern "C" int dl_gemm(const dl_gemm_args* a, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DL_REQUIRE(a != nullptr, "dl_gemm: null args");
  DL_REQUIRE(a->A && a->B && a->C, "dl_gemm: A, B and C must be non-null");
  DL_REQUIRE(a->M >= 0 && a->N >= 0 && a->K >= 1, "dl_gemm: bad shape M=%lld N=%lld K=%lld",
             (long long)a->M, (long long)a->N, (long long)a->K);
  DL_REQUIRE(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), "dl_gemm: extent too large");
  DL_REQUIRE(a->dtype_ab == DL_F32 || a->dtype_ab == DL_BF16, "dl_gemm: dtype_ab must be DL_F32 or DL_BF16");
  DL_REQUIRE(a->dtype_c == DL_F32 || a->dtype_c == DL_BF16, "dl_gemm: dtype_c must be DL_F32 or DL_BF16");
  DL_REQUIRE(a->batch[0] >= 1 && a->batch[1] >= 1 && a->batch[2] >= 1, "dl_gemm: batch extents must be >= 1");
  const int KE_ = a->dtype_ab == DL_F32 ? 32 : 64;
  // conv modes (implicit-GEMM conv1d; see the header)
  const bool conv = a->conv_taps > 0, kred = a->kred != 0;
  int conv_cin = 0, kred_kpb = 0;
  if (conv) {
    DL_REQUIRE(!a->trans_a && !kred, "dl_gemm: conv_taps needs a K-major A");
    DL_REQUIRE(a->K % a->conv_taps == 0 && (a->K / a->conv_taps) % KE_ == 0,
               "dl_gemm: conv needs K = taps * channels with channels a multiple of %d", KE_);
    conv_cin = (int)(a->K / a->conv_taps);
  }
  if (kred) {
    DL_REQUIRE(a->trans_a && a->trans_b, "dl_gemm: kred needs both operands MN-major");
    DL_REQUIRE(a->sa[2] > 0 && a->sb[2] > 0, "dl_gemm: kred needs batch[2] strides for A and B");
    kred_kpb = ceil_div(a->K, KE_);
  }
  // with kred, batch[2] is a reduction dimension: it does not multiply the output tiles
  const long long batch = a->batch[0] * a->batch[1] * (kred ? 1 : a->batch[2]);
  const long long k_total = kred ? (long long)a->batch[2] * kred_kpb * KE_ : a->K;
  DL_REQUIRE(k_total < (1ll << 31), "dl_gemm: reduction extent too large");
  DL_REQUIRE(batch < (1ll << 24), "dl_gemm: batch %lld too large", batch);
  DL_REQUIRE(a->ldc >= a->N, "dl_gemm: ldc < N");
  DL_REQUIRE(a->mul_mode == DL_MUL_NONE || a->mul_aux != nullptr, "dl_gemm: mul_mode set without mul_aux");
  DL_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "dl_gemm: drop_p must be in [0, 1)");
  DL_REQUIRE(a->act >= 0 && a->act <= 2 && a->mul_mode >= 0 && a->mul_mode <= 3, "dl_gemm: bad act / mul_mode");
  DL_REQUIRE(a->pre_mode == 0 || (a->pre_mode == 1 && a->mul_mode == DL_MUL_NONE),
             "dl_gemm: pre_mode must be 0 or 1 (1 excludes mul_aux)");
  if (a->M == 0 || a->N == 0) return 0;
  const bool f32 = a->dtype_ab == DL_F32;
  const int sms = sm_count();

  int bn = a->tile_n;
  if (bn == 0) {
    // widest tile that still gives every SM a tile (wider tiles halve the operand re-reads and
    // measured 1.1 PFLOP/s at 128x256 vs 0.74 at 128x128 on 8192^3)
    const long long mt = ceil_div(a->M, BM);
    auto tiles_at = [&](int w) { return mt * ceil_div(a->N, w) * batch; };
    const int widest = a->N > 128 ? 256 : a->N > 64 ? 128 : 64;
    // weight-gradient shapes get their parallelism from split-K: the widest tile that still leaves
    // four output tiles (wider tiles re-read less of the long-K operands from L2; the split slices
    // land with 16-byte vector reductions).  Measured on B200, K = 16384: 512x512 40.7 us at BN=64,
    // 24.6 at 256; 256x256 16.7 / 12.1 (128) / 14.1; 128x256 best at 64.
    const bool split_candidate = a->dtype_c == DL_F32 && !a->bias && !a->preact_out && !a->mul_aux &&
                                 !a->residual && a->act == DL_ACT_NONE && a->drop_p == 0.f &&
                                 a->split_k != 1 && ceil_div(k_total, KE_) >= 16;
    bn = 0;
    if (split_candidate) {
      int w = widest;
      while (w > 64 && tiles_at(w) < 4) w >>= 1;
      if (tiles_at(w) * 2 <= sms) bn = w;
    }
    if (bn == 0) {
      // otherwise minimise waves x per-tile operand traffic (a 128 x w tile loads 128 + w rows per
      // k-block): wide tiles unless they leave most SMs idle.  8192^3: 1.11 PFLOP/s at 256, 0.73 at 128.
      long long best = -1;
      for (int w = widest; w >= 64; w >>= 1) {
        const long long cost = ceil_div(tiles_at(w), (long long)sms) * (BM + w);
        if (best < 0 || cost < best) { best = cost; bn = w; }
      }
    }
  }
  DL_REQUIRE(bn == 64 || bn == 128 || bn == 256, "dl_gemm: tile_n must be 0, 64, 128 or 256");

  // split-K: weight-gradient shaped problems (few output tiles, very long K) would otherwise
  // occupy a handful of SMs.  Only for plain fp32 outputs: slices are accumulated with atomics
  // into a zero-initialised C.
  // CTA pairs (cta_group::2, see Cfg): bf16 operands, tiles at least 128 wide, at least two row tiles.
  // DL_GEMM_CTA2: 0 = never, 1 = where it measured faster (default), 2 = wherever it is legal.
  static const int cta2_mode = [] { const char* e = getenv("DL_GEMM_CTA2"); return e ? atoi(e) : 1; }();
  static const int dbg_mode0 = [] { const char* e = getenv("DL_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  const int mt_all = ceil_div(a->M, BM);
  bool cta2 = cta2_mode != 0 && !f32 && bn >= 128 && mt_all >= 2 && (dbg_mode0 == 0 || dbg_mode0 == 5) && a->tile_n >= 0;
  const int nkb = ceil_div(k_total, KE_);
  if (cta2 && cta2_mode == 1) {
    // Measured on B200 over the step's shapes (profiles/r2t_gemm_cta_pairs.txt): a pair launch costs ~1 us more
    // (cluster scheduling, two cluster barriers) and saves a third of the operand traffic per K-block, so it
    // pays for 256-wide tiles once the launch runs more than one wave of tiles or a unit is at least 24
    // K-blocks long (16384 x 512 x 2048: 35.7 -> 33.1 us, 8192^3: 1.29 -> 1.46 PFLOP/s); the 1-wave, K <= 1024
    // launches and the short split-K slices stay single-CTA.
    const long long tiles1 = (long long)ceil_div(a->N, bn) * mt_all * batch;
    long long splits1 = a->split_k > 1 ? a->split_k : 1;
    if (a->split_k == 0 && tiles1 * 2 <= sms && nkb >= 16) {
      splits1 = sms / tiles1;
      if (splits1 > nkb / 4) splits1 = nkb / 4;
    }
    cta2 = bn == 256 && ((mt_all % 2 == 0) || mt_all >= 9) && (tiles1 > sms || nkb / splits1 >= 24);
  }
  const int unit_sms = cta2 ? max_cta_pairs() : sms;     // schedulable units: CTAs, or co-resident CTA pairs
  int splits = 1;
  const long long tiles = (long long)ceil_div(a->N, bn) * (cta2 ? ceil_div(mt_all, 2) : mt_all) * batch;
  // batched outputs can be split too when C is one dense [batch, M, N] block (a single memset)
  const bool dense_c = batch == 1 || (a->ldc == a->N && a->sc[0] == a->M * a->N &&
                                      (a->batch[1] == 1 || a->sc[1] == a->M * a->N * a->batch[0]) &&
                                      (kred || a->batch[2] == 1 || a->sc[2] == a->M * a->N * a->batch[0] * a->batch[1]));
  const bool plain = a->dtype_c == DL_F32 && !a->bias && !a->preact_out && !a->mul_aux && !a->residual &&
                     a->act == DL_ACT_NONE && a->drop_p == 0.f && dense_c;
  if (a->split_k > 1) {
    DL_REQUIRE(plain, "dl_gemm: split_k needs a plain fp32 output (no epilogue operands; batched C must be dense)");
    splits = a->split_k;
  } else if (a->split_k == 0 && plain && tiles * 2 <= unit_sms && nkb >= 16) {
    // one (tile, K-slice) unit per CTA and never more units than SMs: a 149th unit would run as a second
    // wave on its own (measured, 1024 x 256 x 16384: 19 slices = 152 units 20.1 us, 18 slices = 144 units
    // 14.7 us).  More, shorter slices do not pay either: every unit ends in a 128 x BN fp32 reduction
    // into L2 (~5 us per wave of 128 KB tiles), the same L2 the operand feed is bound by.
    splits = (int)(unit_sms / tiles);
    if (splits > nkb / 4) splits = nkb / 4;
  }
  if (splits > nkb) splits = nkb;
  if (splits < 1) splits = 1;
  if (a->accumulate) {
    DL_REQUIRE(plain, "dl_gemm: accumulate needs a plain fp32 output (no epilogue operands)");
  }
  if (splits > 1 && !a->accumulate) {
    if (batch == 1)
      DL_CUDA(cudaMemset2DAsync(a->C, (size_t)a->ldc * 4, 0, (size_t)a->N * 4, (size_t)a->M, stream));
    else
      DL_CUDA(cudaMemsetAsync(a->C, 0, (size_t)(a->M * a->N * batch) * 4, stream));
  }

  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, a->A, f32, a->trans_a != 0, a->M, conv ? conv_cin : a->K, a->lda, a->batch, a->sa, BM, "A");
  if (rc) return rc;
  rc = make_operand_map(&tmB, a->B, f32, a->trans_b != 0, a->N, a->K, a->ldb, a->batch, a->sb, cta2 ? bn / 2 : bn, "B");
  if (rc) return rc;

  GemmParams p;
  p.C = a->C; p.bias = a->bias; p.preact = a->preact_out; p.aux = a->mul_aux; p.res = a->residual;
  p.ldc = a->ldc;
  p.ldr = a->ldr != 0 ? a->ldr : a->ldc;
  for (int i = 0; i < 3; ++i) {
    p.sc[i] = a->sc[i];
    p.sr[i] = a->ldr != 0 ? a->sr[i] : a->sc[i];
    p.a_on[i] = a->sa[i] != 0;
    p.b_on[i] = a->sb[i] != 0;
  }
  p.drop_seed = a->drop_seed; p.drop_p = a->drop_p; p.drop_step = (const long long*)a->drop_seed_step;
  p.M = (int)a->M; p.N = (int)a->N; p.K = (int)k_total;
  p.nb0 = (int)a->batch[0]; p.nb1 = (int)a->batch[1];
  p.conv_cin = conv_cin; p.conv_left = a->conv_left;
  p.kred_kpb = kred_kpb; p.kred_shift = a->kred_shift; p.kred_tap = a->kred == 1;
  if (kred) { p.a_on[2] = p.b_on[2] = 1; }
  if (a->accumulate && splits == 1) {     // C += result through the residual path (same layout)
    p.res = a->C;
    p.ldr = a->ldc;
    for (int i = 0; i < 3; ++i) p.sr[i] = a->sc[i];
  }
  p.splits = splits;
  p.colsum_a = a->colsum_a;
  if (a->colsum_a) {
    DL_REQUIRE(!f32 && a->trans_a && batch == 1 && !conv && !kred,
               "dl_gemm: colsum_a needs bf16 operands, an MN-major A (trans_a = 1) and no batch / conv modes");
  }
  p.a_mn = a->trans_a != 0; p.b_mn = a->trans_b != 0;
  p.c_bf16 = a->dtype_c == DL_BF16;
  p.act = a->act; p.mul_mode = a->mul_mode; p.alpha = a->alpha;
  p.pre_mode = a->pre_mode;
  static const int dbg_mode = [] { const char* e = getenv("DL_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg_mode;
  p.idesc = 0;
  p.tma_in = 0;
  p.in_on[0] = p.in_on[1] = p.in_on[2] = 0;
  // Epilogue through shared memory + bulk tensor stores (see Cfg): bf16 output tiles of width 256 whose
  // main loop is short enough (K <= 1024) for the epilogue to matter, exactly one elementwise input.
  static const bool tma_epi_off = [] { const char* e = getenv("DL_GEMM_NO_TMA_EPI"); return e && atoi(e) != 0; }();
  auto al8 = [](long long v) { return v % 8 == 0; };
  // Measured on B200 (same box, DL_GEMM_NO_TMA_EPI A/B): with an elementwise input the slab path wins
  // (16384x1024x256 * aux 29.0 -> 26.6 us, 16384x2048x512 * aux 56.1 -> 49.7 us: the input's HBM latency
  // hides behind the MMAs instead of sitting in front of every 16-column chunk); without one the
  // slab's store -> reuse chain is longer than a K = 256 main loop and the direct stores win, so those
  // launches keep them.
  const bool tma_epi = !tma_epi_off && (a->mul_aux || a->residual) && !(a->preact_out && a->pre_mode == 1) &&
                       !f32 && p.c_bf16 && bn == 256 && splits == 1 && !conv && !kred &&
                       (dbg_mode == 0 || dbg_mode == 5) && nkb <= 16 && ((uintptr_t)a->C & 15) == 0 && al8(a->ldc) && al8(a->sc[0]) &&
                       al8(a->sc[1]) && al8(a->sc[2]) && !(a->mul_aux && a->residual) &&
                       (!a->mul_aux || ((uintptr_t)a->mul_aux & 15) == 0) &&
                       (!a->residual || (((uintptr_t)a->residual & 15) == 0 && al8(p.ldr) && al8(p.sr[0]) &&
                                         al8(p.sr[1]) && al8(p.sr[2])));
  if (tma_epi) {
    CUtensorMap tmC, tmIn;
    rc = make_epi_map(&tmC, a->C, a->N, a->M, a->ldc, a->batch, p.sc, "C");
    if (rc) return rc;
    if (a->mul_aux) {
      p.tma_in = 1;
      for (int i = 0; i < 3; ++i) p.in_on[i] = p.sc[i] != 0;
      rc = make_epi_map(&tmIn, a->mul_aux, a->N, a->M, a->ldc, a->batch, p.sc, "mul_aux");
    } else if (a->residual) {
      p.tma_in = 2;
      for (int i = 0; i < 3; ++i) p.in_on[i] = p.sr[i] != 0;
      rc = make_epi_map(&tmIn, a->residual, a->N, a->M, p.ldr, a->batch, p.sr, "residual");
    }
    if (rc) return rc;
    if (cta2) return launch<256, false, false, true, true>(tmA, tmB, p, batch, stream, &tmC, p.tma_in ? &tmIn : nullptr);
    return launch<256, false, false, true>(tmA, tmB, p, batch, stream, &tmC, p.tma_in ? &tmIn : nullptr);
  }
  if (cta2) {
    if (bn == 128) return launch<128, false, false, false, true>(tmA, tmB, p, batch, stream);
    return launch<256, false, false, false, true>(tmA, tmB, p, batch, stream);
  }
  if (f32 && a->precise) {
    if (bn == 64) return launch<64, true, true>(tmA, tmB, p, batch, stream);
    if (bn == 128) return launch<128, true, true>(tmA, tmB, p, batch, stream);
    return launch<256, true, true>(tmA, tmB, p, batch, stream);
  }
  if (f32) {
    if (bn == 64) return launch<64, true>(tmA, tmB, p, batch, stream);
    if (bn == 128) return launch<128, true>(tmA, tmB, p, batch, stream);
    return launch<256, true>(tmA, tmB, p, batch, stream);
  }
  if (bn == 64) return launch<64, false>(tmA, tmB, p, batch, stream);
  if (bn == 128) return launch<128, false>(tmA, tmB, p, batch, stream);
  return launch<256, false>(tmA, tmB, p, batch, stream);
}
