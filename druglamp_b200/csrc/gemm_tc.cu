// dl_gemm: TMA-fed tcgen05 GEMM with fp32 accumulators in tensor memory and a fused epilogue.
//
// One CTA computes one 128 x BN tile of C.  Warp roles (192 threads):
//   warp 0      TMA producer: fills a ring of kStages {A tile, B tile} buffers (SWIZZLE_128B)
//   warp 1      allocates TMEM, issues tcgen05.mma (one thread), commits to mbarriers
//   warps 2..5  epilogue: tcgen05.ld the accumulator quadrant they own (TMEM lanes 32*(warp%4)..),
//               transpose through shared memory so global traffic is coalesced, apply
//               bias / activation / auxiliary multiply / dropout / residual, store C.
// A K-block is 128 bytes of K (64 bf16 or 32 tf32 values) = four tcgen05.mma instructions.
// Both operands may be K-major or MN-major (transposed storage); the shared-memory tile is
// always "rows x 128 B" so only the descriptors and TMA boxes differ (see ptx.cuh).
// Up to three batch dimensions (e.g. head, query-set, pair) map onto a 5-D tensor map.
#include <mutex>

#include "../../include/druglamp_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace dl {
void count_launch(int n = 1);

namespace {

constexpr int BM = 128;
constexpr int kGemmThreads = 192;

struct GemmParams {
  void* C;
  const float* bias;
  void* preact;
  const void* aux;
  const void* res;
  long long ldc, sc[3];
  long long ldr, sr[3];
  unsigned long long drop_seed;
  float drop_p;
  int M, N, K;
  int nb0, nb1;                 // extents of the two fastest batch dims
  int a_on[3], b_on[3];         // 0 when that batch stride is a broadcast
  int a_mn, b_mn;
  int c_bf16;
  int act, mul_mode;
  float alpha;
  uint32_t idesc;
};

// SPLIT (fp32 operands only): "3xTF32".  kind::tf32 reads just the top 19 bits of each fp32
// operand, which costs ~1e-3 relative accuracy.  In SPLIT mode the four otherwise idle epilogue
// warps rewrite every landed stage in place as hi = x & 0xffffe000 and lo = x - hi (exact), and
// the issuer runs three MMAs per K step (hi*hi + lo*hi + hi*lo): fp32-grade products on the
// tensor cores.  The conversion is elementwise, so it is independent of the swizzled layout.
template <int BN, bool TF32, bool SPLIT = false>
struct Cfg {
  static constexpr int ELEM = TF32 ? 4 : 2;
  static constexpr int KE = 128 / ELEM;   // K elements per K-block
  static constexpr int UK = 32 / ELEM;    // K elements per tcgen05.mma
  static constexpr int MNB = 128 / ELEM;  // MN elements per 128-byte block (MN-major operands)
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int LOAD_BYTES = A_BYTES + B_BYTES;             // what TMA delivers per stage
  static constexpr int STAGE_BYTES = SPLIT ? 2 * LOAD_BYTES : LOAD_BYTES;
  static constexpr int STAGES = SPLIT ? ((BN == 256) ? 2 : (BN == 128 ? 3 : 4))
                                      : ((BN == 256) ? 4 : (BN == 128 ? 3 : 4));
  static constexpr int EPI_BYTES = 4 * 32 * 33 * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static_assert(EPI_BYTES <= STAGES * STAGE_BYTES, "epilogue staging aliases the stage ring");
};

template <int BN, bool TF32, bool SPLIT>
__global__ void __launch_bounds__(kGemmThreads)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  using C = Cfg<BN, TF32, SPLIT>;
  static_assert(!SPLIT || TF32, "SPLIT applies to fp32 operands");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_addr - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  const uint32_t bar_full = ptx::smem_u32(bars);                    // [STAGES]
  const uint32_t bar_empty = bar_full + 8 * C::STAGES;              // [STAGES]
  const uint32_t bar_acc = bar_empty + 8 * C::STAGES;               // accumulator ready
  const uint32_t bar_conv = bar_acc + 8;                            // [STAGES] hi/lo split done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int b0 = blockIdx.z % p.nb0;
  const int b1 = (blockIdx.z / p.nb0) % p.nb1;
  const int b2 = blockIdx.z / (p.nb0 * p.nb1);
  const int nkb = (p.K + C::KE - 1) / C::KE;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
      ptx::mbar_init(bar_conv + 8 * s, 128);
    }
    ptx::mbar_init(bar_acc, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<BN>(ptx::smem_u32(tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      const int a2 = p.a_on[0] ? b0 : 0, a3 = p.a_on[1] ? b1 : 0, a4 = p.a_on[2] ? b2 : 0;
      const int c2 = p.b_on[0] ? b0 : 0, c3 = p.b_on[1] ? b1 : 0, c4 = p.b_on[2] ? b2 : 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t full = bar_full + 8 * s;
        ptx::mbar_arrive_expect_tx(full, C::LOAD_BYTES);
        const uint32_t sA = base_addr + s * C::STAGE_BYTES, sB = sA + C::A_BYTES;
        const int k0 = kb * C::KE;
        if (!p.a_mn) {
          ptx::tma_load_5d(sA, &tmA, full, k0, m0, a2, a3, a4);
        } else {
#pragma unroll
          for (int blk = 0; blk < BM / C::MNB; ++blk)
            ptx::tma_load_5d(sA + blk * C::KE * 128, &tmA, full, m0 + blk * C::MNB, k0, a2, a3, a4);
        }
        if (!p.b_mn) {
          ptx::tma_load_5d(sB, &tmB, full, k0, n0, c2, c3, c4);
        } else {
#pragma unroll
          for (int blk = 0; blk < BN / C::MNB; ++blk)
            ptx::tma_load_5d(sB + blk * C::KE * 128, &tmB, full, n0 + blk * C::MNB, k0, c2, c3, c4);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        ptx::mbar_wait((SPLIT ? bar_conv : bar_full) + 8 * s, ph);
        ptx::tc_fence_after();
        const uint32_t sA = base_addr + s * C::STAGE_BYTES, sB = sA + C::A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // MN-major: bf16 uses SWIZZLE_128B (8-row K atom), tf32 must use 128B_BASE32B (4-row)
          constexpr uint32_t kMnSbo = TF32 ? 512 : 1024, kMnType = TF32 ? 1 : 2;
          const uint64_t ad = p.a_mn ? ptx::make_smem_desc(sA + k * C::UK * 128, C::KE * 128, kMnSbo, kMnType)
                                     : ptx::make_smem_desc(sA + k * 32, 16, 1024);
          const uint64_t bd = p.b_mn ? ptx::make_smem_desc(sB + k * C::UK * 128, C::KE * 128, kMnSbo, kMnType)
                                     : ptx::make_smem_desc(sB + k * 32, 16, 1024);
          ptx::mma_ss<TF32>(tmem, ad, bd, p.idesc, (uint32_t)((kb | k) != 0));
          if constexpr (SPLIT) {
            // descriptors address 16-byte units: the lo copies sit LOAD_BYTES above the hi ones
            constexpr uint64_t kLo = (uint64_t)(C::LOAD_BYTES >> 4);
            ptx::mma_ss<TF32>(tmem, ad + kLo, bd, p.idesc, 1u);
            ptx::mma_ss<TF32>(tmem, ad, bd + kLo, p.idesc, 1u);
          }
        }
        ptx::mma_commit(bar_empty + 8 * s);   // frees the stage once these MMAs have read it
      }
      ptx::mma_commit(bar_acc);
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    float* t = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
    if constexpr (SPLIT) {
      const int tid = threadIdx.x - 64;           // 0..127
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        ptx::mbar_wait(bar_full + 8 * s, ph);
        uint4* hi = reinterpret_cast<uint4*>(smem + s * C::STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + s * C::STAGE_BYTES + C::LOAD_BYTES);
#pragma unroll 4
        for (int i = tid; i < C::LOAD_BYTES / 16; i += 128) {
          const uint4 u = hi[i];
          const uint4 h = make_uint4(u.x & 0xffffe000u, u.y & 0xffffe000u, u.z & 0xffffe000u,
                                     u.w & 0xffffe000u);
          lo[i] = make_float4(__uint_as_float(u.x) - __uint_as_float(h.x),
                              __uint_as_float(u.y) - __uint_as_float(h.y),
                              __uint_as_float(u.z) - __uint_as_float(h.z),
                              __uint_as_float(u.w) - __uint_as_float(h.w));
          hi[i] = h;
        }
        ptx::fence_proxy_async();                 // generic-proxy writes -> visible to the MMA
        ptx::mbar_arrive(bar_conv + 8 * s);
      }
    }
    ptx::mbar_wait(bar_acc, 0);
    ptx::tc_fence_after();
    const long long cbase = (long long)b0 * p.sc[0] + (long long)b1 * p.sc[1] + (long long)b2 * p.sc[2];
    const long long rbase = (long long)b0 * p.sr[0] + (long long)b1 * p.sr[1] + (long long)b2 * p.sr[2];
    const float drop_inv = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    float* Cf = reinterpret_cast<float*>(p.C);
    __nv_bfloat16* Ch = reinterpret_cast<__nv_bfloat16*>(p.C);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int col = n0 + c0 + lane;
      const bool col_ok = col < p.N;
      const float bias_v = (p.bias != nullptr && col_ok) ? p.bias[col] : 0.f;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int row = m0 + q * 32 + r;
        if (row < p.M && col_ok) {
          const long long off = cbase + (long long)row * p.ldc + col;
          const long long roff = rbase + (long long)row * p.ldr + col;
          float keep = 1.f;
          if (p.drop_p > 0.f) {
            const unsigned long long e = ((unsigned long long)blockIdx.z * p.M + row) * p.N + col;
            keep = hash_uniform(p.drop_seed, e) >= p.drop_p ? drop_inv : 0.f;
          }
          float x = t[r * 33 + lane] * p.alpha + bias_v;
          if (p.c_bf16) {
            if (p.preact) reinterpret_cast<__nv_bfloat16*>(p.preact)[off] = __float2bfloat16_rn(x);
            if (p.act == DL_ACT_GELU) x = gelu_erf(x);
            else if (p.act == DL_ACT_RELU) x = fmaxf(x, 0.f);
            if (p.mul_mode != DL_MUL_NONE) {
              const float a = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.aux)[off]);
              x *= (p.mul_mode == DL_MUL_GELU_GRAD) ? gelu_erf_grad(a)
                   : (p.mul_mode == DL_MUL_RELU_MASK) ? (a > 0.f ? 1.f : 0.f) : a;
            }
            x *= keep;
            if (p.res) x += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res)[roff]);
            Ch[off] = __float2bfloat16_rn(x);
          } else {
            if (p.preact) reinterpret_cast<float*>(p.preact)[off] = x;
            if (p.act == DL_ACT_GELU) x = gelu_erf(x);
            else if (p.act == DL_ACT_RELU) x = fmaxf(x, 0.f);
            if (p.mul_mode != DL_MUL_NONE) {
              const float a = reinterpret_cast<const float*>(p.aux)[off];
              x *= (p.mul_mode == DL_MUL_GELU_GRAD) ? gelu_erf_grad(a)
                   : (p.mul_mode == DL_MUL_RELU_MASK) ? (a > 0.f ? 1.f : 0.f) : a;
            }
            x *= keep;
            if (p.res) x += reinterpret_cast<const float*>(p.res)[roff];
            Cf[off] = x;
          }
        }
      }
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<BN>(tmem);
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

template <int BN, bool TF32, bool SPLIT = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, long long batch,
           cudaStream_t stream) {
  using C = Cfg<BN, TF32, SPLIT>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BN, TF32, SPLIT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  });
  if (attr_err != cudaSuccess)
    return set_error((int)attr_err, "dl_gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  p.idesc = ptx::make_idesc(TF32, p.a_mn != 0, p.b_mn != 0, BM, BN);
  dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, BM), (unsigned)batch);
  gemm_tc_kernel<BN, TF32, SPLIT><<<grid, kGemmThreads, C::SMEM, stream>>>(tmA, tmB, p);
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
  const long long batch = a->batch[0] * a->batch[1] * a->batch[2];
  DL_REQUIRE(batch <= 65535, "dl_gemm: batch %lld exceeds 65535", batch);
  DL_REQUIRE(a->ldc >= a->N, "dl_gemm: ldc < N");
  DL_REQUIRE(a->mul_mode == DL_MUL_NONE || a->mul_aux != nullptr, "dl_gemm: mul_mode set without mul_aux");
  DL_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "dl_gemm: drop_p must be in [0, 1)");
  DL_REQUIRE(a->act >= 0 && a->act <= 2 && a->mul_mode >= 0 && a->mul_mode <= 3, "dl_gemm: bad act / mul_mode");
  if (a->M == 0 || a->N == 0) return 0;

  int bn = a->tile_n;
  if (bn == 0) {
    bn = (a->N <= 64) ? 64 : 128;
    // keep at least ~one wave of CTAs when the problem is small
    if (bn == 128 && (long long)ceil_div(a->N, 128) * ceil_div(a->M, BM) * batch < sm_count()) bn = 64;
  }
  DL_REQUIRE(bn == 64 || bn == 128 || bn == 256, "dl_gemm: tile_n must be 0, 64, 128 or 256");
  const bool f32 = a->dtype_ab == DL_F32;

  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, a->A, f32, a->trans_a != 0, a->M, a->K, a->lda, a->batch, a->sa, BM, "A");
  if (rc) return rc;
  rc = make_operand_map(&tmB, a->B, f32, a->trans_b != 0, a->N, a->K, a->ldb, a->batch, a->sb, bn, "B");
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
  p.drop_seed = a->drop_seed; p.drop_p = a->drop_p;
  p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K;
  p.nb0 = (int)a->batch[0]; p.nb1 = (int)a->batch[1];
  p.a_mn = a->trans_a != 0; p.b_mn = a->trans_b != 0;
  p.c_bf16 = a->dtype_c == DL_BF16;
  p.act = a->act; p.mul_mode = a->mul_mode; p.alpha = a->alpha;
  p.idesc = 0;
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
