/* libdruglamp_sm100.so -- C ABI of the B200 (sm_100a) kernels behind DrugLAMP's cross-modal
 * fusion + contrastive hot path.
 *
 * The reference (Lzcstan/DrugLAMP) is pure Python/PyTorch and has no FFI layer of its own: the
 * arithmetic below is what its nn.Modules hand to ATen/cuBLAS/DGL.
 * Each entry point cites the reference call site it replaces.  The host-side mirror of the
 * reference's module API (druglamp_b200/*.py) binds these with ctypes; INTEGRATION.md shows the
 * reference-side patch.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All pointers are DEVICE pointers unless noted.
 *  - Caller owns every buffer; the library never allocates, frees or retains them.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 *    device synchronisation and is re-entrant.
 *  - Return value: 0 = OK, <0 = argument / shape / alignment error, >0 = cudaError_t.
 *    dl_last_error() returns the thread-local message of the last failing call.
 *  - dtype codes: DL_F32 = 0, DL_BF16 = 1.  Matrices are row-major.
 *  - No CPU fallback exists: without an sm_100 device every compute call fails.
 */
#ifndef DRUGLAMP_SM100_H_
#define DRUGLAMP_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DL_F32 0
#define DL_BF16 1

#define DL_ACT_NONE 0
#define DL_ACT_GELU 1 /* exact erf GELU (nn.GELU default) */
#define DL_ACT_RELU 2

#define DL_MUL_NONE 0
#define DL_MUL_GELU_GRAD 1 /* out *= gelu'(aux)          */
#define DL_MUL_RELU_MASK 2 /* out *= (aux > 0)            */
#define DL_MUL_VALUE 3     /* out *= aux                  */

int dl_version(void);
const char* dl_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
int64_t dl_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05.mma, TMA-fed, accumulators in TMEM) with a fused epilogue.
 *
 *   C[b] = epilogue( alpha * op(A[b]) . op(B[b]) )          b = b_hi * batch_lo + b_lo
 *     op(A) is M x K:  trans_a = 0 -> A stored [M, K] (K contiguous), 1 -> stored [K, M]
 *     op(B) is K x N:  trans_b = 0 -> B stored [N, K] (K contiguous, nn.Linear weight layout),
 *                      1 -> stored [K, N]
 *   epilogue(v):  v += bias[n];  if preact_out: preact_out = v;  v = act(v);
 *                 v *= f(mul_aux) (mul_mode);  v += residual;  C = v
 *   dtype_ab: DL_BF16 (kind::f16) or DL_F32 (kind::tf32); fp32 accumulation always.
 *   dtype_c applies to C, preact_out, mul_aux and residual (all share C's layout: ldc, sc_*).
 *   bias is fp32.  lda/ldb/ldc and batch strides are in ELEMENTS; every row stride and batch
 *   stride of A and B must be a multiple of 16 bytes and the base pointers 16-byte aligned
 *   (TMA).  A batch stride of 0 broadcasts that operand.
 *
 * Replaces: every nn.Linear / torch.bmm on the hot path -- PGCA in/out projections and
 * q.k^T / p.v (reference model/PGCA/guided_cross_attention_model.py:147-162,290,311-314),
 * MHLA lin1/lin2 (model/PMMA/encoder.py:128-131), PMMA projections, scores, fc/out and FFN
 * (model/PMMA/attention.py:91-98,58-83; model/PMMA/mlp.py:45-49; model/PMMA/embed.py:49),
 * GraphConv / residual / init_transform matmuls (model/basic_model.py:149,620,431), and the
 * CrossModality latent products (model/cross_modality.py:151-162).
 */
typedef struct dl_gemm_args {
  const void* A;
  const void* B;
  void* C;
  const float* bias;    /* [N] or NULL */
  void* preact_out;     /* or NULL */
  const void* mul_aux;  /* or NULL */
  const void* residual; /* or NULL; may alias C */
  int64_t M, N, K;
  int64_t lda, ldb, ldc;
  int64_t batch_lo, batch_hi;
  int64_t sa_lo, sa_hi, sb_lo, sb_hi, sc_lo, sc_hi;
  float alpha;
  int32_t dtype_ab, dtype_c;
  int32_t trans_a, trans_b;
  int32_t act, mul_mode;
  int32_t tile_n; /* 0 = auto, else 64 / 128 / 256 */
} dl_gemm_args;

int dl_gemm(const dl_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRUGLAMP_SM100_H_ */
