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
#define DL_ACT_GELU 1 /* nn.GELU: erf form for fp32 activations; bf16 activations use
                         the tanh form on the hardware tanh (|diff| <= 4.8e-4, a fraction of a bf16 ulp) */
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
 *   C[b] = epilogue( alpha * op(A[b]) . op(B[b]) )          b = (b0, b1, b2): three batch dims
 *     op(A) is M x K:  trans_a = 0 -> A stored [M, K] (K contiguous), 1 -> stored [K, M]
 *     op(B) is K x N:  trans_b = 0 -> B stored [N, K] (K contiguous, nn.Linear weight layout),
 *                      1 -> stored [K, N]
 *   epilogue(v):  v += bias[n];  if preact_out: preact_out = v;  v = act(v);
 *                 v *= f(mul_aux) (mul_mode);  v = dropout(v);  v += residual;  C = v
 *   dropout keeps element e = (b*M + m)*N + n iff hash(drop_seed, e) >= drop_p and scales by
 *   1/(1-drop_p) (same mask as dl_dropout on a contiguous [batch*M, N] tensor).
 *   residual has its own layout (ldr, sr[]; ldr = 0 means "same as C"); a batch stride
 *   of 0 with ldr != 0 broadcasts it (positional embeddings).
 *   dtype_ab: DL_BF16 (kind::f16) or DL_F32 (kind::tf32); fp32 accumulation always.  With
 *   precise = 1 fp32 operands are split in shared memory into hi + lo tf32 parts and three MMAs
 *   per K step recover fp32-grade accuracy (what the fp32 parity mode uses).
 *   dtype_c applies to C, preact_out, mul_aux and residual (all share C's layout: ldc, sc[]).
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
  int64_t batch[3];     /* batch extents, fastest first; b = (b2*batch[1] + b1)*batch[0] + b0 */
  int64_t sa[3], sb[3], sc[3]; /* batch strides (elements) of A, B, C */
  int64_t ldr, sr[3];   /* residual layout; ldr = 0 -> same as C */
  uint64_t drop_seed;
  float drop_p;
  float alpha;
  int32_t dtype_ab, dtype_c;
  int32_t trans_a, trans_b;
  int32_t act, mul_mode;
  int32_t tile_n;  /* 0 = auto, else 64 / 128 / 256 */
  int32_t precise; /* DL_F32 operands only: 1 = 3xTF32 split (fp32-grade products), 0 = plain TF32 */
  int32_t split_k; /* 0 = auto, 1 = off, n > 1 = n K-slices accumulated atomically into a zeroed fp32 C
                      (plain outputs only: no bias / activation / residual; batched C must be dense) */
  /* Implicit-GEMM conv1d ('same' padding, channels-last activations [batch, L, C]) -- ProteinCNN
   * (model/basic_model.py:155-180).  conv_taps = k > 0: A is the K-major activation [L, cin] per
   * batch, K = k * cin enumerates (tap, channel) and the A tile of tap t is rows m + t - conv_left,
   * rows outside [0, L) read as zero; B = weights [cout, k * cin].  kred = 1 (weight gradient): both
   * operands MN-major, K = L rows per batch[2] entry, batch[2] is REDUCED over, batch[0] is the
   * tap and shifts B's rows by b0 + kred_shift (zero outside [0, L)); C is [taps, M, N].
   * kred = 2: the same K-reduction over batch[2] without the tap shift (batch[0] is an ordinary
   * batch dim): sums over the stacked query sets in the paired attention's dK / dV. */
  int32_t conv_taps, conv_left;
  int32_t kred, kred_shift;
  int32_t accumulate; /* 1: C += result (plain fp32 outputs only) -- weight gradients land directly in
                         the flat gradient buffer instead of a temporary plus an add kernel */
  float* colsum_a;    /* or NULL.  bf16 operands, trans_a = 1, no batch: colsum_a[m] += sum_k A[k, m],
                         computed from the A tiles while they sit in shared memory.  For a weight
                         gradient dW = dY^T X (A = dY) this is the bias gradient, so nn.Linear's
                         backward needs no separate pass over dY (fp32 [M], accumulated into). */
  const int64_t* drop_seed_step; /* or NULL.  Device-resident step counter: the dropout seed of this launch
                         is splitmix64(drop_seed ^ f(*drop_seed_step)), read on the device, so a
                         captured CUDA graph draws a fresh mask on every replay (the same pointer is given
                         to the backward's dl_act_bwd / dl_gemm of that step). */
  int32_t pre_mode;   /* what preact_out receives.  0: the pre-activation v (after bias).  1: the derivative of
                         this epilogue's elementwise part, d dropout(act(v)) / dv = act'(v) * keep / (1 - p): the
                         backward's dX GEMM then multiplies by it (DL_MUL_VALUE) instead of re-evaluating
                         gelu' and the dropout hash per element -- the epilogue of the 4x-wide FFN gradient
                         (model/PMMA/mlp.py:45-49) is instruction-bound, the forward has the values at hand. */
} dl_gemm_args;

int dl_gemm(const dl_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused scaled-dot-product attention (bf16; tcgen05 score / p.v tiles with the softmax between
 * them done out of tensor memory; no score or probability map in HBM).  One call serves
 *   - GuidedCrossAttention's core (model/PGCA/guided_cross_attention_model.py:290,307-311):
 *     H = 1, d = 128, raw != NULL returns the scaled pre-softmax logits (:307,:319-320);
 *   - the paired attention of PMMA (model/PMMA/attention.py:44-88): S2 = 2 stacked query sets
 *     (own + other stream) against ONE key/value set, output columns [set*o_ss + h*d, +d) =
 *     cat(attn, attn_p) as :81 builds it;
 *   - plain multi-head self attention (model/PMMA/attention.py:109-122).
 * Layouts (elements, bf16, unit column stride; every stride a multiple of 8 elements):
 *   q[set, b, row, h*d + c] at q + set*q_ss + b*q_sb + row*q_ld;   k / v [b, key, h*d + c] likewise;
 *   o, d_o [b, row, set*o_ss + h*d + c];   dq / dk / dv mirror q / k / v with their own strides.
 *   lse: fp32 [S2, B, H, Lq], log2-domain row log-sum-exp of the scaled logits (saved by forward,
 *   read by backward).  raw: bf16 (B, H, Lq, raw_ld).
 * Limits: d in {64, 128}, Lk <= 512 (one query tile's whole score row lives in tensor memory).
 * Backward workspace: dvec fp32 [S2, B, H, Lq].  dq is written by bulk tensor stores (first key tile)
 * and bf16 reduce-adds (further key tiles); dq_accumulate != 0 makes every tile a reduce-add onto the
 * existing dq (the second K/V set of the paired block shares its query gradient with the first).
 */
typedef struct dl_attn_args {
  const void* q;
  const void* k;
  const void* v;
  void* o;            /* forward: output; backward: the forward's output */
  float* lse;         /* forward: output; backward: input */
  void* raw;          /* forward only, or NULL */
  const void* d_o;    /* backward */
  void* dq;
  void* dk;
  void* dv;
  float* dvec;        /* backward workspace */
  int64_t B, H, S2, Lq, Lk, d;
  int64_t q_ld, q_sb, q_ss;
  int64_t k_ld, k_sb, v_ld, v_sb;
  int64_t o_ld, o_sb, o_ss;
  int64_t dq_ld, dq_sb, dq_ss, dk_ld, dk_sb, dv_ld, dv_sb;
  int64_t raw_ld;
  float scale;
  int32_t dq_accumulate;
} dl_attn_args;

/* out[m, n] = aux[m, n] * sum_{k < K} g[m, k] * w[k, n]  (bf16; K <= 16; N a multiple of 8; g rows padded to
 * 8 (K <= 8) or 16 columns, row stride ldg; w [K, N] and aux / out [M, N] contiguous).  The input gradient of a
 * layer with a handful of outputs times the stored activation derivative: MultiHeadLinearAttention's lin2
 * (model/PMMA/encoder.py:128-131, 1024 -> 8 heads) in its backward -- HBM-bound, no tensor cores. */
int dl_smallk_mul(const void* g, const void* w, const void* aux, void* out, int64_t M, int32_t N, int32_t K,
                  int64_t ldg, void* stream);

int dl_attn_fwd(const dl_attn_args* args, void* stream);
int dl_attn_bwd(const dl_attn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused position-wise feed-forward network of a PMMA block (bf16, model width D = 256):
 * Mlp.forward (model/PMMA/mlp.py:44-50: fc1 -> GELU -> dropout -> fc2 -> dropout) together with the
 * block's residual add (model/PMMA/block.py:45-47,59-61), two chained tcgen05 GEMMs in ONE kernel per
 * direction; the 4x-wide hidden activation stays in shared / tensor memory between them.
 *
 *   dl_ffn_fwd:  hidden = dropout(gelu(x w1^T + b1), seed1)      [M, Dh]  (NULL: not stored, forward-only)
 *                dact   = d hidden / d (x w1^T + b1)              [M, Dh]  (NULL: not stored)
 *                y      = dropout(hidden w2^T + b2, seed2) + residual      [M, D]
 *   dl_ffn_bwd:  x = g, the gradient of (hidden w2^T + b2) (i.e. dL/dy with the seed2 mask applied,
 *                dl_act_bwd);  hidden (OUT) = dpre = (g w2) * dact   [M, Dh]  (the dW1 operand);
 *                y (OUT) = dX = dpre w1                               [M, D]
 *   The weight / bias gradients stay dl_gemm launches (dW2 = g^T hidden, dW1 = dpre^T x).
 *
 * x, y, residual: bf16 [M, D] with row strides ldx, ldy, ldr; hidden, dact: bf16 [M, Dh] with row
 * stride ldh; w1: bf16 [Dh, D] contiguous, w2: bf16 [D, Dh] contiguous (nn.Linear layouts); b1, b2
 * fp32.  Row strides are multiples of 16 elements, base pointers 32-byte aligned.  Dh is a multiple
 * of 128.  Dropout masks are the ones dl_gemm / dl_act_bwd draw for the same (seed, element index),
 * so either direction may be mixed with the unfused launches.
 */
typedef struct dl_ffn_args {
  const void* x;
  const void* w1;
  const void* w2;
  const float* b1;       /* forward */
  const float* b2;       /* forward */
  const void* residual;  /* forward, or NULL */
  void* hidden;
  void* dact;            /* forward: out (or NULL); backward: in */
  void* y;
  int64_t M, D, Dh;
  int64_t ldx, ldh, ldy, ldr;
  float drop_p;          /* forward */
  uint64_t seed1, seed2;
  const int64_t* drop_seed_step; /* or NULL; see dl_gemm_args.drop_seed_step */
} dl_ffn_args;

int dl_ffn_fwd(const dl_ffn_args* args, void* stream);
int dl_ffn_bwd(const dl_ffn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound; one warp per row, 128-bit vector loads, warp-shuffle reductions).
 * `dtype` is the activation dtype (x, y, dy, dx); statistics and parameters are fp32.
 */

/* nn.LayerNorm over the last dim.  cols in {128,256,512,1024}.  mean/rstd: [rows] (saved for bwd;
 * may be NULL).  Replaces PMMA pre-norms and encoder_norm (model/PMMA/block.py:23-27,
 * model/PMMA/encoder.py:31,55; eps 1e-6) and v/x_gca_norm (model/basic_model.py:115,118). */
int dl_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                     float* rstd, int64_t rows, int32_t cols, float eps, int32_t dtype,
                     void* stream);
/* dgamma/dbeta ([cols] fp32; may be NULL): overwritten, or added to when accumulate != 0 (the
 * parameter's .grad buffer itself, torch's AccumulateGrad semantics without the extra add).
 * dx_add ([rows, cols], may be NULL; may alias dx): added to dx -- in a pre-norm residual block
 * (model/PMMA/block.py:33-47) the gradient of the skip connection joins the LayerNorm's input
 * gradient here instead of in a separate add pass. */
int dl_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                     const float* rstd, void* dx, const void* dx_add, float* dgamma, float* dbeta,
                     int64_t rows, int32_t cols, int32_t accumulate, int32_t dtype, void* stream);

/* softmax over the last dim of a [rows, cols<=1024] matrix with row stride ld; in-place allowed.
 * Replaces F.softmax in PGCA (model/PGCA/guided_cross_attention_model.py:308) and
 * Attention.softmax (model/PMMA/attention.py:42,60,71,112). */
int dl_softmax_fwd(const void* s, void* p, int64_t rows, int32_t cols, int64_t ld, int32_t dtype,
                   void* stream);
/* ds = scale * p * (dp - sum(p*dp)); in-place on dp allowed. */
int dl_softmax_bwd(const void* p, const void* dp, void* ds, int64_t rows, int32_t cols, int64_t ld,
                   float scale, int32_t dtype, void* stream);

/* out[c] (+)= sum_r x[r, c]  (bias gradients); out is fp32 [cols]: overwritten, or accumulated
 * into when accumulate != 0 (gradients written straight into the flat gradient buffer). */
int dl_colsum(const void* x, float* out, int64_t rows, int32_t cols, int64_t ld, int32_t accumulate,
              int32_t dtype, void* stream);

/* y = x * keep/(1-p), keep(i) = hash(seed, i) >= p (nn.Dropout of PMMA, model/PMMA/mlp.py:47,49,
 * model/PMMA/embed.py:42,52; the mask is recomputed from the seed in backward).  seed_step (device
 * pointer or NULL) advances the seed per training step: see dl_gemm_args.drop_seed_step. */
int dl_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, const int64_t* seed_step, int32_t dtype,
               void* stream);
/* y = act(x) elementwise (the ReLU inside Mean2Embed, model/cross_modality.py:166-171). */
int dl_act_fwd(const void* x, void* y, int64_t n, int32_t act, int32_t dtype, void* stream);
/* F.normalize(x, dim=-1) = l2norm (utils.py:443-444): y = x / max(||x||, eps); norm: [rows] fp32. */
int dl_l2norm_fwd(const void* x, void* y, float* norm, int64_t rows, int32_t cols, float eps,
                  int32_t dtype, void* stream);
int dl_l2norm_bwd(const void* dy, const void* y, const float* norm, void* dx, int64_t rows,
                  int32_t cols, int32_t dtype, void* stream);
/* g = dy * act'(pre) * dropout_mask(seed): backward of the dl_gemm epilogue's act + dropout. */
int dl_act_bwd(const void* dy, const void* pre, void* g, int64_t n, int32_t act, float p,
               uint64_t seed, const int64_t* seed_step, int32_t dtype, void* stream);
/* dst[r, 0:row_bytes) = src[r, 0:row_bytes), r < rows, rows src_pitch / dst_pitch bytes apart (all multiples
 * of 16): the column blocks of torch.cat(..., dim=-1) (model/DrugLAMP.py:57,66, model/PMMA/encoder.py:46) and
 * of its gradient, without PyTorch's element-wise strided copy. */
int dl_copy_rows(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t row_bytes,
                 int64_t rows, void* stream);

/* One AdamW step over a flat fp32 parameter buffer (torch.optim.AdamW semantics; the reference
 * builds AdamW in main.py:158-160).  grad is multiplied by grad_scale first (1/world_size after a
 * sum all-reduce); *step (device int64) is incremented; shadow_bf16 (optional) receives the
 * updated parameters in bf16 for the tensor-core GEMMs.  active_blocks (or NULL = all): one byte per
 * 64 consecutive elements; a 0 skips those elements entirely -- parameters that received no gradient
 * this step are neither decayed nor have their moments moved, as torch.optim.AdamW skips
 * `p.grad is None` (three AdamWs over the same parameters, main.py:158-160; SSL / CM heads get no
 * gradient from the classification loss, trainer.py:196-200).  step_blocks (or NULL): one int32 step
 * count per 64 elements, advanced for the active blocks and used for THEIR bias correction -- torch keeps
 * state['step'] per parameter, which matters when the set of parameters with a gradient changes from
 * step to step (classification / SSL / 2C2P epochs of trainer.py:190-191); NULL = *step for everyone.
 * tick != 0: *step is incremented first (one call per optimiser step does this); tick == 0: *step is left
 * alone and the bias correction uses *step + 1 -- a RANGE of the flat buffer updated ahead of the call that
 * ticks (train.TrainStep updates the PMMA parameters while the rest of the backward still runs; the dropout
 * kernels of that backward keep reading the un-ticked counter). */
int dl_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  void* shadow_bf16, int64_t n, int64_t* step, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float grad_scale, const uint8_t* active_blocks,
                  int32_t* step_blocks, int32_t tick, void* stream);
int dl_cast(const void* x, int32_t dtype_in, void* y, int32_t dtype_out, int64_t n, void* stream);
/* y[i] = dropout(x[i] + pe[i % period])  (model/PMMA/embed.py:51-52). */
int dl_add_pe(const void* x, const float* pe, void* y, int64_t n, int64_t period, float p,
              uint64_t seed, const int64_t* seed_step, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Molecular GCN (model/basic_model.py:545-638 GraphConv norm='both'; :411-436 GCNLayer).
 */

/* out[i,:] = norm_dst[i] * sum_{e in indptr[i]..indptr[i+1]} norm_src[indices[e]] * h[indices[e],:]
 * CSR by destination = DGL update_all(copy_u, sum) with both degree norms folded in
 * (model/basic_model.py:596-603,612-618,623-630).  Backward = same call on the transposed CSR
 * with the norms swapped.  feats must be 128. */
int dl_spmm_norm(const int32_t* indptr, const int32_t* indices, const float* norm_src,
                 const float* norm_dst, const void* h, void* out, int64_t n_rows, int32_t feats,
                 int32_t dtype, void* stream);

/* Graph carrier built on the device from the batched molecule's edge list (what DGL hands
 * GraphConv: g.edges(), in_degrees(), out_degrees(); model/basic_model.py:579-603,623-630):
 * CSR by destination (indptr / indices = source ids) and by source (indptr_t / indices_t) in stable
 * edge order -- identical to a host argsort(stable) -- plus norm_dst = clamp(in_deg, 1)^-1/2 and
 * norm_src = clamp(out_deg, 1)^-1/2.  Duplicate edges are kept (App. A5).  flags[0] = number of
 * zero-in-degree nodes (the DGLError condition of :580-590), flags[1] = edges with an endpoint outside
 * [0, n_nodes) (skipped).  indptr*: [n_nodes + 1], indices*: [n_edges], workspace: 2*n_nodes +
 * 2*n_edges int32.  No host synchronisation: the input pipeline can run it on its copy stream. */
int dl_csr_build(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n_nodes,
                 int32_t* indptr, int32_t* indices, int32_t* indptr_t, int32_t* indices_t,
                 float* norm_src, float* norm_dst, int32_t* flags, int32_t* workspace, void* stream);

/* nn.BatchNorm1d over [rows, cols] (model/basic_model.py:401,434 bn_layer; also
 * cross_modality.py:168).  training!=0: batch statistics, running buffers updated with
 * `momentum` (unbiased variance) and *num_batches_tracked += 1 when given; training==0: running
 * statistics.  mean/rstd: [cols] outputs saved for backward.  workspace: 2*cols doubles.
 * y == NULL: statistics only (mean / rstd and the running buffers); the normalisation is then applied by
 * the consumer (dl_bn_transpose).
 * last_row_weight w > 1 (training only): the LAST row stands for w identical rows -- the molecular
 * GCN's virtual nodes, 92 % of the 512 slots per molecule, all carry one and the same row at every
 * layer, so the layer is evaluated once for them: batch statistics count that row w
 * times (N = rows - 1 + w).  In the backward the last row's dy is the SUM of the w copies' gradients
 * and its dx the sum of their input gradients.  w <= 1 turns it off. */
int dl_batchnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                     float* rstd, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, double* workspace, int64_t rows, int32_t cols,
                     float eps, float momentum, int32_t training, float last_row_weight, int32_t dtype,
                     void* stream);
/* accumulate != 0: dgamma/dbeta are added to (the parameters' .grad buffers).  relu_mask != 0: x is
 * a ReLU output (ProteinCNN conv -> ReLU -> BN, model/basic_model.py:174-178) and dx is also
 * multiplied by (x > 0), so the activation's backward needs no pass of its own. */
int dl_batchnorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                     const float* rstd, void* dx, float* dgamma, float* dbeta, double* workspace,
                     int64_t rows, int32_t cols, int32_t training, int32_t accumulate,
                     int32_t relu_mask, float last_row_weight, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Model glue.
 */

/* x: (B, S*L, C) fp32.  bit_out (B, S*L) = float(x.sum(-1) == 0)            [may be NULL]
 *                       cat_out (B, S*L, C+1) = cat(x, bit) fp32           [may be NULL]
 *                       pooled  (B, L, C+1) = cat(x, bit).view(B,S,L,C+1).mean(1)  [may be NULL]
 * One pass over x (model/DrugLAMP.py:11-19,39-40).  S = 1 gives the plain fill-bit concat.
 * pooled rows have leading dimension ld_pooled (0 = C+1); columns C+1..ld_pooled-1 are written
 * as zeros, so a 641- or 385-wide result can feed dl_gemm without a padding copy. */
int dl_fillbit_pool(const float* x, float* bit_out, float* cat_out, void* pooled,
                    int32_t pooled_dtype, int64_t B, int32_t S, int32_t L, int32_t C,
                    int32_t ld_pooled, void* stream);

/* Device side of the collate (utils.py:304-324, called from multimodality_collate_func :326-334):
 * rows [sum_b R_b, C] fp32 holds every sample's embedding rows back to back, offsets [B+1] int32
 * their starts.  out (B, maxsize, C) fp32 is written exactly as the reference's host loops do:
 * repeat = 0: tail_pad   -- rows once, zeros after (R_b > maxsize is truncated);
 * repeat = 1: repeat_pad -- rows tiled floor(maxsize / R_b) times, zeros after.
 * Lets the host ship 1.5 MB/pair instead of the 6.9 MB/pair dense tensors over PCIe. */
int dl_expand_rows(const float* rows, const int32_t* offsets, float* out, int64_t B, int32_t maxsize,
                   int32_t C, int32_t repeat, void* stream);

/* y[b, j, :] = mean_s x[b, s*L + j, :]; y rows have stride ldy (model/DrugLAMP.py:35-37). */
int dl_site_pool_fwd(const void* x, void* y, int64_t B, int32_t S, int32_t L, int32_t C,
                     int64_t ldy, int32_t dtype, void* stream);
int dl_site_pool_bwd(const void* dy, void* dx, int64_t B, int32_t S, int32_t L, int32_t C,
                     int64_t ldy, int32_t dtype, void* stream);

/* y[b, c, r] = x[b, r, c]: the channels-last ProteinCNN output laid out as the reference's
 * (B, C, L) buffer, which the reference then reinterprets with .view(B, L, C)
 * (model/basic_model.py:179). */
int dl_transpose(const void* x, void* y, int64_t B, int32_t R, int32_t C, int32_t dtype, void* stream);

/* The same layout change with ProteinCNN's last BatchNorm1d (model/basic_model.py:178) applied on the way:
 * y[b, c, r] = (x[b, r, c] - mean[c]) * rstd[c] * gamma[c] + beta[c]  (mean == NULL: plain transpose;
 * gamma / beta may be NULL).  mean / rstd come from dl_batchnorm_fwd called with y == NULL (statistics
 * only).  64 x 64 tiles, 16-byte accesses both ways: R and C multiples of 8 (bf16) / 4 (fp32) elements. */
int dl_bn_transpose(const void* x, void* y, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, int64_t B, int32_t R, int32_t C, int32_t dtype, void* stream);

/* Gradient of the channels-last activation x (B, L, C) through  transpose -> (B, C, L) ->
 * .view(B, L, C) -> .view(B, S, L/S, C).mean(1)  (model/basic_model.py:179, model/DrugLAMP.py:35-37) in one
 * pass: dx[b, l, c] = g[b, ((c*L + l) / C) % (L/S), (c*L + l) % C] / S with g (B, L/S, C) contiguous.
 * Replaces dl_site_pool_bwd + dl_transpose (two passes over the 9x larger tensor). */
int dl_site_pool_view_bwd(const void* g, void* dx, int64_t B, int32_t S, int32_t L, int32_t C, int32_t dtype,
                          void* stream);

/* ProteinCNN input (model/basic_model.py:171-173: nn.Embedding(27, 127, padding_idx=0), then
 * torch.cat with the fill mask): out[r, 0:127] = table[tokens[r], :], out[r, 127] = fill[r], written
 * in the compute dtype.  tokens: [rows] int64 (DL_TOK_I64) or the float64 the reference collate
 * delivers (DL_TOK_F64, utils.py:403-407), clamped to [0, vocab).  width must be 128. */
#define DL_TOK_I64 0
#define DL_TOK_F64 1
int dl_embed_fill_fwd(const void* tokens, int32_t tok_dtype, const float* fill, const float* table,
                      void* out, int64_t rows, int32_t vocab, int32_t width, int32_t dtype,
                      void* stream);
/* dtable[v, c] += sum_{r: tokens[r] == v} g[r, c] for c < 127; the padding_idx row (pass -1 for
 * none) receives nothing, like nn.Embedding's backward.  dtable: [vocab, 127] fp32, accumulated. */
int dl_embed_fill_bwd(const void* tokens, int32_t tok_dtype, const void* g, float* dtable,
                      int64_t rows, int32_t vocab, int32_t width, int32_t padding_idx, int32_t dtype,
                      void* stream);

/* y = LayerNorm(v + gate(v)), gate = MultiHeadLinearAttention's softmax-over-sequence gating
 * through its .view(B*H, L, E/H) reinterpretation (model/PMMA/encoder.py:132-140) applied to
 * logits = lin2(act(lin1(v))) of shape (B, L, H); residual and LayerNorm from
 * model/DrugLAMP.py:63-71.  p_out (B,H,L), mean/rstd (B*L) are saved for backward.
 * gamma == NULL selects the gating alone, y = gate(v) (MultiHeadLinearAttention.forward as the
 * reference module returns it; beta/mean/rstd unused). */
int dl_mhla_gate_ln_fwd(const void* v, const void* logits, const float* gamma, const float* beta,
                        void* y, float* p_out, float* mean, float* rstd, int64_t B, int32_t L,
                        int32_t E, int32_t H, float eps, int32_t dtype, void* stream);
/* dv = gradient through the gating/residual path only (add the lin1/lin2 path to it). */
int dl_mhla_gate_ln_bwd(const void* dy, const void* v, const float* p, const float* mean,
                        const float* rstd, const float* gamma, void* dv, void* dlogits,
                        float* dgamma, float* dbeta, int64_t B, int32_t L, int32_t E, int32_t H,
                        int32_t dtype, void* stream);

/* CrossModality triplet loss on cos = P_lat . D_lat^T (unit rows) and labels G in {0,1}
 * (model/cross_modality.py:15-47 with utils.py:571-574 distance).  acc: 2 doubles workspace
 * (hinge sum, term count) kept for backward; loss: 1 float. */
int dl_cm_triplet_fwd(const float* cos, const int8_t* G, int64_t P, int64_t D, float margin,
                      double* acc, float* loss, void* stream);
int dl_cm_triplet_bwd(const float* cos, const int8_t* G, int64_t P, int64_t D, float margin,
                      const double* acc, const float* gout, float* dcos, void* stream);

/* F.cross_entropy(logits, labels, ignore_index) with mean reduction over the non-ignored rows --
 * the two MLM heads of SSL.prot_mlm (model/self_supervised_learning.py:78-101).
 * logit[r,c] = x[r*ld + c] + extra[r] * wextra[c] (extra/wextra optional: the fill-bit column of
 * llm_to_logits applied without materialising cat(xp, bit)).  acc: 2 doubles workspace. */
int dl_cross_entropy_fwd(const void* x, const int64_t* labels, const float* extra,
                         const float* wextra, int64_t rows, int32_t C, int64_t ld,
                         int64_t ignore_index, double* acc, float* loss, int32_t dtype, void* stream);
int dl_cross_entropy_bwd(const void* x, const int64_t* labels, const float* extra,
                         const float* wextra, int64_t rows, int32_t C, int64_t ld,
                         int64_t ignore_index, const double* acc, const float* gout, void* dx,
                         float* dextra, int32_t dtype, void* stream);

/* binary_cross_entropy: prob = sigmoid(score); loss = BCELoss(prob, y) mean
 * (model/basic_model.py:17-22). */
int dl_bce_fwd(const float* score, const float* y, float* prob, float* loss, int64_t n,
               void* stream);
int dl_bce_bwd(const float* prob, const float* y, const float* gout, float* dscore, int64_t n,
               void* stream);

/* One dense layer on at most 64 rows, fp32, in ONE launch: Y = BatchNorm1d(act(X op(W) + bias)) -- a layer
 * of the decoder head (model/basic_model.py:196-215: fc -> GELU -> BatchNorm1d over the PAIRS of the
 * batch; the head is 0.04 % of the FLOPs but a serial chain between forward and backward).
 *   w_kn = 0: W is [N, K] (nn.Linear weight, y = x W^T);  w_kn = 1: W is [K, N] (dX = g W)
 *   pre (optional): X op(W) + bias before the activation, kept for the backward
 *   bn != 0: BatchNorm1d over the M rows -- training != 0: batch statistics, running_mean / running_var
 *   (optional) updated with `momentum` (unbiased variance), *num_batches_tracked += 1 (optional);
 *   training == 0: the running statistics.  mean / rstd: [N] outputs for the backward.
 * CUDA-core fp32 FMA (the product is latency bound at M <= 64); a CTA owns 8 output columns for all rows,
 * which makes the batch statistics CTA-local. */
typedef struct dl_small_linear_args {
  const float* X; const float* W; const float* bias;
  float* pre; float* Y;
  const float* gamma; const float* beta;
  float* mean; float* rstd; float* running_mean; float* running_var; int64_t* num_batches_tracked;
  int64_t M, N, K;
  int64_t ldx, ldw, ldy;          /* row strides of X, W and of Y / pre (elements) */
  int32_t w_kn, act, bn, training;
  float eps, momentum;
} dl_small_linear_args;
int dl_small_linear(const dl_small_linear_args* args, void* stream);
/* Backward of the layer's BatchNorm1d + activation (column-local): g = d loss / d pre from dy = d loss / d Y;
 * dgamma, dbeta, dbias (each optional): parameter gradients, added to when accumulate != 0.  dX and dW
 * follow as g W (dl_small_linear, w_kn = 1) and g^T X (dl_gemm).  Contiguous [M, N] tensors, M <= 64. */
int dl_head_bn_act_bwd(const float* dy, const float* pre, const float* gamma, const float* mean,
                       const float* rstd, float* g, float* dgamma, float* dbeta, float* dbias,
                       int64_t M, int64_t N, int32_t act, int32_t bn, int32_t training,
                       int32_t accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRUGLAMP_SM100_H_ */
