// Small-M dense layers: the decoder head (model/basic_model.py:196-215: three Linear -> GELU ->
// BatchNorm1d over the PAIRS of the batch, then Linear -> 1) works on <= 64 rows.  On the tensor-core
// GEMM such a layer is a chain of launch-latency-bound kernels (a 64 x 1024 x 1024 product occupies 16
// CTAs for 32 serial K-blocks: 21 us; then memset, column statistics, normalise), and the whole head
// sits on the critical path between the forward and the backward with the GPU idle.  Here one layer
// is ONE launch forward (product + bias + GELU + BatchNorm statistics / running buffers / normalise)
// and two backward (BatchNorm + GELU backward with the bias / gamma / beta gradients; dX).
//
// fp32 throughout, CUDA-core FMA: M <= 64 rows make the product HBM/L2-latency bound (the weights are
// read once, 4 MB), not FLOP bound.  A CTA owns 8 output columns for all rows -- so BatchNorm's batch
// statistics are CTA-local -- and streams K through shared memory with a 4-stage cp.async pipeline;
// its 8 warps split every K tile, lanes own rows (l, l + 32), partials meet in shared memory.
#include "../../include/druglamp_sm100.h"
#include "common.cuh"

namespace dl {
void count_launch(int n = 1);

namespace {

constexpr int kRows = 64;        // max M
constexpr int kNC = 8;           // output columns per CTA
constexpr int kKT = 128;         // K tile
constexpr int kXS = kKT + 4;     // smem row stride of the X tile (floats): 16-byte reads of 8 lanes hit 8 banks groups
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kStageFloats = kRows * kXS + kNC * kXS;     // X tile + W tile (NT: [8][kXS]; NN: [kKT][8] fits too)
constexpr int kSmemBytes = kStages * kStageFloats * 4;
static_assert(kKT * kNC <= kNC * kXS, "NN weight tile must fit the NT slot");
static_assert(8 * kRows * kNC * 4 <= kSmemBytes, "partial sums alias the pipeline buffers");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

struct HeadArgs {
  const float* X;      // [M, K]
  const float* W;      // NT: [N, K] (ldw = K stride); NN: [K, N]
  const float* bias;   // [N] or null
  float* pre;          // [M, N] value after the bias, before the activation (null: not kept)
  float* Y;            // [M, N]
  const float* gamma;  // BatchNorm (bn != 0)
  const float* beta;
  float* mean;         // [N] saved statistics (training) / the running ones used (eval)
  float* rstd;
  float* running_mean;
  float* running_var;
  long long* nbt;
  int M, N, K;
  long long ldx, ldw, ldy;
  int act, bn, training;
  float eps, momentum;
};

// Y = epilogue(X op(W)).  NN = false: W[n, k] (y = x W^T);  NN = true: W[k, n] (dx = g W).
template <bool NN>
__global__ void __launch_bounds__(kThreads, 1) small_linear_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float smem[];
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * kNC;
  const int ntiles = (a.K + kKT - 1) / kKT;
  const bool x_vec = (a.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.X) & 15) == 0);
  const bool w_vec = (a.ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
  pdl_wait();

  auto load_tile = [&](int t) {
    float* xs = smem + (t % kStages) * kStageFloats;
    float* ws = xs + kRows * kXS;
    const int k0 = t * kKT;
    // X tile: 64 rows x 32 float4; rows >= M and columns >= K are zero
#pragma unroll
    for (int i = 0; i < (kRows * kKT / 4) / kThreads; ++i) {
      const int slot = tid + i * kThreads;
      const int r = slot >> 5, kq = (slot & 31) * 4;
      float* dst = xs + r * kXS + kq;
      const int k = k0 + kq;
      if (r < a.M && k + 3 < a.K && x_vec) {
        cp_async16(dst, a.X + (long long)r * a.ldx + k);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = (r < a.M && k + j < a.K) ? a.X[(long long)r * a.ldx + k + j] : 0.f;
      }
    }
    if constexpr (!NN) {
      // W tile [8 columns][kKT]: 8 x 32 float4 = one per thread
      const int c = tid >> 5, kq = (tid & 31) * 4;
      float* dst = ws + c * kXS + kq;
      const int n = n0 + c, k = k0 + kq;
      if (n < a.N && k + 3 < a.K && w_vec) {
        cp_async16(dst, a.W + (long long)n * a.ldw + k);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = (n < a.N && k + j < a.K) ? a.W[(long long)n * a.ldw + k + j] : 0.f;
      }
    } else {
      // W tile [kKT][8 columns]: 128 x 2 float4 = one per thread
      const int kk = tid >> 1, cq = (tid & 1) * 4;
      float* dst = ws + kk * kNC + cq;
      const int k = k0 + kk, n = n0 + cq;
      if (k < a.K && n + 3 < a.N && w_vec) {
        cp_async16(dst, a.W + (long long)k * a.ldw + n);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = (k < a.K && n + j < a.N) ? a.W[(long long)k * a.ldw + n + j] : 0.f;
      }
    }
  };

  float acc[2][kNC];
#pragma unroll
  for (int c = 0; c < kNC; ++c) acc[0][c] = acc[1][c] = 0.f;

  for (int t = 0; t < kStages - 1; ++t) {
    if (t < ntiles) load_tile(t);
    cp_commit();
  }
  for (int t = 0; t < ntiles; ++t) {
    cp_wait<kStages - 2>();
    __syncthreads();                               // tile t landed for everyone; slot (t-1) % S is free
    if (t + kStages - 1 < ntiles) load_tile(t + kStages - 1);
    cp_commit();
    const float* xs = smem + (t % kStages) * kStageFloats;
    const float* ws = xs + kRows * kXS;
    // warp w: k in [16 w, 16 w + 16) of the tile, four at a time
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int k = warp * 16 + g * 4;
      const float4 xa = *reinterpret_cast<const float4*>(xs + lane * kXS + k);
      const float4 xb = *reinterpret_cast<const float4*>(xs + (lane + 32) * kXS + k);
      if constexpr (!NN) {
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(ws + c * kXS + k);
          acc[0][c] = fmaf(xa.x, w.x, fmaf(xa.y, w.y, fmaf(xa.z, w.z, fmaf(xa.w, w.w, acc[0][c]))));
          acc[1][c] = fmaf(xb.x, w.x, fmaf(xb.y, w.y, fmaf(xb.z, w.z, fmaf(xb.w, w.w, acc[1][c]))));
        }
      } else {
        const float xav[4] = {xa.x, xa.y, xa.z, xa.w}, xbv[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + (k + j) * kNC);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + (k + j) * kNC + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int c = 0; c < kNC; ++c) {
            acc[0][c] = fmaf(xav[j], wv[c], acc[0][c]);
            acc[1][c] = fmaf(xbv[j], wv[c], acc[1][c]);
          }
        }
      }
    }
  }
  cp_wait<0>();
  __syncthreads();
  // partial sums of the 8 warps -> shared memory [warp][row][col] (aliases the pipeline buffers)
  float* red = smem;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float* dst = red + (warp * kRows + lane + 32 * h) * kNC;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[h][4], acc[h][5], acc[h][6], acc[h][7]);
  }
  __syncthreads();
  // thread -> (row = tid & 63, columns 2 (tid >> 6), +1)
  const int row = tid & 63, c0 = (tid >> 6) * 2;
  float v[2] = {0.f, 0.f};
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const float2 p = *reinterpret_cast<const float2*>(red + (w * kRows + row) * kNC + c0);
    v[0] += p.x; v[1] += p.y;
  }
  const bool row_ok = row < a.M;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int n = n0 + c0 + j;
    if (n < a.N) {
      if (a.bias) v[j] += a.bias[n];
      if (a.pre && row_ok) a.pre[(long long)row * a.ldy + n] = v[j];
      if (a.act == DL_ACT_GELU) v[j] = gelu_erf(v[j]);
      else if (a.act == DL_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
    }
  }
  if (a.bn) {
    // column statistics over the M rows: the activations of this CTA's 8 columns meet in shared
    // memory, warp c reduces column c (two-pass variance: 64 values, no cancellation)
    __syncthreads();
    float* av = smem + 8 * kRows * kNC;            // [row][col], behind the partials
    float* st = av + kRows * kNC;                  // [col][2]: mean, rstd
    av[row * kNC + c0] = row_ok ? v[0] : 0.f;
    av[row * kNC + c0 + 1] = row_ok ? v[1] : 0.f;
    __syncthreads();
    if (warp < kNC) {
      const int n = n0 + warp;
      float mean, rstd;
      if (a.training) {
        const float x0 = av[lane * kNC + warp], x1 = av[(lane + 32) * kNC + warp];
        float s = x0 + x1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        mean = s / (float)a.M;
        const float d0 = lane < a.M ? x0 - mean : 0.f, d1 = lane + 32 < a.M ? x1 - mean : 0.f;
        float q = d0 * d0 + d1 * d1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float var = q / (float)a.M;
        rstd = rsqrtf(var + a.eps);
        if (lane == 0 && n < a.N && a.running_mean) {
          a.running_mean[n] = (1.f - a.momentum) * a.running_mean[n] + a.momentum * mean;
          const float unb = a.M > 1 ? q / (float)(a.M - 1) : var;
          a.running_var[n] = (1.f - a.momentum) * a.running_var[n] + a.momentum * unb;
        }
      } else {
        mean = n < a.N ? a.running_mean[n] : 0.f;
        rstd = n < a.N ? rsqrtf(a.running_var[n] + a.eps) : 0.f;
      }
      if (lane == 0) {
        st[2 * warp] = mean; st[2 * warp + 1] = rstd;
        if (n < a.N) { a.mean[n] = mean; a.rstd[n] = rstd; }
      }
    }
    if (blockIdx.x == 0 && tid == 0 && a.training && a.nbt) *a.nbt += 1;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + c0 + j;
      if (n < a.N) {
        const float g = a.gamma ? a.gamma[n] : 1.f, b = a.beta ? a.beta[n] : 0.f;
        v[j] = (v[j] - st[2 * (c0 + j)]) * st[2 * (c0 + j) + 1] * g + b;
      }
    }
  }
  if (row_ok) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + c0 + j;
      if (n < a.N) a.Y[(long long)row * a.ldy + n] = v[j];
    }
  }
}

// BatchNorm1d + activation backward of one head layer, column-local: block = 32 columns (lane =
// column, coalesced rows), warp w owns rows w, w + 8, ...
//   a = act(pre), xh = (a - mean) rstd;  training: da = gamma rstd (dy - mean(dy) - xh mean(dy xh)),
//   eval: da = gamma rstd dy;  g = da act'(pre);  dgamma += sum dy xh, dbeta += sum dy, dbias += sum g.
__global__ void __launch_bounds__(256) head_bn_act_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ pre, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ g,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int M, int N,
    int act, int bn, int training, int accumulate) {
  __shared__ float red[3][8][32];
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool on = n < N;
  pdl_wait();
  float mu = 0.f, rs = 1.f, gm = 1.f;
  if (on && bn) { mu = mean[n]; rs = rstd[n]; gm = gamma ? gamma[n] : 1.f; }
  float s1 = 0.f, s2 = 0.f;
  float dyv[8], av[8], dav[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp + 8 * i;
    dyv[i] = 0.f; av[i] = 0.f; dav[i] = 0.f;
    if (on && r < M) {
      const float p = pre[(long long)r * N + n];
      dyv[i] = dy[(long long)r * N + n];
      if (act == DL_ACT_GELU) { av[i] = gelu_erf(p); dav[i] = gelu_erf_grad(p); }
      else if (act == DL_ACT_RELU) { av[i] = fmaxf(p, 0.f); dav[i] = p > 0.f ? 1.f : 0.f; }
      else { av[i] = p; dav[i] = 1.f; }
      const float xh = (av[i] - mu) * rs;
      av[i] = xh;
      s1 += dyv[i];
      s2 += dyv[i] * xh;
    }
  }
  red[0][warp][lane] = s1; red[1][warp][lane] = s2;
  __syncthreads();
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { t1 += red[0][w][lane]; t2 += red[1][w][lane]; }
  const float m1 = t1 / (float)M, m2 = t2 / (float)M;
  float sg = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp + 8 * i;
    if (on && r < M) {
      float da = dyv[i];
      if (bn) da = training ? gm * rs * (dyv[i] - m1 - av[i] * m2) : gm * rs * dyv[i];
      const float gv = da * dav[i];
      g[(long long)r * N + n] = gv;
      sg += gv;
    }
  }
  red[2][warp][lane] = sg;
  __syncthreads();
  if (warp == 0 && on) {
    float tg = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tg += red[2][w][lane];
    if (dbias) dbias[n] = accumulate ? dbias[n] + tg : tg;
    if (bn && dgamma) dgamma[n] = accumulate ? dgamma[n] + t2 : t2;
    if (bn && dbeta) dbeta[n] = accumulate ? dbeta[n] + t1 : t1;
  }
}

}  // namespace
}  // namespace dl

using namespace dl;

extern "C" int dl_small_linear(const dl_small_linear_args* p, void* stream) {
  DL_REQUIRE(p != nullptr, "dl_small_linear: null args");
  DL_REQUIRE(p->X && p->W && p->Y, "dl_small_linear: X, W and Y must be non-null");
  DL_REQUIRE(p->M >= 1 && p->M <= kRows, "dl_small_linear: M must be in [1, %d] (got %lld)", kRows, (long long)p->M);
  DL_REQUIRE(p->N >= 1 && p->K >= 1 && p->N < (1ll << 31) && p->K < (1ll << 31), "dl_small_linear: bad N / K");
  DL_REQUIRE(p->act >= 0 && p->act <= 2, "dl_small_linear: bad act");
  if (p->bn) {
    DL_REQUIRE(p->mean && p->rstd, "dl_small_linear: BatchNorm needs mean / rstd outputs");
    DL_REQUIRE(p->training || (p->running_mean && p->running_var), "dl_small_linear: eval-mode BatchNorm needs running statistics");
    DL_REQUIRE(!p->running_mean == !p->running_var, "dl_small_linear: running_mean and running_var go together");
  }
  HeadArgs a;
  a.X = p->X; a.W = p->W; a.bias = p->bias; a.pre = p->pre; a.Y = p->Y;
  a.gamma = p->gamma; a.beta = p->beta; a.mean = p->mean; a.rstd = p->rstd;
  a.running_mean = p->running_mean; a.running_var = p->running_var; a.nbt = (long long*)p->num_batches_tracked;
  a.M = (int)p->M; a.N = (int)p->N; a.K = (int)p->K;
  a.ldx = p->ldx; a.ldw = p->ldw; a.ldy = p->ldy;
  a.act = p->act; a.bn = p->bn; a.training = p->training; a.eps = p->eps; a.momentum = p->momentum;
  cudaStream_t st = (cudaStream_t)stream;
  static cudaError_t attr_err = [] {
    cudaError_t e = cudaFuncSetAttribute(small_linear_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(small_linear_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    return e;
  }();
  if (attr_err != cudaSuccess)
    return set_error((int)attr_err, "dl_small_linear: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  const int grid = (a.N + kNC - 1) / kNC;
  if (p->w_kn) DL_LAUNCH((small_linear_kernel<true>), grid, kThreads, kSmemBytes, st, a);
  else DL_LAUNCH((small_linear_kernel<false>), grid, kThreads, kSmemBytes, st, a);
  DL_LAUNCH_CHECK("small_linear_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_head_bn_act_bwd(const float* dy, const float* pre, const float* gamma, const float* mean,
                                  const float* rstd, float* g, float* dgamma, float* dbeta, float* dbias,
                                  int64_t M, int64_t N, int32_t act, int32_t bn, int32_t training,
                                  int32_t accumulate, void* stream) {
  DL_REQUIRE(dy && pre && g, "dl_head_bn_act_bwd: null pointer");
  DL_REQUIRE(M >= 1 && M <= kRows && N >= 1 && N < (1ll << 31), "dl_head_bn_act_bwd: M must be in [1, %d]", kRows);
  DL_REQUIRE(!bn || (mean && rstd), "dl_head_bn_act_bwd: BatchNorm needs the saved mean / rstd");
  cudaStream_t st = (cudaStream_t)stream;
  DL_LAUNCH(head_bn_act_bwd_kernel, (int)((N + 31) / 32), 256, 0, st, dy, pre, gamma, mean, rstd, g, dgamma, dbeta,
            dbias, (int)M, (int)N, act, bn, training, accumulate);
  DL_LAUNCH_CHECK("head_bn_act_bwd_kernel");
  count_launch();
  return 0;
}
