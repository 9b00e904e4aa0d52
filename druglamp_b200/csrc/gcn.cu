// Molecular GCN kernels (reference model/basic_model.py:545-638 GraphConv, :411-436 GCNLayer):
//   * dl_spmm_norm   -- degree-normalised segment-sum over in-edges (DGL update_all(copy_u,sum)
//                       with D_out^-1/2 on the source side and D_in^-1/2 on the destination side
//                       folded in).  CSR by destination, one warp per destination row, each lane
//                       moves a 128-bit vector (32 lanes x 4 = 128 features).  The same kernel
//                       with the transposed CSR and swapped norms is the backward.
//   * BatchNorm1d over all N node rows: column statistics in fp64 partials (one atomic per column
//     per block), a finalize step that also updates the running buffers, and fused apply /
//     backward kernels.
// All of it is HBM-bound integer-indexed work: no tensor cores here.
#include "../../include/druglamp_sm100.h"
#include "common.cuh"

namespace dl {
void count_launch(int n = 1);
namespace {

// out[i, :] = nd[i] * sum_{e in [indptr[i], indptr[i+1])} ns[idx[e]] * h[idx[e], :]
template <typename T>
__global__ void __launch_bounds__(256)
spmm_norm_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                 const float* __restrict__ norm_src, const float* __restrict__ norm_dst,
                 const T* __restrict__ h, T* __restrict__ out, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int e0 = indptr[row], e1 = indptr[row + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = e0; e < e1; ++e) {
    const int s = indices[e];           // warp-uniform broadcast load
    const float w = norm_src[s];
    const float4 v = ld4<T>(h + (size_t)s * 128 + lane * 4);
    acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
  }
  const float nd = norm_dst[row];
  acc.x *= nd; acc.y *= nd; acc.z *= nd; acc.w *= nd;
  st4<T>(out + (size_t)row * 128 + lane * 4, acc);
}

// ---------------------------------------------------------------- BatchNorm statistics
// sums[0:C] += sum_r a[r,c] ; sums[C:2C] += sum_r a[r,c]*b[r,c]
//   forward : a = x, b = x            -> sum, sum of squares
//   backward: a = dy, b = xhat (from x, mean, rstd) -> sum dy, sum dy*xhat
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
bn_colstats_kernel(const T* __restrict__ a, const T* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ rstd, double* __restrict__ sums, long long rows,
                   int cols, long long rows_per_block) {
  __shared__ double red[2][8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  double s1 = 0.0, s2 = 0.0;
  if (c < cols) {
    const float mu = BWD ? mean[c] : 0.f, rs = BWD ? rstd[c] : 0.f;
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      const float av = ldf<T>(a, r * cols + c);
      const float bv = BWD ? (ldf<T>(x, r * cols + c) - mu) * rs : av;
      s1 += av;
      s2 += (double)av * bv;
    }
  }
  red[0][threadIdx.y][threadIdx.x] = s1;
  red[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t1 += red[0][i][threadIdx.x]; t2 += red[1][i][threadIdx.x]; }
    atomicAdd(sums + c, t1);
    atomicAdd(sums + cols + c, t2);
  }
}

// mean / biased var -> rstd; running = (1-m)*running + m*{mean, unbiased var}
__global__ void bn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ nbt,
                                   long long rows, int cols, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) {
    const double mu = sums[c] / (double)rows;
    double var = sums[cols + c] / (double)rows - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  }
  if (c == 0 && nbt) *nbt += 1;
}

// y = (x - mean) * rstd * gamma + beta      (eval: mean/rstd derived from running stats)
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, T* __restrict__ y, long long n4,
                                int cols) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % cols);
    const float4 v = ld4<T>(x + i * 4);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gamma) g = *reinterpret_cast<const float4*>(gamma + c);
    if (beta) b = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = (v.x - mu.x) * rs.x * g.x + b.x;
    o.y = (v.y - mu.y) * rs.y * g.y + b.y;
    o.z = (v.z - mu.z) * rs.z * g.z + b.z;
    o.w = (v.w - mu.w) * rs.w * g.w + b.w;
    st4<T>(y + i * 4, o);
  }
}

// training: dx = gamma*rstd * (dy - sum_dy/N - xhat * sum_dyxhat/N); eval: dx = gamma*rstd*dy
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const double* __restrict__ sums,
                                    T* __restrict__ dx, long long n4, int cols, double inv_rows,
                                    int training) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % cols);
    const float4 dv = ld4<T>(dy + i * 4);
    const float4 xv = ld4<T>(x + i * 4);
    const float d[4] = {dv.x, dv.y, dv.z, dv.w};
    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = gamma ? gamma[c + k] : 1.f;
      const float rs = rstd[c + k];
      if (training) {
        const float xh = (xx[k] - mean[c + k]) * rs;
        const float m1 = (float)(sums[c + k] * inv_rows), m2 = (float)(sums[cols + c + k] * inv_rows);
        o[k] = g * rs * (d[k] - m1 - xh * m2);
      } else {
        o[k] = g * rs * d[k];
      }
    }
    st4<T>(dx + i * 4, make_float4(o[0], o[1], o[2], o[3]));
  }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) {
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dgamma) dgamma[c] = (float)sums[cols + c];
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var, float* __restrict__ mean,
                                     float* __restrict__ rstd, int cols, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) {
    mean[c] = running_mean[c];
    rstd[c] = rsqrtf(running_var[c] + eps);
  }
}

void stats_grid(long long rows, int cols, dim3* grid, long long* rpb) {
  const int xb = ceil_div(cols, 32);
  long long yb = (long long)sm_count() * 4 / xb;
  if (yb < 1) yb = 1;
  const long long maxy = (rows + 63) / 64;
  if (yb > maxy) yb = maxy < 1 ? 1 : maxy;
  *rpb = (rows + yb - 1) / yb;
  *grid = dim3(xb, (unsigned)yb);
}

int ew_grid4(long long n4) {
  long long b = (n4 + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace
}  // namespace dl

using namespace dl;

extern "C" int dl_spmm_norm(const int32_t* indptr, const int32_t* indices, const float* norm_src,
                            const float* norm_dst, const void* h, void* out, int64_t n_rows,
                            int32_t feats, int32_t dtype, void* stream) {
  DL_REQUIRE(indptr && indices && norm_src && norm_dst && h && out, "dl_spmm_norm: null pointer");
  DL_REQUIRE(feats == 128, "dl_spmm_norm: feature width must be 128 (got %d)", feats);
  DL_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "dl_spmm_norm: bad row count");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(n_rows, 8);
  if (dtype == DL_BF16)
    spmm_norm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(indptr, indices, norm_src, norm_dst, (const __nv_bfloat16*)h, (__nv_bfloat16*)out, (int)n_rows);
  else
    spmm_norm_kernel<float><<<grid, 256, 0, st>>>(indptr, indices, norm_src, norm_dst, (const float*)h, (float*)out, (int)n_rows);
  DL_LAUNCH_CHECK("spmm_norm_kernel");
  count_launch();
  return 0;
}

// workspace: 2*cols doubles (zeroed here).  training != 0: batch statistics (+ running update when
// running_mean != NULL); training == 0: statistics from the running buffers.
extern "C" int dl_batchnorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                float* mean, float* rstd, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, double* workspace, int64_t rows,
                                int32_t cols, float eps, float momentum, int32_t training,
                                int32_t dtype, void* stream) {
  DL_REQUIRE(x && y && mean && rstd, "dl_batchnorm_fwd: null pointer");
  DL_REQUIRE(cols > 0 && cols % 4 == 0 && rows >= 1, "dl_batchnorm_fwd: cols must be a positive multiple of 4 and rows >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (training) {
    DL_REQUIRE(workspace != nullptr, "dl_batchnorm_fwd: workspace required in training mode");
    DL_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * cols, st));
    dim3 grid; long long rpb;
    stats_grid(rows, cols, &grid, &rpb);
    if (dtype == DL_BF16)
      bn_colstats_kernel<__nv_bfloat16, false><<<grid, dim3(32, 8), 0, st>>>((const __nv_bfloat16*)x, nullptr, nullptr, nullptr, workspace, rows, cols, rpb);
    else
      bn_colstats_kernel<float, false><<<grid, dim3(32, 8), 0, st>>>((const float*)x, nullptr, nullptr, nullptr, workspace, rows, cols, rpb);
    DL_LAUNCH_CHECK("bn_colstats_kernel");
    bn_finalize_kernel<<<ceil_div(cols, 128), 128, 0, st>>>(workspace, mean, rstd, running_mean, running_var, (long long*)num_batches_tracked, rows, cols, eps, momentum);
    DL_LAUNCH_CHECK("bn_finalize_kernel");
    count_launch(2);
  } else {
    DL_REQUIRE(running_mean && running_var, "dl_batchnorm_fwd: eval mode needs running statistics");
    bn_eval_stats_kernel<<<ceil_div(cols, 128), 128, 0, st>>>(running_mean, running_var, mean, rstd, cols, eps);
    DL_LAUNCH_CHECK("bn_eval_stats_kernel");
    count_launch();
  }
  const long long n4 = rows * cols / 4;
  if (dtype == DL_BF16)
    bn_apply_kernel<__nv_bfloat16><<<ew_grid4(n4), 256, 0, st>>>((const __nv_bfloat16*)x, mean, rstd, gamma, beta, (__nv_bfloat16*)y, n4, cols);
  else
    bn_apply_kernel<float><<<ew_grid4(n4), 256, 0, st>>>((const float*)x, mean, rstd, gamma, beta, (float*)y, n4, cols);
  DL_LAUNCH_CHECK("bn_apply_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_batchnorm_bwd(const void* dy, const void* x, const float* gamma,
                                const float* mean, const float* rstd, void* dx, float* dgamma,
                                float* dbeta, double* workspace, int64_t rows, int32_t cols,
                                int32_t training, int32_t dtype, void* stream) {
  DL_REQUIRE(dy && x && mean && rstd && dx && workspace, "dl_batchnorm_bwd: null pointer");
  DL_REQUIRE(cols > 0 && cols % 4 == 0 && rows >= 1, "dl_batchnorm_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  DL_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * cols, st));
  dim3 grid; long long rpb;
  stats_grid(rows, cols, &grid, &rpb);
  if (dtype == DL_BF16)
    bn_colstats_kernel<__nv_bfloat16, true><<<grid, dim3(32, 8), 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, workspace, rows, cols, rpb);
  else
    bn_colstats_kernel<float, true><<<grid, dim3(32, 8), 0, st>>>((const float*)dy, (const float*)x, mean, rstd, workspace, rows, cols, rpb);
  DL_LAUNCH_CHECK("bn_colstats_kernel(bwd)");
  if (dgamma || dbeta) {
    bn_param_grad_kernel<<<ceil_div(cols, 128), 128, 0, st>>>(workspace, dgamma, dbeta, cols);
    DL_LAUNCH_CHECK("bn_param_grad_kernel");
    count_launch();
  }
  const long long n4 = rows * cols / 4;
  if (dtype == DL_BF16)
    bn_bwd_apply_kernel<__nv_bfloat16><<<ew_grid4(n4), 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, gamma, workspace, (__nv_bfloat16*)dx, n4, cols, 1.0 / (double)rows, training);
  else
    bn_bwd_apply_kernel<float><<<ew_grid4(n4), 256, 0, st>>>((const float*)dy, (const float*)x, mean, rstd, gamma, workspace, (float*)dx, n4, cols, 1.0 / (double)rows, training);
  DL_LAUNCH_CHECK("bn_bwd_apply_kernel");
  count_launch(2);
  return 0;
}
