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
  pdl_trigger();
  pdl_wait();
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

// ---------------------------------------------------------------- CSR construction on the device
// (src, dst) edge list -> CSR by destination and by source in STABLE edge order (bit-identical to a
// host argsort(stable) + bincount), the two deg.clamp(1)^-1/2 vectors of GraphConv
// (model/basic_model.py:596-603,623-630) and the validity flags.  Four small launches, no host sync:
// count degrees (atomics) -> scan -> scatter edge ids (atomic cursors) -> per-row sort by edge id.
__global__ void csr_count_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                                 long long n_edges, long long n_nodes, int* __restrict__ cnt_in,
                                 int* __restrict__ cnt_out, int* __restrict__ flags) {
  pdl_trigger();
  pdl_wait();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges;
       e += (long long)gridDim.x * blockDim.x) {
    const long long s = src[e], d = dst[e];
    if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes) { atomicAdd(flags + 1, 1); continue; }
    atomicAdd(cnt_in + d, 1);
    atomicAdd(cnt_out + s, 1);
  }
}

// one block: exclusive scan of both degree arrays, norms, zero-in-degree count; the degree arrays are
// zeroed afterwards (they become the scatter cursors)
__global__ void __launch_bounds__(1024)
csr_scan_kernel(int* __restrict__ cnt_in, int* __restrict__ cnt_out, long long n_nodes,
                int* __restrict__ indptr, int* __restrict__ indptr_t, float* __restrict__ norm_src,
                float* __restrict__ norm_dst, int* __restrict__ flags) {
  pdl_trigger();
  pdl_wait();
  __shared__ int part[2][1024];
  const int t = threadIdx.x;
  const long long chunk = (n_nodes + 1023) / 1024;
  const long long i0 = t * chunk, i1 = i0 + chunk < n_nodes ? i0 + chunk : n_nodes;
  int a = 0, b = 0, zero_in = 0;
  for (long long i = i0; i < i1; ++i) {
    a += cnt_in[i];
    b += cnt_out[i];
    zero_in += cnt_in[i] == 0;
  }
  part[0][t] = a;
  part[1][t] = b;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {      // Hillis-Steele inclusive scan of the per-thread sums
    const int va = t >= o ? part[0][t - o] : 0, vb = t >= o ? part[1][t - o] : 0;
    __syncthreads();
    part[0][t] += va;
    part[1][t] += vb;
    __syncthreads();
  }
  int pa = part[0][t] - a, pb = part[1][t] - b;
  for (long long i = i0; i < i1; ++i) {
    const int ci = cnt_in[i], co = cnt_out[i];
    indptr[i] = pa;
    indptr_t[i] = pb;
    pa += ci;
    pb += co;
    norm_dst[i] = __fdiv_rn(1.0f, __fsqrt_rn((float)(ci > 1 ? ci : 1)));    // IEEE: equals clamp(deg,1).pow(-0.5)
    norm_src[i] = __fdiv_rn(1.0f, __fsqrt_rn((float)(co > 1 ? co : 1)));
    cnt_in[i] = 0;
    cnt_out[i] = 0;
  }
  if (t == 1023) {
    indptr[n_nodes] = part[0][1023];
    indptr_t[n_nodes] = part[1][1023];
  }
  if (zero_in) atomicAdd(flags, zero_in);
}

__global__ void csr_scatter_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                                   long long n_edges, long long n_nodes, const int* __restrict__ indptr,
                                   const int* __restrict__ indptr_t, int* __restrict__ cur_in,
                                   int* __restrict__ cur_out, int* __restrict__ eid_in, int* __restrict__ eid_out) {
  pdl_trigger();
  pdl_wait();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges;
       e += (long long)gridDim.x * blockDim.x) {
    const long long s = src[e], d = dst[e];
    if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes) continue;
    eid_in[indptr[d] + atomicAdd(cur_in + d, 1)] = (int)e;
    eid_out[indptr_t[s] + atomicAdd(cur_out + s, 1)] = (int)e;
  }
}

// one thread per node and direction: order the row's edge ids (rows are a handful of entries:
// insertion sort), then replace them by the neighbour at the other end
__global__ void csr_finish_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                                  long long n_nodes, const int* __restrict__ indptr,
                                  const int* __restrict__ indptr_t, int* __restrict__ eid_in,
                                  int* __restrict__ eid_out, int* __restrict__ indices, int* __restrict__ indices_t) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n_nodes) return;
  const bool t = i >= n_nodes;
  const long long row = t ? i - n_nodes : i;
  const int* ptr = t ? indptr_t : indptr;
  int* eid = t ? eid_out : eid_in;
  const long long* other = t ? dst : src;
  int* out = t ? indices_t : indices;
  const int e0 = ptr[row], e1 = ptr[row + 1];
  for (int a = e0 + 1; a < e1; ++a) {
    const int v = eid[a];
    int b = a - 1;
    while (b >= e0 && eid[b] > v) { eid[b + 1] = eid[b]; --b; }
    eid[b + 1] = v;
  }
  for (int a = e0; a < e1; ++a) out[a] = (int)other[eid[a]];
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
  pdl_trigger();
  pdl_wait();
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
  pdl_trigger();
  pdl_wait();
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
  pdl_trigger();
  pdl_wait();
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
  pdl_trigger();
  pdl_wait();
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

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, float* __restrict__ dgamma, int acc,
                                     float* __restrict__ dbeta, int cols) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) {
    if (dbeta) dbeta[c] = (acc ? dbeta[c] : 0.f) + (float)sums[c];
    if (dgamma) dgamma[c] = (acc ? dgamma[c] : 0.f) + (float)sums[cols + c];
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var, float* __restrict__ mean,
                                     float* __restrict__ rstd, int cols, float eps) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) {
    mean[c] = running_mean[c];
    rstd[c] = rsqrtf(running_var[c] + eps);
  }
}

// ---------------------------------------------------------------- vectorised BatchNorm passes
// Column-slice layout shared by the three kernels below: `tpr` threads cover one row slice with
// 16-byte accesses (thread = V fixed columns, so the per-column constants live in registers),
// 256 / tpr rows per pass, blockIdx.y strides over row blocks.  Needs cols % V == 0.
struct ColSlice {
  int tpr, xb;
  unsigned yb;
  long long rpb;
};

template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
bn_colstats_vec_kernel(const T* __restrict__ a, const T* __restrict__ x, const float* __restrict__ mean,
                       const float* __restrict__ rstd, double* __restrict__ sums, long long rows,
                       int cols, long long rows_per_block, int tpr) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = VecWidth<T>::N;
  __shared__ float red[2][256][V + 1];
  const int vcol = threadIdx.x % tpr, rsub = threadIdx.x / tpr, rpp = 256 / tpr;
  const int c = (blockIdx.x * tpr + vcol) * V;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s1[V], s2[V], mu[V], rs[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { s1[j] = 0.f; s2[j] = 0.f; mu[j] = 0.f; rs[j] = 0.f; }
  if (c < cols) {
    if (BWD) {
#pragma unroll
      for (int j = 0; j < V; ++j) { mu[j] = mean[c + j]; rs[j] = rstd[c + j]; }
    }
    long long r = r0 + rsub;
    for (; r + rpp < r1; r += 2 * rpp) {             // two rows (four 16-byte loads) in flight
      float av[2][V], xv[2][V];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        ldv(a + (r + u * rpp) * cols + c, av[u]);
        if (BWD) ldv(x + (r + u * rpp) * cols + c, xv[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float bv = BWD ? (xv[u][j] - mu[j]) * rs[j] : av[u][j];
          s1[j] += av[u][j];
          s2[j] = fmaf(av[u][j], bv, s2[j]);
        }
    }
    for (; r < r1; r += rpp) {
      float av[V], xv[V];
      ldv(a + r * cols + c, av);
      if (BWD) ldv(x + r * cols + c, xv);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float bv = BWD ? (xv[j] - mu[j]) * rs[j] : av[j];
        s1[j] += av[j];
        s2[j] = fmaf(av[j], bv, s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) { red[0][threadIdx.x][j] = s1[j]; red[1][threadIdx.x][j] = s2[j]; }
  __syncthreads();
  // per-thread fp32 partials cover at most a few dozen rows; everything above that is fp64
  for (int o = threadIdx.x; o < tpr * V; o += 256) {
    const int vc = o / V, j = o % V;
    const int col = (blockIdx.x * tpr + vc) * V + j;
    if (col < cols) {
      double t1 = 0.0, t2 = 0.0;
      for (int i = 0; i < rpp; ++i) { t1 += red[0][i * tpr + vc][j]; t2 += red[1][i * tpr + vc][j]; }
      atomicAdd(sums + col, t1);
      atomicAdd(sums + cols + col, t2);
    }
  }
}

// The LAST row stands for `w` identical rows (the molecular GCN's virtual nodes):
// add the (w - 1) missing copies to the column sums.
template <typename T>
__global__ void bn_lastrow_fix_kernel(const T* __restrict__ xlast, double* __restrict__ sums, int cols, double wm1) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) {
    const double v = (double)Cvt<T>::to_f(xlast[c]);
    sums[c] += wm1 * v;
    sums[cols + c] += wm1 * v * v;
  }
}

// y = (x - mean) * rstd * gamma + beta.  FINALIZE: mean / rstd come from the fp64 column sums
// (biased variance); the first row block also stores them for the backward pass and updates the
// running statistics (momentum, unbiased variance) and num_batches_tracked.
template <typename T, bool FINALIZE>
__global__ void __launch_bounds__(256)
bn_apply_vec_kernel(const T* __restrict__ x, const double* __restrict__ sums, float* __restrict__ mean,
                    float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ running_mean,
                    float* __restrict__ running_var, long long* __restrict__ nbt, T* __restrict__ y,
                    long long rows, int cols, long long rows_per_block, int tpr, float eps,
                    float momentum, double count) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = VecWidth<T>::N;
  const int vcol = threadIdx.x % tpr, rsub = threadIdx.x / tpr, rpp = 256 / tpr;
  const int c = (blockIdx.x * tpr + vcol) * V;
  if (FINALIZE && nbt && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *nbt += 1;
  if (c >= cols) return;
  float sc[V], sh[V], mu_[V];
  const bool writer = blockIdx.y == 0 && rsub == 0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float mu, rs;
    if (FINALIZE) {
      const double m = sums[c + j] / count;            // count = rows, or the weighted row count
      double var = sums[cols + c + j] / count - m * m;
      if (var < 0.0) var = 0.0;
      mu = (float)m;
      rs = (float)(1.0 / sqrt(var + (double)eps));
      if (writer) {
        mean[c + j] = mu;
        rstd[c + j] = rs;
        if (running_mean) {
          const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
          running_mean[c + j] = (1.f - momentum) * running_mean[c + j] + momentum * mu;
          running_var[c + j] = (1.f - momentum) * running_var[c + j] + momentum * (float)unb;
        }
      }
    } else {
      mu = mean[c + j];
      rs = rstd[c + j];
    }
    sc[j] = rs * (gamma ? gamma[c + j] : 1.f);
    sh[j] = beta ? beta[c + j] : 0.f;
    mu_[j] = mu;
  }
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  long long r = r0 + rsub;
  for (; r + 3 * rpp < r1; r += 4 * rpp) {
    float v[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) ldv(x + (r + u * rpp) * cols + c, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[u][j] = fmaf(v[u][j] - mu_[j], sc[j], sh[j]);
      stv(y + (r + u * rpp) * cols + c, v[u]);
    }
  }
  for (; r < r1; r += rpp) {
    float v[V];
    ldv(x + r * cols + c, v);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = fmaf(v[j] - mu_[j], sc[j], sh[j]);
    stv(y + r * cols + c, v);
  }
}

// training: dx = gamma*rstd * (dy - sum_dy/N - xhat * sum_dyxhat/N); eval: dx = gamma*rstd*dy.
// The first row block also writes dbeta = sum dy, dgamma = sum dy*xhat.
// relu != 0: x is the output of a ReLU (ProteinCNN: conv -> ReLU -> BN); its backward mask
// (x > 0) is applied to dx here instead of in a separate pass over the gradient.
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_vec_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                        const double* __restrict__ sums, T* __restrict__ dx, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, long long rows, int cols, long long rows_per_block,
                        int tpr, int training, int acc, int relu, double count, float last_w) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = VecWidth<T>::N;
  const int vcol = threadIdx.x % tpr, rsub = threadIdx.x / tpr, rpp = 256 / tpr;
  const int c = (blockIdx.x * tpr + vcol) * V;
  if (c >= cols) return;
  // dx = k0 * (dy - m1 - (x - mu) * k1) with per-column constants
  float k0[V], k1[V], m1[V], mu_[V];
  const bool writer = blockIdx.y == 0 && rsub == 0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const float g = gamma ? gamma[c + j] : 1.f, rs = rstd[c + j], mu = mean[c + j];
    const double sdy = sums[c + j], sdyx = sums[cols + c + j];
    if (writer) {
      if (dbeta) dbeta[c + j] = (acc ? dbeta[c + j] : 0.f) + (float)sdy;
      if (dgamma) dgamma[c + j] = (acc ? dgamma[c + j] : 0.f) + (float)sdyx;
    }
    k0[j] = g * rs;
    mu_[j] = mu;
    m1[j] = training ? (float)(sdy / count) : 0.f;
    k1[j] = training ? rs * (float)(sdyx / count) : 0.f;
  }
  // last_w > 1: the last row stands for last_w identical rows and its dy is already the SUM of their
  // gradients, so the batch-statistics correction applies last_w times to it
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  long long r = r0 + rsub;
  for (; r + rpp < r1; r += 2 * rpp) {
    float d[2][V], xv[2][V];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ldv(dy + (r + u * rpp) * cols + c, d[u]);
      ldv(x + (r + u * rpp) * cols + c, xv[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float w = (r + u * rpp == rows - 1) ? last_w : 1.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float t = k0[j] * (d[u][j] - w * (m1[j] + (xv[u][j] - mu_[j]) * k1[j]));
        d[u][j] = (relu && !(xv[u][j] > 0.f)) ? 0.f : t;
      }
      stv(dx + (r + u * rpp) * cols + c, d[u]);
    }
  }
  for (; r < r1; r += rpp) {
    float d[V], xv[V];
    ldv(dy + r * cols + c, d);
    ldv(x + r * cols + c, xv);
    const float w = (r == rows - 1) ? last_w : 1.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float t = k0[j] * (d[j] - w * (m1[j] + (xv[j] - mu_[j]) * k1[j]));
      d[j] = (relu && !(xv[j] > 0.f)) ? 0.f : t;
    }
    stv(dx + r * cols + c, d);
  }
}

// geometry for the vectorised passes; false when the shape / alignment does not qualify
bool col_slice(const void* p0, const void* p1, const void* p2, long long rows, int cols, int vec,
               int max_rows_per_thread, ColSlice* g) {
  if (cols % vec != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2)) & 15)
    return false;
  int tpr = 1;
  while (tpr < 256 && tpr * vec < cols) tpr <<= 1;
  g->tpr = tpr;
  g->xb = ceil_div(cols, tpr * vec);
  const int rpp = 256 / tpr;
  long long yb = (long long)sm_count() * 2 / g->xb;
  const long long need = (rows + (long long)max_rows_per_thread * rpp - 1) / ((long long)max_rows_per_thread * rpp);
  if (yb < need) yb = need;
  const long long maxy = (rows + 2 * rpp - 1) / (2 * rpp);
  if (yb > maxy) yb = maxy;
  if (yb < 1) yb = 1;
  if (yb > 65535) return false;
  g->yb = (unsigned)yb;
  g->rpb = (rows + yb - 1) / yb;
  return true;
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
    DL_LAUNCH((spmm_norm_kernel<__nv_bfloat16>), grid, 256, 0, st, indptr, indices, norm_src, norm_dst, (const __nv_bfloat16*)h, (__nv_bfloat16*)out, (int)n_rows);
  else
    DL_LAUNCH((spmm_norm_kernel<float>), grid, 256, 0, st, indptr, indices, norm_src, norm_dst, (const float*)h, (float*)out, (int)n_rows);
  DL_LAUNCH_CHECK("spmm_norm_kernel");
  count_launch();
  return 0;
}

// workspace: 2*cols doubles (zeroed here).  training != 0: batch statistics (+ running update when
// running_mean != NULL); training == 0: statistics from the running buffers.
extern "C" int dl_csr_build(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n_nodes,
                            int32_t* indptr, int32_t* indices, int32_t* indptr_t, int32_t* indices_t,
                            float* norm_src, float* norm_dst, int32_t* flags, int32_t* workspace,
                            void* stream) {
  using namespace dl;
  DL_REQUIRE(src && dst && indptr && indices && indptr_t && indices_t && norm_src && norm_dst && flags && workspace,
             "dl_csr_build: null pointer");
  DL_REQUIRE(n_nodes >= 1 && n_edges >= 0 && n_nodes < (1ll << 30) && n_edges < (1ll << 31), "dl_csr_build: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  int* cnt_in = workspace;
  int* cnt_out = workspace + n_nodes;
  int* eid_in = workspace + 2 * n_nodes;
  int* eid_out = eid_in + n_edges;
  DL_CUDA(cudaMemsetAsync(workspace, 0, (size_t)(2 * n_nodes) * sizeof(int), st));
  DL_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), st));
  const int eg = n_edges > 0 ? (int)((n_edges + 255) / 256 < 1184 ? (n_edges + 255) / 256 : 1184) : 1;
  DL_LAUNCH(csr_count_kernel, eg, 256, 0, st, (const long long*)src, (const long long*)dst, (long long)n_edges,
            (long long)n_nodes, cnt_in, cnt_out, flags);
  DL_LAUNCH_CHECK("csr_count_kernel");
  DL_LAUNCH(csr_scan_kernel, 1, 1024, 0, st, cnt_in, cnt_out, (long long)n_nodes, indptr, indptr_t, norm_src,
            norm_dst, flags);
  DL_LAUNCH_CHECK("csr_scan_kernel");
  DL_LAUNCH(csr_scatter_kernel, eg, 256, 0, st, (const long long*)src, (const long long*)dst, (long long)n_edges,
            (long long)n_nodes, (const int*)indptr, (const int*)indptr_t, cnt_in, cnt_out, eid_in, eid_out);
  DL_LAUNCH_CHECK("csr_scatter_kernel");
  DL_LAUNCH(csr_finish_kernel, (int)((2 * n_nodes + 255) / 256), 256, 0, st, (const long long*)src,
            (const long long*)dst, (long long)n_nodes, (const int*)indptr, (const int*)indptr_t, eid_in, eid_out,
            indices, indices_t);
  DL_LAUNCH_CHECK("csr_finish_kernel");
  count_launch(4);
  return 0;
}

extern "C" int dl_batchnorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                float* mean, float* rstd, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, double* workspace, int64_t rows,
                                int32_t cols, float eps, float momentum, int32_t training,
                                float last_row_weight, int32_t dtype, void* stream) {
  DL_REQUIRE(x && mean && rstd, "dl_batchnorm_fwd: null pointer");
  const bool weighted = last_row_weight > 1.f;
  if (y == nullptr) {
    // statistics only (mean / rstd, running buffers): the normalisation itself is fused into the
    // consumer (dl_bn_transpose)
    DL_REQUIRE(!weighted, "dl_batchnorm_fwd: statistics-only mode does not take last_row_weight");
    DL_REQUIRE(cols > 0 && cols % 4 == 0 && rows >= 1, "dl_batchnorm_fwd: cols must be a positive multiple of 4 and rows >= 1");
    cudaStream_t st0 = (cudaStream_t)stream;
    if (!training) {
      DL_REQUIRE(running_mean && running_var, "dl_batchnorm_fwd: eval mode needs running statistics");
      DL_LAUNCH(bn_eval_stats_kernel, ceil_div(cols, 128), 128, 0, st0, running_mean, running_var, mean, rstd, cols, eps);
      DL_LAUNCH_CHECK("bn_eval_stats_kernel");
      count_launch();
      return 0;
    }
    DL_REQUIRE(workspace != nullptr, "dl_batchnorm_fwd: workspace required in training mode");
    DL_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * cols, st0));
    ColSlice g0;
    if (col_slice(x, nullptr, nullptr, rows, cols, dtype == DL_BF16 ? 8 : 4, 32, &g0)) {
      const dim3 grid0(g0.xb, g0.yb);
      if (dtype == DL_BF16)
        DL_LAUNCH((bn_colstats_vec_kernel<__nv_bfloat16, false>), grid0, 256, 0, st0, (const __nv_bfloat16*)x, nullptr, nullptr, nullptr, workspace, rows, cols, g0.rpb, g0.tpr);
      else
        DL_LAUNCH((bn_colstats_vec_kernel<float, false>), grid0, 256, 0, st0, (const float*)x, nullptr, nullptr, nullptr, workspace, rows, cols, g0.rpb, g0.tpr);
      DL_LAUNCH_CHECK("bn_colstats_vec_kernel");
    } else {
      dim3 grid0; long long rpb0;
      stats_grid(rows, cols, &grid0, &rpb0);
      if (dtype == DL_BF16)
        DL_LAUNCH((bn_colstats_kernel<__nv_bfloat16, false>), grid0, dim3(32, 8), 0, st0, (const __nv_bfloat16*)x, nullptr, nullptr, nullptr, workspace, rows, cols, rpb0);
      else
        DL_LAUNCH((bn_colstats_kernel<float, false>), grid0, dim3(32, 8), 0, st0, (const float*)x, nullptr, nullptr, nullptr, workspace, rows, cols, rpb0);
      DL_LAUNCH_CHECK("bn_colstats_kernel");
    }
    DL_LAUNCH(bn_finalize_kernel, ceil_div(cols, 128), 128, 0, st0, workspace, mean, rstd, running_mean, running_var, (long long*)num_batches_tracked, rows, cols, eps, momentum);
    DL_LAUNCH_CHECK("bn_finalize_kernel");
    count_launch(2);
    return 0;
  }
  const double count = weighted ? (double)rows - 1.0 + (double)last_row_weight : (double)rows;
  DL_REQUIRE(cols > 0 && cols % 4 == 0 && rows >= 1, "dl_batchnorm_fwd: cols must be a positive multiple of 4 and rows >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  ColSlice g;
  const int vec = dtype == DL_BF16 ? 8 : 4;
  if (col_slice(x, y, nullptr, rows, cols, vec, 32, &g)) {
    const dim3 grid(g.xb, g.yb);
    if (training) {
      DL_REQUIRE(workspace != nullptr, "dl_batchnorm_fwd: workspace required in training mode");
      DL_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * cols, st));
      if (dtype == DL_BF16) {
        DL_LAUNCH((bn_colstats_vec_kernel<__nv_bfloat16, false>), grid, 256, 0, st, (const __nv_bfloat16*)x, nullptr, nullptr, nullptr, workspace, rows, cols, g.rpb, g.tpr);
        if (weighted)
          DL_LAUNCH((bn_lastrow_fix_kernel<__nv_bfloat16>), ceil_div(cols, 128), 128, 0, st, (const __nv_bfloat16*)x + (rows - 1) * cols, workspace, cols, (double)last_row_weight - 1.0);
        DL_LAUNCH((bn_apply_vec_kernel<__nv_bfloat16, true>), grid, 256, 0, st, (const __nv_bfloat16*)x, workspace, mean, rstd, gamma, beta, running_mean, running_var, (long long*)num_batches_tracked, (__nv_bfloat16*)y, rows, cols, g.rpb, g.tpr, eps, momentum, count);
      } else {
        DL_LAUNCH((bn_colstats_vec_kernel<float, false>), grid, 256, 0, st, (const float*)x, nullptr, nullptr, nullptr, workspace, rows, cols, g.rpb, g.tpr);
        if (weighted)
          DL_LAUNCH((bn_lastrow_fix_kernel<float>), ceil_div(cols, 128), 128, 0, st, (const float*)x + (rows - 1) * cols, workspace, cols, (double)last_row_weight - 1.0);
        DL_LAUNCH((bn_apply_vec_kernel<float, true>), grid, 256, 0, st, (const float*)x, workspace, mean, rstd, gamma, beta, running_mean, running_var, (long long*)num_batches_tracked, (float*)y, rows, cols, g.rpb, g.tpr, eps, momentum, count);
      }
      if (weighted) count_launch();
      DL_LAUNCH_CHECK("bn_colstats_vec_kernel / bn_apply_vec_kernel");
      count_launch(2);
    } else {
      DL_REQUIRE(running_mean && running_var, "dl_batchnorm_fwd: eval mode needs running statistics");
      DL_LAUNCH(bn_eval_stats_kernel, ceil_div(cols, 128), 128, 0, st, running_mean, running_var, mean, rstd, cols, eps);
      if (dtype == DL_BF16)
        DL_LAUNCH((bn_apply_vec_kernel<__nv_bfloat16, false>), grid, 256, 0, st, (const __nv_bfloat16*)x, nullptr, mean, rstd, gamma, beta, nullptr, nullptr, nullptr, (__nv_bfloat16*)y, rows, cols, g.rpb, g.tpr, eps, momentum, count);
      else
        DL_LAUNCH((bn_apply_vec_kernel<float, false>), grid, 256, 0, st, (const float*)x, nullptr, mean, rstd, gamma, beta, nullptr, nullptr, nullptr, (float*)y, rows, cols, g.rpb, g.tpr, eps, momentum, count);
      DL_LAUNCH_CHECK("bn_eval_stats_kernel / bn_apply_vec_kernel");
      count_launch(2);
    }
    return 0;
  }
  DL_REQUIRE(!weighted, "dl_batchnorm_fwd: last_row_weight needs cols %% (16 / element size) == 0 and 16-byte aligned tensors");
  if (training) {
    DL_REQUIRE(workspace != nullptr, "dl_batchnorm_fwd: workspace required in training mode");
    DL_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * cols, st));
    dim3 grid; long long rpb;
    stats_grid(rows, cols, &grid, &rpb);
    if (dtype == DL_BF16)
      DL_LAUNCH((bn_colstats_kernel<__nv_bfloat16, false>), grid, dim3(32, 8), 0, st, (const __nv_bfloat16*)x, nullptr, nullptr, nullptr, workspace, rows, cols, rpb);
    else
      DL_LAUNCH((bn_colstats_kernel<float, false>), grid, dim3(32, 8), 0, st, (const float*)x, nullptr, nullptr, nullptr, workspace, rows, cols, rpb);
    DL_LAUNCH_CHECK("bn_colstats_kernel");
    DL_LAUNCH(bn_finalize_kernel, ceil_div(cols, 128), 128, 0, st, workspace, mean, rstd, running_mean, running_var, (long long*)num_batches_tracked, rows, cols, eps, momentum);
    DL_LAUNCH_CHECK("bn_finalize_kernel");
    count_launch(2);
  } else {
    DL_REQUIRE(running_mean && running_var, "dl_batchnorm_fwd: eval mode needs running statistics");
    DL_LAUNCH(bn_eval_stats_kernel, ceil_div(cols, 128), 128, 0, st, running_mean, running_var, mean, rstd, cols, eps);
    DL_LAUNCH_CHECK("bn_eval_stats_kernel");
    count_launch();
  }
  const long long n4 = rows * cols / 4;
  if (dtype == DL_BF16)
    DL_LAUNCH((bn_apply_kernel<__nv_bfloat16>), ew_grid4(n4), 256, 0, st, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, (__nv_bfloat16*)y, n4, cols);
  else
    DL_LAUNCH((bn_apply_kernel<float>), ew_grid4(n4), 256, 0, st, (const float*)x, mean, rstd, gamma, beta, (float*)y, n4, cols);
  DL_LAUNCH_CHECK("bn_apply_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_batchnorm_bwd(const void* dy, const void* x, const float* gamma,
                                const float* mean, const float* rstd, void* dx, float* dgamma,
                                float* dbeta, double* workspace, int64_t rows, int32_t cols,
                                int32_t training, int32_t accumulate, int32_t relu_mask,
                                float last_row_weight, int32_t dtype, void* stream) {
  DL_REQUIRE(dy && x && mean && rstd && dx && workspace, "dl_batchnorm_bwd: null pointer");
  const bool weighted = last_row_weight > 1.f;
  const double count = weighted ? (double)rows - 1.0 + (double)last_row_weight : (double)rows;
  const float last_w = weighted ? last_row_weight : 1.f;
  DL_REQUIRE(cols > 0 && cols % 4 == 0 && rows >= 1, "dl_batchnorm_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  DL_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * cols, st));
  ColSlice g;
  if (col_slice(dy, x, dx, rows, cols, dtype == DL_BF16 ? 8 : 4, 32, &g)) {
    const dim3 vgrid(g.xb, g.yb);
    if (dtype == DL_BF16) {
      DL_LAUNCH((bn_colstats_vec_kernel<__nv_bfloat16, true>), vgrid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, workspace, rows, cols, g.rpb, g.tpr);
      DL_LAUNCH((bn_bwd_apply_vec_kernel<__nv_bfloat16>), vgrid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, gamma, workspace, (__nv_bfloat16*)dx, dgamma, dbeta, rows, cols, g.rpb, g.tpr, training, accumulate, relu_mask, count, last_w);
    } else {
      DL_LAUNCH((bn_colstats_vec_kernel<float, true>), vgrid, 256, 0, st, (const float*)dy, (const float*)x, mean, rstd, workspace, rows, cols, g.rpb, g.tpr);
      DL_LAUNCH((bn_bwd_apply_vec_kernel<float>), vgrid, 256, 0, st, (const float*)dy, (const float*)x, mean, rstd, gamma, workspace, (float*)dx, dgamma, dbeta, rows, cols, g.rpb, g.tpr, training, accumulate, relu_mask, count, last_w);
    }
    DL_LAUNCH_CHECK("bn_colstats_vec_kernel / bn_bwd_apply_vec_kernel");
    count_launch(2);
    return 0;
  }
  DL_REQUIRE(!relu_mask && !weighted, "dl_batchnorm_bwd: relu_mask / last_row_weight need cols %% (16 / element size) == 0 and 16-byte aligned tensors");
  dim3 grid; long long rpb;
  stats_grid(rows, cols, &grid, &rpb);
  if (dtype == DL_BF16)
    DL_LAUNCH((bn_colstats_kernel<__nv_bfloat16, true>), grid, dim3(32, 8), 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, workspace, rows, cols, rpb);
  else
    DL_LAUNCH((bn_colstats_kernel<float, true>), grid, dim3(32, 8), 0, st, (const float*)dy, (const float*)x, mean, rstd, workspace, rows, cols, rpb);
  DL_LAUNCH_CHECK("bn_colstats_kernel(bwd)");
  if (dgamma || dbeta) {
    DL_LAUNCH(bn_param_grad_kernel, ceil_div(cols, 128), 128, 0, st, workspace, dgamma, accumulate, dbeta, cols);
    DL_LAUNCH_CHECK("bn_param_grad_kernel");
    count_launch();
  }
  const long long n4 = rows * cols / 4;
  if (dtype == DL_BF16)
    DL_LAUNCH((bn_bwd_apply_kernel<__nv_bfloat16>), ew_grid4(n4), 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, gamma, workspace, (__nv_bfloat16*)dx, n4, cols, 1.0 / (double)rows, training);
  else
    DL_LAUNCH((bn_bwd_apply_kernel<float>), ew_grid4(n4), 256, 0, st, (const float*)dy, (const float*)x, mean, rstd, gamma, workspace, (float*)dx, n4, cols, 1.0 / (double)rows, training);
  DL_LAUNCH_CHECK("bn_bwd_apply_kernel");
  count_launch(2);
  return 0;
}
