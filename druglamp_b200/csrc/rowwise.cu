// HBM-bound row kernels of the fusion path: LayerNorm fwd/bwd, row softmax fwd/bwd, column sums
// (bias gradients), dropout, dtype casts and the positional-embedding add.  One warp owns one
// row; loads are 8/16-byte vectors where the row length allows; statistics are fp32 and
// reduced with warp shuffles.  Column reductions finish with one fp32 atomic per block.
#include "../../include/druglamp_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace dl {
void count_launch(int n = 1);
namespace {

constexpr int kWarpsPerBlock = 8;

__host__ int row_grid(long long rows) {
  long long blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long long cap = (long long)sm_count() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// ------------------------------------------------------------------ LayerNorm
// One warp per row, C = 128..1024 columns.  A lane owns EPL = C/32 elements, fetched as 16-byte
// vectors (8 bf16 / 4 fp32; 8-byte for the 128-wide bf16 case): vector k of lane l covers columns
// k*32*V + l*V .. +V, so every warp access is one contiguous 512-byte (or 256-byte) run.  Two rows
// are in flight per warp iteration so the loads of one overlap the shuffle reductions of the other.
template <typename T, int C>
struct LnLayout {
  static constexpr int EPL = C / 32;
  static constexpr int VMAX = 16 / (int)sizeof(T);
  static constexpr int V = VMAX < EPL ? VMAX : EPL;
  static constexpr int NV = EPL / V;
};

template <typename T, int V>
__device__ __forceinline__ void ln_load(const T* p, float* out) {
  if constexpr (V == 8) {
    float t[8];
    ldv(reinterpret_cast<const __nv_bfloat16*>(p), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = t[i];
  } else {
    const float4 f = ld4<T>(p);
    out[0] = f.x; out[1] = f.y; out[2] = f.z; out[3] = f.w;
  }
}
template <typename T, int V>
__device__ __forceinline__ void ln_store(T* p, const float* in) {
  if constexpr (V == 8) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = in[i];
    stv(reinterpret_cast<__nv_bfloat16*>(p), t);
  } else {
    st4<T>(p, make_float4(in[0], in[1], in[2], in[3]));
  }
}

template <typename T, int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, long long rows, float eps) {
  pdl_trigger();
  pdl_wait();
  using LY = LnLayout<T, C>;
  constexpr int EPL = LY::EPL, V = LY::V, NV = LY::NV;
  constexpr int R = EPL <= 16 ? 2 : 1;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r0 = warp0; r0 < rows; r0 += R * nwarps) {
    float v[R][EPL], s[R];
    bool on[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long r = r0 + u * nwarps;
      on[u] = r < rows;
      s[u] = 0.f;
      if (on[u]) {
#pragma unroll
        for (int k = 0; k < NV; ++k) ln_load<T, V>(x + r * C + k * 32 * V + lane * V, v[u] + k * V);
      } else {
#pragma unroll
        for (int e = 0; e < EPL; ++e) v[u][e] = 0.f;
      }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) s[u] += v[u][e];
    }
#pragma unroll
    for (int u = 0; u < R; ++u) mean[u] = warp_sum(s[u]) * (1.f / C);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      float q = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) { const float d = v[u][e] - mean[u]; q += d * d; }
      s[u] = q;
    }
#pragma unroll
    for (int u = 0; u < R; ++u) rstd[u] = rsqrtf(warp_sum(s[u]) * (1.f / C) + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = k * 32 * V + lane * V;
      float g[V], b[V];
#pragma unroll
      for (int e = 0; e < V; e += 4) {
        const float4 gg = *reinterpret_cast<const float4*>(gamma + c0 + e);
        const float4 bb = *reinterpret_cast<const float4*>(beta + c0 + e);
        g[e] = gg.x; g[e + 1] = gg.y; g[e + 2] = gg.z; g[e + 3] = gg.w;
        b[e] = bb.x; b[e + 1] = bb.y; b[e + 2] = bb.z; b[e + 3] = bb.w;
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (!on[u]) continue;
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = (v[u][k * V + e] - mean[u]) * rstd[u] * g[e] + b[e];
        ln_store<T, V>(y + (r0 + u * nwarps) * C + c0, o);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (!on[u]) continue;
        if (mean_out) mean_out[r0 + u * nwarps] = mean[u];
        if (rstd_out) rstd_out[r0 + u * nwarps] = rstd[u];
      }
    }
  }
}

template <typename T, int C, int CL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, C <= 512 ? 2 : 1)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, T* __restrict__ dx, const T* __restrict__ dx_add,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows) {
  pdl_trigger();
  pdl_wait();
  using LY = LnLayout<T, C>;
  constexpr int EPL = LY::EPL, V = LY::V, NV = LY::NV;
  // rows in flight per warp: the grid is capped (atomics), so memory-level parallelism has to come from
  // the warp itself -- 4 rows (C <= 256) / 2 rows (C = 512) of x and dy are loaded before the first
  // reduction starts
  constexpr int R = EPL <= 8 ? 4 : (EPL <= 16 ? 2 : 1);
  __shared__ float red[kWarpsPerBlock][C];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + w;
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  float ag[EPL], ab[EPL], gm[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) { ag[e] = 0.f; ab[e] = 0.f; }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int e = 0; e < V; e += 4) {
      const float4 gg = *reinterpret_cast<const float4*>(gamma + k * 32 * V + lane * V + e);
      gm[k * V + e] = gg.x; gm[k * V + e + 1] = gg.y; gm[k * V + e + 2] = gg.z; gm[k * V + e + 3] = gg.w;
    }
  }
  for (long long r0 = warp0; r0 < rows; r0 += R * nwarps) {
    float xh[R][EPL], dg[R][EPL], s1[R], s2[R], rs[R];
    bool on[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long r = r0 + u * nwarps;
      on[u] = r < rows;
      s1[u] = s2[u] = 0.f;
      rs[u] = 0.f;
      if (on[u]) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          ln_load<T, V>(x + r * C + k * 32 * V + lane * V, xh[u] + k * V);
          ln_load<T, V>(dy + r * C + k * 32 * V + lane * V, dg[u] + k * V);
        }
        const float mean = mean_in[r];
        rs[u] = rstd_in[r];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const float h = (xh[u][e] - mean) * rs[u], d = dg[u][e];
          xh[u][e] = h;
          ag[e] += d * h;
          ab[e] += d;
          dg[u][e] = d * gm[e];
          s1[u] += dg[u][e];
          s2[u] += dg[u][e] * h;
        }
      }
    }
    float c1[R], c2[R];
#pragma unroll
    for (int u = 0; u < R; ++u) { c1[u] = warp_sum(s1[u]) * (1.f / C); c2[u] = warp_sum(s2[u]) * (1.f / C); }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (!on[u]) continue;
      const long long r = r0 + u * nwarps;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = k * 32 * V + lane * V;
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = rs[u] * (dg[u][k * V + e] - c1[u] - xh[u][k * V + e] * c2[u]);
        if (dx_add) {                        // the residual branch's gradient joins here
          float a[V];
          ln_load<T, V>(dx_add + r * C + c0, a);
#pragma unroll
          for (int e = 0; e < V; ++e) o[e] += a[e];
        }
        ln_store<T, V>(dx + r * C + c0, o);
      }
    }
  }
  // Column partials: per-warp registers -> per-block sums in shared memory -> (CL > 1) summed over the CL
  // blocks of a cluster through distributed shared memory -> ONE 16-byte vector reduction per four columns
  // per cluster.  One atomicAdd per column per block was the kernel's tail: 296 blocks x 512 same-address
  // atomics serialise in L2 for longer than the 25 MB of row traffic take (12.5 us at 2.0 TB/s).
  __shared__ __align__(16) float tot[2][C];
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
      for (int e = 0; e < V; ++e) red[w][k * 32 * V + lane * V + e] = pass == 0 ? ag[k * V + e] : ab[k * V + e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kWarpsPerBlock * 32) {
      float t = 0.f;
#pragma unroll
      for (int ww = 0; ww < kWarpsPerBlock; ++ww) t += red[ww][c];
      tot[pass][c] = t;
    }
  }
  if constexpr (CL > 1) ptx::cluster_sync_all();
  else __syncthreads();
  if (CL == 1 || ptx::cluster_ctarank() == 0) {
    for (int i = threadIdx.x; i < 2 * (C / 4); i += kWarpsPerBlock * 32) {
      const int pass = i / (C / 4), c = (i % (C / 4)) * 4;
      float* dst = pass == 0 ? dgamma : dbeta;
      if (!dst) continue;
      float4 t = *reinterpret_cast<const float4*>(&tot[pass][c]);
      if constexpr (CL > 1) {
        const uint32_t a = ptx::smem_u32(&tot[pass][c]);
#pragma unroll
        for (int r = 1; r < CL; ++r) {
          float4 u;
          asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(u.x), "=f"(u.y), "=f"(u.z), "=f"(u.w) : "r"(ptx::mapa(a, (uint32_t)r)));
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
      }
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                     :: "l"(dst + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
      } else {
        atomicAdd(dst + c, t.x); atomicAdd(dst + c + 1, t.y); atomicAdd(dst + c + 2, t.z); atomicAdd(dst + c + 3, t.w);
      }
    }
  }
  if constexpr (CL > 1) ptx::cluster_sync_all();     // the other blocks' partials stay alive until block 0 has read them
}

// ------------------------------------------------------------------ row softmax
// cols <= 32 * kMaxPerLane; scalar coalesced accesses (lane stride 1)
constexpr int kMaxPerLane = 32;

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_fwd_kernel(const T* __restrict__ s, T* __restrict__ p, long long rows, int cols,
                   long long ld) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int n = (cols + 31) / 32;
  for (long long r = warp0; r < rows; r += nwarps) {
    float v[kMaxPerLane];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
      if (j < n) {
        const int c = j * 32 + lane;
        v[j] = c < cols ? ldf<T>(s, r * ld + c) : -INFINITY;
        m = fmaxf(m, v[j]);
      }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
      if (j < n) {
        v[j] = __expf(v[j] - m);
        sum += v[j];
      }
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
      if (j < n) {
        const int c = j * 32 + lane;
        if (c < cols) stf<T>(p, r * ld + c, v[j] * inv);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_bwd_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ ds,
                   long long rows, int cols, long long ld, float scale) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int n = (cols + 31) / 32;
  for (long long r = warp0; r < rows; r += nwarps) {
    float pv[kMaxPerLane], dv[kMaxPerLane];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
      if (j < n) {
        const int c = j * 32 + lane;
        pv[j] = c < cols ? ldf<T>(p, r * ld + c) : 0.f;
        dv[j] = c < cols ? ldf<T>(dp, r * ld + c) : 0.f;
        dot += pv[j] * dv[j];
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
      if (j < n) {
        const int c = j * 32 + lane;
        if (c < cols) stf<T>(ds, r * ld + c, scale * pv[j] * (dv[j] - dot));
      }
    }
  }
}

// Vectorised variants for cols == NCH * 256 (256 and 512 are the attention widths): each lane owns
// NCH runs of 8 consecutive elements = 16-byte loads/stores for bf16, 2 x 16 bytes for fp32.
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&x)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&x)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
    x[2 * k] = f.x; x[2 * k + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&x)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&x)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(x[0], x[1], x[2], x[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(x[4], x[5], x[6], x[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&x)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_fwd_vec_kernel(const T* __restrict__ s, T* __restrict__ p, long long rows, long long ld) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < rows; r += nwarps) {
    float v[NCH][8];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      ld8<T>(s + r * ld + c * 256 + lane * 8, v[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, v[c][j]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[c][j] = __expf(v[c][j] - m); sum += v[c][j]; }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] *= inv;
      st8<T>(p + r * ld + c * 256 + lane * 8, v[c]);
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_bwd_vec_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ ds,
                       long long rows, long long ld, float scale) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < rows; r += nwarps) {
    float pv[NCH][8], dv[NCH][8];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      ld8<T>(p + r * ld + c * 256 + lane * 8, pv[c]);
      ld8<T>(dp + r * ld + c * 256 + lane * 8, dv[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot += pv[c][j] * dv[c][j];
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dv[c][j] = scale * pv[c][j] * (dv[c][j] - dot);
      st8<T>(ds + r * ld + c * 256 + lane * 8, dv[c]);
    }
  }
}

// Ragged variants: any cols <= NCH * 256 on rows padded to a multiple of 8 elements (a key length
// such as 290 stored with a 296-element row stride).  Same lane -> 8-element-run mapping; runs that
// lie fully inside the row use the 16-byte accesses, the one run that straddles `cols` is handled
// element by element (so nothing outside [0, cols) is read as data or written), later runs are skipped.
template <typename T>
__device__ __forceinline__ void ld8_ragged(const T* row, int col0, int cols, float fill, float (&x)[8]) {
  if (col0 + 8 <= cols) {
    ld8<T>(row + col0, x);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = col0 + j < cols ? ldf<T>(row, col0 + j) : fill;
  }
}
template <typename T>
__device__ __forceinline__ void st8_ragged(T* row, int col0, int cols, const float (&x)[8]) {
  if (col0 + 8 <= cols) {
    st8<T>(row + col0, x);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (col0 + j < cols) stf<T>(row, col0 + j, x[j]);
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_fwd_ragged_kernel(const T* __restrict__ s, T* __restrict__ p, long long rows, int cols,
                          long long ld) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < rows; r += nwarps) {
    float v[NCH][8];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      ld8_ragged<T>(s + r * ld, c * 256 + lane * 8, cols, -INFINITY, v[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, v[c][j]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[c][j] = __expf(v[c][j] - m); sum += v[c][j]; }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] *= inv;
      st8_ragged<T>(p + r * ld, c * 256 + lane * 8, cols, v[c]);
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_bwd_ragged_kernel(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ ds,
                          long long rows, int cols, long long ld, float scale) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < rows; r += nwarps) {
    float pv[NCH][8], dv[NCH][8];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      ld8_ragged<T>(p + r * ld, c * 256 + lane * 8, cols, 0.f, pv[c]);
      ld8_ragged<T>(dp + r * ld, c * 256 + lane * 8, cols, 0.f, dv[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot += pv[c][j] * dv[c][j];
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dv[c][j] = scale * pv[c][j] * (dv[c][j] - dot);
      st8_ragged<T>(ds + r * ld, c * 256 + lane * 8, cols, dv[c]);
    }
  }
}

template <typename T>
static void softmax_fwd_ragged(const void* s, void* p, long long rows, int cols, long long ld, int grid,
                               int th, cudaStream_t st) {
  switch ((cols + 255) / 256) {
    case 1: DL_LAUNCH((softmax_fwd_ragged_kernel<T, 1>), grid, th, 0, st, (const T*)s, (T*)p, rows, cols, ld); break;
    case 2: DL_LAUNCH((softmax_fwd_ragged_kernel<T, 2>), grid, th, 0, st, (const T*)s, (T*)p, rows, cols, ld); break;
    case 3: DL_LAUNCH((softmax_fwd_ragged_kernel<T, 3>), grid, th, 0, st, (const T*)s, (T*)p, rows, cols, ld); break;
    default: DL_LAUNCH((softmax_fwd_ragged_kernel<T, 4>), grid, th, 0, st, (const T*)s, (T*)p, rows, cols, ld); break;
  }
}
template <typename T>
static void softmax_bwd_ragged(const void* p, const void* dp, void* ds, long long rows, int cols,
                               long long ld, float scale, int grid, int th, cudaStream_t st) {
  switch ((cols + 255) / 256) {
    case 1: DL_LAUNCH((softmax_bwd_ragged_kernel<T, 1>), grid, th, 0, st, (const T*)p, (const T*)dp, (T*)ds, rows, cols, ld, scale); break;
    case 2: DL_LAUNCH((softmax_bwd_ragged_kernel<T, 2>), grid, th, 0, st, (const T*)p, (const T*)dp, (T*)ds, rows, cols, ld, scale); break;
    case 3: DL_LAUNCH((softmax_bwd_ragged_kernel<T, 3>), grid, th, 0, st, (const T*)p, (const T*)dp, (T*)ds, rows, cols, ld, scale); break;
    default: DL_LAUNCH((softmax_bwd_ragged_kernel<T, 4>), grid, th, 0, st, (const T*)p, (const T*)dp, (T*)ds, rows, cols, ld, scale); break;
  }
}

// ------------------------------------------------------------------ column sums
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long rows, int cols,
              long long ld, long long rows_per_block) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  if (c < cols)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += ldf<T>(x, r * ld + c);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// Vectorised variant (cols and ld multiples of the 16-byte vector, 16-byte aligned base): `tpr`
// threads cover one row slice with 16-byte loads, 256 / tpr rows per pass, four passes in flight.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, float* __restrict__ out, long long rows, int cols,
                  long long ld, long long rows_per_block, int tpr) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = 16 / sizeof(T);
  __shared__ float red[256][V + 1];
  const int vcol = threadIdx.x % tpr, rsub = threadIdx.x / tpr, rpp = 256 / tpr;
  const int c = (blockIdx.x * tpr + vcol) * V;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  if (c < cols) {
    long long r = r0 + rsub;
    for (; r + 3 * rpp < r1; r += 4 * rpp) {
      float v[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) ldv(x + (r + u * rpp) * ld + c, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += v[u][j];
    }
    for (; r < r1; r += rpp) {
      float v[V];
      ldv(x + r * ld + c, v);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  // four consecutive columns per thread: one 16-byte vector reduction instead of four scalar atomics (every
  // block of a column slice adds to the same addresses, and same-address atomics serialise in L2)
  const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int o = threadIdx.x; o < tpr * V / 4; o += 256) {
    const int vc = (o * 4) / V, j = (o * 4) % V;
    const int col = (blockIdx.x * tpr + vc) * V + j;
    if (col < cols) {                     // cols % V == 0, so the four columns are valid together
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < rpp; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] += red[i * tpr + vc][j + q];
      if (vec_ok) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                     :: "l"(out + col), "f"(t[0]), "f"(t[1]), "f"(t[2]), "f"(t[3]) : "memory");
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(out + col + q, t[q]);
      }
    }
  }
}

// ------------------------------------------------------------------ dropout / cast / add

template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float p,
                               unsigned long long seed, const long long* __restrict__ ctr) {
  pdl_trigger();
  pdl_wait();
  seed = drop_seed_at(seed, ctr);
  const uint32_t thr = drop_threshold(p);
  const float inv = drop_scale(thr);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float keep = drop_keep(seed, (unsigned long long)i, thr) ? inv : 0.f;
    stf<T>(y, i, ldf<T>(x, i) * keep);
  }
}

template <typename T>
__global__ void act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, int act) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = ldf<T>(x, i);
    stf<T>(y, i, act == DL_ACT_GELU ? gelu_fwd<T>(v) : (act == DL_ACT_RELU ? fmaxf(v, 0.f) : v));
  }
}

// F.normalize(x, dim=-1): y = x / max(||x||_2, eps); one warp per row, any cols
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ norm_out,
                  long long rows, int cols, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { const float v = ldf<T>(x, r * cols + c); s += v * v; }
  const float nrm = fmaxf(sqrtf(warp_sum(s)), eps);
  const float inv = 1.f / nrm;
  for (int c = lane; c < cols; c += 32) stf<T>(y, r * cols + c, ldf<T>(x, r * cols + c) * inv);
  if (lane == 0 && norm_out) norm_out[r] = nrm;
}

// dx = (dy - y * <dy, y>) / ||x||
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, const float* __restrict__ norm,
                  T* __restrict__ dx, long long rows, int cols) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= rows) return;
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot += ldf<T>(dy, r * cols + c) * ldf<T>(y, r * cols + c);
  dot = warp_sum(dot);
  const float inv = 1.f / norm[r];
  for (int c = lane; c < cols; c += 32)
    stf<T>(dx, r * cols + c, (ldf<T>(dy, r * cols + c) - ldf<T>(y, r * cols + c) * dot) * inv);
}

// g = dy * act'(pre) * dropout_mask  (backward of the GEMM epilogue's act + dropout)
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ pre,
                               T* __restrict__ g, long long n, int act, float p,
                               unsigned long long seed, const long long* __restrict__ ctr) {
  pdl_trigger();
  pdl_wait();
  seed = drop_seed_at(seed, ctr);
  const uint32_t thr = drop_threshold(p);
  const float inv = p > 0.f ? drop_scale(thr) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = ldf<T>(dy, i);
    if (act == DL_ACT_GELU) v *= gelu_grad<T>(ldf<T>(pre, i));
    else if (act == DL_ACT_RELU) v = ldf<T>(pre, i) > 0.f ? v : 0.f;
    if (p > 0.f) v *= drop_keep(seed, (unsigned long long)i, thr) ? inv : 0.f;
    stf<T>(g, i, v);
  }
}

// 16-byte vector variant (n % V == 0, 16-byte aligned pointers): V = 8 bf16 / 4 fp32 elements per
// thread and step, one dropout hash per two elements.
template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ pre, T* __restrict__ g,
                   long long nvec, int act, float p, unsigned long long seed,
                   const long long* __restrict__ ctr) {
  pdl_trigger();
  pdl_wait();
  seed = drop_seed_at(seed, ctr);
  constexpr int V = VecWidth<T>::N;
  const uint32_t thr = drop_threshold(p);
  const float inv = p > 0.f ? drop_scale(thr) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float v[V];
    ldv(dy + i * V, v);
    if (act != DL_ACT_NONE) {
      float a[V];
      ldv(pre + i * V, a);
#pragma unroll
      for (int j = 0; j < V; ++j)
        v[j] = act == DL_ACT_GELU ? v[j] * gelu_grad<T>(a[j]) : (a[j] > 0.f ? v[j] : 0.f);
    }
    if (p > 0.f) {
      const unsigned long long e2 = (unsigned long long)i * (V / 2);      // i * V is even
#pragma unroll
      for (int k = 0; k < V / 2; ++k) {
        const uint32_t h = drop_hash(seed, e2 + k);
        v[2 * k] *= (h & 0xffffu) >= thr ? inv : 0.f;
        v[2 * k + 1] *= (h >> 16) >= thr ? inv : 0.f;
      }
    }
    stv(g + i * V, v);
  }
}

// AdamW over one flat fp32 buffer (torch.optim.AdamW semantics: decoupled weight decay, bias
// correction from the device-side step counter) that also refreshes the bf16 shadow the GEMMs read.
__global__ void adamw_tick_kernel(long long* step) {
  pdl_trigger();
  pdl_wait(); *step += 1; }

// per-parameter step counts (torch keeps state['step'] per parameter and only advances it when the
// parameter has a gradient): one counter per 64-element block of the flat buffer
__global__ void adamw_tick_blocks_kernel(int* __restrict__ steps, const unsigned char* __restrict__ active,
                                         long long nblocks) {
  pdl_trigger();
  pdl_wait();
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks;
       b += (long long)gridDim.x * blockDim.x)
    if (active == nullptr || active[b]) steps[b] += 1;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                             float* __restrict__ m, float* __restrict__ v,
                             __nv_bfloat16* __restrict__ shadow, long long n,
                             const long long* __restrict__ step, float lr, float b1, float b2,
                             float eps, float wd, float grad_scale,
                             const unsigned char* __restrict__ active, const int* __restrict__ step_blocks,
                             int step_offset) {
  pdl_trigger();
  pdl_wait();
  const float t = (float)(*step + step_offset);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    // parameters that received no gradient this step are skipped entirely (no decay, moments
    // untouched), like torch.optim.AdamW skips p.grad is None
    if (active != nullptr && active[i >> 6] == 0) continue;
    if (step_blocks != nullptr) {            // this parameter's own step count
      const float tb = (float)step_blocks[i >> 6];
      step_size = lr / (1.f - powf(b1, tb));
      inv_sqrt_bc2 = rsqrtf(1.f - powf(b2, tb));
    }
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

// dst[r, 0:bytes) = src[r, 0:bytes) for rows with independent pitches: the column halves of a
// concatenation (forward) and of its gradient (backward), 16 bytes per thread per step
__global__ void copy_rows_kernel(const uint4* __restrict__ src, long long src_pitch16, uint4* __restrict__ dst,
                                 long long dst_pitch16, long long rows, int row16) {
  pdl_trigger();
  pdl_wait();
  const long long n = rows * row16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row16;
    const int c = (int)(i - r * row16);
    dst[r * dst_pitch16 + c] = src[r * src_pitch16 + c];
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  pdl_trigger();
  pdl_wait();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x)
    st4<TO>(y + i * 4, ld4<TI>(x + i * 4));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    stf<TO>(y, i, ldf<TI>(x, i));
  }
}

// y[b, i] = (x[b, i] + pe[i]) * dropout
template <typename T>
__global__ void add_pe_kernel(const T* __restrict__ x, const float* __restrict__ pe,
                              T* __restrict__ y, long long n, long long period, float p,
                              unsigned long long seed, const long long* __restrict__ ctr) {
  pdl_trigger();
  pdl_wait();
  seed = drop_seed_at(seed, ctr);
  const uint32_t thr = drop_threshold(p);
  const float inv = p > 0.f ? drop_scale(thr) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = ldf<T>(x, i) + pe[i % period];
    if (p > 0.f) v *= drop_keep(seed, (unsigned long long)i, thr) ? inv : 0.f;
    stf<T>(y, i, v);
  }
}

// 16-byte vector variant (n, period multiples of V; aligned pointers)
template <typename T>
__global__ void __launch_bounds__(256)
add_pe_vec_kernel(const T* __restrict__ x, const float* __restrict__ pe, T* __restrict__ y,
                  long long nvec, long long period, float p, unsigned long long seed,
                  const long long* __restrict__ ctr) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = VecWidth<T>::N;
  seed = drop_seed_at(seed, ctr);
  const uint32_t thr = drop_threshold(p);
  const float inv = p > 0.f ? drop_scale(thr) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float v[V];
    ldv(x + i * V, v);
    const float* q = pe + (i * V) % period;
#pragma unroll
    for (int j = 0; j < V; j += 4) {
      const float4 e = *reinterpret_cast<const float4*>(q + j);
      v[j] += e.x; v[j + 1] += e.y; v[j + 2] += e.z; v[j + 3] += e.w;
    }
    if (p > 0.f) {
      const unsigned long long e2 = (unsigned long long)i * (V / 2);
#pragma unroll
      for (int k = 0; k < V / 2; ++k) {
        const uint32_t h = drop_hash(seed, e2 + k);
        v[2 * k] *= (h & 0xffffu) >= thr ? inv : 0.f;
        v[2 * k + 1] *= (h >> 16) >= thr ? inv : 0.f;
      }
    }
    stv(y + i * V, v);
  }
}

int ew_grid(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

template <typename T>
int ln_fwd_dispatch(const void* x, const float* g, const float* b, void* y, float* mean,
                    float* rstd, long long rows, int cols, float eps, cudaStream_t st) {
  const int grid = row_grid(rows), th = kWarpsPerBlock * 32;
  const T* xx = (const T*)x;
  T* yy = (T*)y;
  switch (cols / 128) {
    case 1: DL_LAUNCH((layernorm_fwd_kernel<T, 128>), grid, th, 0, st, xx, g, b, yy, mean, rstd, rows, eps); break;
    case 2: DL_LAUNCH((layernorm_fwd_kernel<T, 256>), grid, th, 0, st, xx, g, b, yy, mean, rstd, rows, eps); break;
    case 4: DL_LAUNCH((layernorm_fwd_kernel<T, 512>), grid, th, 0, st, xx, g, b, yy, mean, rstd, rows, eps); break;
    case 8: DL_LAUNCH((layernorm_fwd_kernel<T, 1024>), grid, th, 0, st, xx, g, b, yy, mean, rstd, rows, eps); break;
    default: return set_error(-1, "dl_layernorm_fwd: cols must be 128, 256, 512 or 1024 (got %d)", cols);
  }
  DL_LAUNCH_CHECK("layernorm_fwd_kernel");
  count_launch();
  return 0;
}

// cluster size for the column-partial reduction: 8 blocks share one set of atomics (1 = no cluster)
constexpr int kLnBwdCluster = 8;

template <typename T, int C>
int ln_bwd_launch(const T* dy, const T* x, const float* g, const float* mean, const float* rstd, T* dx, const T* add,
                  float* dg, float* db, long long rows, cudaStream_t st) {
  // two blocks per SM (as many as fit with 8 x C floats of shared memory and ~100 registers), a whole number of
  // clusters; rows are strided over all warps
  int grid = row_grid(rows);
  if (grid > sm_count() * 2) grid = sm_count() * 2;
  const int th = kWarpsPerBlock * 32;
  // Measured on B200 (tools/ln_bench.py, 16384 x 256 bf16): one atomicAdd per column per block 12.5 us; 16-byte
  // vector reductions per block 7.0 us; the same per 8-block cluster 11.8 us (cluster scheduling + two cluster
  // barriers cost more than the 8x fewer reductions save) -- so clusters are opt-in (DL_LN_CLUSTER=1).
  static const bool use_cluster = [] { const char* e = getenv("DL_LN_CLUSTER"); return e && atoi(e) != 0; }();
  if (use_cluster && grid >= kLnBwdCluster && (dg || db)) {
    grid -= grid % kLnBwdCluster;
    (void)launch_cluster_k(layernorm_bwd_kernel<T, C, kLnBwdCluster>, dim3(grid), dim3(th), (size_t)0, (unsigned)kLnBwdCluster,
                           st, dy, x, g, mean, rstd, dx, add, dg, db, rows);
  } else {
    DL_LAUNCH((layernorm_bwd_kernel<T, C, 1>), grid, th, 0, st, dy, x, g, mean, rstd, dx, add, dg, db, rows);
  }
  return 0;
}

template <typename T>
int ln_bwd_dispatch(const void* dy, const void* x, const float* g, const float* mean,
                    const float* rstd, void* dx, const void* dx_add, float* dg, float* db, long long rows,
                    int cols, cudaStream_t st) {
  const T *dyy = (const T*)dy, *xx = (const T*)x;
  T* dxx = (T*)dx;
  const T* add = (const T*)dx_add;
  switch (cols / 128) {
    case 1: ln_bwd_launch<T, 128>(dyy, xx, g, mean, rstd, dxx, add, dg, db, rows, st); break;
    case 2: ln_bwd_launch<T, 256>(dyy, xx, g, mean, rstd, dxx, add, dg, db, rows, st); break;
    case 4: ln_bwd_launch<T, 512>(dyy, xx, g, mean, rstd, dxx, add, dg, db, rows, st); break;
    case 8: ln_bwd_launch<T, 1024>(dyy, xx, g, mean, rstd, dxx, add, dg, db, rows, st); break;
    default: return set_error(-1, "dl_layernorm_bwd: cols must be 128, 256, 512 or 1024 (got %d)", cols);
  }
  DL_LAUNCH_CHECK("layernorm_bwd_kernel");
  count_launch();
  return 0;
}

}  // namespace
}  // namespace dl

using namespace dl;

extern "C" int dl_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                float* mean, float* rstd, int64_t rows, int32_t cols, float eps,
                                int32_t dtype, void* stream) {
  DL_REQUIRE(x && gamma && beta && y, "dl_layernorm_fwd: null pointer");
  DL_REQUIRE(rows >= 0 && cols > 0 && cols % 128 == 0, "dl_layernorm_fwd: bad shape rows=%lld cols=%d", (long long)rows, cols);
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == DL_BF16 ? ln_fwd_dispatch<__nv_bfloat16>(x, gamma, beta, y, mean, rstd, rows, cols, eps, st)
                          : ln_fwd_dispatch<float>(x, gamma, beta, y, mean, rstd, rows, cols, eps, st);
}

extern "C" int dl_layernorm_bwd(const void* dy, const void* x, const float* gamma,
                                const float* mean, const float* rstd, void* dx, const void* dx_add,
                                float* dgamma, float* dbeta, int64_t rows, int32_t cols,
                                int32_t accumulate, int32_t dtype, void* stream) {
  DL_REQUIRE(dy && x && gamma && mean && rstd && dx, "dl_layernorm_bwd: null pointer");
  DL_REQUIRE(rows >= 0 && cols > 0 && cols % 128 == 0, "dl_layernorm_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (dgamma && !accumulate) DL_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * cols, st));
  if (dbeta && !accumulate) DL_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * cols, st));
  if (rows == 0) return 0;
  return dtype == DL_BF16 ? ln_bwd_dispatch<__nv_bfloat16>(dy, x, gamma, mean, rstd, dx, dx_add, dgamma, dbeta, rows, cols, st)
                          : ln_bwd_dispatch<float>(dy, x, gamma, mean, rstd, dx, dx_add, dgamma, dbeta, rows, cols, st);
}

extern "C" int dl_softmax_fwd(const void* s, void* p, int64_t rows, int32_t cols, int64_t ld,
                              int32_t dtype, void* stream) {
  DL_REQUIRE(s && p, "dl_softmax_fwd: null pointer");
  DL_REQUIRE(cols > 0 && cols <= 32 * kMaxPerLane && ld >= cols, "dl_softmax_fwd: cols must be in [1, %d]", 32 * kMaxPerLane);
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = row_grid(rows), th = kWarpsPerBlock * 32;
  const bool al = ((uintptr_t)s & 15) == 0 && ((uintptr_t)p & 15) == 0 && ld % 8 == 0;
  if (al && (cols == 256 || cols == 512)) {
    if (dtype == DL_BF16) {
      if (cols == 256) DL_LAUNCH((softmax_fwd_vec_kernel<__nv_bfloat16, 1>), grid, th, 0, st, (const __nv_bfloat16*)s, (__nv_bfloat16*)p, rows, ld);
      else DL_LAUNCH((softmax_fwd_vec_kernel<__nv_bfloat16, 2>), grid, th, 0, st, (const __nv_bfloat16*)s, (__nv_bfloat16*)p, rows, ld);
    } else {
      if (cols == 256) DL_LAUNCH((softmax_fwd_vec_kernel<float, 1>), grid, th, 0, st, (const float*)s, (float*)p, rows, ld);
      else DL_LAUNCH((softmax_fwd_vec_kernel<float, 2>), grid, th, 0, st, (const float*)s, (float*)p, rows, ld);
    }
  } else if (al && cols >= 64 && cols <= 1024) {
    if (dtype == DL_BF16) softmax_fwd_ragged<__nv_bfloat16>(s, p, rows, cols, ld, grid, th, st);
    else softmax_fwd_ragged<float>(s, p, rows, cols, ld, grid, th, st);
  } else if (dtype == DL_BF16)
    DL_LAUNCH((softmax_fwd_kernel<__nv_bfloat16>), grid, th, 0, st, (const __nv_bfloat16*)s, (__nv_bfloat16*)p, rows, cols, ld);
  else
    DL_LAUNCH((softmax_fwd_kernel<float>), grid, th, 0, st, (const float*)s, (float*)p, rows, cols, ld);
  DL_LAUNCH_CHECK("softmax_fwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_softmax_bwd(const void* p, const void* dp, void* ds, int64_t rows, int32_t cols,
                              int64_t ld, float scale, int32_t dtype, void* stream) {
  DL_REQUIRE(p && dp && ds, "dl_softmax_bwd: null pointer");
  DL_REQUIRE(cols > 0 && cols <= 32 * kMaxPerLane && ld >= cols, "dl_softmax_bwd: cols must be in [1, %d]", 32 * kMaxPerLane);
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = row_grid(rows), th = kWarpsPerBlock * 32;
  const bool al = ((uintptr_t)p & 15) == 0 && ((uintptr_t)dp & 15) == 0 && ((uintptr_t)ds & 15) == 0 && ld % 8 == 0;
  if (al && (cols == 256 || cols == 512)) {
    if (dtype == DL_BF16) {
      if (cols == 256) DL_LAUNCH((softmax_bwd_vec_kernel<__nv_bfloat16, 1>), grid, th, 0, st, (const __nv_bfloat16*)p, (const __nv_bfloat16*)dp, (__nv_bfloat16*)ds, rows, ld, scale);
      else DL_LAUNCH((softmax_bwd_vec_kernel<__nv_bfloat16, 2>), grid, th, 0, st, (const __nv_bfloat16*)p, (const __nv_bfloat16*)dp, (__nv_bfloat16*)ds, rows, ld, scale);
    } else {
      if (cols == 256) DL_LAUNCH((softmax_bwd_vec_kernel<float, 1>), grid, th, 0, st, (const float*)p, (const float*)dp, (float*)ds, rows, ld, scale);
      else DL_LAUNCH((softmax_bwd_vec_kernel<float, 2>), grid, th, 0, st, (const float*)p, (const float*)dp, (float*)ds, rows, ld, scale);
    }
  } else if (al && cols >= 64 && cols <= 1024) {
    if (dtype == DL_BF16) softmax_bwd_ragged<__nv_bfloat16>(p, dp, ds, rows, cols, ld, scale, grid, th, st);
    else softmax_bwd_ragged<float>(p, dp, ds, rows, cols, ld, scale, grid, th, st);
  } else if (dtype == DL_BF16)
    DL_LAUNCH((softmax_bwd_kernel<__nv_bfloat16>), grid, th, 0, st, (const __nv_bfloat16*)p, (const __nv_bfloat16*)dp, (__nv_bfloat16*)ds, rows, cols, ld, scale);
  else
    DL_LAUNCH((softmax_bwd_kernel<float>), grid, th, 0, st, (const float*)p, (const float*)dp, (float*)ds, rows, cols, ld, scale);
  DL_LAUNCH_CHECK("softmax_bwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_colsum(const void* x, float* out, int64_t rows, int32_t cols, int64_t ld,
                         int32_t accumulate, int32_t dtype, void* stream) {
  DL_REQUIRE(x && out && cols > 0 && ld >= cols, "dl_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) DL_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  if (rows <= 0) return 0;
  const int vec = dtype == DL_BF16 ? 8 : 4;
  if (cols % vec == 0 && ld % vec == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    int tpr = 1;
    while (tpr < 256 && tpr * vec < cols) tpr <<= 1;
    const int xb = ceil_div(cols, tpr * vec), rpp = 256 / tpr;
    long long yb = (long long)sm_count() * 2 / xb;
    if (yb < 1) yb = 1;
    if (yb > (rows + 4 * rpp - 1) / (4 * rpp)) yb = (rows + 4 * rpp - 1) / (4 * rpp);
    const long long rpb = (rows + yb - 1) / yb;
    dim3 grid(xb, (unsigned)yb);
    if (dtype == DL_BF16)
      DL_LAUNCH((colsum_vec_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, out, rows, cols, ld, rpb, tpr);
    else
      DL_LAUNCH((colsum_vec_kernel<float>), grid, 256, 0, st, (const float*)x, out, rows, cols, ld, rpb, tpr);
    DL_LAUNCH_CHECK("colsum_vec_kernel");
    count_launch();
    return 0;
  }
  const int xb = ceil_div(cols, 32);
  long long yb = (long long)sm_count() * 4 / xb;
  if (yb < 1) yb = 1;
  if (yb > (rows + 63) / 64) yb = (rows + 63) / 64;
  const long long rpb = (rows + yb - 1) / yb;
  dim3 grid(xb, (unsigned)yb), block(32, 8);
  if (dtype == DL_BF16)
    DL_LAUNCH((colsum_kernel<__nv_bfloat16>), grid, block, 0, st, (const __nv_bfloat16*)x, out, rows, cols, ld, rpb);
  else
    DL_LAUNCH((colsum_kernel<float>), grid, block, 0, st, (const float*)x, out, rows, cols, ld, rpb);
  DL_LAUNCH_CHECK("colsum_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed,
                          const int64_t* seed_step, int32_t dtype,
                          void* stream) {
  DL_REQUIRE(x && y && p >= 0.f && p < 1.f, "dl_dropout: bad arguments");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(n, 256);
  if (dtype == DL_BF16)
    DL_LAUNCH((dropout_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, p, seed, (const long long*)seed_step);
  else
    DL_LAUNCH((dropout_kernel<float>), grid, 256, 0, st, (const float*)x, (float*)y, n, p, seed, (const long long*)seed_step);
  DL_LAUNCH_CHECK("dropout_kernel");
  count_launch();
  return 0;
}

// out[m, n] = aux[m, n] * sum_{k < K} g[m, k] * w[k, n]   (bf16, K <= 16): the input gradient of a layer with a
// handful of outputs -- MHLA's lin2 (model/PMMA/encoder.py:130: 1024 -> 8 heads) -- times the stored
// activation derivative.  As a K = 8 tensor-core GEMM it was a 64-wide K-block of zeros behind a TMA epilogue
// (24 us for 67 MB of aux + out); here each thread keeps its 8 columns of w in registers and streams rows.
template <int KMAX>
__global__ void __launch_bounds__(128)
smallk_mul_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ w,
                  const __nv_bfloat16* __restrict__ aux, __nv_bfloat16* __restrict__ out, long long M, int N, int K,
                  long long ldg, long long rows_per_block) {
  pdl_trigger();
  pdl_wait();
  const int nv = blockIdx.x * 128 + threadIdx.x;          // 8-column vector of the row
  if (nv * 8 >= N) return;
  float wr[KMAX][8];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) ldv(w + (long long)k * N + nv * 8, wr[k]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) wr[k][j] = 0.f;
    }
  }
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  for (long long r = r0; r < r1; r += 2) {
    float a[2][8], gv[2][KMAX];
    const bool two = r + 1 < r1;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        ldv(aux + (r + u) * N + nv * 8, a[u]);
#pragma unroll
        for (int k0 = 0; k0 < KMAX; k0 += 8) ldv(g + (r + u) * ldg + k0, *reinterpret_cast<float(*)[8]>(&gv[u][k0]));
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k >= K) gv[u][k] = 0.f;             // padding columns of g may hold anything
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(gv[u][k], wr[k][j], o[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= a[u][j];
      stv(out + (r + u) * N + nv * 8, o);
    }
  }
}

extern "C" int dl_smallk_mul(const void* g, const void* w, const void* aux, void* out, int64_t M, int32_t N, int32_t K,
                             int64_t ldg, void* stream) {
  DL_REQUIRE(g && w && aux && out && M >= 0 && N >= 8 && N % 8 == 0 && K >= 1 && K <= 16, "dl_smallk_mul: bad shape (N %% 8 == 0, K <= 16)");
  DL_REQUIRE(ldg % 8 == 0 && ldg >= (K <= 8 ? 8 : 16), "dl_smallk_mul: g rows must be whole 16-byte vectors (ldg %% 8 == 0, padded to 8 / 16 columns)");
  DL_REQUIRE((((uintptr_t)g | (uintptr_t)w | (uintptr_t)aux | (uintptr_t)out) & 15) == 0, "dl_smallk_mul: tensors must be 16-byte aligned");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int xb = ceil_div(N / 8, 128);
  long long yb = (long long)sm_count() * 8 / xb;
  if (yb > (M + 7) / 8) yb = (M + 7) / 8;
  if (yb < 1) yb = 1;
  const long long rpb = (M + yb - 1) / yb;
  dim3 grid(xb, (unsigned)yb);
  if (K <= 8)
    DL_LAUNCH((smallk_mul_kernel<8>), grid, 128, 0, st, (const __nv_bfloat16*)g, (const __nv_bfloat16*)w, (const __nv_bfloat16*)aux, (__nv_bfloat16*)out, (long long)M, N, K, (long long)ldg, rpb);
  else
    DL_LAUNCH((smallk_mul_kernel<16>), grid, 128, 0, st, (const __nv_bfloat16*)g, (const __nv_bfloat16*)w, (const __nv_bfloat16*)aux, (__nv_bfloat16*)out, (long long)M, N, K, (long long)ldg, rpb);
  DL_LAUNCH_CHECK("smallk_mul_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_act_fwd(const void* x, void* y, int64_t n, int32_t act, int32_t dtype, void* stream) {
  DL_REQUIRE(x && y && act >= 0 && act <= 2, "dl_act_fwd: bad arguments");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(n, 256);
  if (dtype == DL_BF16)
    DL_LAUNCH((act_fwd_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, act);
  else
    DL_LAUNCH((act_fwd_kernel<float>), grid, 256, 0, st, (const float*)x, (float*)y, n, act);
  DL_LAUNCH_CHECK("act_fwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_l2norm_fwd(const void* x, void* y, float* norm, int64_t rows, int32_t cols,
                             float eps, int32_t dtype, void* stream) {
  DL_REQUIRE(x && y && norm && cols > 0, "dl_l2norm_fwd: bad arguments");
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(rows, kWarpsPerBlock), th = kWarpsPerBlock * 32;
  if (dtype == DL_BF16)
    DL_LAUNCH((l2norm_fwd_kernel<__nv_bfloat16>), grid, th, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, norm, rows, cols, eps);
  else
    DL_LAUNCH((l2norm_fwd_kernel<float>), grid, th, 0, st, (const float*)x, (float*)y, norm, rows, cols, eps);
  DL_LAUNCH_CHECK("l2norm_fwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_l2norm_bwd(const void* dy, const void* y, const float* norm, void* dx,
                             int64_t rows, int32_t cols, int32_t dtype, void* stream) {
  DL_REQUIRE(dy && y && norm && dx && cols > 0, "dl_l2norm_bwd: bad arguments");
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(rows, kWarpsPerBlock), th = kWarpsPerBlock * 32;
  if (dtype == DL_BF16)
    DL_LAUNCH((l2norm_bwd_kernel<__nv_bfloat16>), grid, th, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, norm, (__nv_bfloat16*)dx, rows, cols);
  else
    DL_LAUNCH((l2norm_bwd_kernel<float>), grid, th, 0, st, (const float*)dy, (const float*)y, norm, (float*)dx, rows, cols);
  DL_LAUNCH_CHECK("l2norm_bwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_act_bwd(const void* dy, const void* pre, void* g, int64_t n, int32_t act, float p,
                          uint64_t seed, const int64_t* seed_step, int32_t dtype, void* stream) {
  DL_REQUIRE(dy && g && (act == DL_ACT_NONE || pre) && p >= 0.f && p < 1.f, "dl_act_bwd: bad arguments");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = dtype == DL_BF16 ? 8 : 4;
  if (n % vec == 0 && (((uintptr_t)dy | (uintptr_t)pre | (uintptr_t)g) & 15) == 0) {
    const long long nvec = n / vec;
    const int vgrid = ew_grid(nvec, 256);
    if (dtype == DL_BF16)
      DL_LAUNCH((act_bwd_vec_kernel<__nv_bfloat16>), vgrid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)pre, (__nv_bfloat16*)g, nvec, act, p, seed, (const long long*)seed_step);
    else
      DL_LAUNCH((act_bwd_vec_kernel<float>), vgrid, 256, 0, st, (const float*)dy, (const float*)pre, (float*)g, nvec, act, p, seed, (const long long*)seed_step);
    DL_LAUNCH_CHECK("act_bwd_vec_kernel");
    count_launch();
    return 0;
  }
  const int grid = ew_grid(n, 256);
  if (dtype == DL_BF16)
    DL_LAUNCH((act_bwd_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)pre, (__nv_bfloat16*)g, n, act, p, seed, (const long long*)seed_step);
  else
    DL_LAUNCH((act_bwd_kernel<float>), grid, 256, 0, st, (const float*)dy, (const float*)pre, (float*)g, n, act, p, seed, (const long long*)seed_step);
  DL_LAUNCH_CHECK("act_bwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             void* shadow_bf16, int64_t n, int64_t* step, float lr, float beta1,
                             float beta2, float eps, float weight_decay, float grad_scale,
                             const uint8_t* active_blocks, int32_t* step_blocks, int32_t tick,
                             void* stream) {
  DL_REQUIRE(param && grad && exp_avg && exp_avg_sq && step, "dl_adamw_step: null pointer");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (tick) {
    DL_LAUNCH(adamw_tick_kernel, 1, 1, 0, st, (long long*)step);
    DL_LAUNCH_CHECK("adamw_tick_kernel");
    count_launch();
  }
  if (step_blocks) {
    const long long nb = (n + 63) / 64;
    DL_LAUNCH(adamw_tick_blocks_kernel, ew_grid(nb, 256), 256, 0, st, (int*)step_blocks,
              (const unsigned char*)active_blocks, nb);
    DL_LAUNCH_CHECK("adamw_tick_blocks_kernel");
    count_launch();
  }
  DL_LAUNCH(adamw_kernel, ew_grid(n, 256), 256, 0, st, param, grad, exp_avg, exp_avg_sq,
                                                (__nv_bfloat16*)shadow_bf16, n, (const long long*)step,
                                                lr, beta1, beta2, eps, weight_decay, grad_scale,
                                                (const unsigned char*)active_blocks, (const int*)step_blocks,
                                                tick ? 0 : 1);
  DL_LAUNCH_CHECK("adamw_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_copy_rows(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch,
                            int64_t row_bytes, int64_t rows, void* stream) {
  DL_REQUIRE(src && dst, "dl_copy_rows: null pointer");
  DL_REQUIRE(row_bytes > 0 && rows >= 0 && src_pitch >= row_bytes && dst_pitch >= row_bytes, "dl_copy_rows: bad extents");
  DL_REQUIRE((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)src_pitch | (uintptr_t)dst_pitch | (uintptr_t)row_bytes) & 15) == 0,
             "dl_copy_rows: pointers, pitches and the row length must be multiples of 16 bytes");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = rows * (row_bytes / 16);
  DL_LAUNCH(copy_rows_kernel, ew_grid(n, 256), 256, 0, st, (const uint4*)src, (long long)(src_pitch / 16), (uint4*)dst,
            (long long)(dst_pitch / 16), (long long)rows, (int)(row_bytes / 16));
  DL_LAUNCH_CHECK("copy_rows_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_cast(const void* x, int32_t dtype_in, void* y, int32_t dtype_out, int64_t n,
                       void* stream) {
  DL_REQUIRE(x && y, "dl_cast: null pointer");
  DL_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "dl_cast: pointers must be 16-byte aligned");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid((n + 3) / 4, 256);
  if (dtype_in == DL_F32 && dtype_out == DL_BF16)
    DL_LAUNCH((cast_kernel<float, __nv_bfloat16>), grid, 256, 0, st, (const float*)x, (__nv_bfloat16*)y, n);
  else if (dtype_in == DL_BF16 && dtype_out == DL_F32)
    DL_LAUNCH((cast_kernel<__nv_bfloat16, float>), grid, 256, 0, st, (const __nv_bfloat16*)x, (float*)y, n);
  else if (dtype_in == DL_F32 && dtype_out == DL_F32)
    DL_LAUNCH((cast_kernel<float, float>), grid, 256, 0, st, (const float*)x, (float*)y, n);
  else
    DL_LAUNCH((cast_kernel<__nv_bfloat16, __nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
  DL_LAUNCH_CHECK("cast_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_add_pe(const void* x, const float* pe, void* y, int64_t n, int64_t period,
                         float p, uint64_t seed, const int64_t* seed_step, int32_t dtype, void* stream) {
  DL_REQUIRE(x && pe && y && period > 0 && p >= 0.f && p < 1.f, "dl_add_pe: bad arguments");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = dtype == DL_BF16 ? 8 : 4;
  if (n % vec == 0 && period % vec == 0 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)pe) & 15) == 0) {
    const long long nvec = n / vec;
    const int vgrid = ew_grid(nvec, 256);
    if (dtype == DL_BF16)
      DL_LAUNCH((add_pe_vec_kernel<__nv_bfloat16>), vgrid, 256, 0, st, (const __nv_bfloat16*)x, pe, (__nv_bfloat16*)y, nvec, period, p, seed, (const long long*)seed_step);
    else
      DL_LAUNCH((add_pe_vec_kernel<float>), vgrid, 256, 0, st, (const float*)x, pe, (float*)y, nvec, period, p, seed, (const long long*)seed_step);
    DL_LAUNCH_CHECK("add_pe_vec_kernel");
    count_launch();
    return 0;
  }
  const int grid = ew_grid(n, 256);
  if (dtype == DL_BF16)
    DL_LAUNCH((add_pe_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, pe, (__nv_bfloat16*)y, n, period, p, seed, (const long long*)seed_step);
  else
    DL_LAUNCH((add_pe_kernel<float>), grid, 256, 0, st, (const float*)x, pe, (float*)y, n, period, p, seed, (const long long*)seed_step);
  DL_LAUNCH_CHECK("add_pe_kernel");
  count_launch();
  return 0;
}
