// Model-glue kernels of the DrugLAMP hot path that are not GEMMs:
//   * dl_fillbit_pool   -- fill-bit mask (x.sum(-1)==0), concat and 9-way site mean in ONE pass
//                          over the LLM embeddings (reference model/DrugLAMP.py:11-19,39-40)
//   * dl_site_pool_*    -- view(B,9,256,C).mean(1) of the CNN output and its backward (:35-37)
//   * dl_mhla_gate_ln_* -- MHLA's softmax-over-sequence gating through the memory-reinterpreting
//                          .view (model/PMMA/encoder.py:127-140) fused with the
//                          residual add and the following LayerNorm (model/DrugLAMP.py:63-71)
//   * dl_cm_triplet_*   -- dense form of ccpp_p_tri_loss (model/cross_modality.py:15-47)
//   * dl_bce_*          -- sigmoid + BCELoss (model/basic_model.py:17-22)
#include "../../include/druglamp_sm100.h"
#include "common.cuh"

namespace dl {
void count_launch(int n = 1);
namespace {

// ------------------------------------------------------------------ fill bit + concat + site mean
// x: (B, S*L, C) fp32.  One WARP per pooled row (b, j): it streams the row's S site rows (lane =
// float4 chunks lane, lane+32, ...; NK chunks per lane, all of a site row's loads in flight at once),
// sums each with shuffles (the fill bit is "row sums to exactly 0", DrugLAMP.py:11-19) and keeps the
// running site mean in registers: no shared-memory reduction, no block barrier.  For the concat
// output the row bounces through a per-warp shared-memory line so the (C+1)-strided rows are
// written with consecutive lanes on consecutive floats.  pooled rows have leading dimension
// ldp >= C+1; columns C+1..ldp-1 are zero (TMA-aligned rows for the GEMM that consumes them).
template <typename TO, int NK>
__global__ void __launch_bounds__(256)
fillbit_pool_kernel(const float* __restrict__ x, float* __restrict__ bit_out,
                    float* __restrict__ cat_out, TO* __restrict__ pooled, int S, int L, int C, int ldp,
                    long long n_rows) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float fb_smem[];                 // [8 warps][C] concat staging
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long pr = (long long)blockIdx.x * 8 + w;   // pooled row index b * L + j
  if (pr >= n_rows) return;
  const int b = (int)(pr / L), j = (int)(pr - (long long)b * L);
  const int nchunk = C >> 2;
  float* rowbuf = fb_smem + w * C;
  float4 acc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float bits = 0.f;
#pragma unroll 2
  for (int s = 0; s < S; ++s) {
    const size_t t = (size_t)b * S * L + (size_t)s * L + j;
    const float* row = x + t * C;
    float4 v[NK];
    float psum = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int ch = lane + 32 * k;
      v[k] = ch < nchunk ? *reinterpret_cast<const float4*>(row + ch * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      psum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      acc[k].x += v[k].x; acc[k].y += v[k].y; acc[k].z += v[k].z; acc[k].w += v[k].w;
    }
    const float bit = warp_sum(psum) == 0.f ? 1.f : 0.f;
    bits += bit;
    if (lane == 0 && bit_out) bit_out[t] = bit;
    if (cat_out) {
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int ch = lane + 32 * k;
        if (ch < nchunk) *reinterpret_cast<float4*>(rowbuf + ch * 4) = v[k];
      }
      __syncwarp();
      float* crow = cat_out + t * (C + 1);
      for (int c = lane; c < C; c += 32) crow[c] = rowbuf[c];
      if (lane == 0) crow[C] = bit;
      __syncwarp();
    }
  }
  if (pooled) {
    const float inv = 1.f / (float)S;
    TO* prow = pooled + (size_t)pr * ldp;
    // ldp % 4 == 0 and 8-byte (bf16) / 16-byte (fp32) aligned rows: vector stores for the C columns
    const bool vec = (ldp & 3) == 0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int ch = lane + 32 * k;
      if (ch < nchunk) {
        const float4 o = make_float4(acc[k].x * inv, acc[k].y * inv, acc[k].z * inv, acc[k].w * inv);
        if (vec) {
          st4<TO>(prow + ch * 4, o);
        } else {
          stf<TO>(prow, ch * 4 + 0, o.x); stf<TO>(prow, ch * 4 + 1, o.y);
          stf<TO>(prow, ch * 4 + 2, o.z); stf<TO>(prow, ch * 4 + 3, o.w);
        }
      }
    }
    for (int c = C + lane; c < ldp; c += 32) stf<TO>(prow, c, c == C ? bits * inv : 0.f);
  }
}

// ------------------------------------------------------------------ site pooling
template <typename T>
__global__ void site_pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int S, long long LC4,
                                     long long n4, long long ldy4, long long C4) {
  pdl_trigger();
  pdl_wait();
  // y[b, j, c] = mean_s x[b, s*L + j, c];  index in float4 units; y row stride ldy (elements)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / LC4, r = i % LC4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
      const float4 v = ld4<T>(x + ((b * S + s) * LC4 + r) * 4);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const float inv = 1.f / (float)S;
    const long long row = (b * LC4 + r) / C4, c4 = r % C4;
    st4<T>(y + (row * ldy4 + c4) * 4, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
  }
}

template <typename T>
__global__ void site_pool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int S,
                                     long long LC4, long long n4, long long ldy4, long long C4) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / LC4, r = i % LC4;
    const long long row = (b * LC4 + r) / C4, c4 = r % C4;
    float4 g = ld4<T>(dy + (row * ldy4 + c4) * 4);
    const float inv = 1.f / (float)S;
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    for (int s = 0; s < S; ++s) st4<T>(dx + ((b * S + s) * LC4 + r) * 4, g);
  }
}

// ------------------------------------------------------------------ MHLA gate + residual + LN
//   p[h, l]  = softmax over l of logits[b, l, h]
//   u[b,l,e] = v[b,l,e] * (1 + p[hh, ll]),  f = l*E + e, hh = f / (L*hd), ll = (f % (L*hd)) / hd
//   y        = LayerNorm_E(u) * gamma + beta
// The reinterpreting view gives head hh exactly the rows [hh*L/H, (hh+1)*L/H) when H divides L, so
// the work of one batch element splits into H independent blocks (grid (B, H)): block (b, hh)
// needs only head hh's softmax and owns all of its gates, forward and backward.  Otherwise one
// block per batch element (grid (B, 1)) handles every head.  256 threads = 8 warps per block.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
mhla_gate_ln_fwd_kernel(const T* __restrict__ v, const T* __restrict__ logits,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        T* __restrict__ y, float* __restrict__ p_out, float* __restrict__ mean_out,
                        float* __restrict__ rstd_out, int L, int H, float eps) {
  pdl_trigger();
  pdl_wait();
  constexpr int E = VEC * 128;
  extern __shared__ float sp[];                 // [H][L]
  const bool plain = gamma == nullptr;
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int hd = E / H;
  const bool sliced = gridDim.y > 1;
  const int h_begin = sliced ? blockIdx.y : 0, h_end = sliced ? blockIdx.y + 1 : H;
  const int l_begin = sliced ? blockIdx.y * (L / H) : 0, l_end = sliced ? l_begin + L / H : L;
  // phase 1: column softmax, warp per head
  for (int h = h_begin + w; h < h_end; h += 8) {
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) m = fmaxf(m, ldf<T>(logits, ((size_t)b * L + l) * H + h));
    m = warp_max(m);
    float s = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float e = __expf(ldf<T>(logits, ((size_t)b * L + l) * H + h) - m);
      sp[h * L + l] = e;
      s += e;
    }
    const float inv = 1.f / warp_sum(s);
    for (int l = lane; l < L; l += 32) {
      const float pv = sp[h * L + l] * inv;
      sp[h * L + l] = pv;
      if (p_out) p_out[((size_t)b * H + h) * L + l] = pv;
    }
  }
  __syncthreads();
  // phase 2: gate + residual + LayerNorm, warp per row
  for (int l = l_begin + w; l < l_end; l += 8) {
    const size_t rbase = ((size_t)b * L + l) * E;
    float4 u[VEC];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int e0 = (j * 32 + lane) * 4;
      const long long f = (long long)l * E + e0;
      const int hh = (int)(f / ((long long)L * hd)), ll = (int)((f % ((long long)L * hd)) / hd);
      const float g = plain ? sp[hh * L + ll] : 1.f + sp[hh * L + ll];
      const float4 x = ld4<T>(v + rbase + e0);
      u[j] = make_float4(x.x * g, x.y * g, x.z * g, x.w * g);
      s += u[j].x + u[j].y + u[j].z + u[j].w;
      if (plain) st4<T>(y + rbase + e0, u[j]);
    }
    if (plain) continue;            // gamma == NULL: gating only (MultiHeadLinearAttention.forward)
    const float mean = warp_sum(s) * (1.f / E);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float a = u[j].x - mean, bb = u[j].y - mean, c = u[j].z - mean, d = u[j].w - mean;
      q += a * a + bb * bb + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / E) + eps);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int e0 = (j * 32 + lane) * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + e0);
      const float4 bt = *reinterpret_cast<const float4*>(beta + e0);
      st4<T>(y + rbase + e0, make_float4((u[j].x - mean) * rstd * g.x + bt.x,
                                         (u[j].y - mean) * rstd * g.y + bt.y,
                                         (u[j].z - mean) * rstd * g.z + bt.z,
                                         (u[j].w - mean) * rstd * g.w + bt.w));
    }
    if (lane == 0) {
      mean_out[(size_t)b * L + l] = mean;
      rstd_out[(size_t)b * L + l] = rstd;
    }
  }
}

// backward: dv (direct path), dlogits, dgamma/dbeta (atomics into zeroed buffers)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
mhla_gate_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ v,
                        const float* __restrict__ p, const float* __restrict__ mean_in,
                        const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                        T* __restrict__ dv, T* __restrict__ dlogits, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, int L, int H) {
  pdl_trigger();
  pdl_wait();
  constexpr int E = VEC * 128;
  extern __shared__ float smem[];
  float* sp = smem;                 // [H][L] probabilities
  float* sdp = smem + H * L;        // [H][L] d(loss)/d(p)
  float* red = sdp + H * L;         // [8][128] column partial staging
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int hd = E / H;
  for (int i = threadIdx.x; i < H * L; i += 256) sp[i] = p[(size_t)b * H * L + i];
  __syncthreads();
  float4 ag[VEC], ab[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) ag[j] = ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int lanes_per_chunk = hd / 4;           // lanes sharing one (hh, ll) gate
  const bool plain = gamma == nullptr;          // gating only: no residual, no LayerNorm
  const bool sliced = gridDim.y > 1;            // block (b, hh): the rows and the gates of head hh
  const int h_begin = sliced ? blockIdx.y : 0, h_end = sliced ? blockIdx.y + 1 : H;
  const int l_begin = sliced ? blockIdx.y * (L / H) : 0, l_end = sliced ? l_begin + L / H : L;
  for (int l = l_begin + w; l < l_end; l += 8) {
    const size_t rbase = ((size_t)b * L + l) * E;
    const float mean = plain ? 0.f : mean_in[(size_t)b * L + l];
    const float rstd = plain ? 1.f : rstd_in[(size_t)b * L + l];
    float4 xv[VEC], xh[VEC], dg[VEC];
    float gate[VEC];
    int gidx[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int e0 = (j * 32 + lane) * 4;
      const long long f = (long long)l * E + e0;
      const int hh = (int)(f / ((long long)L * hd)), ll = (int)((f % ((long long)L * hd)) / hd);
      gidx[j] = hh * L + ll;
      gate[j] = plain ? sp[gidx[j]] : 1.f + sp[gidx[j]];
      xv[j] = ld4<T>(v + rbase + e0);
      const float4 d = ld4<T>(dy + rbase + e0);
      const float4 g = plain ? make_float4(1.f, 1.f, 1.f, 1.f) : *reinterpret_cast<const float4*>(gamma + e0);
      xh[j] = make_float4((xv[j].x * gate[j] - mean) * rstd, (xv[j].y * gate[j] - mean) * rstd,
                          (xv[j].z * gate[j] - mean) * rstd, (xv[j].w * gate[j] - mean) * rstd);
      dg[j] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s1 += dg[j].x + dg[j].y + dg[j].z + dg[j].w;
      s2 += dg[j].x * xh[j].x + dg[j].y * xh[j].y + dg[j].z * xh[j].z + dg[j].w * xh[j].w;
      ag[j].x += d.x * xh[j].x; ag[j].y += d.y * xh[j].y; ag[j].z += d.z * xh[j].z; ag[j].w += d.w * xh[j].w;
      ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
    }
    const float c1 = plain ? 0.f : warp_sum(s1) * (1.f / E);
    const float c2 = plain ? 0.f : warp_sum(s2) * (1.f / E);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int e0 = (j * 32 + lane) * 4;
      float4 du;
      du.x = rstd * (dg[j].x - c1 - xh[j].x * c2);
      du.y = rstd * (dg[j].y - c1 - xh[j].y * c2);
      du.z = rstd * (dg[j].z - c1 - xh[j].z * c2);
      du.w = rstd * (dg[j].w - c1 - xh[j].w * c2);
      st4<T>(dv + rbase + e0, make_float4(du.x * gate[j], du.y * gate[j], du.z * gate[j], du.w * gate[j]));
      float t = du.x * xv[j].x + du.y * xv[j].y + du.z * xv[j].z + du.w * xv[j].w;
      for (int o = 1; o < lanes_per_chunk; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if ((lane % lanes_per_chunk) == 0) sdp[gidx[j]] = t;     // each (hh,ll) is owned by one chunk
    }
  }
  __syncthreads();
  // softmax-over-L backward, warp per head: dlogit[l,h] = p (dp - sum_l p dp)
  for (int h = h_begin + w; h < h_end; h += 8) {
    float dot = 0.f;
    for (int l = lane; l < L; l += 32) dot += sp[h * L + l] * sdp[h * L + l];
    dot = warp_sum(dot);
    for (int l = lane; l < L; l += 32)
      stf<T>(dlogits, ((size_t)b * L + l) * H + h, sp[h * L + l] * (sdp[h * L + l] - dot));
  }
  // LayerNorm parameter gradients
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    for (int pass = 0; pass < 2; ++pass) {
      const float4 a = pass == 0 ? ag[j] : ab[j];
      __syncthreads();
      red[w * 128 + lane * 4 + 0] = a.x; red[w * 128 + lane * 4 + 1] = a.y;
      red[w * 128 + lane * 4 + 2] = a.z; red[w * 128 + lane * 4 + 3] = a.w;
      __syncthreads();
      if (threadIdx.x < 128) {
        float t = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) t += red[ww * 128 + threadIdx.x];
        float* dst = pass == 0 ? dgamma : dbeta;
        if (dst) atomicAdd(dst + j * 128 + threadIdx.x, t);
      }
    }
  }
}

// ------------------------------------------------------------------ CrossModality triplet loss
// cos: (P, D) fp32 cosine similarities of the unit latents, G: (P, D) int8 labels.
// loss = [ sum_i sum_{p in pos_i} sum_{n in neg_i} relu(s_in - s_ip + m)
//        + sum_{i: pos_i empty} sum_n relu(s_in - sigmoid(1) + m) ] / max(#terms, 1),  s = sigmoid(cos)
// acc[0] += sum of hinge terms, acc[1] += number of terms (doubles).
template <bool BWD>
__global__ void __launch_bounds__(256)
cm_triplet_kernel(const float* __restrict__ cos, const int8_t* __restrict__ G, int D, float margin,
                  double* __restrict__ acc, const float* __restrict__ gout, float* __restrict__ dcos) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* sig = sm;                          // [D]
  int* pos = reinterpret_cast<int*>(sm + D);  // [D] compacted positive columns
  __shared__ int npos_s;
  __shared__ float red[32];
  const int i = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) npos_s = 0;
  __syncthreads();
  for (int j = tid; j < D; j += 256) {
    sig[j] = 1.f / (1.f + __expf(-cos[(size_t)i * D + j]));
    if (G[(size_t)i * D + j] == 1) pos[atomicAdd(&npos_s, 1)] = j;
  }
  __syncthreads();
  const int npos = npos_s;
  const int nneg = D - npos;                // labels are {0,1}: everything not positive is negative
  const float s_self = 1.f / (1.f + __expf(-1.f));
  float coef = 0.f;
  if (BWD) {
    const double cnt = acc[1];
    coef = gout[0] / (float)(cnt > 1.0 ? cnt : 1.0);
  }
  float local = 0.f;
  if (nneg > 0) {
    for (int j = tid; j < D; j += 256) {
      if (G[(size_t)i * D + j] == 1) continue;
      const float sn = sig[j];
      if (npos > 0) {
        int active = 0;
        for (int k = 0; k < npos; ++k) {
          const float t = sn - sig[pos[k]] + margin;
          if (t > 0.f) { local += t; ++active; }
        }
        if (BWD) dcos[(size_t)i * D + j] = coef * (float)active * sn * (1.f - sn);
      } else {
        const float t = sn - s_self + margin;
        if (t > 0.f) local += t;
        if (BWD) dcos[(size_t)i * D + j] = t > 0.f ? coef * sn * (1.f - sn) : 0.f;
      }
    }
  } else if (BWD) {
    for (int j = tid; j < D; j += 256) dcos[(size_t)i * D + j] = 0.f;
  }
  if (BWD) {
    // positives: d/ds_ip = -(number of negatives n with s_in - s_ip + m > 0)
    for (int k = tid; k < npos; k += 256) {
      const int jp = pos[k];
      const float sp = sig[jp];
      int active = 0;
      if (nneg > 0)
        for (int j = 0; j < D; ++j)
          if (G[(size_t)i * D + j] != 1 && sig[j] - sp + margin > 0.f) ++active;
      dcos[(size_t)i * D + jp] = -coef * (float)active * sp * (1.f - sp);
    }
  } else {
    const float total = block_sum(local, red);
    if (tid == 0) {
      atomicAdd(acc, (double)total);
      const double terms = nneg > 0 ? (npos > 0 ? (double)npos * nneg : (double)nneg) : 0.0;
      atomicAdd(acc + 1, terms);
    }
  }
}

__global__ void cm_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss) {
  pdl_trigger();
  pdl_wait();
  const double cnt = acc[1] > 1.0 ? acc[1] : 1.0;
  loss[0] = (float)(acc[0] / cnt);
}

// ------------------------------------------------------------------ batched 2-D transpose
// y[b, c, r] = x[b, r, c]; 32x32 tiles through padded shared memory, coalesced both ways.
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ x, T* __restrict__ y, int R, int Cc) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * R * Cc;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    if (r < R && c < Cc) tile[ty + i][tx] = ldf<T>(x, base + (size_t)r * Cc + c);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (r < R && c < Cc) stf<T>(y, base + (size_t)c * R + r, tile[tx][ty + i]);
  }
}

// 64 x 64 tiles with 16-byte global accesses both ways: y[b, c, r] = in[b, r, c].
//   SRC 0   in[b, r, c] = x[b, r, c], optionally BatchNorm-normalised per input column c:
//           (x - mean[c]) * rstd[c] * gamma[c] + beta[c] -- the last BatchNorm of ProteinCNN is applied
//           while its output is laid out as the reference's (B, C, L) buffer (model/basic_model.py:178-179)
//   SRC 1   in[b, r, c] = g[b, (f / R) % P, f % R] * scale with f = r * Cc + c: the gradient that reaches
//           the channels-last (B, L = Cc, C = R) activation through transpose -> .view(B, L, C) ->
//           .view(B, S, P, C).mean(1) (model/basic_model.py:179, model/DrugLAMP.py:35-37), scale = 1 / S
template <typename T, int SRC>
__global__ void __launch_bounds__(256)
transpose64_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int R, int Cc, int P, float scale) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = VecWidth<T>::N;
  constexpr int TPR = 64 / V, RPP = 256 / TPR;
  __shared__ float tile[64][65];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const size_t b = blockIdx.z;
  const int tv = threadIdx.x % TPR, tr = threadIdx.x / TPR;
  {
    const int c = c0 + tv * V;
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sc[k] = 1.f; sh[k] = 0.f; }
    if (SRC == 0 && mean != nullptr && c < Cc) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float a = rstd[c + k] * (gamma ? gamma[c + k] : 1.f);
        sc[k] = a;
        sh[k] = (beta ? beta[c + k] : 0.f) - mean[c + k] * a;
      }
    }
#pragma unroll
    for (int i = 0; i < 64; i += RPP) {
      const int r = r0 + tr + i;
      if (r < R && c < Cc) {
        float v[V];
        if (SRC == 0) {
          ldv(x + (b * R + r) * Cc + c, v);
#pragma unroll
          for (int k = 0; k < V; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
        } else {
          const long long f = (long long)r * Cc + c;
          const long long rv = f / R;
          ldv(x + (b * P + (size_t)(rv % P)) * R + (size_t)(f - rv * R), v);
#pragma unroll
          for (int k = 0; k < V; ++k) v[k] *= scale;
        }
#pragma unroll
        for (int k = 0; k < V; ++k) tile[tv * V + k][tr + i] = v[k];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 64; i += RPP) {
    const int c = c0 + tr + i, r = r0 + tv * V;
    if (c < Cc && r < R) {
      float v[V];
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = tile[tr + i][tv * V + k];
      stv(y + (b * Cc + c) * R + r, v);
    }
  }
}

// ------------------------------------------------------------------ cross entropy (MLM heads)
// logit[r, c] = x[r*ld + c] + (extra ? extra[r] * wextra[c] : 0);  rows with label == ignore are
// skipped.  One warp per row, classes strided over lanes (C = 27 for the MLM heads).
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const T* __restrict__ x, const long long* __restrict__ labels,
                     const float* __restrict__ extra, const float* __restrict__ wextra,
                     long long rows, int C, long long ld, long long ignore_index,
                     double* __restrict__ acc, const float* __restrict__ gout, T* __restrict__ dx,
                     float* __restrict__ dextra_dot) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long lab = labels[r];
  const bool skip = lab == ignore_index;
  if (skip) {
    if (BWD) {
      for (int c = lane; c < C; c += 32) stf<T>(dx, r * ld + c, 0.f);
      if (dextra_dot && lane == 0) dextra_dot[r] = 0.f;
    }
    return;
  }
  const float ex = extra ? extra[r] : 0.f;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, ldf<T>(x, r * ld + c) + (extra ? ex * wextra[c] : 0.f));
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(ldf<T>(x, r * ld + c) + (extra ? ex * wextra[c] : 0.f) - m);
  s = warp_sum(s);
  if (!BWD) {
    if (lane == 0) {
      const float xl = ldf<T>(x, r * ld + lab) + (extra ? ex * wextra[lab] : 0.f);
      atomicAdd(acc, (double)(m + __logf(s) - xl));
      atomicAdd(acc + 1, 1.0);
    }
  } else {
    const float coef = gout[0] / (float)acc[1];
    const float inv = 1.f / s;
    float dot = 0.f;   // sum_c dlogit[c] * wextra[c]  -> gradient w.r.t. extra[r]
    for (int c = lane; c < C; c += 32) {
      const float p = __expf(ldf<T>(x, r * ld + c) + (extra ? ex * wextra[c] : 0.f) - m) * inv;
      const float d = coef * (p - (c == lab ? 1.f : 0.f));
      stf<T>(dx, r * ld + c, d);
      if (extra) dot += d * wextra[c];
    }
    if (dextra_dot) {
      dot = warp_sum(dot);
      if (lane == 0) dextra_dot[r] = dot;
    }
  }
}

__global__ void ce_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss) {
  pdl_trigger();
  pdl_wait();
  loss[0] = (float)(acc[0] / acc[1]);
}

// ------------------------------------------------------------------ BCE
__global__ void bce_fwd_kernel(const float* __restrict__ score, const float* __restrict__ y,
                               float* __restrict__ prob, float* __restrict__ loss, int n) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  float local = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float pr = 1.f / (1.f + expf(-score[i]));
    prob[i] = pr;
    const float l1 = fmaxf(logf(pr), -100.f), l0 = fmaxf(logf(1.f - pr), -100.f);
    local -= y[i] * l1 + (1.f - y[i]) * l0;
  }
  const float t = block_sum(local, red);
  if (threadIdx.x == 0) loss[0] = t / (float)n;
}

__global__ void bce_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ y,
                               const float* __restrict__ gout, float* __restrict__ dscore, int n) {
  pdl_trigger();
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float pr = prob[i];
    const float dn = (pr - y[i]) / fmaxf((1.f - pr) * pr, 1e-12f);   // BCELoss backward
    dscore[i] = gout[0] * dn * pr * (1.f - pr) / (float)n;           // through the sigmoid
  }
}

// ---------------------------------------------------------------- packed collate -> dense
// The reference's collate pads every sample's LLM embedding rows on the host (utils.py:304-324):
// tail_pad copies the R_b rows once, repeat_pad tiles them floor(maxsize / R_b) times, the rest is
// zeros.  Here the host ships the R_b rows only and this kernel writes the identical dense
// (B, maxsize, C) tensor: out[b, r, :] = rows[off_b + r mod R_b] for r < reps_b * R_b, else 0.
// One warp per output row, 16-byte vectors; the source rows are re-read from L2.
__global__ void __launch_bounds__(256)
expand_rows_kernel(const float* __restrict__ rows, const int* __restrict__ offsets,
                   float* __restrict__ out, int B, int maxsize, int C4, int repeat) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * maxsize;
  const long long w0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 8;
  for (long long o = w0; o < total; o += nw) {
    const int b = (int)(o / maxsize), r = (int)(o - (long long)b * maxsize);
    const int off = offsets[b], R = offsets[b + 1] - off;
    const int filled = R <= 0 ? 0 : (repeat ? (maxsize / R) * R : (R < maxsize ? R : maxsize));
    float4* dst = reinterpret_cast<float4*>(out) + o * C4;
    if (r < filled) {
      const float4* src = reinterpret_cast<const float4*>(rows) + (long long)(off + r % R) * C4;
      for (int c = lane; c < C4; c += 32) dst[c] = src[c];
    } else {
      for (int c = lane; c < C4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ---------------------------------------------------------------- ProteinCNN input
// out[r, 0:127] = table[tok[r], :], out[r, 127] = fill[r]   (embedding gather + concat + cast in
// one pass; model/basic_model.py:171-173).  The 27 x 127 fp32 table sits in shared memory; one
// warp writes one 128-wide row per step, lane = 4 consecutive columns.
constexpr int kEmbW = 128;          // embedding_dim: 127 table columns + the fill bit
constexpr int kEmbMaxVocab = 32;

template <typename TokT> __device__ __forceinline__ int tok_index(const TokT* t, long long i) { return (int)t[i]; }

template <typename T, typename TokT>
__global__ void __launch_bounds__(256)
embed_fill_fwd_kernel(const TokT* __restrict__ tok, const float* __restrict__ fill,
                      const float* __restrict__ table, T* __restrict__ out, long long rows, int vocab) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tab[kEmbMaxVocab * (kEmbW - 1) + 4];
  for (int i = threadIdx.x; i < vocab * (kEmbW - 1); i += 256) tab[i] = table[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long w0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 8;
  for (long long r = w0; r < rows; r += nw) {
    int v = tok_index(tok, r);
    v = v < 0 ? 0 : (v >= vocab ? vocab - 1 : v);
    const float* src = tab + v * (kEmbW - 1) + lane * 4;
    float4 o;
    o.x = src[0]; o.y = src[1]; o.z = src[2];
    o.w = lane == 31 ? fill[r] : src[3];
    st4<T>(out + r * kEmbW + lane * 4, o);
  }
}

// dtable[v, c] += sum over rows with token v of g[r, c], c < 127 (the fill-bit column has no
// parameter).  Every warp owns a private 27 x 128 fp32 table in shared memory and adds its rows
// with plain vector read-modify-writes (lane = 4 fixed columns, so no two lanes share an address);
// the block then folds its 8 tables and issues one atomic per table entry.
template <typename T, typename TokT>
__global__ void __launch_bounds__(256)
embed_fill_bwd_kernel(const TokT* __restrict__ tok, const T* __restrict__ g, float* __restrict__ dtable,
                      long long rows, int vocab, int padding_idx) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float acc[];                     // [8 warps][vocab][128]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per_warp = vocab * kEmbW;
  for (int i = threadIdx.x; i < 8 * per_warp; i += 256) acc[i] = 0.f;
  __syncthreads();
  float* mine = acc + w * per_warp + lane * 4;
  const long long w0 = (long long)blockIdx.x * 8 + w, nw = (long long)gridDim.x * 8;
  long long r = w0;
  for (; r + 7 * nw < rows; r += 8 * nw) {           // eight rows of loads in flight
    float4 gv[8];
    int v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      gv[u] = ld4<T>(g + (r + u * nw) * kEmbW + lane * 4);
      v[u] = tok_index(tok, r + u * nw);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (v[u] < 0 || v[u] >= vocab) continue;
      float4* a = reinterpret_cast<float4*>(mine + v[u] * kEmbW);
      float4 t = *a;
      t.x += gv[u].x; t.y += gv[u].y; t.z += gv[u].z; t.w += gv[u].w;
      *a = t;
    }
  }
  for (; r < rows; r += nw) {
    const float4 gv = ld4<T>(g + r * kEmbW + lane * 4);
    const int v = tok_index(tok, r);
    if (v < 0 || v >= vocab) continue;
    float4* a = reinterpret_cast<float4*>(mine + v * kEmbW);
    float4 t = *a;
    t.x += gv.x; t.y += gv.y; t.z += gv.z; t.w += gv.w;
    *a = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < per_warp; i += 256) {
    const int v = i / kEmbW, c = i % kEmbW;
    if (c == kEmbW - 1 || v == padding_idx) continue;
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += acc[ww * per_warp + i];
    if (t != 0.f) atomicAdd(dtable + v * (kEmbW - 1) + c, t);
  }
}

int ew_grid(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace
}  // namespace dl

using namespace dl;

extern "C" int dl_fillbit_pool(const float* x, float* bit_out, float* cat_out, void* pooled,
                               int32_t pooled_dtype, int64_t B, int32_t S, int32_t L, int32_t C,
                               int32_t ld_pooled, void* stream) {
  DL_REQUIRE(x != nullptr, "dl_fillbit_pool: null input");
  DL_REQUIRE(B >= 0 && S >= 1 && L >= 1 && C >= 4 && C % 4 == 0 && C <= 2048,
             "dl_fillbit_pool: need C %% 4 == 0 and C <= 2048 (got B=%lld S=%d L=%d C=%d)", (long long)B, S, L, C);
  DL_REQUIRE(((uintptr_t)x & 15) == 0, "dl_fillbit_pool: x must be 16-byte aligned");
  DL_REQUIRE(B <= 65535, "dl_fillbit_pool: B too large");
  if (ld_pooled == 0) ld_pooled = C + 1;
  DL_REQUIRE(ld_pooled >= C + 1, "dl_fillbit_pool: ld_pooled %d < C + 1", ld_pooled);
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_rows = B * (long long)L;
  const unsigned grid = (unsigned)((n_rows + 7) / 8);
  const int smem = cat_out ? 8 * C * (int)sizeof(float) : 0;
  const int nk = (C / 4 + 31) / 32;                 // float4 chunks per lane: 3 (C=384), 5 (640), <= 16
#define DL_FILLBIT(TT, NKK)                                                                        \
  do {                                                                                             \
    if (smem > 48 * 1024)                                                                          \
      DL_CUDA(cudaFuncSetAttribute(fillbit_pool_kernel<TT, NKK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    DL_LAUNCH((fillbit_pool_kernel<TT, NKK>), grid, 256, smem, st, x, bit_out, cat_out, (TT*)pooled, S, L, C, ld_pooled, n_rows); \
  } while (0)
#define DL_FILLBIT_T(TT)                                                                           \
  do {                                                                                             \
    if (nk <= 3) DL_FILLBIT(TT, 3); else if (nk <= 5) DL_FILLBIT(TT, 5);                           \
    else if (nk <= 8) DL_FILLBIT(TT, 8); else DL_FILLBIT(TT, 16);                                  \
  } while (0)
  if (pooled_dtype == DL_BF16) DL_FILLBIT_T(__nv_bfloat16); else DL_FILLBIT_T(float);
#undef DL_FILLBIT_T
#undef DL_FILLBIT
  DL_LAUNCH_CHECK("fillbit_pool_kernel");
  count_launch();
  return 0;
}

namespace {
template <typename T, typename TokT>
int embed_fwd_launch(const void* tok, const float* fill, const float* table, void* out, long long rows,
                     int vocab, cudaStream_t st) {
  long long blocks = (rows + 63) / 64;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  DL_LAUNCH((embed_fill_fwd_kernel<T, TokT>), (int)blocks, 256, 0, st, (const TokT*)tok, fill, table, (T*)out, rows, vocab);
  DL_LAUNCH_CHECK("embed_fill_fwd_kernel");
  count_launch();
  return 0;
}
template <typename T, typename TokT>
int embed_bwd_launch(const void* tok, const void* g, float* dtable, long long rows, int vocab,
                     int padding_idx, cudaStream_t st) {
  const int smem = 8 * vocab * kEmbW * (int)sizeof(float);
  static int configured = 0;       // the attribute is per kernel instantiation
  if (configured < smem) {
    DL_CUDA(cudaFuncSetAttribute(embed_fill_bwd_kernel<T, TokT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  long long blocks = (rows + 255) / 256;
  if (blocks > sm_count()) blocks = sm_count();
  DL_LAUNCH((embed_fill_bwd_kernel<T, TokT>), (int)blocks, 256, smem, st, (const TokT*)tok, (const T*)g, dtable, rows, vocab, padding_idx);
  DL_LAUNCH_CHECK("embed_fill_bwd_kernel");
  count_launch();
  return 0;
}
}  // namespace

extern "C" int dl_embed_fill_fwd(const void* tokens, int32_t tok_dtype, const float* fill,
                                 const float* table, void* out, int64_t rows, int32_t vocab,
                                 int32_t width, int32_t dtype, void* stream) {
  DL_REQUIRE(tokens && fill && table && out, "dl_embed_fill_fwd: null pointer");
  DL_REQUIRE(width == kEmbW, "dl_embed_fill_fwd: width must be 128 (127 embedding columns + fill bit), got %d", width);
  DL_REQUIRE(vocab >= 1 && vocab <= kEmbMaxVocab, "dl_embed_fill_fwd: vocab must be in [1, %d]", kEmbMaxVocab);
  DL_REQUIRE(tok_dtype == DL_TOK_I64 || tok_dtype == DL_TOK_F64, "dl_embed_fill_fwd: tok_dtype must be DL_TOK_I64 or DL_TOK_F64");
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DL_BF16)
    return tok_dtype == DL_TOK_I64 ? embed_fwd_launch<__nv_bfloat16, long long>(tokens, fill, table, out, rows, vocab, st)
                                   : embed_fwd_launch<__nv_bfloat16, double>(tokens, fill, table, out, rows, vocab, st);
  return tok_dtype == DL_TOK_I64 ? embed_fwd_launch<float, long long>(tokens, fill, table, out, rows, vocab, st)
                                 : embed_fwd_launch<float, double>(tokens, fill, table, out, rows, vocab, st);
}

extern "C" int dl_embed_fill_bwd(const void* tokens, int32_t tok_dtype, const void* g, float* dtable,
                                 int64_t rows, int32_t vocab, int32_t width, int32_t padding_idx,
                                 int32_t dtype, void* stream) {
  DL_REQUIRE(tokens && g && dtable, "dl_embed_fill_bwd: null pointer");
  DL_REQUIRE(width == kEmbW, "dl_embed_fill_bwd: width must be 128, got %d", width);
  DL_REQUIRE(vocab >= 1 && vocab <= kEmbMaxVocab, "dl_embed_fill_bwd: vocab must be in [1, %d]", kEmbMaxVocab);
  DL_REQUIRE(tok_dtype == DL_TOK_I64 || tok_dtype == DL_TOK_F64, "dl_embed_fill_bwd: tok_dtype must be DL_TOK_I64 or DL_TOK_F64");
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DL_BF16)
    return tok_dtype == DL_TOK_I64 ? embed_bwd_launch<__nv_bfloat16, long long>(tokens, g, dtable, rows, vocab, padding_idx, st)
                                   : embed_bwd_launch<__nv_bfloat16, double>(tokens, g, dtable, rows, vocab, padding_idx, st);
  return tok_dtype == DL_TOK_I64 ? embed_bwd_launch<float, long long>(tokens, g, dtable, rows, vocab, padding_idx, st)
                                 : embed_bwd_launch<float, double>(tokens, g, dtable, rows, vocab, padding_idx, st);
}

extern "C" int dl_expand_rows(const float* rows, const int32_t* offsets, float* out, int64_t B,
                              int32_t maxsize, int32_t C, int32_t repeat, void* stream) {
  DL_REQUIRE(rows && offsets && out, "dl_expand_rows: null pointer");
  DL_REQUIRE(B >= 0 && maxsize >= 1 && C >= 4 && C % 4 == 0, "dl_expand_rows: need C %% 4 == 0 (got B=%lld maxsize=%d C=%d)", (long long)B, maxsize, C);
  DL_REQUIRE((((uintptr_t)rows | (uintptr_t)out) & 15) == 0, "dl_expand_rows: rows and out must be 16-byte aligned");
  DL_REQUIRE(B * (long long)maxsize < (1ll << 31), "dl_expand_rows: too many rows");
  if (B == 0) return 0;
  long long blocks = (B * (long long)maxsize + 7) / 8;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  DL_LAUNCH(expand_rows_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, rows, offsets, out, (int)B, maxsize, C / 4, repeat);
  DL_LAUNCH_CHECK("expand_rows_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_site_pool_fwd(const void* x, void* y, int64_t B, int32_t S, int32_t L, int32_t C,
                                int64_t ldy, int32_t dtype, void* stream) {
  DL_REQUIRE(x && y && C % 4 == 0 && ldy % 4 == 0 && ldy >= C, "dl_site_pool_fwd: bad arguments");
  const long long LC4 = (long long)L * C / 4, n4 = B * LC4;
  if (n4 <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DL_BF16)
    DL_LAUNCH((site_pool_fwd_kernel<__nv_bfloat16>), ew_grid(n4, 256), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, S, LC4, n4, ldy / 4, C / 4);
  else
    DL_LAUNCH((site_pool_fwd_kernel<float>), ew_grid(n4, 256), 256, 0, st, (const float*)x, (float*)y, S, LC4, n4, ldy / 4, C / 4);
  DL_LAUNCH_CHECK("site_pool_fwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_site_pool_bwd(const void* dy, void* dx, int64_t B, int32_t S, int32_t L,
                                int32_t C, int64_t ldy, int32_t dtype, void* stream) {
  DL_REQUIRE(dy && dx && C % 4 == 0 && ldy % 4 == 0 && ldy >= C, "dl_site_pool_bwd: bad arguments");
  const long long LC4 = (long long)L * C / 4, n4 = B * LC4;
  if (n4 <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DL_BF16)
    DL_LAUNCH((site_pool_bwd_kernel<__nv_bfloat16>), ew_grid(n4, 256), 256, 0, st, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, S, LC4, n4, ldy / 4, C / 4);
  else
    DL_LAUNCH((site_pool_bwd_kernel<float>), ew_grid(n4, 256), 256, 0, st, (const float*)dy, (float*)dx, S, LC4, n4, ldy / 4, C / 4);
  DL_LAUNCH_CHECK("site_pool_bwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_mhla_gate_ln_fwd(const void* v, const void* logits, const float* gamma,
                                   const float* beta, void* y, float* p_out, float* mean,
                                   float* rstd, int64_t B, int32_t L, int32_t E, int32_t H,
                                   float eps, int32_t dtype, void* stream) {
  DL_REQUIRE(v && logits && y && p_out, "dl_mhla_gate_ln_fwd: null pointer");
  DL_REQUIRE(gamma == nullptr || (beta && mean && rstd), "dl_mhla_gate_ln_fwd: LayerNorm mode needs beta, mean, rstd");
  DL_REQUIRE(E % 128 == 0 && E <= 512 && H >= 1 && E % H == 0 && (E / H) % 4 == 0 && 32 % ((E / H) / 4) == 0,
             "dl_mhla_gate_ln_fwd: unsupported E=%d H=%d", E, H);
  DL_REQUIRE((long long)H * L * 4 <= 160 * 1024, "dl_mhla_gate_ln_fwd: H*L too large for shared memory");
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)H * L * sizeof(float);
#define DL_MHLA_FWD(TT, VV)                                                                         \
  do {                                                                                              \
    if (smem > 48 * 1024)                                                                           \
      DL_CUDA(cudaFuncSetAttribute(mhla_gate_ln_fwd_kernel<TT, VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    DL_LAUNCH((mhla_gate_ln_fwd_kernel<TT, VV>), dim3((unsigned)B, L % H == 0 ? H : 1), 256, smem, st, (const TT*)v, (const TT*)logits, gamma, beta, (TT*)y, p_out, mean, rstd, L, H, eps); \
  } while (0)
  if (dtype == DL_BF16) {
    if (E == 128) DL_MHLA_FWD(__nv_bfloat16, 1); else if (E == 256) DL_MHLA_FWD(__nv_bfloat16, 2);
    else if (E == 384) DL_MHLA_FWD(__nv_bfloat16, 3); else DL_MHLA_FWD(__nv_bfloat16, 4);
  } else {
    if (E == 128) DL_MHLA_FWD(float, 1); else if (E == 256) DL_MHLA_FWD(float, 2);
    else if (E == 384) DL_MHLA_FWD(float, 3); else DL_MHLA_FWD(float, 4);
  }
#undef DL_MHLA_FWD
  DL_LAUNCH_CHECK("mhla_gate_ln_fwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_mhla_gate_ln_bwd(const void* dy, const void* v, const float* p, const float* mean,
                                   const float* rstd, const float* gamma, void* dv, void* dlogits,
                                   float* dgamma, float* dbeta, int64_t B, int32_t L, int32_t E,
                                   int32_t H, int32_t dtype, void* stream) {
  DL_REQUIRE(dy && v && p && dv && dlogits, "dl_mhla_gate_ln_bwd: null pointer");
  DL_REQUIRE(gamma == nullptr || (mean && rstd), "dl_mhla_gate_ln_bwd: LayerNorm mode needs mean, rstd");
  DL_REQUIRE(E % 128 == 0 && E <= 512 && H >= 1 && E % H == 0 && (E / H) % 4 == 0 && 32 % ((E / H) / 4) == 0,
             "dl_mhla_gate_ln_bwd: unsupported E=%d H=%d", E, H);
  cudaStream_t st = (cudaStream_t)stream;
  if (dgamma) DL_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * E, st));
  if (dbeta) DL_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * E, st));
  if (B <= 0) return 0;
  const size_t smem = ((size_t)2 * H * L + 8 * 128) * sizeof(float);
  DL_REQUIRE(smem <= 200 * 1024, "dl_mhla_gate_ln_bwd: H*L too large for shared memory");
#define DL_MHLA_BWD(TT, VV)                                                                         \
  do {                                                                                              \
    if (smem > 48 * 1024)                                                                           \
      DL_CUDA(cudaFuncSetAttribute(mhla_gate_ln_bwd_kernel<TT, VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    DL_LAUNCH((mhla_gate_ln_bwd_kernel<TT, VV>), dim3((unsigned)B, L % H == 0 ? H : 1), 256, smem, st, (const TT*)dy, (const TT*)v, p, mean, rstd, gamma, (TT*)dv, (TT*)dlogits, dgamma, dbeta, L, H); \
  } while (0)
  if (dtype == DL_BF16) {
    if (E == 128) DL_MHLA_BWD(__nv_bfloat16, 1); else if (E == 256) DL_MHLA_BWD(__nv_bfloat16, 2);
    else if (E == 384) DL_MHLA_BWD(__nv_bfloat16, 3); else DL_MHLA_BWD(__nv_bfloat16, 4);
  } else {
    if (E == 128) DL_MHLA_BWD(float, 1); else if (E == 256) DL_MHLA_BWD(float, 2);
    else if (E == 384) DL_MHLA_BWD(float, 3); else DL_MHLA_BWD(float, 4);
  }
#undef DL_MHLA_BWD
  DL_LAUNCH_CHECK("mhla_gate_ln_bwd_kernel");
  count_launch();
  return 0;
}

// acc: 2 doubles of workspace (zeroed here); loss: 1 float
extern "C" int dl_cm_triplet_fwd(const float* cos, const int8_t* G, int64_t P, int64_t D,
                                 float margin, double* acc, float* loss, void* stream) {
  DL_REQUIRE(cos && G && acc && loss, "dl_cm_triplet_fwd: null pointer");
  DL_REQUIRE(P >= 0 && D >= 1 && D <= 24576, "dl_cm_triplet_fwd: D must be in [1, 24576]");
  cudaStream_t st = (cudaStream_t)stream;
  DL_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  const size_t smem = (size_t)D * 8;
  if (P > 0) {
    if (smem > 48 * 1024)
      DL_CUDA(cudaFuncSetAttribute(cm_triplet_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DL_LAUNCH((cm_triplet_kernel<false>), (unsigned)P, 256, smem, st, cos, G, (int)D, margin, acc, nullptr, nullptr);
    DL_LAUNCH_CHECK("cm_triplet_kernel");
    count_launch();
  }
  DL_LAUNCH(cm_finalize_kernel, 1, 1, 0, st, acc, loss);
  DL_LAUNCH_CHECK("cm_finalize_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_cm_triplet_bwd(const float* cos, const int8_t* G, int64_t P, int64_t D,
                                 float margin, const double* acc, const float* gout, float* dcos,
                                 void* stream) {
  DL_REQUIRE(cos && G && acc && gout && dcos, "dl_cm_triplet_bwd: null pointer");
  DL_REQUIRE(P >= 0 && D >= 1 && D <= 24576, "dl_cm_triplet_bwd: D must be in [1, 24576]");
  if (P == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)D * 8;
  if (smem > 48 * 1024)
    DL_CUDA(cudaFuncSetAttribute(cm_triplet_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DL_LAUNCH((cm_triplet_kernel<true>), (unsigned)P, 256, smem, st, cos, G, (int)D, margin, const_cast<double*>(acc), gout, dcos);
  DL_LAUNCH_CHECK("cm_triplet_kernel(bwd)");
  count_launch();
  return 0;
}

extern "C" int dl_transpose(const void* x, void* y, int64_t B, int32_t R, int32_t Cc, int32_t dtype,
                            void* stream) {
  DL_REQUIRE(x && y && R >= 1 && Cc >= 1 && B >= 0 && B <= 65535, "dl_transpose: bad arguments");
  if (B == 0) return 0;
  dim3 grid(ceil_div(Cc, 32), ceil_div(R, 32), (unsigned)B);
  DL_REQUIRE(grid.y <= 65535, "dl_transpose: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DL_BF16)
    DL_LAUNCH((transpose_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, R, Cc);
  else
    DL_LAUNCH((transpose_kernel<float>), grid, 256, 0, st, (const float*)x, (float*)y, R, Cc);
  DL_LAUNCH_CHECK("transpose_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_bn_transpose(const void* x, void* y, const float* mean, const float* rstd, const float* gamma,
                               const float* beta, int64_t B, int32_t R, int32_t Cc, int32_t dtype, void* stream) {
  DL_REQUIRE(x && y && R >= 1 && Cc >= 1 && B >= 0 && B <= 65535, "dl_bn_transpose: bad arguments");
  DL_REQUIRE(mean == nullptr || rstd != nullptr, "dl_bn_transpose: mean needs rstd");
  const int V = dtype == DL_BF16 ? 8 : 4;
  DL_REQUIRE(R % V == 0 && Cc % V == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0,
             "dl_bn_transpose: extents must be multiples of %d elements and the tensors 16-byte aligned", V);
  if (B == 0) return 0;
  dim3 grid(ceil_div(Cc, 64), ceil_div(R, 64), (unsigned)B);
  DL_REQUIRE(grid.y <= 65535, "dl_bn_transpose: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DL_BF16)
    DL_LAUNCH((transpose64_kernel<__nv_bfloat16, 0>), grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, mean, rstd, gamma, beta, R, Cc, 0, 1.f);
  else
    DL_LAUNCH((transpose64_kernel<float, 0>), grid, 256, 0, st, (const float*)x, (float*)y, mean, rstd, gamma, beta, R, Cc, 0, 1.f);
  DL_LAUNCH_CHECK("transpose64_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_site_pool_view_bwd(const void* g, void* dx, int64_t B, int32_t S, int32_t L, int32_t C,
                                     int32_t dtype, void* stream) {
  DL_REQUIRE(g && dx && S >= 1 && L >= 1 && C >= 1 && B >= 0 && B <= 65535, "dl_site_pool_view_bwd: bad arguments");
  const int V = dtype == DL_BF16 ? 8 : 4;
  DL_REQUIRE(L % S == 0 && L % V == 0 && C % V == 0 && (((uintptr_t)g | (uintptr_t)dx) & 15) == 0,
             "dl_site_pool_view_bwd: L must be a multiple of S, L and C multiples of %d elements, tensors 16-byte aligned", V);
  if (B == 0) return 0;
  // the (B, C, L) gradient is never materialised: "input" rows = channels (R = C), columns = positions (Cc = L)
  dim3 grid(ceil_div(L, 64), ceil_div(C, 64), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = 1.f / (float)S;
  if (dtype == DL_BF16)
    DL_LAUNCH((transpose64_kernel<__nv_bfloat16, 1>), grid, 256, 0, st, (const __nv_bfloat16*)g, (__nv_bfloat16*)dx, nullptr, nullptr, nullptr, nullptr, C, L, L / S, scale);
  else
    DL_LAUNCH((transpose64_kernel<float, 1>), grid, 256, 0, st, (const float*)g, (float*)dx, nullptr, nullptr, nullptr, nullptr, C, L, L / S, scale);
  DL_LAUNCH_CHECK("transpose64_kernel");
  count_launch();
  return 0;
}

// acc: 2 doubles (loss sum, number of non-ignored rows), zeroed here; loss: 1 float (mean)
extern "C" int dl_cross_entropy_fwd(const void* x, const int64_t* labels, const float* extra,
                                    const float* wextra, int64_t rows, int32_t C, int64_t ld,
                                    int64_t ignore_index, double* acc, float* loss, int32_t dtype,
                                    void* stream) {
  DL_REQUIRE(x && labels && acc && loss && C >= 1 && ld >= C, "dl_cross_entropy_fwd: bad arguments");
  DL_REQUIRE((extra == nullptr) == (wextra == nullptr), "dl_cross_entropy_fwd: extra and wextra go together");
  cudaStream_t st = (cudaStream_t)stream;
  DL_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  if (rows > 0) {
    const int grid = ceil_div(rows, 8);
    if (dtype == DL_BF16)
      DL_LAUNCH((cross_entropy_kernel<__nv_bfloat16, false>), grid, 256, 0, st, (const __nv_bfloat16*)x, (const long long*)labels, extra, wextra, rows, C, ld, ignore_index, acc, nullptr, nullptr, nullptr);
    else
      DL_LAUNCH((cross_entropy_kernel<float, false>), grid, 256, 0, st, (const float*)x, (const long long*)labels, extra, wextra, rows, C, ld, ignore_index, acc, nullptr, nullptr, nullptr);
    DL_LAUNCH_CHECK("cross_entropy_kernel");
    count_launch();
  }
  DL_LAUNCH(ce_finalize_kernel, 1, 1, 0, st, acc, loss);
  DL_LAUNCH_CHECK("ce_finalize_kernel");
  count_launch();
  return 0;
}

// dx: same layout as x; dextra: [rows] fp32 or NULL (gradient w.r.t. extra)
extern "C" int dl_cross_entropy_bwd(const void* x, const int64_t* labels, const float* extra,
                                    const float* wextra, int64_t rows, int32_t C, int64_t ld,
                                    int64_t ignore_index, const double* acc, const float* gout,
                                    void* dx, float* dextra, int32_t dtype, void* stream) {
  DL_REQUIRE(x && labels && acc && gout && dx && C >= 1 && ld >= C, "dl_cross_entropy_bwd: bad arguments");
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(rows, 8);
  if (dtype == DL_BF16)
    DL_LAUNCH((cross_entropy_kernel<__nv_bfloat16, true>), grid, 256, 0, st, (const __nv_bfloat16*)x, (const long long*)labels, extra, wextra, rows, C, ld, ignore_index, const_cast<double*>(acc), gout, (__nv_bfloat16*)dx, dextra);
  else
    DL_LAUNCH((cross_entropy_kernel<float, true>), grid, 256, 0, st, (const float*)x, (const long long*)labels, extra, wextra, rows, C, ld, ignore_index, const_cast<double*>(acc), gout, (float*)dx, dextra);
  DL_LAUNCH_CHECK("cross_entropy_kernel(bwd)");
  count_launch();
  return 0;
}

extern "C" int dl_bce_fwd(const float* score, const float* y, float* prob, float* loss, int64_t n,
                          void* stream) {
  DL_REQUIRE(score && y && prob && loss && n >= 1 && n < (1 << 30), "dl_bce_fwd: bad arguments");
  DL_LAUNCH(bce_fwd_kernel, 1, 256, 0, (cudaStream_t)stream, score, y, prob, loss, (int)n);
  DL_LAUNCH_CHECK("bce_fwd_kernel");
  count_launch();
  return 0;
}

extern "C" int dl_bce_bwd(const float* prob, const float* y, const float* gout, float* dscore,
                          int64_t n, void* stream) {
  DL_REQUIRE(prob && y && gout && dscore && n >= 1 && n < (1 << 30), "dl_bce_bwd: bad arguments");
  DL_LAUNCH(bce_bwd_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream, prob, y, gout, dscore, (int)n);
  DL_LAUNCH_CHECK("bce_bwd_kernel");
  count_launch();
  return 0;
}
