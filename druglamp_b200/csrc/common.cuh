// Shared device/host helpers for libdruglamp_sm100.so (sm_100a only).
//
// Error convention of the C ABI (include/druglamp_sm100.h): every entry point returns int,
// 0 = OK, negative = argument/shape/alignment error, positive = cudaError_t; the message
// is kept in a thread-local buffer readable through dl_last_error().  Nothing throws.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dl {

int set_error(int code, const char* fmt, ...);

#define DL_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return dl::set_error(-1, __VA_ARGS__);         \
  } while (0)

#define DL_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return dl::set_error((int)_e, "%s failed: %s", #expr, cudaGetErrorString(_e));     \
  } while (0)

#define DL_LAUNCH_CHECK(name)                                                            \
  do {                                                                                   \
    cudaError_t _e = cudaPeekAtLastError();                                              \
    if (_e != cudaSuccess) {                                                             \
      cudaGetLastError();                                                                \
      return dl::set_error((int)_e, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    }                                                                                    \
  } while (0)

int sm_count();   // cached multiProcessorCount of the current device

// ------------------------------------------------------------------ programmatic dependent launch
// Every kernel of the library is launched with the programmatic-stream-serialization attribute
// and begins with pdl_trigger(); pdl_wait():  its blocks may be scheduled (and run their
// prologue: barrier init, TMEM allocation, tensor-map prefetch, shared-memory zeroing) while the
// previous kernel in the stream is still draining; pdl_wait() then blocks until that kernel has
// completed and its writes are visible, so no global access ever overtakes a producer.  A step is
// ~480 launches of 5-20 us each: this hides the launch latency between them.  DL_NO_PDL=1 turns
// the attribute off (the device-side instructions are then no-ops).
bool pdl_enabled();

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// the same with thread-block clusters of `cluster_x` CTAs along x (CTA pairs for cta_group::2 kernels)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, unsigned cluster_x,
                                    cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// DL_LAUNCH((kernel<T>), grid, block, smem, stream, args...): parenthesise templated kernel names
#define DL_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (void)dl::launch_k(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

// ------------------------------------------------------------------ dtype helpers
enum : int { DT_F32 = 0, DT_BF16 = 1 };

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

template <typename T> __device__ __forceinline__ float ldf(const T* p, size_t i) { return Cvt<T>::to_f(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, size_t i, float v) { p[i] = Cvt<T>::from_f(v); }

// 4 consecutive elements <-> float4 (8-byte access for bf16, 16-byte for f32)
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// one 16-byte access <-> 4 (fp32) or 8 (bf16) floats
template <typename T> struct VecWidth { static constexpr int N = 16 / sizeof(T); };
__device__ __forceinline__ void ldv(const float* p, float (&x)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
__device__ __forceinline__ void ldv(const __nv_bfloat16* p, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    x[2 * i] = f.x; x[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void stv(float* p, const float (&x)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
}
__device__ __forceinline__ void stv(__nv_bfloat16* p, const float (&x)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------ math
// erf by Abramowitz & Stegun 7.1.26 (|abs error| <= 1.5e-7, i.e. fp32 rounding level) with one
// ex2 and one fast reciprocal: ~12 instructions instead of erff's ~40 -- the GELU epilogues of the
// FFN GEMMs evaluate it 16.8 M times per launch.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + fast_erf(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// GELU for bf16 activations: the tanh form on the hardware tanh (1 MUFU + 5 FMA-pipe instructions
// instead of 2 MUFU + ~16).  |gelu_tanh - gelu_erf| <= 4.8e-4 and |d/dx difference| <= 8.7e-4 over
// all x -- a fraction of the bf16 rounding step (3.9e-3 relative) the result is stored with.  fp32
// activations keep the erf form above.  T selects by storage type so forward and backward agree.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T> __device__ __forceinline__ float gelu_fwd(float x) {
  if constexpr (sizeof(T) == 2) {
    const float u = x * x;
    const float t = tanh_approx(x * fmaf(u, 0.044715f * 0.7978845608f, 0.7978845608f));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
  } else {
    return gelu_erf(x);
  }
}
// value and derivative together (one tanh): the forward epilogue that stores d/dx for the backward
template <typename T> __device__ __forceinline__ void gelu_fwd_grad(float x, float& y, float& d) {
  if constexpr (sizeof(T) == 2) {
    const float u = x * x;
    const float t = tanh_approx(x * fmaf(u, 0.044715f * 0.7978845608f, 0.7978845608f));
    const float hx = 0.5f * x;
    y = fmaf(hx, t, hx);
    const float dz = fmaf(u, 3.0f * 0.044715f * 0.7978845608f, 0.7978845608f);
    d = fmaf(hx * fmaf(-t, t, 1.0f), dz, fmaf(0.5f, t, 0.5f));
  } else {
    y = gelu_erf(x);
    d = gelu_erf_grad(x);
  }
}
// ---- packed fp32 pairs (sm_100: FFMA2 / FMUL2 do two fp32 lanes per issue slot) -----------------------
// The GEMM epilogues that evaluate GELU + its derivative + two dropout factors per element are bound by
// instruction issue (36 per element before, ncu: profiles/r2m_gemm_top_ncu.md); the floating-point part
// of that goes through these.  Same roundings as the scalar code: fma.rn / mul.rn per lane.
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mul.rn.f32x2 rd, ra, rb;\n"
      " mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }

// value and derivative of GELU for two elements (bf16 storage: tanh form, packed; fp32: the scalar erf form)
template <typename T> __device__ __forceinline__ void gelu_fwd_grad2(float2 x, float2& y, float2& d) {
  if constexpr (sizeof(T) == 2) {
    constexpr float c1 = 0.7978845608f, c2 = 0.044715f * 0.7978845608f;
    const float2 u = f2_mul(x, x);
    const float2 arg = f2_mul(x, f2_fma(u, f2(c2), f2(c1)));
    const float2 t = make_float2(tanh_approx(arg.x), tanh_approx(arg.y));
    const float2 hx = f2_mul(x, f2(0.5f)), nhx = f2_mul(x, f2(-0.5f));
    y = f2_fma(hx, t, hx);
    const float2 dz = f2_fma(u, f2(3.0f * c2), f2(c1));
    const float2 a = f2_fma(f2_mul(nhx, t), t, hx);            // hx (1 - t^2)
    d = f2_fma(a, dz, f2_fma(t, f2(0.5f), f2(0.5f)));
  } else {
    gelu_fwd_grad<T>(x.x, y.x, d.x);
    gelu_fwd_grad<T>(x.y, y.y, d.y);
  }
}
template <typename T> __device__ __forceinline__ float2 gelu_fwd2(float2 x) {
  if constexpr (sizeof(T) == 2) {
    constexpr float c1 = 0.7978845608f, c2 = 0.044715f * 0.7978845608f;
    const float2 u = f2_mul(x, x);
    const float2 arg = f2_mul(x, f2_fma(u, f2(c2), f2(c1)));
    const float2 t = make_float2(tanh_approx(arg.x), tanh_approx(arg.y));
    const float2 hx = f2_mul(x, f2(0.5f));
    return f2_fma(hx, t, hx);
  } else {
    return make_float2(gelu_fwd<T>(x.x), gelu_fwd<T>(x.y));
  }
}

template <typename T> __device__ __forceinline__ float gelu_grad(float x) {
  if constexpr (sizeof(T) == 2) {
    const float u = x * x;
    const float t = tanh_approx(x * fmaf(u, 0.044715f * 0.7978845608f, 0.7978845608f));
    const float dz = fmaf(u, 3.0f * 0.044715f * 0.7978845608f, 0.7978845608f);
    const float a = 0.5f * x * fmaf(-t, t, 1.0f);
    return fmaf(a, dz, fmaf(0.5f, t, 0.5f));
  } else {
    return gelu_erf_grad(x);
  }
}

// Dropout masks are a pure function of (seed, logical element index), so backward regenerates them
// instead of storing them.  One counter-based 32-bit hash decides TWO neighbouring elements (its
// 16-bit halves against a 16-bit threshold; p is honoured to 1/65536 and the survivors are scaled by
// the exact inverse of the realised keep probability): the GEMM epilogues, which are instruction
// bound, pay half a hash per element.
__device__ __forceinline__ uint32_t drop_hash(unsigned long long seed, unsigned long long pair) {
  // 32-bit avalanche (three multiplies) of (seed, pair index)
  uint32_t h = (uint32_t)pair ^ (uint32_t)seed;
  h = (h ^ (uint32_t)(pair >> 32) * 0x9E3779B1u) * 0x85EBCA77u + (uint32_t)(seed >> 32);
  h ^= h >> 15; h *= 0xC2B2AE3Du;
  h ^= h >> 13; h *= 0x27D4EB2Fu;
  h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 65536.f + 0.5f); }
__host__ __device__ __forceinline__ float drop_scale(uint32_t thr) { return 65536.f / (float)(65536u - thr); }
// The effective seed of a launch: the host seed advanced by a device-resident step counter (or
// unchanged when there is none), so a CUDA-graph replay draws a fresh mask every step while the
// forward and the backward of one step -- which read the counter before the optimiser bumps it --
// regenerate the same one.
// The step is mixed in through splitmix64 (not added with the multiplier the host uses to space
// consecutive launch seeds): launch i at step t must not reuse the mask of launch i+1 at step t-1.
__device__ __forceinline__ unsigned long long drop_seed_at(unsigned long long seed, const long long* ctr) {
  if (!ctr) return seed;
  unsigned long long z = seed ^ ((unsigned long long)(*ctr) * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned long long idx, uint32_t thr) {
  const uint32_t h = drop_hash(seed, idx >> 1);
  return ((idx & 1ull) ? (h >> 16) : (h & 0xffffu)) >= thr;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` is a shared float[32]; every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  return warp_max(t);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace dl
