// Raw tcgen05.mma issue rate, cta_group::1 (128 x 256 x 16) vs cta_group::2 (256 x 256 x 16 over a CTA pair),
// operands resident in shared memory (no loads): cycles per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cta2_mma_rate cta2_mma_rate.cu && ./cta2_mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../druglamp_b200/csrc/ptx.cuh"
using namespace dl;

template <bool CTA2>
__global__ void __launch_bounds__(128, 1) probe(int iters, int nmma, long long* out, int pingpong = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + 16384;             // A: 128 x 64 bf16, B: 256 (or 128) x 64 bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  const uint32_t bar = ptx::smem_u32(bars);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5;
  const int rank = CTA2 ? (int)ptx::cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const uint32_t bar2 = bar + 8;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_barrier_init(); }
  if (warp == 1) { if constexpr (CTA2) ptx::tmem_alloc_2sm<512>(ptx::smem_u32(slot)); else ptx::tmem_alloc<512>(ptx::smem_u32(slot)); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 32 && rank == 0) {
    const uint32_t idesc = ptx::make_idesc(false, false, false, CTA2 ? 256 : 128, 256);
    uint32_t ph = 0;
    long long t0 = 0;
    for (int it = 0; it < iters + 1; ++it) {
      if (it == 1) t0 = clock64();
      for (int m = 0; m < nmma; ++m) {
        const int k = m & 3;
        const uint64_t ad = ptx::make_smem_desc(sA + k * 32, 16, 1024), bd = ptx::make_smem_desc(sB + k * 32, 16, 1024);
        if constexpr (CTA2) ptx::mma_ss_2sm(tmem + (it & 1) * 256, ad, bd, idesc, (uint32_t)(m != 0));
        else ptx::mma_ss<false>(tmem + (it & 1) * 256, ad, bd, idesc, (uint32_t)(m != 0));
      }
      if constexpr (CTA2) ptx::mma_commit_2sm(bar); else ptx::mma_commit(bar);
      ptx::mbar_wait(pingpong ? bar2 : bar, ph);      // pingpong: via the peer's thread and its remote arrive
      ph ^= 1u;
    }
    out[blockIdx.x] = clock64() - t0;
  }
  if constexpr (CTA2) { if (threadIdx.x == 32 && rank == 1) { uint32_t ph = 0; for (int it = 0; it < iters + 1; ++it) { ptx::mbar_wait(bar, ph); ph ^= 1u; if (pingpong) ptx::mbar_arrive_remote(ptx::mapa(bar2, 0)); } } }
  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); if constexpr (CTA2) ptx::tmem_dealloc_2sm<512>(tmem); else ptx::tmem_dealloc<512>(tmem); }
}

template <bool CTA2>
void run(int grid, int iters, int nmma, int pingpong = 0) {
  long long* d; cudaMalloc(&d, grid * sizeof(long long)); cudaMemset(d, 0, grid * sizeof(long long));
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe<CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTA2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CTA2>, iters, nmma, d, pingpong);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[512]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  printf("cta_group::%d grid %3d pingpong %d: %d x %d MMAs (N=256, K=16): %.1f cycles per MMA, %.0f per round  (%s / %s)\n", CTA2 ? 2 : 1, grid, pingpong, iters, nmma,
         (double)mx / ((double)iters * nmma), (double)mx / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
  cudaFree(d);
}

int main() {
  for (int nmma : {4, 16, 64}) {
    run<false>(1, 200, nmma); run<false>(148, 200, nmma);
    run<true>(2, 200, nmma); run<true>(148, 200, nmma); run<true>(148, 200, nmma, 1);
  }
  return 0;
}
