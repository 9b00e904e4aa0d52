// Does a tcgen05.commit between groups of MMAs cost tensor-pipe time?  R groups of G MMAs, a commit after every
// group onto a ring of 8 barriers, no waiting in between (one wait at the end).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../druglamp_b200/csrc/ptx.cuh"
using namespace dl;

__device__ __forceinline__ void commit_2sm_single(uint32_t bar) {   // no multicast: arrives on one CTA's barrier
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// mode 0: cta_group::1; 1: cta_group::2 + multicast commit; 2: cta_group::2 + single-CTA commit; 3: cta_group::2, commit only at the end
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(int rounds, int G, long long* out) {
  constexpr bool CTA2 = MODE != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int STG = 49152, NST = 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * STG);
  const uint32_t bar = ptx::smem_u32(bars);          // [8] ring + [1] final
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5;
  const int rank = CTA2 ? (int)ptx::cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < NST * STG / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(bar + 8 * i, 1 << 19); ptx::mbar_init(bar + 64, 1); ptx::fence_barrier_init(); }
  if (warp == 1) { if constexpr (CTA2) ptx::tmem_alloc_2sm<512>(ptx::smem_u32(slot)); else ptx::tmem_alloc<512>(ptx::smem_u32(slot)); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 32 && rank == 0) {
    const uint32_t idesc = ptx::make_idesc(false, false, false, CTA2 ? 256 : 128, 256);
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      const uint32_t sA = base + (r % NST) * STG, sB = sA + 16384;
      for (int m = 0; m < G; ++m) {
        const int k = m & 3;
        const uint64_t ad = ptx::make_smem_desc(sA + k * 32, 16, 1024), bd = ptx::make_smem_desc(sB + k * 32, 16, 1024);
        if constexpr (CTA2) ptx::mma_ss_2sm(tmem, ad, bd, idesc, 1u); else ptx::mma_ss<false>(tmem, ad, bd, idesc, 1u);
      }
      const uint32_t b = bar + 8 * (r & 7);
      if constexpr (MODE == 0) ptx::mma_commit(b);
      else if constexpr (MODE == 1) ptx::mma_commit_2sm(b);
      else if constexpr (MODE == 2) commit_2sm_single(b);
    }
    if constexpr (MODE == 0) ptx::mma_commit(bar + 64); else ptx::mma_commit_2sm(bar + 64);
    ptx::mbar_wait(bar + 64, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  if constexpr (CTA2) { if (threadIdx.x == 32 && rank == 1) ptx::mbar_wait(bar + 64, 0); }
  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); if constexpr (CTA2) ptx::tmem_dealloc_2sm<512>(tmem); else ptx::tmem_dealloc<512>(tmem); }
}

template <int MODE>
void run(int grid, int rounds, int G) {
  long long* d; cudaMalloc(&d, grid * sizeof(long long)); cudaMemset(d, 0, grid * sizeof(long long));
  const int smem = 4 * 49152 + 128 + 1024;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = MODE ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe<MODE>, rounds, G, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[512]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  printf("mode %d grid %3d: %d groups of %d MMAs: %.1f cycles per MMA, %.0f per group  (%s / %s)\n", MODE, grid, rounds, G,
         (double)mx / ((double)rounds * G), (double)mx / rounds, cudaGetErrorString(e), cudaGetErrorString(e2));
  cudaFree(d);
}

int main() {
  for (int G : {4, 8}) {
    run<0>(148, 512, G); run<1>(148, 512, G); run<2>(148, 512, G); run<3>(148, 512, G);
  }
  return 0;
}
