// Micro-benchmark: bandwidth of cp.async.bulk shared::cta -> shared::cluster (peer CTA of a pair), 8 warps each
// streaming 4 KB blocks, completion counted on the DESTINATION CTA's mbarrier (complete_tx).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__global__ void __cluster_dims__(2, 1, 1) k(float* out, int iters, int blk_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[8];
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
  asm volatile("fence.mbarrier_init.release.cluster;");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  long long t0 = clock64();
  uint8_t* mine = smem + warp * 2 * blk_bytes;
  if (rank == 1) {
    // producer: send `iters` blocks to the peer's slot (2 slots alternate; no back-pressure: peak rate)
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        uint32_t src = smem_u32(mine + (i & 1) * blk_bytes);
        uint32_t dst = mapa(src, 0), mb = mapa(smem_u32(&bar[warp]), 0);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(src), "r"(blk_bytes), "r"(mb) : "memory");
      }
    }
  } else {
    // consumer: expect all bytes in `iters` phases
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp])), "r"(blk_bytes) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(i & 1) : "memory");
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
  float* d; cudaMalloc(&d, 64);
  for (int blk : {2048, 4096, 8192}) {
    int smem = 8 * 2 * blk;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int iters = 2000;
    k<<<148, 256, smem>>>(d, iters, blk); cudaDeviceSynchronize();
    k<<<148, 256, smem>>>(d, iters, blk);
    cudaError_t e = cudaDeviceSynchronize();
    float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    printf("block %5d B x 8 warps: %.1f B/clk per receiving SM (%s)\n", blk, 8.0 * iters * blk / cyc, cudaGetErrorString(e));
  }
  return 0;
}
