// Probe (GPU): latencies that set the weight-ring slot cycle: bulk copy L2->SMEM, remote mbarrier round trip.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (++spins > (1u << 24)) { printf("timeout\n"); __trap(); }
  }
}
__global__ void __cluster_dims__(2, 1, 1) lat(const uint8_t* src, long long* out, int nblk) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 65536);
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint8_t* my = src + (size_t)(blockIdx.x % nblk) * 65536;
    // (1) bulk copy latency, sizes 4K/16K/32K/64K, second pass (L2 warm)
    int sizes[4] = {4096, 16384, 32768, 65536};
    uint32_t ph = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < 4; ++i) {
        long long t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(sizes[i]) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)), "l"(my), "r"(sizes[i]), "r"(smem_u32(&bars[0])) : "memory");
        mbar_wait(&bars[0], ph); ph ^= 1;
        long long t1 = clock64();
        if (blockIdx.x == 0) out[pass * 4 + i] = t1 - t0;
      }
    // two 16 KB copies in flight at once
    {
      long long t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(32768) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)), "l"(my), "r"(16384), "r"(smem_u32(&bars[0])) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + 16384)), "l"(my + 16384), "r"(16384), "r"(smem_u32(&bars[0])) : "memory");
      mbar_wait(&bars[0], ph); ph ^= 1;
      if (blockIdx.x == 0) out[8] = clock64() - t0;
    }
    // (2) remote arrive round trip: rank0 -> rank1 bar[1]; rank1 -> rank0 bar[2]
    if (rank == 0) {
      long long t0 = clock64();
      for (int r = 0; r < 8; ++r) {
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(smem_u32(&bars[1]), 1)) : "memory");
        mbar_wait(&bars[2], r & 1);
      }
      if (blockIdx.x == 0) out[9] = (clock64() - t0) / 8;
    } else {
      for (int r = 0; r < 8; ++r) {
        mbar_wait(&bars[1], r & 1);
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(smem_u32(&bars[2]), 0)) : "memory");
      }
    }
    // (3) local arrive + wait round trip (same thread)
    {
      long long t0 = clock64();
      for (int r = 0; r < 8; ++r) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[3])) : "memory");
        mbar_wait(&bars[3], r & 1);
      }
      if (blockIdx.x == 0) out[10] = (clock64() - t0) / 8;
    }
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  uint8_t* src; long long* out;
  const int nblk = 64;
  cudaMalloc(&src, (size_t)nblk * 65536); cudaMemset(src, 1, (size_t)nblk * 65536);
  cudaMalloc(&out, 16 * 8); cudaMemset(out, 0, 128);
  cudaFuncSetAttribute(lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 128);
  for (int grid : {2, 148}) {
    lat<<<grid, 32, 65536 + 128>>>(src, out, nblk);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, out, 128, cudaMemcpyDeviceToHost);
    printf("grid %3d (%s): bulk copy cycles cold 4K/16K/32K/64K = %lld %lld %lld %lld | warm = %lld %lld %lld %lld | 2x16K together %lld | remote arrive RTT %lld | local arrive+wait %lld\n",
           grid, cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10]);
  }
  return 0;
}
