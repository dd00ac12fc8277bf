// Probe: thread <-> (lane, column) mapping of tcgen05.ld.16x256b, incl. the lane offset 16 inside a warp's quadrant.
// TMEM is filled with value = lane*1000 + column through tcgen05.st.32x32b, then read back with 16x256b.x2 (16 columns).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(float* out) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
  // fill: lane L = warp*32 + lane, columns 0..15
  for (int c = 0; c < 16; ++c) {
    float v = (float)((warp * 32 + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(__float_as_uint(v)));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    const uint32_t a = taddr + ((uint32_t)(half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[((warp * 2 + half) * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb));
}
int main() {
  float* d; cudaMalloc(&d, 4 * 2 * 32 * 8 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  static float h[4 * 2 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 2; ++w) for (int half = 0; half < 2; ++half) for (int l = 0; l < 32; l += 1) {
    if (!(l < 6 || l == 31)) continue;
    printf("warp %d half %d thread %2d:", w, half, l);
    for (int i = 0; i < 8; ++i) printf(" %6.0f", h[((w * 2 + half) * 32 + l) * 8 + i]);
    printf("\n");
  }
  return 0;
}
