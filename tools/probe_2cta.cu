// Probe (GPU, standalone): semantics of tcgen05.mma.cta_group::2 needed by the paired sub-band kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_2cta probe_2cta.cu && ./probe_2cta
// Checks: 2-CTA TMEM alloc, M=256 and M=128 (cta_group::2) accumulator layouts, B operand written into the
// PEER's shared memory with st.shared::cluster (+ fence.proxy.async) and published with a remote mbarrier
// arrive, multicast tcgen05.commit to both CTAs.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (++spins > (1u << 24)) { printf("probe: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SW128
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ __forceinline__ int swz_off(int row, int k) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2;
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

struct Sm {
  uint8_t a[128 * 128];   // A half: 128 rows x 64 k fp16, SW128
  uint8_t b[128 * 128];   // B half: up to 128 rows x 64 k (rows >= 32 are only used by the timing runs)
  uint64_t ready, done1, done2, done3;
  long long cyc[16];
  uint32_t tmem;
};

// A [256][64], B [64][64] fp16 row-major in global; out1 [2][128][64], out2 [2][128][32]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __half* A, const __half* B, float* out1, float* out2) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  Sm& s = *reinterpret_cast<Sm*>(base);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&s.ready, 2 * 128);
    mbar_init(&s.done1, 1);
    mbar_init(&s.done2, 1);
    mbar_init(&s.done3, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s.tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  // own A half -> own smem
  for (int i = tid; i < 128 * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<__half*>(s.a + swz_off(r, k)) = A[(size_t)(rank * 128 + r) * 64 + k];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s.tmem;
  // the PEER's B half -> the peer's smem through DSMEM (as the paired kernel's epilogue will do with h)
  const uint32_t peer = rank ^ 1u;
  const uint32_t b_peer = mapa(smem_u32(s.b), peer);
  for (int i = tid; i < 32 * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    const __half v = B[(size_t)(peer * 32 + r) * 64 + k];
    asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(b_peer + swz_off(r, k)), "h"(__half_as_ushort(v)) : "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  mbar_arrive_cluster(mapa(smem_u32(&s.ready), 0));  // every thread of both CTAs arrives on the LEADER's barrier

  if (rank == 0 && warp == 1) {
    mbar_wait(&s.ready, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (lane == 0) {
      const uint64_t ad = make_desc(smem_u32(s.a)), bd = make_desc(smem_u32(s.b));
      // (1) M=256, N=64: each CTA gets its 128 rows x 64 columns
      const uint32_t idesc256 = (1u << 4) | ((64u >> 3) << 17) | ((256u >> 4) << 24);
      for (int k = 0; k < 4; ++k) mma2(tmem, ad + 2 * k, bd + 2 * k, idesc256, k ? 1u : 0u);
      commit_mc(&s.done1, 3);
      // (2) M=128 (64 rows per CTA = first 64 rows of each A half), N=64, at TMEM column 64
      const uint32_t idesc128 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      for (int k = 0; k < 4; ++k) mma2(tmem + 64, ad + 2 * k, bd + 2 * k, idesc128, k ? 1u : 0u);
      commit_mc(&s.done2, 3);
      // (3) timing: 128 MMAs round-robin over 4 accumulators (as the LSTM kernel issues them), then one commit
      uint32_t ph3 = 0;
      int cfg = 0;
      for (int Mm = 256; Mm >= 128; Mm -= 128)
        for (int Nn = 64; Nn <= 256; Nn *= 2) {
          const uint32_t idesc = (1u << 4) | (((uint32_t)Nn >> 3) << 17) | (((uint32_t)Mm >> 4) << 24);
          const uint32_t dstep = (Mm == 256) ? Nn : Nn / 2;
          const int nacc = (4 * dstep <= 512) ? 4 : (512 / dstep);
          long long t0 = clock64();
          for (int i = 0; i < 128; ++i) mma2(tmem + (i % nacc) * dstep, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s.done3)) : "memory");
          mbar_wait(&s.done3, ph3); ph3 ^= 1;
          s.cyc[cfg++] = clock64() - t0;
        }
      for (int i = 0; i < cfg; ++i) printf("timing cfg %d: %lld cycles for 128 MMAs = %lld per MMA\n", i, s.cyc[i], s.cyc[i] / 128);
    }
  }
  mbar_wait(&s.done1, 0);
  mbar_wait(&s.done2, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out1[((size_t)rank * 128 + tid) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + 64 + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out2[((size_t)rank * 128 + tid) * 32 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  std::vector<__half> A(256 * 64), B(64 * 64);
  std::vector<float> Af(256 * 64), Bf(64 * 64);
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { Af[i] = (rand() % 200 - 100) / 64.0f; A[i] = __float2half(Af[i]); Af[i] = __half2float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (rand() % 200 - 100) / 64.0f; B[i] = __float2half(Bf[i]); Bf[i] = __half2float(B[i]); }
  std::vector<float> D(256 * 64);
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 64; ++n) {
      float s = 0;
      for (int k = 0; k < 64; ++k) s += Af[m * 64 + k] * Bf[n * 64 + k];
      D[m * 64 + n] = s;
    }
  __half *dA, *dB; float *o1, *o2;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2);
  cudaMalloc(&o1, 2 * 128 * 64 * 4); cudaMalloc(&o2, 2 * 128 * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(o1, 0, 2 * 128 * 64 * 4); cudaMemset(o2, 0, 2 * 128 * 32 * 4);
  const int smem = sizeof(Sm) + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<2, 128, smem>>>(dA, dB, o1, o2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> h1(2 * 128 * 64), h2(2 * 128 * 32);
  cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost);
  // (1) expect out1[rank][lane][n] = D[rank*128 + lane][n]
  double e1 = 0;
  for (int r = 0; r < 2; ++r)
    for (int l = 0; l < 128; ++l)
      for (int n = 0; n < 64; ++n) e1 = fmax(e1, fabs(h1[(r * 128 + l) * 64 + n] - D[(r * 128 + l) * 64 + n]));
  printf("M=256: max err vs lane=row, col=n : %g\n", e1);
  // (2) M=128 over the pair: logical rows: CTA r holds rows r*64..r*64+63 where logical row i of CTA r = A row r*128 + i
  // candidate layout (cute 2x2 atom): lane = m + 64*(n / 32), col = n % 32
  double e2 = 0;
  for (int r = 0; r < 2; ++r)
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < 64; ++n) {
        const int lane = m + 64 * (n / 32), col = n % 32;
        e2 = fmax(e2, fabs(h2[(r * 128 + lane) * 32 + col] - D[(r * 128 + m) * 64 + n]));
      }
  printf("M=128 (2 CTA): max err vs lane = m + 64*(n/32), col = n%%32 : %g\n", e2);
  if (e2 > 1e-2) {  // print a few values to infer the layout
    for (int r = 0; r < 2; ++r)
      for (int l = 0; l < 128; l += 16) {
        printf("r%d lane %3d:", r, l);
        for (int c = 0; c < 4; ++c) printf(" %8.3f", h2[(r * 128 + l) * 32 + c]);
        printf("   | D[%d][0..3] =", r * 128 + (l % 64));
        for (int c = 0; c < 4; ++c) printf(" %8.3f", D[(r * 128 + (l % 64)) * 64 + c]);
        printf("\n");
      }
  }
  return 0;
}
