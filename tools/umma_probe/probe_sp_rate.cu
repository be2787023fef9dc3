// Developer probe: issue rate of tcgen05.mma (dense, K = 16) against tcgen05.mma.sp (2:4 sparse A, K = 32) for kind::f16,
// CTA pairs (cta_group::2, M = 256, N = 256), operands resident in shared memory, every pair of the grid at once.
// Prints clocks per instruction seen by the issuing thread and the dense-equivalent PFLOP/s of the whole grid.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate(int sparse, int iters, int n, long long* clocks, int tf32)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  unsigned char* sa = smem; unsigned char* sb = smem + 16384;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x;   // bf16 pairs with small exponents (finite, varied bits)
    ((uint32_t*)smem)[i] = (h & 0x007F007Fu) | 0x3F003F00u | (h & 0x80008000u);
  }
  if (0 == tid) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (0 == warp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = slot;
  {  // metadata columns 256..259 of this CTA: kept elements at positions 0, 1
    const uint32_t t = tm + ((uint32_t)(warp * 32) << 16) + 256;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(t), "r"(0x44444444u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (0 == rank && 0 == tid) {
    // bf16: A K-major, B MN-major; tf32: both K-major (32-byte k-steps inside 128-byte rows)
    const uint32_t base = tf32 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24))
                               : ((1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int ks = it & 3;
      const uint64_t da = mkdesc(smem_u32(sa) + ks * 32, 16, 1024, 2);
      if (tf32) {
        const uint64_t db = mkdesc(smem_u32(sb) + ks * (sparse ? 64 : 32), 16, 1024, 2);   // B K-major: 128 n-rows x 128 B
        if (sparse) {
          const uint32_t te = tm + 256 + (ks & ~1), idesc = base | (1u << 2) | (uint32_t)(ks & 1);
          asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\ttcgen05.mma.sp.cta_group::2.kind::tf32 [%0], %1, %2, [%3], %4, q;\n\t}\n"
                       ::"r"(tm), "l"(da), "l"(db), "r"(te), "r"(idesc), "r"(it > 0 ? 1u : 0u) : "memory");
        }
        else asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, q;\n\t}\n"
                          ::"r"(tm), "l"(da), "l"(db), "r"(base), "r"(it > 0 ? 1u : 0u) : "memory");
      }
      else if (sparse) {
        const uint64_t db = mkdesc(smem_u32(sb) + ks * 4096, 16384, 1024, 2);   // B: [2 column blocks of 64][128 k][128 B]
        const uint32_t te = tm + 256 + (ks & ~1), idesc = base | (1u << 2) | (uint32_t)(ks & 1);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\ttcgen05.mma.sp.cta_group::2.kind::f16 [%0], %1, %2, [%3], %4, q;\n\t}\n"
                     ::"r"(tm), "l"(da), "l"(db), "r"(te), "r"(idesc), "r"(it > 0 ? 1u : 0u) : "memory");
      }
      else {
        const uint64_t db = mkdesc(smem_u32(sb) + ks * 2048, 8192, 1024, 2);    // B: [2 column blocks of 64][64 k][128 B]
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n\t}\n"
                     ::"r"(tm), "l"(da), "l"(db), "r"(base), "r"(it > 0 ? 1u : 0u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    clocks[blockIdx.x >> 1] = clock64() - t0;
  }
  else if (0 == tid) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  if (0 == warp) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512) : "memory");
}
int main(int argc, char** argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  const int SM = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  long long* dclk; cudaMalloc(&dclk, 128 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int tf32 = 0; tf32 < 2; ++tf32) for (int pairs : { 74 }) for (int n : { 256 }) for (int sparse = 0; sparse < 2; ++sparse) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      rate<<<2 * pairs, 128, SM>>>(sparse, iters, n, dclk, tf32);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
      if (0 == rep) continue;
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long clk; cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
      const double flops = 2.0 * 256 * n * (sparse ? 32 : 16) * (tf32 ? 0.5 : 1.0) * (double)iters * pairs;
      printf("%s pairs=%2d N=%d %s: %.1f clocks per MMA, %.3f ms, dense-equivalent %.3f PFLOP/s\n", tf32 ? "tf32" : "bf16", pairs, n, sparse ? "sparse (2x K)" : "dense       ",
             (double)clk / iters, ms, flops / (ms * 1e-3) * 1e-15);
    }
  }
  return 0;
}
