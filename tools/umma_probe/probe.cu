// Developer probe: one 128x128x32 TF32 tcgen05.mma chain with hand-filled shared-memory operands, to pin the
// shared-memory descriptor conventions (K-major / MN-major, LBO / SBO) used by spmdm_compute_tc.cu.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 2)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
struct P { int b_mode; uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, idesc; uint32_t b_layout, f_n, f_k; };
__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, P p)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sa = smem; unsigned char* sb = smem + 16384;
  for (int i = tid; i < 32768 / 4; i += 128) ((uint32_t*)smem)[i] = 0;
  __syncthreads();
  for (int i = tid; i < 128 * 32; i += 128) {   // A[m][k], K-major SW128
    const int m = i / 32, k = i % 32;
    const uint32_t off = (m >> 3) * 1024 + (m & 7) * 128 + ((((k >> 2) ^ (m & 7)) & 7) << 4) + ((k & 3) << 2);
    *(float*)(sa + off) = A[m * 32 + k];
  }
  for (int i = tid; i < 32 * 128; i += 128) {   // B[k][n]
    const int k = i / 128, n = i % 128;
    uint32_t off;
    if (2 == p.b_mode) off = (n >> 5) * p.f_n + (k >> 2) * p.f_k + (k & 3) * 128 + (((((n & 31) >> 3) ^ (k & 3)) & 3) << 5) + ((n & 7) << 2);   // MN-major, 128B swizzle with 32B atoms
    else if (1 == p.b_mode) off = (n >> 5) * 4096 + k * 128 + (((((n & 31) >> 2) ^ (k & 7)) & 7) << 4) + ((n & 3) << 2);   // MN-major
    else off = (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 2) ^ (n & 7)) & 7) << 4) + ((k & 3) << 2);                   // K-major
    *(float*)(sb + off) = B[k * 128 + n];
  }
  if (0 == tid) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (0 == warp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = slot;
  if (0 == tid) {
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t da = mkdesc(smem_u32(sa) + ks * p.a_kstep, p.a_lbo, p.a_sbo);
      const uint64_t db = mkdesc(smem_u32(sb) + ks * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}\n"
                   ::"r"(tm), "l"(da), "l"(db), "r"(p.idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int cb = 0; cb < 128; cb += 32) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(tm + ((uint32_t)(warp * 32) << 16) + cb));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + cb + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (0 == warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(128) : "memory");
}
int main()
{
  std::vector<float> A(128 * 32), B(32 * 128), D(128 * 128), E(128 * 128, 0.f);
  for (int i = 0; i < 128 * 32; ++i) A[i] = (float)((i * 7 + 3) % 11 - 5);
  for (int i = 0; i < 32 * 128; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += A[m * 32 + k] * B[k * 128 + n]; E[m * 128 + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024);
  const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);
  struct { const char* name; P p; } cases[] = {
    { "B MN SW128_32B desc lbo=4096 sbo=512 ", { 2, 16, 1024, 4096, 512, 32, 1024, base | (1u << 16), 1, 4096, 512 } },
    { "B MN SW128_32B desc lbo=512 sbo=4096 ", { 2, 16, 1024, 512, 4096, 32, 1024, base | (1u << 16), 1, 4096, 512 } },
    { "B MN SW128_32B kgroup-major fill     ", { 2, 16, 1024, 512, 2048, 32, 4096, base | (1u << 16), 1, 512, 2048 } },
    { "B MN SW128_32B kgroup-major, swapped ", { 2, 16, 1024, 2048, 512, 32, 4096, base | (1u << 16), 1, 512, 2048 } },
    { "B K-major  lbo=16   sbo=1024 kstep=32", { 0, 16, 1024, 16, 1024, 32, 32, base, 2, 0, 0 } },
  };
  for (auto& c : cases) {
    cudaMemset(dD, 0xFF, D.size() * 4);
    probe<<<1, 128, 32768 + 1024>>>(dA, dB, dD, c.p);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 128 * 128; ++i) { double d = fabs((double)D[i] - E[i]); if (!(d <= 1e-3)) ++bad; if (d > maxerr) maxerr = d; }
    printf("%s: %s  bad=%d maxerr=%g  D[0..3]=%g %g %g %g  E=%g %g %g %g  D[5*128+77]=%g E=%g\n", c.name, cudaGetErrorString(e), bad, maxerr,
           D[0], D[1], D[2], D[3], E[0], E[1], E[2], E[3], D[5 * 128 + 77], E[5 * 128 + 77]);
  }
  return 0;
}
