// Developer probe: tcgen05.mma.sp kind::f16 (2:4 structured-sparse A), single CTA, M = 128, N = 128, four K = 32 atoms.
// A: compressed, K-major SWIZZLE_128B (128 rows x 64 kept bf16 = one 128-byte swizzle row per A row and 128 logical k);
// B: MN-major SWIZZLE_128B, 128 k x 128 columns; metadata: four TMEM columns written with tcgen05.st.
// Phase 1 checks the hypothesised metadata layout (CUTLASS cute/atom/mma_traits_sm100.hpp, tmem_e_frg, read not included)
// on random 2:4 patterns; phase 2 maps every (lane, column, nibble) of the metadata to the (row, group) it steers by
// multiplying with an identity B, so that the layout can be read off the output whatever it is.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
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
constexpr int MCOL = 128;   // metadata columns start here (accumulator: columns 0..127)
__device__ __forceinline__ void st_meta(uint32_t taddr, const uint32_t (&w)[4])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
    : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// mode 0: one multiply with the metadata words given per lane (meta[lane*4 + c]); D written out.
// mode 1: discovery: experiment e = (lane L, column c, nibble q) sets that nibble to 0xE; res[e] = 1 + row*32 + group that changed
// (0: none), cnt[e] = how many (row, group) cells changed.  mode 2: nibble semantics: lane 0, column 0, nibble 0 takes every value v;
// pat[v*4 + i] = D[row r0][4*g0 + i] for the (r0, g0) found by experiment (0, 0, 0).
struct P { uint32_t idesc; int mode; int id2_from_col; uint32_t r0g0; int meta_by_cp; };
__global__ void __launch_bounds__(128, 1) probe(const uint16_t* Ac, const uint16_t* B, const uint32_t* meta, float* D, uint32_t* res, uint32_t* cnt, float* pat, P p)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sa = smem; unsigned char* sb = smem + 16384;
  for (int i = tid; i < 128 * 64; i += 128) {   // compressed A[m][j], K-major SW128: rows of 64 bf16
    const int m = i / 64, k = i % 64;
    const uint32_t off = (m >> 3) * 1024 + (m & 7) * 128 + ((((k >> 3) ^ (m & 7)) & 7) << 4) + ((k & 7) << 1);
    *(uint16_t*)(sa + off) = Ac[m * 64 + k];
  }
  for (int i = tid; i < 128 * 128; i += 128) {  // B[k][n], MN-major SW128: [n block of 64][128 k][128 B]
    const int k = i / 128, n = i % 128;
    const uint32_t off = (n >> 6) * 16384 + k * 128 + (((((n & 63) >> 3) ^ (k & 7)) & 7) << 4) + ((n & 7) << 1);
    *(uint16_t*)(sb + off) = B[k * 128 + n];
  }
  if (0 == tid) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (0 == warp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = slot;
  const uint32_t my_t = tm + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;
  unsigned char* simg = smem + 16384 + 32768;      // metadata image: lane L at byte 16 * L (four 32-bit columns)
  auto multiply = [&](const uint32_t (&w)[4]) {
    if (p.meta_by_cp) {
      *(uint4*)(simg + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    else st_meta(my_t + MCOL, w);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (0 == tid) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (p.meta_by_cp) {   // 128 lanes x 128 bits from shared memory: no swizzle, rows of 16 bytes, 8-row groups 128 bytes apart
        const uint64_t dm = mkdesc(smem_u32(simg), 16, 128, 0);
        asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;\n" ::"r"(tm + MCOL), "l"(dm) : "memory");
      }
      for (int ks = 0; ks < 4; ++ks) {   // K = 32 logical (16 kept) per MMA
        const uint64_t da = mkdesc(smem_u32(sa) + ks * 32, 16, 1024, 2);
        const uint64_t db = mkdesc(smem_u32(sb) + ks * 4096, 16384, 1024, 2);
        const uint32_t acc = ks > 0 ? 1u : 0u;
        uint32_t te = tm + MCOL + ks, idesc = p.idesc;
        if (p.id2_from_col) { idesc |= (te & 1u); te &= ~1u; }
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\ttcgen05.mma.sp.cta_group::1.kind::f16 [%0], %1, %2, [%3], %4, q;\n\t}\n"
                     ::"r"(tm), "l"(da), "l"(db), "r"(te), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  };
  uint32_t w[4];
  if (0 == p.mode) {
    for (int c = 0; c < 4; ++c) w[c] = meta[tid * 4 + c];
    multiply(w);
    for (int cb = 0; cb < 128; cb += 32) {
      uint32_t r[32];
      ld32(my_t + cb, r);
      for (int j = 0; j < 32; ++j) D[tid * 128 + cb + j] = __uint_as_float(r[j]);
    }
  }
  else {
    // baseline: every nibble 0x4 (kept elements at positions 0 and 1 of their group); B = identity, A_c[m][j] = j + 1
    const int n_exp = (1 == p.mode) ? 128 * 4 * 8 : 16;
    float basev[128];                        // this row's output under the baseline metadata (whatever 0x4 means)
    for (int c = 0; c < 4; ++c) w[c] = 0x44444444u;
    multiply(w);
    for (int cb = 0; cb < 128; cb += 32) { uint32_t r[32]; ld32(my_t + cb, r); for (int j = 0; j < 32; ++j) basev[cb + j] = __uint_as_float(r[j]); }
    if (1 == p.mode && tid < 2) for (int j = 0; j < 16; ++j) pat[tid * 16 + j] = basev[j];
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    for (int e = 0; e < n_exp; ++e) {
      for (int c = 0; c < 4; ++c) w[c] = 0x44444444u;
      int L, c, q; uint32_t v;
      if (1 == p.mode) { L = e >> 5; c = (e >> 3) & 3; q = e & 7; v = 0xEu; }
      else { L = 0; c = 0; q = 0; v = (uint32_t)e; }
      if (tid == L) w[c] = (w[c] & ~(0xFu << (4 * q))) | (v << (4 * q));
      multiply(w);
      uint32_t changed = 0, first = 0;
      for (int cb = 0; cb < 128; cb += 32) {
        uint32_t r[32];
        ld32(my_t + cb, r);
        if (1 == p.mode) {
          for (int gl = 0; gl < 8; ++gl) {
            const int g = (cb >> 2) + gl;
            bool same = true;
            for (int i = 0; i < 4; ++i) same = same && __uint_as_float(r[4 * gl + i]) == basev[4 * g + i];
            if (!same) { if (0 == changed) first = 1u + (uint32_t)tid * 32u + (uint32_t)g; ++changed; }
          }
        }
        else {
          const uint32_t r0 = (p.r0g0 - 1) >> 5, g0 = (p.r0g0 - 1) & 31;
          if ((uint32_t)tid == r0 && (uint32_t)cb == (g0 >> 3) * 32) for (int i = 0; i < 4; ++i) pat[e * 4 + i] = __uint_as_float(r[4 * (g0 & 7) + i]);
        }
      }
      if (1 == p.mode && changed) { atomicAdd(&cnt[e], changed); res[e] = first; }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (0 == warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(256) : "memory");
}
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
int main()
{
  const int M = 128, K = 128, N = 128, KC = 64;
  std::vector<uint16_t> Ac(M * KC), B(K * N);
  std::vector<float> Ad(M * K, 0.f), Bf(K * N), D(M * N), E(M * N, 0.f);
  std::vector<uint32_t> meta(128 * 4, 0);
  // random 2:4 pattern; hypothesised layout: lane = m%8 + 8*((k/16)%2) + 16*(m/16), column = k/32,
  // bit = 16*((m/8)%2) + 4*((k%16)/4); nibble = idx0 | idx1 << 2 with idx0 < idx1
  uint32_t rng = 12345u;
  auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
  for (int m = 0; m < M; ++m) for (int g = 0; g < K / 4; ++g) {
    int i0 = rnd() % 4, i1 = rnd() % 4;
    if (i0 == i1) i1 = (i0 + 1) % 4;
    if (i0 > i1) { int t = i0; i0 = i1; i1 = t; }
    const float v0 = (float)((int)(rnd() % 9) - 4), v1 = (float)((int)(rnd() % 9) - 4);
    Ac[m * KC + 2 * g] = f2bf(v0); Ac[m * KC + 2 * g + 1] = f2bf(v1);
    Ad[m * K + 4 * g + i0] = v0; Ad[m * K + 4 * g + i1] = v1;
    const int k = 4 * g;
    const int lane = (m % 8) + 8 * ((k / 16) % 2) + 16 * (m / 16), col = k / 32, bit = 16 * ((m / 8) % 2) + 4 * ((k % 16) / 4);
    meta[lane * 4 + col] |= (uint32_t)(i0 | (i1 << 2)) << bit;
  }
  for (int i = 0; i < K * N; ++i) { Bf[i] = (float)((i * 5 + 1) % 13 - 6); B[i] = f2bf(Bf[i]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += Ad[m * K + k] * Bf[k * N + n]; E[m * N + n] = s; }
  uint16_t *dA, *dB; float *dD, *dPat; uint32_t *dMeta, *dRes, *dCnt;
  cudaMalloc(&dA, Ac.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dMeta, meta.size() * 4);
  cudaMalloc(&dRes, 4096 * 4); cudaMalloc(&dCnt, 4096 * 4); cudaMalloc(&dPat, 64 * 4);
  cudaMemcpy(dA, Ac.data(), Ac.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dMeta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice);
  const int SM = 16384 + 32768 + 2048 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  const uint32_t base = (1u << 2) | (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  for (int id2c = 1; id2c >= -1; --id2c) {
    cudaMemset(dD, 0xFF, D.size() * 4);
    if (0 == id2c) continue;     // an odd metadata column as address faults (seen once; kept out of the run)
    P p = { base, 0, id2c < 0 ? 1 : id2c, 0, id2c < 0 ? 1 : 0 };
    if (id2c < 0) printf("metadata through tcgen05.cp from a shared-memory image: ");
    probe<<<1, 128, SM>>>(dA, dB, dMeta, dD, dRes, dCnt, dPat, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("phase 1 (id2_from_col=%d): %s\n", id2c, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs((double)D[i] - E[i]); if (!(d <= 1e-3)) ++bad; if (d > maxerr) maxerr = d; }
    printf("phase 1 hypothesis (id2_from_col=%d): bad=%d of %d maxerr=%g  D[0..3]=%g %g %g %g  E=%g %g %g %g\n", id2c, bad, M * N, maxerr, D[0], D[1], D[2], D[3], E[0], E[1], E[2], E[3]);
  }
  // phase 2: identity B, A_c[m][j] = j + 1
  for (int m = 0; m < M; ++m) for (int j = 0; j < KC; ++j) Ac[m * KC + j] = f2bf((float)(j + 1));
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) B[k * N + n] = f2bf(k == n ? 1.f : 0.f);
  cudaMemcpy(dA, Ac.data(), Ac.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  for (int id2c = 1; id2c >= 1; --id2c) {
    cudaMemset(dRes, 0, 4096 * 4); cudaMemset(dCnt, 0, 4096 * 4); cudaMemset(dPat, 0, 64 * 4);
    P p = { base, 1, id2c, 0, 0 };
    probe<<<1, 128, SM>>>(dA, dB, dMeta, dD, dRes, dCnt, dPat, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("phase 2: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> res(4096), cnt(4096);
    { float b[32]; cudaMemcpy(b, dPat, sizeof b, cudaMemcpyDeviceToHost);
      printf("baseline (nibbles 0x4), row 0 k 0..15:"); for (int j = 0; j < 16; ++j) printf(" %g", b[j]);
      printf("\n                        row 1 k 0..15:"); for (int j = 0; j < 16; ++j) printf(" %g", b[16 + j]); printf("\n"); }
    cudaMemcpy(res.data(), dRes, 4096 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(cnt.data(), dCnt, 4096 * 4, cudaMemcpyDeviceToHost);
    int agree = 0, none = 0, multi = 0;
    for (int e2 = 0; e2 < 4096; ++e2) {
      const int L = e2 >> 5, c = (e2 >> 3) & 3, q = e2 & 7;
      if (0 == res[e2]) { ++none; continue; }
      if (cnt[e2] > 1) ++multi;
      const int row = (res[e2] - 1) >> 5, g = (res[e2] - 1) & 31;
      const int hrow = (L % 8) + 16 * (L / 16) + 8 * (q / 4), hg = 8 * c + 4 * ((L / 8) % 2) + (q % 4);
      if (row == hrow && g == hg) ++agree;
    }
    printf("phase 2 (id2_from_col=%d): %d of 4096 nibbles agree with the hypothesis, %d steer nothing, %d steer several cells\n", id2c, agree, none, multi);
    char name[64]; snprintf(name, sizeof name, "gpurun_out/sp_map_id2c%d.txt", id2c);
    if (FILE* f = fopen(name, "w")) {
      for (int e2 = 0; e2 < 4096; ++e2) fprintf(f, "lane %3d col %d nib %d -> row %3d group %2d (cells %u)\n", e2 >> 5, (e2 >> 3) & 3, e2 & 7,
                                                  res[e2] ? (int)((res[e2] - 1) >> 5) : -1, res[e2] ? (int)((res[e2] - 1) & 31) : -1, cnt[e2]);
      fclose(f);
    }
    // nibble semantics
    if (res[0]) {
      P p2 = { base, 2, id2c, res[0], 0 };
      probe<<<1, 128, SM>>>(dA, dB, dMeta, dD, dRes, dCnt, dPat, p2);
      cudaDeviceSynchronize();
      float pat[64]; cudaMemcpy(pat, dPat, sizeof pat, cudaMemcpyDeviceToHost);
      for (int v = 0; v < 16; ++v) printf("  nibble 0x%X -> group reads %g %g %g %g (kept elements 1, 2)\n", v, pat[4 * v], pat[4 * v + 1], pat[4 * v + 2], pat[4 * v + 3]);
    }
  }
  return 0;
}
