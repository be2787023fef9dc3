// K4h: tensor-core branch of the spmdm compute step for bf16 inputs (fp32 accumulate), sm_100a.
//
//   C[128 rows, 256 cols] = beta*C + sum_kb densify(slice(kb, mb))[128 x 128] * B[kb*128 .. +128, 256 cols]
//
// The single-CTA predecessor of the CTA-pair kernel (spmdm_compute_tc16p.cu), kept as its fallback and for comparison
// (LIBXSMM_B200_TC16_PAIR=0).  bf16 x bf16 products are exact in the tensor core, so there is no operand split: one
// tcgen05.mma (kind::f16, M = 128, N = 256, K = 16) per 16 k; the time does not depend on the density.  Measured on
// B200, 4096^3: 164 us (K4p: 93 us).  With its stages switched off one at a time: MMA + TMA alone 128 us (B crosses
// L2 -> SM once per 128 output rows: 1.07 GB at 8.4 TB/s), zero-fill + its barrier 13 us, the epilogue 25 us.
//   * A: the slice block is scattered into a zeroed K-major SWIZZLE_128B tile, one 64-k half at a time
//     (128 rows x 128 B = 16 KiB), four buffers deep so that the workers run ahead of the tensor core (the
//     rebuild -> multiply -> release cycle of one buffer is latency bound).  A thread keeps its share of the k-block's nonzeros in registers (the slicing kernel's packed
//     word: bf16 value | half | position) and scatters them once per half.
//   * B: 64 x 256 tiles by TMA (four boxes of 64 columns, SWIZZLE_128B; MN-major operand: pinned with
//     tools/umma_probe/probe16.cu) through a 4-stage ring; for transb = 'T' (B stored n x k) one box of
//     256 rows x 64 k, a K-major operand.
//   * accumulator: 256 TMEM columns, carried over the whole K loop (products are exact and the 1e-2 contract
//     leaves four orders of magnitude for the tensor core's accumulation rounding).
//   * epilogue: tcgen05.ld (thread = row), + beta*C, 16-byte streaming stores; C stored n x m is written one
//     full line per instruction.
// Like the fp32 branch this kernel does not keep the reference's rounding sequence; LIBXSMM_B200_SPMDM_TC=0
// selects the order-preserving CUDA-core kernels.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace xb {

constexpr int T16_BM = 128;
constexpr int T16_BN = 256;
constexpr int T16_KH = 64;                        // k per step = one 128-byte swizzle row of bf16
constexpr int T16_NB = 4;                         // B stages
constexpr int T16_NA = 4;                         // A half-tile buffers (the rebuild -> multiply -> release cycle is latency bound)
constexpr int T16_WORKERS = 4;                    // worker warps = the four TMEM lane quarters
constexpr int T16_WT = T16_WORKERS * 32;
constexpr int T16_THREADS = (2 + T16_WORKERS) * 32;
constexpr int T16_NQ = 8;                         // nonzeros a worker thread keeps in registers per k-block
constexpr int T16_A_HALF = T16_BM * 128;          // 16 KiB
constexpr int T16_B_STAGE = T16_KH * T16_BN * 2;  // 32 KiB
constexpr int T16_SMEM_A = 0;
constexpr int T16_SMEM_B = T16_NA * T16_A_HALF;
constexpr int T16_SMEM_BAR = T16_SMEM_B + T16_NB * T16_B_STAGE;
constexpr int T16_SMEM_BYTES = T16_SMEM_BAR + 256;

__global__ void __launch_bounds__(T16_THREADS, 1)
spmdm_compute_tc16_kernel(const __grid_constant__ CUtensorMap tmB, const ComputeArgs p)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = (uint64_t*)(smem + T16_SMEM_BAR);
  uint64_t* b_full = bar;                 // [4] TMA landed
  uint64_t* b_free = bar + T16_NB;        // [4] MMAs that read the stage have completed
  uint64_t* a_ready = bar + 2 * T16_NB;   // [NA] workers built the A half
  uint64_t* a_free = a_ready + T16_NA;    // [NA] MMAs that read the A half have completed
  uint64_t* acc_full = a_free + T16_NA;   // all MMAs of the tile have completed
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const Geom& g = p.g;
  if (p.tc_twin > 0 && xb_total_nnz(p.sl.slice_nnz, g.mb * g.kb) < p.tc_min_nnz) return;   // very sparse: the CUDA-core twin does this multiply
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_mb = (g.bm + T16_BM - 1) / T16_BM;
  const int mbi = p.mb_first + (int)blockIdx.y / tiles_per_mb;
  const int ml0 = ((int)blockIdx.y % tiles_per_mb) * T16_BM;
  const int rows_in_block = min(g.bm, g.m - mbi * g.bm);
  if (ml0 >= rows_in_block) return;
  const int tile_rows = min(T16_BM, rows_in_block - ml0);
  const int n0 = (int)blockIdx.x * T16_BN;
  const int nsteps = g.kb * 2;
  const uint32_t sbase = smem_u32(smem);

  if (0 == tid) {
#pragma unroll
    for (int i = 0; i < T16_NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_free[i], 1); }
#pragma unroll
    for (int i = 0; i < T16_NA; ++i) { mbar_init(&a_ready[i], T16_WORKERS); mbar_init(&a_free[i], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (1 == warp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (0 == warp) {
    // ---------------- TMA producer ----------------
    if (0 == lane) {
      tma_prefetch_desc(&tmB);
      for (int t = 0; t < nsteps; ++t) {
        const int s = t % T16_NB, f = t / T16_NB;
        if (f > 0) mbar_wait(&b_free[s], (f - 1) & 1);
        mbar_arrive_expect_tx(&b_full[s], T16_B_STAGE);
        unsigned char* dst = smem + T16_SMEM_B + s * T16_B_STAGE;
        if (p.transb) tma_load_2d(dst, &tmB, t * T16_KH, n0, &b_full[s]);   // B stored n x k: 256 n-rows x 64 k
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(dst + j * (T16_KH * 128), &tmB, n0 + 64 * j, t * T16_KH, &b_full[s]);
        }
      }
    }
  }
  else if (1 == warp) {
    // ---------------- MMA issuer ----------------
    if (0 == lane) {
      // D = F32, A = B = BF16, A K-major, B MN-major ('N') or K-major ('T'), N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((p.transb ? 0u : 1u) << 16) | ((uint32_t)(T16_BN >> 3) << 17) | ((uint32_t)(T16_BM >> 4) << 24);
      const uint32_t b_kstep = p.transb ? 32u : 2048u, b_lbo = p.transb ? 16u : (uint32_t)(T16_KH * 128), b_sbo = 1024u;
      for (int t = 0; t < nsteps; ++t) {
        const int ab = t % T16_NA, s = t % T16_NB;
        mbar_wait(&a_ready[ab], (t / T16_NA) & 1);
        mbar_wait(&b_full[s], (t / T16_NB) & 1);
        tc_fence_after();
        const uint32_t a_base = sbase + T16_SMEM_A + ab * T16_A_HALF;
        const uint32_t b_base = sbase + T16_SMEM_B + s * T16_B_STAGE;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da = tc_smem_desc(a_base + ks * 32, 16, 1024, 2);
          const uint64_t db = tc_smem_desc(b_base + ks * b_kstep, b_lbo, b_sbo, 2);
          tc_mma_bf16(tmem_d, da, db, idesc, (t > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&b_free[s]);
        tc_commit(&a_free[ab]);
      }
      tc_commit(acc_full);
    }
  }
  else {
    // ---------------- workers: densify A, epilogue ----------------
    const int wt = tid - 64;                      // 0..127
    const size_t cap = (size_t)g.bm * g.bk;
    // packed nonzero (written by the slicing kernel, xb_tc16_pack): bf16 value << 16 | half << 15 | (byte offset inside the 16 KiB half) >> 1
    uint32_t pk[T16_NQ];
    // fetch() only ISSUES the loads of the next k-block's nonzeros (raw registers, no use): they are consumed a
    // step later, so the global-memory latency never stalls a worker
    uint32_t rw[T16_NQ];
    int first = 0, last = 0, nfirst = 0, nlast = 0;
    int pf = 0, pl = 0, pm = 0;        // raw row pointers of the k-block after next (first, end, last row's start)
    auto fetch_ptrs = [&](int kbf) {   // issue only
      const uint16_t* ro = p.sl.rowidx + (size_t)(kbf * g.mb + mbi) * (g.bm + 1) + ml0;
      pf = (int)__ldg(ro); pl = (int)__ldg(ro + tile_rows); pm = (int)__ldg(ro + tile_rows - 1);
    };
    auto fetch = [&](int kbf) {        // uses the pointers issued one k-block earlier; issues the nonzero loads
      const uint32_t* pw = p.sl.tcpk + (size_t)(kbf * g.mb + mbi) * cap;
      nfirst = pf;
      nlast = (pl < pf) ? pm : pl;     // wrapped u16 counter of a full slice: last row reads as empty
#pragma unroll
      for (int i = 0; i < T16_NQ; ++i) {
        const int q = nfirst + wt + i * T16_WT;
        rw[i] = 0;
        if (q < nlast) rw[i] = __ldg(pw + q);
      }
      if (kbf + 1 < g.kb) fetch_ptrs(kbf + 1);
    };
    fetch_ptrs(0);
    fetch(0);
    for (int t = 0; t < nsteps; ++t) {
      const int kb = t >> 1, h = t & 1, ab = t % T16_NA;
      if (t >= T16_NA) mbar_wait(&a_free[ab], ((t / T16_NA) - 1) & 1);
      unsigned char* abuf = smem + T16_SMEM_A + ab * T16_A_HALF;
      {
        uint4* z = (uint4*)abuf;
#pragma unroll
        for (int i = 0; i < T16_A_HALF / 16 / T16_WT; ++i) z[wt + i * T16_WT] = make_uint4(0, 0, 0, 0);
        asm volatile("bar.sync 1, %0;\n" ::"n"(T16_WT) : "memory");   // zero-fill complete before the scatter
      }
      if (0 == h) {   // the raw registers fetched a step ago become this k-block's packed nonzeros
        first = nfirst; last = nlast;
#pragma unroll
        for (int i = 0; i < T16_NQ; ++i) pk[i] = rw[i];
        if (kb + 1 < g.kb) fetch(kb + 1);   // the raw registers are free again: two steps of cover for the next k-block's loads
      }
      auto put = [&](uint32_t w) {
        if ((int)((w >> 15) & 1u) == h) *(uint16_t*)(abuf + ((w & 0x7FFFu) << 1)) = (uint16_t)(w >> 16);
      };
#pragma unroll
      for (int i = 0; i < T16_NQ; ++i) {
        if (first + wt + i * T16_WT < last) put(pk[i]);
      }
      if (first + T16_NQ * T16_WT < last) {   // denser than T16_NQ*T16_WT nonzeros per tile: the rest straight from memory
        const uint32_t* pw = p.sl.tcpk + (size_t)(kb * g.mb + mbi) * cap;
#pragma unroll 4
        for (int q = first + wt + T16_NQ * T16_WT; q < last; q += T16_WT) put(__ldg(pw + q));
      }
      fence_proxy_async();
      __syncwarp();
      if (0 == lane) mbar_arrive(&a_ready[ab]);
    }
    // epilogue: warp (2 + q') owns TMEM lanes 32*(warp % 4) .. +31
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const size_t crow = (size_t)(mbi * g.bm + ml0 + row - p.row_origin);
#pragma unroll 1
    for (int cb = 0; cb < T16_BN; cb += 32) {
      uint32_t v[32];
      tc_ld32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cb, v);
      if (row < tile_rows) {
        if (p.transc) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = n0 + cb + j;
            if (col < p.ncols) {
              float* dst = p.c + (size_t)col * p.ldc + crow;
              *dst = (0.f != p.beta) ? fmaf(p.beta, *dst, __uint_as_float(v[j])) : __uint_as_float(v[j]);
            }
          }
        }
        else {
          float* dst = p.c + crow * p.ldc + n0 + cb;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int col = n0 + cb + j;
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (col + 3 < p.ncols) {
              if (0.f != p.beta) {
                const float4 cin = *(const float4*)(dst + j);
                o.x = fmaf(p.beta, cin.x, o.x); o.y = fmaf(p.beta, cin.y, o.y); o.z = fmaf(p.beta, cin.z, o.z); o.w = fmaf(p.beta, cin.w, o.w);
              }
              st_global_cs_f4(dst + j, o);
            }
            else {
              const float e[4] = { o.x, o.y, o.z, o.w };
#pragma unroll
              for (int t2 = 0; t2 < 4; ++t2) if (col + t2 < p.ncols) dst[j + t2] = (0.f != p.beta) ? fmaf(p.beta, dst[j + t2], e[t2]) : e[t2];
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (1 == warp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(256) : "memory");
  }
}

bool make_tensor_map_2d_sw128(CUtensorMap* map, const void* base, int elem_bytes, unsigned long long cols, unsigned long long rows,
                              unsigned long long row_pitch_bytes, unsigned box_cols, unsigned box_rows, bool atom32);

// returns false when the panel does not qualify (caller falls back to the CUDA-core kernels)
bool launch_compute_tc16(const ComputeArgs& a, cudaStream_t stream)
{
  if (0 == a.aux_valid) return false;   // these kernels rebuild A from the packed per-nonzero words the slicing pass writes
  if (!a.is_bf16) return false;
  if (!a.transc && (0 != ((uintptr_t)a.c & 15) || 0 != (a.ldc & 3))) return false;
  CUtensorMap map;
  if (a.transb) {
    if (!make_tensor_map_2d_sw128(&map, a.b, 2, (unsigned long long)a.g.k, (unsigned long long)a.ncols, (unsigned long long)a.ldb * 2, 64, T16_BN, false)) return false;
  }
  else if (!make_tensor_map_2d_sw128(&map, a.b, 2, (unsigned long long)a.ncols, (unsigned long long)a.g.k, (unsigned long long)a.ldb * 2, 64, T16_KH, false)) return false;
  ensure_smem_optin((const void*)spmdm_compute_tc16_kernel, T16_SMEM_BYTES);
  const int tiles_per_mb = (a.g.bm + T16_BM - 1) / T16_BM;
  const dim3 grid((unsigned)((a.ncols + T16_BN - 1) / T16_BN), (unsigned)(a.mb_count * tiles_per_mb), 1);
  count_launch(1);
  note_compute_kernel("spmdm_compute_tc16_kernel");
  spmdm_compute_tc16_kernel<<<grid, T16_THREADS, T16_SMEM_BYTES, stream>>>(map, a);
  XB_CUDA(cudaGetLastError());
  return true;
}

}  // namespace xb
