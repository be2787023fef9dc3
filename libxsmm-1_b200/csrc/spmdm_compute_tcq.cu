// K4q: CTA-pair (tcgen05 cta_group::2) form of the fp32 tensor-core branch, for slices that carry their dense tile
// image (SliceArena::dense, written by the slicing kernel for fp32 / transa = 'N' / complete k-blocks).
//
//   C[256 rows, 256 cols] = beta*C + sum_kb A_image(kb, rows)[256 x 128] * B[kb*128 .. +128, 256 cols]      (3xTF32)
//
// The single-CTA kernel (K4, spmdm_compute_tc.cu) with the image needs 8 KiB of shared-memory operand reads per
// 128 x 128 x 8 MMA (64 clocks): with the TMA writes, the b_lo pass and the A copies it moves 176 KiB per 32-k chunk
// through a 128 B/clk port and is bound by exactly that (96 us for 2048^3).  A CTA pair computes 128 x 256 per CTA
// with each CTA staging its 128 rows of A and its 128 columns of B: the same bytes per CTA and chunk now feed twice
// the MMA work (12 x 128 clocks), which puts the tensor pipe, not the port, on the critical path.
//   * A: the image half (a_hi chunk 0, 1, a_lo chunk 0, 1 = 64 KiB, already in the SWIZZLE_128B operand layout) by two
//     TMA boxes of 256 x 128 B, completion counted on the leader's barrier (cta_group::2).
//   * B: 32 x 128 fp32 per CTA and chunk by TMA (four boxes of 32 columns, 32-byte-atom swizzle; transb = 'T': one
//     K-major box), b_lo = b - trunc_tf32(b) written by the worker warps of the CTA that holds the columns.
//   * MMA: the leader's elected thread, M = 256, N = 256, K = 8, three per k-step (hi*hi, hi*lo, lo*hi);
//     tcgen05.commit ... multicast releases stages / halves / accumulators in both CTAs.
//   * every k-block's accumulator (256 TMEM columns, two buffers) is drained into registers by the sixteen worker
//     warps of each CTA (thread = row, 64 columns) -- the tensor core truncates on every accumulation -- and the
//     epilogue adds beta*C and stores from those registers.
// Contract 1e-5 relative, not the reference's rounding sequence (like K4).
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace xb {

constexpr int Q_BM = 128;                       // rows per CTA
constexpr int Q_BN = 256;                       // columns per pair tile
constexpr int Q_BNH = 128;                      // columns of B staged per CTA
constexpr int Q_KC = 32;                        // k per B chunk = one 128-byte swizzle row of fp32
constexpr int Q_NB = 3;                         // B stages
constexpr int Q_WORKERS = 16;
constexpr int Q_WT = Q_WORKERS * 32;
constexpr int Q_THREADS = (2 + Q_WORKERS) * 32;
constexpr int Q_A_CHUNK = Q_BM * 128;           // 16 KiB
constexpr int Q_A_HALF = 4 * Q_A_CHUNK;         // 64 KiB: hi chunk 0, 1, lo chunk 0, 1
constexpr int Q_B_CHUNK = Q_KC * Q_BNH * 4;     // 16 KiB
constexpr int Q_SMEM_A = 0;
constexpr int Q_SMEM_B = 2 * Q_A_HALF;          // stage s: raw at +s*32K, lo at +s*32K+16K
constexpr int Q_SMEM_BAR = Q_SMEM_B + Q_NB * 2 * Q_B_CHUNK;
constexpr int Q_SMEM_BYTES = Q_SMEM_BAR + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Q_THREADS, 1)
spmdm_compute_tcq_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA, const ComputeArgs p)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = (uint64_t*)(smem + Q_SMEM_BAR);
  uint64_t* b_full = bar;            // [3] each CTA: its chunk landed
  uint64_t* b_split = bar + 3;       // [3] leader: both CTAs wrote b_lo
  uint64_t* b_free = bar + 6;        // [3] both: MMAs that read the stage have completed
  uint64_t* a_ready = bar + 9;       // [2] leader: both CTAs' A halves landed
  uint64_t* a_free = bar + 11;       // [2] both: MMAs that read the half have completed
  uint64_t* acc_full = bar + 13;     // [2] both: the k-block's MMAs have completed
  uint64_t* acc_free = bar + 15;     // [2] leader: the workers of both CTAs have drained the buffer
  uint32_t* tmem_slot = (uint32_t*)(bar + 17);

  const Geom& g = p.g;
  if (p.tc_twin > 0) pdl_wait();        // the twin decision reads the slices' counts
  if (p.tc_twin > 0 && xb_total_nnz(p.sl.slice_nnz, g.mb * g.kb) < p.tc_min_nnz) return;   // sparse: the CUDA-core twin multiplies (uniform over the grid)
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int tiles_per_mb = (g.bm + Q_BM - 1) / Q_BM;
  const int ctiles_m = p.mb_count * tiles_per_mb;
  const int pair_m = (ctiles_m + 1) >> 1;
  const int pidx = (int)blockIdx.x >> 1;                      // pair tile
  const int ct = 2 * (pidx % pair_m) + (int)rank;             // this CTA's 128-row tile
  const int n0 = (pidx / pair_m) * Q_BN;
  const int mbi = p.mb_first + ct / tiles_per_mb;
  const int ml0 = (ct % tiles_per_mb) * Q_BM;
  const int tile_rows = (ct < ctiles_m) ? max(0, min(Q_BM, min(g.bm, g.m - mbi * g.bm) - ml0)) : 0;
  const int nsteps = g.kb * 2;       // one step = one half k-block = two B chunks
  const uint32_t sbase = smem_u32(smem);

  if (0 == tid) {
#pragma unroll
    for (int i = 0; i < Q_NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_split[i], 2 * Q_WORKERS); mbar_init(&b_free[i], 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_ready[i], 1); mbar_init(&a_free[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], 2 * Q_WORKERS);
    }
    mbar_fence_init();
  }
  if (1 == warp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  if (p.tc_twin <= 0) pdl_wait();       // barriers, tensor memory and the cluster hand-shake were set up while the kernel in front was still finishing

  if (0 == warp) {
    // ---------------- producer: this CTA's A half per step, its B chunk per 32 k ----------------
    if (0 == lane) {
      tma_prefetch_desc(&tmB);
      tma_prefetch_desc(&tmA);
      // row of the image (128 bytes each) where this CTA's tile starts for k-block 0; a k-block further on the row
      // blocks of all mb lie in between (slice s = kb * mb_count + mb), a half is 512 rows
      const int img_tiles = (g.bm + 127) / 128;
      const bool has_rows = tile_rows > 0;
      for (int c = 0; c < nsteps * 2; ++c) {
        if (0 == (c & 1)) {
          const int t = c >> 1, kb = t >> 1, h = t & 1;
          if (kb > 0) mbar_wait(&a_free[h], (kb - 1) & 1);
          if (0 == rank) mbar_arrive_expect_tx(&a_ready[h], 2 * Q_A_HALF);
          const uint32_t lbar = map_to_cta(&a_ready[h], 0);
          // a CTA without rows (odd number of row tiles) still delivers its 64 KiB: rows of tile 0 (any valid address),
          // multiplied into accumulator rows nobody stores
          const long long tile_id = has_rows ? ((long long)(kb * g.mb + mbi) * img_tiles + (ml0 >> 7)) : 0;
          const int row0 = (int)(tile_id * 1024 + h * 512);
          unsigned char* dst = smem + Q_SMEM_A + h * Q_A_HALF;
          tma_load_2d_pair(dst, &tmA, 0, row0, lbar);
          tma_load_2d_pair(dst + 32768, &tmA, 0, row0 + 256, lbar);
        }
        const int s = c % Q_NB, f = c / Q_NB;
        if (f > 0) mbar_wait(&b_free[s], (f - 1) & 1);
        // B chunks complete on the barrier of the CTA that holds them: its own workers wait there, write b_lo and
        // report to the leader's b_split, which is all the MMA thread needs
        mbar_arrive_expect_tx(&b_full[s], Q_B_CHUNK);
        unsigned char* dst = smem + Q_SMEM_B + s * 2 * Q_B_CHUNK;
        const int nc = n0 + (int)rank * Q_BNH;
        if (p.transb) tma_load_2d(dst, &tmB, c * Q_KC, nc, &b_full[s]);   // B stored n x k: one box of 128 n-rows x 32 k
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(dst + j * (Q_KC * 128), &tmB, nc + 32 * j, c * Q_KC, &b_full[s]);
        }
      }
    }
  }
  else if (1 == warp) {
    // ---------------- MMA issuer (leader CTA) ----------------
    if (0 == rank && 0 == lane) {
      // D = F32, A = B = TF32, A K-major, B MN-major ('N') or K-major ('T'), N = 256, M = 256 (pair)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((p.transb ? 0u : 1u) << 16) | ((uint32_t)(Q_BN >> 3) << 17) | ((uint32_t)((2 * Q_BM) >> 4) << 24);
      const uint32_t b_kstep = p.transb ? 32u : 1024u, b_lbo = p.transb ? 16u : (uint32_t)(Q_KC * 128), b_sbo = p.transb ? 1024u : 512u, b_lay = p.transb ? 2u : 1u;
      for (int t = 0; t < nsteps; ++t) {
        const int kb = t >> 1, h = t & 1;
        const uint32_t tmem_acc = tmem_d + (uint32_t)((kb & 1) * Q_BN);
        if (0 == h && kb >= 2) mbar_wait(&acc_free[kb & 1], ((kb >> 1) - 1) & 1);
        mbar_wait(&a_ready[h], kb & 1);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = 2 * t + j, s = c % Q_NB;
          mbar_wait(&b_split[s], (c / Q_NB) & 1);
          tc_fence_after();
          const uint32_t a_hi = sbase + Q_SMEM_A + h * Q_A_HALF + j * Q_A_CHUNK;
          const uint32_t a_lo = a_hi + 2 * Q_A_CHUNK;
          const uint32_t b_hi = sbase + Q_SMEM_B + s * 2 * Q_B_CHUNK;
          const uint32_t b_lo = b_hi + Q_B_CHUNK;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dah = tc_smem_desc(a_hi + ks * 32, 16, 1024, 2);
            const uint64_t dal = tc_smem_desc(a_lo + ks * 32, 16, 1024, 2);
            const uint64_t dbh = tc_smem_desc(b_hi + ks * b_kstep, b_lbo, b_sbo, b_lay);
            const uint64_t dbl = tc_smem_desc(b_lo + ks * b_kstep, b_lbo, b_sbo, b_lay);
            tc_mma_tf32_pair(tmem_acc, dah, dbh, idesc, (h > 0 || j > 0 || ks > 0) ? 1u : 0u);
            tc_mma_tf32_pair(tmem_acc, dah, dbl, idesc, 1u);
            tc_mma_tf32_pair(tmem_acc, dal, dbh, idesc, 1u);
          }
          tc_commit_pair(&b_free[s]);
        }
        tc_commit_pair(&a_free[h]);
        if (1 == h) tc_commit_pair(&acc_full[kb & 1]);
      }
    }
  }
  else {
    // ---------------- workers: b_lo of this CTA's columns, drain, epilogue ----------------
    const int w = warp - 2;                       // 0..15
    const int wt = tid - 64;                      // 0..511
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int cgrp = (w >> 2) * 64;               // this thread's 64 of the 256 columns
    const uint32_t lead_split0 = map_to_cta(&b_split[0], 0), lead_accfree0 = map_to_cta(&acc_free[0], 0);
    float run[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) run[j] = 0.f;
    auto drain = [&](int kbd) {                   // run += accumulator of k-block kbd
      mbar_wait(&acc_full[kbd & 1], (kbd >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 4; ++q) {      // 16 columns at a time: 64 running sums per thread leave little room
        uint32_t v[16];
        tc_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((kbd & 1) * Q_BN + cgrp + 16 * q), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) run[16 * q + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (0 == lane) mbar_arrive_cluster(lead_accfree0 + (kbd & 1) * 8);
    };
    for (int t = 0; t < nsteps; ++t) {
      const int kb = t >> 1, h = t & 1;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 2 * t + j, s = c % Q_NB;
        mbar_wait(&b_full[s], (c / Q_NB) & 1);
        const uint4* src = (const uint4*)(smem + Q_SMEM_B + s * 2 * Q_B_CHUNK);
        uint4* dst = (uint4*)(smem + Q_SMEM_B + s * 2 * Q_B_CHUNK + Q_B_CHUNK);
#pragma unroll
        for (int i = 0; i < Q_B_CHUNK / 16 / Q_WT; ++i) {
          const uint4 b = src[wt + i * Q_WT];
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(b.x) - __uint_as_float(b.x & 0xFFFFE000u));
          l.y = __float_as_uint(__uint_as_float(b.y) - __uint_as_float(b.y & 0xFFFFE000u));
          l.z = __float_as_uint(__uint_as_float(b.z) - __uint_as_float(b.z & 0xFFFFE000u));
          l.w = __float_as_uint(__uint_as_float(b.w) - __uint_as_float(b.w & 0xFFFFE000u));
          dst[wt + i * Q_WT] = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (0 == lane) mbar_arrive_cluster(lead_split0 + s * 8);
      }
      if (0 == h && kb > 0) drain(kb - 1);
    }
    drain(g.kb - 1);
    // epilogue from registers
    const size_t crow = (size_t)(mbi * g.bm + ml0 + row - p.row_origin);
    if (row < tile_rows) {
      if (p.transc) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const int col = n0 + cgrp + j;
          if (col < p.ncols) {
            float* dst = p.c + (size_t)col * p.ldc + crow;
            *dst = (0.f != p.beta) ? fmaf(p.beta, *dst, run[j]) : run[j];
          }
        }
      }
      else {
        float* dst = p.c + crow * p.ldc + n0 + cgrp;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const int col = n0 + cgrp + j;
          float4 o = make_float4(run[j], run[j + 1], run[j + 2], run[j + 3]);
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
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's shared memory and barriers stay alive until every MMA and remote arrive has landed
  if (1 == warp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(512) : "memory");
  }
}

bool make_tensor_map_2d(CUtensorMap* map, const void* base, int elem_bytes, unsigned long long cols, unsigned long long rows,
                        unsigned long long row_pitch_bytes, unsigned box_cols, unsigned box_rows);
bool make_tensor_map_2d_sw128(CUtensorMap* map, const void* base, int elem_bytes, unsigned long long cols, unsigned long long rows,
                              unsigned long long row_pitch_bytes, unsigned box_cols, unsigned box_rows, bool atom32);

// returns false when the panel does not qualify (caller falls back to the single-CTA kernel)
bool launch_compute_tcq(const ComputeArgs& a, cudaStream_t stream)
{
  if (a.is_bf16 || 0 == a.dense_valid || 0 == a.sl.dense) return false;
  if (!a.transc && (0 != ((uintptr_t)a.c & 15) || 0 != (a.ldc & 3))) return false;
  CUtensorMap mapB, mapA;
  if (a.transb) {
    if (!make_tensor_map_2d_sw128(&mapB, a.b, 4, (unsigned long long)a.g.k, (unsigned long long)a.ncols, (unsigned long long)a.ldb * 4, 32, Q_BNH, false)) return false;
  }
  else if (!make_tensor_map_2d_sw128(&mapB, a.b, 4, (unsigned long long)a.ncols, (unsigned long long)a.g.k, (unsigned long long)a.ldb * 4, 32, Q_KC, true)) return false;
  // the image as a plain matrix of 128-byte rows (the bytes already are in the swizzled operand layout)
  const unsigned long long img_rows = (unsigned long long)a.g.mb * a.g.kb * ((a.g.bm + 127) / 128) * 1024ull;
  if (img_rows > 0x7FFFFFFFull) return false;
  if (!make_tensor_map_2d(&mapA, a.sl.dense, 4, 32, img_rows, 128, 32, 256)) return false;
  ensure_smem_optin((const void*)spmdm_compute_tcq_kernel, Q_SMEM_BYTES);
  const int tiles_per_mb = (a.g.bm + Q_BM - 1) / Q_BM;
  const int pair_m = (a.mb_count * tiles_per_mb + 1) / 2;
  const int total = pair_m * ((a.ncols + Q_BN - 1) / Q_BN);
  if (total <= 0) return true;
  count_launch(1);
  note_compute_kernel("spmdm_compute_tcq_kernel");
  XB_CUDA(launch_pdl(spmdm_compute_tcq_kernel, dim3(2u * (unsigned)total), dim3(Q_THREADS), Q_SMEM_BYTES, stream, mapB, mapA, a));   // may be scheduled while the slicing kernel drains (pdl_wait in the kernel)
  return true;
}

}  // namespace xb
