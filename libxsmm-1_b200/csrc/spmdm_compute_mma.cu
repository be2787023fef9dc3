// K2m: sliced SpMM for bf16 inputs in the SPARSE regime on the warp-level tensor-core path (sm_100a).
//
// A CUDA-core formulation pays, per nonzero and 256 columns, 8 FFMA plus 8 bf16->fp32 unpack operations and
// is instruction-issue bound (profiles/).  Here a warp owns 16 consecutive rows (the M dimension of
// mma.sync.m16n8k16) and gathers its nonzeros 16 at a time into the K dimension:
//     A[r][j] = value of nonzero j if it belongs to row r, else 0         (16 x 16, built in registers)
//     B[j][:] = the row of the staged B tile that nonzero j points at      (gathered by ldmatrix row addresses)
// so that one instruction applies 16 nonzeros to 8 columns; the bf16 operands go to the tensor core as they
// are (no unpack) and products are exact.  tcgen05 cannot do this: its operands are regular shared-memory
// tiles, and at 1 % density a 128-row tile touches nearly every k, so nothing could be skipped.
// B tiles (128 x 256 bf16) are staged by TMA into an mbarrier ring exactly like in spmdm_compute_tma.cu;
// row pointers are prefetched two k-blocks ahead, nonzeros (column, value, row) one ahead.
// Rounding: fp32 accumulation inside the tensor core, NOT the reference's fma chain; the bar is the 1e-2
// relative contract for bf16 inputs (observed ~1e-7).  Zero entries of A multiply rows of B that the
// reference would not read for that output row: identical for finite B.
// MEASURED (B200, C2 = bf16 4096^3 at 1 %): 217 us (no bank-class packing) / 239 us (packing) against 177 us of
// the CUDA-core kernel -- HMMA.16816 issues at ~2.5 clk per SM at best on this part and the gather leaves the
// pipe 30-40 % busy, so this kernel is OPT-IN (LIBXSMM_B200_SPMDM_MMA=1) and kept as a documented experiment.
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace xb {

#ifndef MM_PACK
#define MM_PACK 0
#endif
#ifndef MM_CW_
#define MM_CW_ 16
#define MM_BN_ 128
#define MM_STAGES_ 4
#endif
constexpr int MM_CW = MM_CW_;             // warps per CTA
constexpr int MM_R = 16;                  // rows per warp
constexpr int MM_TM = MM_CW * MM_R;       // rows per CTA
constexpr int MM_BN = MM_BN_;             // columns per CTA
constexpr int MM_NT = MM_BN / 8;          // n-tiles of 8 columns
constexpr int MM_NCB = MM_BN / 64;        // column blocks per stage
constexpr int MM_CBLK = 128 * 128;        // one column block of a stage: 128 k-rows x 128 B (64 columns), SWIZZLE_128B
constexpr int MM_STAGEB = MM_NCB * MM_CBLK;
constexpr int MM_STAGES = MM_STAGES_;
constexpr int MM_ZERO = MM_STAGES * MM_STAGEB;            // 8 zero rows of 128 B for unused k slots (one per bank class)
constexpr int MM_SLOTS = MM_PACK ? 256 : 32;              // slots per warp (packing: 8 classes x up to 32 ranks)
constexpr int MM_SCR = MM_ZERO + 1024;                    // per-warp slot table: MM_SLOTS x {row address, value|row}
constexpr int MM_BAR = MM_SCR + MM_CW * MM_SLOTS * 8;
constexpr int MM_SMEM = MM_BAR + 2 * MM_STAGES * 8;

__device__ __forceinline__ void mm_ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(MM_CW * 32, 1)
spmdm_compute_mma_kernel(const __grid_constant__ CUtensorMap tmB, const ComputeArgs p)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = (uint64_t*)(smem + MM_BAR);
  uint64_t* empty = full + MM_STAGES;

  const Geom& g = p.g;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int tiles_per_mb = (g.bm + MM_TM - 1) / MM_TM;
  const int mbi = p.mb_first + (int)blockIdx.y / tiles_per_mb;
  const int ml0 = ((int)blockIdx.y % tiles_per_mb) * MM_TM;
  const int rows_in_block = min(g.bm, g.m - mbi * g.bm);
  if (ml0 >= rows_in_block) return;
  const int n0 = (int)blockIdx.x * MM_BN;
  const uint32_t sbase = smem_u32(smem);

  if (0 == tid) {
#pragma unroll
    for (int s = 0; s < MM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], MM_CW); }
    mbar_fence_init();
  }
  for (int i = tid; i < 1024 / 4; i += MM_CW * 32) ((uint32_t*)(smem + MM_ZERO))[i] = 0;
  __syncthreads();

  const bool is_producer = (0 == tid);
  auto produce = [&](int t) {
    if (t < g.kb) {
      const int s = t % MM_STAGES, f = t / MM_STAGES;
      if (f > 0) mbar_wait(&empty[s], (f - 1) & 1);
      mbar_arrive_expect_tx(&full[s], MM_STAGEB);
#pragma unroll
      for (int cb = 0; cb < MM_NCB; ++cb) tma_load_2d(smem + (size_t)s * MM_STAGEB + cb * MM_CBLK, &tmB, n0 + 64 * cb, t * g.bk, &full[s]);
    }
  };
  if (is_producer) {
    tma_prefetch_desc(&tmB);
    for (int t = 0; t < MM_STAGES - 1; ++t) produce(t);
  }

  const int wrow0 = ml0 + warp * MM_R;
  const int nvalid = max(0, min(MM_R, rows_in_block - wrow0));
  const size_t cap = (size_t)g.bm * g.bk;
  float acc[MM_NT][4];
#pragma unroll
  for (int t = 0; t < MM_NT; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }

  const size_t rp_stride = (size_t)g.mb * (g.bm + 1);
  const size_t nz_stride = (size_t)g.mb * cap;
  const bool rp_lane = (lane <= MM_R) && (nvalid > 0);
  const uint16_t* rpp = p.sl.rowidx + (size_t)mbi * (g.bm + 1) + wrow0 + min(lane, nvalid);
  const uint16_t* cpn = p.sl.colidx + (size_t)mbi * cap + lane;
  const uint16_t* rin = p.sl.tcoff + (size_t)mbi * cap + lane;     // bf16 slices: block-local row of every nonzero
  const float* vpn = p.sl.values + (size_t)mbi * cap + lane;
  uint2* scr = (uint2*)(smem + MM_SCR) + warp * MM_SLOTS;

  auto monotone = [&](int rp) -> int {   // wrapped u16 end pointer of a completely full slice reads as "empty row"
    const int nxt = __shfl_down_sync(0xffffffffu, rp, 1);
    if (__any_sync(0xffffffffu, lane < MM_R && nxt < rp)) {
#pragma unroll
      for (int d = 1; d <= MM_R; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, rp, d);
        if (lane >= d) rp = max(rp, t);
      }
    }
    return rp;
  };

  // lane q of a chunk holds nonzero (first + q): column, block-local row, value; has = valid
  int rp_cur = rp_lane ? (int)__ldg(rpp) : 0;
  rpp += rp_stride;
  int rp_nxt = (rp_lane && 1 < g.kb) ? (int)__ldg(rpp) : 0;
  rpp += rp_stride;
  rp_cur = monotone(rp_cur);
  uint32_t col_c = 0, row_c = 0; float val_c = 0.f; bool has_c = false;
  {
    const int s0 = __shfl_sync(0xffffffffu, rp_cur, 0), e0 = __shfl_sync(0xffffffffu, rp_cur, MM_R);
    if (s0 + lane < e0) { col_c = __ldg(cpn + s0); row_c = __ldg(rin + s0); val_c = __ldg(vpn + s0); has_c = true; }
  }
  cpn += nz_stride; rin += nz_stride; vpn += nz_stride;

  for (int kb = 0; kb < g.kb; ++kb) {
    if (is_producer) produce(kb + MM_STAGES - 1);
    const int rp_nn = (rp_lane && kb + 2 < g.kb) ? (int)__ldg(rpp) : 0;
    rpp += rp_stride;
    rp_nxt = monotone(rp_nxt);
    uint32_t col_n = 0, row_n = 0; float val_n = 0.f; bool has_n = false;
    {
      const int s1 = __shfl_sync(0xffffffffu, rp_nxt, 0), e1 = __shfl_sync(0xffffffffu, rp_nxt, MM_R);
      if (s1 + lane < e1) { col_n = __ldg(cpn + s1); row_n = __ldg(rin + s1); val_n = __ldg(vpn + s1); has_n = true; }
    }
    const int first = __shfl_sync(0xffffffffu, rp_cur, 0);
    const int total = __shfl_sync(0xffffffffu, rp_cur, MM_R) - first;
    const int s = kb % MM_STAGES;
    mbar_wait(&full[s], (kb / MM_STAGES) & 1);
    const uint32_t stage = sbase + (uint32_t)s * MM_STAGEB;

    for (int p0 = 0; p0 < total; p0 += 32) {
      if (0 != p0) {   // denser rows: later chunks are fetched in line
        has_c = (p0 + lane < total);
        col_c = 0; row_c = 0; val_c = 0.f;
        if (has_c) {
          col_c = __ldg(cpn - nz_stride + first + p0);
          row_c = __ldg(rin - nz_stride + first + p0);
          val_c = __ldg(vpn - nz_stride + first + p0);
        }
      }
      // ---- slot assignment: the 8 B-rows that one ldmatrix phase gathers must fall into different banks.
      // In the SWIZZLE_128B tile the bank class of row k is k & 7, so slot = class + 8 * (rank inside the class);
      // the order of nonzeros inside an instruction is free (the tensor core sums over its K dimension).
#if MM_PACK
      const uint32_t cls = col_c & 7u;
      // lanes of the same class: three ballots on the class bits (MATCH.ANY has a much longer latency)
      const uint32_t bv = __ballot_sync(0xffffffffu, has_c);
      const uint32_t b0 = __ballot_sync(0xffffffffu, cls & 1u), b1 = __ballot_sync(0xffffffffu, cls & 2u), b2 = __ballot_sync(0xffffffffu, cls & 4u);
      const uint32_t peers = bv & ((cls & 1u) ? b0 : ~b0) & ((cls & 2u) ? b1 : ~b1) & ((cls & 4u) ? b2 : ~b2);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      const uint32_t slot = cls + 8u * rank;
      const int nslots = 8 * (int)__reduce_max_sync(0xffffffffu, has_c ? rank + 1u : 0u);
#else
      const uint32_t slot = (uint32_t)lane;
      const int nslots = min(32, total - p0);
#endif
      const int ngroups = (nslots + 15) >> 4;
      // slot table: {address of the B row (column block 0), value bits << 16 | row in warp}; empty slots -> zero row of the class
      for (int i = lane; i < ngroups * 16; i += 32) scr[i] = make_uint2(sbase + MM_ZERO + (uint32_t)(i & 7) * 128u, 0u);
      __syncwarp();
      if (has_c) scr[slot] = make_uint2(stage + col_c * 128u, (__float_as_uint(val_c) & 0xFFFF0000u) | ((row_c - (uint32_t)wrow0) & 15u));
      __syncwarp();
      for (int gi = 0; gi < ngroups; ++gi) {
        // ---- A fragment: slots 2*tig, 2*tig+1 (k 0..7 half) and 2*tig+8, 2*tig+9 for rows gid and gid+8 ----
        uint32_t a[4];
        {
          const uint4 u = *(const uint4*)(smem + MM_SCR + ((size_t)warp * MM_SLOTS + gi * 16 + 2 * tig) * 8);       // slots 2tig, 2tig+1
          const uint4 w = *(const uint4*)(smem + MM_SCR + ((size_t)warp * MM_SLOTS + gi * 16 + 2 * tig + 8) * 8);   // slots 2tig+8, 2tig+9
          const uint32_t v0 = u.y, v1 = u.w, v2 = w.y, v3 = w.w;
          const uint32_t r0 = v0 & 15u, r1 = v1 & 15u, r2 = v2 & 15u, r3 = v3 & 15u;
          const uint32_t lo = (uint32_t)gid, hi = (uint32_t)gid + 8u;
          a[0] = ((r0 == lo) ? (v0 >> 16) : 0u) | ((r1 == lo) ? (v1 & 0xFFFF0000u) : 0u);
          a[1] = ((r0 == hi) ? (v0 >> 16) : 0u) | ((r1 == hi) ? (v1 & 0xFFFF0000u) : 0u);
          a[2] = ((r2 == lo) ? (v2 >> 16) : 0u) | ((r3 == lo) ? (v3 & 0xFFFF0000u) : 0u);
          a[3] = ((r2 == hi) ? (v2 >> 16) : 0u) | ((r3 == hi) ? (v3 & 0xFFFF0000u) : 0u);
        }
        // ---- B rows: lane l addresses row (l & 7) of matrix (l >> 3): slot ((l >> 3) & 1)*8 + (l & 7); matrices 2, 3
        // are the next n-tile.  16-byte unit u of a row sits at unit (u ^ class): R = row | ((hi ^ class) << 4), and
        // the unit of n-tile t (even) is then R ^ ((t & 7) << 4).
        const int myslot = gi * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const uint32_t rowaddr = scr[myslot].x;
        const uint32_t R = rowaddr | ((((uint32_t)(lane >> 4)) ^ ((rowaddr >> 7) & 7u)) << 4);   // bank class of a row = row index & 7
        const uint32_t cbs = (rowaddr >= sbase + MM_ZERO) ? 0u : (uint32_t)MM_CBLK;   // the zero rows have no column blocks
#pragma unroll
        for (int t = 0; t < MM_NT; t += 2) {
          uint32_t b0, b1, b2, b3;
          mm_ldsm4t((R + (uint32_t)(t >> 3) * cbs) ^ (uint32_t)((t & 7) << 4), b0, b1, b2, b3);
          mm_mma(acc[t], a, b0, b1);
          mm_mma(acc[t + 1], a, b2, b3);
        }
      }
      __syncwarp();
    }
    __syncwarp();
    if (0 == lane) mbar_arrive(&empty[s]);
    cpn += nz_stride; rin += nz_stride; vpn += nz_stride;
    rp_cur = rp_nxt; rp_nxt = rp_nn;
    col_c = col_n; row_c = row_n; val_c = val_n; has_c = has_n;
  }

  // ---- epilogue: C fragment (rows gid, gid+8; columns 2*tig, 2*tig+1 of every n-tile) + beta*C -------------
  const bool vec2 = (0 == (p.ldc & 1)) && (0 == ((uintptr_t)p.c & 7));
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = gid + half * 8;
    if (r < nvalid) {
      float* crow = p.c + (size_t)(mbi * g.bm + wrow0 + r - p.row_origin) * p.ldc;
#pragma unroll
      for (int t = 0; t < MM_NT; ++t) {
        const int col = n0 + t * 8 + 2 * tig;
        float x = acc[t][2 * half], y = acc[t][2 * half + 1];
        if (col + 1 < p.ncols && vec2) {
          if (0.f != p.beta) { const float2 cin = *(const float2*)(crow + col); x = fmaf(p.beta, cin.x, x); y = fmaf(p.beta, cin.y, y); }
          *(float2*)(crow + col) = make_float2(x, y);
        }
        else {
          if (col < p.ncols) crow[col] = (0.f != p.beta) ? fmaf(p.beta, crow[col], x) : x;
          if (col + 1 < p.ncols) crow[col + 1] = (0.f != p.beta) ? fmaf(p.beta, crow[col + 1], y) : y;
        }
      }
    }
  }
}

bool make_tensor_map_2d_sw128(CUtensorMap* map, const void* base, int elem_bytes, unsigned long long cols, unsigned long long rows,
                              unsigned long long row_pitch_bytes, unsigned box_cols, unsigned box_rows, bool atom32);

// returns false when the panel does not qualify
bool launch_compute_mma(const ComputeArgs& a, cudaStream_t stream)
{
  if (!a.is_bf16 || a.transb || a.transc) return false;
  CUtensorMap map;
  if (!make_tensor_map_2d_sw128(&map, a.b, 2, (unsigned long long)a.ncols, (unsigned long long)a.g.k, (unsigned long long)a.ldb * 2, 64, 128, false)) return false;
  static bool configured = false;
  if (!configured) {
    XB_CUDA(cudaFuncSetAttribute(spmdm_compute_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM));
    configured = true;
  }
  const int tiles_per_mb = (a.g.bm + MM_TM - 1) / MM_TM;
  const dim3 grid((unsigned)((a.ncols + MM_BN - 1) / MM_BN), (unsigned)(a.mb_count * tiles_per_mb), 1);
  count_launch(1);
  spmdm_compute_mma_kernel<<<grid, MM_CW * 32, MM_SMEM, stream>>>(map, a);
  XB_CUDA(cudaGetLastError());
  return true;
}

}  // namespace xb
