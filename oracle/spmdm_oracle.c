/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's SPMDM path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The product (libxsmm-1_b200/csrc) never links or calls it.
 *
 * Pinned: tests/test_oracle_vs_ref.py checks every function here bit-for-bit against
 * the UNMODIFIED reference compiled by oracle/build_ref.sh (AVX2 instantiation, bn = 48),
 * and tests/golden/ holds outputs of that reference for use where it is not present.
 *
 * What is restated (all paths relative to /root/reference):
 *   geometry  src/libxsmm_spmdm.c:540-608
 *   slicing   src/template/libxsmm_spmdm_createSparseSlice_fp32_thread.tpl.c:47-141
 *             src/template/libxsmm_spmdm_createSparseSlice_bfloat16_thread.tpl.c:47-142
 *   compute   src/template/libxsmm_spmdm_compute_fp32_thread.tpl.c:38-558
 *             src/template/libxsmm_spmdm_compute_bfloat16_thread.tpl.c:38-585
 * It is a plain scalar program: one output element at a time, but with the SAME
 * rounding sequence per element as the vectorised reference (see orc_spmdm_compute).
 *
 * Flat slice layout used by the oracle, the reference dump (oracle/ref_driver.c) and the
 * parity tests: slice s = kb*mb_count + mb (reference src/libxsmm_spmdm.c:126-127) owns
 *   rowidx[s*(bm+1) ...], colidx[s*bm*bk ...], values[s*bm*bk ...].
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

enum { G_M, G_N, G_K, G_BM, G_BN, G_BK, G_MB, G_NB, G_KB };

static float bf16_widen(uint16_t h) { union { uint32_t u; float f; } x; x.u = (uint32_t)h << 16; return x.f; }
static int is_t(char c) { return 'T' == c || 't' == c; }

/*
 * Block geometry (reference src/libxsmm_spmdm.c:552-608).  bn is what the reference binds
 * to its ISA instantiation (96 AVX-512, 48 AVX2, 6 scalar; :557-583) and is an input here.
 * The row-block height bm starts at 512 or 256 and is shrunk one row at a time while the
 * product of "largest block / average block" and "busiest thread / average thread" exceeds
 * 1.1 (:589-608).  The doubles are evaluated in the reference's association order.
 * Returns the per-thread scratch size in bytes (:148-155).
 */
int orc_spmdm_geometry(int M, int N, int K, int max_threads, int bn, int* g)
{
  int bm = (M >= 4096 || M <= 1024) ? 512 : 256;
  const int bk = 128;
  int mb = (M + bm - 1) / bm;
  const int nb = (N + bn - 1) / bn;
  const int kb = (K + bk - 1) / bk;
  for (;;) {
    const int biggest = bm * bn;
    const double mean_block = (double)((size_t)M * N) / ((size_t)mb * nb);
    const double skew_block = biggest / mean_block;
    const int busiest = (mb * nb + max_threads - 1) / max_threads;
    const double mean_thread = (double)mb * nb / max_threads;
    const double skew_thread = busiest / mean_thread;
    if (!(32 < bm && skew_block * skew_thread > 1.1)) break;
    --bm;
    mb = (M + bm - 1) / bm;
  }
  g[G_M] = M; g[G_N] = N; g[G_K] = K; g[G_BM] = bm; g[G_BN] = bn; g[G_BK] = bk;
  g[G_MB] = mb; g[G_NB] = nb; g[G_KB] = kb;
  {
    size_t sz = ((size_t)bm * bn + (size_t)bk * bn) * sizeof(float);
    sz = (sz + 4095) & ~(size_t)4095;
    return (int)sz;
  }
}

/*
 * Dense A -> CSR slices, all mb*kb blocks (reference createSparseSlice templates).
 *   dtype 0: a is float;  dtype 1: a is bf16 bits, widened by <<16 before the test.
 *   transa 'N': element (i,k) at a[i*K+k];  'T': at a[k*M+i]   (fp32 tpl.c:49-59).
 *   simd_w: vector width of the instantiation mirrored (8 AVX2, 16 AVX-512, 1 scalar).
 *     Columns handled by the vector loops use an ORDERED not-equal compare
 *     (_CMP_NEQ_OQ, src/libxsmm_spmdm_begin_avx2.h:54) which drops NaN; the scalar
 *     remainder loop uses "!(0 == v)" which keeps NaN.  The vector region ends at
 *     ncols/W*W for fp32 (tpl.c:63-64,106-128) and at ncols/(4W)*(4W) for bf16
 *     (bf16 tpl.c:61,119-130).  -0.0 is dropped everywhere, denormals and Inf are kept.
 *   The running count is a uint16_t and is used as the store position (tpl.c:72), so it
 *   wraps at 65536 exactly like the reference.
 * Entries at and beyond the slice's count are left untouched (callers pre-zero).
 */
void orc_spmdm_slices(const int* g, int dtype, char transa, const void* a, int simd_w,
                      uint16_t* rowidx, uint16_t* colidx, float* values)
{
  const int M = g[G_M], K = g[G_K], bm = g[G_BM], bk = g[G_BK], mbc = g[G_MB], kbc = g[G_KB];
  const size_t cap = (size_t)bm * bk;
  int kb, mb, i, k;
  if (simd_w < 1) simd_w = 1;
  for (kb = 0; kb < kbc; ++kb) for (mb = 0; mb < mbc; ++mb) {
    const int s = kb * mbc + mb;
    const int nrows = ((mb + 1) * bm > M) ? (M - mb * bm) : bm;
    const int ncols = ((kb + 1) * bk > K) ? (K - kb * bk) : bk;
    int vec_end = (0 == dtype) ? (ncols / simd_w * simd_w) : (ncols / (4 * simd_w) * (4 * simd_w));
    uint16_t* ro = rowidx + (size_t)s * (bm + 1);
    uint16_t* co = colidx + s * cap;
    float* va = values + s * cap;
    uint16_t cnt = 0;
    if (1 == simd_w) vec_end = 0;
    for (i = 0; i < nrows; ++i) {
      ro[i] = cnt;
      for (k = 0; k < ncols; ++k) {
        const size_t at = is_t(transa) ? ((size_t)(kb * bk + k) * M + (size_t)mb * bm + i)
                                       : ((size_t)(mb * bm + i) * K + (size_t)kb * bk + k);
        const float v = (0 == dtype) ? ((const float*)a)[at] : bf16_widen(((const uint16_t*)a)[at]);
        const int keep = (k < vec_end) ? (v < 0.f || v > 0.f) : !(v == 0.f);
        if (keep) { co[cnt] = (uint16_t)k; va[cnt] = v; ++cnt; }
      }
    }
    ro[nrows] = cnt;
  }
}

/*
 * C = beta*C + A_slices * B for the whole problem, evaluated block by block the way the
 * reference's compute template does, one output element at a time.
 *
 * Rounding sequence per output element (m, n) -- this is what "same result" means:
 *   start   : 0 if beta == 0, C if beta == 1, else the single product beta*C
 *             (fp32 tpl.c:81-212).  For bf16 the reference reads *beta as an unsigned
 *             16-bit INTEGER and converts that to float (bf16 tpl.c:91,113,164), so the
 *             caller passes beta already converted the same way.
 *   CHAIN   : column inside a full-width block (num_n == bn): one fused multiply-add per
 *             nonzero, slices in ascending kb, nonzeros in stored order (tpl.c:309-370).
 *   PARTIAL : column inside the last, narrower block and below last_n_start: per kb a
 *             fresh sum from zero by fused multiply-adds, then ONE add into the running
 *             value (tpl.c:372-434).
 *   TAIL    : remaining columns of the narrow block: "run += b*v" per nonzero
 *             (tpl.c:395-398), which GCC at -O2 contracts into an fma inside the
 *             fma-enabled function (tail_fma = 1); tail_fma = 0 keeps product and sum
 *             separately rounded.
 * transb 'T': B stored N x K (tpl.c:228-252); transc 'T': C stored N x M (:103-110,510-532).
 * The bf16 variant widens B on staging (bf16 tpl.c:246-279); C stays fp32.
 * Row counts are taken as int differences of the u16 row pointers, so a wrapped slice
 * (end < start) contributes nothing, as in the reference (tpl.c:292-297).
 * Quirk Q17 (AVX-512 instantiation only): staging beta*C for transc == 'T' and beta not in
 * {0, 1} transposes SIMD_WIDTH x SIMD_WIDTH blocks and then scales EIGHT rows of each block,
 * whatever the vector width (fp32 tpl.c:163-172, bf16 tpl.c:173-182: n_block_size*0 .. *7).
 * With 16-wide vectors rows 8..15 of every complete 16 x 16 block therefore start from C
 * instead of beta*C.  Mirrored here for simd_w == 16 unless flags bit 1 is set
 * (flags = tail_fma | 2: the scaling every other path of the reference performs).
 */
void orc_spmdm_compute(const int* g, int dtype, char transb, char transc, float beta,
                       const uint16_t* rowidx, const uint16_t* colidx, const float* values,
                       const void* b, float* c, int simd_w, int flags)
{
  const int tail_fma = flags & 1, fix_q17 = flags & 2;
  const int M = g[G_M], N = g[G_N], K = g[G_K], bm = g[G_BM], bn = g[G_BN], bk = g[G_BK];
  const int mbc = g[G_MB], nbc = g[G_NB], kbc = g[G_KB];
  const size_t cap = (size_t)bm * bk;
  int mb, nb, ml, nl, kb, j;
  if (simd_w < 1) simd_w = 1;
  const int fused = (simd_w > 1);
  for (mb = 0; mb < mbc; ++mb) for (nb = 0; nb < nbc; ++nb) {
    const int m0 = mb * bm, n0 = nb * bn;
    const int num_m = ((m0 + bm) > M ? M : (m0 + bm)) - m0;
    const int num_n = ((n0 + bn) > N ? N : (n0 + bn)) - n0;
    const int narrow = (num_n != bn);
    int full_regs = num_n / simd_w, tail_from;
    if (full_regs > 0 && (full_regs % 2)) --full_regs;
    tail_from = full_regs * simd_w;
    for (ml = 0; ml < num_m; ++ml) for (nl = 0; nl < num_n; ++nl) {
      const size_t cat = is_t(transc) ? ((size_t)(n0 + nl) * M + m0 + ml) : ((size_t)(m0 + ml) * N + n0 + nl);
      const int mode = !narrow ? 0 : (nl < tail_from ? 1 : 2);
      const int q17 = is_t(transc) && 16 == simd_w && !fix_q17 && (ml % 16) >= 8
                      && ml < (num_m / 16) * 16 && nl < (num_n / 16) * 16;   /* row 8..15 of a complete 16 x 16 block: not scaled */
      float run = (0.f == beta) ? 0.f : ((1.f == beta || q17) ? c[cat] : beta * c[cat]);
      for (kb = 0; kb < kbc; ++kb) {
        const int s = kb * mbc + mb;
        const uint16_t* ro = rowidx + (size_t)s * (bm + 1);
        const int start = ro[ml], count = (int)ro[ml + 1] - start;
        const uint16_t* co = colidx + s * cap + start;
        const float* va = values + s * cap + start;
        float sum = 0.f;
        for (j = 0; j < count; ++j) {
          const size_t kk = (size_t)kb * bk + co[j];
          const size_t bat = is_t(transb) ? ((size_t)(n0 + nl) * K + kk) : (kk * N + n0 + nl);
          const float bv = (0 == dtype) ? ((const float*)b)[bat] : bf16_widen(((const uint16_t*)b)[bat]);
          if (!fused) {   /* scalar instantiation: _MM_FMADD_FP32(x, y, z) is ((x)*(y))+(z) in a function without FMA target (libxsmm_spmdm_begin.h:51): two roundings */
            volatile float p = va[j] * bv;
            if (1 == mode) sum = p + sum; else run = p + run;
          }
          else if (0 == mode) run = fmaf(va[j], bv, run);
          else if (1 == mode) sum = fmaf(va[j], bv, sum);
          else if (tail_fma) run = fmaf(bv, va[j], run);
          else { volatile float p = bv * va[j]; run = run + p; }
        }
        if (1 == mode) run = sum + run;
      }
      c[cat] = run;
    }
  }
}

/* plain double-precision check value (not the reference's arithmetic; sanity only) */
double orc_spmdm_gold_element(const int* g, int dtype, char transa, char transb, const void* a, const void* b, int m, int n)
{
  const int M = g[G_M], N = g[G_N], K = g[G_K];
  double s = 0; int k;
  for (k = 0; k < K; ++k) {
    const size_t aat = is_t(transa) ? ((size_t)k * M + m) : ((size_t)m * K + k);
    const size_t bat = is_t(transb) ? ((size_t)n * K + k) : ((size_t)k * N + n);
    const double av = (0 == dtype) ? ((const float*)a)[aat] : bf16_widen(((const uint16_t*)a)[aat]);
    const double bv = (0 == dtype) ? ((const float*)b)[bat] : bf16_widen(((const uint16_t*)b)[bat]);
    s += av * bv;
  }
  return s;
}
