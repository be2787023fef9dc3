/*
 * TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * OpenMP driver around the UNMODIFIED reference library built by oracle/build_ref.sh
 * (oracle/_ref/libxsmm_ref.so).  It calls the reference exactly the way its own
 * samples do:
 *   spmdm   : samples/spmdm/spmdm.c:88-111  (omp-for over createSparseSlice blocks,
 *             implicit barrier, omp-for over compute blocks)
 *   fsspmdm : samples/pyfr/pyfr_driver_asp_reg.c:268-272,297-308 (one handle with
 *             N = panel width, ldb = ldc = full row pitch; omp-for over column panels)
 * and either dumps the results (parity oracle / golden fixtures) or times them
 * (bench.py cpu_baseline and --impl reference).
 */
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "libxsmm_spmdm.h"
#include "libxsmm_fsspmdm.h"

/* mirrors the private layout at reference src/libxsmm_main.h:695-715 (read-only peek) */
typedef struct { int M, N, K, ldb, ldc, N_chunksize; void* a_dense; void* kernel; } refdrv_fs_peek;

/* geometry only: fills geom[9] = m n k bm bn bk mb nb kb, returns scratch bytes per thread */
int refdrv_spmdm_geometry(int M, int N, int K, int max_threads, int* geom)
{
  libxsmm_spmdm_handle h; libxsmm_CSR_sparseslice* s = 0;
  libxsmm_spmdm_init(M, N, K, max_threads, &h, &s);
  geom[0] = h.m; geom[1] = h.n; geom[2] = h.k; geom[3] = h.bm; geom[4] = h.bn; geom[5] = h.bk;
  geom[6] = h.mb; geom[7] = h.nb; geom[8] = h.kb;
  libxsmm_spmdm_destroy(&h);
  return h.memory_for_scratch_per_thread;
}

/*
 * dtype: 0 = fp32, 1 = bfloat16.  A, B typed accordingly; beta points at a float (fp32)
 * or at an unsigned short (bf16, raw bits as the reference reads them).
 * reps multiplies (slice + compute); times[3*r + {0,1,2}] = slice, compute, total seconds.
 * If rowidx != NULL the slices of the LAST rep are copied out as flat arrays with
 * strides (bm+1), bm*bk, bm*bk per slice; entries at and beyond rowidx[nrows] are
 * zero-filled here so that dumps are deterministic (the reference leaves scratch there).
 */
int refdrv_spmdm_run(int dtype, int M, int N, int K, int threads, int max_threads,
                     char transa, char transb, char transc,
                     const void* A, const void* B, const void* beta, float* C,
                     int reps, double* times,
                     int* geom, unsigned short* rowidx, unsigned short* colidx, float* values)
{
  libxsmm_spmdm_handle h; libxsmm_CSR_sparseslice* s = 0;
  const float alpha_f = 1.f; const libxsmm_bfloat16 alpha_h = 0x3F80;
  int r, nslice, ncomp;
  if (threads <= 0) threads = omp_get_max_threads();
  if (max_threads <= 0) max_threads = threads;
  libxsmm_spmdm_init(M, N, K, max_threads, &h, &s);
  if (0 == h.base_ptr_scratch_A || 0 == h.memory_for_scratch_per_thread) return -1;
  nslice = libxsmm_spmdm_get_num_createSparseSlice_blocks(&h);
  ncomp = libxsmm_spmdm_get_num_compute_blocks(&h);
  if (geom) {
    geom[0] = h.m; geom[1] = h.n; geom[2] = h.k; geom[3] = h.bm; geom[4] = h.bn; geom[5] = h.bk;
    geom[6] = h.mb; geom[7] = h.nb; geom[8] = h.kb;
  }
  for (r = 0; r < reps; ++r) {
    double t0 = 0, t1 = 0, t2 = 0;
#   pragma omp parallel num_threads(threads)
    {
      const int tid = omp_get_thread_num(), nth = omp_get_num_threads();
      int i;
#     pragma omp barrier
#     pragma omp master
      t0 = omp_get_wtime();
#     pragma omp for schedule(dynamic)
      for (i = 0; i < nslice; ++i) {
        if (0 == dtype) libxsmm_spmdm_createSparseSlice_fp32_thread(&h, transa, (const float*)A, s, i, tid, nth);
        else libxsmm_spmdm_createSparseSlice_bfloat16_thread(&h, transa, (const libxsmm_bfloat16*)A, s, i, tid, nth);
      }
#     pragma omp master
      t1 = omp_get_wtime();
#     pragma omp for schedule(dynamic)
      for (i = 0; i < ncomp; ++i) {
        if (0 == dtype) libxsmm_spmdm_compute_fp32_thread(&h, transa, transb, &alpha_f, s, (const float*)B, transc, (const float*)beta, C, i, tid, nth);
        else libxsmm_spmdm_compute_bfloat16_thread(&h, transa, transb, &alpha_h, s, (const libxsmm_bfloat16*)B, transc, (const libxsmm_bfloat16*)beta, C, i, tid, nth);
      }
#     pragma omp master
      t2 = omp_get_wtime();
    }
    if (times) { times[3 * r] = t1 - t0; times[3 * r + 1] = t2 - t1; times[3 * r + 2] = t2 - t0; }
  }
  if (rowidx) {
    int i, kb, mb;
    const size_t cap = (size_t)h.bm * h.bk;
    for (kb = 0; kb < h.kb; ++kb) for (mb = 0; mb < h.mb; ++mb) {
      const int sl = kb * h.mb + mb;
      const int nrows = ((mb + 1) * h.bm > h.m) ? (h.m - mb * h.bm) : h.bm;
      unsigned short* ro = rowidx + (size_t)sl * (h.bm + 1);
      unsigned int nnz = 0;
      memset(ro, 0, sizeof(unsigned short) * (h.bm + 1));
      memcpy(ro, s[sl].rowidx, sizeof(unsigned short) * (nrows + 1));
      /* true count (rowidx is u16 and wraps at 65536): sum of per-row differences */
      for (i = 0; i < nrows; ++i) nnz += (unsigned short)(s[sl].rowidx[i + 1] - s[sl].rowidx[i]);
      memset(colidx + sl * cap, 0, sizeof(unsigned short) * cap);
      memset(values + sl * cap, 0, sizeof(float) * cap);
      memcpy(colidx + sl * cap, s[sl].colidx, sizeof(unsigned short) * nnz);
      memcpy(values + sl * cap, s[sl].values, sizeof(float) * nnz);
    }
  }
  libxsmm_spmdm_destroy(&h);
  return 0;
}

/*
 * dbl: 1 = dfsspmdm, 0 = sfsspmdm.  A is M x K with pitch lda; B is K x N, C is M x N, both
 * with pitch ld.  panel = width of the column panel one execute() handles (the reference
 * driver uses 48; BASELINE.md asks for 64).  info[0] = 1 if the sparse_reg branch was taken
 * (a_dense == NULL), info[1] = N_chunksize.  times[r] = seconds of rep r.
 */
int refdrv_fsspmdm_run(int dbl, int M, int N, int K, int lda, int ld, double beta,
                       const void* A, const void* B, void* C,
                       int panel, int threads, int reps, double* times, int* info)
{
  void* h; int r; long z;
  if (threads <= 0) threads = omp_get_max_threads();
  if (panel <= 0 || 0 != (panel % 16) || 0 != (N % panel)) return -2;
  h = dbl ? (void*)libxsmm_dfsspmdm_create(M, panel, K, lda, ld, ld, 1.0, beta, (const double*)A)
          : (void*)libxsmm_sfsspmdm_create(M, panel, K, lda, ld, ld, 1.f, (float)beta, (const float*)A);
  if (0 == h) return -1;
  if (info) { info[0] = (0 == ((const refdrv_fs_peek*)h)->a_dense); info[1] = ((const refdrv_fs_peek*)h)->N_chunksize; }
  for (r = 0; r < reps; ++r) {
    const double t0 = omp_get_wtime();
#   pragma omp parallel for num_threads(threads) schedule(static)
    for (z = 0; z < N; z += panel) {
      if (dbl) libxsmm_dfsspmdm_execute((const libxsmm_dfsspmdm*)h, (const double*)B + z, (double*)C + z);
      else libxsmm_sfsspmdm_execute((const libxsmm_sfsspmdm*)h, (const float*)B + z, (float*)C + z);
    }
    if (times) times[r] = omp_get_wtime() - t0;
  }
  if (dbl) libxsmm_dfsspmdm_destroy((libxsmm_dfsspmdm*)h); else libxsmm_sfsspmdm_destroy((libxsmm_sfsspmdm*)h);
  return 0;
}

int refdrv_max_threads(void) { return omp_get_max_threads(); }

