/*
 * TEST INFRASTRUCTURE ONLY (never linked into the product).  Second part of the driver around the UNMODIFIED reference
 * (oracle/_ref/libxsmm_ref.so), compiled against the reference's OWN headers (the first part, ref_driver.c, is compiled
 * against this repository's drop-in headers on purpose).
 */
/*
 * CSR "A sparse" x dense SoA kernels: libxsmm_create_xcsr_soa, called the way
 * samples/edge/asparse_srsoa.c:148-160 does -- descriptor (m, n, k, lda = 0, ldb, ldc, alpha = 1, beta, 'N','N', no prefetch),
 * then kernel(values, B, C) once per element.  B is [k][ldb][soa], C is [m][ldc][soa] (soa = 8 doubles / 16 floats on an
 * AVX-512 host, 4 / 8 otherwise: returned in *soa_width_used by probing the target the way the generator selects it).
 * n_elem elements, strides in scalars.  Returns 0, or -1 if the kernel could not be generated.
 */
#include "libxsmm.h"
int refdrv_csr_soa_run_ex(int dbl, int M, int N, int K, int lda, int ldb, int ldc, double beta,
                          const unsigned int* rowptr, const unsigned int* colidx, const void* values,
                          const void* X, void* C, long n_elem, long stride_x, long stride_c, int* soa_width_used);

int refdrv_csr_soa_run(int dbl, int M, int N, int K, int ldb, int ldc, double beta,
                       const unsigned int* rowptr, const unsigned int* colidx, const void* values,
                       const void* B, void* C, long n_elem, long stride_b, long stride_c, int* soa_width_used)
{ return refdrv_csr_soa_run_ex(dbl, M, N, K, 0, ldb, ldc, beta, rowptr, colidx, values, B, C, n_elem, stride_b, stride_c, soa_width_used); }

/* lda == 0: A sparse (CSR over m), X = B dense [k][ldb][soa], kernel(values, X, C).
 * ldb == 0: B sparse (CSR over k), X = A dense [m][lda][soa], kernel(X, values, C)  (samples/edge/bsparse_srsoa.c:160-176). */
int refdrv_csr_soa_run_ex(int dbl, int M, int N, int K, int lda, int ldb, int ldc, double beta,
                          const unsigned int* rowptr, const unsigned int* colidx, const void* values,
                          const void* B, void* C, long n_elem, long stride_b, long stride_c, int* soa_width_used)
{
  libxsmm_descriptor_blob blob;
  const int flags = LIBXSMM_GEMM_FLAGS('N', 'N');
  long e;
  libxsmm_init();
  if (soa_width_used) {
    const int avx512 = (libxsmm_get_target_archid() >= LIBXSMM_X86_AVX512);
    *soa_width_used = dbl ? (avx512 ? 8 : 4) : (avx512 ? 16 : 8);
  }
  if (dbl) {
    const libxsmm_gemm_descriptor* d = libxsmm_gemm_descriptor_dinit(&blob, LIBXSMM_GEMM_PRECISION_F64, M, N, K, lda, ldb, ldc, 1.0, beta, flags, LIBXSMM_GEMM_PREFETCH_NONE);
    libxsmm_dmmfunction kern = (0 != d) ? libxsmm_create_xcsr_soa(d, rowptr, colidx, values).dmm : 0;
    if (0 == kern) return -1;
    for (e = 0; e < n_elem; ++e) {
      if (0 == lda) kern((const double*)values, (const double*)B + e * stride_b, (double*)C + e * stride_c);
      else kern((const double*)B + e * stride_b, (const double*)values, (double*)C + e * stride_c);
    }
    libxsmm_release_kernel((const void*)kern);
  }
  else {
    const libxsmm_gemm_descriptor* d = libxsmm_gemm_descriptor_dinit(&blob, LIBXSMM_GEMM_PRECISION_F32, M, N, K, lda, ldb, ldc, 1.0, beta, flags, LIBXSMM_GEMM_PREFETCH_NONE);
    libxsmm_smmfunction kern = (0 != d) ? libxsmm_create_xcsr_soa(d, rowptr, colidx, values).smm : 0;
    if (0 == kern) return -1;
    for (e = 0; e < n_elem; ++e) {
      if (0 == lda) kern((const float*)values, (const float*)B + e * stride_b, (float*)C + e * stride_c);
      else kern((const float*)B + e * stride_b, (const float*)values, (float*)C + e * stride_c);
    }
    libxsmm_release_kernel((const void*)kern);
  }
  return 0;
}

/*
 * Dense SMM dispatch used on a row-major panel, exactly as samples/pyfr/pyfr_gemm_rm.c:98-122:
 *     kernel = libxsmm_dmmdispatch(nblock, M, K, &ld_panel, &lda_op, &ld_panel, &alpha, &beta, NULL, &prefetch_none)
 *     for (i = 0; i < N; i += nblock) kernel(B + i, A, C + i)
 * A: M x K row-major operator (pitch lda_op); B: K x N panel, C: M x N panel (pitch ld).  N % nblock == 0.
 */
int refdrv_mm_rm_run(int dbl, int M, int N, int K, int lda_op, int ld, double beta, int nblock, const void* A, const void* B, void* C)
{
  const int prefetch = LIBXSMM_GEMM_PREFETCH_NONE;
  int i;
  if (nblock <= 0 || 0 != (N % nblock)) return -2;
  libxsmm_init();
  if (dbl) {
    const double alpha = 1.0;
    libxsmm_dmmfunction kern = libxsmm_dmmdispatch(nblock, M, K, &ld, &lda_op, &ld, &alpha, &beta, NULL, &prefetch);
    if (0 == kern) return -1;
    for (i = 0; i < N; i += nblock) kern((const double*)B + i, (const double*)A, (double*)C + i);
  }
  else {
    const float alpha = 1.f, fbeta = (float)beta;
    libxsmm_smmfunction kern = libxsmm_smmdispatch(nblock, M, K, &ld, &lda_op, &ld, &alpha, &fbeta, NULL, &prefetch);
    if (0 == kern) return -1;
    for (i = 0; i < N; i += nblock) kern((const float*)B + i, (const float*)A, (float*)C + i);
  }
  return 0;
}

/* the same kernel timed the way an element loop would run it on the host: OpenMP over elements, reps repetitions */
#include <omp.h>
int refdrv_csr_soa_bench_ex(int dbl, int M, int N, int K, int lda, int ldb, int ldc, double beta,
                            const unsigned int* rowptr, const unsigned int* colidx, const void* values,
                            const void* X, void* C, long n_elem, long stride_x, long stride_c, int threads, int reps, double* times)
{
  libxsmm_descriptor_blob blob;
  const int flags = LIBXSMM_GEMM_FLAGS('N', 'N');
  const libxsmm_gemm_descriptor* d;
  libxsmm_xmmfunction kern;
  int r; long e;
  libxsmm_init();
  if (threads <= 0) threads = omp_get_max_threads();
  d = dbl ? libxsmm_gemm_descriptor_dinit(&blob, LIBXSMM_GEMM_PRECISION_F64, M, N, K, lda, ldb, ldc, 1.0, beta, flags, LIBXSMM_GEMM_PREFETCH_NONE)
          : libxsmm_gemm_descriptor_dinit(&blob, LIBXSMM_GEMM_PRECISION_F32, M, N, K, lda, ldb, ldc, 1.0, beta, flags, LIBXSMM_GEMM_PREFETCH_NONE);
  if (0 == d) return -1;
  kern = libxsmm_create_xcsr_soa(d, rowptr, colidx, values);
  if (0 == kern.dmm) return -1;
  for (r = 0; r < reps; ++r) {
    const double t0 = omp_get_wtime();
#   pragma omp parallel for num_threads(threads) schedule(static)
    for (e = 0; e < n_elem; ++e) {
      if (0 == lda) {   /* A sparse: kernel(values, B, C) */
        if (dbl) kern.dmm((const double*)values, (const double*)X + e * stride_x, (double*)C + e * stride_c);
        else kern.smm((const float*)values, (const float*)X + e * stride_x, (float*)C + e * stride_c);
      }
      else {            /* B sparse: kernel(A, values, C) */
        if (dbl) kern.dmm((const double*)X + e * stride_x, (const double*)values, (double*)C + e * stride_c);
        else kern.smm((const float*)X + e * stride_x, (const float*)values, (float*)C + e * stride_c);
      }
    }
    if (times) times[r] = omp_get_wtime() - t0;
  }
  libxsmm_release_kernel((const void*)kern.dmm);
  return 0;
}

int refdrv_csr_soa_bench(int dbl, int M, int N, int K, int ldb, int ldc, double beta,
                         const unsigned int* rowptr, const unsigned int* colidx, const void* values,
                         const void* B, void* C, long n_elem, long stride_b, long stride_c, int threads, int reps, double* times)
{ return refdrv_csr_soa_bench_ex(dbl, M, N, K, 0, ldb, ldc, beta, rowptr, colidx, values, B, C, n_elem, stride_b, stride_c, threads, reps, times); }

/* libxsmm_create_xcsc_soa, called like samples/edge/bsparse_scsoa.c:327-354: descriptor (m, n, k, lda, 0, ldc), kernel(A, values, C) */
int refdrv_csc_soa_run(int dbl, int M, int N, int K, int lda, int ldc, double beta,
                       const unsigned int* colptr, const unsigned int* rowidx, const void* values,
                       const void* A, void* C, long n_elem, long stride_a, long stride_c, int* soa_width_used)
{
  libxsmm_descriptor_blob blob;
  const int flags = LIBXSMM_GEMM_FLAGS('N', 'N');
  const libxsmm_gemm_descriptor* d;
  libxsmm_xmmfunction kern;
  long e;
  libxsmm_init();
  if (soa_width_used) {
    const int avx512 = (libxsmm_get_target_archid() >= LIBXSMM_X86_AVX512);
    *soa_width_used = dbl ? (avx512 ? 8 : 4) : (avx512 ? 16 : 8);
  }
  d = libxsmm_gemm_descriptor_dinit(&blob, dbl ? LIBXSMM_GEMM_PRECISION_F64 : LIBXSMM_GEMM_PRECISION_F32, M, N, K, lda, 0, ldc, 1.0, beta, flags, LIBXSMM_GEMM_PREFETCH_NONE);
  if (0 == d) return -1;
  kern = libxsmm_create_xcsc_soa(d, colptr, rowidx, values);
  if (0 == kern.dmm) return -1;
  for (e = 0; e < n_elem; ++e) {
    if (dbl) kern.dmm((const double*)A + e * stride_a, (const double*)values, (double*)C + e * stride_c);
    else kern.smm((const float*)A + e * stride_a, (const float*)values, (float*)C + e * stride_c);
  }
  libxsmm_release_kernel((const void*)kern.dmm);
  return 0;
}
