/*
 * libxsmm_fsspmdm.h -- drop-in declaration of the LIBXSMM fixed-size sparse-A times
 * dense-B interface for the B200-native implementation (libxsmm_b200.so).
 *
 * Replaces:  reference include/libxsmm_fsspmdm.h:37-57 (same guard, opaque handle
 * types and six entry points; reference .abi.txt:47-49,352-354).
 *
 * C[0:M, 0:N] (row pitch ldc) = A[0:M, 0:K] * B[0:K, 0:N] (row pitch ldb) + beta*C.
 * A (row pitch lda) is read once at create time and baked into the operator.
 * B and C of execute() may be HOST or DEVICE pointers.
 */
#ifndef LIBXSMM_FSSPMDM_H
#define LIBXSMM_FSSPMDM_H

#if !defined(LIBXSMM_API)
# if defined(__cplusplus)
#   define LIBXSMM_API extern "C" __attribute__((visibility("default")))
# else
#   define LIBXSMM_API extern __attribute__((visibility("default")))
# endif
#endif
#if !defined(LIBXSMM_TYPEDEFS_H) && !defined(LIBXSMM_B200_TYPEDEFS)
# define LIBXSMM_B200_TYPEDEFS
typedef unsigned short libxsmm_bfloat16;
typedef int libxsmm_blasint;
#endif

/* reference include/libxsmm_fsspmdm.h:38-39 (layout private: src/libxsmm_main.h:695-715) */
typedef struct libxsmm_dfsspmdm libxsmm_dfsspmdm;
typedef struct libxsmm_sfsspmdm libxsmm_sfsspmdm;

/* reference include/libxsmm_fsspmdm.h:41-44, src/libxsmm_fsspmdm.c:48-151.
 * Contract (asserted by the reference, checked here): N % 16 == 0, N >= 16, alpha == 1,
 * beta in {0, 1}, K <= lda, N <= ldb, N <= ldc.  Returns NULL on violation. */
LIBXSMM_API libxsmm_dfsspmdm* libxsmm_dfsspmdm_create(
  libxsmm_blasint M, libxsmm_blasint N, libxsmm_blasint K,
  libxsmm_blasint lda, libxsmm_blasint ldb, libxsmm_blasint ldc,
  const double alpha, const double beta, const double* a_dense);
/* reference include/libxsmm_fsspmdm.h:46, src/libxsmm_fsspmdm.c:260-274 */
LIBXSMM_API void libxsmm_dfsspmdm_execute(const libxsmm_dfsspmdm* handle, const double* B, double* C);
/* reference include/libxsmm_fsspmdm.h:48, src/libxsmm_fsspmdm.c:294-310 */
LIBXSMM_API void libxsmm_dfsspmdm_destroy(libxsmm_dfsspmdm* handle);

/* reference include/libxsmm_fsspmdm.h:50-53, src/libxsmm_fsspmdm.c:154-257 */
LIBXSMM_API libxsmm_sfsspmdm* libxsmm_sfsspmdm_create(
  libxsmm_blasint M, libxsmm_blasint N, libxsmm_blasint K,
  libxsmm_blasint lda, libxsmm_blasint ldb, libxsmm_blasint ldc,
  const float alpha, const float beta, const float* a_dense);
/* reference include/libxsmm_fsspmdm.h:55, src/libxsmm_fsspmdm.c:277-291 */
LIBXSMM_API void libxsmm_sfsspmdm_execute(const libxsmm_sfsspmdm* handle, const float* B, float* C);
/* reference include/libxsmm_fsspmdm.h:57, src/libxsmm_fsspmdm.c:313-329 */
LIBXSMM_API void libxsmm_sfsspmdm_destroy(libxsmm_sfsspmdm* handle);

#endif /*LIBXSMM_FSSPMDM_H*/
