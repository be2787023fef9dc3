/*
 * libxsmm_spmdm.h -- drop-in declaration of the LIBXSMM SPMDM interface for the
 * B200-native implementation (libxsmm_b200.so).
 *
 * Replaces:  reference include/libxsmm_spmdm.h:37-132 (same include guard, same
 * struct layouts, same eight entry points, same argument meaning).  Written from
 * the interface description, not copied: only names, field order and prototypes
 * are shared because they ARE the ABI (reference .abi.txt:376-383).
 *
 * Pointer arguments (a, b, c) of the *_thread entries may be HOST or DEVICE
 * pointers; the library tells them apart with cudaPointerGetAttributes.  The two
 * char* members of the handle carry DEVICE addresses (slice arena, staging arena).
 * Stream-ordered whole-problem entries live in libxsmm_b200.h.
 */
#ifndef LIBXSMM_SPMDM_H
#define LIBXSMM_SPMDM_H

#include <stdint.h>

#if !defined(LIBXSMM_API)
# if defined(__cplusplus)
#   define LIBXSMM_API extern "C" __attribute__((visibility("default")))
# else
#   define LIBXSMM_API extern __attribute__((visibility("default")))
# endif
#endif
#if !defined(LIBXSMM_TYPEDEFS_H) && !defined(LIBXSMM_B200_TYPEDEFS)
# define LIBXSMM_B200_TYPEDEFS
/* reference include/libxsmm_typedefs.h:119 and :37-47 (LP64 build) */
typedef unsigned short libxsmm_bfloat16;
typedef int libxsmm_blasint;
#endif

/* reference include/libxsmm_spmdm.h:37-40 */
typedef enum libxsmm_spmdm_datatype {
  LIBXSMM_SPMDM_DATATYPE_F32,
  LIBXSMM_SPMDM_DATATYPE_BFLOAT16
} libxsmm_spmdm_datatype;

/* reference include/libxsmm_spmdm.h:42-60.  A is m x k (sparse), B is k x n, C is m x n.
 * bm x bk is the slice shape (bk = 128), bn the width of one legacy compute block. */
typedef struct libxsmm_spmdm_handle {
  int m, n, k;
  int bm, bn, bk;
  int mb, nb, kb;
  libxsmm_spmdm_datatype datatype;
  char* base_ptr_scratch_A;            /* DEVICE: CSR slice arena */
  char* base_ptr_scratch_B_scratch_C;  /* DEVICE: per-tid staging arena */
  int memory_for_scratch_per_thread;
} libxsmm_spmdm_handle;

/* reference include/libxsmm_spmdm.h:66-71.  One bm x bk block of A as CSR with
 * block-local 16-bit indices.  The three pointers are DEVICE pointers. */
typedef struct libxsmm_CSR_sparseslice {
  uint16_t* rowidx;
  uint16_t* colidx;
  float*    values;
} libxsmm_CSR_sparseslice;

/* reference include/libxsmm_spmdm.h:74-78, src/libxsmm_spmdm.c:540-627 */
LIBXSMM_API void libxsmm_spmdm_init(int M, int N, int K, int max_threads,
  libxsmm_spmdm_handle* handle, libxsmm_CSR_sparseslice** libxsmm_output_csr);

/* reference include/libxsmm_spmdm.h:80-81, src/libxsmm_spmdm.c:182-185 */
LIBXSMM_API void libxsmm_spmdm_destroy(libxsmm_spmdm_handle* handle);

/* reference include/libxsmm_spmdm.h:83-87, src/libxsmm_spmdm.c:188-197 */
LIBXSMM_API int libxsmm_spmdm_get_num_createSparseSlice_blocks(const libxsmm_spmdm_handle* handle);
LIBXSMM_API int libxsmm_spmdm_get_num_compute_blocks(const libxsmm_spmdm_handle* handle);

/* reference include/libxsmm_spmdm.h:90-96, src/libxsmm_spmdm.c:253-272 */
LIBXSMM_API void libxsmm_spmdm_createSparseSlice_fp32_thread(
  const libxsmm_spmdm_handle* handle, char transa, const float* a,
  libxsmm_CSR_sparseslice* libxsmm_output_csr_a, int block_id, int tid, int nthreads);

/* reference include/libxsmm_spmdm.h:98-104, src/libxsmm_spmdm.c:328-347 */
LIBXSMM_API void libxsmm_spmdm_createSparseSlice_bfloat16_thread(
  const libxsmm_spmdm_handle* handle, char transa, const libxsmm_bfloat16* a,
  libxsmm_CSR_sparseslice* libxsmm_output_csr_a, int block_id, int tid, int nthreads);

/* reference include/libxsmm_spmdm.h:106-119, src/libxsmm_spmdm.c:418-442 (alpha is ignored) */
LIBXSMM_API void libxsmm_spmdm_compute_fp32_thread(
  const libxsmm_spmdm_handle* handle, char transa, char transb, const float* alpha,
  libxsmm_CSR_sparseslice* a_sparse, const float* b, char transc, const float* beta,
  float* c, int block_id, int tid, int nthreads);

/* reference include/libxsmm_spmdm.h:121-132, src/libxsmm_spmdm.c:513-537 (alpha is ignored;
 * *beta is read as an INTEGER bit pattern exactly like the reference does) */
LIBXSMM_API void libxsmm_spmdm_compute_bfloat16_thread(
  const libxsmm_spmdm_handle* handle, char transa, char transb, const libxsmm_bfloat16* alpha,
  libxsmm_CSR_sparseslice* a_sparse, const libxsmm_bfloat16* b, char transc,
  const libxsmm_bfloat16* beta, float* c, int block_id, int tid, int nthreads);

#endif /*LIBXSMM_SPMDM_H*/
