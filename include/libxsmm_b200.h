/*
 * libxsmm_b200.h -- stream-ordered, device-pointer entry points of the B200-native SPMDM/FSSPMDM
 * library (libxsmm_b200.so) plus a few service calls.  These are ADDITIONS to the reference
 * interface declared in libxsmm_spmdm.h / libxsmm_fsspmdm.h: one call covers the whole problem
 * (all createSparseSlice blocks / all compute blocks) and is asynchronous on `stream`.
 * `stream` is a cudaStream_t passed as void* so that this header needs no CUDA headers.
 */
#ifndef LIBXSMM_B200_H
#define LIBXSMM_B200_H

#include <stddef.h>
#include "libxsmm_spmdm.h"
#include "libxsmm_fsspmdm.h"

/* All block ids of reference samples/spmdm/spmdm.c:100-104 in one launch.  d_a is a DEVICE pointer. */
LIBXSMM_API void libxsmm_spmdm_createSparseSlice_fp32_stream(const libxsmm_spmdm_handle* handle, char transa,
  const float* d_a, libxsmm_CSR_sparseslice* libxsmm_output_csr_a, void* stream);
LIBXSMM_API void libxsmm_spmdm_createSparseSlice_bfloat16_stream(const libxsmm_spmdm_handle* handle, char transa,
  const libxsmm_bfloat16* d_a, libxsmm_CSR_sparseslice* libxsmm_output_csr_a, void* stream);

/* All block ids of reference samples/spmdm/spmdm.c:106-110 in one launch.  d_b, d_c are DEVICE pointers;
 * alpha and beta are HOST pointers read at call time (alpha is ignored like in the reference). */
LIBXSMM_API void libxsmm_spmdm_compute_fp32_stream(const libxsmm_spmdm_handle* handle, char transa, char transb,
  const float* alpha, libxsmm_CSR_sparseslice* a_sparse, const float* d_b, char transc, const float* beta,
  float* d_c, void* stream);
LIBXSMM_API void libxsmm_spmdm_compute_bfloat16_stream(const libxsmm_spmdm_handle* handle, char transa, char transb,
  const libxsmm_bfloat16* alpha, libxsmm_CSR_sparseslice* a_sparse, const libxsmm_bfloat16* d_b, char transc,
  const libxsmm_bfloat16* beta, float* d_c, void* stream);

/* Whole multiply with HOST buffers: copies A and B (and C if *beta != 0) to the device, slices,
 * multiplies and copies C back; returns when C is complete.  This is what one repetition of the
 * reference sample does (samples/spmdm/spmdm.c:88-111).  datatype selects the element type of a, b
 * (the reference never fills handle->datatype, so it is an argument here); beta points at a float
 * (F32) or at a libxsmm_bfloat16 bit pattern (BFLOAT16).  Host buffers should be page-locked
 * (libxsmm_b200_host_alloc) for full PCIe speed; pageable memory works but is slower. */
LIBXSMM_API void libxsmm_spmdm_exec_host(const libxsmm_spmdm_handle* handle, libxsmm_CSR_sparseslice* slices,
  libxsmm_spmdm_datatype datatype, char transa, char transb, char transc, const void* a, const void* b,
  const void* beta, float* c);
/* Same two phases on DEVICE matrices, asynchronous on stream (slice creation, then compute). */
LIBXSMM_API void libxsmm_spmdm_exec_stream(const libxsmm_spmdm_handle* handle, libxsmm_CSR_sparseslice* slices,
  libxsmm_spmdm_datatype datatype, char transa, char transb, char transc, const void* d_a, const void* d_b,
  const void* beta, float* d_c, void* stream);

/* execute() on DEVICE pointers, asynchronous on stream (reference src/libxsmm_fsspmdm.c:260-291). */
LIBXSMM_API void libxsmm_dfsspmdm_execute_stream(const libxsmm_dfsspmdm* handle, const double* d_B, double* d_C, void* stream);
LIBXSMM_API void libxsmm_sfsspmdm_execute_stream(const libxsmm_sfsspmdm* handle, const float* d_B, float* d_C, void* stream);

/* 1 if the reference would have taken its sparse_reg branch for this operator (rows without
 * nonzeros are then left untouched, reference src/generator_spgemm_csr_asparse_reg.c:229,287). */
LIBXSMM_API int libxsmm_dfsspmdm_is_sparse(const libxsmm_dfsspmdm* handle);
LIBXSMM_API int libxsmm_sfsspmdm_is_sparse(const libxsmm_sfsspmdm* handle);
/* 1 if the operator was baked into a specialised kernel at create time, 0 = generic kernel. */
LIBXSMM_API int libxsmm_dfsspmdm_is_baked(const libxsmm_dfsspmdm* handle);
LIBXSMM_API int libxsmm_sfsspmdm_is_baked(const libxsmm_sfsspmdm* handle);
/* 1 if execute() of this float operator runs on the tensor cores (dense operators; the reference's dense
 * SMM branch, src/libxsmm_fsspmdm.c:240-248). */
LIBXSMM_API int libxsmm_sfsspmdm_is_tensor_core(const libxsmm_sfsspmdm* handle);

/* Sticky error state (the reference entry points return void and are mute unless LIBXSMM_VERBOSE). */
LIBXSMM_API int libxsmm_b200_last_error(void);
LIBXSMM_API const char* libxsmm_b200_last_error_string(void);
LIBXSMM_API void libxsmm_b200_clear_error(void);

/* Number of kernels this library has launched so far in this process. */
LIBXSMM_API unsigned long long libxsmm_b200_launch_count(void);
/* name of the kernel the last spmdm compute / fsspmdm execute call enqueued for the bulk of its work (static string) */
LIBXSMM_API const char* libxsmm_b200_last_compute_kernel(void);

/* Page-locked host memory for callers that want full PCIe speed on the host-pointer paths. */
LIBXSMM_API void* libxsmm_b200_host_alloc(size_t bytes);
LIBXSMM_API void libxsmm_b200_host_free(void* p);

/* Thin device-memory helpers so that C callers and the test harness need no CUDA headers. */
LIBXSMM_API void* libxsmm_b200_device_alloc(size_t bytes);
LIBXSMM_API void libxsmm_b200_device_free(void* p);
LIBXSMM_API int libxsmm_b200_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes);
LIBXSMM_API int libxsmm_b200_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes);
LIBXSMM_API int libxsmm_b200_memset(void* dst_device, int value, size_t bytes);
LIBXSMM_API int libxsmm_b200_synchronize(void);
LIBXSMM_API int libxsmm_b200_device_count(void);
LIBXSMM_API int libxsmm_b200_set_device(int device);

/* Streams, events, asynchronous copies and graph capture, again so that plain C callers (and the
 * ctypes harness) can drive the stream entries without CUDA headers.  A stream/event/graph is an
 * opaque pointer (cudaStream_t / cudaEvent_t / cudaGraphExec_t underneath). */
LIBXSMM_API void* libxsmm_b200_stream_create(void);
LIBXSMM_API void libxsmm_b200_stream_destroy(void* stream);
LIBXSMM_API int libxsmm_b200_stream_synchronize(void* stream);
LIBXSMM_API void* libxsmm_b200_event_create(void);
LIBXSMM_API void libxsmm_b200_event_destroy(void* event);
LIBXSMM_API int libxsmm_b200_event_record(void* event, void* stream);
LIBXSMM_API int libxsmm_b200_event_synchronize(void* event);
LIBXSMM_API float libxsmm_b200_event_elapsed_ms(void* start, void* stop);
LIBXSMM_API int libxsmm_b200_memcpy_h2d_async(void* dst_device, const void* src_host, size_t bytes, void* stream);
LIBXSMM_API int libxsmm_b200_memcpy_d2h_async(void* dst_host, const void* src_device, size_t bytes, void* stream);
LIBXSMM_API int libxsmm_b200_memset_async(void* dst_device, int value, size_t bytes, void* stream);
/* Capture everything enqueued on stream between begin and end into an executable graph. */
LIBXSMM_API int libxsmm_b200_graph_begin(void* stream);
LIBXSMM_API void* libxsmm_b200_graph_end(void* stream);
LIBXSMM_API int libxsmm_b200_graph_launch(void* graph_exec, void* stream);
LIBXSMM_API void libxsmm_b200_graph_destroy(void* graph_exec);

/* Fused caller step: C = A.B (+ beta.C) on DEVICE matrices in ONE stream-ordered call -- handle lookup, slice
 * creation and compute.  This is the wrapper TensorFlow's sparse_matmul_op keeps around the reference
 * (reference documentation/tensorflow.md:241-250; in-tree model samples/spmdm/spmdm.c:74-154): handles are cached
 * by (M, N, K, max_threads, stream), at most 16, least recently used evicted.  Arguments as
 * libxsmm_spmdm_exec_stream; returns 0 or the library's error code. */
LIBXSMM_API int libxsmm_b200_sparse_matmul(libxsmm_spmdm_datatype datatype, char transa, char transb, char transc,
  int M, int N, int K, int max_threads, const void* d_a, const void* d_b, const void* beta, float* d_c, void* stream);
LIBXSMM_API int libxsmm_b200_sparse_matmul_cache_entries(void);
LIBXSMM_API void libxsmm_b200_sparse_matmul_cache_clear(void);

/* MatrixMarket coordinate files (the operator format of samples/pyfr/mats and samples/edge/mats; reference reader
 * src/generator_spgemm_csr_reader.c:46-169): read into CSR (malloc'ed arrays, release with libxsmm_b200_csr_free;
 * returns 0 or a negative error code), or straight into a fixed operator (lda = K = columns of the file, alpha = 1).
 * Host-only parsing; create_mtx then calls the ordinary create(). */
LIBXSMM_API int libxsmm_b200_csr_read_mtx(const char* path, unsigned int** row_ptr, unsigned int** col_idx, double** values,
  unsigned int* rows, unsigned int* cols, unsigned int* nnz);
LIBXSMM_API void libxsmm_b200_csr_free(unsigned int* row_ptr, unsigned int* col_idx, double* values);
LIBXSMM_API libxsmm_dfsspmdm* libxsmm_b200_dfsspmdm_create_mtx(const char* path, int N, int ldb, int ldc, double beta, int* M, int* K);
LIBXSMM_API libxsmm_sfsspmdm* libxsmm_b200_sfsspmdm_create_mtx(const char* path, int N, int ldb, int ldc, float beta, int* M, int* K);

/* CSR x dense SoA kernels: the GPU counterpart of
 *     kernel = libxsmm_create_xcsr_soa(descriptor(m, n, k, lda, ldb, ldc, alpha = 1, beta), row_ptr, column_idx, values)
 *     kernel(a, b, c)                                        once per mesh element
 * (reference src/template/libxsmm.h:283-293, src/libxsmm_main.c:2423-2447, src/generator_spgemm_csr_asparse_soa.c,
 * src/generator_spgemm_csr_bsparse_soa.c; callers samples/edge/asparse_srsoa.c:148-160, samples/edge/bsparse_srsoa.c:160-176).
 * As in the reference's descriptor exactly one of lda / ldb is 0 and names the SPARSE operand (CSR):
 *   lda == 0  A sparse (CSR over its m rows), B dense [k][ldb][soa_width], C [m][ldc][soa_width]:
 *             C[m][n][s] = (beta == 0 ? 0 : C[m][n][s]) + sum over row m's nonzeros z, in CSR order, of values[z] * B[column_idx[z]][n][s];
 *             rows of A WITHOUT nonzeros leave their C rows untouched, like the reference's emitted code;
 *   ldb == 0  B sparse (CSR over its k rows), A dense [m][lda][soa_width], C [m][ldc][soa_width]:
 *             C[m][n][s] = (beta == 0 ? 0 : C[m][n][s]) + sum over k ascending, over row k's nonzeros (k, n), of A[m][k][s] * values[z];
 *             columns 0 .. ncols-1 of C are written, ncols = 1 + the largest column index any nonzero of B holds (an empty
 *             column below ncols gives beta * C; columns from ncols on are never touched, not even for beta = 0 -- the
 *             reference's generator, generator_spgemm_csr_bsparse_soa.c:161-167); nonzeros with column >= n are not
 *             multiplied; ncols > ldc is rejected.
 * One fused multiply-add per nonzero; beta is 0 or 1, as for the reference's descriptor.  soa_width is a property of the caller's
 * tensors (the reference's generator fixes it per host: 8 doubles / 16 floats with AVX-512, 4 / 8 otherwise).  One element is far
 * too small for a launch, so execute is BATCHED: n_elements elements, the dense operand of element e at d_X + e * stride_x and its
 * C at d_C + e * stride_c (strides in scalars), asynchronous on `stream`.  The sparse operand's values are fixed at create (the
 * reference's kernel re-reads them at every call; its callers pass the same array). */
typedef struct libxsmm_b200_csr_soa libxsmm_b200_csr_soa;
LIBXSMM_API libxsmm_b200_csr_soa* libxsmm_b200_dcsr_soa_create(int M, int N, int K, int lda, int ldb, int ldc, int soa_width, double beta,
  const unsigned int* row_ptr, const unsigned int* column_idx, const double* values);
LIBXSMM_API libxsmm_b200_csr_soa* libxsmm_b200_scsr_soa_create(int M, int N, int K, int lda, int ldb, int ldc, int soa_width, float beta,
  const unsigned int* row_ptr, const unsigned int* column_idx, const float* values);
/* libxsmm_create_xcsc_soa + kernel(a, values, c) (reference src/template/libxsmm.h:283-293, src/libxsmm_main.c:2450-2474,
 * src/generator_spgemm_csc_bsparse_soa.c; caller samples/edge/bsparse_scsoa.c:327-354): B sparse in CSC (column_ptr over its n
 * columns, row_idx = k), A dense [m][lda][soa_width], C [m][ldc][soa_width];
 *     C[m][n][s] = (beta == 0 ? 0 : C[m][n][s]) + sum over k ascending of A[m][k][s] * (the first entry of column n with row index k).
 * All n columns are written.  The handle is executed (d_X = A) and destroyed with the csr_soa entries below. */
LIBXSMM_API libxsmm_b200_csr_soa* libxsmm_b200_dcsc_soa_create(int M, int N, int K, int lda, int ldc, int soa_width, double beta,
                                                               const unsigned int* column_ptr, const unsigned int* row_idx, const double* values);
LIBXSMM_API libxsmm_b200_csr_soa* libxsmm_b200_scsc_soa_create(int M, int N, int K, int lda, int ldc, int soa_width, float beta,
                                                               const unsigned int* column_ptr, const unsigned int* row_idx, const float* values);
LIBXSMM_API void libxsmm_b200_csr_soa_execute(const libxsmm_b200_csr_soa* handle, const void* d_X, void* d_C, long long n_elements,
  long long stride_x, long long stride_c, void* stream);
LIBXSMM_API int libxsmm_b200_csr_soa_is_baked(const libxsmm_b200_csr_soa* handle);
LIBXSMM_API void libxsmm_b200_csr_soa_destroy(libxsmm_b200_csr_soa* handle);

/* Dense SMM dispatch for row-major operators with very many columns: the GPU counterpart of
 *     kernel = libxsmm_dmmdispatch(nblock, M_op, K_op, &ldb_panel, &lda_op, &ldc_panel, alpha, beta, flags, prefetch)
 *     for (i = 0; i < N; i += nblock) kernel(B_panel + i, A_op, C_panel + i)
 * (reference src/libxsmm_main.c:2166-2195; caller samples/pyfr/pyfr_gemm_rm.c:98-122).  Same argument meaning as
 * libxsmm_[sd]mmdispatch: column-major C(m x n) = A(m x k) B(k x n) + beta C with NULL lda / ldb / ldc = m / k / m and
 * NULL alpha / beta = 1; like the reference's JIT it exists for alpha = 1 and beta in {0, 1} only (NULL otherwise).
 * execute covers m_total rows of A and C -- all the chunks of the caller's loop -- in one asynchronous call: d_a, d_c
 * are DEVICE pointers, b (the small k x n operand = the row-major operator, pitch ldb) may be a host or a device
 * pointer and is read at every call; the handle keeps the kernel baked for the operand's current content and re-bakes when
 * it changes.  Per output element: an in-order chain of fused multiply-adds over k starting from C (beta = 1) or 0. */
typedef struct libxsmm_b200_mm libxsmm_b200_mm;
LIBXSMM_API libxsmm_b200_mm* libxsmm_b200_dmmdispatch(int m, int n, int k, const int* lda, const int* ldb, const int* ldc,
  const double* alpha, const double* beta);
LIBXSMM_API libxsmm_b200_mm* libxsmm_b200_smmdispatch(int m, int n, int k, const int* lda, const int* ldb, const int* ldc,
  const float* alpha, const float* beta);
LIBXSMM_API void libxsmm_b200_mm_execute(libxsmm_b200_mm* handle, const void* d_a, const void* b, void* d_c, long long m_total, void* stream);
LIBXSMM_API const char* libxsmm_b200_mm_kernel(const libxsmm_b200_mm* handle);   /* which kernel the handle currently holds */
LIBXSMM_API void libxsmm_b200_mm_release(libxsmm_b200_mm* handle);

/* Host-only planning entries (no CUDA call is made; they work on a machine without a GPU).
 * geometry: the block geometry libxsmm_spmdm_init would choose (reference src/libxsmm_spmdm.c:552-608)
 *   for bn = 48 | 96 | 6; geom[9] = m n k bm bn bk mb nb kb.  Returns 0 on success.
 * plan: what create() decides for an operator (reference src/libxsmm_fsspmdm.c:88-143 and
 *   src/generator_spgemm_csr_asparse_reg.c:111-150): info[6] = nnz, unique values, 1 if the reference
 *   takes its sparse_reg branch, bytes of x86 code it would emit (0 if not evaluated), N_chunksize, and the
 *   form of the kernel create() bakes (0 none: generic kernel, 1 B rows in registers, 2 B strip in shared memory).
 * kernel_source: the CUDA source create() would bake for the operator (free with free_string). */
LIBXSMM_API int libxsmm_b200_spmdm_geometry(int M, int N, int K, int max_threads, int bn, int* geom);
LIBXSMM_API int libxsmm_b200_fsspmdm_plan(int is_double, int M, int N, int K, int lda, int ldb, int ldc, double beta,
  const void* a_dense, long long* info);
LIBXSMM_API char* libxsmm_b200_fsspmdm_kernel_source(int is_double, int M, int N, int K, int lda, int ldb, int ldc,
  double beta, const void* a_dense);
LIBXSMM_API void libxsmm_b200_free_string(char* s);

#endif /*LIBXSMM_B200_H*/
