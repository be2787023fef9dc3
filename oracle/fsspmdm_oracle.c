/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's fixed-size
 * sparse-A x dense-B path (libxsmm_[sd]fsspmdm).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this; the product never links or calls it.
 *
 * Pinned: tests/test_oracle_vs_ref.py compares it bit-for-bit with the UNMODIFIED reference
 * compiled by oracle/build_ref.sh (both the JIT'd sparse_reg branch and the dense SMM
 * branch), including the real PyFR operators under samples/pyfr/mats.
 *
 * What is restated (paths relative to /root/reference):
 *   create    src/libxsmm_fsspmdm.c:48-151 (double), :154-257 (float)
 *   branch    src/generator_spgemm_csr_asparse_reg.c:111-150,187-191; src/libxsmm_main.c:1414-1423
 *   execute   src/libxsmm_fsspmdm.c:260-291 and the emitted kernel, generator :227-300
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

/*
 * Plan for a double operator: CSR in row-major scan order keeping a != 0.0 (so NaN is
 * kept, -0.0 is dropped; src/libxsmm_fsspmdm.c:88-116), then the generator's "unique
 * value" table (generator :125-150): value u is mapped to the LAST table entry z with
 * !(t[z] < v) && !(t[z] > v) -- for a NaN v that is the last entry of the table so far,
 * i.e. the NaN is REPLACED by that entry's value; otherwise v is appended.
 * out_val[u] is the value the emitted kernel actually multiplies with.
 * Returns nnz; *n_unique gets the table size (0 if nnz == 0).
 */
int orc_dfsspmdm_plan(int M, int K, int lda, const double* a,
                      uint32_t* rowptr, uint32_t* colidx, double* out_val, int* n_unique)
{
  int i, j, n = 0, nu = 0, u, z;
  double* table;
  for (i = 0; i < M; ++i) {
    rowptr[i] = (uint32_t)n;
    for (j = 0; j < K; ++j) {
      const double v = a[(size_t)i * lda + j];
      if (v != 0.0) { out_val[n] = v; colidx[n] = (uint32_t)j; ++n; }
    }
  }
  rowptr[M] = (uint32_t)n;
  if (0 == n) { *n_unique = 0; return 0; }
  table = (double*)malloc(sizeof(double) * n);
  table[0] = out_val[0]; nu = 1;
  for (u = 1; u < n; ++u) {
    int hit = -1;
    for (z = 0; z < nu; ++z) if (!(table[z] < out_val[u]) && !(table[z] > out_val[u])) hit = z;
    if (hit < 0) { table[nu] = out_val[u]; ++nu; }
    else out_val[u] = table[hit];
  }
  free(table);
  *n_unique = nu;
  return n;
}

/*
 * Size in bytes of the x86 kernel the reference would emit for this operator; the sparse
 * branch only exists while it fits the 128 KiB JIT buffer (src/libxsmm_main.c:66-67,1256).
 * Byte counts follow the reference's own encoder (src/generator_x86_instructions.c): an
 * EVEX memory operand is 6 bytes + displacement, a prefetcht2 is 3 bytes + displacement
 * (+1 if the base register needs REX), displacement = 0 bytes if 0, 1 byte if it is a
 * multiple of the access size (64 for zmm, 1 for prefetch... see encoder :52-77) within
 * [-128,127] units, else 4.  Validated against the reference by probing the fallback
 * threshold (tests/test_oracle_vs_ref.py::test_fsspmdm_code_size_rule).
 */
static int disp_bytes(long off, int unit, int forced)
{
  if (0 == off && !forced) return 0;
  if (0 == (off % unit) && off / unit <= 127 && off / unit >= -128) return 1;
  return 4;
}

long orc_dfsspmdm_code_size(int M, const uint32_t* rowptr, const uint32_t* colidx, int n_unique,
                            int ldb, int ldc, int beta_one)
{
  /* prologue/epilogue (push/pop of callee-saved registers, ret) and the per-unique-value
   * constant loads: vbroadcast-free full-vector load of a 64-byte immediate block that the
   * encoder places in the instruction stream with a jump over it. */
  long sz = 0;
  int m; uint32_t u;
  sz += 40;                                  /* open_stream + close_stream */
  sz += (long)n_unique * (64 + 2 + 10);       /* jmp over 64-byte constant + rip-relative vmovupd */
  for (m = 0; m < M; ++m) {
    const uint32_t cnt = rowptr[m + 1] - rowptr[m];
    const long coff = (long)m * ldc * 8;
    if (0 == cnt) continue;
    sz += beta_one ? (6 + disp_bytes(coff, 64, 0)) : 6;     /* vmovupd load of C, or vpxord */
    sz += 3 + disp_bytes(coff + 64, 1, 0);                  /* prefetcht2 C + 64 */
    for (u = rowptr[m]; u < rowptr[m + 1]; ++u) {
      const long boff = (long)colidx[u] * ldb * 8;
      sz += 6 + disp_bytes(boff, 64, 0);                    /* vfmadd231pd zmm, zmm, [B+off] */
      sz += 3 + disp_bytes(boff + 64, 1, 0);                /* prefetcht2 B + off + 64 */
    }
    sz += 6 + disp_bytes(coff, 64, 0);                      /* vmovupd store of C */
  }
  return sz;
}

/*
 * Apply the operator.  branch 1 = sparse_reg kernel: rows WITHOUT nonzeros are not touched
 * at all, not even for beta == 0 (generator :229,287); every other row starts from C
 * (beta == 1) or 0 and takes one fused multiply-add per nonzero in CSR order with the
 * table value (generator :260-275).  branch 0 = dense SMM fallback
 * (src/libxsmm_fsspmdm.c:133-143,270-272): every row, every k in ascending order, zeros
 * included, fused multiply-adds starting from C or 0.
 * B is K x N with pitch ldb, C is M x N with pitch ldc.
 */
void orc_dfsspmdm_execute(int M, int N, int K, int lda, int ldb, int ldc, double beta,
                          const double* a, const double* B, double* C, int branch)
{
  int m, n, k;
  if (branch) {
    uint32_t* rowptr = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t* colidx = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M * K + 1));
    double* val = (double*)malloc(sizeof(double) * ((size_t)M * K + 1));
    int nu; uint32_t u;
    orc_dfsspmdm_plan(M, K, lda, a, rowptr, colidx, val, &nu);
    for (m = 0; m < M; ++m) {
      if (rowptr[m + 1] == rowptr[m]) continue;
      for (n = 0; n < N; ++n) {
        double acc = (1.0 == beta) ? C[(size_t)m * ldc + n] : 0.0;
        for (u = rowptr[m]; u < rowptr[m + 1]; ++u) acc = fma(val[u], B[(size_t)colidx[u] * ldb + n], acc);
        C[(size_t)m * ldc + n] = acc;
      }
    }
    free(rowptr); free(colidx); free(val);
  }
  else {
    for (m = 0; m < M; ++m) for (n = 0; n < N; ++n) {
      double acc = (1.0 == beta) ? C[(size_t)m * ldc + n] : 0.0;
      for (k = 0; k < K; ++k) acc = fma(a[(size_t)m * lda + k], B[(size_t)k * ldb + n], acc);
      C[(size_t)m * ldc + n] = acc;
    }
  }
}

/* float operator: the reference never obtains a sparse kernel for float
 * (src/libxsmm_main.c:1418 admits F64 only), so this is always the dense branch. */
void orc_sfsspmdm_execute(int M, int N, int K, int lda, int ldb, int ldc, float beta,
                          const float* a, const float* B, float* C)
{
  int m, n, k;
  for (m = 0; m < M; ++m) for (n = 0; n < N; ++n) {
    float acc = (1.f == beta) ? C[(size_t)m * ldc + n] : 0.f;
    for (k = 0; k < K; ++k) acc = fmaf(a[(size_t)m * lda + k], B[(size_t)k * ldb + n], acc);
    C[(size_t)m * ldc + n] = acc;
  }
}

/*
 * Which branch the reference takes for a double operator on an AVX-512 host
 * (the JIT follows the runtime CPUID target; on a host without AVX-512 it is always dense):
 * sparse iff nnz > 0, at most 31 table entries, and the emitted code fits 128 KiB.
 */
int orc_dfsspmdm_branch(int M, int K, int lda, const double* a, int ldb, int ldc, double beta, int host_avx512)
{
  uint32_t* rowptr = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
  uint32_t* colidx = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M * K + 1));
  double* val = (double*)malloc(sizeof(double) * ((size_t)M * K + 1));
  int nu = 0, branch = 0;
  const int nnz = orc_dfsspmdm_plan(M, K, lda, a, rowptr, colidx, val, &nu);
  if (host_avx512 && nnz > 0 && nu <= 31) {
    branch = (orc_dfsspmdm_code_size(M, rowptr, colidx, nu, ldb, ldc, 1.0 == beta) <= 131072);
  }
  free(rowptr); free(colidx); free(val);
  return branch;
}

/*
 * CSR "A sparse" x dense SoA kernel: what the code emitted by
 * src/generator_spgemm_csr_asparse_soa.c:213-420 computes for ONE element:
 *     C[m][n][s] = (beta == 0 ? 0 : C[m][n][s]) + sum_z a[z] * B[col[z]][n][s],   z over row m's nonzeros in CSR order,
 * B laid out [k][ldb][soa], C [m][ldc][soa].  One fused multiply-add per nonzero, in order, starting from the loaded C
 * (any beta != 0 loads C: the generator only tests the BETA_0 flag, :253-267) or from zero.  A row without nonzeros is
 * skipped entirely (:249: "if (l_row_elements > 0)"): its C entries are left untouched even for beta = 0.  The operator
 * VALUES are read at run time from the kernel's first argument (:277-285); the N chunking (:176-180) does not change
 * the arithmetic.  dbl: 1 = double, 0 = float.  n_elem elements, strides in scalars.
 */
void orc_csr_soa_execute(int dbl, int M, int N, int K, int ldb, int ldc, int soa, double beta,
                         const uint32_t* rowptr, const uint32_t* colidx, const void* values,
                         const void* B, void* C, long n_elem, long stride_b, long stride_c)
{
  long e; int m, n, s; uint32_t z;
  (void)K;
  for (e = 0; e < n_elem; ++e) for (m = 0; m < M; ++m) {
    if (rowptr[m + 1] == rowptr[m]) continue;
    for (n = 0; n < N; ++n) for (s = 0; s < soa; ++s) {
      const size_t cat = (size_t)e * stride_c + ((size_t)m * ldc + n) * soa + s;
      if (dbl) {
        double acc = (0.0 == beta) ? 0.0 : ((double*)C)[cat];
        for (z = rowptr[m]; z < rowptr[m + 1]; ++z)
          acc = fma(((const double*)values)[z], ((const double*)B)[(size_t)e * stride_b + ((size_t)colidx[z] * ldb + n) * soa + s], acc);
        ((double*)C)[cat] = acc;
      }
      else {
        float acc = (0.0 == beta) ? 0.f : ((float*)C)[cat];
        for (z = rowptr[m]; z < rowptr[m + 1]; ++z)
          acc = fmaf(((const float*)values)[z], ((const float*)B)[(size_t)e * stride_b + ((size_t)colidx[z] * ldb + n) * soa + s], acc);
        ((float*)C)[cat] = acc;
      }
    }
  }
}

/*
 * The same entry with B sparse (descriptor lda > 0, ldb == 0; src/generator_spgemm_csr_bsparse_soa.c:60-250; caller
 * samples/edge/bsparse_srsoa.c): A dense [m][lda][soa], B in CSR over its K rows, C [m][ldc][soa],
 *     C[m][n][s] = (beta == 0 ? 0 : C[m][n][s]) + sum over k ascending, over the nonzeros (k, n) of B's row k, of A[m][k][s] * b.
 * The emitted code works on columns 0 .. ncols-1 only, ncols = 1 + the LARGEST column index any nonzero of B holds
 * (:161-167): it loads (or zeroes) all ncols accumulators chunk by chunk, walks k = 0 .. K-1 with one fused multiply-add
 * per nonzero of row k, and stores all ncols accumulators (:187-205, :208-292, :295-305).  So a column WITHOUT nonzeros
 * is written (zero for beta = 0) when a later column has one, and columns >= ncols are never touched, not even for
 * beta = 0 (PyFR's tet4_5_stiffT_1 has 56 columns, the last 21 empty: C keeps its old content there).  Nonzeros with
 * column >= N take no part in the multiply (:213,233) but still count towards ncols.
 */
void orc_csr_soa_bsparse_execute(int dbl, int M, int N, int K, int lda, int ldc, int soa, double beta,
                                 const uint32_t* rowptr, const uint32_t* colidx, const void* values,
                                 const void* A, void* C, long n_elem, long stride_a, long stride_c)
{
  long e; int m, n, s, k, ncols = 0; uint32_t z;
  for (z = 0; z < rowptr[K]; ++z) if ((int)colidx[z] >= ncols) ncols = (int)colidx[z] + 1;
  for (e = 0; e < n_elem; ++e) for (m = 0; m < M; ++m) for (n = 0; n < ncols; ++n) for (s = 0; s < soa; ++s) {
    const size_t cat = (size_t)e * stride_c + ((size_t)m * ldc + n) * soa + s;
    if (dbl) {
      double acc = (0.0 == beta) ? 0.0 : ((double*)C)[cat];
      for (k = 0; k < K; ++k) for (z = rowptr[k]; z < rowptr[k + 1]; ++z) if ((int)colidx[z] == n && n < N)
        acc = fma(((const double*)A)[(size_t)e * stride_a + ((size_t)m * lda + k) * soa + s], ((const double*)values)[z], acc);
      ((double*)C)[cat] = acc;
    }
    else {
      float acc = (0.0 == beta) ? 0.f : ((float*)C)[cat];
      for (k = 0; k < K; ++k) for (z = rowptr[k]; z < rowptr[k + 1]; ++z) if ((int)colidx[z] == n && n < N)
        acc = fmaf(((const float*)A)[(size_t)e * stride_a + ((size_t)m * lda + k) * soa + s], ((const float*)values)[z], acc);
      ((float*)C)[cat] = acc;
    }
  }
}

/*
 * libxsmm_create_xcsc_soa (src/libxsmm_main.c:2450-2474; src/generator_spgemm_csc_bsparse_soa.c:143-435; caller
 * samples/edge/bsparse_scsoa.c:327-354): B sparse in CSC (colptr over its N columns, rowidx = k), A dense [m][lda][soa],
 * C [m][ldc][soa].  Per chunk of columns the emitted code loads (or zeroes) the accumulators (:206-224), walks k = 0 .. K-1
 * and for every column of the chunk searches the column for the FIRST entry whose row index is k (:270-279; the KNM
 * "qmadd" variant :231-268 is not what any host here generates), one fused multiply-add each (:345-400), then stores
 * (:406-415).  The "max column" loop (:185-190) tests colptr[n'] == colptr[N] for every n' WITHOUT stopping, so its last
 * assignment is always N: all N columns are written, also trailing empty ones.
 */
void orc_csc_soa_execute(int dbl, int M, int N, int K, int lda, int ldc, int soa, double beta,
                         const uint32_t* colptr, const uint32_t* rowidx, const void* values,
                         const void* A, void* C, long n_elem, long stride_a, long stride_c)
{
  long e; int m, n, s, k; uint32_t z;
  for (e = 0; e < n_elem; ++e) for (m = 0; m < M; ++m) for (n = 0; n < N; ++n) for (s = 0; s < soa; ++s) {
    const size_t cat = (size_t)e * stride_c + ((size_t)m * ldc + n) * soa + s;
    if (dbl) {
      double acc = (0.0 == beta) ? 0.0 : ((double*)C)[cat];
      for (k = 0; k < K; ++k) for (z = colptr[n]; z < colptr[n + 1]; ++z) if ((int)rowidx[z] == k) {
        acc = fma(((const double*)A)[(size_t)e * stride_a + ((size_t)m * lda + k) * soa + s], ((const double*)values)[z], acc);
        break;
      }
      ((double*)C)[cat] = acc;
    }
    else {
      float acc = (0.0 == beta) ? 0.f : ((float*)C)[cat];
      for (k = 0; k < K; ++k) for (z = colptr[n]; z < colptr[n + 1]; ++z) if ((int)rowidx[z] == k) {
        acc = fmaf(((const float*)A)[(size_t)e * stride_a + ((size_t)m * lda + k) * soa + s], ((const float*)values)[z], acc);
        break;
      }
      ((float*)C)[cat] = acc;
    }
  }
}
