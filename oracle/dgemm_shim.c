/* TEST INFRASTRUCTURE ONLY.  A plain column-major dgemm ('N','N' only) so that the UNMODIFIED reference driver
 * samples/pyfr/pyfr_driver_asp_reg.c -- which calls MKL's dgemm as a comparator (:314-316; MKL is not in this image) --
 * links.  Own code, not part of the product, not timed. */
void dgemm(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha,
           const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc)
{
  int i, j, l;
  (void)transa; (void)transb;
  for (j = 0; j < *n; ++j) {
    for (i = 0; i < *m; ++i) c[(long)j * *ldc + i] = (0.0 == *beta) ? 0.0 : *beta * c[(long)j * *ldc + i];
    for (l = 0; l < *k; ++l) {
      const double t = *alpha * b[(long)j * *ldb + l];
      for (i = 0; i < *m; ++i) c[(long)j * *ldc + i] += t * a[(long)l * *lda + i];
    }
  }
}
