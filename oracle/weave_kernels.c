/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's two inline scipy.weave kernels
 * (the only native code in GPclust).  Used by tests/ and by bench.py's CPU legs; never by the product.
 *
 *   gpo_softmax              restates GPclust/utilities.py:61-102 (`softmax_weave`)
 *   gpo_multiple_mahalanobis restates GPclust/utilities.py:147-170 (`multiple_mahalanobis`)
 *
 * Same loop order and the same OpenMP row-parallel pragma the reference uses when GPy's config has
 * [parallel] openmp = True (utilities.py:54,143).  One deviation, stated here: the reference's
 * `double entropy_i[N]` is a stack VLA (utilities.py:64) that overflows the stack at large N; it is
 * heap-allocated by the caller here.  `tmp` (utilities.py:148) is private per iteration, as the
 * reference's `private(...tmp)` clause intends.  The reference's weave sources are Python strings with
 * blitz converters and cannot be compiled from where they lie, hence this "port".
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC (oracle/build.py).
 */
#include <math.h>
#include <stdlib.h>

/* X: N x D row-major.  phi, log_phi: N x D outputs.  entropy_i: N scratch.  Returns H. */
double gpo_softmax(const double *X, double *phi, double *log_phi, double *entropy_i, long N, long D) {
  long i, j;
  double max_x, norm, log_norm, entropy;
#pragma omp parallel for private(i, j, max_x, norm, log_norm)
  for (i = 0; i < N; i++) {
    const double *x = X + i * D;
    double *p = phi + i * D, *lp = log_phi + i * D;
    max_x = x[0];
    for (j = 1; j < D; j++) {
      if (x[j] > max_x) max_x = x[j];
    }
    norm = 0.0;
    for (j = 0; j < D; j++) {
      lp[j] = x[j] - max_x;
      p[j] = exp(lp[j]);
      norm += p[j];
    }
    log_norm = log(norm);
    entropy_i[i] = 0.0;
    for (j = 0; j < D; j++) {
      p[j] /= norm;
      lp[j] -= log_norm;
      entropy_i[i] -= p[j] * lp[j];
    }
  }
  entropy = 0.0;
  for (i = 0; i < N; i++) entropy += entropy_i[i];
  return entropy;
}

/* X1: N1 x D, X2: N2 x D, L: D x D lower, result: N1 x N2 (overwritten). */
void gpo_multiple_mahalanobis(const double *X1, const double *X2, const double *L, long N1, long N2, long D, double *result) {
  long n;
#pragma omp parallel
  {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)D);
    long m, i, j;
#pragma omp for
    for (n = 0; n < N1; n++) {
      for (m = 0; m < N2; m++) {
        double r = 0.0;
        for (i = 0; i < D; i++) {
          tmp[i] = X1[n * D + i] - X2[m * D + i];
          for (j = 0; j < i; j++) tmp[i] -= L[i * D + j] * tmp[j];
          tmp[i] /= L[i * D + i];
        }
        for (i = 0; i < D; i++) r += tmp[i] * tmp[i];
        result[n * N2 + m] = r;
      }
    }
    free(tmp);
  }
}
