/* GSL-compat shim implementation — TEST INFRASTRUCTURE ONLY (see gsl/gsl_shim_all.h).
 *
 * Restates, from their published descriptions, the gsl-1.14 algorithms behind the call sites the
 * reference's hot path uses:
 *   gsl_linalg_cholesky_decomp  mvdens.c:255   row-by-row (Cholesky-Banachiewicz), diagonal through a
 *                                              scaled 2-norm, L^T mirrored into the upper triangle,
 *                                              GSL_EDOM when a pivot is <= 0
 *   gsl_blas_dtrsv / dtrmv      mvdens.c:341,302  reference CBLAS, row-major, lower, no-trans, non-unit
 *   gsl_sf_lngamma              mvdens.c:376-377  (libm lgamma; agrees to rounding for the arguments used)
 *   gsl_ran_gaussian            mvdens.c:295   polar Box-Muller on uniform_pos
 *   gsl_ran_flat                mvdens.c:775   a(1-u)+bu
 *   gsl_ran_chisq               mvdens.c:307   2*Gamma(nu/2,1), Marsaglia-Tsang
 *   gsl_rng mt19937             pmc_rc.c:734-763
 * Nothing here is linked into libpmcb200.so.
 */
#include "gsl/gsl_shim_all.h"
#include <string.h>
#include <stdio.h>

gsl_error_handler_t *gsl_set_error_handler_off(void) { return NULL; }

/* ---- views and level-1 helpers ------------------------------------------------------------- */
gsl_vector_view gsl_vector_view_array(double *base, size_t n) {
  gsl_vector_view v;
  v.vector.size = n; v.vector.stride = 1; v.vector.data = base; v.vector.block = NULL; v.vector.owner = 0;
  return v;
}
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2) {
  gsl_matrix_view m;
  m.matrix.size1 = n1; m.matrix.size2 = n2; m.matrix.tda = n2; m.matrix.data = base;
  m.matrix.block = NULL; m.matrix.owner = 0;
  return m;
}
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b) {
  if (a->size != b->size) return GSL_EBADLEN;
  for (size_t i = 0; i < a->size; ++i) a->data[i * a->stride] -= b->data[i * b->stride];
  return GSL_SUCCESS;
}
int gsl_vector_scale(gsl_vector *a, const double x) {
  for (size_t i = 0; i < a->size; ++i) a->data[i * a->stride] *= x;
  return GSL_SUCCESS;
}
void gsl_vector_set_zero(gsl_vector *v) {
  for (size_t i = 0; i < v->size; ++i) v->data[i * v->stride] = 0.0;
}
void gsl_matrix_set_zero(gsl_matrix *m) {
  for (size_t i = 0; i < m->size1; ++i)
    for (size_t j = 0; j < m->size2; ++j) m->data[i * m->tda + j] = 0.0;
}
int gsl_matrix_scale(gsl_matrix *a, const double x) {
  for (size_t i = 0; i < a->size1; ++i)
    for (size_t j = 0; j < a->size2; ++j) a->data[i * a->tda + j] *= x;
  return GSL_SUCCESS;
}

/* ---- Cholesky (gsl-1.14 linalg/cholesky.c algorithm) ---------------------------------------- */
static double shim_nrm2(const double *x, size_t n) {
  /* reference-BLAS dnrm2: scaled sum of squares */
  double scale = 0.0, ssq = 1.0;
  if (n == 0) return 0.0;
  if (n == 1) return fabs(x[0]);
  for (size_t i = 0; i < n; ++i) {
    const double v = x[i];
    if (v != 0.0) {
      const double ax = fabs(v);
      if (scale < ax) { ssq = 1.0 + ssq * (scale / ax) * (scale / ax); scale = ax; }
      else            { ssq += (ax / scale) * (ax / scale); }
    }
  }
  return scale * sqrt(ssq);
}
static double shim_quiet_sqrt(double x) { return (x >= 0) ? sqrt(x) : (0.0 / 0.0); }

int gsl_linalg_cholesky_decomp(gsl_matrix *A) {
  const size_t M = A->size1, T = A->tda;
  double *a = A->data;
  int status = 0;
  if (M != A->size2) return GSL_ENOTSQR;
  {
    const double a00 = a[0];
    a[0] = shim_quiet_sqrt(a00);
    if (a00 <= 0) status = GSL_EDOM;
  }
  if (M > 1) {
    const double l10 = a[T] / a[0];
    const double diag = a[T + 1] - l10 * l10;
    a[T] = l10;
    a[T + 1] = shim_quiet_sqrt(diag);
    if (diag <= 0) status = GSL_EDOM;
  }
  for (size_t k = 2; k < M; ++k) {
    const double akk = a[k * T + k];
    for (size_t i = 0; i < k; ++i) {
      double sum = 0.0;
      if (i > 0) for (size_t j = 0; j < i; ++j) sum += a[i * T + j] * a[k * T + j];
      a[k * T + i] = (a[k * T + i] - sum) / a[i * T + i];
    }
    {
      const double nrm = shim_nrm2(a + k * T, k);
      const double diag = akk - nrm * nrm;
      a[k * T + k] = shim_quiet_sqrt(diag);
      if (diag <= 0) status = GSL_EDOM;
    }
  }
  for (size_t i = 1; i < M; ++i)
    for (size_t j = 0; j < i; ++j) a[j * T + i] = a[i * T + j];
  return status == GSL_EDOM ? GSL_EDOM : GSL_SUCCESS;
}

int gsl_linalg_cholesky_invert(gsl_matrix *C) {
  /* A^-1 = L^-T L^-1 from the factor stored by gsl_linalg_cholesky_decomp */
  const size_t n = C->size1, T = C->tda;
  double *a = C->data;
  double *li = (double *)calloc(n * n, sizeof(double));
  if (!li) return GSL_ENOMEM;
  for (size_t c = 0; c < n; ++c) {          /* columns of L^-1 by forward substitution */
    for (size_t i = c; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (size_t j = c; j < i; ++j) s -= a[i * T + j] * li[j * n + c];
      li[i * n + c] = s / a[i * T + i];
    }
  }
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j <= i; ++j) {
      double s = 0.0;
      for (size_t k = i; k < n; ++k) s += li[k * n + i] * li[k * n + j];
      a[i * T + j] = s; a[j * T + i] = s;
    }
  free(li);
  return GSL_SUCCESS;
}

/* ---- level-2 BLAS, the one case the reference uses ------------------------------------------- */
int gsl_blas_dtrsv(CBLAS_UPLO_t Uplo, CBLAS_TRANSPOSE_t TransA, CBLAS_DIAG_t Diag, const gsl_matrix *A, gsl_vector *X) {
  const size_t N = A->size1, lda = A->tda, inc = X->stride;
  double *x = X->data;
  if (Uplo != CblasLower || TransA != CblasNoTrans) { fprintf(stderr, "gsl shim: dtrsv case not provided\n"); abort(); }
  if (N == 0) return GSL_SUCCESS;
  if (Diag == CblasNonUnit) x[0] = x[0] / A->data[0];
  for (size_t i = 1; i < N; ++i) {
    double tmp = x[i * inc];
    for (size_t j = 0; j < i; ++j) tmp -= A->data[lda * i + j] * x[j * inc];
    x[i * inc] = (Diag == CblasNonUnit) ? tmp / A->data[lda * i + i] : tmp;
  }
  return GSL_SUCCESS;
}
int gsl_blas_dtrmv(CBLAS_UPLO_t Uplo, CBLAS_TRANSPOSE_t TransA, CBLAS_DIAG_t Diag, const gsl_matrix *A, gsl_vector *X) {
  const size_t N = A->size1, lda = A->tda, inc = X->stride;
  double *x = X->data;
  if (Uplo != CblasLower || TransA != CblasNoTrans) { fprintf(stderr, "gsl shim: dtrmv case not provided\n"); abort(); }
  for (size_t i = N; i > 0 && i--;) {
    double temp = 0.0;
    for (size_t j = 0; j < i; ++j) temp += x[j * inc] * A->data[lda * i + j];
    if (Diag == CblasNonUnit) x[i * inc] = temp + x[i * inc] * A->data[lda * i + i];
    else x[i * inc] += temp;
  }
  return GSL_SUCCESS;
}

double gsl_sf_lngamma(double x) { return lgamma(x); }

/* ---- mt19937 ---------------------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397
typedef struct { unsigned long mt[MT_N]; int mti; } mt_state_t;
static void mt_set(void *vstate, unsigned long int s) {
  mt_state_t *st = (mt_state_t *)vstate;
  if (s == 0) s = 4357;
  st->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < MT_N; ++i)
    st->mt[i] = (1812433253UL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  st->mti = MT_N;
}
static unsigned long int mt_get(void *vstate) {
  mt_state_t *st = (mt_state_t *)vstate;
  unsigned long *const mt = st->mt;
  unsigned long k;
  if (st->mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; ++kk) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    for (; kk < MT_N - 1; ++kk) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    {
      unsigned long y = (mt[MT_N - 1] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
      mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    st->mti = 0;
  }
  k = mt[st->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  return k & 0xffffffffUL;
}
static double mt_get_double(void *vstate) { return mt_get(vstate) / 4294967296.0; }
static const gsl_rng_type mt_type = {"mt19937", 0xffffffffUL, 0, sizeof(mt_state_t), &mt_set, &mt_get, &mt_get_double};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;
const gsl_rng_type *gsl_rng_default = &mt_type;
unsigned long int gsl_rng_default_seed = 0;
static const gsl_rng_type *shim_types[2] = {&mt_type, NULL};
const gsl_rng_type **gsl_rng_types_setup(void) { return shim_types; }

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  if (!r) return NULL;
  r->state = calloc(1, T->size);
  if (!r->state) { free(r); return NULL; }
  r->type = T;
  gsl_rng_set(r, gsl_rng_default_seed);
  return r;
}
void gsl_rng_set(const gsl_rng *r, unsigned long int seed) { r->type->set(r->state, seed); }
void gsl_rng_free(gsl_rng *r) { if (r) { free(r->state); free(r); } }
unsigned long int gsl_rng_get(const gsl_rng *r) { return r->type->get(r->state); }
double gsl_rng_uniform(const gsl_rng *r) { return r->type->get_double(r->state); }
double gsl_rng_uniform_pos(const gsl_rng *r) {
  double x;
  do { x = r->type->get_double(r->state); } while (x == 0);
  return x;
}

/* ---- variates --------------------------------------------------------------------------------- */
double gsl_ran_gaussian(const gsl_rng *r, const double sigma) {
  double x, y, r2;
  do {
    x = -1 + 2 * gsl_rng_uniform_pos(r);
    y = -1 + 2 * gsl_rng_uniform_pos(r);
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);
  return sigma * y * sqrt(-2.0 * log(r2) / r2);
}
double gsl_ran_flat(const gsl_rng *r, const double a, const double b) {
  const double u = gsl_rng_uniform(r);
  return a * (1 - u) + b * u;
}
double gsl_ran_gamma(const gsl_rng *r, const double a, const double b) {
  if (a < 1) {
    const double u = gsl_rng_uniform_pos(r);
    return gsl_ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
  }
  {
    double x, v, u;
    const double d = a - 1.0 / 3.0;
    const double c = (1.0 / 3.0) / sqrt(d);
    for (;;) {
      do { x = gsl_ran_gaussian(r, 1.0); v = 1.0 + c * x; } while (v <= 0);
      v = v * v * v;
      u = gsl_rng_uniform_pos(r);
      if (u < 1 - 0.0331 * x * x * x * x) break;
      if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
    }
    return b * d * v;
  }
}
double gsl_ran_chisq(const gsl_rng *r, const double nu) { return 2 * gsl_ran_gamma(r, nu / 2, 1.0); }

gsl_ran_discrete_t *gsl_ran_discrete_preproc(size_t K, const double *P) {
  gsl_ran_discrete_t *g = (gsl_ran_discrete_t *)malloc(sizeof(*g));
  double tot = 0, c = 0;
  g->K = K; g->cum = (double *)malloc(sizeof(double) * K);
  for (size_t k = 0; k < K; ++k) tot += P[k];
  for (size_t k = 0; k < K; ++k) { c += P[k] / tot; g->cum[k] = c; }
  return g;
}
void gsl_ran_discrete_free(gsl_ran_discrete_t *g) { if (g) { free(g->cum); free(g); } }
size_t gsl_ran_discrete(const gsl_rng *r, const gsl_ran_discrete_t *g) {
  const double u = gsl_rng_uniform(r);
  size_t lo = 0, hi = g->K - 1;
  while (lo < hi) { size_t mid = (lo + hi) / 2; if (g->cum[mid] < u) lo = mid + 1; else hi = mid; }
  return lo;
}
void gsl_ran_multinomial(const gsl_rng *r, const size_t K, const unsigned int N, const double p[], unsigned int n[]) {
  gsl_ran_discrete_t *g = gsl_ran_discrete_preproc(K, p);
  memset(n, 0, sizeof(unsigned int) * K);
  for (unsigned int t = 0; t < N; ++t) n[gsl_ran_discrete(r, g)]++;
  gsl_ran_discrete_free(g);
}
