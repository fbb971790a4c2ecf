/* GSL-compat shim — TEST INFRASTRUCTURE ONLY (used to compile the unmodified reference into
 * oracle/_ref/; never linked into the product library).
 *
 * pmclib depends on an un-vendored GSL (>= 1.14, wscript:131-133).  GSL is absent from this image, so this
 * header declares just the GSL surface the hot-path files reference, with
 * layout-compatible public structs, and gsl_shim.c restates the published algorithms of gsl-1.14
 * (row-dot Cholesky that mirrors L^T into the upper triangle, reference-CBLAS dtrsv/dtrmv, mt19937,
 * polar Box-Muller gaussian, Marsaglia-Tsang gamma).
 */
#ifndef PMCB200_GSL_SHIM_ALL_H
#define PMCB200_GSL_SHIM_ALL_H
#include <stddef.h>
#include <errno.h>
#include <stdlib.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_EDOM = 1, GSL_ERANGE = 2, GSL_EINVAL = 4, GSL_ENOMEM = 8, GSL_EBADLEN = 19, GSL_ENOTSQR = 20 };

typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler_off(void);

typedef struct { size_t size; double *data; } gsl_block;
typedef struct { size_t size; size_t stride; double *data; gsl_block *block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } _gsl_vector_view;
typedef _gsl_vector_view gsl_vector_view;
typedef struct { size_t size; int *data; } gsl_block_int;
typedef struct { size_t size; size_t stride; int *data; gsl_block_int *block; int owner; } gsl_vector_int;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; gsl_block *block; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } _gsl_matrix_view;
typedef _gsl_matrix_view gsl_matrix_view;

gsl_vector_view gsl_vector_view_array(double *base, size_t n);
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2);
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b);
int gsl_vector_scale(gsl_vector *a, const double x);
void gsl_vector_set_zero(gsl_vector *v);
void gsl_matrix_set_zero(gsl_matrix *m);
int gsl_matrix_scale(gsl_matrix *a, const double x);

int gsl_linalg_cholesky_decomp(gsl_matrix *A);
int gsl_linalg_cholesky_invert(gsl_matrix *cholesky);

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER_t;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO_t;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG_t;
int gsl_blas_dtrsv(CBLAS_UPLO_t Uplo, CBLAS_TRANSPOSE_t TransA, CBLAS_DIAG_t Diag, const gsl_matrix *A, gsl_vector *X);
int gsl_blas_dtrmv(CBLAS_UPLO_t Uplo, CBLAS_TRANSPOSE_t TransA, CBLAS_DIAG_t Diag, const gsl_matrix *A, gsl_vector *X);

double gsl_sf_lngamma(double x);

typedef struct {
  const char *name;
  unsigned long int max;
  unsigned long int min;
  size_t size;
  void (*set)(void *state, unsigned long int seed);
  unsigned long int (*get)(void *state);
  double (*get_double)(void *state);
} gsl_rng_type;
typedef struct { const gsl_rng_type *type; void *state; } gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_default;
extern unsigned long int gsl_rng_default_seed;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
void gsl_rng_free(gsl_rng *r);
unsigned long int gsl_rng_get(const gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
double gsl_rng_uniform_pos(const gsl_rng *r);
const gsl_rng_type **gsl_rng_types_setup(void);

double gsl_ran_gaussian(const gsl_rng *r, const double sigma);
double gsl_ran_flat(const gsl_rng *r, const double a, const double b);
double gsl_ran_gamma(const gsl_rng *r, const double a, const double b);
double gsl_ran_chisq(const gsl_rng *r, const double nu);
typedef struct { size_t K; double *cum; } gsl_ran_discrete_t;
gsl_ran_discrete_t *gsl_ran_discrete_preproc(size_t K, const double *P);
void gsl_ran_discrete_free(gsl_ran_discrete_t *g);
size_t gsl_ran_discrete(const gsl_rng *r, const gsl_ran_discrete_t *g);
void gsl_ran_multinomial(const gsl_rng *r, const size_t K, const unsigned int N, const double p[], unsigned int n[]);

#ifdef __cplusplus
}
#endif
#endif
