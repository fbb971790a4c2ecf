/* pmc_oracle.h — CPU restatement of pmclib's PMC iteration.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.  It is the checker for the CUDA path, never the thing measured or shipped; libpmcb200.so
 * does not link, load or call it.
 *
 * Parity status: PINNED against the reference itself.  The reference has no tests or golden vectors for this
 * path, so the pin is the unmodified reference compiled here (oracle/_ref/libpmcref.so, built by
 * oracle/Makefile from /root/reference + the GSL shim); tests/test_oracle_vs_ref.py checks every function below
 * against it, and the .npz fixtures under tests/golden/ (made by tests/golden/make_golden.py from _ref) carry the same pins to
 * machines where /root/reference does not exist.
 *
 * Everything is plain C on flat arrays.  A mixture is (K, d, wght[K], mean[K*d], L[K*d*d], detL[K], df[K]) with
 * L the row-major lower Cholesky factor exactly as the reference stores it after mvdens_cholesky_decomp
 * (upper triangle = mirror of L^T; the restatement never reads the upper part).
 */
#ifndef PMC_ORACLE_H
#define PMC_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference constants: pmc.h:45-47,71 ; mvdens.h:70-81 ; maths_base.h:49 (ln2pi is TRUNCATED there, on purpose kept) */
#define ORC_PI 3.141592653589793
#define ORC_LN2PI 1.837877066409
#define ORC_DYNMAX 500.0
#define ORC_MINCOUNT 20
#define ORC_PERP_ZERO 1e-20

/* error codes = the reference's (pmc.h:23-43, mvdens.h:54-68) */
#define ORC_OK 0
#define ORC_MV_CHOLESKY (-706)
#define ORC_MV_NEGWEIGHT (-705)
#define ORC_MV_OUTOFBOUND (-703)
#define ORC_PMC_NEGWEIGHT (-6005)
#define ORC_PMC_UNDEF (-6008)
#define ORC_PMC_TOOMANYSTEPS (-6011)
#define ORC_PMC_NOSAMPLEP (-6017)
#define ORC_PMC_INFINITE (-6019)

typedef struct {
  int K, d;
  double *wght;        /* [K] */
  double *mean;        /* [K*d] */
  double *L;           /* [K*d*d] */
  double *detL;        /* [K] */
  int *df;             /* [K], -1 = Gaussian */
  size_t *band_limit;  /* [K] or NULL (= d) */
} orc_mix;

enum { ORC_TGT_BANANA = 1, ORC_TGT_MVDENS = 2, ORC_TGT_MIX = 3, ORC_TGT_COMBINE = 5 };
typedef struct orc_target_s {
  int kind, d;
  double b, sig1;  /* bentGauss */
  orc_mix mix;     /* kinds 2 (K=1, wght ignored) and 3 */
  /* kind 5, combine_lkl (distribution.c:350-396): the sum of nterms densities, term t on the coordinates
   * dims[t][0..terms[t].d) of x; this is what add_gaussian_prior (distribution.c:548-606) builds */
  int nterms;
  const struct orc_target_s *terms;
  const int *const *dims;
} orc_target;

/* mvdens.c:229-277 + gsl_linalg_cholesky_decomp: in place, restores a[] and returns ORC_MV_CHOLESKY if not PD */
int orc_cholesky(int d, double *a, double *detL);
/* mvdens.c:319-349 */
double orc_scalar_product(int d, const double *mean, const double *L, const double *x, double *tmp);
/* mvdens.c:354-385 */
double orc_mvdens_log_pdf(int d, const double *mean, const double *L, double detL, int df, const double *x, double *tmp);
/* mvdens.c:784-825; tmp[d], ld[K] */
double orc_mix_log_pdf(const orc_mix *m, const double *x, double *tmp, double *ld, int *rc);
/* mvdens.c:736-751: renormalises wght in place when |sum-1|>1e-6, fills cw */
int orc_cwght(int K, double *wght, double *cw);
/* mvdens.c:766-782 with the uniform injected */
size_t orc_ranindx(const double *cw, size_t K, double z);
/* mvdens.c:279-317 with the d normals g[] (and the chi^2 variate u when df!=-1) injected */
void orc_mvdens_ran_from(int d, const double *mean, const double *L, int df, const double *g, double u, double *x);
/* parabox.c:112-136; faces[nfaces*(d+1)] */
int orc_isinbox(int d, int nfaces, const double *faces, const double *x);
/* examples/target/sampleTarget.c:89-110 */
double orc_bentgauss(int d, double b, double sig1, const double *x);
double orc_target_log_pdf(const orc_target *t, const double *x, double *tmp, double *ld, int *rc);

/* pmc.c:484-600 (t = t_iter_tempered); returns the number of ok samples, <0 on error */
long orc_weights(long N, int d, const double *X, const orc_mix *prop, const orc_target *tgt, double t,
                 double *weights, double *log_rho, short *flg, double *maxW, double *maxR);
/* pmc.c:641-697 */
int orc_normalize(long N, double *weights, short *flg, double maxW, double *logSum);
/* pmc.c:1041-1077 (MC_UNORM) */
int orc_perplexity_ess(long N, const double *weights, const short *flg, double *perp, double *ess);
/* pmc.c:1088-1105 */
int orc_evidence(long N, const short *flg, double logSum, double *ln_evi);
/* pmc.c:1323-1469 + prepare_update :1535-1579 + cleanup_after_update :1471-1533.
 * m is updated in place (L then holds the NEW Cholesky factors of live components, zeros for dead ones);
 * chol_ok[k] (may be NULL) = 1 where a factor was produced.  *retry is psim->retry (in/out). */
int orc_update_rb(long N, const double *X, const size_t *indices, const double *weights, const double *log_rho,
                  const short *flg, double maxR, orc_mix *m, int *retry);
/* pmc.c:1194-1310 (hard-assignment EM) */
int orc_update(long N, const double *X, const size_t *indices, const double *weights, const short *flg,
               orc_mix *m, int *retry);
/* pmc.c:1471-1533 on its own: m->L holds COVARIANCES of the updated components on entry */
int orc_cleanup_after_update(orc_mix *m, const size_t *count, double minwgt, int *retry, const double *snapshot);

/* pmc.c:266-298 with a private splitmix64/polar-method stream (statistically equivalent to the reference's
 * gsl_rng stream; used as the CPU baseline "port" and for statistical checks).  Returns N or <0. */
long orc_simulate(long N, orc_mix *m, int nfaces, const double *faces, unsigned long long seed,
                  double *X, size_t *indices, short *flg, double *cw_scratch);

/* filter_histogram (pmc.c:709-770), mean_from_psim (pmc.c:1108-1120) */
int orc_clip_weights(long N, double *weights, short *flg, int n, double *thr); /* pmc.c:1134-1188 */
long orc_filter_histogram(long N, const double *weights, short *flg, int nbins, int cl, int cr);
double orc_mean_from_psim(const double *X, const double *w, const short *flg, long N, int d, int a);

#ifdef __cplusplus
}
#endif
#endif
