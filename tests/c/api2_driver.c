/* api2_driver.c — the rest of pmclib's API around the PMC path, called the way a pmclib user calls it (names and
 * signatures of mvdens.h:113-155, pmc.h:192-199, distribution.h:66-119, errorlist.h:48-49, parabox.h), linked against
 * libpmcb200.so.
 *
 *   api2_driver wire  <out.bin>                      serialize / deserialize / copy / size / small helpers -> text report
 *   api2_driver ran   <out.bin> <M> <nu>             M draws of mix_mvdens_ran (3-component Gaussian mixture, d = 4) and M
 *                                                    draws of mvdens_ran (Student-t, nu degrees of freedom) -> doubles
 *   api2_driver file  <in.lc> <out.bin> <N> <d>      pmc_simu_from_file with a 2-component proposal -> arrays
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pmclib_abi.h"

typedef struct { uint64_t s; } sm_state;
static unsigned long sm_get(void *vs) {
  sm_state *st = (sm_state *)vs;
  uint64_t z = (st->s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return (unsigned long)((z ^ (z >> 31)) & 0xffffffffUL);
}
static double sm_get_double(void *vs) { return sm_get(vs) / 4294967296.0; }
static void sm_set(void *vs, unsigned long seed) { ((sm_state *)vs)->s = seed; }
static const gsl_rng_type sm_type = {"splitmix", 0xffffffffUL, 0, sizeof(sm_state), sm_set, sm_get, sm_get_double};

#define CHECK(err, what) do { if (isError(*(err))) { fprintf(stderr, "FAILED in %s\n", what); printError(stderr, *(err)); return 10; } } while (0)

/* the test mixture: d = 4, three components with full covariances */
static mix_mvdens *test_mix(error **err) {
  const int K = 3, d = 4;
  mix_mvdens *m = mix_mvdens_alloc(K, d, err);
  if (isError(*err)) return NULL;
  const double w[3] = {0.5, 0.3, 0.2};
  for (int k = 0; k < K; ++k) {
    m->wght[k] = w[k];
    for (int a = 0; a < d; ++a) {
      m->comp[k]->mean[a] = 3.0 * k - 2.0 + 0.5 * a;
      for (int b = 0; b < d; ++b) m->comp[k]->std[a * d + b] = (a == b ? 1.0 + 0.5 * k + 0.1 * a : 0.2 / (1.0 + abs(a - b)));
    }
  }
  return m;
}

static int mode_wire(const char *outname) {
  error *_err = initError(), **err = &_err;
  FILE *out = fopen(outname, "w");
  if (!out) return 5;
  mix_mvdens *m = test_mix(err); CHECK(err, "test_mix");
  size_t sz = 0;
  void *buf = serialize_mix_mvdens(m, &sz, err); CHECK(err, "serialize");
  fprintf(out, "size %zu expected %zu\n", sz, mix_mvdens_size(m->ncomp, m->ndim));
  mix_mvdens *m2 = deserialize_mix_mvdens(buf, sz, err); CHECK(err, "deserialize");
  size_t sz2 = 0;
  void *buf2 = serialize_mix_mvdens(m2, &sz2, err); CHECK(err, "serialize 2");
  fprintf(out, "roundtrip %d\n", sz == sz2 && memcmp(buf, buf2, sz) == 0);
  /* the wire format itself (mvdens.c:596-660): K, d, weights, then per component df, band_limit, mean, factor */
  {
    const char *p = (const char *)buf;
    size_t K, d;
    memcpy(&K, p, sizeof K); memcpy(&d, p + sizeof K, sizeof d);
    double w0;
    memcpy(&w0, p + 2 * sizeof(size_t), sizeof w0);
    int df0;
    memcpy(&df0, p + 2 * sizeof(size_t) + K * sizeof(double), sizeof df0);
    fprintf(out, "wire K %zu d %zu w0 %.17g df0 %d chol %d\n", K, d, w0, df0, m2->comp[0]->chol);
  }
  /* a bad size must raise mv_serialize (-709) */
  mix_mvdens *bad = deserialize_mix_mvdens(buf, sz - 8, err);
  fprintf(out, "badsize null %d code %d\n", bad == NULL, getErrorValue(*err));
  purgeError(err);
  free(buf); free(buf2);
  /* shallow copy (mvdens.c:418-431) */
  mix_mvdens *m3 = mix_mvdens_alloc(m->ncomp, m->ndim, err); CHECK(err, "alloc");
  mvdens **own = malloc(sizeof(mvdens *) * m3->ncomp);
  memcpy(own, m3->comp, sizeof(mvdens *) * m3->ncomp);
  mix_mvdens_copy(m3, m, err); CHECK(err, "copy");
  fprintf(out, "copy weights %d shares_components %d\n", memcmp(m3->wght, m->wght, sizeof(double) * m->ncomp) == 0, m3->comp[1] == m->comp[1]);
  memcpy(m3->comp, own, sizeof(mvdens *) * m3->ncomp);  /* give it back its own components before freeing */
  free(own);
  fprintf(out, "enc %.17g\n", effective_number_of_components(m, err)); CHECK(err, "enc");
  fprintf(out, "lndet %.17g det %.17g\n", ln_determinant(m->comp[0]->std, m->ndim, err), determinant(m->comp[0]->std, m->ndim)); CHECK(err, "lndet");
  /* mvdens_inverse: Sigma * Sigma^-1 = 1 */
  {
    const int d = 4;
    mvdens *g = mvdens_alloc(d, err); CHECK(err, "mvdens_alloc");
    double S[16];
    for (int a = 0; a < d; ++a) for (int b = 0; b < d; ++b) S[a * d + b] = g->std[a * d + b] = (a == b ? 2.0 + a : 0.3);
    mvdens_inverse(g, err); CHECK(err, "inverse");
    double worst = 0;
    for (int a = 0; a < d; ++a) for (int b = 0; b < d; ++b) {
      double s = 0;
      for (int k = 0; k < d; ++k) s += S[a * d + k] * g->std[k * d + b];
      const double e = fabs(s - (a == b));
      if (e > worst) worst = e;
    }
    fprintf(out, "inverse worst %.3e chol %d\n", worst, g->chol);
    mvdens_free(&g);
  }
  /* text dump / read back (mvdens.c:121-188) */
  {
    char name[4096];
    snprintf(name, sizeof name, "%s.mvd", outname);
    mvdens_chdump(name, m->comp[1], err); CHECK(err, "chdump");
    FILE *f = fopen(name, "r");
    mvdens *g = mvdens_read_and_chol(f, err); CHECK(err, "read_and_chol");
    fclose(f);
    double worst = 0;
    for (size_t i = 0; i < m->ndim * m->ndim; ++i) { const double e = fabs(g->std[i] - m->comp[1]->std[i]); if (e > worst) worst = e; }
    fprintf(out, "dump_read worst %.3e chol %d df %d\n", worst, g->chol, g->df);
    mvdens_free(&g);
  }
  /* names (distribution.c:129-191) */
  {
    distribution *t = init_bentGauss_distribution(3, 0.1, 10.0, err); CHECK(err, "bentGauss");
    char *names[3] = {(char *)"omega", (char *)"sigma8", (char *)"h"};
    distribution_set_names(t, names, err); CHECK(err, "set_names");
    fprintf(out, "name sigma8 %d h %d\n", distribution_get_name(t, (char *)"sigma8", err), distribution_get_name(t, (char *)"h", err));
    int miss = distribution_get_name(t, (char *)"nope", err);
    fprintf(out, "name missing %d code %d\n", miss, getErrorValue(*err));
    purgeError(err);
    int idim[1];
    double loc[1] = {0.7}, var[1] = {0.01};
    char *pn[1] = {(char *)"h"};
    distribution *c = add_gaussian_prior_name(t, 1, pn, loc, var, err); CHECK(err, "prior_name");
    (void)idim;
    double x[3] = {0.3, 0.8, 0.72};
    const double v = distribution_lkl(c, x, err); CHECK(err, "combine lkl");
    const double want = distribution_lkl(t, x, err) - 0.5 * (0.02 * 0.02 / 0.01) - 0.5 * log(2 * M_PI) - log(0.1);
    fprintf(out, "prior_name value %.17g want %.17g\n", v, want);
    free_distribution(&c);
  }
  /* error chain pickling (errorlist.c:126-169) */
  {
    error *e = newError(-42, (char *)"inner", (char *)"first", NULL, NULL);
    e = newError(-43, (char *)"outer", (char *)"second", e, NULL);
    void *pk = NULL;
    const int n = pickleError(e, &pk);
    error *back = unpickleError(pk, n);
    fprintf(out, "pickle n %d top %d next %d text %s\n", n, back->errValue, back->next ? back->next->errValue : 0, back->next ? back->next->errText : "");
    free(pk);
  }
  /* parabox printing (parabox.c:147-186) */
  {
    parabox *pb = init_parabox(2, err); CHECK(err, "parabox");
    add_slab(pb, 0, -1.0, 2.0, err); CHECK(err, "slab");
    print_parabox(pb, out);
    free_parabox(&pb);
  }
  /* sort_indices (pmc.c:881-1034) */
  {
    double list[6] = {0.3, -1.0, 7.5, 0.1, 2.0, -3.0};
    size_t *ind = sort_indices(list, 6, NULL, NULL, 1, err); CHECK(err, "sort");
    fprintf(out, "sort");
    for (int i = 0; i < 6; ++i) fprintf(out, " %zu", ind[i]);
    fprintf(out, "\n");
    free(ind);
  }
  mix_mvdens_free(&m3);
  mix_mvdens_free(&m2);
  mix_mvdens_free(&m);
  fclose(out);
  return 0;
}

static int mode_ran(const char *outname, long M, int nu) {
  error *_err = initError(), **err = &_err;
  sm_state st = {777};
  gsl_rng rng = {&sm_type, &st};
  mix_mvdens *m = test_mix(err); CHECK(err, "test_mix");
  const int d = (int)m->ndim;
  double *x = malloc(sizeof(double) * (size_t)M * d), *t = malloc(sizeof(double) * (size_t)M * d);
  size_t *idx = malloc(sizeof(size_t) * M);
  for (long i = 0; i < M; ++i) {
    double *r = mix_mvdens_ran(x + (size_t)i * d, idx + i, m, &rng, err); CHECK(err, "mix_mvdens_ran");
    if (r != x + (size_t)i * d) return 11;
  }
  double *fresh = mix_mvdens_ran(NULL, NULL, m, &rng, err); CHECK(err, "mix_mvdens_ran alloc");  /* dest == NULL allocates */
  free(fresh);
  mvdens *g = mvdens_alloc(d, err); CHECK(err, "mvdens_alloc");
  for (int a = 0; a < d; ++a) { g->mean[a] = 1.0 + a; for (int b = 0; b < d; ++b) g->std[a * d + b] = (a == b ? 2.0 : 0.5); }
  g->df = nu;
  for (long i = 0; i < M; ++i) { mvdens_ran(t + (size_t)i * d, g, &rng, err); CHECK(err, "mvdens_ran"); }
  FILE *out = fopen(outname, "wb");
  if (!out) return 5;
  fwrite(x, sizeof(double), (size_t)M * d, out);
  fwrite(idx, sizeof(size_t), M, out);
  fwrite(t, sizeof(double), (size_t)M * d, out);
  fclose(out);
  printf("init_cwght %d cw_last %.17g\n", m->init_cwght, m->cwght[m->ncomp - 1]);
  return 0;
}

static int mode_file(const char *inname, const char *outname, int N, int d) {
  error *_err = initError(), **err = &_err;
  mix_mvdens *m = mix_mvdens_alloc(2, d, err); CHECK(err, "alloc");
  m->wght[0] = 0.6; m->wght[1] = 0.4;
  for (int k = 0; k < 2; ++k)
    for (int a = 0; a < d; ++a) {
      m->comp[k]->mean[a] = k ? 1.0 : -1.0;
      for (int b = 0; b < d; ++b) m->comp[k]->std[a * d + b] = (a == b ? 2.0 + k : 0.1);
    }
  FILE *f = fopen(inname, "r");
  if (!f) return 5;
  pmc_simu *ps = pmc_simu_from_file(f, N, d, 0, m, 0, err); CHECK(err, "pmc_simu_from_file");
  fclose(f);
  FILE *out = fopen(outname, "wb");
  if (!out) return 5;
  double sc[5] = {(double)ps->nsamples, ps->logSum, ps->maxW, ps->maxR, (double)ps->isLog};
  fwrite(sc, sizeof sc, 1, out);
  fwrite(ps->X, sizeof(double), (size_t)ps->nsamples * d, out);
  fwrite(ps->indices, sizeof(size_t), ps->nsamples, out);
  fwrite(ps->weights, sizeof(double), ps->nsamples, out);
  fwrite(ps->log_rho, sizeof(double), ps->nsamples, out);
  fwrite(ps->flg, sizeof(short), ps->nsamples, out);
  fclose(out);
  /* a header that disagrees with the arguments must raise tls_Nparam */
  f = fopen(inname, "r");
  pmc_simu *bad = pmc_simu_from_file(f, N, d + 1, 0, m, 0, err);
  fclose(f);
  printf("bad_header null %d code %d\n", bad == NULL, getErrorValue(*err));
  purgeError(err);
  pmc_simu_free(&ps);
  mix_mvdens_free(&m);
  return 0;
}

/* dump <in.bin> <prefix>: the on-disk formats (byte-compared with files the reference's own functions wrote) */
static int mode_dump(const char *inname, const char *prefix) {
  error *_err = initError(), **err = &_err;
  FILE *in = fopen(inname, "rb");
  if (!in) return 3;
  long hdr[4]; /* N, d, K, isLog */
  double logSum;
  if (fread(hdr, sizeof hdr, 1, in) != 1 || fread(&logSum, sizeof logSum, 1, in) != 1) return 4;
  const long N = hdr[0];
  const int d = (int)hdr[1], K = (int)hdr[2];
  pmc_simu *ps = pmc_simu_init(N, d, err); CHECK(err, "pmc_simu_init");
  mix_mvdens *m = mix_mvdens_alloc(K, d, err); CHECK(err, "mix_mvdens_alloc");
  int ok = fread(ps->X, sizeof(double), (size_t)N * d, in) == (size_t)N * d && fread(ps->indices, sizeof(size_t), N, in) == (size_t)N &&
           fread(ps->weights, sizeof(double), N, in) == (size_t)N && fread(ps->log_rho, sizeof(double), N, in) == (size_t)N &&
           fread(ps->flg, sizeof(short), N, in) == (size_t)N && fread(m->wght, sizeof(double), K, in) == (size_t)K;
  for (int k = 0; k < K && ok; ++k) ok = fread(m->comp[k]->mean, sizeof(double), d, in) == (size_t)d;
  for (int k = 0; k < K && ok; ++k) ok = fread(m->comp[k]->std, sizeof(double), (size_t)d * d, in) == (size_t)d * d;
  for (int k = 0; k < K && ok; ++k) ok = fread(&m->comp[k]->df, sizeof(int), 1, in) == 1;
  fclose(in);
  if (!ok) return 4;
  ps->logSum = logSum; ps->isLog = (int)hdr[3];
  char name[4096];
  FILE *f;
  snprintf(name, sizeof name, "%s.mix", prefix); f = fopen(name, "w"); mix_mvdens_dump(f, m); fclose(f);
  snprintf(name, sizeof name, "%s.mvd", prefix); f = fopen(name, "w"); mvdens_dump(f, m->comp[0]); fclose(f);
  snprintf(name, sizeof name, "%s.psim", prefix); f = fopen(name, "w"); pmc_simu_dump(f, ps, err); fclose(f); CHECK(err, "pmc_simu_dump");
  snprintf(name, sizeof name, "%s.out", prefix); out_pmc_simu(name, ps, err); CHECK(err, "out_pmc_simu");
  /* and back: mix_mvdens_dwnp reads what mix_mvdens_dump wrote */
  snprintf(name, sizeof name, "%s.mix", prefix); f = fopen(name, "r");
  mix_mvdens *back = mix_mvdens_dwnp(f, err); CHECK(err, "mix_mvdens_dwnp");
  fclose(f);
  double worst = 0;
  for (int k = 0; k < K; ++k)
    for (int i = 0; i < d * d; ++i) { const double e = fabs(back->comp[k]->std[i] - m->comp[k]->std[i]); if (e > worst) worst = e; }
  printf("dwnp K %zu d %zu worst %.3e df0 %d\n", back->ncomp, back->ndim, worst, back->comp[0]->df);
  pmc_simu_free(&ps);
  return 0;
}

/* point <in.bin> <out.bin>: the point-wise entry points on M points of dimension d: mvdens_log_pdf (Gaussian and
 * Student-t), scalar_product, mix_mvdens_log_pdf, bentGauss_lkl, isinBox */
static int mode_point(const char *inname, const char *outname) {
  error *_err = initError(), **err = &_err;
  FILE *in = fopen(inname, "rb");
  if (!in) return 3;
  long hdr[2]; /* M, d */
  if (fread(hdr, sizeof hdr, 1, in) != 1) return 4;
  const long M = hdr[0];
  const int d = (int)hdr[1];
  double *X = malloc(sizeof(double) * (size_t)M * d), *mean = malloc(sizeof(double) * d), *cov = malloc(sizeof(double) * d * d);
  if (fread(X, sizeof(double), (size_t)M * d, in) != (size_t)M * d || fread(mean, sizeof(double), d, in) != (size_t)d ||
      fread(cov, sizeof(double), (size_t)d * d, in) != (size_t)d * d) return 4;
  fclose(in);
  mvdens *g = mvdens_alloc(d, err); CHECK(err, "mvdens_alloc");
  memcpy(g->mean, mean, sizeof(double) * d); memcpy(g->std, cov, sizeof(double) * d * d);
  mvdens *t = mvdens_alloc(d, err); CHECK(err, "mvdens_alloc");
  memcpy(t->mean, mean, sizeof(double) * d); memcpy(t->std, cov, sizeof(double) * d * d);
  t->df = 4;
  mix_mvdens *m = mix_mvdens_alloc(2, d, err); CHECK(err, "mix_mvdens_alloc");
  m->wght[0] = 0.25; m->wght[1] = 0.75;
  for (int k = 0; k < 2; ++k) {
    for (int a = 0; a < d; ++a) m->comp[k]->mean[a] = mean[a] + (k ? 0.5 : -0.5);
    memcpy(m->comp[k]->std, cov, sizeof(double) * d * d);
  }
  bentGauss *bg = init_bentGauss(d, 0.1, 10.0, err); CHECK(err, "bentGauss");
  parabox *pb = init_parabox(d, err); CHECK(err, "parabox");
  add_slab(pb, 0, -1.0, 1.5, err); add_hibound(pb, 1, 2.0, err); CHECK(err, "box");
  double *out = malloc(sizeof(double) * (size_t)M * 7);
  for (long i = 0; i < M; ++i) {
    double *x = X + (size_t)i * d;
    out[7 * i + 0] = mvdens_log_pdf(g, x, err);
    out[7 * i + 1] = scalar_product(g, x, err);
    out[7 * i + 2] = mvdens_log_pdf(t, x, err);
    out[7 * i + 3] = scalar_product(t, x, err);  /* must not depend on df nor change it */
    out[7 * i + 4] = mix_mvdens_log_pdf(m, x, err);
    out[7 * i + 5] = bentGauss_lkl(bg, x, err);
    out[7 * i + 6] = (double)isinBox(pb, x, err);
    CHECK(err, "point-wise");
  }
  printf("df_after %d chol %d\n", t->df, t->chol);
  FILE *f = fopen(outname, "wb");
  fwrite(out, sizeof(double), (size_t)M * 7, f);
  fclose(f);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 4 && !strcmp(argv[1], "point")) return mode_point(argv[2], argv[3]);
  if (argc >= 4 && !strcmp(argv[1], "dump")) return mode_dump(argv[2], argv[3]);
  if (argc >= 3 && !strcmp(argv[1], "wire")) return mode_wire(argv[2]);
  if (argc >= 5 && !strcmp(argv[1], "ran")) return mode_ran(argv[2], atol(argv[3]), atoi(argv[4]));
  if (argc >= 6 && !strcmp(argv[1], "file")) return mode_file(argv[2], argv[3], atoi(argv[4]), atoi(argv[5]));
  fprintf(stderr, "usage: see the header of api2_driver.c\n");
  return 2;
}
