/* Flat C driver over the UNMODIFIED reference (compiled into oracle/_ref/libpmcref.so by oracle/Makefile).
 * TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product library never does.
 *
 * Every function builds the reference's own objects (pmc_simu, mix_mvdens, distribution, parabox) and calls
 * the reference's own functions.  pmc_simu_pmc_step is not used because it fails in this snapshot
 * (pmc.c:475-476 passes t_iter_tempered=-1, pmc.c:508 rejects it); the step is spelled out with
 * generic_get_importance_weight_and_deduced_verb(..., t, quiet=1) as SURVEY.md §8c prescribes.
 * Return value: 0 or the reference's (negative) error code.
 */
#include "pmc.h"
#include "sampleTarget.h"
#include <time.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

/* rc_* are referenced by sampleTarget.c:rcinit_bentGauss only (Lua config glue, never called here) */
confFile *rc_alias(confFile *rc, char *root, error **err) { (void)rc; (void)root; (void)err; abort(); }
void rc_close(confFile **rc) { (void)rc; abort(); }
long rc_get_integer(confFile *rc, char *key, error **err) { (void)rc; (void)key; (void)err; abort(); }
double rc_get_real(confFile *rc, char *key, error **err) { (void)rc; (void)key; (void)err; abort(); }

enum { REF_TGT_BANANA = 1, REF_TGT_MVDENS = 2, REF_TGT_MIX = 3 };

static int take_err(error **err) {
  int v = 0;
  if (isError(*err)) { v = getErrorValue(*err); purgeError(err); if (v == 0) v = -1; }
  return v;
}

/* mixture from flat arrays: wght[K], mean[K*d], cov[K*d*d] (covariance, chol=0 -> reference factorises), df[K] */
/* band limit applied to every mixture make_mix builds (0 = leave the default, band_limit = ndim) */
static size_t g_band_limit = 0;
void ref_set_band_limit(size_t b) { g_band_limit = b; }

static mix_mvdens *make_mix(int K, int d, const double *wght, const double *mean, const double *cov,
                            const int *df, int is_chol, error **err) {
  mix_mvdens *m = mix_mvdens_alloc(K, d, err);
  if (isError(*err)) return NULL;
  if (g_band_limit > 0) mix_mvdens_set_band_limit(m, g_band_limit);
  for (int k = 0; k < K; ++k) {
    m->wght[k] = wght[k];
    memcpy(m->comp[k]->mean, mean + (size_t)k * d, sizeof(double) * d);
    memcpy(m->comp[k]->std, cov + (size_t)k * d * d, sizeof(double) * d * d);
    m->comp[k]->df = df ? df[k] : -1;
    if (is_chol) { m->comp[k]->chol = 1; m->comp[k]->detL = determinant(m->comp[k]->std, d); }
  }
  return m;
}

static void export_mix(mix_mvdens *m, double *wght, double *mean, double *std, int *chol, double *detL) {
  const int K = m->ncomp, d = m->ndim;
  for (int k = 0; k < K; ++k) {
    wght[k] = m->wght[k];
    memcpy(mean + (size_t)k * d, m->comp[k]->mean, sizeof(double) * d);
    memcpy(std + (size_t)k * d * d, m->comp[k]->std, sizeof(double) * d * d);
    if (chol) chol[k] = m->comp[k]->chol;
    if (detL) detL[k] = m->comp[k]->detL;
  }
}

/* target: kind 1 = bentGauss(tp[0]=b, tp[1]=sig1sq); kind 2 = single mvdens; kind 3 = mix_mvdens.
 * For kinds 2/3 the density is given as (tK, tw, tmean, tcov, tdf). */
static distribution *make_target(int kind, int d, const double *tp, int tK, const double *tw, const double *tmean,
                                 const double *tcov, const int *tdf, error **err) {
  if (kind == REF_TGT_BANANA) return init_bentGauss_distribution(d, tp[0], tp[1], err);
  if (kind == REF_TGT_MVDENS) {
    mvdens *g = mvdens_alloc(d, err);
    if (isError(*err)) return NULL;
    memcpy(g->mean, tmean, sizeof(double) * d);
    memcpy(g->std, tcov, sizeof(double) * d * d);
    g->df = tdf ? tdf[0] : -1;
    return init_distribution(d, g, mvdens_log_pdf_void, mvdens_free_void, NULL, err);
  }
  if (kind == REF_TGT_MIX) {
    mix_mvdens *m = make_mix(tK, d, tw, tmean, tcov, tdf, 0, err);
    if (isError(*err)) return NULL;
    return init_distribution(d, m, mix_mvdens_log_pdf_void, mix_mvdens_free_void, NULL, err);
  }
  return NULL;
}

static parabox *make_box(int d, int nfaces, const double *faces, error **err) {
  if (nfaces <= 0 || !faces) return NULL;
  parabox *pb = init_parabox(d, err);
  if (isError(*err)) return NULL;
  for (int f = 0; f < nfaces; ++f) {
    add_face(pb, (double *)(faces + (size_t)f * (d + 1)), faces[(size_t)f * (d + 1) + d], err);
    if (isError(*err)) return NULL;
  }
  return pb;
}

/* ---- single functions ------------------------------------------------------------------------ */
int ref_cholesky(int d, double *a, double *detL) {
  error *e = initError(), **err = &e;
  mvdens *g = mvdens_alloc(d, err);
  memcpy(g->std, a, sizeof(double) * d * d);
  mvdens_cholesky_decomp(g, err);
  int rc = take_err(err);
  memcpy(a, g->std, sizeof(double) * d * d);
  *detL = g->detL;
  mvdens_free(&g);
  endError(err);
  return rc;
}

int ref_mix_log_pdf(long N, int d, int K, const double *X, const double *wght, const double *mean,
                    const double *cov, const int *df, double *out) {
  error *e = initError(), **err = &e;
  mix_mvdens *m = make_mix(K, d, wght, mean, cov, df, 0, err);
  int rc = take_err(err);
  for (long i = 0; i < N && !rc; ++i) { out[i] = mix_mvdens_log_pdf(m, X + i * d, err); rc = take_err(err); }
  if (m) mix_mvdens_free(&m);
  endError(err);
  return rc;
}

int ref_mvdens_log_pdf(long N, int d, const double *X, const double *mean, const double *cov, int df, double *out) {
  error *e = initError(), **err = &e;
  mvdens *g = mvdens_alloc(d, err);
  memcpy(g->mean, mean, sizeof(double) * d);
  memcpy(g->std, cov, sizeof(double) * d * d);
  g->df = df;
  int rc = 0;
  for (long i = 0; i < N && !rc; ++i) { out[i] = mvdens_log_pdf(g, X + i * d, err); rc = take_err(err); }
  mvdens_free(&g);
  endError(err);
  return rc;
}

int ref_target_log_pdf(long N, int d, const double *X, int kind, const double *tp, int tK, const double *tw,
                       const double *tmean, const double *tcov, const int *tdf, double *out) {
  error *e = initError(), **err = &e;
  distribution *t = make_target(kind, d, tp, tK, tw, tmean, tcov, tdf, err);
  int rc = take_err(err);
  for (long i = 0; i < N && !rc; ++i) { out[i] = distribution_lkl(t, X + i * d, err); rc = take_err(err); }
  if (t) free_distribution(&t);
  endError(err);
  return rc;
}

/* A target with Gaussian priors, built by the reference's own add_gaussian_prior / add_gaussian_prior_2
 * (distribution.c:548-606) around make_target's density, evaluated through distribution_lkl -> combine_lkl (:350-396).
 * full_cov = 0: var[nprior] (diagonal); 1: var[nprior*nprior]. */
int ref_prior_log_pdf(long N, int d, const double *X, int kind, const double *tp, int tK, const double *tw,
                      const double *tmean, const double *tcov, const int *tdf, int nprior, const int *idim,
                      const double *loc, const double *var, int full_cov, double *out) {
  error *e = initError(), **err = &e;
  distribution *t = make_target(kind, d, tp, tK, tw, tmean, tcov, tdf, err);
  int rc = take_err(err);
  distribution *c = NULL;
  if (!rc) {
    c = full_cov ? add_gaussian_prior_2(t, nprior, (int *)idim, (double *)loc, (double *)var, err)
                 : add_gaussian_prior(t, nprior, (int *)idim, (double *)loc, (double *)var, err);
    rc = take_err(err);
  }
  for (long i = 0; i < N && !rc; ++i) { out[i] = distribution_lkl(c, X + i * d, err); rc = take_err(err); }
  if (c) free_distribution(&c);
  endError(err);
  return rc;
}

/* the on-disk formats either side of the PMC path, written by the reference's own functions: <prefix>.mix
 * (mix_mvdens_dump, mvdens.c:513-528), .mvd (mvdens_dump of component 0, :114-134), .psim (pmc_simu_dump, pmc.c:342-371),
 * .out (out_pmc_simu, pmc.c:306-325) */
int ref_dump_all(const char *prefix, long N, int d, int K, const double *X, const size_t *indices, const double *w,
                 const double *log_rho, const short *flg, double logSum, int isLog, const double *wght, const double *mean,
                 const double *cov, const int *df) {
  error *e = initError(), **err = &e;
  char name[4096];
  mix_mvdens *m = make_mix(K, d, wght, mean, cov, df, 0, err);
  pmc_simu *ps = pmc_simu_init(N, d, err);
  int rc = take_err(err);
  if (!rc) {
    memcpy(ps->X, X, sizeof(double) * N * d);
    memcpy(ps->indices, indices, sizeof(size_t) * N);
    memcpy(ps->weights, w, sizeof(double) * N);
    memcpy(ps->log_rho, log_rho, sizeof(double) * N);
    memcpy(ps->flg, flg, sizeof(short) * N);
    ps->logSum = logSum; ps->isLog = isLog;
    FILE *f;
    snprintf(name, sizeof name, "%s.mix", prefix); f = fopen(name, "w"); mix_mvdens_dump(f, m); fclose(f);
    snprintf(name, sizeof name, "%s.mvd", prefix); f = fopen(name, "w"); mvdens_dump(f, m->comp[0]); fclose(f);
    snprintf(name, sizeof name, "%s.psim", prefix); f = fopen(name, "w"); pmc_simu_dump(f, ps, err); fclose(f);
    rc = take_err(err);
    snprintf(name, sizeof name, "%s.out", prefix);
    if (!rc) { out_pmc_simu(name, ps, err); rc = take_err(err); }
  }
  if (ps) pmc_simu_free(&ps);
  if (m) mix_mvdens_free(&m);
  endError(err);
  return rc;
}

/* ranindx with injected uniforms (mvdens.c:766-782): a gsl_rng whose get_double replays z[] */
typedef struct { const double *z; long pos; } replay_state;
static double replay_get_double(void *s) { replay_state *r = (replay_state *)s; return r->z[r->pos++]; }
static unsigned long replay_get(void *s) { return (unsigned long)(replay_get_double(s) * 4294967296.0); }
static void replay_set(void *s, unsigned long seed) { (void)s; (void)seed; }
static const gsl_rng_type replay_type = {"replay", 0xffffffffUL, 0, sizeof(replay_state), &replay_set, &replay_get, &replay_get_double};

int ref_ranindx(long N, int K, const double *cw, const double *z, size_t *idx) {
  error *e = initError(), **err = &e;
  replay_state st = {z, 0};
  gsl_rng r = {&replay_type, &st};
  for (long i = 0; i < N; ++i) idx[i] = ranindx((double *)cw, K, &r, err);
  int rc = take_err(err);
  endError(err);
  return rc;
}

int ref_isinbox(long N, int d, int nfaces, const double *faces, const double *X, int *inside) {
  error *e = initError(), **err = &e;
  parabox *pb = make_box(d, nfaces, faces, err);
  int rc = take_err(err);
  for (long i = 0; i < N && !rc; ++i) { inside[i] = isinBox(pb, (double *)(X + i * d), err); rc = take_err(err); }
  if (pb) free_parabox(&pb);
  endError(err);
  return rc;
}

/* simulate_mix_mvdens (pmc.c:266-298) with the shim's mt19937 */
int ref_simulate(long N, int d, int K, const double *wght, const double *mean, const double *cov, const int *df,
                 int nfaces, const double *faces, unsigned long seed, double *X, size_t *indices, double *wght_after) {
  error *e = initError(), **err = &e;
  int rc = 0;
  gsl_rng *r = gsl_rng_alloc(gsl_rng_default);
  gsl_rng_set(r, seed);
  mix_mvdens *m = make_mix(K, d, wght, mean, cov, df, 0, err);
  parabox *pb = make_box(d, nfaces, faces, err);
  pmc_simu *ps = pmc_simu_init(N, d, err);
  rc = take_err(err);
  if (!rc) { simulate_mix_mvdens(ps, m, r, pb, err); rc = take_err(err); }
  if (!rc) {
    memcpy(X, ps->X, sizeof(double) * N * d);
    memcpy(indices, ps->indices, sizeof(size_t) * N);
    if (wght_after) memcpy(wght_after, m->wght, sizeof(double) * K);
  }
  if (ps) pmc_simu_free(&ps);
  if (pb) free_parabox(&pb);
  if (m) mix_mvdens_free(&m);
  gsl_rng_free(r);
  endError(err);
  return rc;
}

/* One PMC iteration on GIVEN samples (X, indices): weights -> normalise -> perplexity/ESS -> evidence ->
 * update_prop_rb (do_update=1) or update_prop (do_update=2) or none (0).
 * in_retry seeds psim->retry.  All outputs are optional except where noted. */
int ref_iteration(long N, int d, int K, const double *X, const size_t *indices,
                  const double *wght, const double *mean, const double *cov, const int *df,
                  int tkind, const double *tp, int tK, const double *tw, const double *tmean, const double *tcov,
                  const int *tdf, double t_temper, int do_update, int in_retry,
                  double *log_rho, double *logw, double *w_norm, short *flg,
                  double *scal /* [8]: nok, maxW, maxR, logSum, perp, ess, ln_evidence, retry_out */,
                  double *o_wght, double *o_mean, double *o_std, int *o_chol, double *o_detL) {
  error *e = initError(), **err = &e;
  int rc = 0;
  mix_mvdens *m = make_mix(K, d, wght, mean, cov, df, 0, err);
  rc = take_err(err);
  distribution *tgt = rc ? NULL : make_target(tkind, d, tp, tK, tw, tmean, tcov, tdf, err);
  if (!rc) rc = take_err(err);
  pmc_simu *ps = rc ? NULL : pmc_simu_init(N, d, err);
  if (!rc) rc = take_err(err);
  if (!rc) {
    pmc_simu_init_classic_pmc(ps, tgt, NULL, m, 0, NULL, NULL, err);  /* psim now owns tgt and m */
    rc = take_err(err);
  }
  if (!rc) {
    memcpy(ps->X, X, sizeof(double) * N * d);
    if (indices) memcpy(ps->indices, indices, sizeof(size_t) * N);
    for (long i = 0; i < N; ++i) ps->flg[i] = 1;
    ps->retry = in_retry;
    size_t nok = generic_get_importance_weight_and_deduced_verb(ps, ps->proposal, &distribution_lkl, &distribution_lkl,
                                                                &distribution_retrieve, ps->target, t_temper, 1, err);
    rc = take_err(err);
    scal[0] = (double)nok; scal[1] = ps->maxW; scal[2] = ps->maxR;
    if (log_rho) memcpy(log_rho, ps->log_rho, sizeof(double) * N);
    if (logw) memcpy(logw, ps->weights, sizeof(double) * N);
  }
  if (!rc) {
    normalize_importance_weight(ps, err);
    rc = take_err(err);
    scal[3] = ps->logSum;
    if (w_norm) memcpy(w_norm, ps->weights, sizeof(double) * N);
  }
  if (!rc) {
    double ess = 0;
    scal[4] = perplexity_and_ess(ps, MC_UNORM, &ess, err);
    rc = take_err(err);
    scal[5] = ess;
  }
  if (!rc) {
    double lne = 0;
    evidence(ps, &lne, err);
    rc = take_err(err);
    scal[6] = lne;
  }
  if (!rc && do_update) {
    if (do_update == 1) update_prop_rb(m, ps, err); else update_prop(m, ps, err);
    rc = take_err(err);
    scal[7] = ps->retry;
    if (o_wght) export_mix(m, o_wght, o_mean, o_std, o_chol, o_detL);
  }
  if (ps && flg) memcpy(flg, ps->flg, sizeof(short) * N);
  if (ps) pmc_simu_free(&ps);
  else { if (tgt) free_distribution(&tgt); if (m) mix_mvdens_free(&m); }
  endError(err);
  return rc;
}

/* update_prop_rb alone on a given, already normalised sample (weights, log_rho, flg, indices, maxR as stored) */
int ref_update_rb(long N, int d, int K, const double *X, const size_t *indices, const double *w_norm,
                  const double *log_rho, const short *flg, double maxR, int in_retry,
                  const double *wght, const double *mean, const double *cov, const int *df, int is_chol,
                  double *o_wght, double *o_mean, double *o_std, int *o_chol, double *o_detL, int *o_retry) {
  error *e = initError(), **err = &e;
  mix_mvdens *m = make_mix(K, d, wght, mean, cov, df, is_chol, err);
  pmc_simu *ps = pmc_simu_init(N, d, err);
  int rc = take_err(err);
  if (!rc) {
    memcpy(ps->X, X, sizeof(double) * N * d);
    memcpy(ps->indices, indices, sizeof(size_t) * N);
    memcpy(ps->weights, w_norm, sizeof(double) * N);
    memcpy(ps->log_rho, log_rho, sizeof(double) * N);
    memcpy(ps->flg, flg, sizeof(short) * N);
    ps->maxR = maxR; ps->isLog = 0; ps->retry = in_retry;
    update_prop_rb(m, ps, err);
    rc = take_err(err);
    *o_retry = ps->retry;
    export_mix(m, o_wght, o_mean, o_std, o_chol, o_detL);
  }
  if (ps) pmc_simu_free(&ps);
  if (m) mix_mvdens_free(&m);
  endError(err);
  return rc;
}

/* ---- timed PMC loop = vanilla_pmc.c:75-128 without the file dumps (CPU baseline) --------------- */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* nworkers <= 1: the single-process vanilla_pmc loop.
 * nworkers  > 1: emulation of vanilla_pmc_mpi (MPI is absent): rank 0 draws all samples, the weight loop runs on
 * nworkers forked processes over the slices of send_simulation (pmc_mpi.c:344-352) in a shared mapping, rank 0
 * alone normalises and updates (pmc_mpi.c:52-61,104-119).
 * times[5*it + {0..4}] = draw, weight, normalise, update, perplexity seconds; stats[3*it+{0,1,2}] = perp, ess, ln_evi.
 * update_mode: 1 = RB update each iteration, 0 = importance only. */
int ref_pmc_run(long N, int d, int K, const double *wght, const double *mean, const double *cov, const int *df,
                int tkind, const double *tp, int tK, const double *tw, const double *tmean, const double *tcov,
                const int *tdf, int nfaces, const double *faces, unsigned long seed, int niter, int update_mode,
                int nworkers, double *times, double *stats,
                double *o_wght, double *o_mean, double *o_std) {
  error *e = initError(), **err = &e;
  int rc = 0;
  gsl_rng *r = gsl_rng_alloc(gsl_rng_default);
  gsl_rng_set(r, seed);
  mix_mvdens *m = make_mix(K, d, wght, mean, cov, df, 0, err);
  distribution *tgt = make_target(tkind, d, tp, tK, tw, tmean, tcov, tdf, err);
  parabox *pb = make_box(d, nfaces, faces, err);
  pmc_simu *ps = pmc_simu_init(N, d, err);
  rc = take_err(err);
  if (rc) { endError(err); return rc; }
  pmc_simu_init_classic_pmc(ps, tgt, pb, m, 0, NULL, NULL, err);
  rc = take_err(err);
  double *sh_w = NULL, *sh_r = NULL, *sh_mx = NULL; short *sh_f = NULL;
  if (nworkers > 1) {
    sh_w = mmap(NULL, sizeof(double) * N, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    sh_r = mmap(NULL, sizeof(double) * N, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    sh_f = mmap(NULL, sizeof(short) * N, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    sh_mx = mmap(NULL, sizeof(double) * 2 * nworkers, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  }
  for (int it = 0; it < niter && !rc; ++it) {
    double t0 = now_s();
    ps->proposal->simulate(ps, ps->proposal->data, r, ps->pb, err);
    rc = take_err(err); if (rc) break;
    double t1 = now_s();
    if (nworkers <= 1) {
      generic_get_importance_weight_and_deduced_verb(ps, ps->proposal, &distribution_lkl, &distribution_lkl,
                                                     &distribution_retrieve, ps->target, 1.0, 1, err);
      rc = take_err(err); if (rc) break;
    } else {
      /* slice layout of send_simulation: master keeps N/P, the first N%P workers get one more */
      long base = N / nworkers, rem = N % nworkers, start = 0;
      pid_t pids[256];
      for (int w = 0; w < nworkers; ++w) {
        long len = base + ((w > 0 && (w - 1) < rem) ? 1 : 0);
        if (w == 0 && nworkers == 1) len = N;
        if (w == nworkers - 1) len = N - start;
        pid_t pid = fork();
        if (pid == 0) {
          pmc_simu loc = *ps;
          loc.nsamples = len; loc.X = ps->X + start * d; loc.weights = sh_w + start; loc.log_rho = sh_r + start;
          loc.flg = sh_f + start; loc.X_ded = ps->X_ded;
          generic_get_importance_weight_and_deduced_verb(&loc, ps->proposal, &distribution_lkl, &distribution_lkl,
                                                         &distribution_retrieve, ps->target, 1.0, 1, err);
          sh_mx[2 * w] = loc.maxW; sh_mx[2 * w + 1] = loc.maxR;
          _exit(isError(*err) ? 1 : 0);
        }
        pids[w] = pid; start += len;
      }
      for (int w = 0; w < nworkers; ++w) { int st; waitpid(pids[w], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st)) rc = -1; }
      memcpy(ps->weights, sh_w, sizeof(double) * N);
      memcpy(ps->log_rho, sh_r, sizeof(double) * N);
      memcpy(ps->flg, sh_f, sizeof(short) * N);
      ps->maxW = sh_mx[0]; ps->maxR = sh_mx[1];
      for (int w = 1; w < nworkers; ++w) { if (sh_mx[2 * w] > ps->maxW) ps->maxW = sh_mx[2 * w]; if (sh_mx[2 * w + 1] > ps->maxR) ps->maxR = sh_mx[2 * w + 1]; }
      ps->isLog = 1;
      if (rc) break;
    }
    double t2 = now_s();
    normalize_importance_weight(ps, err);
    rc = take_err(err); if (rc) break;
    double t3 = now_s();
    if (update_mode) { ps->pmc_update(ps->proposal->data, ps, err); rc = take_err(err); if (rc) break; }
    double t4 = now_s();
    double ess = 0, lne = 0;
    double perp = perplexity_and_ess(ps, MC_UNORM, &ess, err);
    rc = take_err(err); if (rc) break;
    evidence(ps, &lne, err);
    rc = take_err(err); if (rc) break;
    double t5 = now_s();
    if (times) { times[5 * it] = t1 - t0; times[5 * it + 1] = t2 - t1; times[5 * it + 2] = t3 - t2; times[5 * it + 3] = t4 - t3; times[5 * it + 4] = t5 - t4; }
    if (stats) { stats[3 * it] = perp; stats[3 * it + 1] = ess; stats[3 * it + 2] = lne; }
  }
  if (!rc && o_wght) export_mix(m, o_wght, o_mean, o_std, NULL, NULL);
  if (sh_w) { munmap(sh_w, sizeof(double) * N); munmap(sh_r, sizeof(double) * N); munmap(sh_f, sizeof(short) * N); munmap(sh_mx, sizeof(double) * 2 * nworkers); }
  pmc_simu_free(&ps);
  gsl_rng_free(r);
  endError(err);
  return rc;
}

/* filter_histogram / mean_from_psim of the compiled reference on plain arrays */
long ref_filter_histogram(long N, double *weights, short *flg, int nbins, int cl, int cr) {
  error *e = initError(), **err = &e;
  pmc_simu ps;
  memset(&ps, 0, sizeof ps);
  ps.nsamples = N; ps.weights = weights; ps.flg = flg; ps.isLog = 1;
  int *fil = filter_histogram_init(nbins, cl, cr, err);
  long before = 0, after = 0;
  for (long i = 0; i < N; ++i) before += flg[i] == 1;
  filter_histogram(&ps, fil, err);
  int rc = take_err(err);
  for (long i = 0; i < N; ++i) after += flg[i] == 1;
  free(fil);
  endError(err);
  return rc ? rc : before - after;
}
double ref_mean_from_psim(double *X, double *w, short *flg, int N, int d, int a) { return mean_from_psim(X, w, flg, N, d, a); }

int ref_clip_weights(long N, double *weights, short *flg, int n) {
  error *e = initError(), **err = &e;
  pmc_simu ps;
  memset(&ps, 0, sizeof ps);
  ps.nsamples = N; ps.weights = weights; ps.flg = flg;
  clip_weights(&ps, n, NULL, err);
  int rc = take_err(err);
  endError(err);
  return rc;
}
