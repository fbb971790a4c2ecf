/* dropin_driver.c — a pmclib caller written against pmclib's own API (include/pmclib_abi.h mirrors the reference
 * headers), linked against libpmcb200.so.  It does what pmcexec/vanilla_pmc.c does between its Lua parsing and its file
 * dumps (vanilla_pmc.c:75-128): build target / proposal / box, loop pmc_simu_pmc_step, report perplexity / ESS /
 * evidence — and writes everything to a binary file the Python tests compare with the oracle.
 *
 *   dropin_driver <in.bin> <out.bin> <mode>
 *     mode 0: pmc_simu_pmc_step (fused GPU iteration)            mode 1: the same phase by phase through the public
 *     functions (simulate, _verb(t=1), normalize, update_prop_rb, perplexity_and_ess, evidence)
 *     mode 2: like 0 but the target is an OPAQUE host callback (user code), exercising the per-sample callback path
 *     mode 3: no-GPU check: object management and the error chain only; pmc_simu_pmc_step must fail loudly
 *     mode 4: with the shipped histogram filter; mode 5: prepare_update as its own entry point
 *     mode 6: the target carries a Gaussian prior (add_gaussian_prior -> combine_lkl), evaluated on the device
 *     mode 7: libpmc_mpi entry points (pmc_simu_init_mpi, pmc_simu_init_classic_pmc_mpi, pmc_simu_pmc_step_mpi), one process
 *             per GPU; rank and size come from the environment (PMCB200_RANK / PMCB200_WORLD_SIZE); rank 0 writes the file
 *     mode 8: mode 7 with the opaque host callback target (the likelihood loop each rank runs over its shard)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef PMCLIB_REFERENCE_HEADERS
/* compiled against the REFERENCE's own headers (pmclib/include, pmctools/include, examples/target) instead of the
 * mirror in include/pmclib_abi.h: proves that struct layouts and prototypes agree (tests/test_cpu_abi.py) */
#include "pmc.h"
#include "sampleTarget.h"
/* pmc_mpi.h needs <mpi.h>, which this image does not have: the three prototypes the driver uses, as pmc_mpi.h:52-66 has them */
pmc_simu *pmc_simu_init_mpi(long nsamples, int ndim, int nded, error **err);
double pmc_simu_pmc_step_mpi(pmc_simu *psim, gsl_rng *r, error **err);
void pmc_simu_init_classic_pmc_mpi(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data, int prop_print_step,
                                   void *filter_data, filter_func *pmc_filter, error **err);
#else
#include "pmclib_abi.h"
#endif

/* a tiny gsl_rng-compatible generator owned by the caller (the library only calls r->type->get) */
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

/* an "unknown" user likelihood: same formula as bentGauss but the library cannot recognise the pointer */
typedef struct { int d; double b, s1; long calls; } user_tgt;
static double user_lkl(void *v, const double *x, error **err) {
  user_tgt *t = (user_tgt *)v;
  double res = 0, x2 = x[0] * x[0];
  (void)err;
  t->calls++;
  res += x2 / t->s1; x2 -= t->s1; x2 *= t->b; x2 += x[1]; res += x2 * x2;
  for (int i = 2; i < t->d; ++i) res += x[i] * x[i];
  return -res / 2.;
}
static void user_free(void **p) { free(*p); *p = NULL; }

#define CHECK(err, what) do { if (isError(*(err))) { fprintf(stderr, "FAILED in %s\n", what); printError(stderr, *(err)); return 10; } } while (0)
static int rd(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n ? 0 : 1; }

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s in.bin out.bin mode\n", argv[0]); return 2; }
  const int mode = atoi(argv[3]);
  error *_err = initError(), **err = &_err;
  FILE *in = fopen(argv[1], "rb");
  if (!in) return 3;
  long hdr[6]; /* K, d, N, niter, nfaces, seed */
  double tp[2];
  if (rd(in, hdr, sizeof hdr) || rd(in, tp, sizeof tp)) return 4;
  const int K = (int)hdr[0], d = (int)hdr[1];
  const long N = hdr[2];
  const int niter = (int)hdr[3], nfaces = (int)hdr[4];
  double *w = malloc(sizeof(double) * K), *mean = malloc(sizeof(double) * K * d), *cov = malloc(sizeof(double) * K * d * d);
  double *faces = malloc(sizeof(double) * (nfaces > 0 ? nfaces : 1) * (d + 1));
  if (rd(in, w, sizeof(double) * K) || rd(in, mean, sizeof(double) * K * d) || rd(in, cov, sizeof(double) * K * d * d)) return 4;
  if (nfaces > 0 && rd(in, faces, sizeof(double) * nfaces * (d + 1))) return 4;
  fclose(in);

  /* objects, exactly as a pmclib user builds them (SURVEY.md appendix A driver skeleton) */
  mix_mvdens *m = mix_mvdens_alloc(K, d, err); CHECK(err, "mix_mvdens_alloc");
  for (int k = 0; k < K; ++k) {
    m->wght[k] = w[k];
    memcpy(m->comp[k]->mean, mean + (size_t)k * d, sizeof(double) * d);
    memcpy(m->comp[k]->std, cov + (size_t)k * d * d, sizeof(double) * d * d);
  }
  distribution *tgt;
  user_tgt *ut = NULL;
  if (mode == 2 || mode == 8) {
    ut = malloc(sizeof(user_tgt)); ut->d = d; ut->b = tp[0]; ut->s1 = tp[1]; ut->calls = 0;
    tgt = init_distribution(d, ut, user_lkl, user_free, NULL, err);
  } else {
    tgt = init_bentGauss_distribution(d, tp[0], tp[1], err);
  }
  CHECK(err, "target");
  if (mode == 6) { /* distribution.c:548-606: prior on coordinates 0, 3, 7 (pmclib_b200/workloads.py:banana_prior) */
    int idim[3] = {0, 3, 7};
    double loc[3] = {0.5, -1.0, 2.0}, var[3] = {25.0, 4.0, 9.0};
    tgt = add_gaussian_prior(tgt, 3, idim, loc, var, err); CHECK(err, "add_gaussian_prior");
  }
  parabox *pb = NULL;
  if (nfaces > 0) {
    pb = init_parabox(d, err); CHECK(err, "init_parabox");
    for (int f = 0; f < nfaces; ++f) { add_face(pb, faces + (size_t)f * (d + 1), faces[(size_t)f * (d + 1) + d], err); CHECK(err, "add_face"); }
  }
  const int mpi_mode = (mode == 7 || mode == 8 || mode == 9);
  pmc_simu *psim = mpi_mode ? pmc_simu_init_mpi(N, d, 0, err) : pmc_simu_init(N, d, err); CHECK(err, "pmc_simu_init");
  if (mpi_mode) {
    pmc_simu_init_classic_pmc_mpi(psim, tgt, pb, m, 0, NULL, NULL, err);
    m = (mix_mvdens *)psim->proposal->data; /* clients work on the copy rank 0 broadcast */
  } else if (mode == 4) { /* the shipped pmc_filter hook (pmc.c:203-206): histogram clipping of the log-weights */
    int *fil = filter_histogram_init(20, 1, 2, err); CHECK(err, "filter_histogram_init");
    pmc_simu_init_classic_pmc(psim, tgt, pb, m, 0, fil, &filter_histogram, err);
  } else {
    pmc_simu_init_classic_pmc(psim, tgt, pb, m, 0, NULL, NULL, err);
  }
  CHECK(err, "init_classic_pmc");
  sm_state st = {(uint64_t)hdr[5]};
  gsl_rng rng = {&sm_type, &st};

  if (mode == 3) {
    /* error chain behaviour + loud failure without a device */
    double perp = pmc_simu_pmc_step(psim, &rng, err);
    const int code = getErrorValue(*err);
    printf("{\"perp\": %g, \"is_error\": %d, \"code\": %d}\n", perp, isError(*err) ? 1 : 0, code);
    purgeError(err);
    pmc_simu_free(&psim);
    return 0;
  }

  FILE *out = fopen((mpi_mode && psim->mpi_rank != 0) ? "/dev/null" : argv[2], "wb");
  if (!out) return 5;
  for (int it = 0; it < niter; ++it) {
    double perp, ess = 0, lne = 0;
    if (mode == 1 || mode == 5) {
      psim->proposal->simulate(psim, psim->proposal->data, &rng, psim->pb, err); CHECK(err, "simulate");
      generic_get_importance_weight_and_deduced_verb(psim, psim->proposal, &distribution_lkl, &distribution_lkl,
                                                     &distribution_retrieve, psim->target, 1.0, 1, err); CHECK(err, "weights");
    } else {
      /* nothing: the step does it all */
    }
    /* snapshot of what the update will see, for the oracle: written after the step below (X, indices survive it) */
    if (mode == 5) { /* prepare_update as a separate entry point (pmc.c:1535-1579): normalises, counts, factorises, snapshots */
      size_t *count = calloc(K, sizeof(size_t));
      double *buf = NULL;
      prepare_update(m, count, psim, &buf, err); CHECK(err, "prepare_update");
      size_t tot = 0, nokf = 0;
      for (int k = 0; k < K; ++k) tot += count[k];
      for (long i = 0; i < N; ++i) nokf += psim->flg[i] == 1;
      printf("prepare_update count %zu ok %zu isLog %d snapshot %d\n", tot, nokf, psim->isLog, buf != NULL);
      if (buf && memcmp(buf, m->comp[0]->std, sizeof(double) * d * d) != 0) { fprintf(stderr, "snapshot differs\n"); return 12; }
      free(count); free(buf);
    }
    if (mode == 1 || mode == 5) {
      normalize_importance_weight(psim, err); CHECK(err, "normalize");
      psim->pmc_update(psim->proposal->data, psim, err); CHECK(err, "update");
      perp = perplexity_and_ess(psim, MC_UNORM, &ess, err); CHECK(err, "perplexity");
    } else if (mpi_mode) {
      if (mode == 9) {
        /* a hand-written master / client iteration on the message-level functions, laid out as pmc_mpi.c:44-125 */
        size_t nok, mys, all = 0;
        if (psim->mpi_rank == 0) {
          psim->proposal->simulate(psim, psim->proposal->data, &rng, psim->pb, err); CHECK(err, "simulate");
          mys = send_simulation(psim, psim->mpi_size, err); CHECK(err, "send_simulation");
          all = psim->nsamples;
          psim->nsamples = mys;
        } else {
          receive_simulation(psim, psim->mpi_size, psim->mpi_rank, err); CHECK(err, "receive_simulation");
        }
        nok = generic_get_importance_weight_and_deduced(psim, psim->proposal->data, psim->proposal->log_pdf, psim->target->log_pdf,
                                                        psim->target->retrieve, psim->target->data, err); CHECK(err, "weights");
        if (psim->mpi_rank == 0) {
          mys = psim->nsamples;
          psim->nsamples = all;
          nok = receive_importance_weight(psim, psim->mpi_size, nok, mys, err); CHECK(err, "receive_importance_weight");
          printf("message-level iteration %d nok %zu of %ld\n", it, nok, psim->nsamples);
          normalize_importance_weight(psim, err); CHECK(err, "normalize");
          psim->pmc_update(psim->proposal->data, psim, err); CHECK(err, "update");
          perp = perplexity_and_ess(psim, MC_UNORM, &ess, err); CHECK(err, "perplexity");
        } else {
          printf("rank %d slice %ld nok %zu\n", psim->mpi_rank, psim->nsamples, nok);
          send_importance_weight(psim->mpi_rank, psim->mpi_size, psim, nok);
          perp = 0;
        }
        distribution_broadcast(psim->proposal, err); CHECK(err, "distribution_broadcast");
        m = (mix_mvdens *)psim->proposal->data;
      } else {
        perp = pmc_simu_pmc_step_mpi(psim, &rng, err); CHECK(err, "pmc_simu_pmc_step_mpi");
      }
      if (psim->mpi_rank != 0) { /* clients hold no sample (pmc_mpi.c:20-34); their proposal must equal the master's: print a digest */
        double dg = 0;
        for (int k = 0; k < K; ++k) { dg += m->wght[k] * (k + 1); for (int a = 0; a < d; ++a) dg += m->comp[k]->mean[a] * (a + 1) + m->comp[k]->std[a * d + a]; }
        printf("rank %d iteration %d proposal digest %.17g\n", psim->mpi_rank, it, dg);
        continue;
      }
      ess = 0;
      double p2 = perplexity_and_ess(psim, MC_UNORM, &ess, err); CHECK(err, "perplexity_and_ess");
      if (fabs(p2 - perp) > 1e-9 * fabs(perp)) { fprintf(stderr, "perplexity mismatch %g %g\n", perp, p2); return 11; }
      { double dg = 0;
        for (int k = 0; k < K; ++k) { dg += m->wght[k] * (k + 1); for (int a = 0; a < d; ++a) dg += m->comp[k]->mean[a] * (a + 1) + m->comp[k]->std[a * d + a]; }
        printf("rank 0 iteration %d proposal digest %.17g\n", it, dg); }
    } else {
      perp = pmc_simu_pmc_step(psim, &rng, err); CHECK(err, "pmc_simu_pmc_step");
      double p2 = perplexity_and_ess(psim, MC_NORM, &ess, err); CHECK(err, "perplexity_and_ess"); /* vanilla_pmc.c:117 */
      if (fabs(p2 - perp) > 1e-9 * fabs(perp)) { fprintf(stderr, "perplexity mismatch %g %g\n", perp, p2); return 11; }
    }
    evidence(psim, &lne, err); CHECK(err, "evidence");
    /* record: scalars, sample arrays, updated proposal (factors, as the reference leaves them) */
    double sc[8] = {perp, ess, lne, psim->logSum, psim->maxW, psim->maxR, (double)psim->retry, (double)psim->isLog};
    fwrite(sc, sizeof sc, 1, out);
    fwrite(psim->X, sizeof(double), (size_t)N * d, out);
    fwrite(psim->indices, sizeof(size_t), N, out);
    fwrite(psim->weights, sizeof(double), N, out);
    fwrite(psim->log_rho, sizeof(double), N, out);
    fwrite(psim->flg, sizeof(short), N, out);
    for (int k = 0; k < K; ++k) fwrite(&m->wght[k], sizeof(double), 1, out);
    for (int k = 0; k < K; ++k) fwrite(m->comp[k]->mean, sizeof(double), d, out);
    for (int k = 0; k < K; ++k) fwrite(m->comp[k]->std, sizeof(double), (size_t)d * d, out);
    printf("iteration %d perplexity %g effectiveSampleSize %g evidence %g\n", it, perp, ess, exp(lne));
  }
  if (mode == 4) { /* post-processing either side of the path: ASCII dump and per-coordinate mean */
    char name[4096];
    snprintf(name, sizeof name, "%s.txt", argv[2]);
    out_pmc_simu(name, psim, err); CHECK(err, "out_pmc_simu");
    for (int a = 0; a < d; ++a) printf("mean %d %.17g\n", a, mean_from_psim(psim->X, psim->weights, psim->flg, (int)N, d, a));
  }
  if (ut) printf("user callback calls %ld\n", ut->calls);
  fclose(out);
  pmc_simu_free(&psim); /* frees target, proposal (and the mixture through it) */
  endError(err);
  free(w); free(mean); free(cov); free(faces);
  return 0;
}
