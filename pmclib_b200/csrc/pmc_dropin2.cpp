// pmc_dropin2.cpp — the rest of pmclib's public API around the PMC path (include/pmclib_abi.h, second half):
// single-sample draws (mvdens_ran, mix_mvdens_ran), the wire format of a mixture (serialize/deserialize, mix_mvdens_size),
// the text readers (mvdens_dwnp & co.), the resume path (pmc_simu_from_file, fill_read_pmc_simu), parameter names and the
// combined-target machinery of distribution.c (combine_distribution_*, add_gaussian_prior*), small helpers.
//
// Same rules as pmc_dropin.cpp: reference names, signatures, struct layouts and error chain; host code here is object
// management, file parsing and O(d^3) linear algebra; whatever loops over samples goes to the kernels (pmcgpu_*).
// Citations are file:line in CosmoStat/pmclib.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>
#include "../../include/pmcb200.h"
#include "../../include/pmclib_abi.h"
#include "pmc_dropin_internal.h"

namespace {
constexpr int mv_base = -700, mv_negWeight = mv_base - 5, mv_negative = mv_base - 7, mv_serialize = mv_base - 9, mv_io = mv_base - 10;
constexpr int dist_undef = -6701, dist_type = -6702, err_allocate = -102, pmc_io = -6010, tls_file = -4010, tls_overflow = -4011,
              tls_Nparam = -4012;

#define ERR2(e, code, ...) e = newErrorVA(code, __func__, (char *)"(pmc_dropin2.cpp)", (char *)__VA_ARGS__)
#define FWD2(e, ret)                                                            \
  do {                                                                          \
    if (isError(e)) {                                                           \
      e = newError(forwardErr, (char *)__func__, (char *)"", nullptr, e);       \
      return ret;                                                               \
    }                                                                           \
  } while (0)

void *xalloc(size_t n, error **err) {
  void *p = malloc(n ? n : 1);
  if (!p) *err = newErrorVA(err_allocate, __func__, (char *)"", (char *)"Cannot allocate %ld bytes", *err, nullptr, (long)n);
  return p;
}

// one sample from a K-component mixture through the GPU draw kernels (Philox keyed from the caller's gsl_rng)
bool gpu_draw_one(int K, int d, const double *w, const double *mean, const double *L, const double *detL, const int *df,
                  gsl_rng *r, double *x, size_t *index, error **err) {
  std::lock_guard<std::mutex> lk(pmcb200::dropin_point_mutex());
  pmcgpu_ctx *c = pmcb200::dropin_point_ctx(err);
  if (!c) return false;
  pmcb200::dropin_point_invalidate();
  int rc = pmcgpu_set_proposal(c, K, d, w, mean, L, detL, df);
  if (!rc) rc = pmcgpu_set_box(c, d, 0, nullptr);
  if (!rc) rc = pmcgpu_resize(c, 1, d);
  long nrej = 0;
  if (!rc) rc = pmcgpu_draw(c, pmcb200::dropin_seed_from(r), 0, 0, &nrej);
  size_t idx = 0;
  if (!rc) rc = pmcgpu_download(c, x, &idx, nullptr, nullptr, nullptr);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return false; }
  if (index) *index = idx;
  return true;
}

// next whitespace-separated double of a line (tools.c: read_double); returns false at the end of the line
bool next_double(char **p, double *v) {
  char *end = nullptr;
  const double x = strtod(*p, &end);
  if (end == *p) return false;
  *v = x;
  *p = end;
  return true;
}
}  // namespace

extern "C" {

// ====================================================================================================================
// mvdens.c: text readers, small helpers
// ====================================================================================================================
mvdens *mvdens_dwnp(FILE *where, error **err) {  // mvdens.c:146-188
  FILE *f = where ? where : stdin;
  size_t ndim = 0, bdl = 0;
  int df = 0, chol = 0;
  if (fscanf(f, "%zu %d %zu %d\n", &ndim, &df, &bdl, &chol) != 4) {
    ERR2(*err, mv_io, "Error while reading mvdens file: Four integers \n(ndim, dof, bandlimit, cholesky flag) expected on first line", *err, nullptr);
    return nullptr;
  }
  mvdens *g = mvdens_alloc(ndim, err);
  FWD2(*err, nullptr);
  for (size_t j = 0; j < ndim; ++j)
    if (fscanf(f, "%lg", g->mean + j) != 1) { ERR2(*err, mv_io, "Cannot read mvdens (mean element #%d) double expected", *err, nullptr, (int)j); return nullptr; }
  for (size_t j = 0; j < ndim; ++j)
    for (size_t k = 0; k < ndim; ++k)
      if (fscanf(f, "%lg", g->std + j * ndim + k) != 1) {
        ERR2(*err, mv_io, "Cannot read mvdens (covariance element #(%d,%d)) double expected", *err, nullptr, (int)j, (int)k);
        return nullptr;
      }
  g->band_limit = bdl;
  g->df = df;
  g->chol = chol;
  if (chol == 1) g->detL = determinant(g->std, g->ndim);
  return g;
}
mvdens *mvdens_read_and_chol(FILE *where, error **err) {  // mvdens.c:190-204
  mvdens *g = mvdens_dwnp(where, err);
  FWD2(*err, nullptr);
  if (g->chol == 0) { mvdens_cholesky_decomp(g, err); FWD2(*err, nullptr); }
  return g;
}
void mvdens_chdump(const char *name, mvdens *what, error **err) {  // mvdens.c:136-144
  FILE *f = fopen_err(name, "w", err);
  FWD2(*err, );
  mvdens_dump(f, what);
  fclose(f);
}
void mvdens_from_meanvar(mvdens *m, const double *pmean, const double *pvar, double correct) {  // mvdens.c:62-71
  const size_t d = m->ndim;
  for (size_t j = 0; j < d; ++j) {
    for (size_t k = 0; k < d; ++k) m->std[j * d + k] = pvar[j * d + k] * correct;
    if (pmean) m->mean[j] = pmean[j];
  }
}
double ln_determinant(const double *L, size_t ndim, error **err) {  // mvdens.c:215-227
  double s = 0.0;
  for (size_t i = 0; i < ndim * ndim; i += ndim + 1) {
    if (L[i] <= 0) {
      ERR2(*err, mv_negative, "Negative diagonal entry L_{%d,%d} = %g, ln det not defined", *err, nullptr, (int)(i / (ndim + 1)), (int)(i / (ndim + 1)), L[i]);
      return 0.0;
    }
    s += std::log(L[i]);
  }
  return s;
}
// mvdens.c:393-410: the covariance is replaced by its inverse (both triangles), chol = 0; the return value is the product
// of the diagonal of the result, as the reference computes it (determinant() applied to the inverted matrix)
double mvdens_inverse(mvdens *m, error **err) {
  mvdens_cholesky_decomp(m, err);
  FWD2(*err, 0.0);
  const size_t d = m->ndim;
  std::vector<double> Li(d * d, 0.0);
  const double *L = m->std;
  for (size_t c = 0; c < d; ++c) {  // columns of L^-1 by forward substitution
    Li[c * d + c] = 1.0 / L[c * d + c];
    for (size_t i = c + 1; i < d; ++i) {
      double s = 0.0;
      for (size_t j = c; j < i; ++j) s -= L[i * d + j] * Li[j * d + c];
      Li[i * d + c] = s / L[i * d + i];
    }
  }
  for (size_t a = 0; a < d; ++a)
    for (size_t b = a; b < d; ++b) {  // Sigma^-1 = L^-T L^-1
      double s = 0.0;
      for (size_t k = b; k < d; ++k) s += Li[k * d + a] * Li[k * d + b];
      m->std[a * d + b] = m->std[b * d + a] = s;
    }
  m->chol = 0;
  return determinant(m->std, d);
}

// mvdens.c:279-317: one draw x = mu + L g (times sqrt(df/chi2) for a Student-t component); dest == NULL allocates
double *mvdens_ran(double *dest, mvdens *g, gsl_rng *r, error **err) {
  mvdens_cholesky_decomp(g, err);
  FWD2(*err, nullptr);
  double *res = dest ? dest : (double *)xalloc(sizeof(double) * g->ndim, err);
  FWD2(*err, nullptr);
  const double one = 1.0;
  if (!gpu_draw_one(1, (int)g->ndim, &one, g->mean, g->std, &g->detL, &g->df, r, res, nullptr, err)) {
    if (!dest) free(res);
    return nullptr;
  }
  return res;
}

// mvdens.c:730-764: lazy normalisation of the weights and cumulative table on the host struct, then component pick
// and draw on the device
double *mix_mvdens_ran(double *dest, size_t *index, mix_mvdens *m, gsl_rng *r, error **err) {
  if (m->init_cwght == 0) {
    double sum = 0.0;
    for (size_t i = 0; i < m->ncomp; ++i) sum += m->wght[i];
    if (std::fabs(sum - 1.0) > 1e-6) {
      if (sum == 0) { ERR2(*err, mv_negWeight, "All weights to zero !! cannot sample", *err, nullptr); return nullptr; }
      for (size_t i = 0; i < m->ncomp; ++i) m->wght[i] *= 1.0 / sum;
    }
    m->cwght[0] = m->wght[0];
    for (size_t i = 1; i < m->ncomp; ++i) m->cwght[i] = m->cwght[i - 1] + m->wght[i];
    m->init_cwght = 1;
  }
  mix_mvdens_cholesky_decomp(m, err);
  FWD2(*err, nullptr);
  const size_t K = m->ncomp, d = m->ndim;
  std::vector<double> mean(K * d), L(K * d * d), det(K);
  std::vector<int> df(K);
  for (size_t k = 0; k < K; ++k) {
    memcpy(&mean[k * d], m->comp[k]->mean, d * sizeof(double));
    memcpy(&L[k * d * d], m->comp[k]->std, d * d * sizeof(double));
    det[k] = m->comp[k]->detL;
    df[k] = m->comp[k]->df;
  }
  double *res = dest ? dest : (double *)xalloc(sizeof(double) * d, err);
  FWD2(*err, nullptr);
  size_t idx = 0;
  if (!gpu_draw_one((int)K, (int)d, m->wght, mean.data(), L.data(), det.data(), df.data(), r, res, &idx, err)) {
    if (!dest) free(res);
    return nullptr;
  }
  if (index) *index = idx;
  return res;
}

// mvdens.c:418-431: the reference copies weights and the component POINTERS (a shallow copy); so do we
void mix_mvdens_copy(mix_mvdens *target, const mix_mvdens *source, error **err) {
  (void)err;
  target->ncomp = source->ncomp;
  target->ndim = source->ndim;
  target->init_cwght = source->init_cwght;
  memcpy(target->wght, source->wght, sizeof(double) * source->ncomp);
  memcpy(target->cwght, source->cwght, sizeof(double) * source->ncomp);
  memcpy(target->comp, source->comp, sizeof(mvdens *) * source->ncomp);
}
double effective_number_of_components(const mix_mvdens *self, error **err) {  // mvdens.c:500-511
  double enc = 0.0;
  for (size_t i = 0; i < self->ncomp; ++i) enc += self->wght[i] * self->wght[i];
  if (enc < 1.0e-6) { ERR2(*err, mv_negWeight, "All mixture weights are zero", *err, nullptr); return 0.0; }
  return 1.0 / enc;
}

// ---- wire format (mvdens.c:585-720): K, d (size_t) | wght[K] | per component df (int), band_limit (size_t), mean[d],
// std[d*d] (always factorised first) ---------------------------------------------------------------------------------
size_t mix_mvdens_size(size_t ncomp, size_t ndim) {
  return sizeof(size_t) * 2 + ncomp * (sizeof(double) + sizeof(int) + sizeof(size_t)) + ncomp * ndim * (1 + ndim) * sizeof(double);
}
void *serialize_mix_mvdens(mix_mvdens *self, size_t *sz, error **err) {
  const size_t K = self->ncomp, d = self->ndim, total = mix_mvdens_size(K, d);
  char *buf = (char *)xalloc(total, err);
  FWD2(*err, nullptr);
  char *p = buf;
  auto put = [&p](const void *src, size_t n) { memcpy(p, src, n); p += n; };
  put(&K, sizeof K);
  put(&d, sizeof d);
  put(self->wght, sizeof(double) * K);
  for (size_t k = 0; k < K; ++k) {
    mvdens *c = self->comp[k];
    put(&c->df, sizeof(int));
    put(&c->band_limit, sizeof(size_t));
    mvdens_cholesky_decomp(c, err);
    if (isError(*err)) { free(buf); FWD2(*err, nullptr); }
    put(c->mean, sizeof(double) * d);
    put(c->std, sizeof(double) * d * d);
  }
  if ((size_t)(p - buf) != total) {
    ERR2(*err, mv_serialize, "Bad size for serialization (expected %d, got %d)", *err, nullptr, (int)total, (int)(p - buf));
    free(buf);
    return nullptr;
  }
  *sz = total;
  return buf;
}
mix_mvdens *deserialize_mix_mvdens(void *serialized, size_t sz, error **err) {
  const char *p = (const char *)serialized;
  size_t K = 0, d = 0;
  auto get = [&p](void *dst, size_t n) { memcpy(dst, p, n); p += n; };
  get(&K, sizeof K);
  get(&d, sizeof d);
  const size_t total = mix_mvdens_size(K, d);
  if (total != sz) { ERR2(*err, mv_serialize, "Cannot deserialize, bad size (Expected %d got %d)", *err, nullptr, (int)sz, (int)total); return nullptr; }
  mix_mvdens *m = mix_mvdens_alloc(K, d, err);
  FWD2(*err, nullptr);
  get(m->wght, sizeof(double) * K);
  for (size_t k = 0; k < K; ++k) {
    mvdens *c = m->comp[k];
    get(&c->df, sizeof(int));
    get(&c->band_limit, sizeof(size_t));
    get(c->mean, sizeof(double) * d);
    get(c->std, sizeof(double) * d * d);
    c->chol = 1;  // serialize always ships factors
    c->detL = determinant(c->std, d);
  }
  return m;
}

// ====================================================================================================================
// pmc.c: resume path and index sort
// ====================================================================================================================
// pmc.c:400-445: after a sample has been read back: every sample flagged ok, log_rho from the proposal (device, all
// samples in one launch), running maxima, then the normalisation.  Returns what normalize_importance_weight returns.
double fill_read_pmc_simu(pmc_simu *psim, mix_mvdens *proposal, error **err) {
  const long N = psim->nsamples;
  double MR = -1.0e30, MW = -1.0e30;
  for (long i = 0; i < N; ++i) psim->flg[i] = 1;
  if (proposal != nullptr && N > 0) {
    mix_mvdens_cholesky_decomp(proposal, err);
    FWD2(*err, 0.0);
    const size_t K = proposal->ncomp, d = proposal->ndim;
    std::vector<double> mean(K * d), L(K * d * d), det(K);
    std::vector<int> df(K);
    for (size_t k = 0; k < K; ++k) {
      memcpy(&mean[k * d], proposal->comp[k]->mean, d * sizeof(double));
      memcpy(&L[k * d * d], proposal->comp[k]->std, d * d * sizeof(double));
      det[k] = proposal->comp[k]->detL;
      df[k] = proposal->comp[k]->df;
    }
    std::lock_guard<std::mutex> lk(pmcb200::dropin_point_mutex());
    pmcgpu_ctx *c = pmcb200::dropin_point_ctx(err);
    if (!c) return 0.0;
    pmcb200::dropin_point_invalidate();
    int rc = pmcgpu_set_proposal(c, (int)K, (int)d, proposal->wght, mean.data(), L.data(), det.data(), df.data());
    if (!rc) rc = pmcgpu_proposal_log_pdf(c, N, psim->X, psim->log_rho);
    if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0.0; }
    for (long i = 0; i < N; ++i)
      if (psim->log_rho[i] > MR) MR = psim->log_rho[i];
  }
  for (long i = 0; i < N; ++i)
    if (psim->weights[i] > MW) MW = psim->weights[i];  // the file holds log-weights (pmc.c:423-428)
  psim->maxW = MW;
  psim->maxR = MR;
  psim->isLog = 1;
  const double s = normalize_importance_weight(psim, err);
  FWD2(*err, 0.0);
  return s;
}

// pmc.c:373-396 + tools.c:140-290 (read_mkmc_chain / read_header_lc / read_step_lc): "lc" text format, one sample per line:
// log-weight, -component, the npar parameters, the n_ded deduced parameters; '#' lines are headers, of which
// "# npar = %d, n_ded = %d" is checked against the arguments
pmc_simu *pmc_simu_from_file(FILE *PMCSIM, int nsamples, int npar, int n_ded, mix_mvdens *proposal, int nclipw, error **err) {
  pmc_simu *psim = pmc_simu_init_plus_ded(nsamples, npar, n_ded, err);
  FWD2(*err, nullptr);
  long n = 0;
  char *line = nullptr;
  size_t cap = 0;
  while (getline(&line, &cap, PMCSIM) != -1) {
    if (line[0] == '#') {
      int NP = 0, ND = 0;
      if (sscanf(line, "# npar = %d, n_ded = %d\n", &NP, &ND) == 2) {
        if (NP != npar) { ERR2(*err, tls_Nparam, "Chain/sample header (npar=%d) not compatible with config file (npar=%d)", *err, nullptr, NP, npar); break; }
        if (ND != n_ded) { ERR2(*err, tls_Nparam, "Chain/sample header (n_ded=%d) not compatible with config file (n_ded=%d)", *err, nullptr, ND, n_ded); break; }
      }
      continue;
    }
    if (n >= nsamples) {
      ERR2(*err, tls_overflow, "Chain/sample length is larger than expected (n=%d). For pmc: check nsamples, or nsamples*fsfinal in case of last iteration.", *err, nullptr, nsamples);
      break;
    }
    char *p = line;
    double lw = 0, negk = 0;
    if (!next_double(&p, &lw)) { ERR2(*err, tls_file, "Bad line: First entry (mcmc:accept, pmc:weight) not a double.", *err, nullptr); break; }
    if (!next_double(&p, &negk)) { ERR2(*err, tls_file, "Bad line: Second entry (mcmc:logL, pmc:icomp) not a double", *err, nullptr); break; }
    bool bad = false;
    for (int j = 0; j < npar && !bad; ++j) bad = !next_double(&p, psim->X + (size_t)n * npar + j);
    if (bad) { ERR2(*err, tls_file, "Bad line: Parameter entry not a double", *err, nullptr); break; }
    for (int j = 0; j < n_ded && !bad; ++j) bad = !next_double(&p, psim->X_ded + (size_t)n * n_ded + j);
    if (bad) { ERR2(*err, tls_file, "Bad line: Deduced parameter not a double", *err, nullptr); break; }
    psim->weights[n] = lw;
    psim->indices[n] = (size_t)(-negk);
    ++n;
  }
  free(line);
  if (isError(*err)) { pmc_simu_free(&psim); return nullptr; }
  psim->nsamples = n;  // read_mkmc_chain reports the number of lines actually read
  fill_read_pmc_simu(psim, proposal, err);
  if (isError(*err)) { FWD2(*err, nullptr); }
  if (nclipw > 0) { clip_weights(psim, nclipw, stderr, err); FWD2(*err, nullptr); }
  return psim;
}

// pmc.c:829-873: an unfinished function in the reference — it sorts a copy of the log-weights, computes nothing from the
// order and leaves the pmc_simu untouched (the median / variance code is commented out).  Exported so that callers that
// reference pmc.h:174 link; observable behaviour (the warning for normalised weights, no change of psim) is kept.
void filter_importance_weight(pmc_simu *psim, double nsig, error **err) {
  (void)nsig; (void)err;
  if (psim->isLog != 1) fprintf(stderr, "Filtering normalized importance weights... It is probably too late...\n");
}

// pmc.c:772-827 (not declared in pmc.h, but a global symbol of libpmc): weights back to log scale with outliers flagged out.
// Normalises first if needed (GPU), then on the host arrays: mean and standard deviation of the usable weights, every usable
// sample with |w - mean| > nsig is flagged out and reported on stderr, the others get w <- log w; maxW = the largest of them
// (with the reference's unconditional assignment at i == 0), isLog = 1.  The reference compares with nsig itself, not with
// nsig standard deviations (the scaled deviation is computed and never used): kept.
void filter_importance_weight_sig(pmc_simu *psim, double nsig, error **err) {
  if (psim->isLog == 1) {
    normalize_importance_weight(psim, err);
    if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return; }
  }
  double mean = 0.0, max = 0.0;
  size_t cnt = 0;
  for (long i = 0; i < psim->nsamples; ++i)
    if (psim->flg[i] != 0) { mean += psim->weights[i]; ++cnt; }
  mean *= 1. / cnt;
  for (long i = 0; i < psim->nsamples; ++i) {
    if (psim->flg[i] == 0) continue;
    const double dev = fabs(psim->weights[i] - mean);
    if (dev > nsig) {
      psim->flg[i] = 0;
      fprintf(stderr, "Flagging %zu %g %g\n ", (size_t)i, dev, nsig);
      fprintf(stderr, "%1d %10g", 0, psim->weights[i]);  // print_step (tools.c:301-313)
      for (int a = 0; a < psim->ndim; ++a) fprintf(stderr, " %10g", psim->X[(size_t)i * psim->ndim + a]);
      fprintf(stderr, "\n");
      continue;
    }
    psim->weights[i] = log(psim->weights[i]);
    if (i == 0 || psim->weights[i] > max) max = psim->weights[i];
  }
  psim->isLog = 1;
  psim->maxW = max;
}

// pmc.c:881-1034: permutation that sorts list (order > 0 ascending, else descending).  in_ind == NULL allocates the
// identity first.  Ties keep their input order (the reference's quick-sort leaves them in an unspecified order).
size_t *sort_indices(double *list, size_t size, size_t *in_ind, size_t *buf, int order, error **err) {
  (void)buf;
  size_t *ind = in_ind;
  if (!ind) {
    ind = (size_t *)xalloc(sizeof(size_t) * size, err);
    FWD2(*err, nullptr);
    for (size_t i = 0; i < size; ++i) ind[i] = i;
  }
  if (order > 0) std::stable_sort(ind, ind + size, [list](size_t a, size_t b) { return list[a] < list[b]; });
  else std::stable_sort(ind, ind + size, [list](size_t a, size_t b) { return list[a] > list[b]; });
  return ind;
}

// ====================================================================================================================
// distribution.c: names, broadcast hook, combined targets, Gaussian priors
// ====================================================================================================================
void distribution_set_broadcast(distribution *dist, mpi_exchange_func *broadcast, error **err) {  // distribution.c:70-74
  if (!dist) { ERR2(*err, -1, "Bad distribution", *err, nullptr); return; }
  dist->broadcast_mpi = broadcast;
}
void distribution_set_names(distribution *dist, char **name, error **err) {  // distribution.c:175-191
  if (dist->ndef != 0) { ERR2(*err, dist_undef, "cannot set name after having set default parameters", *err, nullptr); return; }
  if (dist->name) free(dist->name);
  const int n = dist->ndim + dist->n_ded;
  dist->name = (_char_name *)xalloc(sizeof(_char_name) * (n > 0 ? n : 1), err);
  FWD2(*err, );
  for (int i = 0; i < n; ++i) snprintf(dist->name[i], sizeof(_char_name), "%s", name[i]);
}
// distribution.c:129-156: index among the FREE parameters, or -1-k for deduced parameter k
int distribution_get_name(distribution *dist, char *name, error **err) {
  if (!dist->name) { ERR2(*err, dist_undef, "The distribution has no parameter name defined", *err, nullptr); return -1; }
  int j = 0;
  for (int i = 0; i < dist->ndim + dist->ndef; ++i) {
    if (dist->ndef > 0 && dist->def[i] == 1) continue;
    if (strcmp(name, dist->name[i]) == 0) return j;
    ++j;
  }
  j = -1;
  for (int i = 0; i < dist->n_ded; ++i) {  // the reference compares with name[i], not name[ndim+i]; kept
    if (strcmp(name, dist->name[i]) == 0) return j;
    --j;
  }
  ERR2(*err, dist_undef, "Cannot find parameter %s", *err, nullptr, name);
  return -1;
}
int *distribution_get_names(distribution *dist, int nnames, char **name, int includeded, error **err) {  // :158-173
  int *idef = (int *)xalloc(sizeof(int) * (nnames > 0 ? nnames : 1), err);
  FWD2(*err, nullptr);
  for (int i = 0; i < nnames; ++i) {
    idef[i] = distribution_get_name(dist, name[i], err);
    if (isError(*err)) { free(idef); FWD2(*err, nullptr); }
    if (idef[i] < 0 && includeded == 0) { ERR2(*err, dist_undef, "Asking for deduced parameter (%s)", *err, nullptr, name[i]); free(idef); return nullptr; }
  }
  return idef;
}
void distribution_set_default_name(distribution *dist, int ndef, char **namedef, double *vdef, error **err) {  // :119-128
  int *idef = distribution_get_names(dist, ndef, namedef, 0, err);
  FWD2(*err, );
  distribution_set_default(dist, ndef, idef, vdef, err);
  free(idef);
  FWD2(*err, );
}

// ---- combine (distribution.c:261-546): sum of log-densities, each on an index-mapped sub-vector ----------------------
distribution *combine_distribution_init(int ndim, int nded, error **err) {
  comb_dist_data *cbd = (comb_dist_data *)xalloc(sizeof(comb_dist_data), err);
  FWD2(*err, nullptr);
  memset(cbd, 0, sizeof *cbd);
  cbd->ndim = ndim;
  cbd->nded = nded;
  const size_t n = (size_t)(ndim + nded);
  cbd->pars = (double *)xalloc(sizeof(double) * n, err);
  cbd->pded = (double *)xalloc(sizeof(double) * n, err);
  cbd->dummy = (double *)xalloc(sizeof(double) * n, err);
  FWD2(*err, nullptr);
  cbd->ndummy = ndim + nded;
  if (nded > 0) {
    cbd->ded_from = (int *)xalloc(sizeof(int) * nded, err);
    FWD2(*err, nullptr);
    for (int i = 0; i < nded; ++i) cbd->ded_from[i] = -1;
  }
  distribution *self = init_distribution_full(ndim, cbd, &combine_lkl, &combine_free, nullptr, nded, &combine_retrieve, err);
  FWD2(*err, nullptr);
  return self;
}
distribution *combine_distribution_simple_init(int ndim, error **err) {
  distribution *c = combine_distribution_init(ndim, 0, err);
  FWD2(*err, nullptr);
  return c;
}
void combine_free(void **pcbd) {
  comb_dist_data *cbd = (comb_dist_data *)*pcbd;
  free(cbd->pars);
  free(cbd->pded);
  free(cbd->dummy);
  free(cbd->ded_from);
  if (cbd->dist) {
    for (int id = 0; id < cbd->ndist; ++id) {
      free_distribution(&cbd->dist[id]);
      free(cbd->dim[id]);
      if (cbd->ded && cbd->ded[id]) free(cbd->ded[id]);
    }
    free(cbd->dist);
    free(cbd->dim);
    free(cbd->ded);
  }
  free(cbd);
  *pcbd = nullptr;
}
// distribution.c:350-396, point-wise on the host: every term goes through distribution_lkl (i.e. through the device
// entry points for the densities this library knows).  Bulk evaluation inside the PMC iteration does not come here:
// a combine of recognised densities is evaluated term by term on the device (pmcgpu_target_combine_*).
double combine_lkl(void *pcbd, const double *pars, error **err) {
  comb_dist_data *cbd = (comb_dist_data *)pcbd;
  double res = 0;
  for (int id = 0; id < cbd->ndist; ++id) {
    distribution *t = cbd->dist[id];
    for (int ip = 0; ip < t->ndim; ++ip) {
      const int ni = cbd->dim[id][ip];
      cbd->pars[ip] = ni >= 0 ? pars[ni] : cbd->pded[-ni - 1];
    }
    res += distribution_lkl(t, cbd->pars, err);
    FWD2(*err, 0);
    if (t->n_ded != 0) {
      distribution_retrieve(t, cbd->dummy, err);
      FWD2(*err, 0);
      for (int ip = 0; ip < t->n_ded; ++ip) {
        const int ni = cbd->dim[id][ip];  // the reference indexes dim here, not ded (distribution.c:387); kept
        if (ni >= 0) cbd->pded[ni] = cbd->dummy[ip];
      }
    }
  }
  return res;
}
void combine_retrieve(const void *pcbd, double *pded, error **err) {
  (void)err;
  const comb_dist_data *cbd = (const comb_dist_data *)pcbd;
  memcpy(pded, cbd->pded, sizeof(double) * (size_t)cbd->nded);
}
void add_to_combine_distribution(distribution *comb, distribution *addon, int *dim_idx, int *ded_idx, error **err) {
  if (comb->log_pdf != &combine_lkl) { ERR2(*err, dist_type, "Bad distribution", *err, nullptr); return; }
  comb_dist_data *cbd = (comb_dist_data *)comb->data;
  if (dim_idx) {
    for (int i = 0; i < addon->ndim; ++i) {
      const int dix = dim_idx[i];
      if (dix >= cbd->ndim) { ERR2(*err, dist_type, "addon need parameter %d, but combine_lkl only has %d dims", *err, nullptr, dix, cbd->ndim); return; }
      if (dix < 0) {
        if (-dix - 1 >= cbd->nded) { ERR2(*err, dist_type, "addon need parameter %d, but combine_lkl only has %d ded", *err, nullptr, -dix - 1, cbd->nded); return; }
        if (cbd->ded_from[-dix - 1] == -1) { ERR2(*err, dist_type, "addon need parameter %d, but combine_lkl has no way to compute it", *err, nullptr, -dix - 1); return; }
      }
    }
  }
  if (ded_idx) {
    for (int i = 0; i < addon->n_ded; ++i)
      if (ded_idx[i] >= 0 && cbd->ded_from[ded_idx[i]] != -1) {
        ERR2(*err, dist_type, "addon want to provide ded %d which is already provided by addon %d", *err, nullptr, ded_idx[i], cbd->ded_from[ded_idx[i]]);
        return;
      }
  }
  if (cbd->ndist % 10 == 0) {  // room for ten more terms
    const size_t n = (size_t)cbd->ndist + 10;
    distribution **nd = (distribution **)realloc(cbd->dist, sizeof(distribution *) * n);
    int **ndim = (int **)realloc(cbd->dim, sizeof(int *) * n);
    int **nded = (int **)realloc(cbd->ded, sizeof(int *) * n);
    if (!nd || !ndim || !nded) { ERR2(*err, err_allocate, "Cannot allocate", *err, nullptr); return; }
    cbd->dist = nd;
    cbd->dim = ndim;
    cbd->ded = nded;
  }
  const int id = cbd->ndist;
  cbd->dist[id] = addon;
  cbd->dim[id] = (int *)xalloc(sizeof(int) * (addon->ndim > 0 ? addon->ndim : 1), err);
  FWD2(*err, );
  for (int i = 0; i < addon->ndim; ++i) cbd->dim[id][i] = dim_idx ? dim_idx[i] : i;
  cbd->ded[id] = nullptr;
  if (addon->n_ded > 0) {
    cbd->ded[id] = (int *)xalloc(sizeof(int) * addon->n_ded, err);
    FWD2(*err, );
    for (int i = 0; i < addon->n_ded; ++i) cbd->ded[id][i] = ded_idx ? ded_idx[i] : i;
    if (addon->n_ded > cbd->ndummy) {
      free(cbd->dummy);
      cbd->ndummy = addon->n_ded;
      cbd->dummy = (double *)xalloc(sizeof(double) * cbd->ndummy, err);
      FWD2(*err, );
    }
    for (int i = 0; i < addon->n_ded; ++i)
      if (cbd->ded[id][i] >= 0 && cbd->ded_from) cbd->ded_from[cbd->ded[id][i]] = id;
  }
  cbd->ndist++;
}
void add_to_combine_distribution_name(distribution *comb, distribution *addon, error **err) {  // distribution.c:405-441
  if (!comb->name || !addon->name) { ERR2(*err, dist_undef, "names undefined", *err, nullptr); return; }
  std::vector<int> dim_idx(addon->ndim > 0 ? addon->ndim : 1), ded_idx(addon->n_ded > 0 ? addon->n_ded : 1);
  int j = 0;
  for (int i = 0; i < addon->ndim + addon->ndef; ++i) {
    if (addon->ndef != 0 && addon->def[i] == 1) continue;
    dim_idx[j++] = distribution_get_name(comb, addon->name[i], err);
    FWD2(*err, );
  }
  for (int i = 0; i < addon->n_ded; ++i) {
    int v = distribution_get_name(comb, addon->name[i + addon->ndim + addon->ndef], err);
    FWD2(*err, );
    if (v >= 0) { ERR2(*err, dist_undef, "parameter %s is not deduced", *err, nullptr, addon->name[i + addon->ndim + addon->ndef]); return; }
    ded_idx[i] = -v - 1;
  }
  add_to_combine_distribution(comb, addon, dim_idx.data(), addon->n_ded ? ded_idx.data() : nullptr, err);
  FWD2(*err, );
}
// distribution.c:548-606: the prior is an mvdens on the listed coordinates, added to a combine (created around orig if needed)
distribution *add_gaussian_prior_2(distribution *orig, int ndim, int *idim, double *loc, double *var, error **err) {
  mvdens *mv = mvdens_alloc(ndim, err);
  FWD2(*err, nullptr);
  memcpy(mv->mean, loc, sizeof(double) * ndim);
  memcpy(mv->std, var, sizeof(double) * ndim * ndim);
  distribution *dmv = init_distribution(ndim, mv, &mvdens_log_pdf_void, &mvdens_free_void, nullptr, err);
  FWD2(*err, nullptr);
  distribution *dres = orig;
  if (orig->log_pdf != &combine_lkl) {
    dres = combine_distribution_init(orig->ndim, orig->n_ded, err);
    FWD2(*err, nullptr);
    add_to_combine_distribution(dres, orig, nullptr, nullptr, err);
    FWD2(*err, nullptr);
  }
  add_to_combine_distribution(dres, dmv, idim, nullptr, err);
  FWD2(*err, nullptr);
  return dres;
}
distribution *add_gaussian_prior(distribution *orig, int ndim, int *idim, double *loc, double *var, error **err) {
  std::vector<double> vv((size_t)ndim * ndim, 0.0);
  for (int i = 0; i < ndim; ++i) vv[(size_t)i * ndim + i] = var[i];
  distribution *d = add_gaussian_prior_2(orig, ndim, idim, loc, vv.data(), err);
  FWD2(*err, nullptr);
  return d;
}
distribution *add_gaussian_prior_name(distribution *orig, int ndim, char **iname, double *loc, double *var, error **err) {
  int *idim = distribution_get_names(orig, ndim, iname, 0, err);
  FWD2(*err, nullptr);
  distribution *d = add_gaussian_prior(orig, ndim, idim, loc, var, err);
  free(idim);
  FWD2(*err, nullptr);
  return d;
}
distribution *add_gaussian_prior_2_name(distribution *orig, int ndim, char **iname, double *loc, double *var, error **err) {
  int *idim = distribution_get_names(orig, ndim, iname, 0, err);
  FWD2(*err, nullptr);
  distribution *d = add_gaussian_prior_2(orig, ndim, idim, loc, var, err);
  free(idim);
  FWD2(*err, nullptr);
  return d;
}

// ====================================================================================================================
// parabox.c / errorlist.c leftovers
// ====================================================================================================================
void print_parabox(parabox *bx, FILE *mstream) {  // parabox.c:147-186
  FILE *out = mstream ? mstream : stdout;
  const int dim = bx->ndim;
  auto rule = [&]() { for (int j = 0; j < dim; ++j) fprintf(out, "----------"); fprintf(out, "------------------\n"); };
  rule();
  fprintf(out, "Face | ");
  for (int j = 0; j < dim; ++j) fprintf(out, "w_%i       ", j);
  fprintf(out, "| threshold \n");
  rule();
  for (int i = 0; i < bx->nfaces; ++i) {
    const double *c = bx->faces + (size_t)i * (dim + 1);
    fprintf(out, "%4i | ", i);
    for (int j = 0; j < dim; ++j) fprintf(out, "% .2e ", c[j]);
    fprintf(out, "| % .2e\n", c[dim]);
  }
  rule();
}
int pickleError(error *err, void **buf) {  // errorlist.c:126-150: the chain as a flat array of nodes
  int len = 0;
  for (error *c = err; c; c = c->next) ++len;
  *buf = malloc(sizeof(error) * (len > 0 ? len : 1));
  char *p = (char *)*buf;
  for (error *c = err; c; c = c->next) { memcpy(p, c, sizeof(error)); p += sizeof(error); }
  return len;
}
error *unpickleError(void *buf, int len) {  // errorlist.c:152-169
  error *first = nullptr, *prev = nullptr;
  for (int i = 0; i < len; ++i) {
    error *cur = (error *)malloc(sizeof(error));
    memcpy(cur, (char *)buf + sizeof(error) * i, sizeof(error));
    cur->next = nullptr;
    if (prev) prev->next = cur; else first = cur;
    prev = cur;
  }
  return first;
}

}  // extern "C"
