// pmc_dropin.cpp — pmclib's own C API (include/pmclib_abi.h) on top of the pmcgpu_* C-ABI.
//
// Host orchestration only: object management with the reference's struct layouts and ownership rules, the reference's
// error chain, and for every function of the PMC path a copy host -> HBM, the kernels, a copy HBM -> host.
// Nothing here loops over samples to do arithmetic, with two deliberate exceptions that the reference's contract
// forces onto the host: (1) an OPAQUE user target callback (allmc.h:21) is called per sample on the host, with the
// reference's per-sample error swallowing (pmc.h:52-66); (2) integer bookkeeping on host-owned arrays (evidence()'s
// flag count, pmc.c:1092-1095).
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/pmcb200.h"
#include "../../include/pmclib_abi.h"
#include "pmc_dropin_internal.h"
#include "pmc_hostonly.h"

namespace {

// reference error codes (pmc.h:23-43, mvdens.h:54-68, parabox.h:9-11, distribution.h:115-117, errorlist.h:147-149)
constexpr int pmc_base = -6000, pmc_allocate = pmc_base - 1, pmc_io = pmc_base - 10, pmc_incompat = pmc_base - 16,
              pmc_undef = pmc_base - 8, pmc_infinite = pmc_base - 19, pmc_isLog = pmc_base - 20;
constexpr int mv_base = -700, mv_negative = mv_base - 7, mv_io = mv_base - 10;
constexpr int pb_infnan = -12002, dist_undef = -6701, err_allocate = -102;

#define ERR_HERE(e, code, ...) e = newErrorVA(code, __func__, (char *)"(pmc_dropin.cpp)", (char *)__VA_ARGS__)
#define FWD(e, ret)                                                                              \
  do {                                                                                           \
    if (isError(e)) {                                                                            \
      e = newError(forwardErr, (char *)__func__, (char *)"", nullptr, e);                        \
      return ret;                                                                                \
    }                                                                                            \
  } while (0)

void *xmalloc(size_t n, error **err) {
  void *p = malloc(n ? n : 1);
  if (!p) *err = newErrorVA(err_allocate, __func__, (char *)"", (char *)"Cannot allocate %ld bytes", *err, nullptr, (long)n);
  return p;
}

int g_resident = -1;
bool resident() {
  if (g_resident < 0) {
    const char *e = getenv("PMCB200_RESIDENT");
    g_resident = (e && atoi(e)) ? 1 : 0;
  }
  return g_resident == 1;
}

// Rendezvous of the per-GPU host threads of one iteration (single-process multi-GPU mode): the all-reduce callback of
// every shard context meets here; the last arrival reduces in rank order (deterministic, identical on all ranks).
struct Rendezvous {
  std::mutex m;
  std::condition_variable cv;
  int n = 1, arrived = 0;
  long gen = 0;
  bool broken = false;
  std::vector<double *> bufs;
  std::vector<double> acc;
};
struct RvRank { Rendezvous *rv; int rank; };
int rv_allreduce(void *user, double *buf, long len, int op) {
  RvRank *me = (RvRank *)user;
  Rendezvous &rv = *me->rv;
  std::unique_lock<std::mutex> lk(rv.m);
  if (rv.broken) return 1;
  rv.bufs[me->rank] = buf;
  if (++rv.arrived == rv.n) {
    rv.acc.assign(buf, buf + len);
    for (long i = 0; i < len; ++i) {
      double a = rv.bufs[0][i];
      for (int r = 1; r < rv.n; ++r) {
        const double v = rv.bufs[r][i];
        a = op ? (v > a ? v : a) : a + v;
      }
      rv.acc[i] = a;
    }
    for (int r = 0; r < rv.n; ++r) memcpy(rv.bufs[r], rv.acc.data(), sizeof(double) * len);
    rv.arrived = 0;
    ++rv.gen;
    rv.cv.notify_all();
    return 0;
  }
  const long g = rv.gen;
  if (!rv.cv.wait_for(lk, std::chrono::seconds(600), [&] { return rv.gen != g || rv.broken; })) rv.broken = true;
  return rv.broken ? 1 : 0;
}

struct Side {
  pmcgpu_ctx *ctx = nullptr;
  long N = 0;
  int d = 0;
  uint32_t iter = 0;
  bool host_stale = false;  // resident mode: HBM holds newer sample arrays than the host
  void *pinned = nullptr;   // psim->data while it is page-locked
  // single-process multi-GPU mode (PMCB200_GPUS > 1): shard[0] == ctx, shard[g] lives on GPU g
  std::vector<pmcgpu_ctx *> shard;
  std::vector<RvRank> rvrank;
  Rendezvous rv;
};
std::map<const pmc_simu *, Side> g_side;
std::map<const distribution *, int> g_registered;
std::mutex g_mu;
uint64_t g_seed_counter = 0x9E3779B97F4A7C15ull;

int device_ordinal() {
  if (const char *e = getenv("PMCB200_DEVICE")) return atoi(e);
  return 0;
}

Side *side_for(pmc_simu *psim, error **err) {
  std::lock_guard<std::mutex> lk(g_mu);
  Side &s = g_side[psim];
  if (!s.ctx) {
    int dev = 0;
    if (const char *e = getenv("PMCB200_DEVICE")) dev = atoi(e);
    const int rc = pmcgpu_create(dev, &s.ctx);
    if (rc) {
      *err = newErrorVA(rc, __func__, (char *)"", (char *)"libpmcb200: no usable CUDA device %d (there is no CPU fallback)", *err, nullptr, dev);
      g_side.erase(psim);
      return nullptr;
    }
  }
  if (s.N != psim->nsamples || s.d != psim->ndim) {
    const int rc = pmcgpu_resize(s.ctx, psim->nsamples, psim->ndim);
    if (rc) {
      *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(s.ctx));
      return nullptr;
    }
    s.N = psim->nsamples;
    s.d = psim->ndim;
  }
  if (s.pinned != psim->data) {  // page-lock the caller's sample block once (it is re-allocated by pmc_simu_realloc)
    const size_t bytes = (size_t)psim->nsamples * ((psim->ndim + 2 + psim->n_ded) * sizeof(double) + sizeof(short) + sizeof(size_t));
    s.pinned = (pmcgpu_pin_host(psim->data, bytes) == 0) ? psim->data : nullptr;
    if (!s.pinned) s.pinned = psim->data;  // pageable copies still work; do not retry every call
  }
  return &s;
}

int gpu_fail(Side *s, int rc, error **err, const char *what) {
  *err = newErrorVA(rc, what, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(s->ctx));
  return rc;
}

// flat copies of a host mixture (components factorised first, as every reference entry point does lazily)
struct FlatMix {
  int K = 0, d = 0;
  std::vector<double> w, mean, L, detL;
  std::vector<int> df;
  std::vector<size_t> band;
  bool banded = false;
};
bool flatten(mix_mvdens *m, FlatMix &f, error **err) {
  mix_mvdens_cholesky_decomp(m, err);
  if (isError(*err)) return false;
  f.K = (int)m->ncomp;
  f.d = (int)m->ndim;
  const size_t d = f.d;
  f.w.assign(m->wght, m->wght + f.K);
  f.mean.resize(f.K * d);
  f.L.resize(f.K * d * d);
  f.detL.resize(f.K);
  f.df.resize(f.K);
  f.band.resize(f.K);
  for (int k = 0; k < f.K; ++k) {
    const mvdens *c = m->comp[k];
    memcpy(&f.mean[k * d], c->mean, d * sizeof(double));
    memcpy(&f.L[k * d * d], c->std, d * d * sizeof(double));
    f.detL[k] = c->detL;
    f.df[k] = c->df;
    f.band[k] = c->band_limit;
    if (c->band_limit < d) f.banded = true;
  }
  return true;
}
void unflatten(mix_mvdens *m, const double *w, const double *mean, const double *L, const double *detL) {
  const size_t d = m->ndim;
  for (size_t k = 0; k < m->ncomp; ++k) {
    mvdens *c = m->comp[k];
    m->wght[k] = w[k];
    memcpy(c->mean, mean + k * d, d * sizeof(double));
    memcpy(c->std, L + k * d * d, d * d * sizeof(double));
    c->chol = 1;  // pmc.c:1509: every component leaves the update flagged as factorised
    if (w[k] > 0) c->detL = detL[k];
  }
  m->init_cwght = 0;  // pmc.c:1552
}

bool push_proposal(Side *s, mix_mvdens *m, error **err) {
  FlatMix f;
  if (!flatten(m, f, err)) return false;
  int rc = pmcgpu_set_proposal(s->ctx, f.K, f.d, f.w.data(), f.mean.data(), f.L.data(), f.detL.data(), f.df.data());
  if (!rc && f.banded) rc = pmcgpu_set_band_limit(s->ctx, f.band.data());
  if (rc) { gpu_fail(s, rc, err, "push_proposal"); return false; }
  return true;
}

bool push_box(Side *s, const parabox *pb, int d, error **err) {
  const int rc = pb ? pmcgpu_set_box(s->ctx, pb->ndim, pb->nfaces, pb->faces) : pmcgpu_set_box(s->ctx, d, 0, nullptr);
  if (rc) { gpu_fail(s, rc, err, "push_box"); return false; }
  return true;
}

// which device density, if any, evaluates this distribution
int device_kind(const distribution *t) {
  if (!t) return PMCGPU_TGT_NONE;
  auto it = g_registered.find(t);
  if (it != g_registered.end()) return it->second;
  if (t->log_pdf == (posterior_log_pdf_func *)&mix_mvdens_log_pdf_void) return PMCGPU_TGT_MIX;
  if (t->log_pdf == (posterior_log_pdf_func *)&mvdens_log_pdf_void) return PMCGPU_TGT_MVDENS;
  if (t->log_pdf == (posterior_log_pdf_func *)&bentGauss_lkl) return PMCGPU_TGT_BANANA;
  if (t->log_pdf == &combine_lkl) {
    // a combine (distribution.c:350-396) is evaluated on the device when every term is a density this library knows,
    // reads only free parameters and neither side carries deduced or conditional parameters; otherwise the whole
    // combine is an opaque host target like any user callback
    const comb_dist_data *cbd = (const comb_dist_data *)t->data;
    if (!cbd || cbd->nded != 0 || t->ndef != 0 || cbd->ndist < 1) return PMCGPU_TGT_HOST;
    for (int id = 0; id < cbd->ndist; ++id) {
      const distribution *a = cbd->dist[id];
      if (!a || a->ndef != 0 || a->n_ded != 0 || a->log_pdf == &combine_lkl) return PMCGPU_TGT_HOST;
      const int k = device_kind(a);
      if (k != PMCGPU_TGT_BANANA && k != PMCGPU_TGT_MVDENS && k != PMCGPU_TGT_MIX) return PMCGPU_TGT_HOST;
      if (k == PMCGPU_TGT_MVDENS && ((const mvdens *)a->data)->band_limit < ((const mvdens *)a->data)->ndim) return PMCGPU_TGT_HOST;
      if (k == PMCGPU_TGT_MIX) {
        const mix_mvdens *mm = (const mix_mvdens *)a->data;
        for (size_t c = 0; c < mm->ncomp; ++c) if (mm->comp[c]->band_limit < mm->ndim) return PMCGPU_TGT_HOST;
      }
      for (int ip = 0; ip < a->ndim; ++ip) if (cbd->dim[id][ip] < 0) return PMCGPU_TGT_HOST;
    }
    return PMCGPU_TGT_COMBINE;
  }
  return PMCGPU_TGT_HOST;
}

bool push_target(Side *s, distribution *t, error **err) {
  const int kind = device_kind(t);
  int rc = 0;
  if (kind == PMCGPU_TGT_BANANA) {
    const bentGauss *bg = (const bentGauss *)t->data;
    rc = pmcgpu_set_target_banana(s->ctx, bg->ndim, bg->b, bg->sig1);
  } else if (kind == PMCGPU_TGT_MVDENS) {
    mvdens *g = (mvdens *)t->data;
    mvdens_cholesky_decomp(g, err);
    if (isError(*err)) return false;
    const double one = 1.0;
    rc = pmcgpu_set_target_mix(s->ctx, kind, 1, (int)g->ndim, &one, g->mean, g->std, &g->detL, &g->df);
  } else if (kind == PMCGPU_TGT_MIX) {
    FlatMix f;
    if (!flatten((mix_mvdens *)t->data, f, err)) return false;
    rc = pmcgpu_set_target_mix(s->ctx, kind, f.K, f.d, f.w.data(), f.mean.data(), f.L.data(), f.detL.data(), f.df.data());
  } else if (kind == PMCGPU_TGT_COMBINE) {
    const comb_dist_data *cbd = (const comb_dist_data *)t->data;
    std::vector<pmcgpu_combine_term> terms(cbd->ndist);
    std::vector<FlatMix> flat(cbd->ndist);
    const double one = 1.0;
    for (int id = 0; id < cbd->ndist; ++id) {
      distribution *a = cbd->dist[id];
      pmcgpu_combine_term &tm = terms[id];
      memset(&tm, 0, sizeof tm);
      tm.kind = device_kind(a);
      tm.ndim = a->ndim;
      tm.dim_idx = cbd->dim[id];
      if (tm.kind == PMCGPU_TGT_BANANA) {
        const bentGauss *bg = (const bentGauss *)a->data;
        tm.b = bg->b;
        tm.sig1 = bg->sig1;
      } else if (tm.kind == PMCGPU_TGT_MVDENS) {
        mvdens *g = (mvdens *)a->data;
        mvdens_cholesky_decomp(g, err);
        if (isError(*err)) return false;
        tm.K = 1; tm.wght = &one; tm.mean = g->mean; tm.L = g->std; tm.detL = &g->detL; tm.df = &g->df;
      } else {
        FlatMix &f = flat[id];
        if (!flatten((mix_mvdens *)a->data, f, err)) return false;
        tm.K = f.K; tm.wght = f.w.data(); tm.mean = f.mean.data(); tm.L = f.L.data(); tm.detL = f.detL.data(); tm.df = f.df.data();
      }
    }
    rc = pmcgpu_set_target_combine(s->ctx, t->ndim, cbd->ndist, terms.data());
  } else {
    rc = pmcgpu_set_target_host(s->ctx, t->ndim);
  }
  if (!rc) {
    if (kind != PMCGPU_TGT_HOST && kind != PMCGPU_TGT_COMBINE && t->ndef != 0) rc = pmcgpu_set_target_defaults(s->ctx, t->ndim + t->ndef, t->def, t->pars);
    else rc = pmcgpu_set_target_defaults(s->ctx, 0, nullptr, nullptr);
  }
  if (rc) { gpu_fail(s, rc, err, "push_target"); return false; }
  return true;
}

// Philox key for one draw call: two 32-bit words from the caller's gsl_rng (so the caller's seed decides the stream),
// or a process-local counter when no generator is given
uint64_t seed_from(gsl_rng *r) {
  if (r && r->type && r->type->get) {
    const uint64_t a = (uint64_t)(r->type->get(r->state) & 0xffffffffu), b = (uint64_t)(r->type->get(r->state) & 0xffffffffu);
    return (a << 32) | b;
  }
  g_seed_counter = g_seed_counter * 6364136223846793005ull + 1442695040888963407ull;
  return g_seed_counter;
}

bool pull_samples(Side *s, pmc_simu *psim, bool X, bool idx, bool w, bool r, bool f, error **err) {
  const int rc = pmcgpu_download(s->ctx, X ? psim->X : nullptr, idx ? psim->indices : nullptr, w ? psim->weights : nullptr,
                                 r ? psim->log_rho : nullptr, f ? psim->flg : nullptr);
  if (rc) { gpu_fail(s, rc, err, "download"); return false; }
  return true;
}
bool push_samples(Side *s, pmc_simu *psim, bool X, bool idx, bool w, bool r, bool f, error **err) {
  if (s->host_stale) return true;  // resident mode: HBM is authoritative
  const int rc = pmcgpu_upload(s->ctx, X ? psim->X : nullptr, idx ? psim->indices : nullptr, w ? psim->weights : nullptr,
                               r ? psim->log_rho : nullptr, f ? psim->flg : nullptr);
  if (rc) { gpu_fail(s, rc, err, "upload"); return false; }
  return true;
}

// a one-shot context for the point-wise density entry points (mvdens_log_pdf & co.)
pmcgpu_ctx *point_ctx(error **err) {
  static pmcgpu_ctx *c = nullptr;
  if (!c) {
    int dev = 0;
    if (const char *e = getenv("PMCB200_DEVICE")) dev = atoi(e);
    const int rc = pmcgpu_create(dev, &c);
    if (rc) {
      *err = newErrorVA(rc, __func__, (char *)"", (char *)"libpmcb200: no usable CUDA device (there is no CPU fallback)", *err, nullptr);
      c = nullptr;
    }
  }
  return c;
}

}  // namespace

namespace pmcb200 {
std::mutex &dropin_point_mutex() {
  static std::mutex m;
  return m;
}
pmcgpu_ctx *dropin_point_ctx(error **err) { return point_ctx(err); }
uint64_t dropin_seed_from(gsl_rng *r) { return seed_from(r); }
bool dropin_push_model(pmcgpu_ctx *ctx, mix_mvdens *m, distribution *target, parabox *pb, int d, error **err) {
  Side tmp;
  tmp.ctx = ctx;
  return push_proposal(&tmp, m, err) && (!target || push_target(&tmp, target, err)) && push_box(&tmp, pb, d, err);
}
void dropin_unflatten(mix_mvdens *m, const double *w, const double *mean, const double *L, const double *detL) { unflatten(m, w, mean, L, detL); }
int dropin_device_kind(const distribution *t) { return device_kind(t); }
// The per-sample loop over an OPAQUE target callback (pmc.c:557-577): errors of a sample are printed (unless quiet),
// purged and the sample flagged out (ParameterErrorVerb, pmc.h:52-66) - except tls_cosmo_par, which aborts the whole
// call with the chain forwarded (pmc.c:559-560).  Returns false in that case.
bool dropin_host_target_values(posterior_log_pdf_func *f, retrieve_ded_func *ret, void *target_data, const double *X, double *X_ded,
                               long N, int d, int nd, int quiet, double *post, short *ok, error **err) {
  constexpr int tls_cosmo_par = -8017;  // tools.h:24,42
  for (long i = 0; i < N; ++i) {
    const double *x = X + (size_t)i * d;
    ok[i] = 1;
    post[i] = f(target_data, x, err);
    if (isError(*err) && getErrorValue(*err) == tls_cosmo_par) {
      *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err);
      return false;
    }
    if (!isError(*err) && ret && nd > 0) ret(target_data, X_ded + (size_t)i * nd, err);
    if (isError(*err)) {
      if (!quiet) {
        fprintf(stderr, "*** Error, parameter ***\n");
        for (int j = 0; j < d; ++j) fprintf(stderr, " %10g", x[j]);
        fprintf(stderr, "\n*** -- ***\n");
        printError(stderr, *err);
      }
      purgeError(err);
      ok[i] = 0;
    }
  }
  return true;
}
int dropin_device_ordinal() { return device_ordinal(); }
}  // namespace pmcb200

extern "C" {

// ====================================================================================================================
// errorlist (pmctools/src/errorlist.c)
// ====================================================================================================================
int _isError(error *err) { return err != nullptr && err->errValue != placeHolderError; }

void purgeError(error **err) {
  error *e = *err;
  while (e) {
    error *n = e->next;
    free(e);
    e = n;
  }
  *err = nullptr;
}

error *newError(int errV, char *where, char *text, error *linkTo, error *addErr) {
  error *er = (error *)malloc(sizeof(error));
  if (!er) {
    fprintf(stderr, "Cannot allocate memory for error management! Aborting\n");
    exit(-1);
  }
  if (snprintf(er->errWhere, WHR_SZ, "%s", where ? where : "") <= 0) snprintf(er->errWhere, WHR_SZ, "PROBLEM IN ERROR TEXT");
  if (snprintf(er->errText, TXT_SZ, "%s", text ? text : "") <= 0) snprintf(er->errText, TXT_SZ, "PROBLEM IN ERROR TEXT");
  er->errValue = errV;
  er->next = nullptr;
  if (_isError(addErr)) er->next = addErr;
  else purgeError(&addErr);
  if (_isError(linkTo)) {
    linkTo->next = er;
    return linkTo;
  }
  purgeError(&linkTo);
  return er;
}

error *newErrorVA(int errV, const char *func, char *where, char *text, error *linkTo, error *addErr, ...) {
  char w[TXT_SZ], t[TXT_SZ];
  snprintf(w, sizeof w, "%s%s", func ? func : "", where ? where : "");
  va_list ap;
  va_start(ap, addErr);
  const int sz = vsnprintf(t, TXT_SZ, text ? text : "", ap);
  va_end(ap);
  if (sz < 0) snprintf(t, sizeof t, "PROBLEM IN ERROR TEXT");
  return newError(errV, w, t, linkTo, addErr);
}

void stringError(char *str, error *err) {
  size_t pos = 0, depth = 0;
  str[0] = '\0';
  for (error *e = err; e; e = e->next, ++depth) {
    for (size_t i = 0; i < 2 * depth && i < 500; ++i) str[pos++] = ' ';
    if (e->errValue == forwardErr) pos += sprintf(str + pos, "%s::ForwardError\n", e->errWhere);
    else pos += sprintf(str + pos, "%s::Error %d (%s)\n", e->errWhere, e->errValue, e->errText);
  }
}

void printError(FILE *flog, error *err) {
  std::vector<char> buf((TXT_SZ + WHR_SZ) * 50);
  stringError(buf.data(), err);
  fprintf(flog ? flog : stderr, "%s", buf.data());
}

int getErrorValue(error *err) {
  if (!err) return noErr;
  if (err->errValue == forwardErr) return err->next ? getErrorValue(err->next) : undefErr;
  return err->errValue;
}

error *initError(void) { return newError(placeHolderError, (char *)"", (char *)"PlaceHolder", nullptr, nullptr); }
void endError(error **err) { purgeError(err); }

// ====================================================================================================================
// parabox (pmclib/src/parabox.c)
// ====================================================================================================================
parabox *init_parabox(int parasize, error **err) {
  parabox *b = (parabox *)xmalloc(sizeof(parabox), err);
  FWD(*err, nullptr);
  b->ndim = parasize;
  b->nfaces = 0;
  b->faces = nullptr;
  return b;
}
void add_face(parabox *bx, double *dirvec, double threshold, error **err) {
  const int nf = bx->nfaces, d = bx->ndim;
  double *nw = (double *)xmalloc(sizeof(double) * (nf + 1) * (d + 1), err);
  FWD(*err, );
  if (nf) memcpy(nw, bx->faces, sizeof(double) * nf * (d + 1));
  memcpy(nw + (size_t)nf * (d + 1), dirvec, sizeof(double) * d);
  nw[(size_t)nf * (d + 1) + d] = threshold;
  free(bx->faces);
  bx->faces = nw;
  bx->nfaces = nf + 1;
}
void add_halfspace(parabox *bx, int coord, double sign, double val, error **err) {
  std::vector<double> dir(bx->ndim, 0.0);
  dir[coord] = sign;
  add_face(bx, dir.data(), val, err);
  FWD(*err, );
}
void add_lobound(parabox *bx, int coord, double vmin, error **err) { add_halfspace(bx, coord, -1.0, -vmin, err); FWD(*err, ); }
void add_hibound(parabox *bx, int coord, double vmax, error **err) { add_halfspace(bx, coord, 1.0, vmax, err); FWD(*err, ); }
void add_slab(parabox *bx, int coord, double vmin, double vmax, error **err) {
  add_halfspace(bx, coord, -1.0, -vmin, err);
  FWD(*err, );
  add_halfspace(bx, coord, 1.0, vmax, err);
  FWD(*err, );
}
// single point: this is the scalar predicate of the reference API, evaluated on the GPU like everything else
int isinBox(const parabox *bx, double *pos, error **err) {
  for (int j = 0; j < bx->ndim; ++j)
    if (!std::isfinite(pos[j])) {
      *err = newErrorVA(pb_infnan, __func__, (char *)"", (char *)"Parameter #%d is not finite", *err, nullptr, j);
      return 0;
    }
  std::lock_guard<std::mutex> lk(pmcb200::dropin_point_mutex());
  pmcgpu_ctx *c = point_ctx(err);
  if (!c) return 0;
  int inside = 0;
  int rc = pmcgpu_set_box(c, bx->ndim, bx->nfaces, bx->faces);
  if (!rc) rc = pmcgpu_isinbox(c, 1, pos, &inside);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  return inside == 1;
}
void free_parabox(parabox **bx) {
  if (!bx || !*bx) return;
  free((*bx)->faces);
  free(*bx);
  *bx = nullptr;
}

// ====================================================================================================================
// distribution (pmclib/src/distribution.c:34-260)
// ====================================================================================================================
distribution *init_distribution_full(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                     simulate_func *simulate, int n_ded, retrieve_ded_func *retrieve, error **err) {
  if (n_ded != 0 && retrieve == nullptr) {
    ERR_HERE(*err, dist_undef, "Target invalid, expect deduced parameters, but no function provided...", *err, nullptr);
    return nullptr;
  }
  distribution *d = (distribution *)xmalloc(sizeof(distribution), err);
  FWD(*err, nullptr);
  memset(d, 0, sizeof(*d));
  d->ndim = ndim;
  d->n_ded = n_ded;
  d->data = data;
  d->log_pdf = log_pdf;
  d->free = freef;
  d->simulate = simulate;
  d->retrieve = retrieve;
  return d;
}
distribution *init_distribution(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                simulate_func *simulate, error **err) {
  distribution *d = init_distribution_full(ndim, data, log_pdf, freef, simulate, 0, nullptr, err);
  FWD(*err, nullptr);
  return d;
}
distribution *init_simple_distribution(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                       error **err) {
  distribution *d = init_distribution_full(ndim, data, log_pdf, freef, nullptr, 0, nullptr, err);
  FWD(*err, nullptr);
  return d;
}
void free_distribution(distribution **pdist) {
  distribution *d = *pdist;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_registered.erase(d);
  }
  if (d->name) free(d->name);
  if (d->free && d->data) d->free(&d->data);
  if (d->ndef != 0) { free(d->pars); free(d->def); }
  free(d);
  *pdist = nullptr;
}
void distribution_set_default(distribution *dist, int ndef, int *idef, double *vdef, error **err) {
  if (ndef >= dist->ndim + dist->ndef) {
    ERR_HERE(*err, dist_undef, "Too many defaults ! (expected at much %d, got %d)", *err, nullptr, dist->ndim + dist->ndef, ndef);
    return;
  }
  if (dist->ndef != 0) {
    free(dist->def);
    free(dist->pars);
    dist->ndim += dist->ndef;
    dist->ndef = 0;
  }
  dist->def = (int *)xmalloc(sizeof(int) * dist->ndim, err);
  FWD(*err, );
  memset(dist->def, 0, sizeof(int) * dist->ndim);
  dist->pars = (double *)xmalloc(sizeof(double) * dist->ndim, err);
  FWD(*err, );
  for (int i = 0; i < ndef; ++i) {
    if (idef[i] >= dist->ndim) { ERR_HERE(*err, dist_undef, "Too many defaults ! (expected at much %d, got %d)", *err, nullptr, dist->ndim, idef[i]); return; }
    if (dist->def[idef[i]] == 1) { ERR_HERE(*err, dist_undef, "par %d already set", *err, nullptr, idef[i]); return; }
    dist->def[idef[i]] = 1;
    dist->pars[idef[i]] = vdef[i];
  }
  dist->ndef = ndef;
  dist->ndim -= ndef;
}
// distribution.c:193-225: scatter the free parameters into the full vector, then the density
// distribution.c:193-209 (a global symbol of libpmc, not declared in distribution.h): the free parameters scattered into the
// full vector of a distribution with conditional ("default") parameters; without defaults the caller's vector itself
double *distribution_fill_pars(distribution *dist, double *pars, error **err) {
  (void)err;
  if (dist->ndef == 0) return pars;
  int ir = 0;
  for (int ik = 0; ik < dist->ndim + dist->ndef; ++ik)
    if (dist->def[ik] == 0) dist->pars[ik] = pars[ir++];
  return dist->pars;
}
double distribution_lkl(void *pdist, const double *pars, error **err) {
  distribution *dist = (distribution *)pdist;
  if (!dist->log_pdf) { ERR_HERE(*err, dist_undef, "undefined log pdf function for distribution", *err, nullptr); return 0; }
  const double *p = distribution_fill_pars(dist, const_cast<double *>(pars), err);
  const double res = dist->log_pdf(dist->data, p, err);
  FWD(*err, 0);
  return res;
}
void distribution_retrieve(const void *pdist, double *pars, error **err) {
  const distribution *dist = (const distribution *)pdist;
  if (dist->retrieve == nullptr && dist->n_ded != 0) { ERR_HERE(*err, dist_undef, "Undefined retrieve function for distribution", *err, nullptr); return; }
  if (dist->n_ded == 0) return;
  dist->retrieve(dist->data, pars, err);
  FWD(*err, );
}

// ====================================================================================================================
// mvdens / mix_mvdens (pmctools/src/mvdens.c)
// ====================================================================================================================
static void set_views(mvdens *g) {
  g->mean_view_container.vector = gsl_vector{g->ndim, 1, g->mean, nullptr, 0};
  g->mean_view = &g->mean_view_container.vector;
  g->std_view_container.matrix = gsl_matrix{g->ndim, g->ndim, g->ndim, g->std, nullptr, 0};
  g->std_view = &g->std_view_container.matrix;
  g->x_tmp_view_container.vector = gsl_vector{g->ndim, 1, g->x_tmp, nullptr, 0};
  g->x_tmp_view = &g->x_tmp_view_container.vector;
}
mvdens *mvdens_init(mvdens *g, size_t size, double *mn, double *std, error **err) {
  if (size == 0) { ERR_HERE(*err, mv_negative, "Invalid size = %d, has to be positive", *err, nullptr, (int)size); return nullptr; }
  g->ndim = size;
  g->mean = mn;
  g->std = std;
  g->band_limit = size;
  g->chol = 0;
  g->df = -1;
  g->own_buf = 0;
  g->detL = -1.0;
  g->x_tmp = (double *)xmalloc(sizeof(double) * size, err);
  FWD(*err, nullptr);
  set_views(g);
  return g;
}
mvdens *mvdens_alloc(size_t size, error **err) {
  mvdens *g = (mvdens *)xmalloc(sizeof(mvdens), err);
  FWD(*err, nullptr);
  g->buf = calloc(size * (size + 1), sizeof(double));
  if (!g->buf) { ERR_HERE(*err, err_allocate, "Cannot allocate", *err, nullptr); return nullptr; }
  g = mvdens_init(g, size, (double *)g->buf, ((double *)g->buf) + size, err);
  FWD(*err, nullptr);
  g->own_buf = 1;
  return g;
}
void mvdens_empty(mvdens *g) { if (g->own_buf == 1) free(g->buf); }
void mvdens_free(mvdens **g) {
  mvdens_empty(*g);
  free((*g)->x_tmp);
  free(*g);
  *g = nullptr;
}
void mvdens_free_void(void **g) { mvdens_free((mvdens **)g); }
void mvdens_set_band_limit(mvdens *self, size_t value) { self->band_limit = value; }
double determinant(const double *L, size_t ndim) {
  double det = 1.0;
  for (size_t i = 0; i < ndim * ndim; i += ndim + 1) det *= L[i];
  return det;
}
// mvdens.c:229-277: O(d^3) on the host, GSL-1.14 semantics, idempotent when chol==1
void mvdens_cholesky_decomp(mvdens *self, error **err) {
  if (self->chol == 1) return;
  double det = 0.0;
  if (pmcgpu_cholesky((int)self->ndim, self->std, &det) != 0) {
    *err = newErrorVA(mv_base - 6, __func__, (char *)"", (char *)"Cholesky decomposition failed", *err, nullptr);
    return;
  }
  self->detL = det;
  self->chol = 1;
}
void mix_mvdens_cholesky_decomp(mix_mvdens *self, error **err) {
  for (size_t i = 0; i < self->ncomp; ++i) {
    mvdens_cholesky_decomp(self->comp[i], err);
    FWD(*err, );
  }
}
mix_mvdens *mix_mvdens_alloc(size_t ncomp, size_t size, error **err) {
  mix_mvdens *m = (mix_mvdens *)xmalloc(sizeof(mix_mvdens), err);
  FWD(*err, nullptr);
  m->ncomp = ncomp;
  m->ndim = size;
  m->wght = (double *)calloc(ncomp, sizeof(double));
  m->cwght = (double *)calloc(ncomp, sizeof(double));
  m->comp = (mvdens **)xmalloc(ncomp * sizeof(mvdens *), err);
  FWD(*err, nullptr);
  m->wght_view_container.vector = gsl_vector{ncomp, 1, m->wght, nullptr, 0};
  m->wght_view = &m->wght_view_container.vector;
  m->cwght_view_container.vector = gsl_vector{ncomp, 1, m->cwght, nullptr, 0};
  m->cwght_view = &m->cwght_view_container.vector;
  const size_t per = size * (size + 1) * sizeof(double);
  m->buf_comp = xmalloc(ncomp * (sizeof(mvdens) + per), err);
  FWD(*err, nullptr);
  char *data = ((char *)m->buf_comp) + ncomp * sizeof(mvdens);
  memset(data, 0, ncomp * per);
  for (size_t i = 0; i < ncomp; ++i) {
    m->comp[i] = mvdens_init((mvdens *)((char *)m->buf_comp + i * sizeof(mvdens)), size, (double *)(data + per * i),
                             (double *)(data + per * i + size * sizeof(double)), err);
    FWD(*err, nullptr);
  }
  m->init_cwght = 0;
  return m;
}
void mix_mvdens_free(mix_mvdens **m) {
  for (size_t i = 0; i < (*m)->ncomp; ++i) { mvdens_empty((*m)->comp[i]); free((*m)->comp[i]->x_tmp); }
  free((*m)->comp);
  free((*m)->buf_comp);
  free((*m)->wght);
  free((*m)->cwght);
  free(*m);
  *m = nullptr;
}
void mix_mvdens_free_void(void **m) { mix_mvdens_free((mix_mvdens **)m); }
void mix_mvdens_set_band_limit(mix_mvdens *self, size_t value) {
  for (size_t i = 0; i < self->ncomp; ++i) self->comp[i]->band_limit = value;
}
size_t ranindx(double *cw, size_t nE, gsl_rng *r, error **err) {  // mvdens.c:766-782, one uniform from the caller's generator
  (void)err;
  const double z = r->type->get_double(r->state);  // gsl_ran_flat(r,0,1) = 0*(1-u) + 1*u
  size_t index = 0;
  while (cw[index] < z && index < nE - 1) index++;
  return index;
}

// Point-wise densities: one upload + one launch + one synchronisation per call (tens of microseconds); bulk evaluation
// should go through generic_get_importance_weight*.  The calls share one context behind a mutex, and the density that
// was uploaded last is remembered, so a likelihood that evaluates the SAME mvdens / mixture sample after sample (the
// usual pattern in pmclib targets) pays for the upload once.
namespace {
struct PointCache {
  FlatMix mix;        // last mixture installed as the proposal of the point context
  bool have_mix = false;
  std::vector<double> gm, gL;  // last mvdens installed as its target
  int gdf = 0;
  bool have_g = false;
};
PointCache g_pc;
bool same_mix(const FlatMix &a, const FlatMix &b) {
  return a.K == b.K && a.d == b.d && a.w == b.w && a.mean == b.mean && a.L == b.L && a.detL == b.detL && a.df == b.df;
}
// installs g (factorised) as the target of the point context unless it is already there
int point_install_mvdens(pmcgpu_ctx *c, mvdens *g) {
  const size_t d = g->ndim;
  if (g_pc.have_g && g_pc.gdf == g->df && g_pc.gm.size() == d && !memcmp(g_pc.gm.data(), g->mean, d * sizeof(double)) &&
      !memcmp(g_pc.gL.data(), g->std, d * d * sizeof(double)))
    return 0;
  const double one = 1.0;
  int rc = pmcgpu_set_target_mix(c, PMCGPU_TGT_MVDENS, 1, (int)d, &one, g->mean, g->std, &g->detL, &g->df);
  if (!rc) rc = pmcgpu_set_target_defaults(c, 0, nullptr, nullptr);
  g_pc.have_g = rc == 0;
  if (!rc) { g_pc.gm.assign(g->mean, g->mean + d); g_pc.gL.assign(g->std, g->std + d * d); g_pc.gdf = g->df; }
  return rc;
}
}  // namespace
}  // extern "C"
namespace pmcb200 {
void dropin_point_invalidate() { g_pc.have_mix = false; g_pc.have_g = false; }
}  // namespace pmcb200
extern "C" {
double mix_mvdens_log_pdf(mix_mvdens *m, const double *x, error **err) {
  std::lock_guard<std::mutex> lk(pmcb200::dropin_point_mutex());
  pmcgpu_ctx *c = point_ctx(err);
  if (!c) return 0;
  FlatMix f;
  if (!flatten(m, f, err)) { FWD(*err, 0); }
  double out = 0.0;
  int rc = 0;
  if (!g_pc.have_mix || !same_mix(f, g_pc.mix)) {
    rc = pmcgpu_set_proposal(c, f.K, f.d, f.w.data(), f.mean.data(), f.L.data(), f.detL.data(), f.df.data());
    g_pc.have_mix = rc == 0;
    if (!rc) g_pc.mix = f;
  }
  if (!rc) rc = pmcgpu_proposal_log_pdf(c, 1, x, &out);
  if (rc) { g_pc.have_mix = false; *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  return out;
}
double mix_mvdens_log_pdf_void(void *m, const double *x, error **err) {
  const double r = mix_mvdens_log_pdf((mix_mvdens *)m, x, err);
  FWD(*err, 0.0);
  return r;
}
double mvdens_log_pdf(mvdens *g, const double *x, error **err) {
  mvdens_cholesky_decomp(g, err);
  FWD(*err, -1);
  std::lock_guard<std::mutex> lk(pmcb200::dropin_point_mutex());
  pmcgpu_ctx *c = point_ctx(err);
  if (!c) return 0;
  double out = 0.0;
  int rc = point_install_mvdens(c, g);
  if (!rc) rc = pmcgpu_target_log_pdf(c, 1, x, &out);
  if (rc) { g_pc.have_g = false; *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  return out;
}
double mvdens_log_pdf_void(void *g, const double *x, error **err) {
  const double r = mvdens_log_pdf((mvdens *)g, x, err);
  FWD(*err, 0.0);
  return r;
}
// (x-mu)^T Sigma^-1 (x-mu) (mvdens.c:319-349): the forward substitution itself on the device, not a value recovered from
// the log-density (that would cancel for small q); g is not modified
double scalar_product(mvdens *g, const double *x, error **err) {
  mvdens_cholesky_decomp(g, err);
  FWD(*err, -1);
  std::lock_guard<std::mutex> lk(pmcb200::dropin_point_mutex());
  pmcgpu_ctx *c = point_ctx(err);
  if (!c) return 0;
  double out = 0.0;
  int rc = point_install_mvdens(c, g);
  if (!rc) rc = pmcgpu_target_scalar_product(c, 1, x, &out);
  if (rc) { g_pc.have_g = false; *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  return out;
}

// text formats (mvdens.c:114-134, 513-583)
void mvdens_dump(FILE *where, mvdens *what) {
  FILE *f = where ? where : stdout;
  const size_t d = what->ndim;
  fprintf(f, "%zu %d %zu %d\n", d, what->df, what->band_limit, what->chol);
  for (size_t j = 0; j < d; ++j) fprintf(f, "% 20.10g ", what->mean[j]);
  fprintf(f, "\n");
  for (size_t j = 0; j < d; ++j) {
    for (size_t k = 0; k < d; ++k) fprintf(f, "% 20.15g ", what->std[j * d + k]);
    fprintf(f, "\n");
  }
}
void mvdens_print(FILE *where, mvdens *what) {
  FILE *f = where ? where : stdout;
  const size_t d = what->ndim;
  if (what->df != -1) fprintf(f, "df = %d\n", what->df);
  if (what->chol == 1) fprintf(f, "Beware! Cholesky decomposed std, df=%d\n", what->df);
  for (size_t j = 0; j < d; ++j) {
    fprintf(f, "% .3g | ", what->mean[j]);
    for (size_t k = 0; k < j; ++k) fprintf(f, "% .3g ", what->std[j * d + k]);
    fprintf(f, "% .3g\n", what->std[j * d + j]);
  }
}
void mix_mvdens_dump(FILE *where, mix_mvdens *what) {
  FILE *f = where ? where : stdout;
  fprintf(f, "%d %d\n", (int)what->ncomp, (int)what->ndim);
  for (size_t i = 0; i < what->ncomp; ++i) {
    fprintf(f, "%.8g\n", what->wght[i]);
    mvdens_dump(f, what->comp[i]);
  }
}
void mix_mvdens_print(FILE *where, mix_mvdens *what) {
  FILE *f = where ? where : stdout;
  for (size_t i = 0; i < what->ncomp; ++i) {
    fprintf(f, "*************************************\n");
    fprintf(f, "Weight of component #%d is %g\n", (int)i, what->wght[i]);
    if (what->wght[i] != 0) mvdens_print(f, what->comp[i]);
  }
}
mix_mvdens *mix_mvdens_dwnp(FILE *where, error **err) {
  FILE *f = where ? where : stdin;
  size_t ncomp, ndim;
  if (fscanf(f, "%zu %zu\n", &ncomp, &ndim) != 2) { ERR_HERE(*err, mv_io, "Cannot read mix_mvdens first line (two integers expected)", *err, nullptr); return nullptr; }
  mix_mvdens *m = mix_mvdens_alloc(ncomp, ndim, err);
  FWD(*err, nullptr);
  for (size_t i = 0; i < ncomp; ++i) {
    size_t nd, df, bdl, chol;
    if (fscanf(f, "%lg\n", &m->wght[i]) != 1 || fscanf(f, "%zu %zu %zu %zu\n", &nd, &df, &bdl, &chol) != 4 || nd != ndim) {
      ERR_HERE(*err, mv_io, "Cannot read mix_mvdens (component %d)", *err, nullptr, (int)i);
      return nullptr;
    }
    mvdens *w = m->comp[i];
    for (size_t j = 0; j < ndim; ++j)
      if (fscanf(f, "%lg", w->mean + j) != 1) { ERR_HERE(*err, mv_io, "Cannot read mix_mvdens component %d (mean)", *err, nullptr, (int)i); return nullptr; }
    for (size_t j = 0; j < ndim * ndim; ++j)
      if (fscanf(f, "%lg", w->std + j) != 1) { ERR_HERE(*err, mv_io, "Cannot read mix_mvdens component %d (variance)", *err, nullptr, (int)i); return nullptr; }
    w->band_limit = bdl;
    w->df = (int)df;
    w->chol = (int)chol;
    if (w->chol == 1) w->detL = determinant(w->std, w->ndim);
  }
  return m;
}

// ====================================================================================================================
// example target (examples/target/sampleTarget.c)
// ====================================================================================================================
bentGauss *init_bentGauss(int ndim, double b, double sig1sq, error **err) {
  if (ndim < 2) { ERR_HERE(*err, -1000, "Number of dimensions too small (got %d, expected at least 2)", *err, nullptr, ndim); return nullptr; }
  bentGauss *bg = (bentGauss *)xmalloc(sizeof(bentGauss), err);
  FWD(*err, nullptr);
  bg->ndim = ndim;
  bg->b = b;
  bg->sig1 = sig1sq;
  return bg;
}
void bentGauss_free(void **pbg) { free(*pbg); *pbg = nullptr; }
// point-wise value; inside a PMC run the density is evaluated by the kernels, not through this pointer
double bentGauss_lkl(void *vbg, double *pars, error **err) {
  const bentGauss *bg = (const bentGauss *)vbg;
  std::lock_guard<std::mutex> lk(pmcb200::dropin_point_mutex());
  pmcgpu_ctx *c = point_ctx(err);
  if (!c) return 0;
  g_pc.have_g = false;  // the point context's target changes
  double out = 0.0;
  int rc = pmcgpu_set_target_banana(c, bg->ndim, bg->b, bg->sig1);
  if (!rc) rc = pmcgpu_set_target_defaults(c, 0, nullptr, nullptr);
  if (!rc) rc = pmcgpu_target_log_pdf(c, 1, pars, &out);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  return out;
}
distribution *init_bentGauss_distribution(int ndim, double b, double sig1, error **err) {
  bentGauss *bg = init_bentGauss(ndim, b, sig1, err);
  FWD(*err, nullptr);
  distribution *d = init_distribution(ndim, bg, (posterior_log_pdf_func *)&bentGauss_lkl, &bentGauss_free, nullptr, err);
  FWD(*err, nullptr);
  return d;
}

// ====================================================================================================================
// pmc_simu (pmclib/src/pmc.c)
// ====================================================================================================================
static void carve(pmc_simu *s) {  // pmc.c:42-56: one block, sub-arrays in this order
  const size_t N = s->nsamples;
  s->X = (double *)s->data;
  s->X_ded = (double *)((char *)s->X + N * s->ndim * sizeof(double));
  s->weights = (double *)((char *)s->X_ded + N * s->n_ded * sizeof(double));
  s->log_rho = (double *)((char *)s->weights + N * sizeof(double));
  s->indices = (size_t *)((char *)s->log_rho + N * sizeof(double));
  s->flg = (short *)((char *)s->indices + N * sizeof(size_t));
}
pmc_simu *pmc_simu_init_plus_ded(long nsamples, int ndim, int n_ded, error **err) {
  pmc_simu *s = (pmc_simu *)xmalloc(sizeof(pmc_simu), err);
  FWD(*err, nullptr);
  memset(s, 0, sizeof(*s));
  s->nsamples = nsamples;
  s->ndim = ndim;
  s->n_ded = n_ded;
  s->data = xmalloc((size_t)nsamples * ((ndim + 2) * sizeof(double) + n_ded * sizeof(double) + sizeof(short) + sizeof(size_t)), err);
  FWD(*err, nullptr);
  carve(s);
  s->prop_print_step = 20;
  return s;
}
pmc_simu *pmc_simu_init(long nsamples, int ndim, error **err) {
  pmc_simu *s = pmc_simu_init_plus_ded(nsamples, ndim, 0, err);
  FWD(*err, nullptr);
  return s;
}
void pmc_simu_free(pmc_simu **melf) {
  pmc_simu *s = *melf;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_side.find(s);
    if (it != g_side.end()) {
      if (it->second.pinned) pmcgpu_unpin_host(it->second.pinned);
      for (size_t g = 1; g < it->second.shard.size(); ++g) pmcgpu_destroy(it->second.shard[g]);
      if (it->second.ctx) pmcgpu_destroy(it->second.ctx);
      g_side.erase(it);
    }
  }
  free(s->data);
  if (s->target) free_distribution(&s->target);
  if (s->proposal) free_distribution(&s->proposal);
  free(s);
  *melf = nullptr;
}
void pmc_simu_realloc(pmc_simu *psim, long newsamples, error **err) {
  if (psim->nsamples == newsamples) return;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_side.find(psim);
    if (it != g_side.end() && it->second.pinned) { pmcgpu_unpin_host(it->second.pinned); it->second.pinned = nullptr; }
  }
  psim->nsamples = newsamples;
  free(psim->data);
  psim->data = xmalloc((size_t)newsamples * ((psim->ndim + 2) * sizeof(double) + psim->n_ded * sizeof(double) + sizeof(short) + sizeof(size_t)), err);
  FWD(*err, );
  carve(psim);
  psim->logSum = 0;
  psim->isLog = 0;
}
void pmc_simu_init_target(pmc_simu *psim, distribution *target, parabox *pb, error **err) {
  if (psim->target) free_distribution(&psim->target);
  if (target->ndim != psim->ndim) { ERR_HERE(*err, pmc_incompat, "Target invalid, got %d dims, while pmc initialized for %d", *err, nullptr, target->ndim, psim->ndim); return; }
  if (target->n_ded != psim->n_ded) { ERR_HERE(*err, pmc_incompat, "Target invalid, got %d deduced parameters, while pmc initialized for %d", *err, nullptr, target->n_ded, psim->n_ded); return; }
  if (target->n_ded != 0 && !target->retrieve) { ERR_HERE(*err, pmc_incompat, "Target invalid, expect deduced parameters, but nu function...", *err, nullptr); return; }
  if (!target->log_pdf) { ERR_HERE(*err, pmc_incompat, "Target invalid, absent log_pdf function", *err, nullptr); return; }
  psim->target = target;
  if (psim->pb) free_parabox(&psim->pb);
  psim->pb = pb;
}
void pmc_simu_init_proposal(pmc_simu *psim, distribution *proposal, int prop_print_step, error **err) {
  if (psim->proposal) free_distribution(&psim->proposal);
  if (proposal->ndim != psim->ndim) { ERR_HERE(*err, pmc_incompat, "Proposal invalid, got %d dims, while pmc initialized for %d", *err, nullptr, proposal->ndim, psim->ndim); return; }
  if (!proposal->log_pdf) { ERR_HERE(*err, pmc_incompat, "Proposal invalid, absent log_pdf function", *err, nullptr); return; }
  if (!proposal->simulate) { ERR_HERE(*err, pmc_incompat, "Proposal invalid, absent simulate function", *err, nullptr); return; }
  psim->proposal = proposal;
  psim->prop_print_step = prop_print_step;
}
void pmc_simu_init_pmc(pmc_simu *psim, void *filter_data, filter_func *pmc_filter, update_func *pmc_update) {
  psim->filter_data = filter_data;
  psim->pmc_filter = pmc_filter;
  psim->pmc_update = pmc_update;
}
distribution *mix_mvdens_distribution(int ndim, void *mxmv, error **err) {
  distribution *d = init_distribution_full(ndim, mxmv, &mix_mvdens_log_pdf_void, &mix_mvdens_free_void, &simulate_mix_mvdens_void, 0, nullptr, err);
  FWD(*err, nullptr);
  return d;
}
void pmc_simu_init_classic_importance(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data, int prop_print_step, error **err) {
  pmc_simu_init_target(psim, target, pb, err);
  FWD(*err, );
  distribution *prop = mix_mvdens_distribution(target->ndim, prop_data, err);
  FWD(*err, );
  pmc_simu_init_proposal(psim, prop, prop_print_step, err);
  FWD(*err, );
}
void pmc_simu_init_classic_pmc(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data, int prop_print_step,
                               void *filter_data, filter_func *pmc_filter, error **err) {
  pmc_simu_init_classic_importance(psim, target, pb, prop_data, prop_print_step, err);
  FWD(*err, );
  pmc_simu_init_pmc(psim, filter_data, pmc_filter, &update_prop_rb_void);
}

// ---- draw (pmc.c:256-298) -------------------------------------------------------------------------------------------
long simulate_mix_mvdens(pmc_simu *psim, mix_mvdens *m, gsl_rng *r, parabox *pb, error **err) {
  Side *s = side_for(psim, err);
  if (!s) return 0;
  // mix_mvdens_ran's lazy normalisation of the weights and cumulative table (mvdens.c:736-751), on the host struct
  if (m->init_cwght == 0) {
    double sum = 0.0;
    for (size_t i = 0; i < m->ncomp; ++i) sum += m->wght[i];
    if (std::fabs(sum - 1.0) > 1e-6) {
      if (sum == 0) { ERR_HERE(*err, mv_base - 5, "All weights to zero !! cannot sample", *err, nullptr); return 0; }
      for (size_t i = 0; i < m->ncomp; ++i) m->wght[i] *= 1.0 / sum;
    }
    m->cwght[0] = m->wght[0];
    for (size_t i = 1; i < m->ncomp; ++i) m->cwght[i] = m->cwght[i - 1] + m->wght[i];
    m->init_cwght = 1;
  }
  if (!push_proposal(s, m, err) || !push_box(s, pb, psim->ndim, err)) return 0;
  long nrej = 0;
  const int rc = pmcgpu_draw(s->ctx, seed_from(r), s->iter++, 0, &nrej);
  if (rc) { gpu_fail(s, rc, err, __func__); return 0; }
  s->host_stale = false;
  if (!pull_samples(s, psim, true, true, false, false, true, err)) return 0;
  return psim->nsamples;
}
long simulate_mix_mvdens_void(struct _pmc_simu_struct_ *psim, void *proposal, gsl_rng *r, parabox *pb, error **err) {
  const long n = simulate_mix_mvdens(psim, (mix_mvdens *)proposal, r, pb, err);
  FWD(*err, 0);
  return n;
}

// ---- importance weights (pmc.c:447-627) -------------------------------------------------------------------------------
size_t generic_get_importance_weight_and_deduced_verb(pmc_simu *psim, void *proposal_data,
                                                      posterior_log_pdf_func *proposal_log_pdf,
                                                      posterior_log_pdf_func *posterior_log_pdf,
                                                      retrieve_ded_func *retrieve_ded, void *target_data,
                                                      double t_iter_tempered, int quiet, error **err) {
  if (t_iter_tempered < 0) {  // pmc.c:508
    ERR_HERE(*err, pmc_infinite, "t_iter_tempered=%g should be > 0", *err, nullptr, t_iter_tempered);
    return (size_t)-1;
  }
  // the proposal must be a mix_mvdens, directly or behind distribution_lkl
  mix_mvdens *m = nullptr;
  if (proposal_log_pdf == &distribution_lkl) {
    distribution *pd = (distribution *)proposal_data;
    if (pd->log_pdf == &mix_mvdens_log_pdf_void) m = (mix_mvdens *)pd->data;
  } else if (proposal_log_pdf == &mix_mvdens_log_pdf_void) {
    m = (mix_mvdens *)proposal_data;
  }
  if (!m) { ERR_HERE(*err, pmc_incompat, "libpmcb200 evaluates mix_mvdens proposals only", *err, nullptr); return 0; }
  Side *s = side_for(psim, err);
  if (!s) return 0;
  if (!push_proposal(s, m, err)) return 0;
  distribution *tdist = (posterior_log_pdf == &distribution_lkl) ? (distribution *)target_data : nullptr;
  const int kind = tdist ? device_kind(tdist) : PMCGPU_TGT_HOST;
  if (!push_samples(s, psim, true, false, false, false, false, err)) return 0;
  long nok = 0;
  double maxW = 0, maxR = 0;
  int rc;
  if (kind != PMCGPU_TGT_HOST) {
    if (!push_target(s, tdist, err)) return 0;
    rc = pmcgpu_weights(s->ctx, t_iter_tempered, 0, &nok, &maxW, &maxR);
  } else {
    // opaque callback: per-sample host calls with the reference's error swallowing (ParameterErrorVerb, pmc.h:52-66)
    if (s->host_stale && !pull_samples(s, psim, true, false, false, false, false, err)) return 0;
    const long N = psim->nsamples;
    const int d = psim->ndim, nd = psim->n_ded;
    std::vector<double> post(N, 0.0);
    std::vector<short> ok(N, 1);
    if (!pmcb200::dropin_host_target_values(posterior_log_pdf, retrieve_ded, target_data, psim->X, psim->X_ded, N, d, nd, quiet,
                                            post.data(), ok.data(), err))
      return 0;
    rc = pmcgpu_set_target_host(s->ctx, d);
    if (!rc) rc = pmcgpu_weights_host_post(s->ctx, post.data(), ok.data(), t_iter_tempered, 0, &nok, &maxW, &maxR);
  }
  if (rc) { gpu_fail(s, rc, err, __func__); return 0; }
  if (!pull_samples(s, psim, false, false, true, true, true, err)) return 0;
  psim->maxW = maxW;
  psim->maxR = maxR;
  psim->isLog = 1;
  return (size_t)nok;
}
size_t generic_get_importance_weight_and_deduced(pmc_simu *psim, void *proposal_data, posterior_log_pdf_func *proposal_log_pdf,
                                                 posterior_log_pdf_func *posterior_log_pdf, retrieve_ded_func *retrieve_ded,
                                                 void *target_data, error **err) {
  // the reference passes t_iter_tempered = -1 here and therefore always fails (pmc.c:475-476 vs :508); the dead
  // code right after the test (pmc.c:509-511) shows the intent: no tempering
  const size_t r = generic_get_importance_weight_and_deduced_verb(psim, proposal_data, proposal_log_pdf, posterior_log_pdf,
                                                                  retrieve_ded, target_data, 1.0, 0, err);
  FWD(*err, 0);
  return r;
}
size_t generic_get_importance_weight(pmc_simu *psim, void *m, posterior_log_pdf_func *proposal_log_pdf,
                                     posterior_log_pdf_func *posterior_log_pdf, void *extra, error **err) {
  const size_t r = generic_get_importance_weight_and_deduced(psim, m, proposal_log_pdf, posterior_log_pdf, nullptr, extra, err);
  FWD(*err, 0);
  return r;
}
size_t get_importance_weight(pmc_simu *psim, mix_mvdens *m, posterior_log_pdf_func *posterior_log_pdf, void *extra, error **err) {
  const size_t r = generic_get_importance_weight_and_deduced(psim, m, &mix_mvdens_log_pdf_void, posterior_log_pdf, nullptr, extra, err);
  FWD(*err, 0);
  return r;
}
size_t get_importance_weight_plus_ded(pmc_simu *psim, mix_mvdens *m, posterior_log_pdf_func *posterior_log_pdf,
                                      retrieve_ded_func *retrieve_ded, void *extra, error **err) {
  const size_t r = generic_get_importance_weight_and_deduced(psim, m, &mix_mvdens_log_pdf_void, posterior_log_pdf, retrieve_ded, extra, err);
  FWD(*err, 0);
  return r;
}

// ---- normalise, perplexity, evidence (pmc.c:641-697, 1041-1105) ----------------------------------------------------------
double normalize_importance_weight(pmc_simu *psim, error **err) {
  if (psim->isLog == 0) {
    fprintf(stderr, "normalize_importance_weight: isLog=0, returning %g\n", psim->logSum);
    return psim->logSum;
  }
  Side *s = side_for(psim, err);
  if (!s) return 0;
  if (!push_samples(s, psim, false, false, true, false, true, err)) return 0;
  double logSum = 0;
  long cnt = 0;
  const int rc = pmcgpu_normalize(s->ctx, psim->maxW, &logSum, &cnt);
  if (rc) { gpu_fail(s, rc, err, __func__); return 0; }
  if (!s->host_stale && !pull_samples(s, psim, false, false, true, false, true, err)) return 0;
  psim->logSum = logSum;
  psim->isLog = 0;
  return std::exp(psim->logSum);
}
double perplexity_and_ess(pmc_simu *psim, int normalize, double *ess, error **err) {
  if (normalize == MC_NORM) {
    normalize_importance_weight(psim, err);
    FWD(*err, 0);
  }
  Side *s = side_for(psim, err);
  if (!s) return 0;
  if (!push_samples(s, psim, false, false, true, false, true, err)) return 0;
  double perp = 0, e = 0;
  long nfl = 0;
  const int rc = pmcgpu_perplexity_ess(s->ctx, &perp, &e, &nfl);
  if (rc) { gpu_fail(s, rc, err, __func__); return 0; }
  *ess = e;
  return perp;
}
double perplexity(pmc_simu *psim, int normalize, error **err) {
  double ess;
  const double p = perplexity_and_ess(psim, normalize, &ess, err);
  FWD(*err, 0);
  return p;
}
double evidence(pmc_simu *psim, double *ln_evi, error **err) {
  if (pmcb200_sync_host(psim, err) != 0) return 0;
  long n = 0;
  for (long i = 0; i < psim->nsamples; ++i) n += (psim->flg[i] != 0);
  if (n == 0) { ERR_HERE(*err, pmc_undef, "All points are flagged", *err, nullptr); return 0; }
  const double lne = psim->logSum - std::log((double)n);
  if (ln_evi) *ln_evi = lne;
  return std::exp(lne);
}

// ---- updates (pmc.c:1194-1579) --------------------------------------------------------------------------------------------
static void update_common(mix_mvdens *m, pmc_simu *psim, int hard, error **err) {
  if (psim->isLog == 1) {  // prepare_update, pmc.c:1541-1545
    normalize_importance_weight(psim, err);
    FWD(*err, );
  }
  Side *s = side_for(psim, err);
  if (!s) return;
  if (!push_proposal(s, m, err)) return;
  if (!push_samples(s, psim, true, true, true, true, true, err)) return;
  const int K = (int)m->ncomp, d = (int)m->ndim;
  std::vector<double> w(K), mean((size_t)K * d), L((size_t)K * d * d), detL(K);
  std::vector<int> alive(K);
  int retry = psim->retry;
  const int rc = hard ? pmcgpu_update_hard(s->ctx, psim->nsamples, &retry, w.data(), mean.data(), L.data(), detL.data(), alive.data())
                      : pmcgpu_update_rb(s->ctx, psim->maxR, psim->nsamples, &retry, w.data(), mean.data(), L.data(), detL.data(), alive.data());
  if (rc) { gpu_fail(s, rc, err, __func__); return; }
  psim->retry = retry;
  if (retry != 0) fprintf(stderr, "RETRY state -> %d\n", retry);
  unflatten(m, w.data(), mean.data(), L.data(), detL.data());
}
void update_prop_rb(mix_mvdens *m, pmc_simu *psim, error **err) { update_common(m, psim, 0, err); }
void update_prop_rb_void(void *m, pmc_simu *psim, error **err) {
  update_prop_rb((mix_mvdens *)m, psim, err);
  FWD(*err, );
}
void update_prop(mix_mvdens *m, pmc_simu *psim, error **err) { update_common(m, psim, 1, err); }

// ---- the two halves of the update as separate entry points (pmc.c:1471-1579), for callers that write their own update ----------
void prepare_update(mix_mvdens *m, size_t *count, pmc_simu *psim, double **pbuf, error **err) {
  double *buf = nullptr;
  if (psim->isLog == 1) {
    normalize_importance_weight(psim, err);
    FWD(*err, );
  }
  m->init_cwght = 0.0;
  {  // count[index] += 1 over flg == 1: histogram on the device, added to the caller's array as the reference does
    Side *s = side_for(psim, err);
    if (!s) return;
    if (!push_proposal(s, m, err)) return;   // sizes the histogram (and factorises, as line 1566 would below)
    if (!push_samples(s, psim, false, true, false, false, true, err)) return;
    std::vector<size_t> c(m->ncomp);
    const int rc = pmcgpu_count_indices(s->ctx, c.data());
    if (rc) { gpu_fail(s, rc, err, __func__); return; }
    for (size_t k = 0; k < m->ncomp; ++k) count[k] += c[k];
  }
  mix_mvdens_cholesky_decomp(m, err);
  FWD(*err, );
  if (psim->retry == 0) {
    const size_t dd = m->ndim * m->ndim;
    buf = (double *)xmalloc(sizeof(double) * m->ncomp * dd, err);
    if (!buf) return;
    for (size_t i = 0; i < m->ncomp; ++i) memcpy(buf + i * dd, m->comp[i]->std, sizeof(double) * dd);
  }
  *pbuf = buf;
}
void cleanup_after_update(mix_mvdens *m, size_t *count, double minwgt, pmc_simu *psim, double *buf, error **err) {
  const int K = (int)m->ncomp, d = (int)m->ndim;
  const size_t dd = (size_t)d * d;
  std::vector<double> w(m->wght, m->wght + K), mean((size_t)K * d), L(K * dd), detL(K);
  for (int k = 0; k < K; ++k) {
    memcpy(&mean[(size_t)k * d], m->comp[k]->mean, d * sizeof(double));
    memcpy(&L[k * dd], m->comp[k]->std, dd * sizeof(double));
    detL[k] = m->comp[k]->detL;
  }
  std::vector<double> zero;
  const double *snap = buf;
  if (!snap) { zero.assign(K * dd, 0.0); snap = zero.data(); }  // only read when retry == 0, where the reference has a buffer
  int retry = psim->retry;
  const int rc = pmcb200::host_cleanup_after_update(K, d, count, minwgt, &retry, snap, w.data(), mean.data(), L.data(), detL.data());
  psim->retry = retry;
  if (retry != 0) fprintf(stderr, "RETRY state -> %d\n", retry);
  for (int k = 0; k < K; ++k) {
    mvdens *c = m->comp[k];
    m->wght[k] = w[k];
    memcpy(c->mean, &mean[(size_t)k * d], d * sizeof(double));
    memcpy(c->std, &L[k * dd], dd * sizeof(double));
    if (w[k] > 0) { c->detL = detL[k]; c->chol = 1; }
  }
  if (rc) ERR_HERE(*err, rc, "Sum of weight is <= 0 after the update", *err, nullptr);
}
int double_cmp(const void *av, const void *bv) {
  const double a = *(const double *)av, b = *(const double *)bv;
  return (a < b) ? -1 : (a > b) ? 1 : 0;
}
// print_step (tools.c:301-313) for every sample
void pmc_simu_print(FILE *where, pmc_simu *psim, double nrm) {
  error *e = initError(), **err = &e;
  const int rc = pmcb200_sync_host(psim, err);
  endError(err);
  if (rc != 0) return;
  FILE *rf = where ? where : stdout;
  for (long i = 0; i < psim->nsamples; ++i) {
    fprintf(rf, "%1d %10g", (int)psim->flg[i], nrm * psim->weights[i]);
    for (int j = 0; j < psim->ndim; ++j) fprintf(rf, " %10g", psim->X[(size_t)i * psim->ndim + j]);
    fprintf(rf, "\n");
  }
}

// ---- the iteration (pmc.c:175-219) -------------------------------------------------------------------------------------------
size_t pmc_simu_importance(pmc_simu *psim, gsl_rng *r, error **err) {
  if (!psim->proposal || !psim->proposal->data) { ERR_HERE(*err, pmc_allocate, "proposal undefined", *err, nullptr); return 0; }
  if (!psim->target || !psim->target->data) { ERR_HERE(*err, pmc_allocate, "target undefined", *err, nullptr); return 0; }
  const bool fast = psim->proposal->simulate == &simulate_mix_mvdens_void && psim->proposal->log_pdf == &mix_mvdens_log_pdf_void &&
                    device_kind(psim->target) != PMCGPU_TGT_HOST;
  if (!fast) {
    psim->proposal->simulate(psim, psim->proposal->data, r, psim->pb, err);
    FWD(*err, 0);
    const size_t nok = generic_get_importance_weight_and_deduced(psim, psim->proposal, &distribution_lkl, &distribution_lkl,
                                                                 &distribution_retrieve, psim->target, err);
    FWD(*err, 0);
    return nok;
  }
  Side *s = side_for(psim, err);
  if (!s) return 0;
  mix_mvdens *m = (mix_mvdens *)psim->proposal->data;
  if (!push_proposal(s, m, err) || !push_target(s, psim->target, err) || !push_box(s, psim->pb, psim->ndim, err)) return 0;
  long nok = 0, nrej = 0;
  double maxW = 0, maxR = 0;
  const bool sink = !resident();
  if (sink) pmcgpu_set_host_sink(s->ctx, psim->X, psim->indices, psim->weights, psim->log_rho, psim->flg);
  const int rc = pmcgpu_draw_weights(s->ctx, seed_from(r), s->iter++, 0, 1.0, &nok, &maxW, &maxR, &nrej);
  pmcgpu_set_host_sink(s->ctx, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc) { gpu_fail(s, rc, err, __func__); return 0; }
  psim->maxW = maxW;
  psim->maxR = maxR;
  psim->isLog = 1;
  s->host_stale = !sink;
  return (size_t)nok;
}

// ---- single-process multi-GPU iteration --------------------------------------------------------------------------------------
// An unmodified (non-MPI) pmclib driver uses every GPU of the node with PMCB200_GPUS=n: the sample is cut into n shards
// (pmc_mpi.c:344-352 layout), one host thread per GPU runs pmcgpu_pmc_step on its shard, the small all-reduces meet in a
// host rendezvous, every shard is copied straight into its slice of the caller's arrays.
int gpus_wanted() {
  static int n = -1;
  if (n < 0) {
    n = 1;
    if (const char *e = getenv("PMCB200_GPUS")) n = atoi(e);
    const int have = pmcgpu_device_count();
    if (n > have) n = have;
    if (n < 1) n = 1;
  }
  return n;
}

double pmc_step_multi(Side *s, pmc_simu *psim, mix_mvdens *m, gsl_rng *r, error **err) {
  const int G = gpus_wanted(), K = (int)m->ncomp, d = (int)m->ndim;
  const long N = psim->nsamples;
  if ((int)s->shard.size() != G) {
    s->shard.assign(1, s->ctx);
    const int dev0 = device_ordinal(), have = pmcgpu_device_count();
    for (int g = 1; g < G; ++g) {
      pmcgpu_ctx *c = nullptr;
      const int rc = pmcgpu_create((dev0 + g) % have, &c);
      if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"libpmcb200: cannot open GPU %d", *err, nullptr, (dev0 + g) % have); s->shard.resize(1); return 0; }
      s->shard.push_back(c);
    }
    s->rv.n = G;
    s->rv.bufs.assign(G, nullptr);
    s->rvrank.resize(G);
    for (int g = 0; g < G; ++g) {
      s->rvrank[g] = RvRank{&s->rv, g};
      pmcgpu_comm_set_callback(s->shard[g], &rv_allreduce, &s->rvrank[g], g, G);
    }
  }
  std::vector<long> first(G), len(G);
  for (int g = 0; g < G; ++g) {
    pmcgpu_shard_range(N, g, G, &first[g], &len[g]);
    Side tmp;
    tmp.ctx = s->shard[g];
    const int rc = pmcgpu_resize(tmp.ctx, len[g], d);
    if (rc) { gpu_fail(&tmp, rc, err, __func__); return 0; }
    if (!push_proposal(&tmp, m, err) || !push_target(&tmp, psim->target, err) || !push_box(&tmp, psim->pb, psim->ndim, err)) return 0;
    const int *fil = psim->pmc_filter ? (const int *)psim->filter_data : nullptr;
    pmcgpu_set_filter_histogram(tmp.ctx, fil ? fil[0] : 0, fil ? fil[1] : 0, fil ? fil[2] : 0);
  }
  s->N = len[0];  // shard[0] is the context the phase-by-phase entry points use: side_for() sizes it back when needed
  const uint64_t seed = seed_from(r);
  const uint32_t iter = s->iter++;
  std::vector<int> rcs(G, 0), retry(G, psim->retry);
  std::vector<pmcgpu_step_result> res(G);
  std::vector<double> w(K), mean((size_t)K * d), L((size_t)K * d * d), detL(K);
  std::vector<int> alive(K);
  s->rv.broken = false;
  s->rv.arrived = 0;
  std::vector<std::thread> th;
  for (int g = 0; g < G; ++g)
    th.emplace_back([&, g]() {
      const bool lead = g == 0;
      pmcgpu_set_host_sink(s->shard[g], psim->X + (size_t)first[g] * d, psim->indices + first[g], psim->weights + first[g],
                           psim->log_rho + first[g], psim->flg + first[g]);
      const int rc = pmcgpu_pmc_step(s->shard[g], seed, iter, first[g], N, 1, &retry[g], &res[g], lead ? w.data() : nullptr,
                                     lead ? mean.data() : nullptr, lead ? L.data() : nullptr, lead ? detL.data() : nullptr,
                                     lead ? alive.data() : nullptr);
      pmcgpu_set_host_sink(s->shard[g], nullptr, nullptr, nullptr, nullptr, nullptr);
      if (rc) {  // release the ranks that may be waiting for this one
        std::lock_guard<std::mutex> lk(s->rv.m);
        s->rv.broken = true;
        s->rv.cv.notify_all();
      }
      rcs[g] = rc;
    });
  for (auto &t : th) t.join();
  for (int g = 0; g < G; ++g) pmcgpu_set_filter_histogram(s->shard[g], 0, 0, 0);
  for (int pass = 0; pass < 2; ++pass)  // the rank that failed first, not the ranks its failure released from the rendezvous
    for (int g = 0; g < G; ++g)
      if (rcs[g] && (pass == 1 || rcs[g] != PMCGPU_ERR_NCCL)) { Side tmp; tmp.ctx = s->shard[g]; gpu_fail(&tmp, rcs[g], err, __func__); return 0; }
  psim->maxW = res[0].maxW;
  psim->maxR = res[0].maxR;
  psim->logSum = res[0].logSum;
  psim->isLog = 0;
  psim->retry = retry[0];
  if (psim->retry != 0) fprintf(stderr, "RETRY state -> %d\n", psim->retry);
  unflatten(m, w.data(), mean.data(), L.data(), detL.data());
  s->host_stale = false;
  return res[0].perplexity;
}

double pmc_simu_pmc_step(pmc_simu *psim, gsl_rng *r, error **err) {
  if (!psim->pmc_update) { ERR_HERE(*err, pmc_allocate, "update function undefined", *err, nullptr); return 0; }
  const bool fast = psim->proposal && psim->target && psim->proposal->simulate == &simulate_mix_mvdens_void &&
                    psim->proposal->log_pdf == &mix_mvdens_log_pdf_void && device_kind(psim->target) != PMCGPU_TGT_HOST &&
                    (psim->pmc_filter == nullptr || psim->pmc_filter == &filter_histogram) &&
                    psim->pmc_update == &update_prop_rb_void;
  if (!fast) {  // the reference's own sequence, phase by phase (opaque target, user filter or user update)
    pmc_simu_importance(psim, r, err);
    FWD(*err, 0);
    if (psim->pmc_filter) { psim->pmc_filter(psim, psim->filter_data, err); FWD(*err, 0); }
    normalize_importance_weight(psim, err);
    FWD(*err, 0);
    psim->pmc_update(psim->proposal->data, psim, err);
    FWD(*err, 0);
    const double perp = perplexity(psim, MC_UNORM, err);
    FWD(*err, 0);
    return perp;
  }
  Side *s = side_for(psim, err);
  if (!s) return 0;
  mix_mvdens *m = (mix_mvdens *)psim->proposal->data;
  if (gpus_wanted() > 1 && !resident()) return pmc_step_multi(s, psim, m, r, err);
  if (!push_proposal(s, m, err) || !push_target(s, psim->target, err) || !push_box(s, psim->pb, psim->ndim, err)) return 0;
  const int K = (int)m->ncomp, d = (int)m->ndim;
  std::vector<double> w(K), mean((size_t)K * d), L((size_t)K * d * d), detL(K);
  std::vector<int> alive(K);
  int retry = psim->retry;
  pmcgpu_step_result res;
  const int *fil = psim->pmc_filter ? (const int *)psim->filter_data : nullptr;  // the shipped filter runs on the device
  pmcgpu_set_filter_histogram(s->ctx, fil ? fil[0] : 0, fil ? fil[1] : 0, fil ? fil[2] : 0);
  // host arrays are filled while the iteration runs (second stream), unless the sample is to stay resident in HBM
  const bool sink = !resident();
  if (sink) pmcgpu_set_host_sink(s->ctx, psim->X, psim->indices, psim->weights, psim->log_rho, psim->flg);
  const int rc = pmcgpu_pmc_step(s->ctx, seed_from(r), s->iter++, 0, psim->nsamples, 1, &retry, &res, w.data(), mean.data(),
                                 L.data(), detL.data(), alive.data());
  pmcgpu_set_host_sink(s->ctx, nullptr, nullptr, nullptr, nullptr, nullptr);
  pmcgpu_set_filter_histogram(s->ctx, 0, 0, 0);
  if (rc) { gpu_fail(s, rc, err, __func__); return 0; }
  psim->maxW = res.maxW;
  psim->maxR = res.maxR;
  psim->logSum = res.logSum;
  psim->isLog = 0;
  psim->retry = retry;
  if (retry != 0) fprintf(stderr, "RETRY state -> %d\n", retry);
  // mix_mvdens_ran would have normalised the weights before the draw; the update then replaces everything
  unflatten(m, w.data(), mean.data(), L.data(), detL.data());
  s->host_stale = !sink;
  return res.perplexity;
}

// ---- binary dump (pmc.c:342-371) ------------------------------------------------------------------------------------------------
void pmc_simu_dump(FILE *where, pmc_simu *psim, error **err) {
  if (pmcb200_sync_host(psim, err) != 0) return;
  const size_t N = psim->nsamples;
  bool ok = fwrite("PMC_SIMD", 1, 8, where) == 8;
  ok = ok && fwrite(&psim->nsamples, sizeof(long), 1, where) == 1 && fwrite(&psim->ndim, sizeof(int), 1, where) == 1 &&
       fwrite(&psim->n_ded, sizeof(int), 1, where) == 1 && fwrite(&psim->isLog, sizeof(int), 1, where) == 1 &&
       fwrite(&psim->logSum, sizeof(double), 1, where) == 1;
  ok = ok && fwrite(psim->flg, sizeof(short), N, where) == N && fwrite(psim->weights, sizeof(double), N, where) == N &&
       fwrite(psim->log_rho, sizeof(double), N, where) == N &&
       fwrite(psim->X, sizeof(double), N * psim->ndim, where) == N * psim->ndim &&
       fwrite(psim->X_ded, sizeof(double), N * psim->n_ded, where) == N * psim->n_ded;
  if (!ok) ERR_HERE(*err, pmc_io, "Cannot write to file", *err, nullptr);
}

// ---- weight post-processing either side of the iteration (pmc.c:699-770, 1108-1120, 306-325) -------------------------------------
int *filter_histogram_init(int nbins, int clipleft, int clipright, error **err) {
  int *self = (int *)malloc(sizeof(int) * 3);
  if (!self) { ERR_HERE(*err, pmc_allocate, "Cannot allocate the filter", *err, nullptr); return nullptr; }
  self[0] = nbins;
  self[1] = clipleft;
  self[2] = clipright;
  return self;
}
void filter_histogram(pmc_simu *psim, void *filter_data, error **err) {
  const int *fil = (const int *)filter_data;
  if (psim->isLog == 0) { ERR_HERE(*err, pmc_isLog, "can only filter un-normalized data", *err, nullptr); return; }
  Side *s = side_for(psim, err);
  if (!s) return;
  if (!push_samples(s, psim, false, false, true, false, true, err)) return;
  long nf = 0;
  const int rc = pmcgpu_filter_histogram(s->ctx, fil[0], fil[1], fil[2], 0, &nf);
  if (rc) { gpu_fail(s, rc, err, __func__); return; }
  if (!s->host_stale) pull_samples(s, psim, false, false, false, false, true, err);
}
// mean of the a-th parameter; the arrays are plain host arrays (no pmc_simu), so they go through a scratch context
double mean_from_psim(double *X, double *w, short *flg, int nsamples, int ndim, int a) {
  static pmcgpu_ctx *scratch = nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!scratch && pmcgpu_create(device_ordinal(), &scratch) != 0) return std::nan("");
  double m = std::nan("");
  if (pmcgpu_resize(scratch, nsamples, ndim) == 0 && pmcgpu_upload(scratch, X, nullptr, w, nullptr, flg) == 0)
    pmcgpu_weighted_mean(scratch, a, &m);
  return m;
}
void out_pmc_simu(const char *name, const pmc_simu *psim, error **err) {
  if (pmcb200_sync_host(const_cast<pmc_simu *>(psim), err) != 0) return;
  FILE *out = fopen(name, "w");
  if (!out) { ERR_HERE(*err, pmc_io, "Cannot open file", *err, nullptr); return; }
  for (long i = 0; i < psim->nsamples; ++i) {
    if (psim->flg[i] == 0) continue;
    for (int j = 0; j < psim->ndim; ++j) fprintf(out, "% .5e ", psim->X[j + i * psim->ndim]);
    for (int j = 0; j < psim->n_ded; ++j) fprintf(out, "% .5e ", psim->X_ded[j + i * psim->n_ded]);
    fprintf(out, "% .5e\n", psim->weights[i]);
  }
  fclose(out);
}

void clip_weights(pmc_simu *psim, int n, FILE *OUT, error **err) {
  if (n == 0) return;
  Side *s = side_for(psim, err);
  if (!s) return;
  if (!push_samples(s, psim, false, false, true, false, true, err)) return;
  double T = 0, S = 0;
  long nc = 0;
  const int rc = pmcgpu_clip_weights(s->ctx, n, &T, &nc, &S);
  if (OUT) fprintf(OUT, "clip_weights: setting %d weights larger than %g to zero (%ld samples)\n", n, T, nc);
  if (rc) { gpu_fail(s, rc, err, __func__); return; }
  if (!s->host_stale) pull_samples(s, psim, false, false, true, false, true, err);
}

// resampling after the last iteration (pmc.c:1585-1626): results are malloc'ed, owned by the caller as in the reference
size_t *sample_from_pmc_simu(const pmc_simu *cpsim, const gsl_rng *rng, error **err) {
  pmc_simu *psim = const_cast<pmc_simu *>(cpsim);
  Side *s = side_for(psim, err);
  if (!s) return nullptr;
  if (!push_samples(s, psim, false, false, true, false, false, err)) return nullptr;
  size_t *isample = (size_t *)xmalloc(sizeof(size_t) * psim->nsamples, err);
  if (!isample) return nullptr;
  const int rc = pmcgpu_sample_indices(s->ctx, seed_from(const_cast<gsl_rng *>(rng)), s->iter++, isample);
  if (rc) { gpu_fail(s, rc, err, __func__); free(isample); return nullptr; }
  return isample;
}
unsigned int *resample_residual(const pmc_simu *cpsim, const gsl_rng *rng, error **err) {
  pmc_simu *psim = const_cast<pmc_simu *>(cpsim);
  Side *s = side_for(psim, err);
  if (!s) return nullptr;
  if (!push_samples(s, psim, false, false, true, false, false, err)) return nullptr;
  unsigned int *n = (unsigned int *)xmalloc(sizeof(unsigned int) * psim->nsamples, err);
  if (!n) return nullptr;
  const int rc = pmcgpu_resample_residual(s->ctx, seed_from(const_cast<gsl_rng *>(rng)), s->iter++, n);
  if (rc) { gpu_fail(s, rc, err, __func__); free(n); return nullptr; }
  return n;
}

// ---- liberrorio helpers that target likelihoods written against pmclib commonly call (io.c:103-131) -----------------------
FILE *fopen_err(const char *fname, const char *mode, error **err) {
  FILE *fp = fopen(fname, mode);
  if (!fp) *err = newErrorVA(-101 /* err_io */, __func__, (char *)"", (char *)"Cannot open file '%s' (mode \"%s\")", *err, nullptr, fname, mode);
  return fp;
}
void *malloc_err(size_t sz, error **err) {
  if (sz <= 0) { *err = newErrorVA(err_allocate, __func__, (char *)"", (char *)"Size %ld has to be positive", *err, nullptr, (long)sz); return nullptr; }
  void *res = malloc(sz);
  if (!res) *err = newErrorVA(err_allocate, __func__, (char *)"", (char *)"Cannot allocate %ld bytes", *err, nullptr, (long)sz);
  return res;
}
void *calloc_err(size_t nl, size_t sz1, error **err) {
  if (nl <= 0) { *err = newErrorVA(err_allocate, __func__, (char *)"", (char *)"Size %ld has to be positive", *err, nullptr, (long)nl); return nullptr; }
  void *res = calloc(nl, sz1);
  if (!res) *err = newErrorVA(err_allocate, __func__, (char *)"", (char *)"Cannot allocate %ld objects of %ld bytes", *err, nullptr, (long)nl, (long)sz1);
  return res;
}

// ---- additions ------------------------------------------------------------------------------------------------------------------
int pmcb200_register_target(distribution *dist, int kind) {
  if (!dist || kind < PMCGPU_TGT_BANANA || kind > PMCGPU_TGT_MIX) return PMCGPU_ERR_USAGE;
  std::lock_guard<std::mutex> lk(g_mu);
  g_registered[dist] = kind;
  return 0;
}
void pmcb200_set_resident(int on) { g_resident = on ? 1 : 0; }
int pmcb200_sync_host(pmc_simu *psim, error **err) {
  Side *s = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_side.find(psim);
    if (it != g_side.end()) s = &it->second;
  }
  if (!s || !s->host_stale) return 0;
  if (!pull_samples(s, psim, true, true, true, true, true, err)) return -1;
  s->host_stale = false;
  return 0;
}
void *pmcb200_context(pmc_simu *psim, error **err) {
  Side *s = side_for(psim, err);
  return s ? (void *)s->ctx : nullptr;
}

}  // extern "C"
