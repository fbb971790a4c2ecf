// pmc_api.cu — the C-ABI of libpmcb200.so (include/pmcb200.h): context, HBM-resident sample store, kernel
// orchestration for one PMC iteration, the optional NCCL all-reduce of the sufficient statistics.
// Host code here is orchestration only; everything that scales with the number of samples runs in the kernels of
// pmc_fused.cu / pmc_generic.cu.  There is no CPU fallback: without a CUDA device pmcgpu_create fails.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>
#include <dlfcn.h>
#include "pmc_kernels.h"
#include "pmc_hostonly.h"

using namespace pmcb200;

namespace {

constexpr int kErrNegWeight = -6005, kErrUndef = -6008, kErrTooManySteps = -6011, kErrNoSample = -6017,
              kErrMvNegWeight = -705, kErrMvOutOfBound = -703;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostMix {
  int K = 0, d = 0;
  std::vector<double> wght, mean, L, detL, logdet, gam, norm, cw;
  std::vector<int> df;
  std::vector<size_t> band;
  bool student = false;
  int n_alive() const { int n = 0; for (double w : wght) n += (w > 0.0); return n; }
};

// pinned host staging that belongs to ONE upload site: the copy out of it is asynchronous; before the site writes the buffer
// again it waits for the event recorded behind its previous copy (normally long complete), not for the whole stream
struct Pinned {
  double *p = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
  cudaError_t ensure(size_t doubles) {
    if (pending) { cudaEventSynchronize(ev); pending = false; }
    if (doubles <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);  // waits for copies in flight
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost((void **)&p, doubles * sizeof(double));
    if (e == cudaSuccess) cap = doubles;
    return e;
  }
  cudaError_t copied(cudaStream_t st) {  // call right after enqueueing the copy out of p
    if (!ev) { cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    pending = true;
    return cudaEventRecord(ev, st);
  }
  void release() {
    if (pending && ev) cudaEventSynchronize(ev);
    pending = false;
    if (ev) cudaEventDestroy(ev);
    ev = nullptr;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct DevMixBuf {
  DevBuf buf;
  MixDev view{};
  Pinned pin;
};

// ---- NCCL through dlopen (no link-time dependency) ---------------------------------------------------------------
struct IdByValue { char internal[PMCGPU_NCCL_ID_BYTES]; };  // ncclUniqueId is passed by value
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, IdByValue, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool load_nccl(std::string &why) {
  if (g_nccl.h) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) { why = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
  g_nccl.GetUniqueId = (int (*)(void *))dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void **, int, IdByValue, int))dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    why = "libnccl.so.2 lacks the expected symbols";
    g_nccl.h = nullptr;
    return false;
  }
  return true;
}
constexpr int kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

}  // namespace

struct pmcgpu_ctx {
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // model
  HostMix prop;
  DevMixBuf dprop;
  DevBuf dprop_pack;
  Pinned pin_ppack, pin_tpack;
  int prop_pack_doubles = 0;
  std::vector<double> prop_pack_host, tgt_pack_host, draw_pack_padded;
  int tkind = PMCGPU_TGT_NONE, td = 0;
  double tb = 0, tsig1 = 1;
  HostMix tmix;
  DevMixBuf dtmix;
  DevBuf dt_pack;
  int t_pack_doubles = 0;
  // combined target (PMCGPU_TGT_COMBINE): terms in device memory, evaluated into hostpost before the weight kernels
  struct CombTerm { int kind = 0; std::vector<int> dim; double b = 0, sig1 = 1; HostMix mix; DevMixBuf dmix; };
  std::vector<CombTerm> comb;
  DevBuf dcomb, dcombdim;
  int comb_kmax = 1;
  int nfull = 0;
  std::vector<int> is_def;
  std::vector<double> def_val;
  DevBuf ddef;
  int nfaces = 0, box_d = 0;
  std::vector<double> faces;
  DevBuf dfaces;
  bool box_slab = false;      // slab box that the fused small-d kernels can test (d <= FUSED_MAXD)
  bool box_slab_any = false;  // slab box of any dimension (tensor-path draw)
  double lo_thr[PMCB200_MAXD], hi_thr[PMCB200_MAXD];
  DevBuf dthr, dredo;
  // store
  long N = 0;
  int d = 0;
  DevBuf X, W, R, Idx, Flg;
  // scratch
  DevBuf ld, blkmax, partials, stats, small, rbg, rbw, rbacc, rbpart, newmean, hostpost, hostok, tmpx, tmpo;
  double *hstage = nullptr;  // pinned
  size_t hstage_doubles = 0;
  // timing of the dominant kernel (CUDA events on the launching stream)
  int timing = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, evd = nullptr;
  double fused_ms = 0.0, stats_ms = 0.0, draw_ms = 0.0;
  int split_draw = 1;   // draw as its own launch: smaller code and registers for both kernels (+6 % measured)
  int split_stats = 1;  // statistics as their own DMMA kernel (the single-kernel MODE 2 is 15 % slower)
  DevBuf resp, partials2, dscal, zig;
  long fused_launches = 0;
  // options
  int fused_version = 0;  // 0 = newest kernel that fits, 1 = pmc_fused.cu only, 2 = pmc_fused2.cu only
  int force_generic = 0, spl = 1, max_attempts = 10000, force_precise = 0;
  double guard_ratio = 200.0;
  // large-d tensor path (pmc_mma.cu)
  struct MmaMix { DevBuf pack, lpack, kmap, dfj; int J = 0; bool valid = false, dirty = false; std::vector<int> alive; double inv_discrepancy = 0.0; };
  MmaMix mprop, mtgt;
  DevBuf Q, Qt, mpart, mstats, mpivot, dperm, dtiles, dcnt, xc;
  HostMix trial;  // scratch of the update (kept to reuse its storage)
  std::vector<double> hW, hG, hs1, hS2;
  DevBuf wpart, cdf, cdfsum, rsout;
  long q_stride = 0;
  int mma_store_ld = 1;     // 0 inside an importance-only iteration: the tensor path's tail does not write the component log-densities back
  bool q_resident = false;  // Q holds the component log-densities of all N samples of the store (left by the weights pass)
  int use_mma = 1, mma_min_d = 2;  // any shape without a fused small-d instantiation goes to the tensor path (30x the literal kernels at d = 7..16)
  double t_temper = 1.0;  // t_iter_tempered of pmcgpu_pmc_step (pmc.c:570: weight = t * post - rloc)
  double mma_inv_tol = 1.0e-13;  // conditioning guard of the explicit inverse factors (pack_mma), option "mma_inv_tol"
  long mma_max_chunk = 0;  // test knob: cap of the samples whitened per pass (0 = as many as memory allows)
  int filt_nbins = 0, filt_cl = 0, filt_cr = 0;  // filter_histogram inside the iteration (0 = off)
  // host sink: arrays of the caller that pmcgpu_pmc_step fills while the iteration is still running
  struct Sink { double *X = nullptr; size_t *idx = nullptr; double *w = nullptr, *lr = nullptr; short *flg = nullptr; bool on = false; } sink;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr;
  bool sink_active = false;  // inside a drawing iteration with a sink attached
  // indices travel to the sink as bytes (1 instead of 8 per sample over PCIe) and are widened into the caller's size_t array by
  // a host thread while the X copy is still in flight (option "sink_idx8", on by default for K <= 256)
  int sink_idx8 = 1;
  DevBuf didx8;
  unsigned char *hidx8 = nullptr;
  size_t hidx8_cap = 0;
  cudaEvent_t ev_idx8 = nullptr;
  std::thread widen;
  double sink_bytes = 0.0;  // device-to-host bytes the sink copies of the current / last call moved (info "sink_bytes")
  // comm
  int rank = 0, nranks = 1;
  void *nccl_comm = nullptr;
  pmcgpu_allreduce_fn cb = nullptr;
  void *cb_user = nullptr;
  DevBuf dcoll, dshare;
  // third-generation small-d iteration (pmc_iter3.cu)
  int gen3 = 1;                 // option "gen3": 0 = second-generation kernels only
  int w3_variant = 3;           // option "w3_variant": 3 = row-scaled solve (default, guarded by gen3_max_ratio), 0 = exact subtraction, 1 / 2 = launch-shape experiments
  int draw3 = 0;                // option "draw3": 1 = grouped draw kernel k_draw3 even where the second-generation draw exists
  double gen3_max_ratio = 100;  // largest |mu' / L_jj| the pre-scaled solve is trusted with (option "gen3_max_ratio")
  bool i3_dirty = true;         // packs must be rebuilt (proposal or target changed)
  double i3_ratio = 0.0;
  double i3_pre_ratio = 0.0;    // largest |mu'_j / L_jj| over the live components about the common shift
  bool i3_pre = false;          // the proposal's whitening pack is in row-scaled form (w3_variant 3 and the ratio guard passed)
  double i3_shift[FUSED_MAXD] = {0};
  std::vector<double> i3_hw, i3_ht, i3_hd, i3_hcw;
  DevBuf i3_gpack, i3_wpart, i3_scal, i3_coll, i3_spart, i3_out, i3_gath;
};

namespace {

int fail(pmcgpu_ctx *c, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CU(call)                                                                                           \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) return fail(c, PMCGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define RC(call) do { int r_ = (call); if (r_) return r_; } while (0)

int stage(pmcgpu_ctx *c, size_t doubles) {
  if (doubles <= c->hstage_doubles) return 0;
  if (c->hstage) cudaFreeHost(c->hstage);
  c->hstage = nullptr;
  c->hstage_doubles = 0;
  CU(cudaMallocHost((void **)&c->hstage, doubles * sizeof(double)));
  c->hstage_doubles = doubles;
  return 0;
}

// small device block: counters[CNT_NUM] (ull) | count_k[64] (ull) | first_vals[4] (double) = 8+64+4 = 76 x 8 bytes
constexpr int SM_COUNTERS = 0, SM_COUNTK = CNT_NUM, SM_FIRST = CNT_NUM + 64, SM_LEN = CNT_NUM + 64 + 4;

int upload_mix(pmcgpu_ctx *c, const HostMix &m, DevMixBuf &o) {
  const size_t K = m.K, d = m.d;
  const size_t nd = K * 6 + K * d + K * d * d;  // alpha,cw,logdet,gam,norm + pad | mean | L
  const size_t bytes = nd * sizeof(double) + K * sizeof(int);
  CU(o.buf.ensure(bytes));
  CU(o.pin.ensure(nd + K));
  double *h = o.pin.p;
  size_t off = 0;
  auto put = [&](const std::vector<double> &v) { memcpy(h + off, v.data(), v.size() * sizeof(double)); size_t r = off; off += v.size(); return r; };
  const size_t oa = put(m.wght), ocw = put(m.cw), old = put(m.logdet), og = put(m.gam), on = put(m.norm);
  off += K;
  const size_t om = put(m.mean), oL = put(m.L);
  memcpy(h + off, m.df.data(), K * sizeof(int));
  CU(cudaMemcpyAsync(o.buf.p, h, bytes, cudaMemcpyHostToDevice, c->stream));  // no host synchronisation: see Pinned
  CU(o.pin.copied(c->stream));
  double *b = o.buf.as<double>();
  o.view = MixDev{(int)K, (int)d, b + oa, b + ocw, b + om, b + oL, b + old, b + og, b + on, reinterpret_cast<const int *>(b + nd)};
  return 0;
}

// derived quantities of a mixture: log det, Student-t constants (mvdens.c:376-377), cumulative weights (mvdens.c:748-750)
void derive_mix(HostMix &m) {
  const int K = m.K, d = m.d;
  m.logdet.assign(K, 0.0);
  m.gam.assign(K, 0.0);
  m.norm.assign(K, 0.0);
  m.cw.assign(K, 0.0);
  m.student = false;
  for (int k = 0; k < K; ++k) {
    m.logdet[k] = std::log(m.detL[k]);
    if (m.df[k] != -1) {
      m.student = true;
      const size_t n = (size_t)m.df[k], p = (size_t)d;
      m.gam[k] = lgamma(0.5 * (n + p));
      m.norm[k] = lgamma(0.5 * n) + 0.5 * p * std::log(n * kPi);
    }
  }
  double cacc = 0.0;
  for (int k = 0; k < K; ++k) { cacc = (k == 0) ? m.wght[0] : cacc + m.wght[k]; m.cw[k] = cacc; }
}

int fill_mix(pmcgpu_ctx *c, HostMix &m, int K, int d, const double *wght, const double *mean, const double *L,
             const double *detL, const int *df) {
  if (K < 1 || d < 1 || d > PMCB200_MAXD) return fail(c, PMCGPU_ERR_USAGE, "mixture shape K=%d d=%d not supported (d <= %d)", K, d, PMCB200_MAXD);
  m.K = K;
  m.d = d;
  m.wght.assign(wght, wght + K);
  m.mean.assign(mean, mean + (size_t)K * d);
  m.L.assign(L, L + (size_t)K * d * d);
  m.detL.resize(K);
  for (int k = 0; k < K; ++k) {
    if (detL) m.detL[k] = detL[k];
    else { double det = 1.0; for (int a = 0; a < d; ++a) det *= L[(size_t)k * d * d + (size_t)a * d + a]; m.detL[k] = det; }
  }
  m.df.assign(K, -1);
  if (df) m.df.assign(df, df + K);
  derive_mix(m);
  return 0;
}

int pack_small(pmcgpu_ctx *c, const HostMix &m, DevBuf &out, int &ndoubles, std::vector<double> &hostcopy, Pinned &pin) {
  ndoubles = 0;
  hostcopy.clear();
  if (m.d > FUSED_MAXD || m.student) return 0;
  const int n = m.K * fused_comp_stride(m.d);
  CU(out.ensure((size_t)n * sizeof(double)));
  CU(pin.ensure(n));
  fused_pack_mixture(m.K, m.d, m.wght.data(), m.mean.data(), m.L.data(), m.logdet.data(), pin.p);
  hostcopy.assign(pin.p, pin.p + n);
  CU(cudaMemcpyAsync(out.p, pin.p, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));  // no host synchronisation: see Pinned
  CU(pin.copied(c->stream));
  ndoubles = n;
  return 0;
}

// inverse factors of the live components in DMMA fragment order (pmc_mma.cu); the O(K d^3) inversion runs on a few
// host threads because at d = 128, K = 64 it is 22 MFLOP of extended-precision work per iteration
int pack_mma(pmcgpu_ctx *c, const HostMix &m, pmcgpu_ctx::MmaMix &o) {
  o.valid = false;
  o.J = 0;
  c->q_resident = false;
  if (!c->use_mma || c->force_generic || m.K == 0 || m.d < c->mma_min_d || mma_nt(m.d) == 0) return 0;
  std::vector<int> alive;
  for (int k = 0; k < m.K; ++k) if (m.wght[k] > 0.0) alive.push_back(k);
  if (alive.empty()) return 0;
  const int J = (int)alive.size(), d = m.d;
  const size_t cs = mma_comp_stride(d), nd = (size_t)J * cs;
  RC(stage(c, 2 * nd + J + 2));
  double *h = c->hstage;
  std::vector<double> disc(J, 0.0);
  parallel_components(J, (double)d * d * d / 3.0, [&](int j) {
    const double *mu = m.mean.data() + (size_t)alive[j] * d, *Lk = m.L.data() + (size_t)alive[j] * d * d;
    mma_pack_component(d, mu, Lk, 1, h + (size_t)j * cs, &disc[j]);  // whitening: inverse factor
    mma_pack_component(d, mu, Lk, 0, h + nd + (size_t)j * cs);       // draw: the factor itself
  });
  // conditioning guard: the product with an explicit inverse and the reference's forward substitution must agree to well
  // inside the 1e-12 parity tolerance on every component, otherwise this mixture stays on the literal kernels
  double worst = 0.0;
  for (int j = 0; j < J; ++j) if (!(disc[j] <= worst)) worst = disc[j];
  o.inv_discrepancy = worst;
  if (!(worst <= c->mma_inv_tol)) return 0;
  int *hi = reinterpret_cast<int *>(h + 2 * nd);
  for (int j = 0; j < J; ++j) { hi[j] = alive[j]; hi[J + j] = m.df[alive[j]]; }
  CU(o.pack.ensure(nd * sizeof(double)));
  CU(o.lpack.ensure(nd * sizeof(double)));
  CU(o.kmap.ensure(J * sizeof(int)));
  CU(o.dfj.ensure(J * sizeof(int)));
  CU(cudaMemcpyAsync(o.pack.p, h, nd * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(o.lpack.p, h + nd, nd * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(o.kmap.p, hi, J * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(o.dfj.p, hi + J, J * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  o.alive = alive;
  o.J = J;
  o.valid = true;
  return 0;
}

int install_proposal(pmcgpu_ctx *c) {
  derive_mix(c->prop);
  int rc = upload_mix(c, c->prop, c->dprop);
  if (rc) return rc;
  c->mprop.valid = false;  // packed for the tensor path on first use (ensure_mma): shapes with a fused kernel never need it
  c->mprop.dirty = true;
  c->q_resident = false;
  c->i3_dirty = true;
  return pack_small(c, c->prop, c->dprop_pack, c->prop_pack_doubles, c->prop_pack_host, c->pin_ppack);
}

TargetDev target_view(const pmcgpu_ctx *c) {
  TargetDev t{};
  t.kind = c->tkind;
  t.d = c->td;
  t.b = c->tb;
  t.sig1 = c->tsig1;
  t.mix = c->dtmix.view;
  t.nfull = c->nfull;
  t.is_def = c->nfull ? c->ddef.as<int>() : nullptr;
  t.def_val = c->nfull ? reinterpret_cast<const double *>(c->ddef.as<char>() + ((c->nfull * sizeof(int) + 15) / 16) * 16) : nullptr;
  return t;
}

// the combined target (PMCGPU_TGT_COMBINE) of n samples at X (device) -> out (device), in chunks that bound the
// per-thread scratch of the mixture terms
int eval_combine(pmcgpu_ctx *c, long n, const double *X, double *out) {
  const long chunk = n < (1L << 20) ? n : (1L << 20);
  CU(c->ld.ensure((size_t)chunk * c->comb_kmax * sizeof(double) + 16));
  for (long base = 0; base < n; base += chunk) {
    const long m = (n - base) < chunk ? (n - base) : chunk;
    launch_combine_target(c->dcomb.as<CombTermDev>(), (int)c->comb.size(), c->td, m, X + (size_t)base * c->td, out + base,
                          c->ld.as<double>(), c->comb_kmax, c->stream);
  }
  CU(cudaGetLastError());
  return 0;
}

FusedPlan plan_for(const pmcgpu_ctx *c) {
  FusedPlan none{};
  if (c->force_generic || c->fused_version == 3 || c->prop.K == 0 || c->prop.student || c->prop.d > FUSED_MAXD || c->nfull != 0) return none;
  if (c->prop_pack_doubles == 0) return none;
  int Kt = 0;
  if (c->tkind == PMCGPU_TGT_MVDENS || c->tkind == PMCGPU_TGT_MIX) {
    if (c->tmix.student || c->t_pack_doubles == 0 || c->tmix.d != c->prop.d) return none;
    Kt = c->tmix.K;
  } else if (c->tkind == PMCGPU_TGT_BANANA) {
    if (c->td != c->prop.d || c->td < 2) return none;
  } else if (c->tkind != PMCGPU_TGT_HOST) {
    return none;
  }
  if (c->nfaces > 0 && !c->box_slab) return none;
  if (c->fused_version != 1) {
    const FusedPlan p2 = fused2_plan(c->prop.d, c->prop.K, Kt, c->tkind, c->sm_count, c->spl);
    if (p2.valid || c->fused_version == 2) return p2;
  }
  return fused_plan(c->prop.d, c->prop.K, Kt, c->sm_count, c->spl);
}

// ---- collectives on small host buffers ------------------------------------------------------------------------
int allreduce_host(pmcgpu_ctx *c, double *buf, long n, int op /*0 sum, 1 max*/) {
  if (c->nranks <= 1) return 0;
  if (c->cb) {
    if (c->cb(c->cb_user, buf, n, op) != 0) return fail(c, PMCGPU_ERR_NCCL, "allreduce callback failed");
    return 0;
  }
  if (!c->nccl_comm) return fail(c, PMCGPU_ERR_NCCL, "nranks=%d but no communicator attached", c->nranks);
  CU(c->dcoll.ensure((size_t)n * sizeof(double)));
  CU(cudaMemcpyAsync(c->dcoll.p, buf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const int r = g_nccl.AllReduce(c->dcoll.p, c->dcoll.p, (size_t)n, kNcclFloat64, op == 0 ? kNcclSum : kNcclMax, c->nccl_comm, c->stream);
  if (r != 0) return fail(c, PMCGPU_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  CU(cudaMemcpyAsync(buf, c->dcoll.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// sum over ranks of a device buffer, in place on the context's stream (NCCL ranks only; returns 1 when the ranks are
// connected through a host callback instead, then the caller reduces on the host)
int allreduce_device_sum(pmcgpu_ctx *c, double *dbuf, size_t n) {
  if (c->nranks <= 1) return 0;
  if (c->cb || !c->nccl_comm) return 1;
  const int r = g_nccl.AllReduce(dbuf, dbuf, n, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream);
  if (r != 0) return fail(c, PMCGPU_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  return 0;
}

// ---- host sink: device->host copies on a second stream, overlapped with the kernels that follow ------------------------
enum { SINK_DRAWN = 1, SINK_RHO = 2, SINK_FINAL = 4 };
void sink_join(pmcgpu_ctx *c) {
  if (c->widen.joinable()) c->widen.join();
}

int sink_copy(pmcgpu_ctx *c, int what) {
  if (!c->sink_active) return 0;
  if (!c->copy_stream) {
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_idx8, cudaEventDisableTiming));
  }
  const size_t N = c->N;
  const pmcgpu_ctx::Sink &k = c->sink;
  bool idx8 = (what & SINK_DRAWN) && k.idx && c->sink_idx8 && c->prop.K <= 256 && N > 0;
  if (idx8) {  // the byte copy of the indices, made on the compute stream before the copy stream is released
    sink_join(c);
    if (c->hidx8_cap < N) {
      if (c->hidx8) cudaFreeHost(c->hidx8);
      c->hidx8 = nullptr;
      c->hidx8_cap = 0;
      if (cudaHostAlloc((void **)&c->hidx8, N, cudaHostAllocDefault) == cudaSuccess) c->hidx8_cap = N;
      else { cudaGetLastError(); idx8 = false; }  // no pinned staging: the indices travel as size_t
    }
    if (idx8 && c->didx8.ensure(N) != cudaSuccess) { cudaGetLastError(); idx8 = false; }
    if (idx8) launch_narrow_idx((long)N, c->Idx.as<unsigned long long>(), c->didx8.as<unsigned char>(), c->sm_count, c->stream);
  }
  CU(cudaEventRecord(c->ev_copy, c->stream));
  CU(cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
  if (idx8) {
    CU(cudaMemcpyAsync(c->hidx8, c->didx8.p, N, cudaMemcpyDeviceToHost, c->copy_stream));
    CU(cudaEventRecord(c->ev_idx8, c->copy_stream));
    const int dev = c->device;
    cudaEvent_t ev = c->ev_idx8;
    const unsigned char *src = c->hidx8;
    size_t *dst = k.idx;
    c->widen = std::thread([dev, ev, src, dst, N]() {
      cudaSetDevice(dev);
      if (cudaEventSynchronize(ev) != cudaSuccess) return;
      for (size_t i = 0; i < N; ++i) dst[i] = src[i];
    });
  }
  if (idx8) c->sink_bytes += (double)N;
  if ((what & SINK_DRAWN) && k.X) c->sink_bytes += (double)N * c->d * sizeof(double);
  if ((what & SINK_DRAWN) && k.idx && !idx8) c->sink_bytes += (double)N * sizeof(size_t);
  if ((what & SINK_RHO) && k.lr) c->sink_bytes += (double)N * sizeof(double);
  if ((what & SINK_FINAL) && k.w) c->sink_bytes += (double)N * sizeof(double);
  if ((what & SINK_FINAL) && k.flg) c->sink_bytes += (double)N * sizeof(short);
  if ((what & SINK_DRAWN) && k.X) CU(cudaMemcpyAsync(k.X, c->X.p, N * c->d * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  if ((what & SINK_DRAWN) && k.idx && !idx8) CU(cudaMemcpyAsync(k.idx, c->Idx.p, N * sizeof(size_t), cudaMemcpyDeviceToHost, c->copy_stream));
  if ((what & SINK_RHO) && k.lr) CU(cudaMemcpyAsync(k.lr, c->R.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  if ((what & SINK_FINAL) && k.w) CU(cudaMemcpyAsync(k.w, c->W.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  if ((what & SINK_FINAL) && k.flg) CU(cudaMemcpyAsync(k.flg, c->Flg.p, N * sizeof(short), cudaMemcpyDeviceToHost, c->copy_stream));
  return 0;
}

// the reference's running maximum with its unconditional i==0 assignment (pmc.c:535,552,579)
double fold_max(double m_all, bool has_first, bool reached0, double v0) {
  if (has_first && reached0) return std::isnan(v0) ? v0 : (m_all > v0 ? m_all : v0);
  return m_all > kNegInit ? m_all : kNegInit;
}

int zero_small(pmcgpu_ctx *c) {
  CU(c->small.ensure(SM_LEN * 8));
  CU(cudaMemsetAsync(c->small.p, 0, SM_LEN * 8, c->stream));
  return 0;
}
int read_small(pmcgpu_ctx *c, unsigned long long *counters, unsigned long long *count_k, double *first) {
  RC(stage(c, SM_LEN));
  CU(cudaMemcpyAsync(c->hstage, c->small.p, SM_LEN * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const unsigned long long *u = reinterpret_cast<const unsigned long long *>(c->hstage);
  if (counters) memcpy(counters, u + SM_COUNTERS, CNT_NUM * 8);
  if (count_k) memcpy(count_k, u + SM_COUNTK, 64 * 8);
  if (first) memcpy(first, c->hstage + SM_FIRST, 4 * 8);
  return 0;
}

struct WeightOut {
  double M_all = -INFINITY, R_all = -INFINITY, S = 0.0;  // NaN-ignoring maxima over local samples; S = sum exp(lw - M_all)
  double first[4] = {0, 0, 0, 0};
  bool has_first = false;
  unsigned long long counters[CNT_NUM] = {0};
  unsigned long long count_k[64] = {0};
  std::vector<double> stats;  // canonical K*E statistics relative to M_all (fused path with do_stats)
};

int ensure_cwght_normalised(pmcgpu_ctx *c) {
  // mix_mvdens_ran renormalises the weights in place before building cwght (mvdens.c:736-751)
  HostMix &m = c->prop;
  double s = 0.0;
  for (int k = 0; k < m.K; ++k) s += m.wght[k];
  if (std::fabs(s - 1.0) > 1e-6) {
    if (s == 0) return fail(c, kErrMvNegWeight, "All weights to zero !! cannot sample");
    const double inv = 1.0 / s;
    for (int k = 0; k < m.K; ++k) m.wght[k] *= inv;
    return install_proposal(c);
  }
  return 0;
}

int check_ready(pmcgpu_ctx *c, bool need_target) {
  if (!c) return PMCGPU_ERR_USAGE;
  if (c->prop.K == 0) return fail(c, PMCGPU_ERR_USAGE, "no proposal set");
  if (c->N <= 0 || c->d != c->prop.d) return fail(c, PMCGPU_ERR_USAGE, "sample store not sized for the proposal (N=%ld d=%d, proposal d=%d)", c->N, c->d, c->prop.d);
  if (need_target && c->tkind == PMCGPU_TGT_NONE) return fail(c, PMCGPU_ERR_USAGE, "no target set");
  if (c->prop.n_alive() == 0) return fail(c, kErrMvOutOfBound, "Indice out of bonds: no component with positive weight");
  return 0;
}

// Fused kernel over the resident store.  do_draw / do_stats select the phases.
int run_fused(pmcgpu_ctx *c, const FusedPlan &p, int do_draw, int do_stats, uint64_t seed, uint32_t iter,
              long first_index, double t_temper, const double *post_in, const short *ok_in, WeightOut &o) {
  FusedArgs a{};
  a.mixpack = c->dprop_pack.as<double>();
  a.mix_doubles = c->prop_pack_doubles;
  a.comp_stride = fused_comp_stride(c->prop.d);
  a.K = c->prop.K;
  a.tgt_kind = c->tkind;
  if (c->tkind == PMCGPU_TGT_MVDENS || c->tkind == PMCGPU_TGT_MIX) {
    a.tgtpack = c->dt_pack.as<double>();
    a.tgt_doubles = c->t_pack_doubles;
    a.tgt_stride = fused_comp_stride(c->prop.d);
    a.Kt = c->tmix.K;
  }
  a.b = c->tb;
  a.sig1 = c->tsig1;
  a.t_temper = t_temper;
  a.post_in = post_in;
  a.ok_in = ok_in;
  a.has_box = c->nfaces > 0;
  for (int j = 0; j < FUSED_MAXD; ++j) { a.lo_thr[j] = c->lo_thr[j]; a.hi_thr[j] = c->hi_thr[j]; a.pivot[j] = 0.0; }
  {  // pivot = mixture mean (weights of live components)
    const HostMix &m = c->prop;
    double sw = 0.0;
    for (int k = 0; k < m.K; ++k) if (m.wght[k] > 0) sw += m.wght[k];
    for (int j = 0; j < m.d; ++j) {
      double s = 0.0;
      for (int k = 0; k < m.K; ++k) if (m.wght[k] > 0) s += m.wght[k] * m.mean[(size_t)k * m.d + j];
      a.pivot[j] = s / sw;
    }
  }
  a.zig = c->zig.as<double2>();
  a.do_draw = do_draw;
  a.do_stats = do_stats;
  a.N = c->N;
  a.first_index = first_index;
  a.seed = seed;
  a.iter = iter;
  a.max_attempts = c->max_attempts;
  a.X = c->X.as<double>();
  a.lw = c->W.as<double>();
  a.log_rho = c->R.as<double>();
  a.indices = c->Idx.as<unsigned long long>();
  a.flg = c->Flg.as<short>();
  const int maxblk = (p.valid == 2) ? fused2_max_blocks(p) : p.grid;
  CU(c->partials.ensure((size_t)maxblk * p.partial_len * sizeof(double)));
  CU(c->stats.ensure((size_t)(p.nstat + 4) * sizeof(double)));
  a.partials = c->partials.as<double>();
  a.partial_len = p.partial_len;
  int rc = zero_small(c);
  if (rc) return rc;
  unsigned long long *sm = c->small.as<unsigned long long>();
  a.counters = sm + SM_COUNTERS;
  a.count_k = sm + SM_COUNTK;
  a.first_vals = reinterpret_cast<double *>(sm + SM_FIRST);
  const int mode = (p.valid == 2) ? (do_stats ? ((c->split_stats || p.SPL > 1 || p.NR == 0) ? 1 : 2) : 0) : -1;
  if (mode == 1) {
    CU(c->resp.ensure((size_t)c->prop.K * c->N * sizeof(double)));
    CU(c->partials2.ensure((size_t)p.grid * 2 * p.partial_len * sizeof(double)));
    CU(c->dscal.ensure(4 * sizeof(double)));
    a.resp = c->resp.as<double>();
  }
  if (c->timing) {
    if (!c->ev0) { CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); CU(cudaEventCreate(&c->ev2)); CU(cudaEventCreate(&c->evd)); }
    CU(cudaEventRecord(c->ev0, c->stream));
  }
  int nblk = p.grid;
  int lrc = 0;
  bool drew_separately = false;
  if (p.valid == 2 && c->split_draw && do_draw && mode != 2) {  // draw as its own (integer-heavy) launch, then the FP64 kernel on X
    FusedArgs ad = a;
    lrc = launch_fused2(p, ad, c->prop_pack_host.data(), c->tgt_pack_host.empty() ? nullptr : c->tgt_pack_host.data(), 3, &nblk, c->stream);
    a.do_draw = 0;
    drew_separately = true;
    if (!lrc) RC(sink_copy(c, SINK_DRAWN));  // X and indices are final: they travel while the FP64 kernels run
  }
  if (c->timing) CU(cudaEventRecord(c->evd, c->stream));
  if (!lrc)
    lrc = (p.valid == 2) ? launch_fused2(p, a, c->prop_pack_host.data(), c->tgt_pack_host.empty() ? nullptr : c->tgt_pack_host.data(), mode, &nblk, c->stream)
                         : launch_fused(p, a, c->stream);
  if (lrc != 0) return fail(c, PMCGPU_ERR_CUDA, "fused kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  RC(sink_copy(c, (do_draw && !drew_separately ? SINK_DRAWN : 0) | SINK_RHO));
  if (c->timing) CU(cudaEventRecord(c->ev1, c->stream));
  int crc = 0;
  if (mode == 1) {
    crc = launch_fused2_max(p, a.partials, nblk, c->dscal.as<double>(), c->stream);
    if (!crc) crc = launch_fused2_stats(p, a.X, a.lw, a.resp, a.flg, c->dscal.as<double>(), a.pivot, c->N, c->partials2.as<double>(), c->stream);
    if (c->timing) CU(cudaEventRecord(c->ev2, c->stream));
    if (!crc) crc = launch_fused2_combine(p, c->partials2.as<double>(), p.grid * 2, c->stats.as<double>(), c->stream);
  } else if (p.valid == 2) {
    crc = launch_fused2_combine(p, a.partials, nblk, c->stats.as<double>(), c->stream);
  } else {
    crc = launch_fused_combine(p, a.partials, c->prop.K, c->stats.as<double>(), c->stream);
  }
  if (crc != 0)
    return fail(c, PMCGPU_ERR_CUDA, "combine launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  RC(stage(c, (size_t)p.nstat + 8 + SM_LEN));
  CU(cudaMemcpyAsync(c->hstage, c->stats.p, (size_t)(p.nstat + 4) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->hstage + p.nstat + 4, c->small.p, SM_LEN * 8, cudaMemcpyDeviceToHost, c->stream));
  if (mode == 1) CU(cudaMemcpyAsync(c->hstage + p.nstat + 4 + SM_LEN, c->dscal.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->timing) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->evd, c->ev1));  // the FP64 kernel (or the single fused kernel)
    c->fused_ms += ms;
    c->fused_launches += 1;
    if (drew_separately) { CU(cudaEventElapsedTime(&ms, c->ev0, c->evd)); c->draw_ms += ms; }
    if (mode == 1) { CU(cudaEventElapsedTime(&ms, c->ev1, c->ev2)); c->stats_ms += ms; }
  }
  o.stats.assign(c->hstage, c->hstage + p.nstat);
  o.M_all = c->hstage[p.nstat];
  o.S = c->hstage[p.nstat + 1];
  o.R_all = (mode == 1) ? c->hstage[p.nstat + 4 + SM_LEN + 1] : c->hstage[p.nstat + 2];
  const unsigned long long *u = reinterpret_cast<const unsigned long long *>(c->hstage + p.nstat + 4);
  memcpy(o.counters, u + SM_COUNTERS, CNT_NUM * 8);
  memcpy(o.count_k, u + SM_COUNTK, 64 * 8);
  memcpy(o.first, c->hstage + p.nstat + 4 + SM_FIRST, 4 * 8);
  o.has_first = (first_index == 0);
  return 0;
}


// ---- third-generation small-d iteration (pmc_iter3.cu) -------------------------------------------------------------
void mixture_pivot(const HostMix &m, double *pivot) {  // mixture mean over the live components
  double sw = 0.0;
  for (int k = 0; k < m.K; ++k) if (m.wght[k] > 0) sw += m.wght[k];
  for (int j = 0; j < m.d; ++j) {
    double s = 0.0;
    for (int k = 0; k < m.K; ++k) if (m.wght[k] > 0) s += m.wght[k] * m.mean[(size_t)k * m.d + j];
    pivot[j] = s / sw;
  }
}

// packs of the current proposal / target for the third-generation kernels; the instantiation may have more components
// than the mixture (K rounded up): the extra ones are packed as dead components
int pack_iter3(pmcgpu_ctx *c, const Iter3Plan &p) {
  if (!c->i3_dirty) return 0;
  const HostMix &m = c->prop;
  const int D = m.d, ws = iter3_wstride(D), ds = iter3_dstride(D);
  for (int j = 0; j < FUSED_MAXD; ++j) c->i3_shift[j] = 0.0;
  mixture_pivot(m, c->i3_shift);
  std::vector<double> wg(p.K, 0.0), mu((size_t)p.K * D, 0.0), L((size_t)p.K * D * D, 0.0), ld(p.K, 0.0);
  memcpy(wg.data(), m.wght.data(), m.K * sizeof(double));
  memcpy(mu.data(), m.mean.data(), (size_t)m.K * D * sizeof(double));
  memcpy(L.data(), m.L.data(), (size_t)m.K * D * D * sizeof(double));
  memcpy(ld.data(), m.logdet.data(), m.K * sizeof(double));
  c->i3_hw.assign((size_t)p.K * ws, 0.0);
  c->i3_hd.assign((size_t)p.K * ds, 0.0);
  // the whitening subtracts each component's mean exactly (no common shift): i3_shift is the pivot of the statistics only
  const double noshift[FUSED_MAXD] = {0};
  // w3_variant 3: the row-scaled solve about the common shift, trusted while no live component has |mu'_j / L_jj| above
  // gen3_max_ratio (its offsets m_j = -mu'_j / L_jj carry a relative rounding error); otherwise the exact subtraction
  c->i3_pre_ratio = iter3_shift_ratio(p.K, D, wg.data(), mu.data(), L.data(), c->i3_shift);
  c->i3_pre = c->w3_variant == 3 && c->i3_pre_ratio <= c->gen3_max_ratio;
  iter3_pack_w(p.K, D, wg.data(), mu.data(), L.data(), ld.data(), c->i3_pre ? c->i3_shift : noshift, c->i3_hw.data(), c->i3_pre ? 1 : 0);
  iter3_pack_d(p.K, D, mu.data(), L.data(), c->i3_hd.data());
  c->i3_ratio = 0.0;
  // ranindx (mvdens.c:776): the index is the number of the first K-1 cumulative weights below z
  c->i3_hcw.assign(64, INFINITY);
  for (int k = 0; k + 1 < m.K; ++k) c->i3_hcw[k] = m.cw[k];
  c->i3_ht.clear();
  if (p.KT > 0 && (c->tkind == PMCGPU_TGT_MVDENS || c->tkind == PMCGPU_TGT_MIX)) {
    const HostMix &t = c->tmix;
    std::vector<double> twg(p.KT, 0.0), tmu((size_t)p.KT * D, 0.0), tL((size_t)p.KT * D * D, 0.0), tld(p.KT, 0.0);
    memcpy(twg.data(), t.wght.data(), t.K * sizeof(double));
    memcpy(tmu.data(), t.mean.data(), (size_t)t.K * D * sizeof(double));
    memcpy(tL.data(), t.L.data(), (size_t)t.K * D * D * sizeof(double));
    memcpy(tld.data(), t.logdet.data(), t.K * sizeof(double));
    c->i3_ht.assign((size_t)p.KT * ws, 0.0);
    iter3_pack_w(p.KT, D, twg.data(), tmu.data(), tL.data(), tld.data(), noshift, c->i3_ht.data());
  }
  CU(c->i3_gpack.ensure(c->i3_hd.size() * sizeof(double)));
  CU(cudaMemcpyAsync(c->i3_gpack.p, c->i3_hd.data(), c->i3_hd.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->i3_dirty = false;
  return 0;
}

Iter3Plan plan3_for(pmcgpu_ctx *c) {
  Iter3Plan none{};
  if (!c->gen3 || c->fused_version == 1 || c->fused_version == 2 || c->force_generic || c->prop.K == 0 || c->prop.student || c->prop.d > FUSED_MAXD || c->nfull != 0)
    return none;
  int Kt = 0;
  if (c->tkind == PMCGPU_TGT_MVDENS || c->tkind == PMCGPU_TGT_MIX) {
    if (c->tmix.student || c->tmix.d != c->prop.d || c->tmix.K < 1) return none;
    Kt = c->tmix.K;
  } else if (c->tkind == PMCGPU_TGT_BANANA) {
    if (c->td != c->prop.d || c->td < 2) return none;
  } else if (c->tkind == PMCGPU_TGT_COMBINE) {
    if (c->td != c->prop.d) return none;  // evaluated by its own kernel, handed to the weights kernel like host values
  } else if (c->tkind != PMCGPU_TGT_HOST) {
    return none;
  }
  if (c->nfaces > 0 && !c->box_slab) return none;
  Iter3Plan p = iter3_plan(c->prop.d, c->prop.K, Kt, c->tkind, c->sm_count);
  if (!p.valid) return none;
  if (pack_iter3(c, p) != 0) return none;
  if (!(c->i3_ratio <= c->gen3_max_ratio)) return none;  // far-off narrow components: the second generation subtracts the means exactly
  return p;
}

struct Iter3Out {
  double scal[I3_SCAL_LEN] = {0};
  std::vector<double> out;  // stats[K3*E] (with_stats) | H, Q, ok, flagged-out | count_k[K3]
  bool clean = false;
};

// reduction over ranks of n doubles of a device buffer, op 0 = sum, 1 = max: NCCL in place on the stream, or the host
// callback through the staging buffer
int allreduce_dev(pmcgpu_ctx *c, double *dbuf, long n, int op) {
  if (c->nranks <= 1) return 0;
  if (!c->cb) {
    if (!c->nccl_comm) return fail(c, PMCGPU_ERR_NCCL, "nranks=%d but no communicator attached", c->nranks);
    const int r = g_nccl.AllReduce(dbuf, dbuf, (size_t)n, kNcclFloat64, op == 0 ? kNcclSum : kNcclMax, c->nccl_comm, c->stream);
    if (r != 0) return fail(c, PMCGPU_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return 0;
  }
  RC(stage(c, (size_t)n));
  CU(cudaMemcpyAsync(c->hstage, dbuf, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->cb(c->cb_user, c->hstage, n, op) != 0) return fail(c, PMCGPU_ERR_NCCL, "allreduce callback failed");
  CU(cudaMemcpyAsync(dbuf, c->hstage, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// draw (optional) -> weights -> maxima/normaliser -> statistics + normalisation -> one device-to-host copy.
// stage_limit: 1 = stop after the weights and their reductions (log-weights stay un-normalised), 2 = full.
// The per-lane-component draw kernel (k_fused2 MODE 3: Philox + ziggurat + mu_k + L_k g + slab box) wherever it is
// instantiated for the proposal's exact K: measured 4.0 ms per 1e8 samples against 8.7 ms of the grouped k_draw3 and 33 ms of
// the literal kernel (profiles/r02_*).  c->small must have been zeroed (zero_small).  *done = 0: no instantiation, nothing launched.
int draw_fused2(pmcgpu_ctx *c, uint64_t seed, uint32_t iter, long first_index, int *done) {
  *done = 0;
  if (!c->prop_pack_doubles || c->prop.student || c->force_generic || c->nfull != 0 || (c->nfaces > 0 && !c->box_slab)) return 0;
  const FusedPlan p2 = fused2_plan_draw(c->prop.d, c->prop.K, c->sm_count);
  if (!p2.valid) return 0;
  const double *hpack = c->prop_pack_host.data();
  if (p2.RK > c->prop.K) {  // a larger instantiation: the extra components are zeros and never picked
    c->draw_pack_padded.assign((size_t)p2.RK * fused_comp_stride(c->prop.d), 0.0);
    memcpy(c->draw_pack_padded.data(), c->prop_pack_host.data(), c->prop_pack_host.size() * sizeof(double));
    hpack = c->draw_pack_padded.data();
  }
  unsigned long long *sm = c->small.as<unsigned long long>();
  FusedArgs a{};
  a.K = c->prop.K;
  a.has_box = c->nfaces > 0;
  for (int j = 0; j < FUSED_MAXD; ++j) { a.lo_thr[j] = c->lo_thr[j]; a.hi_thr[j] = c->hi_thr[j]; a.pivot[j] = 0.0; }
  a.zig = c->zig.as<double2>();
  a.do_draw = 1;
  a.N = c->N; a.first_index = first_index; a.seed = seed; a.iter = iter; a.max_attempts = c->max_attempts;
  a.X = c->X.as<double>(); a.lw = c->W.as<double>(); a.log_rho = c->R.as<double>();
  a.indices = c->Idx.as<unsigned long long>(); a.flg = c->Flg.as<short>();
  CU(c->partials.ensure((size_t)fused2_max_blocks(p2) * p2.partial_len * sizeof(double)));
  a.partials = c->partials.as<double>(); a.partial_len = p2.partial_len;
  a.counters = sm + SM_COUNTERS; a.count_k = sm + SM_COUNTK; a.first_vals = reinterpret_cast<double *>(sm + SM_FIRST);
  int nb2 = 0;
  if (launch_fused2(p2, a, hpack, nullptr, 3, &nb2, c->stream))
    return fail(c, PMCGPU_ERR_CUDA, "draw kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  CU(cudaMemsetAsync(sm + SM_COUNTK, 0, 64 * 8, c->stream));
  CU(cudaMemsetAsync(sm + SM_COUNTERS + CNT_NOK, 0, 8, c->stream));
  *done = 1;
  return 0;
}

int run_iter3(pmcgpu_ctx *c, const Iter3Plan &p, int do_draw, int stage_limit, int with_stats, uint64_t seed, uint32_t iter,
              long first_index, long N_total, double t_temper, const double *post_in, const short *ok_in, Iter3Out &o) {
  const HostMix &m = c->prop;
  const int D = m.d;
  const long N = c->N;
  RC(zero_small(c));
  unsigned long long *sm = c->small.as<unsigned long long>();
  CU(c->i3_wpart.ensure((size_t)p.weights_blocks * 4 * sizeof(double)));
  CU(c->i3_scal.ensure(I3_SCAL_LEN * sizeof(double)));
  CU(c->i3_coll.ensure(I3_COLL_LEN * sizeof(double)));
  const int nout = (with_stats ? p.K * p.E : 0) + 4 + p.K;
  CU(c->i3_out.ensure((size_t)(p.K * p.E + 4 + p.K) * sizeof(double)));
  if (stage_limit >= 2) {
    CU(c->i3_spart.ensure((size_t)p.stats_blocks * p.partial_len * sizeof(double)));
    if (with_stats) CU(c->resp.ensure((size_t)p.K * N * sizeof(double)));
  }
  if (iter3_upload(p, c->i3_hw.data(), (int)c->i3_hw.size(), c->i3_ht.empty() ? nullptr : c->i3_ht.data(), (int)c->i3_ht.size(),
                   c->i3_hd.data(), (int)c->i3_hd.size(), c->i3_hcw.data(), 64, c->stream))
    return fail(c, PMCGPU_ERR_CUDA, "constant upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (c->timing) {
    if (!c->ev0) { CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); CU(cudaEventCreate(&c->ev2)); CU(cudaEventCreate(&c->evd)); }
    CU(cudaEventRecord(c->ev0, c->stream));
  }
  int drew2 = 0;
  if (do_draw && !c->draw3) RC(draw_fused2(c, seed, iter, first_index, &drew2));
  if (drew2) {
    RC(sink_copy(c, SINK_DRAWN));
  } else if (do_draw) {
    Iter3Draw d{};
    d.N = N; d.first_index = first_index; d.seed = seed; d.iter = iter; d.max_attempts = c->max_attempts;
    d.has_box = c->nfaces > 0;
    for (int j = 0; j < FUSED_MAXD; ++j) { d.lo_thr[j] = c->lo_thr[j]; d.hi_thr[j] = c->hi_thr[j]; }
    d.X = c->X.as<double>(); d.indices = c->Idx.as<unsigned long long>(); d.flg = c->Flg.as<short>();
    d.zig = c->zig.as<double2>(); d.gpack = c->i3_gpack.as<double>(); d.counters = sm + SM_COUNTERS;
    if (launch_draw3(p, d, c->stream)) return fail(c, PMCGPU_ERR_CUDA, "draw kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    RC(sink_copy(c, SINK_DRAWN));
  }
  if (c->timing) CU(cudaEventRecord(c->evd, c->stream));
  int wblocks = 0;
  {
    Iter3Weights w{};
    w.X = c->X.as<double>(); w.N = N; w.first_index = first_index; w.tgt_kind = c->tkind; w.b = c->tb; w.sig1 = c->tsig1;
    w.t_temper = t_temper; w.post_in = post_in; w.ok_in = ok_in;
    if (c->tkind == PMCGPU_TGT_COMBINE && !post_in) {  // the sum of the terms, computed by k_combine_target
      CU(c->hostpost.ensure((size_t)N * sizeof(double)));
      RC(eval_combine(c, N, c->X.as<double>(), c->hostpost.as<double>()));
      w.post_in = c->hostpost.as<double>();
      w.ok_in = nullptr;
    }
    if (c->tkind == PMCGPU_TGT_COMBINE) w.tgt_kind = PMCGPU_TGT_HOST;
    for (int j = 0; j < FUSED_MAXD; ++j) w.shift[j] = c->i3_shift[j];
    w.lw = c->W.as<double>(); w.log_rho = c->R.as<double>(); w.flg = c->Flg.as<short>();
    w.resp = (stage_limit >= 2 && with_stats) ? c->resp.as<double>() : nullptr;
    w.partials = c->i3_wpart.as<double>(); w.counters = sm + SM_COUNTERS; w.first_vals = reinterpret_cast<double *>(sm + SM_FIRST);
    w.mix = c->dprop.view; w.tmix = c->dtmix.view; w.variant = c->i3_pre ? 3 : (c->w3_variant == 3 ? 0 : c->w3_variant);
    if (launch_weights3(p, w, &wblocks, c->stream)) return fail(c, PMCGPU_ERR_CUDA, "weights kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (c->timing) CU(cudaEventRecord(c->ev1, c->stream));
  RC(sink_copy(c, SINK_RHO));
  const double *fv = reinterpret_cast<const double *>(sm + SM_FIRST);
  double *coll = c->i3_coll.as<double>(), *scal = c->i3_scal.as<double>();
  const int has_first = first_index == 0;
  if (c->nranks <= 1) {
    if (launch_red3(c->i3_wpart.as<double>(), wblocks, sm + SM_COUNTERS, fv, has_first, coll, scal, 7, (double)N_total, c->stream))
      return fail(c, PMCGPU_ERR_CUDA, "reduction launch failed");
  } else {
    // one collective for the maxima and the normaliser: every rank publishes its 16 local values in its slot of a zeroed
    // [nranks][16] buffer, the sum all-reduce gathers the slots, k_red3 folds them in rank order (see k_red3)
    CU(c->i3_gath.ensure((size_t)c->nranks * 16 * sizeof(double)));
    double *gath = c->i3_gath.as<double>();
    if (launch_red3(c->i3_wpart.as<double>(), wblocks, sm + SM_COUNTERS, fv, has_first, coll, scal, 1 | 2 | 16, (double)N_total, c->stream, gath, c->rank, c->nranks))
      return fail(c, PMCGPU_ERR_CUDA, "reduction launch failed");
    RC(allreduce_dev(c, gath, (long)c->nranks * 16, 0));
    if (launch_red3(c->i3_wpart.as<double>(), wblocks, sm + SM_COUNTERS, fv, has_first, coll, scal, 8 | 4, (double)N_total, c->stream, gath, c->rank, c->nranks))
      return fail(c, PMCGPU_ERR_CUDA, "reduction launch failed");
  }
  if (stage_limit >= 2) {
    Iter3Stats st{};
    st.X = c->X.as<double>(); st.resp = c->resp.as<double>(); st.W = c->W.as<double>(); st.flg = c->Flg.as<short>();
    st.indices = c->Idx.as<unsigned long long>(); st.scal = scal; st.N = N; st.partials = c->i3_spart.as<double>();
    st.with_stats = with_stats;
    for (int j = 0; j < FUSED_MAXD; ++j) st.pivot[j] = c->i3_shift[j];
    if (launch_stats3(p, st, c->stream)) return fail(c, PMCGPU_ERR_CUDA, "statistics kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (c->timing) CU(cudaEventRecord(c->ev2, c->stream));
    RC(sink_copy(c, SINK_FINAL));
    if (launch_comb3(p, c->i3_spart.as<double>(), with_stats, scal, c->i3_out.as<double>(), c->stream))
      return fail(c, PMCGPU_ERR_CUDA, "combine launch failed");
    RC(allreduce_dev(c, c->i3_out.as<double>(), nout, 0));
  }
  RC(stage(c, (size_t)I3_SCAL_LEN + nout));
  CU(cudaMemcpyAsync(c->hstage, scal, I3_SCAL_LEN * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (stage_limit >= 2)
    CU(cudaMemcpyAsync(c->hstage + I3_SCAL_LEN, c->i3_out.p, (size_t)nout * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(o.scal, c->hstage, sizeof o.scal);
  o.clean = o.scal[I3_CLEAN] != 0.0;
  if (stage_limit >= 2) o.out.assign(c->hstage + I3_SCAL_LEN, c->hstage + I3_SCAL_LEN + nout);
  if (c->timing) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->evd, c->ev1));
    c->fused_ms += ms;
    c->fused_launches += 1;
    if (do_draw) { CU(cudaEventElapsedTime(&ms, c->ev0, c->evd)); c->draw_ms += ms; }
    if (stage_limit >= 2) { CU(cudaEventElapsedTime(&ms, c->ev1, c->ev2)); c->stats_ms += ms; }
  }
  (void)D;
  return 0;
}

constexpr long GEN_CHUNK = 1L << 18;

// tensor-path packs are built lazily, the first time a call cannot be served by a fused small-d kernel
int ensure_mma(pmcgpu_ctx *c) {
  if (c->mprop.dirty) { c->mprop.dirty = false; RC(pack_mma(c, c->prop, c->mprop)); }
  if (c->mtgt.dirty) {
    c->mtgt.dirty = false;
    if (c->tmix.K && (c->tkind == PMCGPU_TGT_MVDENS || c->tkind == PMCGPU_TGT_MIX)) RC(pack_mma(c, c->tmix, c->mtgt));
  }
  return 0;
}

// can the large-d tensor path evaluate this configuration?  (mode 1 = target values supplied by the host)
bool mma_usable(const pmcgpu_ctx *c, int mode) {
  if (!c->use_mma || c->force_generic || !c->mprop.valid || c->nfull != 0) return false;
  if (mode == 1 || c->tkind == PMCGPU_TGT_HOST) return true;
  if (c->tkind == PMCGPU_TGT_BANANA) return c->td == c->prop.d;
  if (c->tkind == PMCGPU_TGT_MVDENS || c->tkind == PMCGPU_TGT_MIX) return c->mtgt.valid && c->tmix.d == c->prop.d;
  return false;
}

// weights through pmc_mma.cu: DMMA whitening of every (sample, component) pair, then the literal tail
int run_weights_mma(pmcgpu_ctx *c, int mode, const double *post_in, const short *ok_in, double t_temper,
                    long first_index, WeightOut &o) {
  const long N = c->N;
  const int K = c->prop.K, d = c->prop.d;
  const bool tq = mode == 0 && (c->tkind == PMCGPU_TGT_MVDENS || c->tkind == PMCGPU_TGT_MIX);
  const int Kt = tq ? c->tmix.K : 0;
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  const size_t have = c->Q.cap + c->Qt.cap;
  const size_t budget = std::min<size_t>((free_b + have) / 2, (size_t)64 << 30);
  long chunk = (long)(budget / (sizeof(double) * (size_t)(K + Kt)));
  chunk = std::max<long>(1024, (chunk / 1024) * 1024);
  if (c->mma_max_chunk > 0 && chunk > c->mma_max_chunk) chunk = c->mma_max_chunk;
  if (chunk > N) chunk = N;
  const long qs = chunk;
  CU(c->Q.ensure((size_t)K * qs * sizeof(double)));
  if (tq) CU(c->Qt.ensure((size_t)Kt * qs * sizeof(double)));
  const int nb = (int)((chunk + GEN_WEIGHT_THREADS - 1) / GEN_WEIGHT_THREADS);
  CU(c->blkmax.ensure((size_t)2 * nb * sizeof(double)));
  int rc = zero_small(c);
  if (rc) return rc;
  unsigned long long *sm = c->small.as<unsigned long long>();
  const TargetDev tv = target_view(c);
  RC(stage(c, (size_t)2 * nb + SM_LEN));
  if (c->timing) {
    if (!c->ev0) { CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); CU(cudaEventCreate(&c->ev2)); CU(cudaEventCreate(&c->evd)); }
    CU(cudaEventRecord(c->ev0, c->stream));
  }
  for (long base = 0; base < N; base += chunk) {
    const long n = (N - base) < chunk ? (N - base) : chunk;
    const int nbc = (int)((n + GEN_WEIGHT_THREADS - 1) / GEN_WEIGHT_THREADS);
    const double *Xb = c->X.as<double>() + (size_t)base * d;
    if (launch_q_mma(Xb, n, d, c->mprop.pack.as<double>(), c->mprop.J, c->mprop.kmap.as<int>(), c->Q.as<double>(), qs, c->sm_count, c->stream))
      return fail(c, PMCGPU_ERR_CUDA, "whitening kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (tq && launch_q_mma(Xb, n, d, c->mtgt.pack.as<double>(), c->mtgt.J, c->mtgt.kmap.as<int>(), c->Qt.as<double>(), qs, c->sm_count, c->stream))
      return fail(c, PMCGPU_ERR_CUDA, "whitening kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (c->timing && base == 0) CU(cudaEventRecord(c->ev1, c->stream));
    launch_weights_q(c->dprop.view, tv, mode, post_in ? post_in + base : nullptr, ok_in ? ok_in + base : nullptr, t_temper,
                     first_index + base, n, Xb, c->Q.as<double>(), c->Qt.as<double>(), qs, c->W.as<double>() + base,
                     c->R.as<double>() + base, c->Flg.as<short>() + base, c->blkmax.as<double>(), c->blkmax.as<double>() + nb,
                     sm + SM_COUNTERS, reinterpret_cast<double *>(sm + SM_FIRST), c->mma_store_ld, c->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->hstage, c->blkmax.p, (size_t)2 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int b = 0; b < nbc; ++b) {
      if (c->hstage[b] > o.M_all) o.M_all = c->hstage[b];
      if (c->hstage[nb + b] > o.R_all) o.R_all = c->hstage[nb + b];
    }
    if (c->timing && base == 0) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));  // the whitening kernel(s) of the first chunk
      c->fused_ms += ms;
      c->fused_launches += 1;
    }
  }
  c->q_stride = qs;
  c->q_resident = (chunk == N) && c->mma_store_ld != 0;
  rc = read_small(c, o.counters, nullptr, o.first);
  o.has_first = (first_index == 0);
  return rc;
}

// one-pass pivoted statistics on the tensor path (Gaussian components; Q still holds ld_k for the whole store and the
// weights are normalised).  Outputs in the layout pmcgpu_finalize_update expects.
struct StatView {  // per-component sufficient statistics with explicit strides (doubles)
  const double *W = nullptr, *G = nullptr, *s1 = nullptr, *S2 = nullptr;
  size_t sW = 1, sG = 1, ss1 = 0, sS2 = 0;
};
int stats_mma(pmcgpu_ctx *c, const std::vector<size_t> &count, const std::vector<double> &pivot, std::vector<double> &W,
              StatView &view) {
  const HostMix &m = c->prop;
  const int K = m.K, d = m.d, J = c->mprop.J;
  const long N = c->N;
  const size_t len = 1 + d + (size_t)d * d;
  CU(c->tmpo.ensure((size_t)K * 8));
  CU(c->mpivot.ensure((size_t)d * sizeof(double)));
  RC(stage(c, (size_t)K + d));
  unsigned long long *hc = reinterpret_cast<unsigned long long *>(c->hstage);
  for (int k = 0; k < K; ++k) hc[k] = count[k];
  memcpy(c->hstage + K, pivot.data(), d * sizeof(double));
  CU(cudaMemcpyAsync(c->tmpo.p, hc, (size_t)K * 8, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->mpivot.p, c->hstage + K, d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const int nchunk = mma_stats_chunks(N, J, c->sm_count);
  CU(c->mpart.ensure(mma_stats_partial_doubles(d, J, nchunk) * sizeof(double)));
  CU(c->mstats.ensure((size_t)K * len * sizeof(double)));
  if (c->xc.ensure((size_t)N * d * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return 1; }  // no room for the centred copy: literal path
  CU(cudaMemsetAsync(c->mstats.p, 0, (size_t)K * len * sizeof(double), c->stream));
  if (c->timing && c->ev1) CU(cudaEventRecord(c->ev1, c->stream));
  const bool student = m.student;
  const int nrb = resp_q_blocks(N);
  if (student) {
    if (K > 256) return 1;  // block reduction scratch: 32 x K doubles of shared memory
    CU(c->wpart.ensure((size_t)nrb * K * sizeof(double)));
  }
  launch_resp_q(c->dprop.view, N, c->W.as<double>(), c->R.as<double>(), c->Flg.as<short>(), c->tmpo.as<unsigned long long>(),
                c->Q.as<double>(), c->q_stride, student ? c->wpart.as<double>() : nullptr, c->stream);
  CU(cudaGetLastError());
  if (launch_stats_mma(c->X.as<double>(), c->Flg.as<short>(), N, d, c->Q.as<double>(), c->q_stride, J, c->mprop.kmap.as<int>(),
                       c->mpivot.as<double>(), nchunk, c->xc.as<double>(), c->mpart.as<double>(), c->mstats.as<double>(),
                       c->sm_count, c->stream))
    return fail(c, PMCGPU_ERR_CUDA, "statistics kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (c->timing && c->ev1) CU(cudaEventRecord(c->ev2, c->stream));
  // K (1 + d + d^2) statistics (8.5 MB at d = 128, K = 64): reduced over NVLink before they cross PCIe once
  const int dev_reduced = allreduce_device_sum(c, c->mstats.as<double>(), (size_t)K * len);
  if (dev_reduced < 0) return dev_reduced;
  c->q_resident = false;  // Q now holds responsibilities
  RC(stage(c, (size_t)K * len + (student ? (size_t)nrb * K : 0)));
  CU(cudaMemcpyAsync(c->hstage, c->mstats.p, (size_t)K * len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (student)
    CU(cudaMemcpyAsync(c->hstage + (size_t)K * len, c->wpart.p, (size_t)nrb * K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<double> wsum;
  if (student) {  // weight_d = sum_i w8 (pmc.c:1392), block partials added in block order
    wsum.assign(K, 0.0);
    const double *wp = c->hstage + (size_t)K * len;
    for (int b = 0; b < nrb; ++b)
      for (int k = 0; k < K; ++k) wsum[k] += wp[(size_t)b * K + k];
    int rcw = allreduce_host(c, wsum.data(), K, 0);
    if (rcw) return rcw;
  }
  if (c->timing && c->ev1) { float ms = 0.f; CU(cudaEventElapsedTime(&ms, c->ev1, c->ev2)); c->stats_ms += ms; }
  int rc = dev_reduced == 1 ? allreduce_host(c, c->hstage, (long)((size_t)K * len), 0) : 0;  // host-callback ranks: in place in the pinned staging buffer
  if (rc) return rc;
  // the reduced buffer { W_gamma, s1[d], S2[d*d] } x K stays in the pinned staging area and is read in place by the
  // finalisation (strides = len); only the Student-t weight sums live elsewhere
  if (student) W = wsum;
  view.G = c->hstage;
  view.sG = len;
  view.W = student ? W.data() : c->hstage;
  view.sW = student ? 1 : len;
  view.s1 = c->hstage + 1;
  view.ss1 = len;
  view.S2 = c->hstage + 1 + d;
  view.sS2 = len;
  return 0;
}

int run_weights_generic(pmcgpu_ctx *c, int mode, const double *post_in, const short *ok_in, double t_temper,
                        long first_index, WeightOut &o) {
  if (mode == 0 && c->tkind == PMCGPU_TGT_COMBINE) {
    CU(c->hostpost.ensure((size_t)c->N * sizeof(double)));
    RC(eval_combine(c, c->N, c->X.as<double>(), c->hostpost.as<double>()));
    mode = 1;
    post_in = c->hostpost.as<double>();
    ok_in = nullptr;
  }
  RC(ensure_mma(c));
  if (mma_usable(c, mode)) return run_weights_mma(c, mode, post_in, ok_in, t_temper, first_index, o);
  c->q_resident = false;
  const long N = c->N;
  const int Kmax = c->prop.K > c->tmix.K ? c->prop.K : c->tmix.K;
  const long chunk = N < GEN_CHUNK ? N : GEN_CHUNK;
  CU(c->ld.ensure((size_t)chunk * Kmax * sizeof(double)));
  const int nb = (int)((chunk + GEN_WEIGHT_THREADS - 1) / GEN_WEIGHT_THREADS);
  CU(c->blkmax.ensure((size_t)2 * nb * sizeof(double)));
  int rc = zero_small(c);
  if (rc) return rc;
  unsigned long long *sm = c->small.as<unsigned long long>();
  const TargetDev tv = target_view(c);
  RC(stage(c, (size_t)2 * nb + SM_LEN));
  for (long base = 0; base < N; base += chunk) {
    const long n = (N - base) < chunk ? (N - base) : chunk;
    const int nbc = (int)((n + GEN_WEIGHT_THREADS - 1) / GEN_WEIGHT_THREADS);
    launch_weights_generic(c->dprop.view, tv, mode, post_in ? post_in + base : nullptr, ok_in ? ok_in + base : nullptr,
                           t_temper, first_index + base, n, c->X.as<double>() + (size_t)base * c->d,
                           c->W.as<double>() + base, c->R.as<double>() + base, c->Flg.as<short>() + base,
                           c->ld.as<double>(), Kmax, c->blkmax.as<double>(), c->blkmax.as<double>() + nb,
                           sm + SM_COUNTERS, reinterpret_cast<double *>(sm + SM_FIRST), c->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->hstage, c->blkmax.p, (size_t)2 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int b = 0; b < nbc; ++b) {
      if (c->hstage[b] > o.M_all) o.M_all = c->hstage[b];
      if (c->hstage[nb + b] > o.R_all) o.R_all = c->hstage[nb + b];
    }
  }
  rc = read_small(c, o.counters, nullptr, o.first);
  o.has_first = (first_index == 0);
  return rc;
}

int normalize_impl(pmcgpu_ctx *c, double maxW, double *S_out, long *count_out, bool reduce_ranks) {
  const int nb = c->sm_count * 8;
  CU(c->blkmax.ensure((size_t)2 * nb * sizeof(double)));
  int rc = zero_small(c);
  if (rc) return rc;
  unsigned long long *sm = c->small.as<unsigned long long>();
  launch_norm_exp_sum(c->N, maxW, c->W.as<double>(), c->Flg.as<short>(), c->blkmax.as<double>(), nb, sm + SM_COUNTERS, c->stream);
  CU(cudaGetLastError());
  RC(stage(c, (size_t)nb + SM_LEN));
  CU(cudaMemcpyAsync(c->hstage, c->blkmax.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->hstage + nb, c->small.p, SM_LEN * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  double red[2] = {0.0, 0.0};
  for (int b = 0; b < nb; ++b) red[0] += c->hstage[b];
  red[1] = (double)reinterpret_cast<const unsigned long long *>(c->hstage + nb)[SM_COUNTERS + CNT_NORMCOUNT];
  if (reduce_ranks) { rc = allreduce_host(c, red, 2, 0); if (rc) return rc; }
  *S_out = red[0];
  *count_out = (long)red[1];
  return 0;
}

int scale_perp_impl(pmcgpu_ctx *c, double S, bool scale, double *H, double *Q, long *nexit, bool reduce_ranks) {
  const int nb = c->sm_count * 8;
  CU(c->blkmax.ensure((size_t)2 * nb * sizeof(double)));
  int rc = zero_small(c);
  if (rc) return rc;
  unsigned long long *sm = c->small.as<unsigned long long>();
  if (scale) launch_norm_scale_perp(c->N, S, c->W.as<double>(), c->Flg.as<short>(), c->blkmax.as<double>(), nb, sm + SM_COUNTERS, c->stream);
  else launch_perplexity(c->N, c->W.as<double>(), c->Flg.as<short>(), c->blkmax.as<double>(), nb, sm + SM_COUNTERS, c->stream);
  CU(cudaGetLastError());
  RC(stage(c, (size_t)2 * nb + SM_LEN));
  CU(cudaMemcpyAsync(c->hstage, c->blkmax.p, (size_t)2 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->hstage + 2 * nb, c->small.p, SM_LEN * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  double red[3] = {0.0, 0.0, 0.0};
  for (int b = 0; b < nb; ++b) { red[0] += c->hstage[2 * b]; red[1] += c->hstage[2 * b + 1]; }
  red[2] = (double)reinterpret_cast<const unsigned long long *>(c->hstage + 2 * nb)[SM_COUNTERS + CNT_FLAGGED];
  if (reduce_ranks) { rc = allreduce_host(c, red, 3, 0); if (rc) return rc; }
  *H = red[0];
  *Q = red[1];
  *nexit = (long)red[2];
  return 0;
}

// filter_histogram (pmc.c:709-770) on the resident weights
int filter_histogram_impl(pmcgpu_ctx *c, int nbins, int cl, int cr, long first_index, bool reduce_ranks, long *n_filtered) {
  const int nb = c->sm_count * 8;
  CU(c->blkmax.ensure((size_t)2 * nb * sizeof(double)));
  int rc = zero_small(c);
  if (rc) return rc;
  unsigned long long *sm = c->small.as<unsigned long long>();
  launch_minmax_ok(c->N, c->W.as<double>(), c->Flg.as<short>(), c->blkmax.as<double>(), nb, sm + SM_COUNTERS + CNT_NOK, c->stream);
  CU(cudaGetLastError());
  RC(stage(c, (size_t)2 * nb + SM_LEN + 1));
  CU(cudaMemcpyAsync(c->hstage, c->blkmax.p, (size_t)2 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->hstage + 2 * nb, c->small.p, SM_LEN * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->hstage + 2 * nb + SM_LEN, c->W.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  // red[0] = max, red[1] = -min over ok samples; red[2], red[3] = w0, -w0 from the rank that owns global sample 0;
  // red[4] = 1 if that w0 is NaN (then every comparison of the reference is false and nothing is filtered)
  double red[5] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, 0.0};
  for (int b = 0; b < nb; ++b) {
    if (c->hstage[2 * b] > red[0]) red[0] = c->hstage[2 * b];
    if (c->hstage[2 * b + 1] > red[1]) red[1] = c->hstage[2 * b + 1];
  }
  double cnt = (double)reinterpret_cast<const unsigned long long *>(c->hstage + 2 * nb)[SM_COUNTERS + CNT_NOK];
  if (first_index == 0) {
    const double w0 = c->hstage[2 * nb + SM_LEN];
    if (std::isnan(w0)) red[4] = 1.0;
    else { red[2] = w0; red[3] = -w0; }
  }
  if (reduce_ranks) {
    rc = allreduce_host(c, red, 5, 1);
    if (rc) return rc;
    rc = allreduce_host(c, &cnt, 1, 0);
    if (rc) return rc;
  }
  if (n_filtered) *n_filtered = 0;
  if (cnt == 0.0 || red[4] != 0.0) return 0;  // no ok sample: min = max = 0 and the last loop has nothing to test
  double maxW = red[2], minW = -red[3];        // both start from weights[0] (pmc.c:734-735)
  if (red[0] > maxW) maxW = red[0];
  if (-red[1] < minW) minW = -red[1];
  const double step = (maxW - minW) * 1. / nbins;
  const double bl = minW + step * cl, br = maxW - step * cr;
  rc = zero_small(c);
  if (rc) return rc;
  launch_flag_outside(c->N, c->W.as<double>(), c->Flg.as<short>(), bl, br, nb, sm + SM_COUNTERS + CNT_FLAGGED, c->stream);
  CU(cudaGetLastError());
  unsigned long long counters[CNT_NUM];
  rc = read_small(c, counters, nullptr, nullptr);
  if (rc) return rc;
  double nf = (double)counters[CNT_FLAGGED];
  if (reduce_ranks) { rc = allreduce_host(c, &nf, 1, 0); if (rc) return rc; }
  if (n_filtered) *n_filtered = (long)nf;
  return 0;
}

int count_indices_impl(pmcgpu_ctx *c, std::vector<size_t> &count, bool reduce_ranks) {
  const int K = c->prop.K;
  CU(c->tmpo.ensure((size_t)K * 8));
  CU(cudaMemsetAsync(c->tmpo.p, 0, (size_t)K * 8, c->stream));
  launch_count_indices(c->N, K, c->Idx.as<unsigned long long>(), c->Flg.as<short>(), c->tmpo.as<unsigned long long>(), c->sm_count * 4, c->stream);
  CU(cudaGetLastError());
  RC(stage(c, K));
  CU(cudaMemcpyAsync(c->hstage, c->tmpo.p, (size_t)K * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<double> tmp(K);
  for (int k = 0; k < K; ++k) tmp[k] = (double)reinterpret_cast<const unsigned long long *>(c->hstage)[k];
  if (reduce_ranks) { int rc = allreduce_host(c, tmp.data(), K, 0); if (rc) return rc; }
  count.resize(K);
  for (int k = 0; k < K; ++k) count[k] = (size_t)tmp[k];
  return 0;
}

void export_proposal(const pmcgpu_ctx *c, double *o_wght, double *o_mean, double *o_L, double *o_detL, int *o_alive) {
  const HostMix &m = c->prop;
  if (o_wght) memcpy(o_wght, m.wght.data(), m.K * sizeof(double));
  if (o_mean) memcpy(o_mean, m.mean.data(), (size_t)m.K * m.d * sizeof(double));
  if (o_L) memcpy(o_L, m.L.data(), (size_t)m.K * m.d * m.d * sizeof(double));
  if (o_detL) memcpy(o_detL, m.detL.data(), m.K * sizeof(double));
  if (o_alive) for (int k = 0; k < m.K; ++k) o_alive[k] = m.wght[k] > 0.0;
}

// two-pass update exactly as the reference (pmc.c:1323-1469): weights must be normalised, log_rho/maxR as stored
int update_two_pass(pmcgpu_ctx *c, double maxR, long N_total, int hard, int *retry, const std::vector<size_t> &count) {
  HostMix &m = c->prop;
  const int K = m.K, d = m.d;
  const long N = c->N;
  const long chunk = N < GEN_CHUNK ? N : GEN_CHUNK;
  const size_t dd = (size_t)d * d;
  CU(c->rbg.ensure((size_t)chunk * K * sizeof(double)));
  CU(c->rbw.ensure((size_t)chunk * K * sizeof(double)));
  const int nsplit = (int)std::max<long>(1, std::min<long>(64, chunk / 8192));
  CU(c->rbpart.ensure((size_t)nsplit * K * (dd + d + 2) * sizeof(double)));
  CU(c->rbacc.ensure((size_t)K * (dd + d + 2) * sizeof(double)));
  CU(c->tmpo.ensure((size_t)K * 8));
  CU(cudaMemsetAsync(c->rbacc.p, 0, (size_t)K * (dd + d + 2) * sizeof(double), c->stream));
  {
    RC(stage(c, K));
    unsigned long long *hc = reinterpret_cast<unsigned long long *>(c->hstage);
    for (int k = 0; k < K; ++k) hc[k] = count[k];
    CU(cudaMemcpyAsync(c->tmpo.p, hc, (size_t)K * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  double *acc_mean = c->rbacc.as<double>();           // [K*(d+2)]
  double *acc_cov = acc_mean + (size_t)K * (d + 2);   // [K*d*d]
  // pass A
  for (long base = 0; base < N; base += chunk) {
    const long n = (N - base) < chunk ? (N - base) : chunk;
    launch_rb_resp(c->dprop.view, n, base, c->X.as<double>(), c->W.as<double>(), c->R.as<double>(), c->Flg.as<short>(),
                   c->Idx.as<unsigned long long>(), c->tmpo.as<unsigned long long>(), maxR, hard, c->rbg.as<double>(), c->rbw.as<double>(), c->stream);
    launch_rb_reduce_mean(K, d, n, base, nsplit, c->X.as<double>(), c->rbg.as<double>(), c->rbw.as<double>(), c->rbpart.as<double>(), acc_mean, c->stream);
    CU(cudaGetLastError());
  }
  std::vector<double> am((size_t)K * (d + 2));
  RC(stage(c, am.size()));
  CU(cudaMemcpyAsync(c->hstage, acc_mean, am.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(am.data(), c->hstage, am.size() * sizeof(double));
  int rc = allreduce_host(c, am.data(), (long)am.size(), 0);
  if (rc) return rc;
  // new means (pmc.c:1409-1426)
  std::vector<double> newmean((size_t)K * d, 0.0), W(K, 0.0), G(K, 0.0);
  for (int k = 0; k < K; ++k) {
    G[k] = am[(size_t)k * (d + 2) + d];
    W[k] = am[(size_t)k * (d + 2) + d + 1];
    if (count[k] > (size_t)kMinCount)
      for (int j = 0; j < d; ++j) newmean[(size_t)k * d + j] = am[(size_t)k * (d + 2) + j] / G[k];
  }
  CU(c->newmean.ensure(newmean.size() * sizeof(double)));
  RC(stage(c, newmean.size()));
  memcpy(c->hstage, newmean.data(), newmean.size() * sizeof(double));
  CU(cudaMemcpyAsync(c->newmean.p, c->hstage, newmean.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  // pass B
  for (long base = 0; base < N; base += chunk) {
    const long n = (N - base) < chunk ? (N - base) : chunk;
    if (N > chunk)  // responsibilities of a single chunk are still in the scratch from pass A
      launch_rb_resp(c->dprop.view, n, base, c->X.as<double>(), c->W.as<double>(), c->R.as<double>(), c->Flg.as<short>(),
                     c->Idx.as<unsigned long long>(), c->tmpo.as<unsigned long long>(), maxR, hard, c->rbg.as<double>(), c->rbw.as<double>(), c->stream);
    launch_rb_reduce_cov(K, d, n, base, nsplit, c->X.as<double>(), c->rbg.as<double>(), c->newmean.as<double>(), c->rbpart.as<double>(), acc_cov, c->stream);
    CU(cudaGetLastError());
  }
  std::vector<double> cov(K * dd);
  RC(stage(c, cov.size()));
  CU(cudaMemcpyAsync(c->hstage, acc_cov, cov.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(cov.data(), c->hstage, cov.size() * sizeof(double));
  rc = allreduce_host(c, cov.data(), (long)cov.size(), 0);
  if (rc) return rc;
  // finish (pmc.c:1455-1462) + cleanup
  std::vector<double> snapshot(m.L);
  for (int k = 0; k < K; ++k) {
    double *Ck = m.L.data() + k * dd;
    m.wght[k] = W[k];
    if (count[k] > (size_t)kMinCount) {
      const size_t bl = m.band.empty() ? (size_t)d : m.band[k];
      const double inv = 1.0 / m.wght[k];
      for (int a = 0; a < d; ++a)
        for (int b = a; b < d; ++b) {
          const double v = ((size_t)(b - a) < bl) ? cov[k * dd + (size_t)a * d + b] * inv : 0.0;
          Ck[(size_t)a * d + b] = v;
          Ck[(size_t)b * d + a] = v;
        }
      memcpy(m.mean.data() + (size_t)k * d, newmean.data() + (size_t)k * d, d * sizeof(double));
    } else {
      memset(Ck, 0, dd * sizeof(double));
      memset(m.mean.data() + (size_t)k * d, 0, d * sizeof(double));
    }
  }
  rc = host_cleanup_after_update(K, d, count.data(), 1.0 / (double)N_total, retry, snapshot.data(), m.wght.data(), m.mean.data(), m.L.data(), m.detL.data());
  if (rc) return fail(c, rc, "Sum of weight is <= 0 after the update");
  return 0;
}

// dead components keep whatever detL they had; log() of it must stay finite for the device tables
void sanitize_dead(HostMix &m) {
  for (int k = 0; k < m.K; ++k)
    if (!(m.wght[k] > 0.0) || !(m.detL[k] > 0.0)) { if (!(m.detL[k] > 0.0)) m.detL[k] = 1.0; }
}

}  // namespace

extern "C" {

int pmcgpu_version(void) { return 100; }

int pmcgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int pmcgpu_create(int device, pmcgpu_ctx **out) {
  if (!out) return PMCGPU_ERR_USAGE;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PMCGPU_ERR_NODEVICE;
  if (device < 0 || device >= ndev) return PMCGPU_ERR_NODEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return PMCGPU_ERR_CUDA;
  pmcgpu_ctx *c = new pmcgpu_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return PMCGPU_ERR_CUDA; }
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return PMCGPU_ERR_CUDA; }
  for (int j = 0; j < FUSED_MAXD; ++j) { c->lo_thr[j] = INFINITY; c->hi_thr[j] = INFINITY; }
  {  // ziggurat table: x_0 = V/f(R), x_1 = R, x_i = sqrt(-2 ln(V/x_{i-1} + f(x_{i-1}))), x_C = 0; entries (x_i, x_{i+1}/x_i)
    std::vector<double> x(kZigLayers + 1), tab(2 * kZigLayers);
    auto f = [](double v) { return std::exp(-0.5 * v * v); };
    x[0] = kZigV / f(kZigR);
    x[1] = kZigR;
    for (int i = 2; i < kZigLayers; ++i) x[i] = std::sqrt(-2.0 * std::log(kZigV / x[i - 1] + f(x[i - 1])));
    x[kZigLayers] = 0.0;
    for (int i = 0; i < kZigLayers; ++i) { tab[2 * i] = x[i]; tab[2 * i + 1] = x[i + 1] / x[i]; }
    if (c->zig.ensure(tab.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(c->zig.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
      pmcgpu_destroy(c);
      return PMCGPU_ERR_CUDA;
    }
  }
  if (const char *e = getenv("PMCB200_FORCE_GENERIC")) c->force_generic = atoi(e);
  if (const char *e = getenv("PMCB200_SPL")) c->spl = atoi(e);
  *out = c;
  return 0;
}

void pmcgpu_destroy(pmcgpu_ctx *c) {
  if (c) sink_join(c);
  if (c && c->copy_stream) { cudaSetDevice(c->device); cudaStreamDestroy(c->copy_stream); if (c->ev_copy) cudaEventDestroy(c->ev_copy); if (c->ev_idx8) cudaEventDestroy(c->ev_idx8); }
  if (c && c->hidx8) { cudaSetDevice(c->device); cudaFreeHost(c->hidx8); c->hidx8 = nullptr; }
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
  DevBuf *bufs[] = {&c->dprop.buf, &c->dprop_pack, &c->dtmix.buf, &c->dt_pack, &c->ddef, &c->dfaces, &c->X, &c->W, &c->R,
                    &c->Idx, &c->Flg, &c->ld, &c->blkmax, &c->partials, &c->stats, &c->small, &c->rbg, &c->rbw, &c->rbacc,
                    &c->rbpart, &c->newmean, &c->hostpost, &c->hostok, &c->tmpx, &c->tmpo, &c->dcoll, &c->dshare, &c->dcomb, &c->dcombdim, &c->resp, &c->partials2, &c->dscal, &c->zig,
                    &c->i3_gpack, &c->i3_wpart, &c->i3_scal, &c->i3_coll, &c->i3_spart, &c->i3_out, &c->didx8, &c->i3_gath};
  for (DevBuf *b : bufs) b->release();
  for (auto &t : c->comb) { t.dmix.buf.release(); t.dmix.pin.release(); }
  c->dprop.pin.release(); c->dtmix.pin.release(); c->pin_ppack.release(); c->pin_tpack.release();
  if (c->hstage) cudaFreeHost(c->hstage);
  if (c->ev0) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev2); cudaEventDestroy(c->evd); }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *pmcgpu_last_error(const pmcgpu_ctx *c) { return c ? c->err.c_str() : "null context"; }
void *pmcgpu_stream(pmcgpu_ctx *c) { return c ? (void *)c->stream : nullptr; }

static int repack_mma(pmcgpu_ctx *c) {
  c->mprop.valid = c->mtgt.valid = false;
  c->mprop.dirty = c->prop.K != 0;
  c->mtgt.dirty = c->tmix.K != 0;
  c->q_resident = false;
  return 0;
}

int pmcgpu_set_option(pmcgpu_ctx *c, const char *name, double value) {
  if (!c || !name) return PMCGPU_ERR_USAGE;
  if (!strcmp(name, "force_generic")) { c->force_generic = (int)value; return repack_mma(c); }
  else if (!strcmp(name, "spl")) c->spl = (int)value;
  else if (!strcmp(name, "fused_version")) c->fused_version = (int)value;
  else if (!strcmp(name, "max_attempts")) c->max_attempts = (int)value;
  else if (!strcmp(name, "force_precise")) c->force_precise = (int)value;
  else if (!strcmp(name, "guard_ratio")) c->guard_ratio = value;
  else if (!strcmp(name, "timing")) { c->timing = (int)value; c->fused_ms = 0.0; c->stats_ms = 0.0; c->draw_ms = 0.0; c->fused_launches = 0; }
  else if (!strcmp(name, "split_stats")) c->split_stats = (int)value;
  else if (!strcmp(name, "split_draw")) c->split_draw = (int)value;
  else if (!strcmp(name, "gen3")) c->gen3 = (int)value;
  else if (!strcmp(name, "gen3_max_ratio")) { c->gen3_max_ratio = value; c->i3_dirty = true; }
  else if (!strcmp(name, "sink_idx8")) c->sink_idx8 = (int)value;
  else if (!strcmp(name, "w3_variant")) { c->w3_variant = (int)value; c->i3_dirty = true; }
  else if (!strcmp(name, "draw3")) c->draw3 = (int)value;
  else if (!strcmp(name, "mma")) { c->use_mma = (int)value; return repack_mma(c); }
  else if (!strcmp(name, "mma_min_d")) { c->mma_min_d = (int)value; return repack_mma(c); }
  else if (!strcmp(name, "mma_max_chunk")) c->mma_max_chunk = (long)value;
  else if (!strcmp(name, "mma_inv_tol")) { c->mma_inv_tol = value; return repack_mma(c); }
  else if (!strcmp(name, "t_temper")) { if (!(value >= 0)) return fail(c, -6019, "t_iter_tempered=%g should be > 0", value); c->t_temper = value; }
  else return fail(c, PMCGPU_ERR_USAGE, "unknown option %s", name);
  return 0;
}

// read-only diagnostics: "mma_active" (1 if the proposal is packed for the tensor path), "mma_inverse_discrepancy"
// (largest relative difference between Linv b and the forward substitution over the proposal's components, see
// pmc_mma.cu:inverse_discrepancy; the tensor path is used only while it is below option "mma_inv_tol")
int pmcgpu_get_info(pmcgpu_ctx *c, const char *name, double *value) {
  if (!c || !name || !value) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  if (!strcmp(name, "mma_active")) { RC(ensure_mma(c)); *value = c->mprop.valid ? 1.0 : 0.0; return 0; }
  if (!strcmp(name, "mma_inverse_discrepancy")) { RC(ensure_mma(c)); *value = c->mprop.inv_discrepancy; return 0; }
  // third-generation weights kernel: 1 if the current proposal is packed for the row-scaled solve, and the guard's ratio
  if (!strcmp(name, "sink_bytes")) { *value = c->sink_bytes; return 0; }
  if (!strcmp(name, "w3_rowscaled")) { const Iter3Plan p = plan3_for(c); *value = (p.valid && c->i3_pre) ? 1.0 : 0.0; return 0; }
  if (!strcmp(name, "w3_shift_ratio")) { const Iter3Plan p = plan3_for(c); *value = p.valid ? c->i3_pre_ratio : -1.0; return 0; }
  return fail(c, PMCGPU_ERR_USAGE, "unknown info %s", name);
}

int pmcgpu_get_timing(pmcgpu_ctx *c, double *fused_ms, long *fused_launches, long *total_launches) {
  if (!c) return PMCGPU_ERR_USAGE;
  if (fused_ms) *fused_ms = c->fused_ms;
  if (fused_launches) *fused_launches = c->fused_launches;
  if (total_launches) *total_launches = launch_count();
  return 0;
}

int pmcgpu_get_stats_timing(pmcgpu_ctx *c, double *stats_ms) {
  if (!c || !stats_ms) return PMCGPU_ERR_USAGE;
  *stats_ms = c->stats_ms;
  return 0;
}
int pmcgpu_get_draw_timing(pmcgpu_ctx *c, double *draw_ms) {
  if (!c || !draw_ms) return PMCGPU_ERR_USAGE;
  *draw_ms = c->draw_ms;
  return 0;
}

int pmcgpu_set_proposal(pmcgpu_ctx *c, int K, int d, const double *wght, const double *mean, const double *L,
                        const double *detL, const int *df) {
  if (!c || !wght || !mean || !L) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  int rc = fill_mix(c, c->prop, K, d, wght, mean, L, detL, df);
  if (rc) return rc;
  c->prop.band.clear();
  sanitize_dead(c->prop);
  return install_proposal(c);
}

int pmcgpu_set_band_limit(pmcgpu_ctx *c, const size_t *band_limit) {
  if (!c || c->prop.K == 0) return PMCGPU_ERR_USAGE;
  if (band_limit) c->prop.band.assign(band_limit, band_limit + c->prop.K);
  else c->prop.band.clear();
  return 0;
}

int pmcgpu_set_target_banana(pmcgpu_ctx *c, int d, double b, double sig1) {
  if (!c || d < 2) return fail(c, PMCGPU_ERR_USAGE, "bentGauss needs d >= 2");
  c->i3_dirty = true;
  c->tkind = PMCGPU_TGT_BANANA;
  c->td = d;
  c->tb = b;
  c->tsig1 = sig1;
  c->tmix = HostMix();
  c->mtgt.valid = false;
  return 0;
}

int pmcgpu_set_target_mix(pmcgpu_ctx *c, int kind, int K, int d, const double *wght, const double *mean,
                          const double *L, const double *detL, const int *df) {
  if (!c || (kind != PMCGPU_TGT_MVDENS && kind != PMCGPU_TGT_MIX)) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  std::vector<double> one(1, 1.0);
  if (kind == PMCGPU_TGT_MVDENS) { K = 1; if (!wght) wght = one.data(); }
  int rc = fill_mix(c, c->tmix, K, d, wght, mean, L, detL, df);
  if (rc) return rc;
  c->tkind = kind;
  c->td = d;
  rc = upload_mix(c, c->tmix, c->dtmix);
  if (rc) return rc;
  c->mtgt.valid = false;
  c->mtgt.dirty = true;
  c->i3_dirty = true;
  return pack_small(c, c->tmix, c->dt_pack, c->t_pack_doubles, c->tgt_pack_host, c->pin_tpack);
}

int pmcgpu_set_target_combine(pmcgpu_ctx *c, int d, int nterms, const pmcgpu_combine_term *terms) {
  if (!c || d < 1 || d > PMCB200_MAXD || nterms < 1 || !terms) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  std::vector<pmcgpu_ctx::CombTerm> comb(nterms);
  std::vector<int> alldim;
  int kmax = 1;
  for (int t = 0; t < nterms; ++t) {
    const pmcgpu_combine_term &in = terms[t];
    pmcgpu_ctx::CombTerm &o = comb[t];
    if (in.ndim < 1 || in.ndim > PMCB200_MAXD || !in.dim_idx) return fail(c, PMCGPU_ERR_USAGE, "combined target: term %d has no coordinates", t);
    for (int a = 0; a < in.ndim; ++a)
      if (in.dim_idx[a] < 0 || in.dim_idx[a] >= d) return fail(c, PMCGPU_ERR_USAGE, "combined target: term %d reads coordinate %d of %d", t, in.dim_idx[a], d);
    o.kind = in.kind;
    o.dim.assign(in.dim_idx, in.dim_idx + in.ndim);
    alldim.insert(alldim.end(), o.dim.begin(), o.dim.end());
    if (in.kind == PMCGPU_TGT_BANANA) {
      if (in.ndim < 2) return fail(c, PMCGPU_ERR_USAGE, "combined target: bentGauss term needs ndim >= 2");
      o.b = in.b;
      o.sig1 = in.sig1;
    } else if (in.kind == PMCGPU_TGT_MVDENS || in.kind == PMCGPU_TGT_MIX) {
      const double one = 1.0;
      const int K = in.kind == PMCGPU_TGT_MVDENS ? 1 : in.K;
      RC(fill_mix(c, o.mix, K, in.ndim, (in.kind == PMCGPU_TGT_MVDENS && !in.wght) ? &one : in.wght, in.mean, in.L, in.detL, in.df));
      RC(upload_mix(c, o.mix, o.dmix));
      if (K > kmax) kmax = K;
    } else {
      return fail(c, PMCGPU_ERR_USAGE, "combined target: term %d has kind %d", t, in.kind);
    }
  }
  CU(c->dcombdim.ensure(alldim.size() * sizeof(int)));
  CU(cudaMemcpy(c->dcombdim.p, alldim.data(), alldim.size() * sizeof(int), cudaMemcpyHostToDevice));
  std::vector<CombTermDev> dev(nterms);
  size_t off = 0;
  for (int t = 0; t < nterms; ++t) {
    dev[t].kind = comb[t].kind;
    dev[t].nd = (int)comb[t].dim.size();
    dev[t].dim = c->dcombdim.as<int>() + off;
    off += comb[t].dim.size();
    dev[t].b = comb[t].b;
    dev[t].sig1 = comb[t].sig1;
    dev[t].mix = comb[t].dmix.view;
  }
  CU(c->dcomb.ensure(dev.size() * sizeof(CombTermDev)));
  CU(cudaMemcpy(c->dcomb.p, dev.data(), dev.size() * sizeof(CombTermDev), cudaMemcpyHostToDevice));
  for (auto &t : c->comb) { t.dmix.buf.release(); t.dmix.pin.release(); }
  c->dprop.pin.release(); c->dtmix.pin.release(); c->pin_ppack.release(); c->pin_tpack.release();
  c->comb = std::move(comb);
  c->comb_kmax = kmax;
  c->tkind = PMCGPU_TGT_COMBINE;
  c->td = d;
  c->tmix = HostMix();
  c->mtgt.valid = false;
  c->i3_dirty = true;
  return 0;
}

int pmcgpu_set_target_host(pmcgpu_ctx *c, int d) {
  if (!c) return PMCGPU_ERR_USAGE;
  c->tkind = PMCGPU_TGT_HOST;
  c->td = d;
  c->tmix = HostMix();
  c->mtgt.valid = false;
  return 0;
}

int pmcgpu_set_target_defaults(pmcgpu_ctx *c, int nfull, const int *is_def, const double *value) {
  if (!c) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  if (nfull <= 0) { c->nfull = 0; return 0; }
  if (nfull > PMCB200_MAXD || !is_def || !value) return fail(c, PMCGPU_ERR_USAGE, "bad defaults map");
  c->is_def.assign(is_def, is_def + nfull);
  c->def_val.assign(value, value + nfull);
  const size_t ioff = ((nfull * sizeof(int) + 15) / 16) * 16;
  std::vector<char> blob(ioff + nfull * sizeof(double));
  memcpy(blob.data(), is_def, nfull * sizeof(int));
  memcpy(blob.data() + ioff, value, nfull * sizeof(double));
  CU(c->ddef.ensure(blob.size()));
  CU(cudaMemcpy(c->ddef.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  c->nfull = nfull;
  return 0;
}

int pmcgpu_set_box(pmcgpu_ctx *c, int d, int nfaces, const double *faces) {
  if (!c) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  c->nfaces = 0;
  c->box_slab = false;
  c->box_slab_any = false;
  for (int j = 0; j < PMCB200_MAXD; ++j) { c->lo_thr[j] = INFINITY; c->hi_thr[j] = INFINITY; }
  if (nfaces <= 0 || !faces) return 0;
  if (d < 1 || d > PMCB200_MAXD) return fail(c, PMCGPU_ERR_USAGE, "box dimension %d", d);
  c->faces.assign(faces, faces + (size_t)nfaces * (d + 1));
  c->box_d = d;
  CU(c->dfaces.ensure(c->faces.size() * sizeof(double)));
  CU(cudaMemcpy(c->dfaces.p, c->faces.data(), c->faces.size() * sizeof(double), cudaMemcpyHostToDevice));
  c->nfaces = nfaces;
  // canonical slab form: every face is -e_j or +e_j (then thr - c.x reduces to one exact-coefficient operation and the
  // tightest threshold per axis decides, rounding being monotone) -> the fused kernel can test it bit-exactly
  bool slab = true;
  for (int f = 0; f < nfaces && slab; ++f) {
    const double *cf = faces + (size_t)f * (d + 1);
    int axis = -1;
    for (int j = 0; j < d; ++j) {
      if (cf[j] != 0.0) { if (axis >= 0 || (cf[j] != 1.0 && cf[j] != -1.0)) { slab = false; break; } axis = j; }
    }
    if (!slab) break;
    if (axis < 0) { if (cf[d] < 0) slab = false; continue; }
    if (cf[axis] < 0) { if (cf[d] < c->lo_thr[axis]) c->lo_thr[axis] = cf[d]; }
    else { if (cf[d] < c->hi_thr[axis]) c->hi_thr[axis] = cf[d]; }
  }
  c->box_slab = slab && d <= FUSED_MAXD;
  c->box_slab_any = slab;
  if (slab) {  // thresholds of the -e_j faces, then of the +e_j faces, for the tensor-path draw
    std::vector<double> thr(2 * (size_t)d);
    for (int j = 0; j < d; ++j) { thr[j] = c->lo_thr[j]; thr[d + j] = c->hi_thr[j]; }
    CU(c->dthr.ensure(thr.size() * sizeof(double)));
    CU(cudaMemcpy(c->dthr.p, thr.data(), thr.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  return 0;
}

int pmcgpu_resize(pmcgpu_ctx *c, long N, int d) {
  if (!c || N < 0 || d < 1 || d > PMCB200_MAXD) return fail(c, PMCGPU_ERR_USAGE, "resize(N=%ld, d=%d)", N, d);
  cudaSetDevice(c->device);
  CU(c->X.ensure((size_t)N * d * sizeof(double) + 16));
  CU(c->W.ensure((size_t)N * sizeof(double) + 16));
  CU(c->R.ensure((size_t)N * sizeof(double) + 16));
  CU(c->Idx.ensure((size_t)N * sizeof(unsigned long long) + 16));
  CU(c->Flg.ensure((size_t)N * sizeof(short) + 16));
  c->N = N;
  c->d = d;
  return 0;
}

int pmcgpu_upload(pmcgpu_ctx *c, const double *X, const size_t *indices, const double *weights, const double *log_rho, const short *flg) {
  if (!c) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  const size_t N = c->N;
  if (X) CU(cudaMemcpyAsync(c->X.p, X, N * c->d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (indices) CU(cudaMemcpyAsync(c->Idx.p, indices, N * sizeof(size_t), cudaMemcpyHostToDevice, c->stream));
  if (weights) CU(cudaMemcpyAsync(c->W.p, weights, N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (log_rho) CU(cudaMemcpyAsync(c->R.p, log_rho, N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (flg) CU(cudaMemcpyAsync(c->Flg.p, flg, N * sizeof(short), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int pmcgpu_download(pmcgpu_ctx *c, double *X, size_t *indices, double *weights, double *log_rho, short *flg) {
  if (!c) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  const size_t N = c->N;
  if (X) CU(cudaMemcpyAsync(X, c->X.p, N * c->d * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (indices) CU(cudaMemcpyAsync(indices, c->Idx.p, N * sizeof(size_t), cudaMemcpyDeviceToHost, c->stream));
  if (weights) CU(cudaMemcpyAsync(weights, c->W.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (log_rho) CU(cudaMemcpyAsync(log_rho, c->R.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (flg) CU(cudaMemcpyAsync(flg, c->Flg.p, N * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int pmcgpu_device_arrays(pmcgpu_ctx *c, void **out) {
  if (!c || !out) return PMCGPU_ERR_USAGE;
  out[0] = c->X.p; out[1] = c->Idx.p; out[2] = c->W.p; out[3] = c->R.p; out[4] = c->Flg.p;
  return 0;
}

// Grouped draw on the tensor path (no box): returns 1 if the configuration is not covered (caller uses the generic kernel)
static int draw_mma(pmcgpu_ctx *c, uint64_t seed, uint32_t iter, long first_index) {
  if (c->nfaces > 0 && (!c->box_slab_any || c->box_d != c->d)) return 1;
  RC(ensure_mma(c));
  const pmcgpu_ctx::MmaMix &mm = c->mprop;
  if (!c->use_mma || c->force_generic || !mm.valid || c->N >= (1L << 31)) return 1;
  const int K = c->prop.K, d = c->prop.d, J = mm.J;
  const long N = c->N;
  CU(c->dcnt.ensure((size_t)(2 * K + 1) * sizeof(unsigned int)));
  const bool boxed = c->nfaces > 0;
  if (boxed) CU(c->dredo.ensure((size_t)N * sizeof(unsigned int)));
  CU(c->dperm.ensure((size_t)N * sizeof(unsigned int)));
  CU(cudaMemsetAsync(c->dcnt.p, 0, (size_t)K * sizeof(unsigned int), c->stream));
  if (c->timing) {
    if (!c->ev0) { CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1)); CU(cudaEventCreate(&c->ev2)); CU(cudaEventCreate(&c->evd)); }
    CU(cudaEventRecord(c->ev0, c->stream));
  }
  launch_draw_pick(c->dprop.view.cw, K, seed, iter, first_index, N, c->Idx.as<unsigned long long>(), c->dcnt.as<unsigned int>(), c->sm_count * 4, c->stream);
  CU(cudaGetLastError());
  RC(stage(c, K + 2));
  unsigned int *hc = reinterpret_cast<unsigned int *>(c->hstage);
  CU(cudaMemcpyAsync(hc, c->dcnt.p, (size_t)K * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<unsigned int> cnt(hc, hc + K), off(K, 0);
  std::vector<int> jof(K, -1);
  for (int j = 0; j < J; ++j) jof[mm.alive[j]] = j;
  unsigned int acc = 0;
  for (int k = 0; k < K; ++k) {
    off[k] = acc;
    acc += cnt[k];
    if (cnt[k] && jof[k] < 0) return 1;  // a zero-weight component was picked (z == 0 exactly): literal kernel handles it
  }
  const int TS = mma_draw_tile_samples(d);
  std::vector<int> tiles;
  for (int k = 0; k < K; ++k)
    for (unsigned int s0 = 0; s0 < cnt[k]; s0 += TS) {
      tiles.push_back(jof[k]);
      tiles.push_back((int)(off[k] + s0));
      tiles.push_back((int)std::min<unsigned int>(TS, cnt[k] - s0));
      tiles.push_back(0);
    }
  const int ntiles = (int)(tiles.size() / 4);
  CU(c->dtiles.ensure(tiles.size() * sizeof(int) + 16));
  RC(stage(c, (tiles.size() + K) / 2 + 2));
  int *ht = reinterpret_cast<int *>(c->hstage);
  memcpy(ht, tiles.data(), tiles.size() * sizeof(int));
  memcpy(ht + tiles.size(), off.data(), K * sizeof(unsigned int));
  CU(cudaMemcpyAsync(c->dtiles.p, ht, tiles.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  unsigned int *cursor = c->dcnt.as<unsigned int>() + K;
  CU(cudaMemcpyAsync(cursor, ht + tiles.size(), (size_t)K * sizeof(unsigned int), cudaMemcpyHostToDevice, c->stream));
  launch_draw_perm(N, K, c->Idx.as<unsigned long long>(), cursor, c->dperm.as<unsigned int>(), c->sm_count * 8, c->stream);
  CU(cudaGetLastError());
  unsigned int *nredo = c->dcnt.as<unsigned int>() + 2 * K;
  if (boxed) CU(cudaMemsetAsync(nredo, 0, sizeof(unsigned int), c->stream));
  if (launch_draw_mma(mm.lpack.as<double>(), c->dtiles.p, ntiles, c->dperm.as<unsigned int>(), mm.dfj.as<int>(), d, seed, iter,
                      first_index, c->X.as<double>(), c->Flg.as<short>(), c->small.as<unsigned long long>() + SM_COUNTERS,
                      boxed ? c->dthr.as<double>() : nullptr, boxed ? c->dredo.as<unsigned int>() : nullptr, nredo, c->sm_count,
                      c->stream))
    return fail(c, PMCGPU_ERR_CUDA, "draw kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (boxed) {  // the samples whose first attempt left the box: literal kernel, same Philox stream from attempt 1 on
    RC(stage(c, 1));
    unsigned int *hn = reinterpret_cast<unsigned int *>(c->hstage);
    CU(cudaMemcpyAsync(hn, nredo, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const long nr = (long)hn[0];
    if (nr > 0) {
      BoxDev bx{c->nfaces, c->box_d, c->dfaces.as<double>()};
      launch_draw_generic(c->dprop.view, bx, seed, iter, first_index, nr, c->max_attempts, c->X.as<double>(),
                          c->Idx.as<unsigned long long>(), c->Flg.as<short>(), c->small.as<unsigned long long>() + SM_COUNTERS,
                          c->stream, c->dredo.as<unsigned int>(), 1);
      CU(cudaGetLastError());
    }
  }
  if (c->timing) {
    CU(cudaEventRecord(c->evd, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->evd));
    c->draw_ms += ms;
  }
  return 0;
}

// The draw of the store's N samples.  local_check: apply the MNOK / failed-draw test of pmc.c:281-297 to this rank's own
// counters.  Inside the multi-rank iteration the test runs on the REDUCED counters instead (step_core), so that every
// rank takes the same decision and nobody is left waiting in a collective.
static int draw_impl(pmcgpu_ctx *c, uint64_t seed, uint32_t iter, long first_index, long *n_reject, long *n_fail, bool local_check) {
  int rc = check_ready(c, false);
  if (rc) return rc;
  cudaSetDevice(c->device);
  rc = ensure_cwght_normalised(c);
  if (rc) return rc;
  rc = zero_small(c);
  if (rc) return rc;
  int drew2 = 0;
  if (c->gen3 && c->fused_version != 1 && !c->draw3) {
    rc = draw_fused2(c, seed, iter, first_index, &drew2);
    if (rc) return rc;
  }
  rc = drew2 ? 0 : draw_mma(c, seed, iter, first_index);
  if (rc != 0 && rc != 1) return rc;
  if (rc == 1) {
    BoxDev bx{c->nfaces, c->box_d, c->dfaces.as<double>()};
    if (c->nfaces > 0 && c->box_d != c->d) return fail(c, PMCGPU_ERR_USAGE, "box dimension %d != %d", c->box_d, c->d);
    launch_draw_generic(c->dprop.view, bx, seed, iter, first_index, c->N, c->max_attempts, c->X.as<double>(),
                        c->Idx.as<unsigned long long>(), c->Flg.as<short>(), c->small.as<unsigned long long>() + SM_COUNTERS, c->stream);
    CU(cudaGetLastError());
  }
  unsigned long long cnt[CNT_NUM];
  rc = read_small(c, cnt, nullptr, nullptr);
  if (rc) return rc;
  if (n_reject) *n_reject = (long)cnt[CNT_REJECT];
  if (n_fail) *n_fail = (long)cnt[CNT_DRAWFAIL];
  if (local_check && (cnt[CNT_DRAWFAIL] || (long)cnt[CNT_REJECT] > 5 * c->N))
    return fail(c, kErrTooManySteps, "Too many points (%llu) outside of box", cnt[CNT_REJECT]);
  return 0;
}

int pmcgpu_draw(pmcgpu_ctx *c, uint64_t seed, uint32_t iter, long first_index, long *n_reject) {
  return draw_impl(c, seed, iter, first_index, n_reject, nullptr, true);
}

static int weights_common(pmcgpu_ctx *c, int mode, const double *post, const short *ok, double t_temper,
                          long first_index, long *nok, double *maxW, double *maxR) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  cudaSetDevice(c->device);
  if (t_temper < 0) return fail(c, -6019, "t_iter_tempered=%g should be > 0", t_temper);  // pmc.c:508
  const double *dpost = nullptr;
  const short *dok = nullptr;
  if (mode == 1) {
    if (!post) return PMCGPU_ERR_USAGE;
    CU(c->hostpost.ensure((size_t)c->N * sizeof(double)));
    CU(cudaMemcpyAsync(c->hostpost.p, post, (size_t)c->N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dpost = c->hostpost.as<double>();
    if (ok) {
      CU(c->hostok.ensure((size_t)c->N * sizeof(short)));
      CU(cudaMemcpyAsync(c->hostok.p, ok, (size_t)c->N * sizeof(short), cudaMemcpyHostToDevice, c->stream));
      dok = c->hostok.as<short>();
    }
  }
  WeightOut o;
  const int saved_kind = c->tkind;
  if (mode == 1) c->tkind = PMCGPU_TGT_HOST;
  const Iter3Plan p3 = (c->nranks <= 1) ? plan3_for(c) : Iter3Plan{};  // the phase entry points return rank-local maxima
  if (p3.valid) {
    Iter3Out o3;
    rc = run_iter3(c, p3, 0, 1, 0, 0, 0, first_index, c->N, t_temper, dpost, dok, o3);
    c->tkind = saved_kind;
    if (rc) return rc;
    if (nok) *nok = (long)o3.scal[I3_NOK_W];
    if (maxW) *maxW = o3.scal[I3_MAXW];
    if (maxR) *maxR = o3.scal[I3_MAXR];
    return 0;
  }
  const FusedPlan p = plan_for(c);
  if (p.valid) rc = run_fused(c, p, 0, 0, 0, 0, first_index, t_temper, dpost, dok, o);
  else rc = run_weights_generic(c, mode, dpost, dok, t_temper, first_index, o);
  c->tkind = saved_kind;
  if (rc) return rc;
  if (nok) *nok = (long)o.counters[CNT_NOK];
  double mw = o.M_all;
  if (o.counters[CNT_POSINF]) mw = INFINITY;
  if (maxW) *maxW = fold_max(mw, o.has_first, o.first[3] != 0.0, o.first[2]);
  if (maxR) *maxR = fold_max(o.R_all, o.has_first, o.first[1] != 0.0, o.first[0]);
  return 0;
}

int pmcgpu_weights(pmcgpu_ctx *c, double t_temper, long first_index, long *nok, double *maxW, double *maxR) {
  if (c && c->tkind == PMCGPU_TGT_HOST) return fail(c, PMCGPU_ERR_USAGE, "host target: use pmcgpu_weights_host_post");
  return weights_common(c, 0, nullptr, nullptr, t_temper, first_index, nok, maxW, maxR);
}
int pmcgpu_weights_host_post(pmcgpu_ctx *c, const double *post, const short *ok, double t_temper, long first_index,
                             long *nok, double *maxW, double *maxR) {
  if (c && c->tkind == PMCGPU_TGT_NONE) c->tkind = PMCGPU_TGT_HOST;
  return weights_common(c, 1, post, ok, t_temper, first_index, nok, maxW, maxR);
}

static int draw_weights_core(pmcgpu_ctx *c, uint64_t seed, uint32_t iter, long first_index, double t_temper, long *nok,
                             double *maxW, double *maxR, long *n_reject) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  cudaSetDevice(c->device);
  if (c->tkind == PMCGPU_TGT_HOST) return fail(c, PMCGPU_ERR_USAGE, "host target: use pmcgpu_draw then pmcgpu_weights_host_post");
  if (t_temper < 0) return fail(c, -6019, "t_iter_tempered=%g should be > 0", t_temper);
  rc = ensure_cwght_normalised(c);
  if (rc) return rc;
  const Iter3Plan p3 = (c->nranks <= 1) ? plan3_for(c) : Iter3Plan{};
  if (p3.valid) {
    Iter3Out o3;
    rc = run_iter3(c, p3, 1, 1, 0, seed, iter, first_index, c->N, t_temper, nullptr, nullptr, o3);
    if (rc) return rc;
    RC(sink_copy(c, SINK_FINAL));
    if (n_reject) *n_reject = (long)o3.scal[I3_NREJ];
    if (o3.scal[I3_NFAIL] > 0 || o3.scal[I3_NREJ] > 5.0 * (double)c->N)
      return fail(c, kErrTooManySteps, "Too many points (%ld) outside of box", (long)o3.scal[I3_NREJ]);
    if (nok) *nok = (long)o3.scal[I3_NOK_W];
    if (maxW) *maxW = o3.scal[I3_MAXW];
    if (maxR) *maxR = o3.scal[I3_MAXR];
    return 0;
  }
  const FusedPlan p = plan_for(c);
  WeightOut o;
  if (p.valid) {
    rc = run_fused(c, p, 1, 0, seed, iter, first_index, t_temper, nullptr, nullptr, o);
    if (rc) return rc;
  } else {
    long nrej = 0;
    rc = pmcgpu_draw(c, seed, iter, first_index, &nrej);
    if (rc) return rc;
    RC(sink_copy(c, SINK_DRAWN));
    rc = run_weights_generic(c, 0, nullptr, nullptr, t_temper, first_index, o);
    if (rc) return rc;
    RC(sink_copy(c, SINK_RHO));
    o.counters[CNT_REJECT] = (unsigned long long)nrej;
    o.counters[CNT_DRAWFAIL] = 0;
  }
  RC(sink_copy(c, SINK_FINAL));  // log-weights and flags (pmc_simu_importance leaves the sample un-normalised)
  if (n_reject) *n_reject = (long)o.counters[CNT_REJECT];
  if (o.counters[CNT_DRAWFAIL] || (long)o.counters[CNT_REJECT] > 5 * c->N)
    return fail(c, kErrTooManySteps, "Too many points (%llu) outside of box", o.counters[CNT_REJECT]);
  if (nok) *nok = (long)o.counters[CNT_NOK];
  double mw = o.M_all;
  if (o.counters[CNT_POSINF]) mw = INFINITY;
  if (maxW) *maxW = fold_max(mw, o.has_first, o.first[3] != 0.0, o.first[2]);
  if (maxR) *maxR = fold_max(o.R_all, o.has_first, o.first[1] != 0.0, o.first[0]);
  return 0;
}

int pmcgpu_draw_weights(pmcgpu_ctx *c, uint64_t seed, uint32_t iter, long first_index, double t_temper, long *nok,
                        double *maxW, double *maxR, long *n_reject) {
  if (!c) return PMCGPU_ERR_USAGE;
  c->sink_active = c->sink.on;  // with a host sink attached the arrays travel while the weight kernels run
  c->sink_bytes = 0.0;
  const int rc = draw_weights_core(c, seed, iter, first_index, t_temper, nok, maxW, maxR, n_reject);
  if (c->sink_active) {
    c->sink_active = false;
    cudaSetDevice(c->device);
    const bool copy_ok = !c->copy_stream || cudaStreamSynchronize(c->copy_stream) == cudaSuccess;
    sink_join(c);
    if (!copy_ok && rc == 0)
      return fail(c, PMCGPU_ERR_CUDA, "copy to the host sink failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  return rc;
}

int pmcgpu_proposal_log_pdf(pmcgpu_ctx *c, long n, const double *x, double *out) {
  if (!c || c->prop.K == 0 || n < 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  if (c->prop.n_alive() == 0) return fail(c, kErrMvOutOfBound, "Indice out of bonds");
  const int d = c->prop.d, K = c->prop.K;
  CU(c->tmpx.ensure((size_t)n * d * sizeof(double) + 16));
  CU(c->tmpo.ensure((size_t)n * sizeof(double) + 16));
  CU(c->ld.ensure((size_t)n * K * sizeof(double) + 16));
  CU(cudaMemcpyAsync(c->tmpx.p, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_mix_log_pdf(c->dprop.view, n, c->tmpx.as<double>(), c->tmpo.as<double>(), c->ld.as<double>(), K, c->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->tmpo.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int pmcgpu_target_log_pdf(pmcgpu_ctx *c, long n, const double *x, double *out) {
  if (!c || n < 0 || c->tkind == PMCGPU_TGT_NONE || c->tkind == PMCGPU_TGT_HOST) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  const int dfree = c->nfull ? 0 : c->td;
  int d = dfree;
  if (c->nfull) { d = 0; for (int i = 0; i < c->nfull; ++i) d += c->is_def[i] ? 0 : 1; }
  const int K = c->tmix.K > 0 ? c->tmix.K : 1;
  CU(c->tmpx.ensure((size_t)n * d * sizeof(double) + 16));
  CU(c->tmpo.ensure((size_t)n * sizeof(double) + 16));
  if (c->tkind == PMCGPU_TGT_COMBINE) {
    if (c->nfull) return fail(c, PMCGPU_ERR_USAGE, "combined target with conditional parameters is not supported on the device");
    CU(cudaMemcpyAsync(c->tmpx.p, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RC(eval_combine(c, n, c->tmpx.as<double>(), c->tmpo.as<double>()));
    CU(cudaMemcpyAsync(out, c->tmpo.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
  }
  CU(c->ld.ensure((size_t)n * K * sizeof(double) + 16));
  CU(cudaMemcpyAsync(c->tmpx.p, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_target_log_pdf(target_view(c), d, n, c->tmpx.as<double>(), c->tmpo.as<double>(), c->ld.as<double>(), K, c->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->tmpo.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// scalar_product (mvdens.c:319-349) of the mvdens / first mixture component installed as the target, on n points
int pmcgpu_target_scalar_product(pmcgpu_ctx *c, long n, const double *x, double *out) {
  if (!c || n < 0 || (c->tkind != PMCGPU_TGT_MVDENS && c->tkind != PMCGPU_TGT_MIX) || c->tmix.K < 1) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  const int d = c->tmix.d;
  CU(c->tmpx.ensure((size_t)n * d * sizeof(double) + 16));
  CU(c->tmpo.ensure((size_t)n * sizeof(double) + 16));
  CU(cudaMemcpyAsync(c->tmpx.p, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_scalar_product(c->dtmix.view, 0, n, c->tmpx.as<double>(), c->tmpo.as<double>(), c->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->tmpo.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int pmcgpu_normalize(pmcgpu_ctx *c, double maxW, double *logSum, long *count) {
  if (!c || c->N <= 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  double S = 0.0;
  long cnt = 0;
  int rc = normalize_impl(c, maxW, &S, &cnt, false);
  if (rc) return rc;
  if (count) *count = cnt;
  if (cnt == 0) return fail(c, kErrNoSample, "Not a single point in sample with flag=ok.");
  if (S <= 0) return fail(c, kErrNegWeight, "Sum of weights <= 0 (got %g)", S);
  launch_norm_scale(c->N, S, c->W.as<double>(), c->sm_count * 8, c->stream);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  if (logSum) *logSum = std::log(S) + maxW - kDynMax;
  return 0;
}

int pmcgpu_perplexity_ess(pmcgpu_ctx *c, double *perp, double *ess, long *n_flagged_out) {
  if (!c || c->N <= 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  double H = 0, Q = 0;
  long nexit = 0;
  int rc = scale_perp_impl(c, 1.0, false, &H, &Q, &nexit, false);
  if (rc) return rc;
  if (n_flagged_out) *n_flagged_out = nexit;
  if (nexit == c->N) return fail(c, kErrUndef, "all points are flagged");
  if (ess) *ess = 1 / Q;
  if (perp) *perp = std::exp(-H) / (double)(c->N - nexit);
  return 0;
}

int pmcgpu_filter_histogram(pmcgpu_ctx *c, int nbins, int clipleft, int clipright, long first_index, long *n_filtered) {
  if (!c || c->N <= 0 || nbins <= 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  return filter_histogram_impl(c, nbins, clipleft, clipright, first_index, true, n_filtered);
}
int pmcgpu_set_filter_histogram(pmcgpu_ctx *c, int nbins, int clipleft, int clipright) {
  if (!c) return PMCGPU_ERR_USAGE;
  c->filt_nbins = nbins > 0 ? nbins : 0;
  c->filt_cl = clipleft;
  c->filt_cr = clipright;
  return 0;
}
int pmcgpu_weighted_mean(pmcgpu_ctx *c, int a, double *mean) {
  if (!c || c->N <= 0 || a < 0 || a >= c->d || !mean) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  const int nb = c->sm_count * 8;
  CU(c->blkmax.ensure((size_t)2 * nb * sizeof(double)));
  launch_weighted_coord(c->N, c->d, a, c->X.as<double>(), c->W.as<double>(), c->Flg.as<short>(), c->blkmax.as<double>(), nb, c->stream);
  CU(cudaGetLastError());
  RC(stage(c, nb));
  CU(cudaMemcpyAsync(c->hstage, c->blkmax.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  double s = 0.0;
  for (int b = 0; b < nb; ++b) s += c->hstage[b];
  int rc = allreduce_host(c, &s, 1, 0);
  if (rc) return rc;
  *mean = s;
  return 0;
}

// ---- resampling of the weighted sample (pmc.c:1585-1626) ----------------------------------------------------------------
static int resample_common(pmcgpu_ctx *c, uint64_t seed, uint32_t stream, int mode, void *out_host) {
  if (!c || c->N <= 0 || !out_host) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  const long N = c->N;
  CU(c->cdf.ensure((size_t)N * sizeof(double)));
  CU(c->cdfsum.ensure((size_t)scan_blocks(N) * sizeof(double)));
  const size_t obytes = (size_t)N * (mode == 0 ? sizeof(unsigned long long) : sizeof(unsigned int));
  CU(c->rsout.ensure(obytes));
  if (mode == 1) CU(cudaMemsetAsync(c->rsout.p, 0, obytes, c->stream));
  launch_weight_cdf(N, c->W.as<double>(), mode, c->cdf.as<double>(), c->cdfsum.as<double>(), c->stream);
  CU(cudaGetLastError());
  RC(stage(c, 1));
  CU(cudaMemcpyAsync(c->hstage, c->cdf.as<double>() + (N - 1), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (!(c->hstage[0] > 0.0)) {
    if (mode == 1) { memset(out_host, 0, obytes); return 0; }  // no residual mass: every multiplicity is zero
    return fail(c, kErrNegWeight, "Sum of weights <= 0 (got %g): nothing to sample from", c->hstage[0]);
  }
  launch_resample(N, N, c->cdf.as<double>(), seed, stream, mode, c->rsout.as<unsigned long long>(), c->rsout.as<unsigned int>(), c->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out_host, c->rsout.p, obytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
int pmcgpu_sample_indices(pmcgpu_ctx *c, uint64_t seed, uint32_t stream, size_t *isample) {
  return resample_common(c, seed, stream, 0, isample);
}
int pmcgpu_resample_residual(pmcgpu_ctx *c, uint64_t seed, uint32_t stream, unsigned int *multiplicity) {
  return resample_common(c, seed, stream, 1, multiplicity);
}

// ---- clip_weights (pmc.c:1134-1188) ------------------------------------------------------------------------------------
// The reference keeps the n largest candidates in a small sorted array; its initialisation loop (pmc.c:1144-1151) fills
// that array with m copies of weights[0] (m = number of leading samples i < n with flg[i] != 0) and zeros, and samples
// j < n are never offered.  The threshold is therefore the n-th largest of  {w0 x m} U {0 x (n-m)} U {w_j : j >= n, ok},
// found here by a bitwise search over the order-preserving integer keys (64 counting passes).
static unsigned long long host_dkey(double v) {
  unsigned long long b;
  memcpy(&b, &v, 8);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
int pmcgpu_clip_weights(pmcgpu_ctx *c, int n, double *threshold, long *n_clipped, double *sum_after) {
  if (!c || c->N <= 0 || n < 0) return PMCGPU_ERR_USAGE;
  if (n == 0) { if (n_clipped) *n_clipped = 0; return 0; }
  if (n > c->N) return fail(c, PMCGPU_ERR_USAGE, "clip_weights: n=%d exceeds the sample size %ld", n, c->N);
  cudaSetDevice(c->device);
  const long N = c->N;
  const int nb = c->sm_count * 8;
  RC(stage(c, (size_t)n / 4 + 8 + nb));
  short *hf = reinterpret_cast<short *>(c->hstage + 2);
  CU(cudaMemcpyAsync(c->hstage, c->W.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(hf, c->Flg.p, (size_t)n * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const double w0 = c->hstage[0];
  int m = 0;
  while (m < n && hf[m] != 0) ++m;
  // extras of the reference's initial array, as keys
  const unsigned long long k_w0 = host_dkey(w0), k_zero = host_dkey(0.0);
  auto extras_ge = [&](unsigned long long key) { return (unsigned long long)((k_w0 >= key ? m : 0) + (k_zero >= key ? n - m : 0)); };
  int rc = zero_small(c);
  if (rc) return rc;
  unsigned long long *cnt = c->small.as<unsigned long long>() + SM_COUNTERS;
  unsigned long long key = 0;
  for (int b = 63; b >= 0; --b) {
    const unsigned long long trial = key | (1ull << b);
    CU(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
    launch_count_key_ge(N, n, c->W.as<double>(), c->Flg.as<short>(), trial, cnt, nb, c->stream);
    CU(cudaGetLastError());
    unsigned long long hc = 0;
    CU(cudaMemcpyAsync(&hc, cnt, sizeof hc, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (hc + extras_ge(trial) >= (unsigned long long)n) key = trial;
  }
  unsigned long long bits = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
  double T;
  memcpy(&T, &bits, 8);
  CU(c->blkmax.ensure((size_t)2 * nb * sizeof(double)));
  CU(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
  launch_clip_apply(N, T, c->W.as<double>(), c->Flg.as<short>(), c->blkmax.as<double>(), cnt, nb, c->stream);
  CU(cudaGetLastError());
  unsigned long long hclip = 0;
  CU(cudaMemcpyAsync(c->hstage, c->blkmax.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&hclip, cnt, sizeof hclip, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  double S = 0.0;
  for (int b = 0; b < nb; ++b) S += c->hstage[b];
  if (threshold) *threshold = T;
  if (n_clipped) *n_clipped = (long)hclip;
  if (sum_after) *sum_after = S;
  if (S <= 0) return fail(c, -6015, "Sum of weights (%g) has to be larger than zero", S);  // pmc_infnan, pmc.c:1180
  launch_scale_ok(N, S, c->W.as<double>(), c->Flg.as<short>(), nb, c->stream);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int pmcgpu_count_indices(pmcgpu_ctx *c, size_t *count) {
  if (!c || !count || c->prop.K == 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  std::vector<size_t> cnt;
  int rc = count_indices_impl(c, cnt, false);
  if (rc) return rc;
  memcpy(count, cnt.data(), cnt.size() * sizeof(size_t));
  return 0;
}

static int update_common(pmcgpu_ctx *c, double maxR, long N_total, int hard, int *retry, double *o_wght,
                         double *o_mean, double *o_L, double *o_detL, int *o_alive) {
  int rc = check_ready(c, false);
  if (rc) return rc;
  cudaSetDevice(c->device);
  if (!retry) return PMCGPU_ERR_USAGE;
  std::vector<size_t> count;
  rc = count_indices_impl(c, count, true);
  if (rc) return rc;
  rc = update_two_pass(c, maxR, N_total > 0 ? N_total : c->N, hard, retry, count);
  if (rc) return rc;
  sanitize_dead(c->prop);
  rc = install_proposal(c);
  if (rc) return rc;
  export_proposal(c, o_wght, o_mean, o_L, o_detL, o_alive);
  return 0;
}

int pmcgpu_update_rb(pmcgpu_ctx *c, double maxR, long N_total, int *retry, double *o_wght, double *o_mean, double *o_L,
                     double *o_detL, int *o_alive) {
  return update_common(c, maxR, N_total, 0, retry, o_wght, o_mean, o_L, o_detL, o_alive);
}
int pmcgpu_update_hard(pmcgpu_ctx *c, long N_total, int *retry, double *o_wght, double *o_mean, double *o_L, double *o_detL,
                       int *o_alive) {
  if (c && c->prop.student) return fail(c, PMCGPU_ERR_USAGE, "update_prop with Student-t components yields NaN in the reference (pmc.c:1241-1258); not supported");
  return update_common(c, 0.0, N_total, 1, retry, o_wght, o_mean, o_L, o_detL, o_alive);
}

int pmcgpu_ranindx(pmcgpu_ctx *c, long n, const double *z, size_t *idx) {
  if (!c || c->prop.K == 0 || n < 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  CU(c->tmpx.ensure((size_t)n * sizeof(double) + 16));
  CU(c->tmpo.ensure((size_t)n * sizeof(size_t) + 16));
  CU(cudaMemcpyAsync(c->tmpx.p, z, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_ranindx(c->dprop.view.cw, c->prop.K, n, c->tmpx.as<double>(), c->tmpo.as<unsigned long long>(), c->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(idx, c->tmpo.p, (size_t)n * sizeof(size_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int pmcgpu_isinbox(pmcgpu_ctx *c, long n, const double *x, int *inside) {
  if (!c || n < 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  if (c->nfaces == 0) { for (long i = 0; i < n; ++i) inside[i] = 1; return 0; }
  const int d = c->box_d;
  CU(c->tmpx.ensure((size_t)n * d * sizeof(double) + 16));
  CU(c->tmpo.ensure((size_t)n * sizeof(int) + 16));
  CU(cudaMemcpyAsync(c->tmpx.p, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_isinbox(BoxDev{c->nfaces, d, c->dfaces.as<double>()}, n, c->tmpx.as<double>(), c->tmpo.as<int>(), c->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(inside, c->tmpo.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- the iteration ------------------------------------------------------------------------------------------------
static int step_core(pmcgpu_ctx *c, int do_draw, uint64_t seed, uint32_t iter, long first_index, long N_total,
                     double t_temper, int do_update, int *retry, pmcgpu_step_result *res, double *o_wght, double *o_mean,
                     double *o_L, double *o_detL, int *o_alive, const double *post_host = nullptr, const short *ok_host = nullptr) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  cudaSetDevice(c->device);
  // PMCB200_TRACE=1: wall-clock of the phases of one iteration on stderr (host finalisation included)
  static const bool trace = getenv("PMCB200_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[pmcb200] %-22s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  if (c->tkind == PMCGPU_TGT_HOST && !post_host)
    return fail(c, PMCGPU_ERR_USAGE, "host target: draw, download X, evaluate, then pmcgpu_step_given_post (or drive the phases separately)");
  if (post_host && do_draw) return fail(c, PMCGPU_ERR_USAGE, "target values supplied for samples that are still to be drawn");
  if (N_total <= 0) N_total = c->N;
  const double *dpost = nullptr;
  const short *dok = nullptr;
  if (post_host) {  // target values of the store's samples, evaluated by the caller (opaque likelihood callbacks)
    CU(c->hostpost.ensure((size_t)c->N * sizeof(double)));
    CU(cudaMemcpyAsync(c->hostpost.p, post_host, (size_t)c->N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dpost = c->hostpost.as<double>();
    if (ok_host) {
      CU(c->hostok.ensure((size_t)c->N * sizeof(short)));
      CU(cudaMemcpyAsync(c->hostok.p, ok_host, (size_t)c->N * sizeof(short), cudaMemcpyHostToDevice, c->stream));
      dok = c->hostok.as<short>();
    }
  }
  int retry_local = retry ? *retry : 0;
  pmcgpu_step_result r{};
  if (do_draw) { rc = ensure_cwght_normalised(c); if (rc) return rc; }
  const HostMix &m0 = c->prop;
  const int K = m0.K, d = m0.d;
  // third generation first: one host synchronisation per iteration, normalisation fused into the statistics kernel
  const Iter3Plan p3 = (c->filt_nbins == 0) ? plan3_for(c) : Iter3Plan{};  // a filter changes the flags between weights and statistics
  Iter3Out o3;
  const bool have3 = p3.valid != 0;
  const bool stats3 = have3 && do_update && !c->force_precise;
  const FusedPlan p = have3 ? FusedPlan{} : plan_for(c);
  WeightOut o;
  const bool fused_stats = p.valid && do_update && !c->force_precise && c->filt_nbins == 0;
  if (have3) {
    rc = run_iter3(c, p3, do_draw, 2, stats3 ? 1 : 0, seed, iter, first_index, N_total, t_temper, dpost, dok, o3);
    if (rc) return rc;
  } else if (p.valid) {
    rc = run_fused(c, p, do_draw, fused_stats ? 1 : 0, seed, iter, first_index, t_temper, dpost, dok, o);
    if (rc) return rc;
  } else {
    if (do_draw) {
      long nrej = 0, nfail = 0;
      rc = draw_impl(c, seed, iter, first_index, &nrej, &nfail, c->nranks <= 1);  // ranks: the reduced counters decide below
      if (rc) return rc;
      o.counters[CNT_REJECT] = (unsigned long long)nrej;
      o.counters[CNT_DRAWFAIL] = (unsigned long long)nfail;
      RC(sink_copy(c, SINK_DRAWN));
    }
    const unsigned long long keep_rej = o.counters[CNT_REJECT], keep_fail = o.counters[CNT_DRAWFAIL];
    c->mma_store_ld = do_update ? 1 : 0;  // importance only: nothing reads the component log-densities afterwards
    rc = run_weights_generic(c, dpost ? 1 : 0, dpost, dok, t_temper, first_index, o);
    c->mma_store_ld = 1;
    if (rc) return rc;
    RC(sink_copy(c, SINK_RHO));
    o.counters[CNT_REJECT] = keep_rej;
    o.counters[CNT_DRAWFAIL] = keep_fail;
  }
  tick("draw+weights");
  // ---- global maxima (pmc_mpi.c:502-508 keeps the max over ranks) with the i==0 quirk carried as a flag
  double mx[6] = {0, 0, 0, 0, 0, 0};
  double maxW = 0.0, maxR = 0.0;
  const double M_local = o.M_all;
  if (have3) {
    maxW = o3.scal[I3_MAXW];
    maxR = o3.scal[I3_MAXR];
    r.maxW = maxW;
    r.maxR = maxR;
    r.n_reject = (long)o3.scal[I3_NREJ];
    if (do_draw && (o3.scal[I3_NFAIL] > 0 || o3.scal[I3_NREJ] > 5.0 * (double)N_total))
      return fail(c, kErrTooManySteps, "Too many points (%ld) outside of box", r.n_reject);
  } else {
  mx[0] = o.counters[CNT_POSINF] ? INFINITY : o.M_all;
  mx[1] = o.R_all;
  mx[2] = (o.has_first && o.first[3] != 0.0 && std::isnan(o.first[2])) ? 1.0 : 0.0;  // maxW poisoned by NaN at i==0
  mx[3] = (o.has_first && o.first[1] != 0.0 && std::isnan(o.first[0])) ? 1.0 : 0.0;
  mx[4] = (o.has_first && o.first[3] != 0.0) ? 1.0 : 0.0;  // sample 0 reached the weight assignment
  mx[5] = (o.has_first && o.first[1] != 0.0) ? 1.0 : 0.0;
  rc = allreduce_host(c, mx, 6, 1);
  if (rc) return rc;
  maxW = mx[2] != 0.0 ? NAN : (mx[4] != 0.0 ? mx[0] : (mx[0] > kNegInit ? mx[0] : kNegInit));
  maxR = mx[3] != 0.0 ? NAN : (mx[5] != 0.0 ? mx[1] : (mx[1] > kNegInit ? mx[1] : kNegInit));
  r.maxW = maxW;
  r.maxR = maxR;
  {
    double cnt[2] = {(double)o.counters[CNT_REJECT], (double)o.counters[CNT_DRAWFAIL]};
    rc = allreduce_host(c, cnt, 2, 0);
    if (rc) return rc;
    r.n_reject = (long)cnt[0];
    if (do_draw && (cnt[1] > 0 || cnt[0] > 5.0 * (double)N_total))
      return fail(c, kErrTooManySteps, "Too many points (%ld) outside of box", r.n_reject);
  }
  }
  // ---- pmc_filter hook (pmc.c:203-206): filter_histogram on the log-weights
  if (c->filt_nbins > 0) {
    long nf = 0;
    rc = filter_histogram_impl(c, c->filt_nbins, c->filt_cl, c->filt_cr, first_index, true, &nf);
    if (rc) return rc;
  }
  // ---- normalise (pmc.c:641-697), literally: S = sum exp(lw - maxW + 500)
  double S = 0.0;
  long cnt_ok = 0;
  double H = 0, Q = 0;
  long nexit = 0;
  const int nstat3 = stats3 ? p3.K * p3.E : 0;
  if (have3 && o3.clean) {  // the statistics kernel normalised the weights and summed w log w, w^2 and the counts
    S = o3.scal[I3_SREF];
    cnt_ok = (long)o3.out[nstat3 + 2];
    H = o3.out[nstat3 + 0];
    Q = o3.out[nstat3 + 1];
    nexit = (long)o3.out[nstat3 + 3];
  } else {
    rc = normalize_impl(c, maxW, &S, &cnt_ok, true);
    if (rc) return rc;
  }
  if (cnt_ok == 0) return fail(c, kErrNoSample, "Not a single point in sample with flag=ok.");
  if (!(S > 0)) return fail(c, kErrNegWeight, "Sum of weights <= 0 (got %g)", S);
  r.logSum = std::log(S) + maxW - kDynMax;
  r.nok = cnt_ok;
  if (!(have3 && o3.clean)) {
    rc = scale_perp_impl(c, S, true, &H, &Q, &nexit, true);
    if (rc) return rc;
  }
  if (nexit == N_total) return fail(c, kErrUndef, "all points are flagged");
  r.ess = 1 / Q;
  r.perplexity = std::exp(-H) / (double)(N_total - nexit);
  r.ln_evidence = r.logSum - std::log((double)(N_total - nexit));
  if (!(have3 && o3.clean)) RC(sink_copy(c, SINK_FINAL));  // weights and flags are final; they travel while the update runs
  tick("normalise+perplexity");
  // ---- update
  if (do_update) {
    // a NaN maximum (the i==0 quirk) poisons the reference's responsibilities: only the literal path reproduces that
    bool mma_stats = !p.valid && !have3 && !c->force_precise && c->q_resident && mma_usable(c, 0);
    if (!p.valid && !have3 && !c->force_precise) {
      // q_resident depends on each rank's free memory and the one-pass statistics need a centred copy of the shard: the
      // ranks agree on the path BEFORE any statistics collective (all take the literal update if one of them cannot),
      // otherwise they would issue different all-reduce sequences
      if (mma_stats && c->prop.student && K > 256) mma_stats = false;
      if (mma_stats && c->xc.ensure((size_t)c->N * d * sizeof(double)) != cudaSuccess) { cudaGetLastError(); mma_stats = false; }
      if (c->nranks > 1) {
        double veto = mma_stats ? 0.0 : 1.0;
        rc = allreduce_host(c, &veto, 1, 1);
        if (rc) return rc;
        if (veto != 0.0) mma_stats = false;
      }
    }
    const bool use_stats3 = stats3 && o3.clean;
    bool precise = !(fused_stats || mma_stats || use_stats3) || std::isnan(maxR) || std::isnan(maxW);
    std::vector<double> &W = c->hW, &s1 = c->hs1, &S2 = c->hS2;
    StatView sv;
    std::vector<double> pivot(d);
    std::vector<size_t> count(K);
    if (!precise) {
      {  // pivot = mixture mean (weights of live components)
        double sw = 0.0;
        for (int k = 0; k < K; ++k) if (m0.wght[k] > 0) sw += m0.wght[k];
        for (int j = 0; j < d; ++j) {
          double s = 0.0;
          for (int k = 0; k < K; ++k) if (m0.wght[k] > 0) s += m0.wght[k] * m0.mean[(size_t)k * d + j];
          pivot[j] = s / sw;
        }
      }
      if (fused_stats || use_stats3) {
        const int E = 1 + d + d * (d + 1) / 2;
        std::vector<double> st;
        if (use_stats3) {  // already reduced over ranks on the device
          st.assign(o3.out.begin(), o3.out.begin() + (size_t)K * E);
          for (int k = 0; k < K; ++k) count[k] = (size_t)o3.out[nstat3 + 4 + k];
        } else {
        // reduce the statistics over ranks: rescale to the global maximum, then sum
        st = o.stats;
        const double Mg = mx[0];
        const double sc = (M_local == -INFINITY) ? 0.0 : std::exp(M_local - Mg);
        if (c->nranks > 1) for (double &v : st) v *= sc;
        std::vector<double> cntk(K);
        for (int k = 0; k < K; ++k) cntk[k] = (double)o.count_k[k];
        rc = allreduce_host(c, st.data(), (long)st.size(), 0);
        if (rc) return rc;
        rc = allreduce_host(c, cntk.data(), K, 0);
        if (rc) return rc;
        for (int k = 0; k < K; ++k) count[k] = (size_t)cntk[k];
        }
        // unpack canonical statistics -> W (= G for Gaussian components), s1, S2 about the pivot
        W.resize(K);
        s1.resize((size_t)K * d);
        S2.resize((size_t)K * d * d);
        for (int k = 0; k < K; ++k) {
          const double *sk = st.data() + (size_t)k * E;
          W[k] = sk[0];
          for (int a = 0; a < d; ++a) s1[(size_t)k * d + a] = sk[1 + a];
          int e = 1 + d;
          for (int a = 0; a < d; ++a)
            for (int b = a; b < d; ++b) { S2[(size_t)k * d * d + (size_t)a * d + b] = sk[e]; S2[(size_t)k * d * d + (size_t)b * d + a] = sk[e]; ++e; }
        }
        sv.W = sv.G = W.data();
        sv.s1 = s1.data();
        sv.ss1 = (size_t)d;
        sv.S2 = S2.data();
        sv.sS2 = (size_t)d * d;
      } else {
        rc = count_indices_impl(c, count, true);
        if (rc) return rc;
        rc = stats_mma(c, count, pivot, W, sv);
        if (rc == 1) precise = true;
        else if (rc) return rc;
      }
      tick("statistics");
    }
    if (!precise) {
      // the trial proposal is written next to the current one (whose factors double as the snapshot that
      // cleanup_after_update restores on a first failed factorisation): no 8 MB copies at d = 128, K = 64
      HostMix &trial = c->trial;
      trial.K = K;
      trial.d = d;
      trial.df = c->prop.df;
      trial.band = c->prop.band;
      trial.student = c->prop.student;
      trial.detL = c->prop.detL;
      trial.wght.resize(K);
      trial.mean.resize((size_t)K * d);
      trial.L.resize((size_t)K * d * d);
      int retry_trial = retry_local;
      std::vector<int> alive(K);
      rc = finalize_update_strided(K, d, sv.W, sv.sW, sv.G, sv.sG, sv.s1, sv.ss1, sv.S2, sv.sS2, pivot.data(), 0, count.data(),
                                   N_total, trial.band.empty() ? nullptr : trial.band.data(), &retry_trial, c->prop.L.data(),
                                   trial.wght.data(), trial.mean.data(), trial.L.data(), trial.detL.data(), alive.data());
      // pivot guard: the one-pass covariance loses ~eps*(1 + r) relative accuracy, r = (mu'-pivot)^2 / Sigma'_aa
      std::vector<double> worst_k(K, 0.0);
      if (rc == 0)
        parallel_components(K, (double)d * d * 4.0, [&](int k) {
          if (!(trial.wght[k] > 0.0)) return;
          double worst = 0.0;
          for (int a = 0; a < d; ++a) {
            double var = 0.0;  // Sigma'_aa = sum_j L_aj^2
            for (int j = 0; j <= a; ++j) { const double l = trial.L[(size_t)k * d * d + (size_t)a * d + j]; var += l * l; }
            const double dm = trial.mean[(size_t)k * d + a] - pivot[a];
            const double ratio = dm * dm / var;
            if (!(ratio <= worst)) worst = ratio;
          }
          worst_k[k] = worst;
        });
      double worst = 0.0;
      for (int k = 0; k < K; ++k) if (!(worst_k[k] <= worst)) worst = worst_k[k];
      tick("finalise+guard");
      if (rc != 0 || retry_trial != 0 || !(worst <= c->guard_ratio)) {
        precise = true;  // anything unusual (failed factorisation, ill-conditioned pivot) goes through the literal path
      } else {
        std::swap(c->prop, trial);
        retry_local = retry_trial;
      }
    }
    if (precise) {
      std::vector<size_t> count;
      rc = count_indices_impl(c, count, true);
      if (rc) return rc;
      rc = update_two_pass(c, maxR, N_total, 0, &retry_local, count);
      if (rc) return rc;
      r.used_precise_path = 1;
    }
    sanitize_dead(c->prop);
    rc = install_proposal(c);
    if (rc) return rc;
    tick("install proposal");
  }
  r.retry = retry_local;
  r.n_alive = c->prop.n_alive();
  if (retry) *retry = retry_local;
  if (res) *res = r;
  export_proposal(c, o_wght, o_mean, o_L, o_detL, o_alive);
  tick("export");
  return 0;
}

// the iteration with the host sink armed: whatever happens, no copy into the caller's arrays is in flight on return
static int step_impl(pmcgpu_ctx *c, int do_draw, uint64_t seed, uint32_t iter, long first_index, long N_total,
                     double t_temper, int do_update, int *retry, pmcgpu_step_result *res, double *o_wght, double *o_mean,
                     double *o_L, double *o_detL, int *o_alive, const double *post_host = nullptr, const short *ok_host = nullptr) {
  if (!c) return PMCGPU_ERR_USAGE;
  c->sink_active = c->sink.on && do_draw;
  c->sink_bytes = 0.0;
  const int rc = step_core(c, do_draw, seed, iter, first_index, N_total, t_temper, do_update, retry, res, o_wght, o_mean,
                           o_L, o_detL, o_alive, post_host, ok_host);
  if (c->sink_active) {
    c->sink_active = false;
    cudaSetDevice(c->device);
    const bool copy_ok = !c->copy_stream || cudaStreamSynchronize(c->copy_stream) == cudaSuccess;
    sink_join(c);
    if (!copy_ok && rc == 0)
      return fail(c, PMCGPU_ERR_CUDA, "copy to the host sink failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  return rc;
}

int pmcgpu_set_host_sink(pmcgpu_ctx *c, double *X, size_t *indices, double *weights, double *log_rho, short *flg) {
  if (!c) return PMCGPU_ERR_USAGE;
  c->sink.X = X;
  c->sink.idx = indices;
  c->sink.w = weights;
  c->sink.lr = log_rho;
  c->sink.flg = flg;
  c->sink.on = X || indices || weights || log_rho || flg;
  return 0;
}

int pmcgpu_pmc_step(pmcgpu_ctx *c, uint64_t seed, uint32_t iter, long first_index, long N_total, int do_update,
                    int *retry, pmcgpu_step_result *res, double *o_wght, double *o_mean, double *o_L, double *o_detL,
                    int *o_alive) {
  return step_impl(c, 1, seed, iter, first_index, N_total, c->t_temper, do_update, retry, res, o_wght, o_mean, o_L, o_detL, o_alive);
}
int pmcgpu_step_given(pmcgpu_ctx *c, long first_index, long N_total, double t_temper, int do_update, int *retry,
                      pmcgpu_step_result *res, double *o_wght, double *o_mean, double *o_L, double *o_detL, int *o_alive) {
  return step_impl(c, 0, 0, 0, first_index, N_total, t_temper, do_update, retry, res, o_wght, o_mean, o_L, o_detL, o_alive);
}

int pmcgpu_step_given_post(pmcgpu_ctx *c, const double *post, const short *ok, long first_index, long N_total, double t_temper,
                           int do_update, int *retry, pmcgpu_step_result *res, double *o_wght, double *o_mean, double *o_L,
                           double *o_detL, int *o_alive) {
  if (!c || !post) return PMCGPU_ERR_USAGE;
  const int saved_kind = c->tkind;
  c->tkind = PMCGPU_TGT_HOST;
  const int rc = step_impl(c, 0, 0, 0, first_index, N_total, t_temper, do_update, retry, res, o_wght, o_mean, o_L, o_detL, o_alive, post, ok);
  c->tkind = saved_kind == PMCGPU_TGT_NONE ? PMCGPU_TGT_HOST : saved_kind;
  return rc;
}

int pmcgpu_get_proposal(pmcgpu_ctx *c, double *o_wght, double *o_mean, double *o_L, double *o_detL, int *o_alive) {
  if (!c || c->prop.K == 0) return PMCGPU_ERR_USAGE;
  export_proposal(c, o_wght, o_mean, o_L, o_detL, o_alive);
  return 0;
}

// page-lock / unlock a caller-owned host block so that upload/download run as DMA at PCIe speed
int pmcgpu_pin_host(void *p, size_t bytes) {
  if (!p || !bytes) return PMCGPU_ERR_USAGE;
  const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return PMCGPU_ERR_CUDA; }
  return 0;
}
int pmcgpu_unpin_host(void *p) {
  if (!p) return PMCGPU_ERR_USAGE;
  const cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { cudaGetLastError(); return PMCGPU_ERR_CUDA; }
  return 0;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------------
void pmcgpu_shard_range(long N_total, int rank, int nranks, long *first, long *len) {
  const long base = N_total / nranks, rem = N_total % nranks;
  const long l = base + (rank < rem ? 1 : 0);
  const long f = rank * base + (rank < rem ? rank : rem);
  if (first) *first = f;
  if (len) *len = l;
}

// Every rank contributes the bytes [my_off, my_off + my_len) of a buffer of `total` bytes (ranges of different ranks are
// disjoint); afterwards host_out (where not null) holds the union.  Chunks travel through a zeroed device staging buffer
// and a sum all-reduce of 64-bit words: with disjoint non-zero bytes the sum is the union, exact for any bit pattern.
static int share_bytes(pmcgpu_ctx *c, const void *src, bool src_on_device, size_t total, size_t my_off, size_t my_len, void *host_out) {
  if (total == 0) return 0;
  if (c->nranks <= 1) {
    if (host_out && my_len) {
      CU(cudaMemcpyAsync((char *)host_out + my_off, src, my_len, src_on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
  }
  const size_t CH = (size_t)64 << 20;
  const size_t chunk = total < CH ? ((total + 7) / 8) * 8 : CH;
  CU(c->dshare.ensure(chunk));
  for (size_t o = 0; o < total; o += chunk) {
    const size_t len = (total - o) < chunk ? (total - o) : chunk;
    const size_t words = (len + 7) / 8;
    CU(cudaMemsetAsync(c->dshare.p, 0, words * 8, c->stream));
    const size_t lo = my_off > o ? my_off : o, hi = (my_off + my_len) < (o + len) ? (my_off + my_len) : (o + len);
    if (hi > lo)
      CU(cudaMemcpyAsync((char *)c->dshare.p + (lo - o), (const char *)src + (lo - my_off), hi - lo,
                         src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    if (!c->cb) {
      if (!c->nccl_comm) return fail(c, PMCGPU_ERR_NCCL, "nranks=%d but no communicator attached", c->nranks);
      const int r = g_nccl.AllReduce(c->dshare.p, c->dshare.p, words, kNcclUint64, kNcclSum, c->nccl_comm, c->stream);
      if (r != 0) return fail(c, PMCGPU_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
      if (host_out) CU(cudaMemcpyAsync((char *)host_out + o, c->dshare.p, len, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    } else {  // callback ranks reduce doubles on the host: 16-bit pieces as exact doubles
      std::vector<unsigned short> raw(words * 4);
      CU(cudaMemcpyAsync(raw.data(), c->dshare.p, words * 8, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      std::vector<double> dv(raw.begin(), raw.end());
      if (c->cb(c->cb_user, dv.data(), (long)dv.size(), 0) != 0) return fail(c, PMCGPU_ERR_NCCL, "allreduce callback failed");
      for (size_t i = 0; i < raw.size(); ++i) raw[i] = (unsigned short)dv[i];
      if (host_out) memcpy((char *)host_out + o, raw.data(), len);
    }
  }
  return 0;
}

int pmcgpu_comm_rank(const pmcgpu_ctx *c, int *rank, int *nranks) {
  if (!c) return PMCGPU_ERR_USAGE;
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  return 0;
}
int pmcgpu_comm_allreduce(pmcgpu_ctx *c, double *buf, long n, int op) {
  if (!c || !buf || n < 0) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  return allreduce_host(c, buf, n, op);
}
int pmcgpu_comm_bcast(pmcgpu_ctx *c, void *buf, size_t bytes, int root) {
  if (!c || !buf || root < 0 || root >= (c->nranks > 1 ? c->nranks : 1)) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  if (c->nranks <= 1) return 0;
  return share_bytes(c, buf, false, bytes, 0, c->rank == root ? bytes : 0, buf);
}
int pmcgpu_comm_allgather_bytes(pmcgpu_ctx *c, void *buf, size_t total, size_t my_off, size_t my_len) {
  if (!c || !buf || my_off + my_len > total) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  if (c->nranks <= 1) return 0;
  return share_bytes(c, (const char *)buf + my_off, false, total, my_off, my_len, buf);
}
int pmcgpu_comm_gather_store(pmcgpu_ctx *c, int root, long N_total, long first, double *X, size_t *indices, double *weights,
                             double *log_rho, short *flg) {
  if (!c || N_total < 0 || first < 0 || first + c->N > N_total) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  const bool me = c->nranks <= 1 || c->rank == root;
  const size_t d = (size_t)c->d, N = (size_t)c->N, T = (size_t)N_total, f = (size_t)first;
  // every rank takes part in every collective: whether an array travels is decided by root's pointers
  double want[5] = {me && X ? 1.0 : 0.0, me && indices ? 1.0 : 0.0, me && weights ? 1.0 : 0.0, me && log_rho ? 1.0 : 0.0, me && flg ? 1.0 : 0.0};
  RC(allreduce_host(c, want, 5, 1));
  if (want[0] != 0.0) RC(share_bytes(c, c->X.p, true, T * d * 8, f * d * 8, N * d * 8, me ? X : nullptr));
  if (want[1] != 0.0) RC(share_bytes(c, c->Idx.p, true, T * 8, f * 8, N * 8, me ? indices : nullptr));
  if (want[2] != 0.0) RC(share_bytes(c, c->W.p, true, T * 8, f * 8, N * 8, me ? weights : nullptr));
  if (want[3] != 0.0) RC(share_bytes(c, c->R.p, true, T * 8, f * 8, N * 8, me ? log_rho : nullptr));
  if (want[4] != 0.0) RC(share_bytes(c, c->Flg.p, true, T * 2, f * 2, N * 2, me ? flg : nullptr));
  return 0;
}

int pmcgpu_nccl_unique_id(void *id_out) {
  std::string why;
  if (!id_out) return PMCGPU_ERR_USAGE;
  if (!load_nccl(why)) return PMCGPU_ERR_NCCL;
  return g_nccl.GetUniqueId(id_out) == 0 ? 0 : PMCGPU_ERR_NCCL;
}

int pmcgpu_comm_init(pmcgpu_ctx *c, const void *id, int rank, int nranks) {
  if (!c || !id || rank < 0 || rank >= nranks) return PMCGPU_ERR_USAGE;
  cudaSetDevice(c->device);
  std::string why;
  if (!load_nccl(why)) return fail(c, PMCGPU_ERR_NCCL, "%s", why.c_str());
  IdByValue idv;
  memcpy(idv.internal, id, PMCGPU_NCCL_ID_BYTES);
  void *comm = nullptr;
  const int r = g_nccl.CommInitRank(&comm, nranks, idv, rank);
  if (r != 0) return fail(c, PMCGPU_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  c->nccl_comm = comm;
  c->rank = rank;
  c->nranks = nranks;
  c->cb = nullptr;
  return 0;
}

int pmcgpu_comm_set_callback(pmcgpu_ctx *c, pmcgpu_allreduce_fn fn, void *user, int rank, int nranks) {
  if (!c || rank < 0 || rank >= nranks) return PMCGPU_ERR_USAGE;
  c->cb = fn;
  c->cb_user = user;
  c->rank = rank;
  c->nranks = fn ? nranks : 1;
  return 0;
}

}  // extern "C"
