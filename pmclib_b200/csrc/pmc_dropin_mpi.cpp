// pmc_dropin_mpi.cpp — libpmc_mpi's entry points (pmclib/include/pmc_mpi.h:52-101, pmclib/src/pmc_mpi.c) for one process
// per GPU.
//
// The reference parallelises ONE thing over MPI ranks: the likelihood loop.  The master draws the whole sample, scatters
// slices (send_simulation), every rank weighs its slice, the master gathers (receive_importance_weight), normalises,
// updates and broadcasts the new mixture (send_mix_mvdens).  Here every rank owns a GPU and a shard of the sample from
// the start: it draws its shard (Philox counted by the GLOBAL sample index, so the sample does not depend on the number
// of ranks), weighs it, and the ranks all-reduce the maxima, the normaliser and the K(1+d+d^2) sufficient statistics
// (NCCL over NVLink).  Every rank finalises the same reduced statistics, so all proposals are identical without a
// broadcast.  What the reference's master holds afterwards - the whole pmc_simu in host memory - is delivered to rank 0
// by pmcgpu_comm_gather_store.
//
// MPI itself is not needed (and is absent from this image): rank and size come from the launcher's environment
// (mpirun / srun / torchrun all export them) and the NCCL id travels through a rendezvous file.  An MPI host that wants
// to own the collective instead attaches pmcgpu_comm_set_callback to pmcb200_context(psim).
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/pmcb200.h"
#include "../../include/pmclib_abi.h"
#include "pmc_dropin_internal.h"

namespace {
constexpr int pmc_allocate = -6001, pmc_incompat = -6016, pmc_badComm = -6009;

int env_int(const char *const *names, int dflt) {
  for (; *names; ++names)
    if (const char *e = getenv(*names)) return atoi(e);
  return dflt;
}

struct World {
  int rank = 0, size = 1, local_rank = 0;
  std::string rdv;
};
const World &world() {
  static World w = [] {
    World x;
    static const char *const rk[] = {"PMCB200_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "PMIX_RANK", "SLURM_PROCID", "RANK", nullptr};
    static const char *const sz[] = {"PMCB200_WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", "WORLD_SIZE", nullptr};
    static const char *const lr[] = {"PMCB200_LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID", "LOCAL_RANK", nullptr};
    x.rank = env_int(rk, 0);
    x.size = env_int(sz, 1);
    x.local_rank = env_int(lr, x.rank);
    if (x.size < 1) x.size = 1;
    if (x.rank < 0 || x.rank >= x.size) { x.rank = 0; x.size = 1; }
    if (const char *e = getenv("PMCB200_RENDEZVOUS")) x.rdv = e;
    else {
      // one launcher, one job: its id (or, failing that, the launcher's pid = our parent) names the rendezvous file
      const char *job = getenv("SLURM_JOB_ID");
      if (!job) job = getenv("TORCHELASTIC_RUN_ID");
      if (!job) job = getenv("MASTER_PORT");
      char buf[256];
      if (job) snprintf(buf, sizeof buf, "/tmp/pmcb200_rdv_%d_%s", (int)getuid(), job);
      else snprintf(buf, sizeof buf, "/tmp/pmcb200_rdv_%d_p%d", (int)getuid(), (int)getppid());
      x.rdv = buf;
    }
    return x;
  }();
  return w;
}

// rank 0 publishes `bytes` bytes under `path` (written to a temporary name, then renamed: readers never see a partial
// file); the other ranks wait for a file that is not older than their own start
bool rendezvous_blob(const std::string &path, int rank, void *blob, size_t bytes, double timeout_s) {
  static const time_t t_start = time(nullptr);
  if (rank == 0) {
    const std::string tmp = path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(blob, 1, bytes, f) == bytes;
    fclose(f);
    return ok && rename(tmp.c_str(), path.c_str()) == 0;
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    struct stat st;
    if (stat(path.c_str(), &st) == 0 && (size_t)st.st_size == bytes && st.st_mtime + 30 >= t_start) {
      FILE *f = fopen(path.c_str(), "rb");
      if (f) {
        const bool ok = fread(blob, 1, bytes, f) == bytes;
        fclose(f);
        if (ok) return true;
      }
    }
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
}

struct MpiSide {
  pmcgpu_ctx *ctx = nullptr;
  long nloc = -1;
  int d = 0;
  uint32_t iter = 0;
  bool comm = false;
};
std::map<const pmc_simu *, MpiSide> g_mside;
std::mutex g_mmu;
int g_comm_serial = 0;

MpiSide *mside_for(pmc_simu *psim, error **err) {
  std::lock_guard<std::mutex> lk(g_mmu);
  MpiSide &s = g_mside[psim];
  const World &w = world();
  if (!s.ctx) {
    int ndev = pmcgpu_device_count();
    int dev = getenv("PMCB200_DEVICE") ? pmcb200::dropin_device_ordinal() : (ndev > 0 ? w.local_rank % ndev : 0);
    const int rc = pmcgpu_create(dev, &s.ctx);
    if (rc) {
      *err = newErrorVA(rc, __func__, (char *)"", (char *)"libpmcb200: rank %d has no usable CUDA device %d (there is no CPU fallback)", *err, nullptr, w.rank, dev);
      g_mside.erase(psim);
      return nullptr;
    }
  }
  if (!s.comm && w.size > 1) {
    char id[PMCGPU_NCCL_ID_BYTES];
    memset(id, 0, sizeof id);
    if (w.rank == 0 && pmcgpu_nccl_unique_id(id) != 0) {
      *err = newErrorVA(PMCGPU_ERR_NCCL, __func__, (char *)"", (char *)"libpmcb200: NCCL is not loadable", *err, nullptr);
      return nullptr;
    }
    char path[512];
    snprintf(path, sizeof path, "%s.%d", w.rdv.c_str(), g_comm_serial++);
    if (!rendezvous_blob(path, w.rank, id, sizeof id, 300.0)) {
      *err = newErrorVA(pmc_badComm, __func__, (char *)"", (char *)"libpmcb200: rank %d: no NCCL id at %s (set PMCB200_RENDEZVOUS to a path all ranks share)", *err, nullptr, w.rank, path);
      return nullptr;
    }
    const int rc = pmcgpu_comm_init(s.ctx, id, w.rank, w.size);
    if (rc) {
      *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(s.ctx));
      return nullptr;
    }
    if (w.rank == 0) { pmcgpu_comm_allreduce(s.ctx, (double *)id, 1, 1); unlink(path); }  // everyone has joined: the file can go
    else { double z = 0; pmcgpu_comm_allreduce(s.ctx, &z, 1, 1); }
    s.comm = true;
  }
  return &s;
}

int mfail(MpiSide *s, int rc, error **err, const char *what) {
  *err = newErrorVA(rc, what, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(s->ctx));
  return rc;
}

// One sharded iteration.  do_update = 0: importance run (log-weights stay un-normalised).  Returns false on error.
struct StepOut { double perp = 0; long nok = 0; };
bool sharded_step(pmc_simu *psim, gsl_rng *r, int do_update, StepOut &out, error **err) {
  const World &w = world();
  MpiSide *s = mside_for(psim, err);
  if (!s) return false;
  distribution *prop = psim->proposal, *tgt = psim->target;
  if (!prop || prop->log_pdf != &mix_mvdens_log_pdf_void || !prop->data) {
    *err = newErrorVA(pmc_incompat, __func__, (char *)"", (char *)"libpmcb200 evaluates mix_mvdens proposals only", *err, nullptr);
    return false;
  }
  if (psim->pmc_filter && psim->pmc_filter != &filter_histogram) {
    *err = newErrorVA(pmc_incompat, __func__, (char *)"", (char *)"the sharded iteration runs filter_histogram (or no filter) on the device; other pmc_filter callbacks need the single-process entry points", *err, nullptr);
    return false;
  }
  mix_mvdens *m = (mix_mvdens *)prop->data;
  const int d = psim->ndim, K = (int)m->ncomp;
  // rank 0 owns the sample size and the generator (the reference's master draws everything): both are broadcast
  long hdr[3] = {psim->nsamples, 0, (long)psim->retry};
  if (w.rank == 0) { const uint64_t seed = pmcb200::dropin_seed_from(r); memcpy(&hdr[1], &seed, 8); }
  int rc = pmcgpu_comm_bcast(s->ctx, hdr, sizeof hdr, 0);
  if (rc) { mfail(s, rc, err, "bcast"); return false; }
  const long N = hdr[0];
  uint64_t seed;
  memcpy(&seed, &hdr[1], 8);
  int retry = (int)hdr[2];
  long first = 0, len = N;
  pmcgpu_shard_range(N, w.rank, w.size, &first, &len);
  if (s->nloc != len || s->d != d) {
    rc = pmcgpu_resize(s->ctx, len, d);
    if (rc) { mfail(s, rc, err, "resize"); return false; }
    s->nloc = len;
    s->d = d;
  }
  const int kind = pmcb200::dropin_device_kind(tgt);
  if (!pmcb200::dropin_push_model(s->ctx, m, kind != PMCGPU_TGT_HOST ? tgt : nullptr, psim->pb, d, err)) return false;
  const int *fil = psim->pmc_filter ? (const int *)psim->filter_data : nullptr;
  pmcgpu_set_filter_histogram(s->ctx, fil ? fil[0] : 0, fil ? fil[1] : 0, fil ? fil[2] : 0);
  std::vector<double> wv(K), mean((size_t)K * d), L((size_t)K * d * d), detL(K);
  std::vector<int> alive(K);
  pmcgpu_step_result res;
  memset(&res, 0, sizeof res);
  const uint32_t iter = s->iter++;
  // maxima of an importance run: every rank's local values, folded in rank order with the reference's "v > m" rule
  // (receive_importance_weight, pmc_mpi.c:502-508: a NaN on the master survives, a NaN on a client never wins)
  auto fold_maxima = [&](long nok, double maxW, double maxR, int rc_local) -> bool {
    std::vector<double> slot((size_t)3 * w.size + 1, 0.0);
    if (!rc_local) { slot[3 * w.rank] = maxW; slot[3 * w.rank + 1] = maxR; slot[3 * w.rank + 2] = (double)nok; }
    slot[(size_t)3 * w.size] = rc_local ? 1.0 : 0.0;
    const int rc2 = pmcgpu_comm_allreduce(s->ctx, slot.data(), (long)slot.size(), 0);
    if (rc_local || rc2) { mfail(s, rc_local ? rc_local : rc2, err, "importance"); return false; }
    if (slot[(size_t)3 * w.size] != 0.0) {
      *err = newErrorVA(pmc_badComm, __func__, (char *)"", (char *)"the importance run failed on another rank", *err, nullptr);
      return false;
    }
    double mw = slot[0], mr = slot[1], n = slot[2];
    for (int q = 1; q < w.size; ++q) {
      if (slot[3 * q] > mw) mw = slot[3 * q];
      if (slot[3 * q + 1] > mr) mr = slot[3 * q + 1];
      n += slot[3 * q + 2];
    }
    res.maxW = mw;
    res.maxR = mr;
    res.nok = (long)n;
    return true;
  };
  if (kind != PMCGPU_TGT_HOST && !do_update) {
    long nok = 0, nrej = 0;
    double maxW = 0, maxR = 0;
    rc = pmcgpu_draw_weights(s->ctx, seed, iter, first, 1.0, &nok, &maxW, &maxR, &nrej);
    if (!fold_maxima(nok, maxW, maxR, rc)) return false;
    rc = 0;
  } else if (kind != PMCGPU_TGT_HOST) {
    rc = pmcgpu_pmc_step(s->ctx, seed, iter, first, N, do_update, &retry, &res, wv.data(), mean.data(), L.data(), detL.data(), alive.data());
  } else {
    // opaque likelihood: this rank's callbacks run over this rank's shard, exactly what vanilla_pmc_mpi distributes
    long nrej = 0;
    rc = pmcgpu_set_target_host(s->ctx, d);
    if (!rc) rc = pmcgpu_draw(s->ctx, seed, iter, first, &nrej);
    double bad = rc ? 1.0 : 0.0;  // a rank whose draw failed must not leave the others waiting in a collective
    const int rc2 = pmcgpu_comm_allreduce(s->ctx, &bad, 1, 1);
    if (rc || rc2 || bad != 0.0) {
      if (rc || rc2) mfail(s, rc ? rc : rc2, err, "draw");
      else *err = newErrorVA(pmc_badComm, __func__, (char *)"", (char *)"the draw failed on another rank", *err, nullptr);
      return false;
    }
    const int nd = psim->n_ded;
    std::vector<double> X((size_t)len * d), Xd((size_t)len * (nd > 0 ? nd : 0)), post(len > 0 ? len : 1);
    std::vector<short> ok(len > 0 ? len : 1);
    rc = pmcgpu_download(s->ctx, X.data(), nullptr, nullptr, nullptr, nullptr);
    if (rc) { mfail(s, rc, err, "download"); return false; }
    const bool fine = pmcb200::dropin_host_target_values(&distribution_lkl, &distribution_retrieve, tgt, X.data(), Xd.data(), len, d, nd,
                                                         0, post.data(), ok.data(), err);
    bad = fine ? 0.0 : 1.0;
    pmcgpu_comm_allreduce(s->ctx, &bad, 1, 1);
    if (bad != 0.0) {
      if (fine) *err = newErrorVA(pmc_badComm, __func__, (char *)"", (char *)"the target aborted on another rank", *err, nullptr);
      return false;
    }
    if (w.rank == 0 && nd > 0) memcpy(psim->X_ded + (size_t)first * nd, Xd.data(), sizeof(double) * Xd.size());
    if (do_update) {
      rc = pmcgpu_step_given_post(s->ctx, post.data(), ok.data(), first, N, 1.0, do_update, &retry, &res, wv.data(), mean.data(),
                                  L.data(), detL.data(), alive.data());
    } else {
      long nok = 0;
      double maxW = 0, maxR = 0;
      rc = pmcgpu_weights_host_post(s->ctx, post.data(), ok.data(), 1.0, first, &nok, &maxW, &maxR);
      if (!fold_maxima(nok, maxW, maxR, rc)) return false;
      rc = 0;
    }
  }
  pmcgpu_set_filter_histogram(s->ctx, 0, 0, 0);
  if (rc) { mfail(s, rc, err, "pmc_step"); return false; }
  // the master's pmc_simu holds the whole sample afterwards (receive_importance_weight, pmc_mpi.c:447-520)
  const bool root = w.rank == 0;
  rc = pmcgpu_comm_gather_store(s->ctx, 0, N, first, root ? psim->X : nullptr, root ? psim->indices : nullptr,
                                root ? psim->weights : nullptr, root ? psim->log_rho : nullptr, root ? psim->flg : nullptr);
  if (rc) { mfail(s, rc, err, "gather"); return false; }
  psim->maxW = res.maxW;
  psim->maxR = res.maxR;
  psim->logSum = res.logSum;
  psim->isLog = do_update ? 0 : 1;
  psim->retry = retry;
  if (do_update) pmcb200::dropin_unflatten(m, wv.data(), mean.data(), L.data(), detL.data());  // identical on every rank
  out.perp = res.perplexity;
  out.nok = res.nok;
  return true;
}
}  // namespace

extern "C" {

// pmc_mpi.c:20-34: the master holds the sample, the clients a placeholder (their shards live in HBM)
pmc_simu *pmc_simu_init_mpi(long nsamples, int ndim, int nded, error **err) {
  const World &w = world();
  pmc_simu *psim = pmc_simu_init_plus_ded(w.rank == 0 ? nsamples : 1, ndim, nded, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return nullptr; }
  psim->mpi_rank = w.rank;
  psim->mpi_size = w.size;
  return psim;
}

void pmc_simu_realloc_mpi(pmc_simu *psim, long newsamples, error **err) {  // pmc_mpi.c:36-41
  if (psim->mpi_rank == 0) {
    pmc_simu_realloc(psim, newsamples, err);
    if (isError(*err)) *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err);
  }
}

// pmc_mpi.c:44-95: draw + importance weights over all ranks; the master ends with the un-normalised sample
size_t pmc_simu_importance_mpi(pmc_simu *psim, gsl_rng *r, error **err) {
  if (!psim->proposal || !psim->proposal->data) { *err = newErrorVA(pmc_allocate, __func__, (char *)"", (char *)"proposal undefined", *err, nullptr); return 0; }
  if (!psim->target || !psim->target->data) { *err = newErrorVA(pmc_allocate, __func__, (char *)"", (char *)"target undefined", *err, nullptr); return 0; }
  StepOut o;
  if (!sharded_step(psim, r, 0, o, err)) {
    if (isError(*err)) *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err);
    return 0;
  }
  return (size_t)o.nok;
}

// pmc_mpi.c:97-125: one PMC iteration; the perplexity is returned on the master (0 elsewhere, as the reference does)
double pmc_simu_pmc_step_mpi(pmc_simu *psim, gsl_rng *r, error **err) {
  if (!psim->pmc_update) { *err = newErrorVA(pmc_allocate, __func__, (char *)"", (char *)"update function undefined", *err, nullptr); return 0; }
  if (psim->pmc_update != &update_prop_rb_void) {
    *err = newErrorVA(pmc_incompat, __func__, (char *)"", (char *)"the sharded iteration implements update_prop_rb", *err, nullptr);
    return 0;
  }
  if (!psim->proposal || !psim->proposal->data) { *err = newErrorVA(pmc_allocate, __func__, (char *)"", (char *)"proposal undefined", *err, nullptr); return 0; }
  if (!psim->target || !psim->target->data) { *err = newErrorVA(pmc_allocate, __func__, (char *)"", (char *)"target undefined", *err, nullptr); return 0; }
  StepOut o;
  if (!sharded_step(psim, r, 1, o, err)) {
    if (isError(*err)) *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err);
    return 0;
  }
  return psim->mpi_rank == 0 ? o.perp : 0.0;
}

// pmc_mpi.c:159-175: rank 0's mixture to every rank (serialize_mix_mvdens wire format); clients get a new object
void *mvdens_broadcast_mpi(void *data, error **err) {
  const World &w = world();
  if (w.size <= 1) return data;
  static pmc_simu key;  // one communicator for all broadcasts of the process
  MpiSide *s = mside_for(&key, err);
  if (!s) return nullptr;
  size_t sz = 0;
  void *buf = nullptr;
  if (w.rank == 0) {
    buf = serialize_mix_mvdens((mix_mvdens *)data, &sz, err);
    if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return nullptr; }
  }
  long n = (long)sz;
  int rc = pmcgpu_comm_bcast(s->ctx, &n, sizeof n, 0);
  if (!rc && w.rank != 0) buf = malloc((size_t)n);
  if (!rc) rc = pmcgpu_comm_bcast(s->ctx, buf, (size_t)n, 0);
  if (rc) { free(buf); mfail(s, rc, err, __func__); return nullptr; }
  void *res = data;
  if (w.rank != 0) {
    res = deserialize_mix_mvdens(buf, (size_t)n, err);
    if (isError(*err)) *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err);
  }
  free(buf);
  return res;
}

distribution *init_distribution_full_mpi(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                         simulate_func *simulate, int nded, retrieve_ded_func *retrieve,
                                         mpi_exchange_func *broadcast_mpi, error **err) {  // pmc_mpi.c:177-198
  void *rdata = data;
  if (broadcast_mpi) {
    rdata = broadcast_mpi(data, err);
    if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return nullptr; }
  }
  distribution *dist = init_distribution_full(ndim, rdata, log_pdf, freef, simulate, nded, retrieve, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return nullptr; }
  dist->broadcast_mpi = broadcast_mpi;
  return dist;
}
distribution *init_distribution_mpi(int ndim, void *data, posterior_log_pdf_func *log_pdf, posterior_log_free *freef,
                                    simulate_func *simulate, mpi_exchange_func *broadcast_mpi, error **err) {  // :200-211
  distribution *d = init_distribution_full_mpi(ndim, data, log_pdf, freef, simulate, 0, nullptr, broadcast_mpi, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return nullptr; }
  return d;
}
void distribution_broadcast(distribution *dist, error **err) {  // pmc_mpi.c:213-228
  if (!dist->broadcast_mpi) { *err = newErrorVA(pmc_incompat, __func__, (char *)"", (char *)"Cannot broadcast the distribution", *err, nullptr); return; }
  void *rdata = dist->broadcast_mpi(dist->data, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return; }
  if (rdata != dist->data) {
    if (dist->free) dist->free(&dist->data);
    dist->data = rdata;
  }
}
distribution *mix_mvdens_distribution_mpi(int ndim, void *mxmv, error **err) {  // pmc_mpi.c:230-238
  distribution *d = init_distribution_full_mpi(ndim, mxmv, &mix_mvdens_log_pdf_void, &mix_mvdens_free_void, &simulate_mix_mvdens_void, 0,
                                               nullptr, &mvdens_broadcast_mpi, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return nullptr; }
  return d;
}
void pmc_simu_init_classic_importance_mpi(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data, int prop_print_step,
                                          error **err) {  // pmc_mpi.c:127-143
  pmc_simu_init_target(psim, target, pb, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return; }
  distribution *prop = mix_mvdens_distribution_mpi(target->ndim, prop_data, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return; }
  pmc_simu_init_proposal(psim, prop, prop_print_step, err);
  if (isError(*err)) *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err);
}
void pmc_simu_init_classic_pmc_mpi(pmc_simu *psim, distribution *target, parabox *pb, void *prop_data, int prop_print_step,
                                   void *filter_data, filter_func *pmc_filter, error **err) {  // pmc_mpi.c:145-157
  pmc_simu_init_classic_importance_mpi(psim, target, pb, prop_data, prop_print_step, err);
  if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return; }
  pmc_simu_init_pmc(psim, filter_data, pmc_filter, &update_prop_rb_void);
}

// pmc_mpi.c:558-600 as entry points of their own: rank 0's mixture to the clients
void send_mix_mvdens(mix_mvdens *m, int nproc, error **err) {
  (void)nproc;
  mvdens_broadcast_mpi(m, err);
}
mix_mvdens *receive_mix_mvdens(int myid, int nproc, error **err) {
  (void)myid; (void)nproc;
  return (mix_mvdens *)mvdens_broadcast_mpi(nullptr, err);
}

// ---- the message-level functions of pmc_mpi.c:334-520, for callers that write their own master / client loop ---------
// Same slice layout as the reference: the master keeps the first N/nproc samples, client i the next N/nproc (+1 for the
// first N mod nproc clients).  Each MPI_Send / MPI_Recv pair becomes a collective every rank takes part in: the master
// enters it from send_simulation / receive_importance_weight, the clients from receive_simulation / send_importance_weight.
namespace {
struct Slices { std::vector<long> first, len; };
Slices ref_slices(long N, int nproc) {  // pmc_mpi.c:344-352, 366-372
  Slices s;
  s.first.assign(nproc, 0);
  s.len.assign(nproc, 0);
  const long per = N / nproc;
  long remaining = N - per * nproc, cur = per;
  s.len[0] = per;
  for (int i = 1; i < nproc; ++i) {
    long b = per;
    if (remaining > 0) { b = per + 1; --remaining; }
    s.first[i] = cur;
    s.len[i] = b;
    cur += b;
  }
  return s;
}
pmcgpu_ctx *msg_ctx(pmc_simu *psim, error **err) {
  MpiSide *s = mside_for(psim, err);
  return s ? s->ctx : nullptr;
}
// one array of the master's pmc_simu <-> the clients' slices: everybody ends with the whole array in `full`
bool exchange(pmcgpu_ctx *c, std::vector<char> &full, size_t elem, long N, const Slices &sl, int rank, const void *mine, long nmine, error **err) {
  full.assign((size_t)N * elem + 8, 0);
  if (mine && nmine > 0) memcpy(full.data() + (size_t)sl.first[rank] * elem, mine, (size_t)nmine * elem);
  const int rc = pmcgpu_comm_allgather_bytes(c, full.data(), (size_t)N * elem, (size_t)sl.first[rank] * elem, (size_t)nmine * elem);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return false; }
  return true;
}
}  // namespace

// master: the slices of X, X_ded and flg leave for the clients; returns the master's own sample count
size_t send_simulation(pmc_simu *psim, int nproc, error **err) {
  const World &w = world();
  pmcgpu_ctx *c = msg_ctx(psim, err);
  if (!c) return 0;
  long hdr[3] = {psim->nsamples, psim->ndim, psim->n_ded};
  int rc = pmcgpu_comm_bcast(c, hdr, sizeof hdr, 0);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  const long N = hdr[0];
  const Slices sl = ref_slices(N, nproc);
  std::vector<char> full;
  // the master owns everything: it contributes the whole arrays (first = 0, len = N in the collective)
  Slices all = sl;
  all.first[w.rank] = 0;
  if (!exchange(c, full, sizeof(double) * psim->ndim, N, all, w.rank, psim->X, N, err)) return 0;
  if (psim->n_ded > 0 && !exchange(c, full, sizeof(double) * psim->n_ded, N, all, w.rank, psim->X_ded, N, err)) return 0;
  if (!exchange(c, full, sizeof(short), N, all, w.rank, psim->flg, N, err)) return 0;
  return (size_t)sl.len[0];
}
// client: receives its slice (the pmc_simu is re-allocated to the slice size, pmc_mpi.c:392-395)
void receive_simulation(pmc_simu *psim, int basetag, int myid, error **err) {
  const int nproc = basetag;  // the reference passes mpi_size as the tag base (pmc_mpi.c:66)
  pmcgpu_ctx *c = msg_ctx(psim, err);
  if (!c) return;
  long hdr[3] = {0, 0, 0};
  int rc = pmcgpu_comm_bcast(c, hdr, sizeof hdr, 0);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return; }
  const long N = hdr[0];
  const Slices sl = ref_slices(N, nproc);
  const long mine = sl.len[myid];
  if (mine != psim->nsamples) {
    pmc_simu_realloc(psim, mine, err);
    if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return; }
  }
  std::vector<char> full;
  Slices none = sl;  // a client contributes nothing
  if (!exchange(c, full, sizeof(double) * psim->ndim, N, none, myid, nullptr, 0, err)) return;
  memcpy(psim->X, full.data() + (size_t)sl.first[myid] * psim->ndim * sizeof(double), (size_t)mine * psim->ndim * sizeof(double));
  if (psim->n_ded > 0) {
    if (!exchange(c, full, sizeof(double) * psim->n_ded, N, none, myid, nullptr, 0, err)) return;
    memcpy(psim->X_ded, full.data() + (size_t)sl.first[myid] * psim->n_ded * sizeof(double), (size_t)mine * psim->n_ded * sizeof(double));
  }
  if (!exchange(c, full, sizeof(short), N, none, myid, nullptr, 0, err)) return;
  memcpy(psim->flg, full.data() + (size_t)sl.first[myid] * sizeof(short), (size_t)mine * sizeof(short));
}

namespace {
// the gather of pmc_mpi.c:410-520, entered by all ranks: weights, X_ded, log_rho, flags, and (nok, maxW, maxR) per rank
size_t gather_weights(pmc_simu *psim, int nproc, int rank, size_t nok, long N_master_total, error **err) {
  pmcgpu_ctx *c = msg_ctx(psim, err);
  if (!c) return 0;
  long hdr[1] = {N_master_total};
  int rc = pmcgpu_comm_bcast(c, hdr, sizeof hdr, 0);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  const long N = hdr[0];
  const Slices sl = ref_slices(N, nproc);
  const long mine = sl.len[rank];
  std::vector<double> slot((size_t)3 * nproc, 0.0);
  slot[3 * rank] = (double)nok; slot[3 * rank + 1] = psim->maxW; slot[3 * rank + 2] = psim->maxR;
  rc = pmcgpu_comm_allreduce(c, slot.data(), (long)slot.size(), 0);
  if (rc) { *err = newErrorVA(rc, __func__, (char *)"", (char *)"%s", *err, nullptr, pmcgpu_last_error(c)); return 0; }
  std::vector<char> fw, fd, fr, ff;
  if (!exchange(c, fw, sizeof(double), N, sl, rank, psim->weights, mine, err)) return 0;
  if (psim->n_ded > 0 && !exchange(c, fd, sizeof(double) * psim->n_ded, N, sl, rank, psim->X_ded, mine, err)) return 0;
  if (!exchange(c, fr, sizeof(double), N, sl, rank, psim->log_rho, mine, err)) return 0;
  if (!exchange(c, ff, sizeof(short), N, sl, rank, psim->flg, mine, err)) return 0;
  if (rank != 0) return nok;
  size_t cnt = (size_t)slot[0];
  for (int i = 1; i < nproc; ++i) {
    const long f = sl.first[i], n = sl.len[i];
    const size_t inok = (size_t)slot[3 * i];
    cnt += inok;
    if (inok == 0) {  // pmc_mpi.c:466-474: a client without a single valid sample
      for (long j = 0; j < n; ++j) { psim->weights[f + j] = 0; psim->log_rho[f + j] = 0; psim->flg[f + j] = 0; }
      continue;
    }
    memcpy(psim->weights + f, fw.data() + (size_t)f * sizeof(double), (size_t)n * sizeof(double));
    if (psim->n_ded > 0) memcpy(psim->X_ded + (size_t)f * psim->n_ded, fd.data() + (size_t)f * psim->n_ded * sizeof(double), (size_t)n * psim->n_ded * sizeof(double));
    memcpy(psim->log_rho + f, fr.data() + (size_t)f * sizeof(double), (size_t)n * sizeof(double));
    memcpy(psim->flg + f, ff.data() + (size_t)f * sizeof(short), (size_t)n * sizeof(short));
    if (slot[3 * i + 1] > psim->maxW) psim->maxW = slot[3 * i + 1];  // pmc_mpi.c:502-508
    if (slot[3 * i + 2] > psim->maxR) psim->maxR = slot[3 * i + 2];
  }
  return cnt;
}
}  // namespace

// client: un-normalised weights, deduced parameters, log_rho, flags and the running maxima go to the master
void send_importance_weight(int myid, int basetag, pmc_simu *psim, size_t nok) {
  error *e = initError();
  gather_weights(psim, basetag, myid, nok, 0, &e);
  if (isError(e)) printError(stderr, e);  // the reference's prototype has no error argument
  endError(&e);
}
// master: psim->nsamples must be the whole sample again (pmc_mpi.c:84-86); master_cnt / master_sample describe its own slice
size_t receive_importance_weight(pmc_simu *psim, int nproc, int master_cnt, int master_sample, error **err) {
  (void)master_sample;
  return gather_weights(psim, nproc, 0, (size_t)master_cnt, psim->nsamples, err);
}

/* test hook: rank, size and rendezvous path this process resolved from its environment */
int pmcb200_mpi_world(int *rank, int *size, char *rdv, int rdv_len) {
  const World &w = world();
  if (rank) *rank = w.rank;
  if (size) *size = w.size;
  if (rdv && rdv_len > 0) snprintf(rdv, (size_t)rdv_len, "%s", w.rdv.c_str());
  return 0;
}
/* test hook: the rendezvous primitive (rank 0 publishes, the others read) */
int pmcb200_mpi_rendezvous(const char *path, int rank, void *blob, long bytes, double timeout_s) {
  return rendezvous_blob(path, rank, blob, (size_t)bytes, timeout_s) ? 0 : 1;
}

}  // extern "C"
