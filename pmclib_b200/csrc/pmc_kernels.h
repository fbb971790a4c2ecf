// pmc_kernels.h — declarations shared by the kernel translation units and the C-ABI layer (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/pmcb200.h"
#include "pmc_dev.cuh"

#ifndef PMCB200_MAXD
#define PMCB200_MAXD 128  // largest dimension the generic kernels hold in per-thread local arrays
#endif

namespace pmcb200 {

// cudaFuncSetAttribute is per device: a kernel's opt-in shared-memory size must be set once on every GPU this process
// drives (single-process multi-GPU mode of the drop-in layer), from whichever host thread gets there first.
template <typename Kern>
inline bool set_max_smem_once(Kern kern, size_t bytes, std::atomic<unsigned long long> &done_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done_mask.load(std::memory_order_acquire) & bit) return true;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return false;
  done_mask.fetch_or(bit, std::memory_order_release);
  return true;
}

constexpr int GEN_WEIGHT_THREADS = 128;

enum { CNT_REJECT = 0, CNT_DRAWFAIL = 1, CNT_NOK = 2, CNT_NORMCOUNT = 3, CNT_FLAGGED = 4, CNT_POSINF = 5, CNT_NUM = 8 };

struct TargetDev {
  int kind, d;  // d = dimension the density itself sees (nfull when a defaults map is set)
  double b, sig1;
  MixDev mix;
  int nfull;  // 0 = no defaults map
  const int *is_def;
  const double *def_val;
};

// one term of a combined target (distribution.c:350-396): a density of kind PMCGPU_TGT_BANANA / _MVDENS / _MIX evaluated on
// the coordinates dim[0..nd) of the full parameter vector
struct CombTermDev {
  int kind, nd;
  const int *dim;
  double b, sig1;
  MixDev mix;
};

#ifdef __CUDACC__
// Common end of the one-thread-per-sample weight kernels (k_weights_generic, k_weights_q): block maxima of the log-weight
// and of log_rho with the reference's "v > m" semantics (a NaN never wins, pmc.c:552-554,579) and the count of ok samples.
__device__ __forceinline__ void weight_block_epilogue(double my_w, double my_r, bool okflag, double *blk_maxw,
                                                      double *blk_maxr, unsigned long long *counters) {
  if (my_w != my_w) my_w = -INFINITY;
  if (my_r != my_r) my_r = -INFINITY;
  my_w = warp_max_nan_ignoring(my_w);
  my_r = warp_max_nan_ignoring(my_r);
  __shared__ double sw[32], sr[32];
  __shared__ unsigned int scnt;
  if (threadIdx.x == 0) scnt = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sw[wid] = my_w; sr[wid] = my_r; }
  const unsigned int nok = __popc(__ballot_sync(0xffffffffu, okflag));
  if (lane == 0 && nok) atomicAdd(&scnt, nok);
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double a = lane < nw ? sw[lane] : -INFINITY, b = lane < nw ? sr[lane] : -INFINITY;
    a = warp_max_nan_ignoring(a);
    b = warp_max_nan_ignoring(b);
    if (lane == 0) {
      blk_maxw[blockIdx.x] = a;
      blk_maxr[blockIdx.x] = b;
      if (scnt) atomicAdd(&counters[CNT_NOK], (unsigned long long)scnt);
    }
  }
}
#endif

// number of kernels launched by this library in this process (bench.py reports it as gpu_launches)
long launch_count();
void count_launch(int n = 1);

// ---- generic kernels (pmc_generic.cu) ------------------------------------------------------------------------
void launch_draw_generic(const MixDev &m, const BoxDev &bx, uint64_t seed, uint32_t iter, long first_index, long N,
                         int max_attempts, double *X, unsigned long long *indices, short *flg,
                         unsigned long long *counters, cudaStream_t s, const unsigned int *list = nullptr, int attempt0 = 0);
void launch_weights_generic(const MixDev &m, const TargetDev &tg, int mode, const double *post_in, const short *ok_in,
                            double t_temper, long first_index, long N, const double *X, double *lw, double *log_rho,
                            short *flg, double *ldscratch, int ldstride, double *blk_maxw, double *blk_maxr,
                            unsigned long long *counters, double *first_vals, cudaStream_t s);
void launch_mix_log_pdf(const MixDev &m, long n, const double *X, double *out, double *ld, int ldstride, cudaStream_t s);
void launch_scalar_product(const MixDev &m, int k, long n, const double *X, double *out, cudaStream_t s);
void launch_combine_target(const CombTermDev *terms, int nterms, int d, long n, const double *X, double *out, double *ld,
                           int ldstride, cudaStream_t s);
void launch_target_log_pdf(const TargetDev &tg, int dfree, long n, const double *X, double *out, double *ld,
                           int ldstride, cudaStream_t s);
void launch_ranindx(const double *cw, int K, long n, const double *z, unsigned long long *idx, cudaStream_t s);
void launch_isinbox(const BoxDev &bx, long n, const double *X, int *inside, cudaStream_t s);
void launch_norm_exp_sum(long N, double maxW, double *w, short *flg, double *partial, int nblocks,
                         unsigned long long *counters, cudaStream_t s);
void launch_norm_scale(long N, double S, double *w, int nblocks, cudaStream_t s);
void launch_perplexity(long N, const double *w, const short *flg, double *partial, int nblocks,
                       unsigned long long *counters, cudaStream_t s);
void launch_count_indices(long N, int K, const unsigned long long *indices, const short *flg,
                          unsigned long long *count, int nblocks, cudaStream_t s);
void launch_rb_resp(const MixDev &m, long n, long base, const double *X, const double *w, const double *log_rho,
                    const short *flg, const unsigned long long *indices, const unsigned long long *count, double maxR,
                    int hard, double *gw8, double *w8, cudaStream_t s);
void launch_rb_reduce_mean(int K, int d, long n, long base, int nsplit, const double *X, const double *gw8,
                           const double *w8, double *part, double *out, cudaStream_t s);
void launch_rb_reduce_cov(int K, int d, long n, long base, int nsplit, const double *X, const double *gw8,
                          const double *newmean, double *part, double *out, cudaStream_t s);
void launch_norm_scale_perp(long N, double S, double *w, const short *flg, double *partial, int nblocks,
                            unsigned long long *counters, cudaStream_t s);

void launch_minmax_ok(long N, const double *w, const short *flg, double *partial, int nblocks, unsigned long long *cnt,
                      cudaStream_t s);
void launch_flag_outside(long N, const double *w, short *flg, double bl, double br, int nblocks, unsigned long long *cnt,
                         cudaStream_t s);
void launch_weighted_coord(long N, int d, int a, const double *X, const double *w, const short *flg, double *partial,
                           int nblocks, cudaStream_t s);

// ---- resampling (pmc_post.cu) ------------------------------------------------------------------------------------
long scan_blocks(long N);
void launch_weight_cdf(long N, const double *w, int mode, double *cdf, double *blocksum, cudaStream_t s);
void launch_resample(long N, long ndraw, const double *cdf, uint64_t seed, uint32_t stream, int mode,
                     unsigned long long *isample, unsigned int *counts, cudaStream_t s);

void launch_narrow_idx(long N, const unsigned long long *in, unsigned char *out, int sm_count, cudaStream_t s);
void launch_count_key_ge(long N, long first, const double *w, const short *flg, unsigned long long key,
                         unsigned long long *cnt, int nblocks, cudaStream_t s);
void launch_clip_apply(long N, double T, double *w, short *flg, double *partial, unsigned long long *cnt, int nblocks,
                       cudaStream_t s);
void launch_scale_ok(long N, double S, double *w, const short *flg, int nblocks, cudaStream_t s);

// ---- fused small-d iteration kernel (pmc_fused.cu) -----------------------------------------------------------
constexpr int FUSED_MAXD = 16;

struct FusedArgs {
  // model (device pointers to packed parameter blocks, see pack_mixture_small())
  const double *mixpack;
  int mix_doubles, comp_stride, K;
  const double *tgtpack;  // TGT_MIX / TGT_MVDENS: same packing, Kt components
  int tgt_doubles, tgt_stride, Kt;
  int tgt_kind;
  double b, sig1, t_temper;
  const double *post_in;  // TGT_HOST: target values evaluated by the caller
  const short *ok_in;
  // box in canonical slab form (thresholds of the -e_j and +e_j faces; +inf when absent)
  int has_box;
  double lo_thr[FUSED_MAXD], hi_thr[FUSED_MAXD];
  double pivot[FUSED_MAXD];
  // work
  int do_draw, do_stats;
  long N, first_index;
  uint64_t seed;
  uint32_t iter;
  int max_attempts;
  // sample store
  double *X, *lw, *log_rho;
  unsigned long long *indices;
  short *flg;
  // per-block outputs
  double *partials;  // [grid][partial_len]
  int partial_len;
  unsigned long long *counters;  // CNT_*
  unsigned long long *count_k;   // [K] histogram of indices over ok samples
  double *first_vals;            // values of the sample with global index 0 (4 doubles)
  double *resp;                  // [K][N] responsibilities (split pipeline, MODE 1)
  const double2 *zig;            // ziggurat table (pmc_fused2.cu draw)
};

struct FusedPlan {
  int valid;
  int D, RK, RE, NR, SPL;
  int grid, threads;
  size_t smem_bytes;
  int partial_len;   // doubles per block partial: NR*32*RK*RE accumulators + 4 scalars (m, S, maxR, spare)
  int nstat;         // K*E canonical statistics
};

// picks an instantiation for (d, K) or returns valid=0
FusedPlan fused_plan(int d, int K, int Kt, int sm_count, int want_spl);
// number of doubles of one packed component / whole packed mixture for dimension D
int fused_comp_stride(int D);
// host-side packing of a mixture (reference layout in, packed out); returns number of doubles written
int fused_pack_mixture(int K, int D, const double *wght, const double *mean, const double *L, const double *logdet,
                       double *out);
int launch_fused(const FusedPlan &p, const FusedArgs &a, cudaStream_t s);
// combines the block partials: stats_out[nstat + 4]: canonical statistics rescaled to the global max M, then
// M (max log-weight over ok samples other than global sample 0), S = sum exp(lw - M), maxR-partial, spare
int launch_fused_combine(const FusedPlan &p, const double *partials, int K, double *stats_out, cudaStream_t s);

// second-generation kernel (pmc_fused2.cu): (D, K) templated, mixture in constant memory, DMMA statistics.
// plan.valid == 2; plan.RK / plan.RE carry K / KT.  h_mix / h_tgt are the host copies of the packed blocks.
FusedPlan fused2_plan(int d, int K, int Kt, int tgt_kind, int sm_count, int want_spl);
FusedPlan fused2_plan_draw(int d, int K, int sm_count);
int fused2_max_blocks(const FusedPlan &p);  // upper bound of the grid of any launch shape (size of the partials buffer)
// mode 0: weights only, 1: weights + responsibilities (split pipeline), 2: statistics fused.  Returns blocks launched
// through *nblocks_out.
int launch_fused2(const FusedPlan &p, const FusedArgs &a, const double *h_mix, const double *h_tgt, int mode,
                  int *nblocks_out, cudaStream_t s);
int launch_fused2_max(const FusedPlan &p, const double *partials, int nblocks, double *out2, cudaStream_t s);
int launch_fused2_stats(const FusedPlan &p, const double *X, const double *lw, const double *resp, const short *flg,
                        const double *Mdev, const double *pivot, long N, double *partials, cudaStream_t s);
int launch_fused2_combine(const FusedPlan &p, const double *partials, int nblocks, double *stats_out, cudaStream_t s);

// ---- third-generation small-d iteration (pmc_iter3.cu): grouped draw, table-driven weights kernel, statistics kernel
// with the normalisation fused in, device-side reductions ----------------------------------------------------------------
enum { I3_MAXW = 0, I3_MAXR, I3_M, I3_SREF, I3_INVS, I3_LOGSREF, I3_CLEAN, I3_NOK_W, I3_NREJ, I3_NFAIL, I3_MLOC, I3_SCAL_LEN = 16 };
constexpr int I3_COLL_LEN = 16;  // [0..7] all-reduce(max) block, [8..15] all-reduce(sum) block

struct Iter3Plan {
  int valid;
  int D, K, KT;      // instantiation; K >= the mixture's K (extra components are packed as dead ones)
  int E, NT, MT, AG, CYC, acc_len, partial_len;  // AG > 0: remainder components in product tiles; CYC: cyclic feature layout (S3Cfg)
  int stats_blocks, weights_blocks, sm_count;
};
struct Iter3Weights {
  const double *X;
  long N, first_index;
  int tgt_kind;
  double b, sig1, t_temper;
  const double *post_in;
  const short *ok_in;
  double shift[FUSED_MAXD];
  double *lw, *log_rho, *resp;  // resp == nullptr: no responsibilities
  short *flg;
  double *partials;
  unsigned long long *counters;
  double *first_vals;
  MixDev mix, tmix;
  int variant;  // launch shape of the weights kernel (pmc_iter3.cu:launch_w3)
};
struct Iter3Stats {
  const double *X, *resp;
  double *W;
  short *flg;
  const unsigned long long *indices;
  const double *scal;
  double pivot[FUSED_MAXD];
  long N;
  double *partials;
  int with_stats;  // 0: normalisation, perplexity sums and the component histogram only
};
struct Iter3Draw {
  long N, first_index;
  uint64_t seed;
  uint32_t iter;
  int max_attempts, has_box;
  double lo_thr[FUSED_MAXD], hi_thr[FUSED_MAXD];
  double *X;
  unsigned long long *indices;
  short *flg;
  const double2 *zig;
  const double *gpack;
  unsigned long long *counters;
};
Iter3Plan iter3_plan(int d, int K, int Kt, int tgt_kind, int sm_count);
int iter3_wstride(int D);
int iter3_dstride(int D);
void iter3_pack_w(int K, int D, const double *wght, const double *mean, const double *L, const double *logdet,
                  const double *shift, double *out, int pre = 0);
void iter3_pack_d(int K, int D, const double *mean, const double *L, double *out);
double iter3_shift_ratio(int K, int D, const double *wght, const double *mean, const double *L, const double *shift);
// constant-bank uploads on the stream (any pointer may be null = unchanged)
int iter3_upload(const Iter3Plan &p, const double *h_w, int nw, const double *h_t, int nt, const double *h_d, int nd,
                 const double *h_cw, int ncw, cudaStream_t s);
int launch_draw3(const Iter3Plan &p, const Iter3Draw &d, cudaStream_t s);
int launch_weights3(const Iter3Plan &p, const Iter3Weights &w, int *nblocks, cudaStream_t s);
// phases: 1 = local maxima -> coll[0..7], 2 = normaliser relative to the (global) maximum -> coll[8..15], 4 = scalars
int launch_red3(const double *partials, int nblocks, const unsigned long long *counters, const double *first_vals,
                int has_first, double *coll, double *scal, int phases, double n_total, cudaStream_t s, double *gath = nullptr,
                int rank = 0, int nranks = 1);
int launch_stats3(const Iter3Plan &p, const Iter3Stats &st, cudaStream_t s);
// out: stats[K*E] (with_stats) | H, Q, ok, flagged-out | count_k[K]
int launch_comb3(const Iter3Plan &p, const double *partials, int with_stats, const double *scal, double *out, cudaStream_t s);

// ---- large-d tensor-path kernels (pmc_mma.cu): DMMA whitening and DMMA sufficient statistics, 16 < d <= 128 --------
int mma_nt(int d);                    // number of 8-row tiles the dimension is padded to (0 = not supported)
size_t mma_comp_stride(int d);        // doubles of one packed component
// host: mean + (invert ? inverse factor : factor) of one component in B-fragment order
// inv_discrepancy (invert only, may be null): conditioning probe, see pmc_mma.cu:inverse_discrepancy
void mma_pack_component(int d, const double *mean, const double *L, int invert, double *out, double *inv_discrepancy = nullptr);
// q[kmap[j]][i] = |L_j^-1 (x_i - mu_j)|^2 for the J packed components, i < n; Q row stride qs
int launch_q_mma(const double *X, long n, int d, const double *pack, int J, const int *kmap, double *Q, long qs,
                 int sm_count, cudaStream_t s);
void launch_weights_q(const MixDev &m, const TargetDev &tg, int mode, const double *post_in, const short *ok_in,
                      double t_temper, long first_index, long N, const double *X, double *Q, double *Qt,
                      long qs, double *lw, double *log_rho, short *flg, double *blk_maxw, double *blk_maxr,
                      unsigned long long *counters, double *first_vals, int store_ld, cudaStream_t s);
void launch_lse_q(const MixDev &m, long n, double *Q, long qs, double *out, cudaStream_t s);
// wpart (Student-t mixtures only, else nullptr): [resp_q_blocks(N)][K] block sums of the un-gamma-weighted responsibilities
int resp_q_blocks(long N);
void launch_resp_q(const MixDev &m, long N, const double *w, const double *log_rho, const short *flg,
                   const unsigned long long *count, double *Q, long qs, double *wpart, cudaStream_t s);
// grouped draw: component pick + histogram, permutation, then x = mu + L g tile by tile (tiles: int4 {packed comp, first
// position in perm, samples, 0})
int mma_draw_tile_samples(int d);
void launch_draw_pick(const double *cw, int K, uint64_t seed, uint32_t iter, long first_index, long N,
                      unsigned long long *indices, unsigned int *cnt, int nblocks, cudaStream_t s);
void launch_draw_perm(long N, int K, const unsigned long long *indices, unsigned int *cursor, unsigned int *perm,
                      int nblocks, cudaStream_t s);
// thr: nullptr, or the slab box as [d] thresholds of the -e_j faces followed by [d] of the +e_j faces; samples whose
// first attempt falls outside are appended to redo[] (count in *nredo) with flg = 0
int launch_draw_mma(const double *pack, const void *tiles, int ntiles, const unsigned int *perm, const int *dfj, int d,
                    uint64_t seed, uint32_t iter, long first_index, double *X, short *flg, unsigned long long *counters,
                    const double *thr, unsigned int *redo, unsigned int *nredo, int sm_count, cudaStream_t s);
int mma_stats_chunks(long n, int J, int sm_count);
size_t mma_stats_partial_doubles(int d, int J, int nchunk);
// out[k] = { W_k, s1_k[d], S2_k[d*d] } for the J packed components (rows of other components untouched).  Xc is a
// scratch of n*d doubles that receives the centred sample (zero rows where flg == 0).
int launch_stats_mma(const double *X, const short *flg, long n, int d, const double *G, long qs, int J, const int *kmap,
                     const double *pivot, int nchunk, double *Xc, double *part, double *out, int sm_count, cudaStream_t s);

}  // namespace pmcb200
