// pmc_post.cu — resampling of a weighted PMC sample on the device (the step after the last iteration):
//   sample_from_pmc_simu (pmc.c:1585-1601): N indices drawn i.i.d. with probability proportional to weights[i]
//   resample_residual   (pmc.c:1605-1626): multinomial multiplicities of the residual weights N w_i - int(N w_i)
// The reference uses gsl_ran_discrete (alias method) and gsl_ran_multinomial on a sequential mt19937 stream; as for the
// draw, the GPU version is distributionally equivalent, not stream-identical: inverse-CDF sampling on a deterministic
// prefix sum with one Philox counter per draw (validated statistically, tests/test_gpu_resample.py).
#include "pmc_kernels.h"

namespace pmcb200 {

namespace {

constexpr int SCAN_T = 256, SCAN_PER = 16, SCAN_CHUNK = SCAN_T * SCAN_PER;

// mode 0: w_i ; mode 1: residual N w_i - (int)(N w_i) (pmc.c:1619); non-finite or negative entries count as zero
__device__ __forceinline__ double scan_value(const double *__restrict__ w, long i, long N, int mode) {
  if (i >= N) return 0.0;
  double v = w[i];
  if (mode == 1) { const double nw = (double)N * v; v = nw - (double)(int)nw; }
  return (v > 0.0 && isfinite(v)) ? v : 0.0;
}

// block-local inclusive prefix sums (fixed summation structure -> deterministic) and the block totals
__global__ void __launch_bounds__(SCAN_T)
k_scan_block(long N, const double *__restrict__ w, int mode, double *__restrict__ cdf, double *__restrict__ blocksum) {
  __shared__ double swarp[SCAN_T / 32];
  const long base = (long)blockIdx.x * SCAN_CHUNK + (long)threadIdx.x * SCAN_PER;
  double v[SCAN_PER];
  double run = 0.0;
#pragma unroll
  for (int u = 0; u < SCAN_PER; ++u) { run += scan_value(w, base + u, N, mode); v[u] = run; }
  // exclusive scan of the thread totals: warp shuffle scan, then the warp totals
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) swarp[wid] = inc;
  __syncthreads();
  double woff = 0.0;
  for (int k = 0; k < wid; ++k) woff += swarp[k];
  const double off = woff + (inc - run);
#pragma unroll
  for (int u = 0; u < SCAN_PER; ++u)
    if (base + u < N) cdf[base + u] = off + v[u];
  if (threadIdx.x == SCAN_T - 1) blocksum[blockIdx.x] = off + run;
}
// exclusive scan of the block totals in place (one thread: at most N / 4096 entries, summed in index order)
__global__ void k_scan_sums(int nb, double *__restrict__ blocksum) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double run = 0.0;
    for (int b = 0; b < nb; ++b) { const double t = blocksum[b]; blocksum[b] = run; run += t; }
  }
}
__global__ void k_scan_add(long N, double *__restrict__ cdf, const double *__restrict__ blocksum) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < N) cdf[i] += blocksum[i / SCAN_CHUNK];
}

// one categorical draw per thread: first index whose cumulative weight exceeds u * total
__global__ void k_resample(long N, long ndraw, const double *__restrict__ cdf, uint64_t seed, uint32_t stream, int mode,
                           unsigned long long *__restrict__ isample, unsigned int *__restrict__ counts) {
  const long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (j >= ndraw) return;
  const Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  const uint4 r = ph((uint32_t)j, (uint32_t)((uint64_t)j >> 32), 0x52455341u /* "RESA" */, stream);
  const double u = u01_co(r.x, r.y) * cdf[N - 1];
  long lo = 0, hi = N - 1;  // invariant: answer in [lo, hi]
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  if (mode == 0) isample[j] = (unsigned long long)lo;
  else atomicAdd(&counts[lo], 1u);
}

// ---- clip_weights (pmc.c:1134-1188) ----------------------------------------------------------------------------------
// order-preserving 64-bit key of a double (total order of IEEE-754 values, NaNs excluded by the callers)
__device__ __forceinline__ unsigned long long dkey(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// number of ok samples j >= first whose weight key is >= key (one step of the bitwise search for the n-th largest)
__global__ void k_count_key_ge(long N, long first, const double *__restrict__ w, const short *__restrict__ flg,
                               unsigned long long key, unsigned long long *cnt) {
  unsigned int n = 0;
  for (long j = first + blockIdx.x * (long)blockDim.x + threadIdx.x; j < N; j += (long)gridDim.x * blockDim.x) {
    const double v = w[j];
    if (flg[j] != 0 && v == v && dkey(v) >= key) ++n;
  }
  const unsigned int c = (unsigned int)warp_sum_ll(n);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, (unsigned long long)c);
}
// ok samples with weight >= T are zeroed and flagged out, the others are summed (block partials, fixed order)
__global__ void k_clip_apply(long N, double T, double *__restrict__ w, short *__restrict__ flg, double *__restrict__ partial,
                             unsigned long long *cnt) {
  double acc = 0.0;
  unsigned int n = 0;
  for (long j = blockIdx.x * (long)blockDim.x + threadIdx.x; j < N; j += (long)gridDim.x * blockDim.x) {
    if (flg[j] == 0) continue;
    const double v = w[j];
    if (v >= T) { w[j] = 0.0; flg[j] = 0; ++n; }
    else acc += v;
  }
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const double sw = warp_sum(acc);
  if (lane == 0) sh[wid] = sw;
  __syncthreads();
  if (wid == 0) {
    const double t = warp_sum(lane < nw ? sh[lane] : 0.0);
    if (lane == 0) partial[blockIdx.x] = t;
  }
  const unsigned int c = (unsigned int)warp_sum_ll(n);
  if (lane == 0 && c) atomicAdd(cnt, (unsigned long long)c);
}
__global__ void k_scale_ok(long N, double S, double *__restrict__ w, const short *__restrict__ flg) {
  for (long j = blockIdx.x * (long)blockDim.x + threadIdx.x; j < N; j += (long)gridDim.x * blockDim.x)
    if (flg[j] != 0) w[j] /= S;
}

}  // namespace

void launch_count_key_ge(long N, long first, const double *w, const short *flg, unsigned long long key,
                         unsigned long long *cnt, int nblocks, cudaStream_t s) {
  k_count_key_ge<<<nblocks, 256, 0, s>>>(N, first, w, flg, key, cnt);
  count_launch();
}
void launch_clip_apply(long N, double T, double *w, short *flg, double *partial, unsigned long long *cnt, int nblocks,
                       cudaStream_t s) {
  k_clip_apply<<<nblocks, 256, 0, s>>>(N, T, w, flg, partial, cnt);
  count_launch();
}
void launch_scale_ok(long N, double S, double *w, const short *flg, int nblocks, cudaStream_t s) {
  k_scale_ok<<<nblocks, 256, 0, s>>>(N, S, w, flg);
  count_launch();
}

// component indices (size_t in pmc_simu, values < 256) as bytes: what the host sink sends over PCIe (1 instead of 8 bytes per
// sample); the host widens them back into the caller's size_t array
__global__ void k_narrow_idx(long N, const unsigned long long *__restrict__ in, unsigned char *__restrict__ out) {
  const long stride = (long)gridDim.x * blockDim.x;
  const long n8 = N >> 3;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += stride) {
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(in + 8 * g);
    unsigned long long packed = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const ulonglong2 v = p[q];
      packed |= (v.x & 0xffull) << (16 * q) | (v.y & 0xffull) << (16 * q + 8);
    }
    reinterpret_cast<unsigned long long *>(out)[g] = packed;
  }
  for (long i = 8 * n8 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) out[i] = (unsigned char)in[i];
}
void launch_narrow_idx(long N, const unsigned long long *in, unsigned char *out, int sm_count, cudaStream_t s) {
  if (N <= 0) return;
  long nb = (N / 8 + 255) / 256;
  if (nb > (long)sm_count * 8) nb = (long)sm_count * 8;
  if (nb < 1) nb = 1;
  k_narrow_idx<<<(int)nb, 256, 0, s>>>(N, in, out);
  count_launch();
}

long scan_blocks(long N) { return (N + SCAN_CHUNK - 1) / SCAN_CHUNK; }

// cdf[i] = sum_{j <= i} value_j, value = weights (mode 0) or residual weights (mode 1); blocksum: scan_blocks(N) doubles
void launch_weight_cdf(long N, const double *w, int mode, double *cdf, double *blocksum, cudaStream_t s) {
  if (N <= 0) return;
  const int nb = (int)scan_blocks(N);
  k_scan_block<<<nb, SCAN_T, 0, s>>>(N, w, mode, cdf, blocksum);
  k_scan_sums<<<1, 32, 0, s>>>(nb, blocksum);
  k_scan_add<<<(int)((N + 255) / 256), 256, 0, s>>>(N, cdf, blocksum);
  count_launch(3);
}
void launch_resample(long N, long ndraw, const double *cdf, uint64_t seed, uint32_t stream, int mode,
                     unsigned long long *isample, unsigned int *counts, cudaStream_t s) {
  if (N <= 0 || ndraw <= 0) return;
  k_resample<<<(int)((ndraw + 255) / 256), 256, 0, s>>>(N, ndraw, cdf, seed, stream, mode, isample, counts);
  count_launch();
}

}  // namespace pmcb200
