// pmc_dev.cuh — device-side building blocks shared by the generic and the fused PMC kernels.
// Constants are the reference's, including the truncated ln2pi (maths_base.h:49) that parity depends on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace pmcb200 {

constexpr double kLn2Pi = 1.837877066409;    // maths_base.h:49 (truncated on purpose)
constexpr double kPi = 3.141592653589793;    // pmc.h:45 / mvdens.h:70
constexpr double kDynMax = 500.0;            // pmc.h:47
constexpr double kPerpZero = 1e-20;          // pmc.c:1040
constexpr int kMinCount = 20;                // pmc.h:71
constexpr double kNegInit = -1.0e30;         // pmc.c:535

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based -------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ b;
      c3 = lo0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// 53-bit uniform in [0,1) (what gsl_rng_uniform promises for ranindx, mvdens.c:775) and in (0,1]
__device__ __forceinline__ double u01_co(uint32_t hi, uint32_t lo) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double u01_oc(uint32_t hi, uint32_t lo) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return ((double)(v >> 11) + 1.0) * (1.0 / 9007199254740992.0);
}

// one Philox call -> two independent N(0,1) (Box-Muller, full double precision in the radius and the angle)
__device__ __forceinline__ void box_muller(const uint4 r, double &g0, double &g1) {
  const double u1 = u01_oc(r.x, r.y);
  const double u2 = u01_co(r.z, r.w);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  g0 = rad * c;
  g1 = rad * s;
}

// mvdens.c:766-782: first index whose cumulative weight is >= z (the last index if none)
__device__ __forceinline__ int ranindx(const double *cw, int K, double z) {
  int index = 0;
  while (cw[index] < z && index < K - 1) index++;
  return index;
}

// examples/target/sampleTarget.c:89-110, same operation order
template <typename XT>
__device__ __forceinline__ double bent_gauss(const XT &x, int d, double b, double sig1) {
  double res = 0.0, x2 = x[0] * x[0];
  res += x2 / sig1;
  x2 -= sig1;
  x2 *= b;
  x2 += x[1];
  res += x2 * x2;
  for (int i = 2; i < d; ++i) res += x[i] * x[i];
  return -res / 2.;
}

// Tail of mvdens_log_pdf (mvdens.c:361-383) given q = |L^-1 (x-mu)|^2, same association order.
// logdet = log(detL); Student-t: gam = lngamma(.5(n+p)), norm = lngamma(.5n) + .5 p log(n PI), both from the host.
__device__ __forceinline__ double logpdf_from_q(double q, int d, double logdet, int df, double gam, double norm) {
  if (df != -1) {
    const double n = (double)df;
    const double r = 0.5 * (n + d) * log(1.0 + (1.0 / n) * q);
    return gam - (norm + logdet + r);
  }
  return -0.5 * (d * kLn2Pi + q) - logdet;
}

// device view of a mixture in reference layout (global memory), used by the generic kernels
struct MixDev {
  int K, d;
  const double *alpha;   // [K] wght
  const double *cw;      // [K] cumulative weights (mvdens.c:748-750)
  const double *mean;    // [K*d]
  const double *L;       // [K*d*d] row-major lower Cholesky factors
  const double *logdet;  // [K] log(detL)
  const double *gam;     // [K]
  const double *norm;    // [K]
  const int *df;         // [K]
};

struct BoxDev {
  int nfaces, d;
  const double *faces;  // [nfaces*(d+1)]
};

// parabox.c:112-136, literal (term-by-term subtraction)
template <typename XT>
__device__ __forceinline__ bool in_box(const BoxDev &bx, const XT &x) {
  for (int f = 0; f < bx.nfaces; ++f) {
    const double *c = bx.faces + (size_t)f * (bx.d + 1);
    double thr = c[bx.d];
    for (int j = 0; j < bx.d; ++j) thr = __dsub_rn(thr, __dmul_rn(c[j], x[j]));  // no FMA contraction: decisions stay bit-exact
    if (thr < 0) return false;
  }
  return true;
}


// ---- ziggurat normal sampler (Marsaglia & Tsang 2000 in Doornik's ZIGNOR form, 4096 layers) -------------------------
// A Box-Muller pair costs ~115 DFMA issue slots in double precision (log + sqrt + sincospi, profiles/r01_fp64_peak.md);
// the ziggurat's fast path is one table lookup, one DADD, one DMUL and one compare.  What matters on a GPU is how often a
// WARP leaves the fast path, because one slow lane stalls the other 31: a 10-D sample is 320 candidates per warp, so the
// 256-layer version (1.5 % slow candidates) spent 27 % of the draw kernel in the slow path, the 1024-layer version
// (0.43 %: 75 % of the warps still had a slow lane, ncu profiles/r02_ncu.txt) 20 %, and 4096 layers (0.12 %) bring it
// to 32 % of the warps.  The 64 KB table lives in shared memory.
// Table t[i] = (x_i, x_{i+1}/x_i), x_0 = V/f(R), x_1 = R, x_C = 0 (built on the host by the C-ABI layer).
// R and V solve the closing condition V/x_{C-1} + f(x_{C-1}) = 1 (tools/zig_constants.py, 60-digit bisection; it
// reproduces the published constants for C = 128 and 256).
constexpr int kZigBits = 12;
constexpr int kZigLayers = 1 << kZigBits;
constexpr double kZigR = 4.3859450348713045;
constexpr double kZigV = 0.0003061541032784645;

// candidate from 64 random bits (hi:lo): layer = lo[0..11], sign = lo[12], lo[13..15] are left for the caller,
// |u| = the top 48 bits as a fraction in [0,1)
__device__ __forceinline__ double zig_u(uint32_t lo, uint32_t hi) {
  return __hiloint2double(0x3ff00000 | (hi >> 12), ((hi << 20) | (lo >> 12)) & ~15u) - 1.0;
}
__device__ __forceinline__ bool zig_fast(uint32_t lo, uint32_t hi, const double2 *zt, double &x) {
  const double2 t = zt[lo & (kZigLayers - 1)];
  const double u = zig_u(lo, hi);
  const double ax = u * t.x;
  // the sign bit of the candidate goes straight into the (non-negative) product: a shift and one three-input logic
  // operation instead of a negation on the FP64 pipe and two selects
  x = __hiloint2double(__double2hiint(ax) ^ ((lo << (31 - kZigBits)) & 0x80000000u), __double2loint(ax));
  return u < t.y;
}

// exp() of two arguments at once (defined below)
template <int N>
__device__ __forceinline__ void exp_batch(const double (&x)[N], double (&out)[N]);

// Slow path for a candidate whose rectangle test failed (wedge or tail), then fresh candidates until one is accepted.
// Fresh randomness comes from Philox calls (cb | call) with call = 64, 65, ... of this sample/attempt.
__device__ __forceinline__ double zig_finish(uint32_t lo, uint32_t hi, const double2 *zt, const Philox &ph, uint32_t ilo,
                                             uint32_t ihi, uint32_t cb, uint32_t iter, uint32_t &call) {
  for (int guard = 0; guard < 1000; ++guard) {
    const uint32_t layer = lo & (kZigLayers - 1);
    const double2 t = zt[layer];
    const double u = zig_u(lo, hi);
    const bool neg = (lo & kZigLayers) != 0;
    const double ax = u * t.x;
    if (u < t.y) return neg ? -ax : ax;
    if (layer == 0) {  // tail beyond R (Marsaglia 1964)
      for (int g2 = 0; g2 < 1000; ++g2) {
        const uint4 r = ph(ilo, ihi, cb | call, iter);
        ++call;
        const double xx = log(u01_oc(r.x, r.y)) / kZigR, yy = log(u01_oc(r.z, r.w));
        if (-2.0 * yy >= xx * xx) return neg ? xx - kZigR : kZigR - xx;
      }
      return neg ? -kZigR : kZigR;
    }
    {  // wedge between the rectangle and the density
      const double xi = t.x, xi1 = t.x * t.y;
      const double arg[2] = {-0.5 * (xi * xi - ax * ax), -0.5 * (xi1 * xi1 - ax * ax)};
      double f[2];
      exp_batch<2>(arg, f);
      const uint4 r = ph(ilo, ihi, cb | call, iter);
      ++call;
      if (f[1] + u01_co(r.x, r.y) * (f[0] - f[1]) < 1.0) return neg ? -ax : ax;
      lo = r.z;
      hi = r.w;  // fresh candidate
    }
  }
  return 0.0;
}

// D standard normals for (sample, attempt): Philox calls 1..ceil(D/2) feed the fast path, calls >= 64 the slow path.
// spare receives the 3 unused bits of every candidate (3*D bits, candidate j in bits 3j..3j+2).
template <int D>
__device__ __forceinline__ void zig_normals(const Philox &ph, uint32_t ilo, uint32_t ihi, uint32_t cb, uint32_t iter,
                                            const double2 *zt, double (&g)[D + 1], uint32_t &spare) {
  uint32_t fail = 0, slo = 0, shi = 0;
  spare = 0;
#pragma unroll
  for (int j = 0; j < D; j += 2) {
    const uint4 r = ph(ilo, ihi, cb | (1 + j / 2), iter);
    spare |= ((r.x >> (kZigBits + 1)) & 7u) << ((3 * j) & 31);
    if (!zig_fast(r.x, r.y, zt, g[j])) {
      if (!fail) { slo = r.x; shi = r.y; }
      fail |= 1u << j;
    }
    if (j + 1 < D) {
      spare |= ((r.z >> (kZigBits + 1)) & 7u) << ((3 * (j + 1)) & 31);
      if (!zig_fast(r.z, r.w, zt, g[j + 1])) {
        if (!fail) { slo = r.z; shi = r.w; }
        fail |= 1u << (j + 1);
      }
    }
  }
  uint32_t call = 64;
  bool first = true;
  while (fail) {  // 0.3 % of the candidates; the bits of the first failed slot were kept, later ones are regenerated
    const int j = __ffs(fail) - 1;
    fail &= fail - 1;
    uint32_t lo = slo, hi = shi;
    if (!first) {
      const uint4 r = ph(ilo, ihi, cb | (1 + (j >> 1)), iter);
      lo = (j & 1) ? r.z : r.x;
      hi = (j & 1) ? r.w : r.y;
    }
    first = false;
    const double v = zig_finish(lo, hi, zt, ph, ilo, ihi, cb, iter, call);
#pragma unroll
    for (int jj = 0; jj < D; ++jj)
      if (jj == j) g[jj] = v;
  }
}

// ---- table-driven exp and log ---------------------------------------------------------------------------------------
// exp(x) for x <= 709; arguments below -700 (and NaN) return exp(-700) ~ 1e-304: callers only add the result to sums
// whose largest term is >= 1, and catch non-finite inputs separately.  x = n ln2/64 + r, |r| <= ln2/128; 2^(n/64) from
// the table, exp(r) - 1 = r P4(r) (truncation 3.5e-17 relative).
__device__ __forceinline__ double exp_tab(double x, const double *tab) {
  constexpr double kMagic = 6755399441055744.0;
  constexpr double kInv = 92.332482616893657;
  constexpr double kHi = 6.93147180369123816490e-01 / 64.0, kLo = 1.90821492927058770002e-10 / 64.0;
  // clamp on the high word (x <= 709 by contract: "below -700" is then an unsigned minimum of the sign-magnitude bits, which
  // leaves the low word alone, i.e. clamps to a value in [-700.0000002, -700]; an FP64 max costs a DSETP on the FP64 pipe plus
  // NaN plumbing)
  const double xc = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC085E000u), __double2loint(x));  // one VIMNMX: |x| <= 700.00.. for negative x
  const double t = fma(xc, kInv, kMagic);
  const int ni = __double2loint(t);
  const double n = t - kMagic;
  double r = fma(n, -kHi, xc);
  r = fma(n, -kLo, r);
  const double T = tab[ni & 63];
  double p = fma(r, 1.0 / 120, 1.0 / 24);
  p = fma(p, r, 1.0 / 6);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  const double s = p * r;
  const double v = fma(T, s, T);
  return __hiloint2double(__double2hiint(v) + ((ni >> 6) << 20), __double2loint(v));
}

// log(x) for positive normal x: x = 2^e m, m in [1,2); r = m rc_j - 1 with |r| < 1/256 from the table indexed by the
// top 7 mantissa bits; log1p(r) to degree 6 (truncation 2e-18 absolute)
__device__ __forceinline__ double log_tab(double x, const double2 *tab) {
  constexpr double kLn2 = 0.693147180559945309417;
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 tb = tab[(hi >> 13) & 127];
  const double r = fma(m, tb.x, -1.0);
  double p = fma(r, -1.0 / 6, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3);
  p = fma(p, r, -0.5);
  const double lp = fma(r * r, p, r);
  const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;  // (double)e
  return fma(ed, kLn2, tb.y + lp);
}

// ---- batched exp ---------------------------------------------------------------------------------------------------
// exp() for N independent arguments <= 709 with the Horner recurrences interleaved: each Taylor coefficient is
// materialised once per batch instead of once per call and the N chains hide each other's latency.
// Cody-Waite reduction x = n ln2 + r, |r| <= ln2/2; degree-13 Taylor polynomial (truncation 4e-18 relative);
// 2^n applied by adding n to the exponent field.  Arguments below -708 (and -inf) give 0, NaN gives NaN; the caller
// guarantees x <= 709.
template <int N>
__device__ __forceinline__ void exp_batch(const double (&x)[N], double (&out)[N]) {
  constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to nearest integer in the low word
  constexpr double kL2e = 1.4426950408889634074, kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;
  double r[N], p[N];
  int ni[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double xc = fmax(x[k], -800.0);  // keeps n in range (NaN -> -800, restored below)
    const double t = fma(xc, kL2e, kMagic);
    ni[k] = __double2loint(t);
    const double n = t - kMagic;
    r[k] = fma(n, -kLn2Lo, fma(n, -kLn2Hi, xc));
  }
  constexpr double c[14] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
                            1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = c[13];
#pragma unroll
  for (int j = 12; j >= 0; --j)
#pragma unroll
    for (int k = 0; k < N; ++k) p[k] = fma(p[k], r[k], c[j]);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double v = __hiloint2double(__double2hiint(p[k]) + (ni[k] << 20), __double2loint(p[k]));
    // x >= -708 -> v ; x < -708 or -inf -> 0 ; NaN -> NaN (x + v is NaN, and the comparison below is false for NaN)
    out[k] = (x[k] >= -708.0) ? v : ((x[k] != x[k]) ? x[k] : 0.0);
  }
}
__device__ __forceinline__ double exp_one(double x) {
  double a[1] = {x}, o[1];
  exp_batch<1>(a, o);
  return o[0];
}

// warp reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max_nan_ignoring(double v) {  // v > m semantics: NaN never wins
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    if (t > v || (v != v)) v = t;
  }
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// per-sample Philox stream of the draw: counter = (global sample index, attempt << 12 | call, iteration)
struct DrawStream {
  Philox ph;
  uint32_t i_lo, i_hi, iter, attempt, call;
  __device__ __forceinline__ uint4 next() {
    const uint4 r = ph(i_lo, i_hi, (attempt << 12) | call, iter);
    ++call;
    return r;
  }
};

static __device__ double chisq_mt(DrawStream &st, double nu) {  // 2*Gamma(nu/2,1), Marsaglia-Tsang (as gsl_ran_chisq)
  double a = nu * 0.5, boost = 1.0;
  if (a < 1.0) {
    const uint4 r = st.next();
    boost = pow(u01_oc(r.x, r.y), 1.0 / a);
    a += 1.0;
  }
  const double dd = a - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(dd);
  for (int guard = 0; guard < 1000; ++guard) {
    double x, v, g1;
    do {
      box_muller(st.next(), x, g1);
      v = 1.0 + c * x;
    } while (v <= 0);
    v = v * v * v;
    const uint4 r = st.next();
    const double u = u01_oc(r.x, r.y);
    if (u < 1 - 0.0331 * x * x * x * x) return 2.0 * dd * v * boost;
    if (log(u) < 0.5 * x * x + dd * (1 - v + log(v))) return 2.0 * dd * v * boost;
  }
  return 2.0 * dd * boost;
}


}  // namespace pmcb200
