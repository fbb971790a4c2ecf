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

}  // namespace pmcb200
