// FP64 pipe micro-benchmarks for B200 (sm_100a): sustained DFMA, DMMA (mma.sync f64),
// DFMA+DMMA issued together, and the double-precision transcendentals the PMC path uses.
// These are the roofline denominators for the PMC iteration (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

template <int NACC, int KK>
__global__ void __launch_bounds__(256) k_dmma16(double* out, double av, double bv) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = av + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = bv + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (KK == 4) { double a2[2] = {a[0], a[1]}; dmma1684(c[i], a2, b[0]); }
      if (KK == 8) { double a4[4] = {a[0], a[1], a[2], a[3]}; double b2[2] = {b[0], b[1]}; dmma1688(c[i], a4, b2); }
      if (KK == 16) dmma16816(c[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// Same warp issues NF DFMA and NM DMMA(8x8x4) per iteration: tells whether the two share a pipe.
template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mixed(double* out, double a, double b) {
  double acc[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < NM; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
      if (i < NF) acc[i] = fma(acc[i], a, b);
      if (i < NM) dmma884(c0[i], c1[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// warp-specialised mix: even warps DFMA, odd warps DMMA
template <int NACC>
__global__ void __launch_bounds__(256) k_split(double* out, double a, double b) {
  const int w = threadIdx.x >> 5;
  double acc[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  if (w & 1) {
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma884(acc[i], c1[i], a, b);
    }
  } else {
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

enum { F_EXP, F_LOG, F_SINCOSPI, F_SQRT, F_RCP, F_DIV, F_RSQRT, F_I2D };
template <int F>
__global__ void __launch_bounds__(256) k_func(double* out, double x0) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = x0 + 1e-3 * threadIdx.x + 0.1 * i;
  double s = 0;
  for (int it = 0; it < ITERS / 16; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double r;
      if (F == F_EXP) r = exp(x[i]);
      if (F == F_LOG) r = log(x[i]);
      if (F == F_SINCOSPI) { double sn, cs; sincospi(x[i], &sn, &cs); r = sn + cs; }
      if (F == F_SQRT) r = sqrt(x[i]);
      if (F == F_RCP) r = 1.0 / x[i];
      if (F == F_DIV) r = x0 / x[i];
      if (F == F_RSQRT) r = rsqrt(x[i]);
      if (F == F_I2D) r = (double)(long long)__double_as_longlong(x[i]);
      s += r;
      x[i] += 1e-6;
    }
  }
  if (s == 123.456) out[0] = s;
}

__global__ void k_fill(double2* p, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = make_double2(v, v);
}
__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

template <typename L>
float time_ms(L launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, clk);
  double* out; CK(cudaMalloc(&out, 1024));
  const int blocks = sms * 8, threads = 256;
  const double nthr = (double)blocks * threads;
  auto rep = [&](const char* name, float ms, double flops_per_thread) {
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"tflops\": %.3f}\n", name, ms, nthr * flops_per_thread / ms * 1e-9);
    fflush(stdout);
  };
  rep("dfma_acc4", time_ms([&] { k_dfma<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 2.0 * ITERS * 4);
  rep("dfma_acc8", time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 2.0 * ITERS * 8);
  rep("dfma_acc16", time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 2.0 * ITERS * 16);
  // DMMA m8n8k4: 8*8*4 FMA per warp-instr = 16 flop per thread per instr
  rep("dmma_m8n8k4_acc4", time_ms([&] { k_dmma884<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 16.0 * ITERS * 4);
  rep("dmma_m8n8k4_acc8", time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 16.0 * ITERS * 8);
  rep("dmma_m16n8k4_acc4", time_ms([&] { k_dmma16<4, 4><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 32.0 * ITERS * 4);
  rep("dmma_m16n8k8_acc4", time_ms([&] { k_dmma16<4, 8><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 64.0 * ITERS * 4);
  rep("dmma_m16n8k16_acc4", time_ms([&] { k_dmma16<4, 16><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 128.0 * ITERS * 4);
  rep("dmma_m16n8k16_acc8", time_ms([&] { k_dmma16<8, 16><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), 128.0 * ITERS * 8);
  // mixed in one warp
  rep("mixed_8dfma_8dmma(flops=dfma+dmma)", time_ms([&] { k_mixed<8, 8><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), (2.0 + 16.0) * ITERS * 8);
  rep("mixed_8dfma_2dmma", time_ms([&] { k_mixed<8, 2><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), (2.0 * 8 + 16.0 * 2) * ITERS);
  rep("mixed_8dfma_1dmma", time_ms([&] { k_mixed<8, 1><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), (2.0 * 8 + 16.0 * 1) * ITERS);
  rep("split_warps_dfma_dmma", time_ms([&] { k_split<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); }), (2.0 + 16.0) * 0.5 * ITERS * 8);
  auto repf = [&](const char* name, float ms) {
    double calls = nthr * (ITERS / 16) * 4;
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"gcalls_per_s\": %.2f, \"dfma_equiv_per_call\": null}\n", name, ms, calls / ms * 1e-6);
    fflush(stdout);
  };
  repf("exp", time_ms([&] { k_func<F_EXP><<<blocks, threads>>>(out, 0.37); }));
  repf("log", time_ms([&] { k_func<F_LOG><<<blocks, threads>>>(out, 0.37); }));
  repf("sincospi", time_ms([&] { k_func<F_SINCOSPI><<<blocks, threads>>>(out, 0.37); }));
  repf("sqrt", time_ms([&] { k_func<F_SQRT><<<blocks, threads>>>(out, 0.37); }));
  repf("rcp", time_ms([&] { k_func<F_RCP><<<blocks, threads>>>(out, 0.37); }));
  repf("div", time_ms([&] { k_func<F_DIV><<<blocks, threads>>>(out, 0.37); }));
  repf("rsqrt", time_ms([&] { k_func<F_RSQRT><<<blocks, threads>>>(out, 0.37); }));
  repf("i2d", time_ms([&] { k_func<F_I2D><<<blocks, threads>>>(out, 0.37); }));
  // HBM: fill (write-only) and copy over 4 GiB
  size_t n2 = (size_t)1 << 28;  // double2 elements: 4 GiB
  double2 *a, *b; CK(cudaMalloc(&a, n2 * 16)); CK(cudaMalloc(&b, n2 * 16));
  float ms = time_ms([&] { k_fill<<<sms * 16, 512>>>(a, n2, 1.0); });
  printf("{\"bench\": \"hbm_fill\", \"ms\": %.4f, \"gbs\": %.1f}\n", ms, n2 * 16.0 / ms * 1e-6);
  ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, n2); });
  printf("{\"bench\": \"hbm_copy\", \"ms\": %.4f, \"gbs\": %.1f}\n", ms, 2.0 * n2 * 16.0 / ms * 1e-6);
  // sustained DFMA for ~2 s to see clocks under load
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  int nl = 0;
  for (; nl < 400; ++nl) k_dfma<8><<<blocks, threads>>>(out, 1.0000001, 1e-9);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("{\"bench\": \"dfma_acc8_sustained\", \"ms\": %.2f, \"tflops\": %.3f}\n", ms, nl * nthr * 2.0 * ITERS * 8 / ms * 1e-9);
  return 0;
}
