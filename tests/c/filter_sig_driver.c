/* filter_importance_weight_sig (pmc.c:772-827) on an already normalised sample: the same program is linked once against
 * libpmcb200.so and once against the compiled reference (oracle/_ref/libpmcref.so); the outputs must be identical
 * (tests/test_cpu_abi.py).  No device is needed: the weights are not on log scale, so nothing is normalised. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "pmclib_abi.h"

int main(int argc, char **argv) {
  const long N = argc > 1 ? atol(argv[1]) : 2000;
  const double nsig = argc > 2 ? atof(argv[2]) : 0.002;
  error *e = initError();
  pmc_simu *p = pmc_simu_init(N, 3, &e);
  if (isError(e)) { printError(stderr, e); return 1; }
  unsigned long long s = 88172645463325252ull;
  double sum = 0;
  for (long i = 0; i < N; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const double u = (double)(s >> 11) / 9007199254740992.0;
    p->weights[i] = pow(u, 6.0) + 1e-9;  /* a few dominant weights */
    sum += p->weights[i];
    p->flg[i] = (i % 17 == 3) ? 0 : 1;
    for (int a = 0; a < 3; ++a) p->X[i * 3 + a] = (double)i + 0.25 * a;
  }
  for (long i = 0; i < N; ++i) p->weights[i] /= sum;
  p->isLog = 0;
  filter_importance_weight_sig(p, nsig, &e);
  if (isError(e)) { printError(stderr, e); return 2; }
  printf("isLog %d maxW %.17g\n", p->isLog, p->maxW);
  for (long i = 0; i < N; ++i) printf("%d %.17g\n", (int)p->flg[i], p->weights[i]);
  return 0;
}
