/* liberrorio's io.h helpers (pmc_hostio.cpp): the same program is linked against libpmcb200.so and against the compiled
 * reference; stdout must be identical (tests/test_cpu_abi.py).  argv[1] = a scratch directory. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pmclib_abi.h"

static void code(const char *what, error **e) {
  printf("%s -> error %d\n", what, isError(*e) ? getErrorValue(*e) : 0);
  purgeError(e);
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  char f1[4096], f2[4096], f3[4096], f4[4096], f5[4096];
  snprintf(f1, sizeof f1, "%s/list.txt", argv[1]);
  snprintf(f2, sizeof f2, "%s/bin.dat", argv[1]);
  snprintf(f3, sizeof f3, "%s/missing.txt", argv[1]);
  snprintf(f4, sizeof f4, "%s/bad.txt", argv[1]);
  snprintf(f5, sizeof f5, "%s/mat.txt", argv[1]);
  FILE *f = fopen(f1, "w");
  fprintf(f, "# header line\n1.5 2.5\t-3e2   4 # trailing comment\n\n   \n! another comment\n5 6;7 8\n9\n# end\n10 11 12");
  fclose(f);
  f = fopen(f4, "w");
  fprintf(f, "1 2 x3 4\n");
  fclose(f);
  f = fopen(f5, "w");
  fprintf(f, "3 2\n1\n2\n3\n4\n5\n6\n");
  fclose(f);
  error *e = initError();

  unsigned int nc = 0;
  printf("numberoflines %u\n", numberoflines(f1, &e)); code("numberoflines", &e);
  printf("numberoflines_comments %u", numberoflines_comments(f1, &nc, &e)); printf(" comments %u\n", nc); code("numberoflines_comments", &e);
  numberoflines(f3, &e); code("numberoflines missing", &e);

  size_t n = 0, nl = 0;
  double *d = read_any_list_count(f1, &n, "%lg", sizeof(double), &nl, &e); code("read_any_list_count", &e);
  printf("n %zu lines %zu:", n, nl);
  for (size_t i = 0; d && i < n; ++i) printf(" %g", d[i]);
  printf("\n");
  free(d);
  n = 0;
  d = read_double_list(f1, &n, &e); code("read_double_list", &e);
  printf("double list n %zu last %g\n", n, d ? d[n - 1] : -1.0);
  free(d);
  d = read_double_vector(f1, 5, &e); code("read_double_vector 5", &e);
  if (d) printf("vector: %g %g %g %g %g\n", d[0], d[1], d[2], d[3], d[4]);
  free(d);
  d = read_double_vector(f1, 50, &e); code("read_double_vector 50 (too many)", &e);
  printf("vector 50 is %s\n", d ? "non-null" : "null");
  n = 0;
  int *iv = read_int_list(f1, &n, &e); code("read_int_list", &e);
  printf("int list n %zu:", n);
  for (size_t i = 0; iv && i < n; ++i) printf(" %d", iv[i]);
  printf("\n");
  free(iv);
  n = 0;
  long *lv = read_long_list(f1, &n, &e); code("read_long_list", &e);
  printf("long list n %zu first %ld\n", n, lv ? lv[0] : -1L);
  free(lv);
  n = 0;
  float *fv = read_float_list(f1, &n, &e); code("read_float_list", &e);
  printf("float list n %zu third %g\n", n, fv ? fv[2] : -1.0f);
  free(fv);
  n = 0;
  d = read_double_list(f4, &n, &e); code("read_double_list bad token", &e);
  printf("bad is %s\n", d ? "non-null" : "null");
  n = 0;
  d = read_double_list(f3, &n, &e); code("read_double_list missing", &e);

  FILE *fp = fopen(f1, "r");
  char buf[64];
  for (int k = 0; k < 4; ++k) {
    size_t len = read_line(fp, buf, 64, &e);
    printf("read_line %zu |%s|\n", len, buf);
  }
  code("read_line", &e);
  size_t len = read_line(fp, buf, 4, &e); printf("short buffer returns %ld\n", (long)len); code("read_line short buffer", &e);
  fclose(fp);

  double payload[4] = {1.0, -2.5, 3.25, 1e300};
  write_bin_vector(payload, f2, sizeof payload, &e); code("write_bin_vector", &e);
  double *back = read_bin_vector(f2, sizeof payload, &e); code("read_bin_vector", &e);
  printf("bin equal %d\n", back && memcmp(back, payload, sizeof payload) == 0);
  free(back);
  n = 0;
  back = read_bin_list(f2, &n, &e); code("read_bin_list", &e);
  printf("bin list bytes %zu equal %d\n", n, back && memcmp(back, payload, sizeof payload) == 0);
  free(back);
  back = read_bin_vector(f2, 1000, &e); code("read_bin_vector too long", &e);
  back = read_bin_list(f3, &n, &e); code("read_bin_list missing", &e);

  char *grown = malloc(4);
  memcpy(grown, "abc", 4);
  char *g2 = resize_err(grown, 4, 16, 1, &e); code("resize_err grow", &e);
  printf("resized |%s|\n", g2);
  char *g3 = resize_err(g2, 16, 2, 0, &e); code("resize_err shrink", &e);
  printf("shrunk %c%c\n", g3[0], g3[1]);
  free(g2); free(g3);
  void *z = resize_err(NULL, 0, 8, 1, &e); code("resize_err from null", &e);
  printf("from null is %s\n", z ? "non-null" : "null");
  free(z);
  resize_err(NULL, 0, 0, 1, &e); code("resize_err zero", &e);
  void *r = realloc_err(malloc(8), 32, &e); printf("realloc_err returns %s\n", r ? "non-null" : "null"); code("realloc_err", &e);
  r = realloc_err(NULL, 32, &e); code("realloc_err null input", &e);

  char txt[] = "line one\nline two\n";
  chomp(txt);
  printf("chomp |%s|\n", txt);
  char num[] = "  3.75 42abc";
  char *p = num;
  double x; int iv2;
  int r1 = read_double(&p, &x);
  int r2 = read_int(&p, &iv2);
  printf("read_double %d %g read_int %d %d rest |%s|\n", r1, x, r2, iv2, p);
  printf("is_directory %d %d %d\n", is_directory(argv[1]), is_directory(f1), is_directory(f3));
  int nx = 0, ny = 0;
  double *m = readASCII(f5, &nx, &ny, &e); code("readASCII", &e);
  if (m) printf("readASCII %d x %d: %g %g %g %g %g %g\n", nx, ny, m[0], m[1], m[2], m[3], m[4], m[5]);
  free(m);
  return 0;
}
