// pmc_hostio.cpp — liberrorio's file and memory helpers (pmctools/include/io.h:34-75) that target likelihoods and drivers
// written for pmclib call beside the PMC path: line counting, the white-space separated text readers behind
// read_double_list & co., the raw binary readers / writers, resize_err, the run-time stamps.  Host code only; same
// names, argument meaning, return values and error codes (err_io, err_allocate, io_file) as the reference, checked by
// linking one C program against this library and against the compiled reference (tests/test_cpu_abi.py).
// fopen_err / malloc_err / calloc_err live in pmc_dropin.cpp.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include <sys/stat.h>
#include "pmc_dropin_internal.h"

namespace {
constexpr int kErrIo = -101;   // err_io  (errorlist.h:144)
constexpr int kErrAlloc = -102;  // err_allocate (errorlist.h:145)
constexpr int kIoFile = -202;  // io_file (io.h:26)
constexpr size_t kChunk = 524288;  // records allocated at a time and longest accepted line (io.c)

#define IO_FAIL(code, ...) (*err = newErrorVA((code), __func__, (char *)"", (char *)__VA_ARGS__))
#define IO_FWD(ret) do { if (isError(*err)) { *err = newError(forwardErr, (char *)__func__, (char *)"", nullptr, *err); return ret; } } while (0)
}  // namespace

extern "C" {

// io.c:13-27: number of newline characters
unsigned int numberoflines(const char *name, error **err) {
  FILE *f = fopen_err(name, "r", err);
  IO_FWD(0);
  unsigned int n = 0;
  for (int c = fgetc(f); c != EOF; c = fgetc(f)) n += (c == '\n');
  fclose(f);
  return n;
}

// io.c:30-51: lines that do not start with '#'; *ncomment = lines that do
unsigned int numberoflines_comments(const char *name, unsigned int *ncomment, error **err) {
  FILE *f = fopen_err(name, "r", err);
  IO_FWD(0);
  unsigned int n = 0;
  *ncomment = 0;
  int prev = '\n';
  for (int c = fgetc(f); c != EOF; c = fgetc(f)) {
    if (c == '#' && prev == '\n') { --n; ++*ncomment; }
    if (c == '\n') ++n;
    prev = c;
  }
  fclose(f);
  return n;
}

// io.c:57-69: strtod at *p, the cursor moves behind the number; the return value is that of the reference's expression
// (&endptr != p) && (*p = endptr), i.e. non-zero whenever the new cursor is non-null
int read_double(char **p, double *x) {
  char *end = nullptr;
  *x = strtod(*p, &end);
  *p = end;
  return end != nullptr;
}
int read_int(char **p, int *x) {
  char *end = nullptr;
  *x = (int)strtod(*p, &end);
  *p = end;
  return end != nullptr;
}

// io.c:72-97: "nx ny" header, then nx*ny numbers, x fastest
double *readASCII(char *filename, int *fnx, int *fny, error **err) {
  FILE *f = fopen_err(filename, "r", err);
  IO_FWD(nullptr);
  int n = fscanf(f, "%d %d\n", fnx, fny);
  if (n != 2) { IO_FAIL(kIoFile, "%d elements read from file %s, expected %d", *err, nullptr, n, filename, 2); fclose(f); return nullptr; }
  double *mat = (double *)malloc_err((size_t)(*fnx) * (size_t)(*fny) * sizeof(double), err);
  if (isError(*err)) { fclose(f); IO_FWD(nullptr); }
  for (int j = 0; j < *fny; ++j)
    for (int i = 0; i < *fnx; ++i) {
      double v;
      n = fscanf(f, "%le\n", &v);
      if (n != 1) { IO_FAIL(kIoFile, "%d elements read from file %s, expected %d", *err, nullptr, n, filename, 1); fclose(f); return nullptr; }
      mat[i + (size_t)j * (*fnx)] = v;
    }
  fclose(f);
  return mat;
}

// io.c:136-145.  The reference's function reallocates, raises err_allocate when the INPUT pointer was NULL or the
// reallocation failed, and returns NULL in every case (callers cannot use it to grow a buffer); kept, so that code written
// against it behaves the same.
void *realloc_err(void *ptr, size_t sz, error **err) {
  void *res = realloc(ptr, sz);
  if (ptr == nullptr) IO_FAIL(kErrAlloc, "Cannot reallocate %zu bytes", *err, nullptr, sz);
  if (res == nullptr) IO_FAIL(kErrAlloc, "Cannot reallocate %zu bytes", *err, nullptr, sz);
  return nullptr;
}

// io.c:151-174: a new block of sz_new bytes holding the first min(sz_orig, sz_new) bytes of orig (orig == NULL or
// sz_orig == 0: just the new block); free_orig == 1 releases orig
void *resize_err(void *orig, size_t sz_orig, size_t sz_new, int free_orig, error **err) {
  void *res = malloc_err(sz_new, err);
  IO_FWD(nullptr);
  if (orig == nullptr || sz_orig == 0) return res;
  memcpy(res, orig, sz_orig < sz_new ? sz_orig : sz_new);
  if (free_orig == 1) free(orig);
  return res;
}

// io.c:176-201: one line without its comment ('#', '!' or ';' to the end of the line); returns the stored length + 1
size_t read_line(FILE *fp, char *buff, int bmax, error **err) {
  size_t i = 0;
  bool comment = false;
  while ((long)i < (long)bmax - 1) {
    const int c = fgetc(fp);
    if (feof(fp) || (char)c == '\n') { buff[i] = '\0'; return i + 1; }
    if (comment) continue;
    if (c == '#' || c == '!' || c == ';') { comment = true; continue; }
    buff[i++] = (char)c;
  }
  *err = newErrorVA(kErrIo, __func__, (char *)"", (char *)"One line of file is longer than the max size (%d)", *err, nullptr, bmax);
  return (size_t)-1;
}

// io.c:236-323: white-space separated records of scanf format prs and size sz.  *n > 0 on input: exactly that many records
// are expected (reading stops there, fewer is an error); *n == 0: all of them.  *n = records read, *nlines = lines that
// held at least one record (comment and blank lines do not count).
void *read_any_list_count(const char *fnm, size_t *n, char *prs, size_t sz, size_t *nlines, error **err) {
  const bool clip = *n > 0;
  size_t cap = clip ? *n : kChunk;
  char *res = (char *)malloc_err(cap * sz, err);
  IO_FWD(nullptr);
  FILE *fp = fopen_err(fnm, "r", err);
  IO_FWD(nullptr);
  std::vector<char> line(kChunk);
  size_t count = 0;
  *nlines = 0;
  bool full = false;
  while (!full && feof(fp) == 0) {
    const size_t len = read_line(fp, line.data(), (int)kChunk, err);
    if (isError(*err)) { IO_FAIL(kErrIo, "Reading file '%s'", *err, nullptr, fnm); return nullptr; }
    const size_t before = count;
    size_t j = 0;
    while (j < len) {
      while (j < len && line[j] != '\0' && isspace((unsigned char)line[j])) ++j;
      if (j == len || line[j] == '\0') break;
      const size_t start = j;
      while (j < len && line[j] != '\0' && !isspace((unsigned char)line[j])) ++j;
      line[j++] = '\0';
      if (sscanf(line.data() + start, prs, res + count * sz) != 1) {
        IO_FAIL(kErrIo, "Cannot read file '%s' at index %d", *err, nullptr, fnm, (int)count);
        return nullptr;
      }
      ++count;
      if (count == cap) {
        if (clip) { full = true; break; }
        res = (char *)resize_err(res, cap * sz, cap * 2 * sz, 1, err);
        IO_FWD(nullptr);
        cap *= 2;
      }
    }
    if (count > before) ++*nlines;
  }
  fclose(fp);
  if (clip && count != cap) {
    IO_FAIL(kErrIo, "Not enough data in file '%s' (got %d entries, expected %d)\n", *err, nullptr, fnm, (int)count, (int)cap);
    return nullptr;
  }
  res = (char *)resize_err(res, cap * sz, count * sz, 1, err);
  IO_FWD(nullptr);
  *n = count;
  return res;
}
void *read_any_list(const char *fnm, size_t *n, char *prs, size_t sz, error **err) {
  size_t nlines = 0;
  void *res = read_any_list_count(fnm, n, prs, sz, &nlines, err);
  IO_FWD(nullptr);
  return res;
}
void *read_any_vector(const char *fnm, size_t n, char *prs, size_t sz, error **err) {
  size_t nn = n;
  void *res = read_any_list(fnm, &nn, prs, sz, err);
  IO_FWD(nullptr);
  return res;
}

#define TYPED_READERS(T, name, fmt)                                                                         \
  T *read_##name##_vector(const char *fnm, size_t n, error **err) {                                         \
    T *res = (T *)read_any_vector(fnm, n, (char *)fmt, sizeof(T), err);                                     \
    IO_FWD(nullptr);                                                                                        \
    return res;                                                                                             \
  }                                                                                                         \
  T *read_##name##_list(const char *fnm, size_t *n, error **err) {                                          \
    T *res = (T *)read_any_list(fnm, n, (char *)fmt, sizeof(T), err);                                       \
    IO_FWD(nullptr);                                                                                        \
    return res;                                                                                             \
  }
TYPED_READERS(double, double, "%lg")
TYPED_READERS(float, float, "%g")
TYPED_READERS(int, int, "%d")
TYPED_READERS(long, long, "%ld")
#undef TYPED_READERS

// io.c:389-436: raw bytes
void *read_bin_vector(const char *fnm, size_t n, error **err) {
  FILE *f = fopen_err(fnm, "r", err);
  IO_FWD(nullptr);
  void *res = malloc_err(n, err);
  if (isError(*err)) { fclose(f); IO_FWD(nullptr); }
  const size_t got = fread(res, 1, n, f);
  fclose(f);
  if (got != n) {
    IO_FAIL(kIoFile, "File '%s' does not contain enough data (Expected %d bytes, got %d)", *err, nullptr, fnm, (int)n, (int)got);
    return nullptr;
  }
  return res;
}
void *read_bin_list(const char *fnm, size_t *n, error **err) {
  struct stat st;
  if (stat(fnm, &st) != 0) { IO_FAIL(kIoFile, "File '%s' cannot be stated", *err, nullptr, fnm); return nullptr; }
  *n = (size_t)st.st_size;
  void *res = read_bin_vector(fnm, *n, err);
  IO_FWD(nullptr);
  return res;
}
void write_bin_vector(void *buf, const char *fnm, size_t n, error **err) {
  FILE *f = fopen_err(fnm, "w", err);
  IO_FWD();
  const size_t put = fwrite(buf, 1, n, f);
  if (put != n) {
    IO_FAIL(kIoFile, "Cannot write data in file '%s' (Expected %d bytes, got %d)", *err, nullptr, fnm, (int)n, (int)put);
    return;
  }
  fclose(f);
  fprintf(stderr, "Saved in %s\n", fnm);
}

// io.c:438-442: cut at the LAST newline
void chomp(char *x) {
  char *p = strrchr(x, '\n');
  if (p) *p = '\0';
}

// io.c:490-502
int is_directory(const char *path) {
  struct stat st;
  return (stat(path, &st) != -1 && S_ISDIR(st.st_mode)) ? 1 : 0;
}

// io.c:444-488: run-time stamps of the drivers
time_t start_time(FILE *FOUT) {
  time_t t0;
  time(&t0);
  fprintf(FOUT, "Started at %s\n", ctime(&t0));
  fflush(FOUT);
  return t0;
}
void end_time(time_t t_start, FILE *FOUT) {
  time_t t1;
  time(&t1);
  fprintf(FOUT, "Ended at %s", ctime(&t1));
  const double diff = difftime(t1, t_start);
  const int s = (int)diff;
  fprintf(FOUT, "Computation time %.0fs (= %dd %dh %dm %ds)\n", diff, s / 86400, (s % 86400) / 3600, (s % 3600) / 60, s % 60);
}
clock_t start_clock(FILE *FOUT) {
  time_t t0;
  time(&t0);
  const clock_t c0 = clock();
  fprintf(FOUT, "Started at %s\n", ctime(&t0));
  fflush(FOUT);
  return c0;
}
void end_clock(clock_t c_start, FILE *FOUT) {
  const clock_t c1 = clock();
  time_t t1;
  time(&t1);
  fprintf(FOUT, "Ended at %s", ctime(&t1));
  float diff = ((float)c1 - (float)c_start) / 1000000.0f;
  const int s = (int)diff;
  fprintf(FOUT, "Computation time %.3fs (= %dd %dh %dm %ds %dms)\n", diff, s / 86400, (s % 86400) / 3600, (s % 3600) / 60, s % 60,
          (int)(1000 * diff) - 1000 * s);
}

}  // extern "C"
