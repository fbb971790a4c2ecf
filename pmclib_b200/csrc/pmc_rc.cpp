// pmc_rc.cpp — the config / driver layer of the PMC path without Lua: pmclib's readConf API (pmctools/src/readConf.c) on a
// small evaluator of the Lua subset parameter files are written in, and the pmc_rc functions that turn such a file into
// objects (pmclib/src/pmc_rc.c: target, proposal incl. the draw_from initialisation, parabox, generator, pmc_simu).
// Host code by nature: everything that touches samples still goes through the kernels — the draw_from
// initialisation calls mvdens_ran (device draw), the objects it builds are the ones pmc_simu_pmc_step runs on the GPU.
// Citations are file:line in CosmoStat/pmclib.
#include <dlfcn.h>
#include <strings.h>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../include/pmclib_rc.h"

namespace {
constexpr int pmc_dimension = -6004, pmc_negWeight = -6005, pmc_undef = -6008, pmc_io = -6010, dist_base = -6700, dist_undef = -6701, pb_base = -6500;

#define RCERR(e, code, ...) e = newErrorVA(code, __func__, (char *)"(pmc_rc.cpp)", (char *)__VA_ARGS__)
#define RCFWD(e, ret)                                                       \
  do {                                                                      \
    if (isError(e)) {                                                       \
      e = newError(forwardErr, (char *)__func__, (char *)"", nullptr, e);   \
      return ret;                                                           \
    }                                                                       \
  } while (0)

// ---- values ----------------------------------------------------------------------------------------------------------
struct Node;
using NodeP = std::shared_ptr<Node>;
struct Node {
  enum Kind { NIL, NUM, STR, BOOL, TABLE } kind = NIL;
  double num = 0;
  std::string str;
  bool b = false;
  std::map<std::string, NodeP> fields;  // named entries
  std::map<long, NodeP> arr;            // integer keys (Lua arrays start at 1)
  long length() const {                 // the border Lua's # operator returns for a sequence
    long n = 0;
    while (arr.count(n + 1) && arr.at(n + 1)->kind != NIL) ++n;
    return n;
  }
};
NodeP mk(Node::Kind k) { auto n = std::make_shared<Node>(); n->kind = k; return n; }

// ---- the evaluator ---------------------------------------------------------------------------------------------------
struct Parser {
  const char *p, *end;
  NodeP globals;
  std::string err;
  int line = 1;
  bool fail(const std::string &m) { if (err.empty()) err = "line " + std::to_string(line) + ": " + m; return false; }
  void skip() {
    for (;;) {
      while (p < end && isspace((unsigned char)*p)) { if (*p == '\n') ++line; ++p; }
      if (p + 1 < end && p[0] == '-' && p[1] == '-') {
        if (p + 3 < end && p[2] == '[' && p[3] == '[') {  // --[[ block ]]
          p += 4;
          while (p + 1 < end && !(p[0] == ']' && p[1] == ']')) { if (*p == '\n') ++line; ++p; }
          p = (p + 1 < end) ? p + 2 : end;
        } else {
          while (p < end && *p != '\n') ++p;
        }
        continue;
      }
      break;
    }
  }
  bool isname(char c, bool first) { return isalpha((unsigned char)c) || c == '_' || (!first && isdigit((unsigned char)c)); }
  bool name(std::string &out) {
    skip();
    if (p >= end || !isname(*p, true)) return false;
    const char *s = p;
    while (p < end && isname(*p, false)) ++p;
    out.assign(s, p);
    return true;
  }
  bool lit(char c) { skip(); if (p < end && *p == c) { ++p; return true; } return false; }
  bool peek(char c) { skip(); return p < end && *p == c; }
  bool string(std::string &out) {
    skip();
    if (p >= end || (*p != '"' && *p != '\'')) return false;
    const char q = *p++;
    out.clear();
    while (p < end && *p != q) {
      if (*p == '\\' && p + 1 < end) {
        ++p;
        switch (*p) { case 'n': out += '\n'; break; case 't': out += '\t'; break; default: out += *p; }
        ++p;
      } else {
        if (*p == '\n') ++line;
        out += *p++;
      }
    }
    if (p >= end) return fail("unterminated string");
    ++p;
    return true;
  }
  // an lvalue / reference: name { '.' name | '[' expr ']' }; create = make missing tables on the way (assignment)
  bool path(NodeP &holder, std::string &fkey, long &ikey, bool &is_int, bool create) {
    std::string nm;
    if (!name(nm)) return fail("name expected");
    holder = globals;
    fkey = nm;
    is_int = false;
    for (;;) {
      skip();
      const bool dot = p < end && *p == '.' && !(p + 1 < end && p[1] == '.');
      const bool br = p < end && *p == '[';
      if (!dot && !br) break;
      NodeP cur = is_int ? (holder->arr.count(ikey) ? holder->arr[ikey] : nullptr) : (holder->fields.count(fkey) ? holder->fields[fkey] : nullptr);
      if (!cur || cur->kind != Node::TABLE) {
        if (!create) return fail("attempt to index a non-table value (" + fkey + ")");
        cur = mk(Node::TABLE);
        if (is_int) holder->arr[ikey] = cur; else holder->fields[fkey] = cur;
      }
      holder = cur;
      ++p;
      if (dot) {
        if (!name(nm)) return fail("name expected after '.'");
        fkey = nm;
        is_int = false;
      } else {
        NodeP k;
        if (!expr(k)) return false;
        if (!lit(']')) return fail("']' expected");
        if (k->kind == Node::NUM) { is_int = true; ikey = (long)k->num; }
        else if (k->kind == Node::STR) { is_int = false; fkey = k->str; }
        else return fail("bad table key");
      }
    }
    return true;
  }
  bool table(NodeP &out) {
    out = mk(Node::TABLE);
    long next = 1;
    for (;;) {
      if (lit('}')) return true;
      const char *save = p;
      const int sl = line;
      std::string nm;
      NodeP v;
      if (peek('[')) {
        ++p;
        NodeP k;
        if (!expr(k) || !lit(']') || !lit('=') || !expr(v)) return fail("bad [key] = value entry");
        if (k->kind == Node::NUM) out->arr[(long)k->num] = v; else if (k->kind == Node::STR) out->fields[k->str] = v; else return fail("bad table key");
      } else if (name(nm) && peek('=') && !(p + 1 < end && p[1] == '=')) {
        ++p;
        if (!expr(v)) return false;
        out->fields[nm] = v;
      } else {
        p = save;
        line = sl;
        if (!expr(v)) return false;
        out->arr[next++] = v;
      }
      if (!lit(',') && !lit(';')) {
        if (lit('}')) return true;
        return fail("',' or '}' expected in table constructor");
      }
    }
  }
  bool primary(NodeP &out) {
    skip();
    if (p >= end) return fail("expression expected");
    if (*p == '{') { ++p; return table(out); }
    if (*p == '(') { ++p; if (!expr(out)) return false; if (!lit(')')) return fail("')' expected"); return true; }
    if (*p == '"' || *p == '\'') { out = mk(Node::STR); return string(out->str); }
    if (isdigit((unsigned char)*p) || (*p == '.' && p + 1 < end && isdigit((unsigned char)p[1]))) {
      char *e = nullptr;
      out = mk(Node::NUM);
      out->num = strtod(p, &e);
      if (e == p) return fail("bad number");
      p = e;
      return true;
    }
    if (*p == '-') { ++p; NodeP v; if (!unary(v)) return false; if (v->kind != Node::NUM) return fail("unary minus on a non-number"); out = mk(Node::NUM); out->num = -v->num; return true; }
    const char *save = p;
    std::string nm;
    if (name(nm)) {
      if (nm == "true" || nm == "false") { out = mk(Node::BOOL); out->b = nm == "true"; return true; }
      if (nm == "nil") { out = mk(Node::NIL); return true; }
      p = save;
      NodeP holder;
      std::string fkey;
      long ikey = 0;
      bool is_int = false;
      if (!path(holder, fkey, ikey, is_int, false)) return false;
      if (peek('(')) return fail("function calls are not part of the parameter-file subset (" + fkey + ")");
      NodeP v = is_int ? (holder->arr.count(ikey) ? holder->arr[ikey] : nullptr) : (holder->fields.count(fkey) ? holder->fields[fkey] : nullptr);
      out = v ? v : mk(Node::NIL);  // tables are shared by reference, as in Lua
      return true;
    }
    return fail("unexpected character in expression");
  }
  bool unary(NodeP &out) { return primary(out); }
  bool power(NodeP &out) {
    if (!unary(out)) return false;
    if (lit('^')) { NodeP r; if (!power(r)) return false; if (out->kind != Node::NUM || r->kind != Node::NUM) return fail("arithmetic on a non-number"); NodeP v = mk(Node::NUM); v->num = std::pow(out->num, r->num); out = v; }
    return true;
  }
  bool term(NodeP &out) {
    if (!power(out)) return false;
    for (;;) {
      skip();
      if (p < end && (*p == '*' || *p == '/')) {
        const char op = *p++;
        NodeP r;
        if (!power(r)) return false;
        if (out->kind != Node::NUM || r->kind != Node::NUM) return fail("arithmetic on a non-number");
        NodeP v = mk(Node::NUM);
        v->num = op == '*' ? out->num * r->num : out->num / r->num;
        out = v;
      } else break;
    }
    return true;
  }
  bool expr(NodeP &out) {
    if (!term(out)) return false;
    for (;;) {
      skip();
      if (p + 1 < end && p[0] == '.' && p[1] == '.') {  // string concatenation
        p += 2;
        NodeP r;
        if (!term(r)) return false;
        auto tos = [](const NodeP &n) { if (n->kind == Node::STR) return n->str; char b[64]; snprintf(b, sizeof b, "%.14g", n->num); return std::string(b); };
        NodeP v = mk(Node::STR);
        v->str = tos(out) + tos(r);
        out = v;
      } else if (p < end && (*p == '+' || (*p == '-' && !(p + 1 < end && p[1] == '-')))) {
        const char op = *p++;
        NodeP r;
        if (!term(r)) return false;
        if (out->kind != Node::NUM || r->kind != Node::NUM) return fail("arithmetic on a non-number");
        NodeP v = mk(Node::NUM);
        v->num = op == '+' ? out->num + r->num : out->num - r->num;
        out = v;
      } else break;
    }
    return true;
  }
  bool chunk() {
    for (;;) {
      skip();
      if (p >= end) return true;
      if (*p == ';') { ++p; continue; }
      const char *save = p;
      std::string nm;
      if (name(nm) && nm == "local") { /* scope does not matter for a parameter file */ } else p = save;
      NodeP holder;
      std::string fkey;
      long ikey = 0;
      bool is_int = false;
      if (!path(holder, fkey, ikey, is_int, true)) return false;
      if (!lit('=')) return fail("'=' expected (only assignments are part of the parameter-file subset)");
      NodeP v;
      if (!expr(v)) return false;
      if (is_int) holder->arr[ikey] = v; else holder->fields[fkey] = v;
    }
  }
};

struct RcState {
  NodeP globals = mk(Node::TABLE);
  std::vector<void *> mem;
};

confFile *root_of(confFile *rc) { return rc->alias ? (confFile *)rc->alias : rc; }
RcState *state_of(confFile *rc) { return (RcState *)root_of(rc)->L; }

std::string full_key(confFile *rc, const char *key) {
  std::string k = rc->alias_prefix;
  if (key && key[0]) {
    if (!k.empty() && key[0] != '[') k += ".";
    k += key;
  }
  return k;
}

// "a.b[2].c" -> node (nullptr when absent)
NodeP lookup(RcState *st, const std::string &key) {
  NodeP cur = st->globals;
  size_t i = 0;
  if (key.empty()) return cur;
  while (i < key.size()) {
    if (!cur || cur->kind != Node::TABLE) return nullptr;
    if (key[i] == '.') { ++i; continue; }
    if (key[i] == '[') {
      const size_t j = key.find(']', i);
      if (j == std::string::npos) return nullptr;
      const long ix = strtol(key.c_str() + i + 1, nullptr, 10);
      cur = cur->arr.count(ix) ? cur->arr[ix] : nullptr;
      i = j + 1;
    } else {
      size_t j = i;
      while (j < key.size() && key[j] != '.' && key[j] != '[') ++j;
      const std::string nm = key.substr(i, j - i);
      cur = cur->fields.count(nm) ? cur->fields[nm] : nullptr;
      i = j;
    }
  }
  return cur;
}

bool load_text(RcState *st, const char *text, size_t n, std::string &msg) {
  Parser ps{text, text + n, st->globals};
  if (!ps.chunk()) { msg = ps.err; return false; }
  return true;
}

NodeP need(confFile *rc, const char *key, Node::Kind kind, error **err, const char *what) {
  const std::string fk = full_key(rc, key);
  NodeP n = lookup(state_of(rc), fk);
  if (!n || n->kind == Node::NIL) { RCERR(*err, rc_bad_key, "ReadConf(%s): cannot find parameter %s", *err, nullptr, root_of(rc)->fnm, fk.c_str()); return nullptr; }
  if (n->kind != kind) { RCERR(*err, rc_bad_type, "ReadConf(%s): parameter %s is not %s", *err, nullptr, root_of(rc)->fnm, fk.c_str(), what); return nullptr; }
  return n;
}

// ---- MT19937 behind GSL's gsl_rng layout (gsl-1.14 rng/mt.c: same seeding, same stream) --------------------------------
struct MtState { unsigned long mt[624]; int mti; };
void mt_set(void *vs, unsigned long s) {
  MtState *st = (MtState *)vs;
  if (s == 0) s = 4357;
  st->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < 624; ++i) st->mt[i] = (1812433253UL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  st->mti = 624;
}
unsigned long mt_get(void *vs) {
  MtState *st = (MtState *)vs;
  unsigned long *mt = st->mt;
  auto magic = [](unsigned long y) { return (y & 1) ? 0x9908b0dfUL : 0UL; };
  if (st->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; ++kk) { unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ magic(y); }
    for (; kk < 623; ++kk) { unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ magic(y); }
    { unsigned long y = (mt[623] & 0x80000000UL) | (mt[0] & 0x7fffffffUL); mt[623] = mt[396] ^ (y >> 1) ^ magic(y); }
    st->mti = 0;
  }
  unsigned long k = mt[st->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  return k;
}
double mt_get_double(void *vs) { return mt_get(vs) / 4294967296.0; }
const gsl_rng_type mt_type = {"mt19937", 0xffffffffUL, 0, sizeof(MtState), &mt_set, &mt_get, &mt_get_double};

}  // namespace
extern "C" { static void *rcinit_mix_mvdens_mpi3(confFile *rc, char *root, error **err); }
namespace {
// INIT_FROM_RC (pmc_rc.h:31-63): initfunc or rcinit_<name>, from libpath if given, else from the process, else built in
rcinit_smthng *builtin_init(const char *fn) {
  if (!strcmp(fn, "rcinit_bentGauss")) return &rcinit_bentGauss;
  if (!strcmp(fn, "rcinit_mix_mvdens")) return (rcinit_smthng *)&rcinit_mix_mvdens;
  if (!strcmp(fn, "rcinit_mix_mvdens_mpi")) return &rcinit_mix_mvdens_mpi3;
  if (!strcmp(fn, "rcinit_combine_distributions")) return &rcinit_combine_distributions;
  if (!strcmp(fn, "rcinit_add_gaussian_prior")) return &rcinit_add_gaussian_prior;
  return nullptr;
}
}  // namespace

extern "C" {

// ====================================================================================================================
// readConf (pmctools/src/readConf.c)
// ====================================================================================================================
confFile *rc_open(char *fnm, error **err) {
  FILE *f = fopen(fnm, "rb");
  if (!f) { RCERR(*err, rc_lua_error, "Cannot open parameter file %s", *err, nullptr, fnm); return nullptr; }
  std::string text;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  fclose(f);
  confFile *rc = (confFile *)calloc(1, sizeof(confFile));
  if (!rc) { RCERR(*err, -102, "Cannot allocate", *err, nullptr); return nullptr; }
  RcState *st = new RcState();
  rc->L = st;
  snprintf(rc->fnm, sizeof rc->fnm, "%s", fnm);
  std::string msg;
  if (!load_text(st, text.data(), text.size(), msg)) {
    RCERR(*err, rc_lua_error, "ReadConf(%s): %s", *err, nullptr, fnm, msg.c_str());
    delete st;
    free(rc);
    return nullptr;
  }
  return rc;
}
confFile *rc_init(char *fnm, error **err) { return rc_open(fnm, err); }

// readConf.c:982-1060: -e <chunk> / -e=<chunk> / --extra[=]<chunk>: parameter-file statements evaluated after the file
void rc_do_args(confFile *rc, int argc, char **argv, char shortopt, char *longopt, error **err) {
  if (rc->alias) { RCERR(*err, rc_bad_type, "Does not work with alias", *err, nullptr); return; }
  const size_t ln = strlen(longopt);
  for (int i = 0; i < argc; ++i) {
    const char *a = argv[i], *chunk = nullptr;
    if (a[0] != '-') continue;
    if (a[1] == shortopt && a[2] == '\0') chunk = i + 1 < argc ? argv[i + 1] : nullptr;
    else if (a[1] == shortopt && a[2] == '=') chunk = a + 3;
    else if (!strncmp(a + 1, longopt, ln) && a[ln + 1] == '\0') chunk = i + 1 < argc ? argv[i + 1] : nullptr;
    else if (!strncmp(a + 1, longopt, ln) && a[ln + 1] == '=') chunk = a + ln + 2;
    else if (a[1] == '-' && !strncmp(a + 2, longopt, ln) && a[ln + 2] == '\0') chunk = i + 1 < argc ? argv[i + 1] : nullptr;
    else if (a[1] == '-' && !strncmp(a + 2, longopt, ln) && a[ln + 2] == '=') chunk = a + ln + 3;
    if (!chunk) continue;
    std::string msg;
    if (!load_text(state_of(rc), chunk, strlen(chunk), msg)) { RCERR(*err, rc_lua_error, "ReadConf: cannot evaluate '%s': %s", *err, nullptr, chunk, msg.c_str()); return; }
  }
}
confFile *rc_init_from_args(int argc, char **argv, error **err) {  // readConf.c:1195-1225
  if (argc == 1) { RCERR(*err, -1, "not enough parameters on the command line", *err, nullptr); return nullptr; }
  printf("%s: reading parfile %s\n", __func__, argv[argc - 1]);
  confFile *rc = rc_open(argv[argc - 1], err);
  RCFWD(*err, nullptr);
  rc_do_args(rc, argc, argv, 'e', (char *)"extra", err);
  RCFWD(*err, nullptr);
  const long verb = rc_safeget_integer(rc, (char *)"rc.verbose", 0, err);
  RCFWD(*err, nullptr);
  if (verb > 0) rc->display = stdout;
  if (verb < 0) rc->display = stderr;
  return rc;
}
void rc_toggle_display(confFile *rc, FILE *display, error **err) { (void)err; rc->display = display; }

void rc_close(confFile **prc) {  // readConf.c:188-210: an alias only frees itself
  confFile *rc = *prc;
  if (!rc) return;
  if (!rc->alias) {
    RcState *st = (RcState *)rc->L;
    for (void *m : st->mem) free(m);
    delete st;
  }
  free(rc);
  *prc = nullptr;
}
confFile *rc_alias(confFile *rc, char *prefix, error **err) {  // readConf.c:140-182
  confFile *a = (confFile *)calloc(1, sizeof(confFile));
  if (!a) { RCERR(*err, -102, "Cannot allocate", *err, nullptr); return nullptr; }
  a->alias = root_of(rc);
  snprintf(a->alias_prefix, sizeof a->alias_prefix, "%s", full_key(rc, prefix).c_str());
  a->display = rc->display;
  return a;
}
void rc_add_mem(confFile *rc, void *ptr, error **err) { (void)err; state_of(rc)->mem.push_back(ptr); }
void *rc_add_malloc(confFile *rc, size_t sz, error **err) {
  void *p = malloc(sz ? sz : 1);
  if (!p) { RCERR(*err, -102, "Cannot allocate %ld bytes", *err, nullptr, (long)sz); return nullptr; }
  state_of(rc)->mem.push_back(p);
  return p;
}
void rc_free(confFile *rc, void **smthng, error **err) {
  (void)err;
  auto &mem = state_of(rc)->mem;
  for (size_t i = 0; i < mem.size(); ++i)
    if (mem[i] == *smthng) { free(mem[i]); mem.erase(mem.begin() + i); break; }
  *smthng = nullptr;
}

int rc_has_key(confFile *rc, char *key, error **err) {
  (void)err;
  NodeP n = lookup(state_of(rc), full_key(rc, key));
  return (n && n->kind != Node::NIL) ? 1 : 0;
}
double rc_get_real(confFile *rc, char *key, error **err) {
  NodeP n = need(rc, key, Node::NUM, err, "a number");
  if (!n) return 0;
  if (rc->display) fprintf(rc->display, "ReadConf(%s) %s = %g\n", root_of(rc)->fnm, full_key(rc, key).c_str(), n->num);
  return n->num;
}
long rc_get_integer(confFile *rc, char *key, error **err) {
  NodeP n = need(rc, key, Node::NUM, err, "a number");
  if (!n) return 0;
  if (rc->display) fprintf(rc->display, "ReadConf(%s) %s = %ld\n", root_of(rc)->fnm, full_key(rc, key).c_str(), (long)n->num);
  return (long)n->num;
}
char *rc_get_string(confFile *rc, char *key, error **err) {
  NodeP n = need(rc, key, Node::STR, err, "a string");
  if (!n) return nullptr;
  char *s = (char *)rc_add_malloc(rc, n->str.size() + 1, err);
  if (!s) return nullptr;
  memcpy(s, n->str.c_str(), n->str.size() + 1);
  if (rc->display) fprintf(rc->display, "ReadConf(%s) %s = \"%s\"\n", root_of(rc)->fnm, full_key(rc, key).c_str(), s);
  return s;
}
double rc_safeget_real(confFile *rc, char *key, double safeguard, error **err) { return rc_has_key(rc, key, err) ? rc_get_real(rc, key, err) : safeguard; }
long rc_safeget_integer(confFile *rc, char *key, long safeguard, error **err) { return rc_has_key(rc, key, err) ? rc_get_integer(rc, key, err) : safeguard; }
char *rc_safeget_string(confFile *rc, char *key, char *safeguard, error **err) { return rc_has_key(rc, key, err) ? rc_get_string(rc, key, err) : safeguard; }
int rc_is_array(confFile *rc, char *key, error **err) {
  (void)err;
  NodeP n = lookup(state_of(rc), full_key(rc, key));
  return (n && n->kind == Node::TABLE) ? 1 : 0;
}
size_t rc_array_size(confFile *rc, char *key, error **err) {
  NodeP n = need(rc, key, Node::TABLE, err, "an array");
  return n ? (size_t)n->length() : 0;
}
size_t rc_get_real_array(confFile *rc, char *key, double **parray, error **err) {
  NodeP n = need(rc, key, Node::TABLE, err, "an array");
  if (!n) return 0;
  const long sz = n->length();
  double *a = (double *)rc_add_malloc(rc, sizeof(double) * (sz > 0 ? sz : 1), err);
  if (!a) return 0;
  for (long i = 0; i < sz; ++i) {
    const NodeP &v = n->arr[i + 1];
    if (v->kind != Node::NUM) { RCERR(*err, rc_bad_type, "ReadConf(%s): %s[%ld] is not a number", *err, nullptr, root_of(rc)->fnm, full_key(rc, key).c_str(), i + 1); return 0; }
    a[i] = v->num;
    if (rc->display) fprintf(rc->display, "ReadConf(%s) %s[%ld] = %g\n", root_of(rc)->fnm, full_key(rc, key).c_str(), i, a[i]);
  }
  *parray = a;
  return (size_t)sz;
}
size_t rc_get_integer_array(confFile *rc, char *key, long **parray, error **err) {
  NodeP n = need(rc, key, Node::TABLE, err, "an array");
  if (!n) return 0;
  const long sz = n->length();
  long *a = (long *)rc_add_malloc(rc, sizeof(long) * (sz > 0 ? sz : 1), err);
  if (!a) return 0;
  for (long i = 0; i < sz; ++i) {
    const NodeP &v = n->arr[i + 1];
    if (v->kind != Node::NUM) { RCERR(*err, rc_bad_type, "ReadConf(%s): %s[%ld] is not a number", *err, nullptr, root_of(rc)->fnm, full_key(rc, key).c_str(), i + 1); return 0; }
    a[i] = (long)v->num;
    if (rc->display) fprintf(rc->display, "ReadConf(%s) %s[%ld] = %ld\n", root_of(rc)->fnm, full_key(rc, key).c_str(), i, a[i]);
  }
  *parray = a;
  return (size_t)sz;
}
size_t rc_get_integer_array_asint(confFile *rc, char *key, int **parray, error **err) {
  long *l = nullptr;
  const size_t sz = rc_get_integer_array(rc, key, &l, err);
  RCFWD(*err, 0);
  int *a = (int *)rc_add_malloc(rc, sizeof(int) * (sz > 0 ? sz : 1), err);
  if (!a) return 0;
  for (size_t i = 0; i < sz; ++i) a[i] = (int)l[i];
  *parray = a;
  return sz;
}
size_t rc_get_string_array(confFile *rc, char *key, char ***parray, error **err) {
  NodeP n = need(rc, key, Node::TABLE, err, "an array");
  if (!n) return 0;
  const long sz = n->length();
  char **a = (char **)rc_add_malloc(rc, sizeof(char *) * (sz > 0 ? sz : 1), err);
  if (!a) return 0;
  for (long i = 0; i < sz; ++i) {
    const NodeP &v = n->arr[i + 1];
    if (v->kind != Node::STR) { RCERR(*err, rc_bad_type, "ReadConf(%s): %s[%ld] is not a string", *err, nullptr, root_of(rc)->fnm, full_key(rc, key).c_str(), i + 1); return 0; }
    a[i] = (char *)rc_add_malloc(rc, v->str.size() + 1, err);
    if (!a[i]) return 0;
    memcpy(a[i], v->str.c_str(), v->str.size() + 1);
  }
  *parray = a;
  return (size_t)sz;
}
// readConf.c: a scalar is repeated sz times, an array must have sz entries
size_t rc_get_real_asarray(confFile *rc, char *key, double **parray, int sz, error **err) {
  if (!rc_is_array(rc, key, err)) {
    const double v = rc_get_real(rc, key, err);
    RCFWD(*err, 0);
    double *a = (double *)rc_add_malloc(rc, sizeof(double) * (sz > 0 ? sz : 1), err);
    if (!a) return 0;
    for (int i = 0; i < sz; ++i) a[i] = v;
    *parray = a;
    return (size_t)sz;
  }
  const size_t n = rc_get_real_array(rc, key, parray, err);
  RCFWD(*err, 0);
  if ((int)n != sz) { RCERR(*err, rc_size_error, "'%s' in %s does not have the right number elements (got %d expected %d)", *err, nullptr, key, rc->alias_prefix, (int)n, sz); return 0; }
  return n;
}
// readConf.c:697-760: scalar -> scalar * identity; array of sz -> diagonal; array of sz*sz or sz arrays of sz -> full
size_t rc_get_real_assquarematrix(confFile *rc, char *key, double **res, int sz, error **err) {
  const size_t rsz = (size_t)sz * sz;
  double *m = (double *)rc_add_malloc(rc, rsz * sizeof(double), err);
  if (!m) return 0;
  memset(m, 0, rsz * sizeof(double));
  if (!rc_is_array(rc, key, err)) {
    const double v = rc_get_real(rc, key, err);
    RCFWD(*err, 0);
    for (int i = 0; i < sz; ++i) m[(size_t)i * sz + i] = v;
    *res = m;
    return rsz;
  }
  char sub[1000];
  snprintf(sub, sizeof sub, "%s[1]", key);
  if (rc_is_array(rc, sub, err)) {
    for (int i = 0; i < sz; ++i) {
      double *row = nullptr;
      snprintf(sub, sizeof sub, "%s[%d]", key, i + 1);
      const size_t n = rc_get_real_array(rc, sub, &row, err);
      RCFWD(*err, 0);
      if ((int)n != sz) { RCERR(*err, rc_size_error, "'%s[%d]' in %s does not have the right number elements (got %d expected %d)", *err, nullptr, key, i, rc->alias_prefix, (int)n, sz); return 0; }
      memcpy(m + (size_t)i * sz, row, sizeof(double) * sz);
    }
  } else {
    double *v = nullptr;
    const size_t n = rc_get_real_array(rc, key, &v, err);
    RCFWD(*err, 0);
    if ((int)n == sz) for (int i = 0; i < sz; ++i) m[(size_t)i * sz + i] = v[i];
    else if (n == rsz) memcpy(m, v, sizeof(double) * rsz);
    else { RCERR(*err, rc_size_error, "'%s' in %s does not have the right number elements (got %d expected %d or %d)", *err, nullptr, key, rc->alias_prefix, (int)n, sz, sz * sz); return 0; }
  }
  *res = m;
  return rsz;
}

// ====================================================================================================================
// pmc_rc (pmclib/src/pmc_rc.c)
// ====================================================================================================================
static int rc_dist_set_names(confFile *rca, distribution *target, error **err) {  // pmc_rc.c:16-34
  const int hn = rc_has_key(rca, (char *)"parameter_name", err);
  if (hn == 1) {
    char **names = nullptr;
    const int n = (int)rc_get_string_array(rca, (char *)"parameter_name", &names, err);
    RCFWD(*err, 0);
    if (n != target->ndim + target->n_ded) { RCERR(*err, dist_undef, "not enough parameter names (got %d expected %d)", *err, nullptr, n, target->ndim + target->n_ded); return 0; }
    distribution_set_names(target, names, err);
    RCFWD(*err, 0);
  }
  return hn;
}

distribution *init_distribution_from_rc(confFile *rc, char *root, error **err) {
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  // INIT_FROM_RC
  char fname[2000];
  char *fn = rc_safeget_string(rca, (char *)"initfunc", (char *)"", err);
  RCFWD(*err, nullptr);
  if (fn[0] == '\0') {
    char *dn = rc_get_string(rca, (char *)"name", err);
    RCFWD(*err, nullptr);
    snprintf(fname, sizeof fname, "rcinit_%s", dn);
    fn = fname;
  }
  void *handle = nullptr;
  char *so = rc_safeget_string(rca, (char *)"libpath", (char *)"", err);
  RCFWD(*err, nullptr);
  rcinit_smthng *init = nullptr;
  if (so[0] != '\0') {
    handle = dlopen(so, RTLD_NOW | RTLD_GLOBAL);
    if (handle) init = (rcinit_smthng *)dlsym(handle, fn);
  }
  if (!init) init = builtin_init(fn);  // the densities this library implements itself need no plugin
  if (!init) init = (rcinit_smthng *)dlsym(RTLD_DEFAULT, fn);
  if (!init) { RCERR(*err, dl_err, "dl error : cannot find %s (libpath '%s')", *err, nullptr, fn, so); return nullptr; }
  distribution *target = (distribution *)init(rca, (char *)"", err);
  RCFWD(*err, nullptr);
  if (handle) target->dlhandle = handle;
  rc_dist_set_names(rca, target, err);
  RCFWD(*err, nullptr);
  if (rc_has_key(rca, (char *)"conditional", err) == 1) {  // pmc_rc.c:52-78
    double *vdef = nullptr;
    const int nv = (int)rc_get_real_array(rca, (char *)"conditional.value", &vdef, err);
    RCFWD(*err, nullptr);
    int *idef = nullptr;
    int nd = (int)rc_get_integer_array_asint(rca, (char *)"conditional.id", &idef, err);
    if (isError(*err)) {  // names instead of indices
      purgeError(err);
      char **cdef = nullptr;
      nd = (int)rc_get_string_array(rca, (char *)"conditional.id", &cdef, err);
      RCFWD(*err, nullptr);
      if (nv != nd) { RCERR(*err, dist_base, "conditional.id and conditional.value have different sizes (got %d and %d)", *err, nullptr, nd, nv); return nullptr; }
      distribution_set_default_name(target, nd, cdef, vdef, err);
    } else {
      if (nv != nd) { RCERR(*err, dist_base, "conditional.id and conditional.value have different sizes (got %d and %d)", *err, nullptr, nd, nv); return nullptr; }
      distribution_set_default(target, nd, idef, vdef, err);
    }
    RCFWD(*err, nullptr);
  }
  rc_close(&rca);
  return target;
}

void *rcinit_bentGauss(confFile *rc, char *root, error **err) {  // sampleTarget.c:61-84
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  const long ndim = rc_get_integer(rca, (char *)"ndim", err);
  RCFWD(*err, nullptr);
  const double b = rc_get_real(rca, (char *)"bent", err);
  RCFWD(*err, nullptr);
  const double s1 = rc_get_real(rca, (char *)"sig1_square", err);
  RCFWD(*err, nullptr);
  distribution *d = init_bentGauss_distribution((int)ndim, b, s1, err);
  RCFWD(*err, nullptr);
  rc_close(&rca);
  return d;
}

void *rcinit_combine_distributions(confFile *rc, char *root, error **err) {  // pmc_rc.c:84-171
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  const int ndim = (int)rc_get_integer(rca, (char *)"ndim", err);
  RCFWD(*err, nullptr);
  const int nded = (int)rc_safeget_integer(rca, (char *)"nded", 0, err);
  RCFWD(*err, nullptr);
  distribution *comb = combine_distribution_init(ndim, nded, err);
  RCFWD(*err, nullptr);
  const int hn = rc_dist_set_names(rca, comb, err);
  RCFWD(*err, nullptr);
  const int ndist = (int)rc_array_size(rca, (char *)"distributions", err);
  RCFWD(*err, nullptr);
  for (int i = 0; i < ndist; ++i) {
    char dname[500];
    snprintf(dname, sizeof dname, "distributions[%d].distribution", i + 1);
    distribution *addon = init_distribution_from_rc(rca, dname, err);
    RCFWD(*err, nullptr);
    std::vector<int> dim_idx, ded_idx;
    bool doname = true;
    snprintf(dname, sizeof dname, "distributions[%d].dim_idx", i + 1);
    if (rc_has_key(rca, dname, err) == 1) {
      doname = false;
      long *l = nullptr;
      const int n = (int)rc_get_integer_array(rca, dname, &l, err);
      RCFWD(*err, nullptr);
      if (n != addon->ndim) { RCERR(*err, dist_base, "bad size for %s (got %d expected %d)", *err, nullptr, dname, n, addon->ndim); return nullptr; }
      dim_idx.assign(l, l + n);
    }
    snprintf(dname, sizeof dname, "distributions[%d].ded_idx", i + 1);
    if (rc_has_key(rca, dname, err) == 1) {
      doname = false;
      long *l = nullptr;
      const int n = (int)rc_get_integer_array(rca, dname, &l, err);
      RCFWD(*err, nullptr);
      if (n != addon->n_ded) { RCERR(*err, dist_base, "bad size for %s (got %d expected %d)", *err, nullptr, dname, n, addon->n_ded); return nullptr; }
      ded_idx.assign(l, l + n);
    }
    if (doname && hn == 1) add_to_combine_distribution_name(comb, addon, err);  // (the reference then adds it a second time, pmc_rc.c:155-160; once is the intent)
    else add_to_combine_distribution(comb, addon, dim_idx.empty() ? nullptr : dim_idx.data(), ded_idx.empty() ? nullptr : ded_idx.data(), err);
    RCFWD(*err, nullptr);
  }
  rc_close(&rca);
  return comb;
}

void *rcinit_add_gaussian_prior(confFile *rc, char *root, error **err) {  // pmc_rc.c:233-305
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  distribution *addon = init_distribution_from_rc(rca, (char *)"distribution", err);
  RCFWD(*err, nullptr);
  double *mn = nullptr, *var = nullptr;
  const int ndim = (int)rc_get_real_array(rca, (char *)"mean", &mn, err);
  RCFWD(*err, nullptr);
  rc_get_real_assquarematrix(rca, (char *)"var", &var, ndim, err);
  RCFWD(*err, nullptr);
  distribution *comb = nullptr;
  if (rc_has_key(rca, (char *)"parameter_name", err) == 1) {
    char **names = nullptr;
    rc_get_string_array(rca, (char *)"parameter_name", &names, err);
    RCFWD(*err, nullptr);
    comb = add_gaussian_prior_2_name(addon, ndim, names, mn, var, err);
  } else {
    std::vector<int> idx;
    if (rc_has_key(rca, (char *)"prior_idx", err) == 1) {
      long *l = nullptr;
      const int n = (int)rc_get_integer_array(rca, (char *)"prior_idx", &l, err);
      RCFWD(*err, nullptr);
      if (n != ndim) { RCERR(*err, dist_base, "bad size for %s (got %d expected %d)", *err, nullptr, "prior_idx", n, ndim); return nullptr; }
      idx.assign(l, l + n);
    } else {
      for (int i = 0; i < ndim; ++i) idx.push_back(i);  // the reference passes NULL, which add_to_combine_distribution reads as the identity map
    }
    comb = add_gaussian_prior_2(addon, ndim, idx.data(), mn, var, err);
  }
  RCFWD(*err, nullptr);
  rc_close(&rca);
  return comb;
}

parabox *parabox_from_rc(confFile *rc, char *root, error **err) {  // pmc_rc.c:356-389
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  const int ndim = (int)rc_array_size(rca, (char *)"min", err);
  RCFWD(*err, nullptr);
  const int nmax = (int)rc_array_size(rca, (char *)"max", err);
  RCFWD(*err, nullptr);
  if (ndim != nmax) { RCERR(*err, pb_base, "min and max arrays have different sizes (%d and %d)", *err, nullptr, ndim, nmax); return nullptr; }
  parabox *pb = init_parabox(ndim, err);
  RCFWD(*err, nullptr);
  for (int i = 0; i < ndim; ++i) {
    char par[100];
    snprintf(par, sizeof par, "min[%d]", i + 1);
    const double rmin = rc_get_real(rca, par, err);
    RCFWD(*err, nullptr);
    snprintf(par, sizeof par, "max[%d]", i + 1);
    const double rmax = rc_get_real(rca, par, err);
    RCFWD(*err, nullptr);
    add_slab(pb, i, rmin, rmax, err);
    RCFWD(*err, nullptr);
  }
  rc_close(&rca);
  return pb;
}

mvdens *rc_mvdens(confFile *rc, mvdens *mv, error **err) {  // pmc_rc.c:391-433
  double *mn = nullptr;
  const int ndum = (int)rc_get_real_array(rc, (char *)"mean", &mn, err);
  RCFWD(*err, nullptr);
  mvdens *g = mv;
  if (!g) { g = mvdens_alloc(ndum, err); RCFWD(*err, nullptr); }
  const int ndim = (int)g->ndim;
  if (ndum != ndim) { RCERR(*err, pmc_dimension, "'mean' in %s does not have the right number elements (got %d expected %d)", *err, nullptr, rc->alias_prefix, ndum, ndim); return nullptr; }
  memcpy(g->mean, mn, sizeof(double) * ndim);
  double *var = nullptr;
  rc_get_real_assquarematrix(rc, (char *)"var", &var, ndim, err);
  RCFWD(*err, nullptr);
  memcpy(g->std, var, sizeof(double) * ndim * ndim);
  g->chol = (int)rc_safeget_integer(rc, (char *)"chol", 0, err);
  RCFWD(*err, nullptr);
  if (g->chol) g->detL = determinant(g->std, ndim);
  g->band_limit = (size_t)rc_safeget_integer(rc, (char *)"bdl", ndim, err);
  RCFWD(*err, nullptr);
  g->df = (int)rc_safeget_integer(rc, (char *)"df", -1, err);
  RCFWD(*err, nullptr);
  return g;
}

double *rc_get_weight(confFile *rc, int sz, error **err) {  // pmc_rc.c:435-465
  double *w = nullptr;
  if (rc_has_key(rc, (char *)"weight", err) == 0) {
    w = (double *)rc_add_malloc(rc, sizeof(double) * sz, err);
    RCFWD(*err, nullptr);
    for (int i = 0; i < sz; ++i) w[i] = 1. / sz;
    return w;
  }
  rc_get_real_asarray(rc, (char *)"weight", &w, sz, err);
  RCFWD(*err, nullptr);
  double sw = 0;
  for (int i = 0; i < sz; ++i) {
    if (w[i] < 0) { RCERR(*err, pmc_negWeight, "weight %d is negative (%g)", *err, nullptr, i, w[i]); return nullptr; }
    sw += w[i];
  }
  if (sw == 0) { RCERR(*err, pmc_negWeight, "all weights are 0", *err, nullptr); return nullptr; }
  for (int i = 0; i < sz; ++i) w[i] = w[i] / sw;
  return w;
}

mix_mvdens *rc_mix_mvdens(confFile *rc, error **err) {  // pmc_rc.c:467-599
  mix_mvdens *mmv = nullptr;
  if (rc_has_key(rc, (char *)"fromfile", err) == 1) {
    // the reference reads HDF5 here (never built: HAS_HDF5 is undefined in its build); the text format mix_mvdens_dump
    // writes - what the drivers save after every iteration - is the resume path that exists
    char *name = rc_get_string(rc, (char *)"fromfile", err);
    RCFWD(*err, nullptr);
    FILE *f = fopen_err(name, "r", err);
    RCFWD(*err, nullptr);
    mmv = mix_mvdens_dwnp(f, err);
    fclose(f);
    RCFWD(*err, nullptr);
  } else {
    const int ncomp = (int)rc_get_integer(rc, (char *)"ncomp", err);
    RCFWD(*err, nullptr);
    const int ndim = (int)rc_get_integer(rc, (char *)"ndim", err);
    RCFWD(*err, nullptr);
    mmv = mix_mvdens_alloc(ncomp, ndim, err);
    RCFWD(*err, nullptr);
    double *w = rc_get_weight(rc, ncomp, err);
    RCFWD(*err, nullptr);
    memcpy(mmv->wght, w, sizeof(double) * ncomp);
    if (rc_has_key(rc, (char *)"draw_from", err) == 1) {
      // the means of components 1.. are drawn from component 0 (pmc_rc.c:537-556): mvdens_ran, i.e. device draws
      confFile *rca = rc_alias(rc, (char *)"draw_from", err);
      RCFWD(*err, nullptr);
      gsl_rng *r = rc_gsl_rng(rca, (char *)"", err);
      RCFWD(*err, nullptr);
      rc_mvdens(rca, mmv->comp[0], err);
      RCFWD(*err, nullptr);
      for (int ic = 1; ic < ncomp; ++ic) {
        mvdens_ran(mmv->comp[ic]->mean, mmv->comp[0], r, err);
        RCFWD(*err, nullptr);
        mmv->comp[ic]->chol = mmv->comp[0]->chol;
        mmv->comp[ic]->band_limit = mmv->comp[0]->band_limit;
        mmv->comp[ic]->df = mmv->comp[0]->df;
        mmv->comp[ic]->detL = mmv->comp[0]->detL;
        memcpy(mmv->comp[ic]->std, mmv->comp[0]->std, sizeof(double) * ndim * ndim);
      }
      rc_close(&rca);
      pmcb200_rng_free(r);
    } else {
      const int n = (int)rc_array_size(rc, (char *)"component", err);
      RCFWD(*err, nullptr);
      if (n != ncomp) { RCERR(*err, pmc_dimension, "not the right number of elements in component (got %d expected %d)", *err, nullptr, n, ncomp); return nullptr; }
      for (int ic = 0; ic < ncomp; ++ic) {
        char sc[100];
        snprintf(sc, sizeof sc, "component[%d]", ic + 1);
        confFile *rca = rc_alias(rc, sc, err);
        RCFWD(*err, nullptr);
        rc_mvdens(rca, mmv->comp[ic], err);
        RCFWD(*err, nullptr);
        rc_close(&rca);
      }
    }
  }
  if (rc_has_key(rc, (char *)"save", err) == 1) {
    char *file = rc_get_string(rc, (char *)"save", err);
    RCFWD(*err, nullptr);
    FILE *f = fopen_err(file, "w", err);
    RCFWD(*err, nullptr);
    mix_mvdens_dump(f, mmv);
    fclose(f);
  }
  return mmv;
}

distribution *rcinit_mix_mvdens(confFile *rc, char *root, error **err) {  // pmc_rc.c:601-618
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  mix_mvdens *mmv = rc_mix_mvdens(rca, err);
  RCFWD(*err, nullptr);
  distribution *d = mix_mvdens_distribution((int)mmv->ndim, mmv, err);
  RCFWD(*err, nullptr);
  rc_close(&rca);
  return d;
}

pmc_simu *init_importance_from_rc(confFile *rc, char *root, error **err) {  // pmc_rc.c:620-665
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  distribution *target = init_distribution_from_rc(rca, (char *)"target", err);
  RCFWD(*err, nullptr);
  distribution *proposal = init_distribution_from_rc(rca, (char *)"proposal", err);
  RCFWD(*err, nullptr);
  if (proposal->ndim != target->ndim) { RCERR(*err, pmc_dimension, "incompatible dim for proposal (%d) and target (%d)", *err, nullptr, proposal->ndim, target->ndim); return nullptr; }
  const long nsamples = rc_get_integer(rca, (char *)"nsample", err);
  RCFWD(*err, nullptr);
  pmc_simu *psim = pmc_simu_init_plus_ded(nsamples, target->ndim, target->n_ded, err);
  RCFWD(*err, nullptr);
  parabox *pb = parabox_from_rc(rca, (char *)"pb", err);
  RCFWD(*err, nullptr);
  if (pb->ndim != target->ndim) { RCERR(*err, pmc_dimension, "incompatible dim for pb (got %d expected %d)", *err, nullptr, pb->ndim, target->ndim); return nullptr; }
  pmc_simu_init_target(psim, target, pb, err);
  RCFWD(*err, nullptr);
  const int pstep = (int)rc_safeget_integer(rca, (char *)"print_pc", 0, err);
  RCFWD(*err, nullptr);
  pmc_simu_init_proposal(psim, proposal, pstep, err);
  RCFWD(*err, nullptr);
  rc_close(&rca);
  return psim;
}

pmc_simu *init_pmc_from_rc(confFile *rc, char *root, error **err) {  // pmc_rc.c:712-721
  pmc_simu *psim = init_importance_from_rc(rc, root, err);
  RCFWD(*err, nullptr);
  pmc_simu_init_pmc(psim, nullptr, nullptr, &update_prop_rb_void);
  return psim;
}

gsl_rng *rc_gsl_rng(confFile *rc, char *root, error **err) {  // pmc_rc.c:723-768 (gsl_rng_default = mt19937)
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  if (rc_has_key(rca, (char *)"rng", err) == 1) {
    char *type = rc_get_string(rca, (char *)"rng", err);
    RCFWD(*err, nullptr);
    if (strcasecmp(type, "mt19937") != 0) { RCERR(*err, pmc_undef, "Unknown generator '%s' (this build carries mt19937, the GSL default)", *err, nullptr, type); return nullptr; }
  }
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  void *st = malloc(sizeof(MtState));
  if (!r || !st) { RCERR(*err, -102, "Cannot allocate", *err, nullptr); return nullptr; }
  r->type = &mt_type;
  r->state = st;
  long seed = 0;  // gsl_rng_default_seed
  if (rc_has_key(rca, (char *)"seed", err) == 1) { seed = rc_get_integer(rca, (char *)"seed", err); RCFWD(*err, nullptr); }
  mt_set(st, (unsigned long)seed);
  rc_close(&rca);
  return r;
}
void pmcb200_rng_free(gsl_rng *r) {
  if (!r) return;
  free(r->state);
  free(r);
}

// ---- libpmc_mpi's config functions (pmc_mpi.c:236-330, compiled there only with Lua) ---------------------------------
// rank 0 builds the mixture from the parameter file (the draw_from initialisation draws on ITS GPU), the others receive it
mix_mvdens *rc_mix_mvdens_mpi(confFile *rc, error **err) {  // pmc_mpi.c:236-250
  mix_mvdens *mmv = nullptr;
  int rank = 0, size = 1;
  pmcb200_mpi_world(&rank, &size, nullptr, 0);
  if (rank == 0) {
    mmv = rc_mix_mvdens(rc, err);
    RCFWD(*err, nullptr);
  }
  mmv = (mix_mvdens *)mvdens_broadcast_mpi(mmv, err);
  RCFWD(*err, nullptr);
  return mmv;
}
static void *rcinit_mix_mvdens_mpi3(confFile *rc, char *root, error **err) {
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  mix_mvdens *mmv = rc_mix_mvdens_mpi(rca, err);
  RCFWD(*err, nullptr);
  distribution *dist = init_distribution_full((int)mmv->ndim, mmv, &mix_mvdens_log_pdf_void, &mix_mvdens_free_void, &simulate_mix_mvdens_void, 0, nullptr, err);
  RCFWD(*err, nullptr);
  dist->broadcast_mpi = &mvdens_broadcast_mpi;
  rc_close(&rca);
  return dist;
}
// pmc_mpi.h:96 declares two arguments although INIT_FROM_RC calls every rcinit_* with three (rc, "", err); the parameter
// file name "mix_mvdens_mpi" is served by the three-argument form above, this symbol keeps the header's prototype
distribution *rcinit_mix_mvdens_mpi(confFile *rc, error **err) { return (distribution *)rcinit_mix_mvdens_mpi3(rc, (char *)"", err); }

pmc_simu *init_importance_mpi_from_rc(confFile *rc, char *root, error **err) {  // pmc_mpi.c:274-318
  confFile *rca = rc_alias(rc, root, err);
  RCFWD(*err, nullptr);
  distribution *target = init_distribution_from_rc(rca, (char *)"target", err);
  RCFWD(*err, nullptr);
  distribution *proposal = init_distribution_from_rc(rca, (char *)"proposal", err);
  RCFWD(*err, nullptr);
  if (proposal->ndim != target->ndim) { RCERR(*err, pmc_dimension, "incompatible dim for proposal (%d) and target (%d)", *err, nullptr, proposal->ndim, target->ndim); return nullptr; }
  const long nsamples = rc_get_integer(rca, (char *)"nsample", err);
  RCFWD(*err, nullptr);
  pmc_simu *psim = pmc_simu_init_mpi(nsamples, target->ndim, target->n_ded, err);
  RCFWD(*err, nullptr);
  parabox *pb = parabox_from_rc(rca, (char *)"pb", err);
  RCFWD(*err, nullptr);
  if (pb->ndim != target->ndim) { RCERR(*err, pmc_dimension, "incompatible dim for pb (got %d expected %d)", *err, nullptr, pb->ndim, target->ndim); return nullptr; }
  pmc_simu_init_target(psim, target, pb, err);
  RCFWD(*err, nullptr);
  const int pstep = (int)rc_safeget_integer(rca, (char *)"print_pc", 0, err);
  RCFWD(*err, nullptr);
  pmc_simu_init_proposal(psim, proposal, pstep, err);
  RCFWD(*err, nullptr);
  rc_close(&rca);
  return psim;
}
pmc_simu *init_pmc_mpi_from_rc(confFile *rc, char *root, error **err) {  // pmc_mpi.c:320-330
  pmc_simu *psim = init_importance_mpi_from_rc(rc, root, err);
  RCFWD(*err, nullptr);
  pmc_simu_init_pmc(psim, nullptr, nullptr, &update_prop_rb_void);
  return psim;
}

char *get_itername(confFile *rc, char *key, int iter, error **err) {  // pmc_rc.c:770-794
  if (rc_is_array(rc, key, err)) {
    char keyix[1000];
    snprintf(keyix, sizeof keyix, "%s[%d]", key, iter);
    char *res = rc_get_string(rc, keyix, err);
    RCFWD(*err, nullptr);
    return res;
  }
  char *res = rc_get_string(rc, key, err);
  RCFWD(*err, nullptr);
  const size_t ll = strlen(res);
  char *oes = (char *)rc_add_malloc(rc, ll + 100, err);
  RCFWD(*err, nullptr);
  snprintf(oes, ll + 100, res, iter);
  return oes;
}

}  // extern "C"
