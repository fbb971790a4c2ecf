"""CPU: the C-ABI library loads and exports every declared symbol, fails loudly without a GPU, and its host-only entry
points (Cholesky, finalisation from sufficient statistics, shard layout) agree with the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_py as O
from pmclib_b200 import workloads as W
from pmclib_b200.lib import EXPORTS, LIB_PATH, load_library
from util import compare_mixture, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmcgpu_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = load_library()
    declared = _declared("pmcb200.h")
    assert set(declared) == set(EXPORTS), set(declared) ^ set(EXPORTS)
    for s in declared:
        assert hasattr(lib, s), s
    # the drop-in names of pmclib's own API (include/pmclib_abi.h)
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    defined = set(l.split()[-1] for l in out.splitlines() if l.strip())
    for s in ["pmc_simu_init", "pmc_simu_init_plus_ded", "pmc_simu_free", "pmc_simu_realloc", "pmc_simu_init_target",
              "pmc_simu_init_proposal", "pmc_simu_init_pmc", "pmc_simu_init_classic_importance", "pmc_simu_init_classic_pmc",
              "pmc_simu_importance", "pmc_simu_pmc_step", "simulate_mix_mvdens", "simulate_mix_mvdens_void",
              "get_importance_weight", "get_importance_weight_plus_ded", "generic_get_importance_weight",
              "generic_get_importance_weight_and_deduced", "generic_get_importance_weight_and_deduced_verb",
              "normalize_importance_weight", "perplexity", "perplexity_and_ess", "evidence", "update_prop", "update_prop_rb",
              "update_prop_rb_void", "mix_mvdens_distribution", "pmc_simu_dump", "out_pmc_simu", "filter_histogram_init",
              "filter_histogram", "mean_from_psim", "sample_from_pmc_simu", "resample_residual", "clip_weights", "prepare_update", "cleanup_after_update",
              "double_cmp", "pmc_simu_print", "fopen_err", "malloc_err", "calloc_err", "mix_mvdens_alloc", "mix_mvdens_free",
              "mix_mvdens_free_void", "mix_mvdens_cholesky_decomp", "mix_mvdens_log_pdf", "mix_mvdens_log_pdf_void",
              "mix_mvdens_dump", "mix_mvdens_dwnp", "mix_mvdens_print", "mvdens_alloc", "mvdens_init", "mvdens_free",
              "mvdens_cholesky_decomp", "mvdens_log_pdf", "mvdens_log_pdf_void", "scalar_product", "determinant", "ranindx",
              "init_distribution", "init_distribution_full", "init_simple_distribution", "free_distribution",
              "distribution_lkl", "distribution_retrieve", "distribution_set_default", "init_parabox", "add_face",
              "add_slab", "add_lobound", "add_hibound", "add_halfspace", "isinBox", "free_parabox", "initError",
              "endError", "newError", "newErrorVA", "printError", "stringError", "purgeError", "getErrorValue", "_isError",
              "init_bentGauss_distribution", "bentGauss_lkl"]:
        assert s in defined, s


def test_no_gpu_fails_loudly_and_never_falls_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = load_library()
    h = C.c_void_p()
    assert lib.pmcgpu_create(0, C.byref(h)) == -6024  # PMCGPU_ERR_NODEVICE
    assert not h.value


def test_dropin_driver_links_and_reports_no_device(tmp_path):
    exe = str(tmp_path / "dropin_driver")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "dropin_driver.c"),
                    "-o", exe, "-L" + os.path.dirname(LIB_PATH), "-lpmcb200", "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-lm"], check=True)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run modes are covered by test_gpu_dropin.py")
    K, d, N = 4, 3, 50
    cfg = tmp_path / "in.bin"
    with open(cfg, "wb") as f:
        f.write(np.array([K, d, N, 1, 0, 1], np.int64).tobytes())
        f.write(np.array([0.1, 10.0]).tobytes())
        f.write(np.full(K, 1 / K).tobytes()); f.write(np.zeros((K, d)).tobytes()); f.write(np.tile(np.eye(d), (K, 1, 1)).tobytes())
    r = subprocess.run([exe, str(cfg), str(tmp_path / "out.bin"), "3"], capture_output=True, text=True, check=True)
    assert '"is_error": 1' in r.stdout and "-6024" in r.stdout


def test_host_cholesky_is_the_gsl_algorithm():
    lib = load_library()
    g = np.load(os.path.join(ROOT, "tests", "golden", "misc.npz"))
    a = g["spd"].copy()
    det = C.c_double()
    assert lib.pmcgpu_cholesky(a.shape[0], O.P(a), C.byref(det)) == 0
    assert np.array_equal(a, g["chol"]) and det.value == float(g["chol_det"])  # bit-identical to the reference
    b = g["bad"].copy()
    assert lib.pmcgpu_cholesky(b.shape[0], O.P(b), C.byref(det)) == -706
    assert np.array_equal(b, g["bad_after"])
    rng = np.random.default_rng(0)
    for d in (1, 2, 5, 30, 128):
        m = rng.normal(0, 1, (d, d)); m = m @ m.T + d * np.eye(d)
        x, y = m.copy(), m.copy()
        d1, d2 = C.c_double(), C.c_double()
        assert lib.pmcgpu_cholesky(d, O.P(x), C.byref(d1)) == 0 and O.orc().orc_cholesky(d, O.P(y), C.byref(d2)) == 0
        assert np.array_equal(x, y) and d1.value == d2.value


def numpy_stats(X, w, lr, flg, wl, maxR, pivot):
    """Sufficient statistics of the RB update in plain numpy (test infrastructure): W, s1, S2 about a pivot."""
    m = O.Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"])
    K, d = m.K, m.d
    ok = flg == 1
    Xo, wo, lro = X[ok], w[ok], lr[ok]
    W_ = np.zeros(K); s1 = np.zeros((K, d)); S2 = np.zeros((K, d, d))
    for k in range(K):
        L = np.tril(m.L[k])
        y = np.linalg.solve(L, (Xo - m.mean[k]).T).T
        ld = -0.5 * (d * 1.837877066409 + (y ** 2).sum(1)) - np.log(m.detL[k])
        g = wo * m.wght[k] * np.exp(ld - (lro - maxR + 500.0))
        v = Xo - pivot
        W_[k] = g.sum(); s1[k] = g @ v; S2[k] = (v * g[:, None]).T @ v
    return W_, s1, S2


@pytest.mark.parametrize("name,N", [("shipped", 20000), ("c2", 20000)])
def test_finalize_update_from_statistics(name, N):
    """pmcgpu_finalize_update (host, O(K d^3)) turns reduced statistics into the reference's updated proposal,
    including component deaths (bit-exact pattern) and the retry flag."""
    lib = load_library()
    # well-conditioned cases: with degenerate weights (ESS ~ 1) the one-pass form cancels catastrophically, which is what
    # the pivot guard of pmcgpu_pmc_step detects before falling back to the two-pass kernels (test_gpu_parity.py)
    wl = {"shipped": W.shipped_example, "c2": W.gmm5}[name]()
    X, idx = O.orc_simulate(N, wl, seed=21)
    o = O.orc_iteration(X, idx, wl, do_update=1)
    assert o["rc"] == 0
    K, d = wl["K"], wl["d"]
    pivot = (wl["wght"][:, None] * wl["mean"]).sum(0)
    W_, s1, S2 = numpy_stats(X, o["w_norm"], o["log_rho"], o["flg"], wl, o["maxR"], pivot)
    count = np.bincount(idx[o["flg"] == 1].astype(np.int64), minlength=K).astype(np.uint64)
    m = O.Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"])
    wght, mean, L, detL = m.wght.copy(), m.mean.copy(), m.L.copy(), m.detL.copy()
    alive = np.zeros(K, np.int32)
    retry = C.c_int(0)
    rc = lib.pmcgpu_finalize_update(K, d, O.P(W_), O.P(W_), O.P(np.ascontiguousarray(s1)), O.P(np.ascontiguousarray(S2)), O.P(pivot), 0,
                                    O.P(count), N, None, C.byref(retry), O.P(wght), O.P(mean), O.P(L), O.P(detL), O.P(alive))
    assert rc == 0 and retry.value == o["retry"]
    assert np.array_equal(alive.astype(bool), o["o_wght"] > 0)
    errs = compare_mixture(wght, mean, L, o["o_wght"], o["o_mean"], o["o_std"], rtol=1e-11, what=name)
    print(errs)


def test_finalize_kills_small_components_and_reports_negweight():
    lib = load_library()
    K, d = 3, 2
    W_ = np.array([0.5, 0.3, 0.2]); s1 = np.zeros((K, d)); S2 = np.tile(np.eye(d), (K, 1, 1)) * W_[:, None, None]
    pivot = np.zeros(d)
    wght = np.full(K, 1 / 3); mean = np.zeros((K, d)); L = np.tile(np.eye(d), (K, 1, 1)); detL = np.ones(K)
    alive = np.zeros(K, np.int32); retry = C.c_int(0)
    count = np.array([100, 20, 21], np.uint64)  # count == 20: not updated (>) and pruned by the weight floor (pmc.c:1382,1517)
    rc = lib.pmcgpu_finalize_update(K, d, O.P(W_), O.P(W_), O.P(s1), O.P(S2), O.P(pivot), 0, O.P(count), 1000, None,
                                    C.byref(retry), O.P(wght), O.P(mean), O.P(L), O.P(detL), O.P(alive))
    assert rc == 0 and list(alive) == [1, 0, 1]
    assert wght[1] == 0 and not L[1].any() and abs(wght.sum() - 1) < 1e-15
    assert relerr(wght[[0, 2]], np.array([0.5, 0.2]) / 0.7) < 1e-15
    count[:] = 5
    rc = lib.pmcgpu_finalize_update(K, d, O.P(W_), O.P(W_), O.P(s1), O.P(S2), O.P(pivot), 0, O.P(count), 1000, None,
                                    C.byref(retry), O.P(wght), O.P(mean), O.P(L), O.P(detL), O.P(alive))
    assert rc == -6005
    # non positive-definite statistics: first failure restores the previous factor and raises retry, second kills
    count[:] = 100
    S2b = S2.copy(); S2b[0] = -S2b[0]
    wght = np.full(K, 1 / 3); L = np.tile(2 * np.eye(d), (K, 1, 1)); retry = C.c_int(0)
    rc = lib.pmcgpu_finalize_update(K, d, O.P(W_), O.P(W_), O.P(s1), O.P(S2b), O.P(pivot), 0, O.P(count), 1000, None,
                                    C.byref(retry), O.P(wght), O.P(mean), O.P(L), O.P(detL), O.P(alive))
    assert rc == 0 and retry.value == 1 and np.array_equal(L[0], 2 * np.eye(d)) and wght[0] > 0
    retry = C.c_int(1); wght = np.full(K, 1 / 3)
    rc = lib.pmcgpu_finalize_update(K, d, O.P(W_), O.P(W_), O.P(s1), O.P(S2b), O.P(pivot), 0, O.P(count), 1000, None,
                                    C.byref(retry), O.P(wght), O.P(mean), O.P(L), O.P(detL), O.P(alive))
    assert rc == 0 and retry.value == 0 and wght[0] == 0 and alive[0] == 0


def test_finalize_threaded_large_shape_matches_numpy():
    """At d = 128, K = 64 the host finalisation runs on several threads (pmc_hostonly.h): same result as a plain numpy
    evaluation of mu' = c + s/W, Sigma' = S/W - (s/W)(s/W)^T, L = chol(Sigma'), one failed component restored + retry."""
    lib = load_library()
    wl = W.gauss128()
    K, d = wl["K"], wl["d"]
    rng = np.random.default_rng(5)
    W_ = rng.uniform(0.5, 1.5, K) / K
    shift = rng.normal(0.0, 0.05, (K, d))
    s1 = shift * W_[:, None]
    S2 = np.array([W_[k] * (wl["cov"][k] + np.outer(shift[k], shift[k])) for k in range(K)])
    S2[7] = -S2[7]   # not positive definite: first failure restores the old factor and raises retry
    pivot = rng.normal(0.0, 0.1, d)
    count = np.full(K, 1000, np.uint64)
    m = O.Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"])
    for _ in range(2):  # twice: the scratch buffers persist across calls
        wght, mean, L, detL = m.wght.copy(), m.mean.copy(), m.L.copy(), m.detL.copy()
        alive = np.zeros(K, np.int32)
        retry = C.c_int(0)
        rc = lib.pmcgpu_finalize_update(K, d, O.P(W_), O.P(W_), O.P(np.ascontiguousarray(s1)), O.P(np.ascontiguousarray(S2)),
                                        O.P(pivot), 0, O.P(count), 100000, None, C.byref(retry), O.P(wght), O.P(mean), O.P(L),
                                        O.P(detL), O.P(alive))
        assert rc == 0 and retry.value == 1 and alive.all()
        assert np.array_equal(L[7], m.L[7])
        assert relerr(wght, W_ / W_.sum()) < 1e-14
        for k in range(K):
            if k == 7:
                continue
            assert relerr(mean[k], pivot + shift[k], scale=0.1) < 1e-13
            Lk = np.linalg.cholesky(wl["cov"][k])
            assert relerr(np.tril(L[k]), Lk, scale=1.0) < 1e-12   # error relative to the size of the factor
            assert relerr(detL[k], np.prod(np.diag(Lk))) < 1e-11


def test_shard_range_partitions_like_send_simulation():
    from pmclib_b200.api import shard_range
    for N, P in [(100, 8), (101, 8), (7, 8), (10**8, 3), (1, 1)]:
        parts = [shard_range(N, r, P) for r in range(P)]
        assert parts[0][0] == 0 and sum(n for _, n in parts) == N
        for (f0, n0), (f1, _) in zip(parts, parts[1:]):
            assert f0 + n0 == f1
        lens = [n for _, n in parts]
        assert max(lens) - min(lens) <= 1 and sorted(lens, reverse=True) == lens  # the first N%P ranks get one more


def test_on_disk_formats_are_byte_identical_to_the_reference(tmp_path):
    """mix_mvdens_dump, mvdens_dump, pmc_simu_dump and out_pmc_simu write the bytes the reference's own functions write
    (tests/golden/dump_ref.*, produced by oracle/ref_driver.c:ref_dump_all from the unmodified reference), and
    mix_mvdens_dwnp reads them back.  Host-only entry points: no GPU needed."""
    import subprocess
    gold = os.path.join(ROOT, "tests", "golden", "dump_ref")
    exe = str(tmp_path / "api2_driver")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "api2_driver.c"),
                    "-o", exe, "-L" + os.path.dirname(LIB_PATH), "-lpmcb200", "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-lm"], check=True)
    r = subprocess.run([exe, "dump", gold + ".in", str(tmp_path / "ours")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for ext in (".mix", ".mvd", ".psim", ".out"):
        assert open(str(tmp_path / "ours") + ext, "rb").read() == open(gold + ext, "rb").read(), ext
    assert "dwnp K 2 d 3" in r.stdout and "df0 -1" in r.stdout
    assert float(r.stdout.split("worst")[1].split()[0]) < 1e-14


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pmclib", "include")), reason="reference headers not present on this machine")
def test_c_driver_compiles_against_the_reference_headers(tmp_path):
    """The C driver — a pmclib caller — built with the REFERENCE's own headers (pmc.h, mvdens.h, distribution.h, parabox.h,
    errorlist.h, sampleTarget.h; GSL types from oracle/gslshim) instead of include/pmclib_abi.h, linked against
    libpmcb200.so: prototypes and struct layouts of the drop-in agree with the reference.  Without a GPU the run must fail
    loudly through the reference's error chain (mode 3)."""
    exe = str(tmp_path / "dropin_driver_ref")
    inc = ["-I" + os.path.join(ROOT, "oracle", "gslshim"), "-I" + os.path.join(ROOT, "oracle", "gslshim", "luastub"),
           "-I" + os.path.join(REF, "pmctools", "include"), "-I" + os.path.join(REF, "pmclib", "include"),
           "-I" + os.path.join(REF, "examples", "target")]
    r = subprocess.run(["gcc", "-O2", "-std=gnu99", "-Wall", "-DPMCLIB_REFERENCE_HEADERS", "-D_NOFFTW_", "-D_GNU_SOURCE"] + inc +
                       [os.path.join(ROOT, "tests", "c", "dropin_driver.c"), "-o", exe, "-L" + os.path.dirname(LIB_PATH), "-lpmcb200",
                        "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "conflicting types" not in r.stderr and "incompatible" not in r.stderr, r.stderr
    # sizes and offsets the driver relies on, as the reference headers define them vs the mirror
    probe = tmp_path / "probe.c"
    body = r'''
#include <stdio.h>
#include <stddef.h>
#ifdef PMCLIB_REFERENCE_HEADERS
#include "pmc.h"
#include "sampleTarget.h"
#else
#include "pmclib_abi.h"
#endif
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu ", sizeof(pmc_simu), sizeof(mvdens), sizeof(mix_mvdens), sizeof(distribution), sizeof(parabox), sizeof(error));
  printf("%zu %zu %zu %zu %zu %zu ", offsetof(pmc_simu, X), offsetof(pmc_simu, flg), offsetof(pmc_simu, indices), offsetof(pmc_simu, logSum), offsetof(pmc_simu, proposal), offsetof(pmc_simu, mpi_size));
  printf("%zu %zu %zu %zu %zu ", offsetof(mvdens, std), offsetof(mvdens, band_limit), offsetof(mvdens, chol), offsetof(mvdens, detL), offsetof(mvdens, x_tmp));
  printf("%zu %zu %zu %zu\n", offsetof(mix_mvdens, comp), offsetof(mix_mvdens, cwght), offsetof(distribution, broadcast_mpi), offsetof(distribution, name));
  return 0;
}
'''
    probe.write_text(body)
    outs = []
    for flags in (["-DPMCLIB_REFERENCE_HEADERS", "-D_NOFFTW_", "-D_GNU_SOURCE"] + inc, ["-I" + os.path.join(ROOT, "include")]):
        e = str(tmp_path / ("probe" + str(len(outs))))
        subprocess.run(["gcc", "-std=gnu99"] + flags + [str(probe), "-o", e], check=True, capture_output=True)
        outs.append(subprocess.run([e], capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1], outs
    import torch
    if torch.cuda.is_available():
        return
    K, d, N = 4, 3, 50
    cfg = tmp_path / "in.bin"
    with open(cfg, "wb") as f:
        f.write(np.array([K, d, N, 1, 0, 1], np.int64).tobytes())
        f.write(np.array([0.1, 10.0]).tobytes())
        f.write(np.full(K, 1 / K).tobytes()); f.write(np.zeros((K, d)).tobytes()); f.write(np.tile(np.eye(d), (K, 1, 1)).tobytes())
    r = subprocess.run([exe, str(cfg), str(tmp_path / "out.bin"), "3"], capture_output=True, text=True, check=True)
    assert '"is_error": 1' in r.stdout and "-6024" in r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpmcref.so")), reason="compiled reference not built")
@pytest.mark.parametrize("N,nsig", [(2000, 0.002), (5000, 1.0), (333, 1e-9)])
def test_filter_importance_weight_sig_equals_the_compiled_reference(tmp_path, N, nsig):
    """filter_importance_weight_sig (pmc.c:772-827, a global symbol of libpmc that pmc.h does not declare): the same C program
    (tests/c/filter_sig_driver.c) linked against libpmcb200.so and against the compiled reference prints byte-identical flags,
    log-weights, maxW and stderr reports - including the reference's comparison with nsig itself instead of nsig sigma."""
    src = os.path.join(ROOT, "tests", "c", "filter_sig_driver.c")
    outs = []
    for tag, libdir, lib in (("ours", os.path.dirname(LIB_PATH), "pmcb200"), ("ref", os.path.join(ROOT, "oracle", "_ref"), "pmcref")):
        exe = str(tmp_path / f"fs_{tag}")
        subprocess.run(["gcc", "-O2", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), src, "-o", exe, "-L" + libdir, "-l" + lib,
                        "-Wl,-rpath," + libdir, "-lm"], check=True)
        r = subprocess.run([exe, str(N), repr(nsig)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-500:]
        outs.append((r.stdout, r.stderr))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == outs[1][1]
    assert outs[0][0].startswith("isLog 1 ")


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpmcref.so")), reason="compiled reference not built")
def test_liberrorio_helpers_equal_the_compiled_reference(tmp_path):
    """io.h:34-75 (pmc_hostio.cpp): line counting, the text and binary readers, resize_err / realloc_err, chomp, read_double,
    is_directory, readASCII - tests/c/io_driver.c linked against libpmcb200.so and against the compiled reference prints the
    same values and the same error codes (err_io -101, err_allocate -102, io_file -202), quirks included."""
    src = os.path.join(ROOT, "tests", "c", "io_driver.c")
    outs = []
    for tag, libdir, lib in (("ours", os.path.dirname(LIB_PATH), "pmcb200"), ("ref", os.path.join(ROOT, "oracle", "_ref"), "pmcref")):
        exe = str(tmp_path / f"io_{tag}")
        work = tmp_path / f"work_{tag}"
        work.mkdir()
        subprocess.run(["gcc", "-O1", "-std=gnu99", "-w", "-I" + os.path.join(ROOT, "include"), src, "-o", exe, "-L" + libdir, "-l" + lib,
                        "-Wl,-rpath," + libdir, "-lm"], check=True)
        r = subprocess.run([exe, "."], capture_output=True, text=True, timeout=120, cwd=str(work))
        assert r.returncode == 0, r.stderr[-500:]
        outs.append(r.stdout)
    assert outs[0] == outs[1]
    assert "read_bin_vector too long -> error -202" in outs[0] and "n 10 lines 4:" in outs[0]
