"""GPU: the second half of pmclib's API around the PMC path (pmc_dropin2.cpp), called from a C program written like a
pmclib user (tests/c/api2_driver.c): the mixture wire format, single-sample draws (mvdens_ran, mix_mvdens_ran), the text
readers, pmc_simu_from_file, parameter names, Gaussian priors by name, error pickling."""
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

import oracle_py as O
from pmclib_b200.lib import LIB_PATH
from util import RTOL, WTOL, assert_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api2(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("api2") / "api2_driver")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "api2_driver.c"),
                    "-o", exe, "-L" + os.path.dirname(LIB_PATH), "-lpmcb200", "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-lm"], check=True)
    return exe


def _mix():
    """tests/c/api2_driver.c:test_mix"""
    K, d = 3, 4
    w = np.array([0.5, 0.3, 0.2])
    mean = np.array([[3.0 * k - 2.0 + 0.5 * a for a in range(d)] for k in range(K)])
    cov = np.array([[[1.0 + 0.5 * k + 0.1 * a if a == b else 0.2 / (1.0 + abs(a - b)) for b in range(d)] for a in range(d)] for k in range(K)])
    return w, mean, cov


def test_wire_format_and_helpers(api2, tmp_path):
    out = tmp_path / "wire.txt"
    r = subprocess.run([api2, "wire", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rep = {}
    lines = open(out).read().splitlines()
    for l in lines:
        f = l.split()
        rep[f[0]] = f[1:]
    K, d = 3, 4
    want = 2 * 8 + K * (8 + 4 + 8) + K * d * (1 + d) * 8  # mvdens.c:585-594
    assert rep["size"] == [str(want), "expected", str(want)]
    assert rep["roundtrip"] == ["1"]
    assert rep["wire"][:6] == ["K", "3", "d", "4", "w0", "0.5"] and rep["wire"][-3:] == ["-1", "chol", "1"]
    assert rep["badsize"] == ["null", "1", "code", "-709"]          # mv_serialize
    assert rep["copy"] == ["weights", "1", "shares_components", "1"]  # the reference copies the component POINTERS
    w, mean, cov = _mix()
    assert_close(float(rep["enc"][0]), 1.0 / np.sum(w ** 2), RTOL, what="effective_number_of_components")
    L0 = np.linalg.cholesky(cov[0])
    assert_close(float(rep["lndet"][0]), np.sum(np.log(np.diag(L0))), 1e-13, what="ln_determinant")
    assert_close(float(rep["lndet"][2]), np.prod(np.diag(L0)), 1e-13, what="determinant")
    assert float(rep["inverse"][1]) < 1e-14 and rep["inverse"][3] == "0"
    assert float(rep["dump_read"][1]) < 1e-9 and rep["dump_read"][3] == "1"  # "% .10e" text format
    assert "name sigma8 1 h 2" in lines
    assert any(l.startswith("name missing -1 code -6701") for l in lines)       # dist_undef
    v, wv = float(rep["prior_name"][1]), float(rep["prior_name"][3])
    assert abs(v - wv) < 1e-9  # the reference's ln2pi is truncated to 13 digits (maths_base.h:49)
    # newError(v, where, text, linkTo, NULL) appends to linkTo's chain and returns linkTo (errorlist.c:44-50)
    assert rep["pickle"][:6] == ["n", "2", "top", "-42", "next", "-43"] and rep["pickle"][-1] == "second"
    assert any(l.startswith("Face |") for l in lines) and sum(l.strip().startswith(("0 |", "1 |")) for l in lines) == 2
    assert rep["sort"] == ["5", "1", "3", "0", "4", "2"]


def test_single_sample_draws_are_the_right_distributions(api2, tmp_path):
    """mix_mvdens_ran (mvdens.c:730-764) and mvdens_ran with a Student-t component (mvdens.c:279-317: x = mu + L g
    sqrt(nu / chi2_nu)), one call per sample as pmc_rc.c:537-556 draw_from does.  The Philox stream differs from mt19937,
    so the check is statistical: component frequencies, per-component means, and for the Student-t the exact law of the
    Mahalanobis radius, (x-mu)^T Sigma^-1 (x-mu) / d ~ F(d, nu)."""
    M, nu = 8000, 5
    out = tmp_path / "ran.bin"
    r = subprocess.run([api2, "ran", str(out), str(M), str(nu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "init_cwght 1 cw_last 1" in r.stdout
    raw = np.fromfile(out, np.uint8)
    d = 4
    x = raw[:M * d * 8].view(np.float64).reshape(M, d)
    idx = raw[M * d * 8:M * d * 8 + M * 8].view(np.uint64).astype(np.int64)
    t = raw[M * d * 8 + M * 8:].view(np.float64).reshape(M, d)
    w, mean, cov = _mix()
    cnt = np.bincount(idx, minlength=3)
    chi2 = np.sum((cnt - M * w) ** 2 / (M * w))
    assert stats.chi2.sf(chi2, 2) > 1e-4, (cnt, chi2)
    for k in range(3):
        xs = x[idx == k]
        z = (xs.mean(0) - mean[k]) / np.sqrt(np.diag(cov[k]) / len(xs))
        assert np.all(np.abs(z) < 5.0), (k, z)
        q = np.einsum("ia,ab,ib->i", xs - mean[k], np.linalg.inv(cov[k]), xs - mean[k])
        assert stats.kstest(q, stats.chi2(d).cdf).pvalue > 1e-4
    mu = 1.0 + np.arange(d)
    S = np.where(np.eye(d) > 0, 2.0, 0.5)
    q = np.einsum("ia,ab,ib->i", t - mu, np.linalg.inv(S), t - mu) / d
    ks = stats.kstest(q, stats.f(d, nu).cdf)
    assert ks.pvalue > 1e-4, ks
    # heavy tails really are there: a Gaussian would put ~0.03 % of the mass beyond q*d > 21
    assert np.mean(q * d > 21.0) > 0.01


def test_pmc_simu_from_file(api2, tmp_path):
    """The resume path (pmc.c:373-445): an "lc" text sample (log-weight, -component, parameters per line) is read back,
    log_rho recomputed from the proposal on the device, weights normalised."""
    rng = np.random.default_rng(5)
    N, d = 700, 3
    X = rng.normal(0.0, 1.5, (N, d))
    lw = rng.normal(-3.0, 2.0, N)
    comp = rng.integers(0, 2, N)
    lc = tmp_path / "sample.lc"
    with open(lc, "w") as f:
        f.write("# weight -icomp param\n")
        f.write(f"# npar = {d}, n_ded = 0\n")
        for i in range(N):
            f.write(f"{lw[i]:.17e} {-float(comp[i]):.1f} " + " ".join(f"{v:.17e}" for v in X[i]) + "\n")
    out = tmp_path / "file.bin"
    r = subprocess.run([api2, "file", str(lc), str(out), str(N + 50), str(d)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad_header null 1 code -4012" in r.stdout  # tls_Nparam
    raw = np.fromfile(out, np.uint8)
    sc = raw[:40].view(np.float64)
    n = int(sc[0])
    assert n == N and int(sc[4]) == 0
    pos = 40
    gX = raw[pos:pos + n * d * 8].view(np.float64).reshape(n, d); pos += n * d * 8
    gi = raw[pos:pos + n * 8].view(np.uint64); pos += n * 8
    gw = raw[pos:pos + n * 8].view(np.float64); pos += n * 8
    gr = raw[pos:pos + n * 8].view(np.float64); pos += n * 8
    gf = raw[pos:pos + n * 2].view(np.int16)
    assert np.array_equal(gX, X) and np.array_equal(gi, comp.astype(np.uint64)) and np.all(gf == 1)
    # the proposal of api2_driver.c:mode_file, evaluated by the oracle
    cov = np.array([np.where(np.eye(d) > 0, 2.0 + k, 0.1) for k in range(2)])
    mean = np.array([np.full(d, -1.0), np.full(d, 1.0)])
    m = O.Mix(np.array([0.6, 0.4]), mean, cov, np.full(2, -1, np.int32))
    L = O.orc()
    tmp, ld, rc = np.zeros(d), np.zeros(2), O.C.c_int(0)
    want_r = np.array([L.orc_mix_log_pdf(O.C.byref(m.c), O.P(np.ascontiguousarray(X[i])), O.P(tmp), O.P(ld), O.C.byref(rc)) for i in range(N)])
    assert_close(gr, want_r, RTOL, what="log_rho")
    assert_close(sc[3], want_r.max(), RTOL, what="maxR")
    assert sc[2] == lw.max()
    e = np.exp(lw - lw.max() + 500.0)
    assert_close(gw, e / e.sum(), WTOL, what="normalised weights")
    assert_close(sc[1], np.log(e.sum()) + lw.max() - 500.0, RTOL, scale=10.0, what="logSum")


def test_point_wise_entry_points(api2, tmp_path):
    """mvdens_log_pdf (Gaussian and Student-t), scalar_product, mix_mvdens_log_pdf, bentGauss_lkl and isinBox called one
    point at a time, as an opaque user likelihood would (pmc_dropin.cpp): against the oracle's restatements.  Points close
    to the mean are included: scalar_product is the forward substitution itself, so small q keeps its relative accuracy."""
    rng = np.random.default_rng(3)
    M, d = 160, 6
    a = rng.normal(0, 1, (d, d))
    cov = a @ a.T / d + 0.7 * np.eye(d)
    mean = rng.normal(0, 1, d)
    X = mean + rng.normal(0, 1.2, (M, d))
    X[:20] = mean + rng.normal(0, 1e-4, (20, d))   # tiny q
    with open(tmp_path / "pt.in", "wb") as f:
        f.write(np.array([M, d], np.int64).tobytes()); f.write(X.tobytes()); f.write(mean.tobytes()); f.write(cov.tobytes())
    r = subprocess.run([api2, "point", str(tmp_path / "pt.in"), str(tmp_path / "pt.out")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "df_after 4 chol 1" in r.stdout
    got = np.fromfile(tmp_path / "pt.out").reshape(M, 7)
    L = O.orc()
    g = O.Mix(np.ones(1), mean[None], cov[None], np.array([-1], np.int32))
    t = O.Mix(np.ones(1), mean[None], cov[None], np.array([4], np.int32))
    mm = O.Mix(np.array([0.25, 0.75]), np.array([mean - 0.5, mean + 0.5]), np.array([cov, cov]), np.full(2, -1, np.int32))
    tmp, ld, rc = np.zeros(d), np.zeros(2), O.C.c_int(0)
    want = np.zeros((M, 7))
    for i in range(M):
        x = np.ascontiguousarray(X[i])
        want[i, 0] = L.orc_mvdens_log_pdf(d, O.P(g.mean), O.P(g.L), O.C.c_double(g.detL[0]), -1, O.P(x), O.P(tmp))
        want[i, 1] = L.orc_scalar_product(d, O.P(g.mean), O.P(g.L), O.P(x), O.P(tmp))
        want[i, 2] = L.orc_mvdens_log_pdf(d, O.P(t.mean), O.P(t.L), O.C.c_double(t.detL[0]), 4, O.P(x), O.P(tmp))
        want[i, 3] = want[i, 1]
        want[i, 4] = L.orc_mix_log_pdf(O.C.byref(mm.c), O.P(x), O.P(tmp), O.P(ld), O.C.byref(rc))
        want[i, 5] = L.orc_bentgauss(d, O.C.c_double(0.1), O.C.c_double(10.0), O.P(x))
        want[i, 6] = float(-1.0 <= x[0] <= 1.5 and x[1] <= 2.0)
    for j, name in enumerate(["mvdens_log_pdf", "scalar_product", "mvdens_log_pdf student-t", "scalar_product (t)",
                              "mix_mvdens_log_pdf", "bentGauss_lkl"]):
        if "scalar" in name:
            e = float(np.max(np.abs(got[:, j] - want[:, j]) / want[:, j]))   # element-wise relative, tiny q included
            assert e <= 1e-11, (name, e)   # q = |y|^2 of a forward substitution: error ~ eps * cond(L), independent of |q|
        else:
            e = assert_close(got[:, j], want[:, j], RTOL, what=name)
        print(name, "max rel err", e)
    assert np.array_equal(got[:, 6], want[:, 6])
