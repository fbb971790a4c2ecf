"""GPU parity of the weight post-processing hooks either side of the iteration: filter_histogram, the
shipped pmc_filter hook of pmc_simu_pmc_step (pmc.c:203-206, 709-770), and mean_from_psim (pmc.c:1108-1120)."""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
from pmclib_b200 import workloads as W
from util import PTOL, RTOL, WTOL, assert_close, compare_mixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["plain", "first_flagged", "w0_extreme", "none_ok", "nan_inside", "w0_nan"])
def test_filter_histogram_flags_bit_exact(gpu, case):
    rng = np.random.default_rng(17)
    N, d = 70001, 2
    w = rng.normal(-300.0, 25.0, N)
    flg = (rng.uniform(size=N) > 0.05).astype(np.int16)
    if case == "first_flagged":
        flg[:5] = 0
    if case == "w0_extreme":
        flg[0] = 0
        w[0] = 1e6
    if case == "none_ok":
        flg[:] = 0
    if case == "nan_inside":
        w[100] = np.nan
    if case == "w0_nan":
        w[0] = np.nan
    gpu.set_workload(W.shipped_example())
    gpu.resize(N, d)
    for nb, cl, cr in [(10, 1, 1), (50, 0, 3), (7, 2, 0)]:
        gpu.upload(X=np.zeros((N, d)), weights=w, flg=flg)
        n = gpu.filter_histogram(nb, cl, cr)
        f = flg.copy()
        n_ref = O.orc().orc_filter_histogram(C.c_long(N), O.P(w), O.P(f), nb, cl, cr)
        got = gpu.download(X=False, indices=False, log_rho=False)
        assert n == n_ref
        assert np.array_equal(got["flg"], f)


def test_weighted_mean(gpu):
    rng = np.random.default_rng(3)
    N, d = 50000, 6
    X = rng.normal(size=(N, d)); w = rng.uniform(size=N); w /= w.sum()
    flg = (rng.uniform(size=N) > 0.1).astype(np.int16)
    wl = W.banana(d=d, K=3, seed=1)
    gpu.set_workload(wl)
    gpu.resize(N, d)
    gpu.upload(X=X, weights=w, flg=flg)
    for a in range(d):
        ref = O.orc().orc_mean_from_psim(O.P(X), O.P(w), O.P(flg), C.c_long(N), d, a)
        assert_close(gpu.weighted_mean(a), ref, RTOL, scale=1.0, what=f"mean of coordinate {a}")


@pytest.mark.parametrize("name", ["c1", "c2", "d40"])
def test_iteration_with_filter(gpu, name):
    """pmc_simu_pmc_step with pmc_filter = filter_histogram: weights -> filter -> normalise -> RB update -> perplexity."""
    if name == "d40":
        wl = W.gauss128(d=40, K=6, seed=31)   # large-d tensor path: the statistics kernels see the filtered flags
        N = 9000
    else:
        wl = {"c1": W.banana, "c2": W.gmm5}[name]()
        N = 20000
    X, idx = O.orc_simulate(N, wl, seed=29)
    filt = (20, 1, 2)
    o = O.orc_iteration(X, idx, wl, do_update=1, filt=filt)
    assert o["rc"] == 0 and o["n_filtered"] > 0
    gpu.set_option("force_generic", 0)
    gpu.set_option("mma", 1)
    gpu.set_workload(wl)
    gpu.resize(N, wl["d"])
    gpu.upload(X=X, indices=idx, flg=np.ones(N, np.int16))
    gpu.set_filter_histogram(*filt)
    try:
        r = gpu.step_given(0, N, 1.0, 1, 0)
    finally:
        gpu.set_filter_histogram(0)
    got = gpu.download(X=False, indices=False)
    assert np.array_equal(got["flg"], o["flg"])
    big = max(np.max(np.abs(o["log_rho"])), 1.0)
    assert_close(r.logSum, o["logSum"], RTOL, scale=big, what="logSum")
    assert_close(r.perplexity, o["perp"], PTOL, what="perplexity")
    assert_close(r.ess, o["ess"], PTOL, what="ESS")
    assert r.retry == o["retry"]
    compare_mixture(r.wght, r.mean, r.L, o["o_wght"], o["o_mean"], o["o_std"], what=name)


@pytest.mark.parametrize("case", ["plain", "first_flagged", "early_gap", "w0_largest", "ties"])
@pytest.mark.parametrize("n", [1, 5, 37])
def test_clip_weights(gpu, case, n):
    """clip_weights (pmc.c:1134-1188): threshold and flags bit-exact, re-normalised weights to 1e-12 (the device sums in a
    different order)."""
    from test_cpu_oracle import _clip_case
    w, flg = _clip_case(case, N=50_000, seed=31)
    N = len(w)
    wo, fo = w.copy(), flg.copy()
    thr = C.c_double()
    rc = O.orc().orc_clip_weights(C.c_long(N), O.P(wo), O.P(fo), n, C.byref(thr))
    assert rc == 0
    gpu.set_workload(W.shipped_example())
    gpu.resize(N, 2)
    gpu.upload(X=np.zeros((N, 2)), weights=w, flg=flg)
    T, nclip, S = gpu.clip_weights(n)
    got = gpu.download(X=False, indices=False, log_rho=False)
    assert T == thr.value
    assert np.array_equal(got["flg"], fo)
    assert nclip == int((flg != 0).sum() - (fo != 0).sum())
    assert_close(got["weights"], wo, RTOL, scale=np.max(wo), what="clipped weights")


def test_clip_weights_nothing_left(gpu):
    N = 100
    w = np.full(N, 1.0 / N)
    gpu.set_workload(W.shipped_example())
    gpu.resize(N, 2)
    gpu.upload(X=np.zeros((N, 2)), weights=w, flg=np.ones(N, np.int16))
    with pytest.raises(Exception) as e:
        gpu.clip_weights(3)     # all weights equal the threshold: everything is clipped -> pmc_infnan
    assert "-6015" in str(e.value)
