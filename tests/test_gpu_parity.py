"""GPU parity: the CUDA path (through the pmcgpu_* C-ABI) against the CPU oracle on the same seeded inputs.

Tolerance: 1e-12 relative for floating point (BASELINE.json north_star), bit-exact for indices, flags and counts.
Sizes are chosen so the oracle finishes in seconds; full-size behaviour is covered by test_gpu_properties.py.
"""
import numpy as np
import pytest

import oracle_py as O
from pmclib_b200 import workloads as W
from util import PTOL, RTOL, WTOL, assert_close, compare_mixture, relerr

pytestmark = pytest.mark.gpu

CASES = {
    # name: (workload factory, N)
    "c1_banana_d10_K10": (lambda: W.banana(), 20000),
    "c2_gmm_d5_K20": (lambda: W.gmm5(), 20000),
    "c4_student_d30_K50": (lambda: W.student30(), 4000),
    "c5_gauss_d128_K64": (lambda: W.gauss128(), 4000),
    "shipped_d2_K9": (lambda: W.shipped_example(), 30000),
    "ragged_d10_K10": (lambda: W.banana(seed=77), 1237),   # not a multiple of any tile size
    "tiny_d10_K10": (lambda: W.banana(seed=78), 33),
    "d3_K4": (lambda: W.banana(d=3, K=4, seed=5, box=30.0), 5000),
    "d4_K6": (lambda: W.banana(d=4, K=6, seed=6, mean_sigma=2.0, var=6.0, box=30.0), 6000),
    "d6_K12": (lambda: W.banana(d=6, K=12, seed=7, mean_sigma=2.0, var=6.0, box=30.0), 12000),
    "d8_K20": (lambda: W.banana(d=8, K=20, seed=8, mean_sigma=2.0, var=6.0, box=30.0), 20000),
    "d10_K20": (lambda: W.banana(d=10, K=20, seed=10, mean_sigma=2.0, var=6.0, box=40.0), 30000),
    "d10_K5": (lambda: W.banana(d=10, K=5, seed=11, mean_sigma=2.0, var=6.0, box=40.0), 8000),
    "d5_K10": (lambda: W.banana(d=5, K=10, seed=12, mean_sigma=2.0, var=6.0, box=30.0), 10000),
    "d7_K3_generic_only": (lambda: W.banana(d=7, K=3, seed=9, mean_sigma=2.0, var=6.0, box=30.0), 3000),
    # combined targets (combine_lkl of a density and a Gaussian prior, distribution.c:350-396, 548-606), evaluated by
    # k_combine_target and handed to the weight kernels
    "banana_prior_d10_K10": (lambda: W.banana_prior(), 20000),
    "gmm_prior_d5_K20": (lambda: W.gmm_prior(), 20000),
    "d7_K3_prior": (lambda: W.with_gaussian_prior(W.banana(d=7, K=3, seed=9, mean_sigma=2.0, var=6.0, box=30.0), [2, 5], [0.0, 1.0], [4.0, 9.0]), 3000),
}


def _setup(gpu, wl, X, idx, force_generic, spl=1, ver=0, w3v=3):
    gpu.set_option("w3_variant", w3v)
    gpu.set_option("force_generic", force_generic)
    gpu.set_option("fused_version", ver)
    gpu.set_option("force_precise", 0)
    gpu.set_option("spl", spl)
    gpu.set_workload(wl)
    gpu.resize(len(X), wl["d"])
    gpu.upload(X=X, indices=idx, flg=np.ones(len(X), np.int16))


@pytest.fixture(scope="module")
def oracle_runs():
    cache = {}

    def get(name):
        if name not in cache:
            f, N = CASES[name]
            wl = f()
            X, idx = O.orc_simulate(N, wl, seed=11)
            o = O.orc_iteration(X, idx, wl, do_update=0 if wl["importance_only"] else 1)
            cache[name] = (wl, X, idx, o)
        return cache[name]
    return get


# fused_v3: the third-generation weights kernel with the row-scaled solve (the default, option w3_variant = 3);
# fused_v3_exact: the same kernel with the exact per-component mean subtraction (w3_variant = 0, the guard's fallback)
PATHS = {"generic": (1, 0, 3), "fused_v1": (0, 1, 3), "fused_v2": (0, 2, 3), "fused_v3": (0, 3, 3), "fused_v3_exact": (0, 3, 0)}


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("name", list(CASES))
def test_weights_normalise_perplexity(gpu, oracle_runs, name, path):
    wl, X, idx, o = oracle_runs(name)
    _setup(gpu, wl, X, idx, PATHS[path][0], 1, PATHS[path][1], PATHS[path][2])
    nok, maxW, maxR = gpu.weights(1.0)
    got = gpu.download(X=False, indices=False)
    assert nok == o["nok"]
    assert np.array_equal(got["flg"], np.ones(len(X), np.int16))
    big = max(np.max(np.abs(o["log_rho"])), 1.0)
    assert_close(got["log_rho"], o["log_rho"], RTOL, what="log_rho")
    # a log-weight is a difference of two log-densities: the error is relative to their size, not to the difference
    assert_close(got["weights"], o["logw"], RTOL, scale=big, what="log-weights")
    assert_close(maxW, o["maxW"], RTOL, scale=big, what="maxW")
    assert_close(maxR, o["maxR"], RTOL, what="maxR")
    logSum, cnt = gpu.normalize(maxW)
    assert cnt == o["nok"]
    assert_close(logSum, o["logSum"], RTOL, scale=big, what="logSum")
    got = gpu.download(X=False, indices=False, log_rho=False)
    assert np.array_equal(got["flg"], o["flg"])
    assert_close(got["weights"], o["w_norm"], WTOL, scale=np.max(o["w_norm"]) * 1e-3, what="normalised weights")
    perp, ess, nfl = gpu.perplexity_ess()
    gpu.set_option("w3_variant", 3)  # the context is shared by the whole session: back to the default
    assert nfl == int(np.sum(o["flg"] == 0))
    assert_close(perp, o["perp"], PTOL, what="perplexity")
    assert_close(ess, o["ess"], PTOL, what="ESS")


@pytest.mark.parametrize("name", [n for n in CASES if not CASES[n][0]()["importance_only"]])
def test_update_rb_two_pass(gpu, oracle_runs, name):
    """update_prop_rb on the oracle's own normalised sample (weights, log_rho, flags uploaded as the reference left them)."""
    wl, X, idx, o = oracle_runs(name)
    _setup(gpu, wl, X, idx, 1)
    gpu.upload(weights=o["w_norm"], log_rho=o["log_rho"], flg=o["flg"])
    if o["rc"] != 0:
        with pytest.raises(Exception) as e:
            gpu.update_rb(o["maxR"], len(X), 0)
        assert str(o["rc"]) in str(e.value)
        return
    cnt = gpu.count_indices()
    assert np.array_equal(cnt, np.bincount(idx[o["flg"] == 1].astype(np.int64), minlength=wl["K"]).astype(np.uint64))
    u = gpu.update_rb(o["maxR"], len(X), 0)
    assert u["retry"] == o["retry"]
    errs = compare_mixture(u["wght"], u["mean"], u["L"], o["o_wght"], o["o_mean"], o["o_std"], what=name)
    print(name, "two-pass update errors", errs)


# (kernel generation, samples per lane): the second- and third-generation kernels are instantiated for one sample per lane only
@pytest.mark.parametrize("ver,spl", [(1, 1), (1, 2), (2, 1), (3, 1)])
@pytest.mark.parametrize("name", ["c1_banana_d10_K10", "c2_gmm_d5_K20", "shipped_d2_K9", "ragged_d10_K10", "tiny_d10_K10",
                                  "d3_K4", "d4_K6", "d6_K12", "d8_K20", "d10_K20", "d10_K5", "d5_K10", "d7_K3_generic_only",
                                  "banana_prior_d10_K10", "gmm_prior_d5_K20", "d7_K3_prior"])
def test_fused_iteration_given_samples(gpu, oracle_runs, name, ver, spl):
    """The one-pass fused kernel (weights + statistics) + host finalisation against the oracle's full iteration."""
    wl, X, idx, o = oracle_runs(name)
    _setup(gpu, wl, X, idx, 0, spl, ver)
    if o["rc"] != 0:
        with pytest.raises(Exception):
            gpu.step_given(0, len(X), 1.0, 1, 0)
        return
    r = gpu.step_given(0, len(X), 1.0, 1, 0)
    big = max(np.max(np.abs(o["log_rho"])), 1.0)
    assert r.nok == o["nok"]
    assert_close(r.maxW, o["maxW"], RTOL, scale=big, what="maxW")
    assert_close(r.maxR, o["maxR"], RTOL, what="maxR")
    assert_close(r.logSum, o["logSum"], RTOL, scale=big, what="logSum")
    assert_close(r.ln_evidence, o["ln_evidence"], RTOL, scale=big, what="ln evidence")
    assert_close(r.perplexity, o["perp"], PTOL, what="perplexity")
    assert_close(r.ess, o["ess"], PTOL, what="ESS")
    assert r.retry == o["retry"]
    got = gpu.download(X=False, indices=False)
    assert np.array_equal(got["flg"], o["flg"])
    assert_close(got["log_rho"], o["log_rho"], RTOL, what="log_rho")
    assert_close(got["weights"], o["w_norm"], WTOL, scale=np.max(o["w_norm"]) * 1e-3, what="normalised weights")
    errs = compare_mixture(r.wght, r.mean, r.L, o["o_wght"], o["o_mean"], o["o_std"], what=name)
    print(name, f"ver={ver} spl={spl} precise={r.used_precise_path} one-pass update errors", errs)


def test_importance_only_student(gpu, oracle_runs):
    """configs[3] shape: weights + perplexity + evidence, no update (vanilla_importance.c:56-82)."""
    wl, X, idx, o = oracle_runs("c4_student_d30_K50")
    _setup(gpu, wl, X, idx, 0)
    r = gpu.step_given(0, len(X), 1.0, 0, 0)
    big = max(np.max(np.abs(o["log_rho"])), 1.0)
    assert_close(r.logSum, o["logSum"], RTOL, scale=big, what="logSum")
    assert_close(r.ln_evidence, o["ln_evidence"], RTOL, scale=big, what="ln evidence")
    assert_close(r.perplexity, o["perp"], PTOL, what="perplexity")
    assert_close(r.ess, o["ess"], PTOL, what="ESS")


def test_row_scaled_solve_and_its_guard(gpu):
    """The third-generation weights kernel uses the row-scaled solve while every live component has |mu'_j / L_jj| (about the
    mixture mean) below option gen3_max_ratio, and the exact mean subtraction beyond it: two narrow clusters 1000 widths
    apart trip the guard; both forms are held to the same bar against the oracle."""
    base = W.banana(seed=31, mean_sigma=2.0, var=6.0, box=40.0)
    far = dict(base)
    far["mean"] = base["mean"] * 0.05
    far["mean"][: base["K"] // 2] += 25.0
    far["mean"][base["K"] // 2:] -= 25.0
    far["cov"] = base["cov"] * (0.05 ** 2 / 6.0)
    for wl, expect in ((base, 1), (far, 0)):
        X, idx = O.orc_simulate(20000, wl, seed=17)
        o = O.orc_iteration(X, idx, wl, do_update=0)
        _setup(gpu, wl, X, idx, 0, 1, 3, 3)
        assert int(gpu.get_info("w3_rowscaled")) == expect, gpu.get_info("w3_shift_ratio")
        nok, maxW, maxR = gpu.weights(1.0)
        got = gpu.download(X=False, indices=False)
        assert nok == o["nok"]
        big = max(np.max(np.abs(o["log_rho"])), 1.0)
        assert_close(got["log_rho"], o["log_rho"], RTOL, what="log_rho")
        assert_close(got["weights"], o["logw"], RTOL, scale=big, what="log-weights")
    gpu.set_option("gen3_max_ratio", 1.0)  # the option moves the guard
    gpu.set_workload(base)
    assert int(gpu.get_info("w3_rowscaled")) == 0
    gpu.set_option("gen3_max_ratio", 100.0)
    gpu.set_workload(base)
    assert int(gpu.get_info("w3_rowscaled")) == 1


def test_ranindx_and_box_bit_exact(gpu):
    wl = W.banana()
    gpu.set_option("force_generic", 0)
    gpu.set_workload(wl)
    rng = np.random.default_rng(3)
    m = O.Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"])
    cw = np.zeros(wl["K"])
    O.orc().orc_cwght(wl["K"], O.P(m.wght), O.P(cw))
    z = np.concatenate([rng.random(100000), cw, np.nextafter(cw, 0), np.nextafter(cw, 2), [0.0, 1.0 - 2**-53]])
    ref = np.array([O.orc().orc_ranindx(O.P(cw), wl["K"], float(v)) for v in z], np.uint64)
    assert np.array_equal(gpu.ranindx(z), ref)
    # box: points on, just inside and just outside every face
    d = wl["d"]
    x = rng.normal(0, 25, (20000, d))
    x[:2000, 0] = np.repeat([40.0, np.nextafter(40.0, 50), np.nextafter(40.0, 0), -40.0], 500)
    faces = wl["faces"]
    ref = np.array([O.orc().orc_isinbox(d, len(faces), O.P(faces), O.P(np.ascontiguousarray(r))) for r in x], np.int32)
    assert np.array_equal(gpu.isinbox(x), ref)
    # general (non axis-parallel) faces
    gf = np.concatenate([faces, rng.normal(0, 1, (5, d + 1)) + np.r_[np.zeros(d), 30.0]])
    gpu.set_box(gf)
    ref = np.array([O.orc().orc_isinbox(d, len(gf), O.P(gf), O.P(np.ascontiguousarray(r))) for r in x], np.int32)
    assert np.array_equal(gpu.isinbox(x), ref)


def test_log_pdf_entry_points(gpu, oracle_runs):
    for name in ("c1_banana_d10_K10", "c4_student_d30_K50"):
        wl, X, idx, o = oracle_runs(name)
        gpu.set_workload(wl)
        assert_close(gpu.proposal_log_pdf(X[:3000]), o["log_rho"][:3000], RTOL, what="proposal_log_pdf")
        post = gpu.target_log_pdf(X[:3000])
        assert_close(post - o["log_rho"][:3000], o["logw"][:3000], RTOL, scale=np.max(np.abs(o["log_rho"])), what="target")


@pytest.mark.parametrize("d,K,df,N", [(5, 4, 5, 20000), (10, 6, 9, 30000), (24, 5, 7, 30000)])
def test_student_t_update(gpu, d, K, df, N):
    """Rao-Blackwellised update of a Student-t proposal (gamma-weighted statistics, pmc.c:1386-1398, 1459-1460: the
    covariance is divided by weight_d, not weight_gamma_d) against the oracle's full iteration."""
    wl = W.gauss128(d=d, K=K, seed=40 + d)
    wl["df"] = np.full(K, df, np.int32)
    X, idx = O.orc_simulate(N, wl, seed=19)
    o = O.orc_iteration(X, idx, wl, do_update=1)
    assert o["rc"] == 0
    for mma in (0, 1):
        gpu.set_option("force_generic", 0)
        gpu.set_option("force_precise", 0)
        gpu.set_option("mma", mma)
        gpu.set_workload(wl)
        gpu.resize(N, d)
        gpu.upload(X=X, indices=idx, flg=np.ones(N, np.int16))
        r = gpu.step_given(0, N, 1.0, 1, 0)
        assert_close(r.perplexity, o["perp"], PTOL, what="perplexity")
        assert r.retry == o["retry"]
        errs = compare_mixture(r.wght, r.mean, r.L, o["o_wght"], o["o_mean"], o["o_std"], what=f"student d={d} mma={mma}")
        print(f"student-t d={d} K={K} df={df} mma={mma} precise={r.used_precise_path}", errs)


@pytest.mark.parametrize("name,generic", [("c1_banana_d10_K10", 0), ("c2_gmm_d5_K20", 0), ("d7_K3_generic_only", 0),
                                          ("d7_K3_generic_only", 1), ("c5_gauss_d128_K64", 0)])
def test_tempered_weights(gpu, oracle_runs, name, generic):
    """t_iter_tempered (pmc.c:570): log-weight = t * log-posterior - log-proposal, on every kernel family; the option
    "t_temper" gives pmcgpu_pmc_step the same temperature."""
    wl, X, idx, _ = oracle_runs(name)
    t = 0.37
    o = O.orc_iteration(X, idx, wl, t=t, do_update=1)
    _setup(gpu, wl, X, idx, generic)
    r = gpu.step_given(0, len(X), t, 1, 0)
    big = max(np.max(np.abs(o["log_rho"])), 1.0)
    assert_close(r.maxW, o["maxW"], RTOL, scale=big, what="maxW")
    assert_close(r.logSum, o["logSum"], RTOL, scale=big, what="logSum")
    assert_close(r.perplexity, o["perp"], PTOL, what="perplexity")
    compare_mixture(r.wght, r.mean, r.L, o["o_wght"], o["o_mean"], o["o_std"], what=name)
    # the drawing iteration with the option: same temperature reaches the weight kernels
    gpu.set_workload(wl)
    gpu.set_option("t_temper", t)
    try:
        a = gpu.pmc_step(3, 0, do_update=0)
    finally:
        gpu.set_option("t_temper", 1.0)
    got = gpu.download()
    o2 = O.orc_iteration(got["X"], got["indices"], wl, t=t, do_update=0)
    assert_close(a.logSum, o2["logSum"], RTOL, scale=max(np.max(np.abs(o2["log_rho"])), 1.0), what="logSum (drawn)")
    assert_close(a.perplexity, o2["perp"], PTOL, what="perplexity (drawn)")


@pytest.mark.parametrize("name", ["banana_prior", "gmm_prior"])
def test_combined_target_log_pdf_golden(gpu, name):
    """pmcgpu_target_log_pdf of a combined target against the REAL reference's add_gaussian_prior values (golden)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "combine.npz"))
    wl = getattr(W, name)()
    gpu.set_workload(wl)
    got = gpu.target_log_pdf(g[name + "_X"])
    e = assert_close(got, g[name + "_post"], RTOL, what="combined target")
    print(name, "combined target max rel err", e)
