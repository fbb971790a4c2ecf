"""GPU: the update variants around update_prop_rb — update_prop (hard-assignment EM, pmc.c:1194-1310), band-limited
covariances (pmc.c:1288, 1444) — and conditional ("default") target parameters (distribution.c:193-225), against the
oracle and, through tests/golden/updates.npz, against outputs of the compiled reference."""
import os

import numpy as np
import pytest

import oracle_py as O
from pmclib_b200 import workloads as W
from util import PTOL, RTOL, WTOL, assert_close, compare_mixture

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = [("hard_c1", "banana", 4000, 2, None), ("hard_c2", "gmm5", 6000, 2, None), ("band_c1", "banana", 4000, 1, 3),
         ("hardband_c2", "gmm5", 6000, 2, 2)]


@pytest.mark.parametrize("name,fact,N,upd,band", CASES)
def test_hard_and_band_limited_updates(gpu, name, fact, N, upd, band, capfd):
    g = np.load(os.path.join(GOLD, "updates.npz"))
    wl = getattr(W, fact)()
    if band is not None:
        wl["band_limit"] = np.full(wl["K"], band, np.uint64)
    X, idx = O.orc_simulate(N, wl, seed=401)
    o = O.orc_iteration(X, idx, wl, do_update=0)
    gpu.set_option("force_generic", 0); gpu.set_option("fused_version", 0); gpu.set_option("force_precise", 0)
    gpu.set_workload(wl)
    gpu.set_band_limit(wl.get("band_limit"))
    gpu.resize(N, wl["d"])
    gpu.upload(X=X, indices=idx, weights=o["w_norm"], log_rho=o["log_rho"], flg=o["flg"])
    u = gpu.update_rb(o["maxR"], N, 0, hard=(upd == 2))
    capfd.readouterr()
    assert u["retry"] == int(g[f"{name}_retry"])
    errs = compare_mixture(u["wght"], u["mean"], u["L"], g[f"{name}_o_wght"], g[f"{name}_o_mean"], g[f"{name}_o_std"], what=name)
    print(name, "update errors vs the compiled reference's outputs", errs)
    gpu.set_band_limit(None)


@pytest.mark.parametrize("name,fact,N,band", [("band_c1", "banana", 4000, 3)])
def test_band_limited_fused_iteration(gpu, name, fact, N, band, capfd):
    """The whole iteration (one-pass statistics + host finalisation) with a band limit: the finalisation zeroes the
    out-of-band entries before factorising, and the retry logic follows the reference."""
    g = np.load(os.path.join(GOLD, "updates.npz"))
    wl = getattr(W, fact)()
    wl["band_limit"] = np.full(wl["K"], band, np.uint64)
    X, idx = O.orc_simulate(N, wl, seed=401)
    gpu.set_option("force_generic", 0); gpu.set_option("fused_version", 0); gpu.set_option("force_precise", 0)
    gpu.set_workload(wl)
    gpu.set_band_limit(wl["band_limit"])
    gpu.resize(N, wl["d"])
    gpu.upload(X=X, indices=idx, flg=np.ones(N, np.int16))
    r = gpu.step_given(0, N, 1.0, 1, 0)
    capfd.readouterr()
    assert r.retry == int(g[f"{name}_retry"])
    assert_close(r.perplexity, float(g[f"{name}_perp"]), PTOL, what="perplexity")
    compare_mixture(r.wght, r.mean, r.L, g[f"{name}_o_wght"], g[f"{name}_o_mean"], g[f"{name}_o_std"], what=name)
    gpu.set_band_limit(None)


@pytest.mark.parametrize("path", ["generic", "default"])
def test_conditional_target_parameters(gpu, path):
    """examples/pmc/test_pmc.lua verbatim: a 4-D bentGauss whose parameters 2 and 3 are pinned (distribution_set_default),
    sampled in the 2 free dimensions.  distribution_lkl fills the full vector (distribution.c:193-225) before the target
    sees it: log-weights against the oracle's 4-D target on the filled vectors; the pinned coordinates only add a
    constant, so normalised weights, perplexity and the update equal the oracle's 2-D run."""
    wl = W.shipped_example()
    N = 20000
    X, idx = O.orc_simulate(N, wl, seed=55)
    pins = np.array([0.7, -1.3])
    gpu.set_option("force_generic", 1 if path == "generic" else 0); gpu.set_option("fused_version", 0); gpu.set_option("force_precise", 0)
    gpu.set_workload(wl)
    gpu.set_target_banana(4, wl["target"]["b"], wl["target"]["sig1"])
    gpu.set_target_defaults([0, 0, 1, 1], [0.0, 0.0, pins[0], pins[1]])
    gpu.resize(N, 2)
    gpu.upload(X=X, indices=idx, flg=np.ones(N, np.int16))
    nok, maxW, maxR = gpu.weights(1.0)
    got = gpu.download(X=False, indices=False)
    o = O.orc_iteration(X, idx, wl, do_update=1)
    wl4 = dict(wl, target=dict(wl["target"], d=4))
    full = np.concatenate([X, np.tile(pins, (N, 1))], axis=1)
    post = O.orc_target_log_pdf(full, wl4)
    big = max(np.max(np.abs(o["log_rho"])), 1.0)
    assert nok == N
    assert_close(got["log_rho"], o["log_rho"], RTOL, what="log_rho")
    assert_close(got["weights"], post - o["log_rho"], RTOL, scale=big, what="log-weights with pinned parameters")
    # the point-wise entry agrees too
    assert_close(gpu.target_log_pdf(X[:500]), post[:500], RTOL, what="target_log_pdf with defaults")
    r = gpu.step_given(0, N, 1.0, 1, 0)
    got = gpu.download(X=False, indices=False)
    assert_close(got["weights"], o["w_norm"], WTOL, scale=np.max(o["w_norm"]) * 1e-3, what="normalised weights")
    assert_close(r.perplexity, o["perp"], PTOL, what="perplexity")
    compare_mixture(r.wght, r.mean, r.L, o["o_wght"], o["o_mean"], o["o_std"], what="update with pinned parameters")
    gpu.set_target_defaults(None, None)
