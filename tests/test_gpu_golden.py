"""GPU against the committed golden fixtures (outputs of the unmodified reference, tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from pmclib_b200 import workloads as W
from util import PTOL, RTOL, WTOL, assert_close, compare_mixture

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FACT = {"c1_banana_d10_K10": W.banana, "c2_gmm_d5_K20": W.gmm5, "c4_student_d30_K50": W.student30,
        "c5_gauss_d128_K64": W.gauss128, "shipped_d2_K9": W.shipped_example}


@pytest.mark.parametrize("path", ["generic", "fused"])
@pytest.mark.parametrize("name", sorted(FACT))
def test_iteration_matches_reference_golden(gpu, name, path):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    wl = FACT[name]()
    X, idx = g["X"], g["indices"]
    gpu.set_option("force_generic", 1 if path == "generic" else 0)
    gpu.set_option("fused_version", 0)
    gpu.set_option("spl", 1)
    gpu.set_workload(wl)
    gpu.resize(len(X), wl["d"])
    gpu.upload(X=X, indices=idx, flg=np.ones(len(X), np.int16))
    big = max(np.max(np.abs(g["log_rho"])), 1.0)
    upd = int(g["do_update"])
    if int(g["rc"]) != 0:
        with pytest.raises(Exception) as e:
            gpu.step_given(0, len(X), 1.0, upd, 0)
        assert str(int(g["rc"])) in str(e.value)   # same reference error code (pmc_negWeight for the tiny c5 sample)
        # the importance part is still well defined (the failed update has consumed the proposal, as in the reference)
        gpu.set_workload(wl)
        gpu.upload(X=X, indices=idx, flg=np.ones(len(X), np.int16))
        r = gpu.step_given(0, len(X), 1.0, 0, 0)
    else:
        r = gpu.step_given(0, len(X), 1.0, upd, 0)
    got = gpu.download(X=False, indices=False)
    assert r.nok == int(g["nok"]) and np.array_equal(got["flg"], g["flg"])
    assert_close(got["log_rho"], g["log_rho"], RTOL, what="log_rho")
    assert_close(got["weights"], g["w_norm"], WTOL, scale=np.max(g["w_norm"]) * 1e-3, what="normalised weights")
    assert_close(r.maxW, float(g["maxW"]), RTOL, scale=big, what="maxW")
    assert_close(r.maxR, float(g["maxR"]), RTOL, what="maxR")
    assert_close(r.logSum, float(g["logSum"]), RTOL, scale=big, what="logSum")
    assert_close(r.ln_evidence, float(g["ln_evidence"]), RTOL, scale=big, what="evidence")
    assert_close(r.perplexity, float(g["perp"]), PTOL, what="perplexity")
    assert_close(r.ess, float(g["ess"]), PTOL, what="ess")
    if upd and int(g["rc"]) == 0:
        assert r.retry == int(g["retry"])
        compare_mixture(r.wght, r.mean, r.L, g["o_wght"], g["o_mean"], g["o_std"], what=name)


@pytest.mark.parametrize("path", ["tensor", "generic"])
def test_c5_update_matches_reference_digest(gpu, path):
    """configs[4] shape (d=128, K=64) with every component above MINCOUNT: the update of the unmodified reference, kept as
    a digest (tests/golden/make_golden.py:update_digest), against the DMMA one-pass update and the literal kernels."""
    import hashlib
    import oracle_py as O
    from golden.make_golden import sample_entries
    from util import relerr
    g = np.load(os.path.join(GOLD, "c5_update_d128_K64.npz"))
    wl = W.gauss128()
    X, idx = O.orc_simulate(int(g["N"]), wl, seed=int(g["seed"]))
    assert hashlib.sha256(X.tobytes()).hexdigest() == str(g["x_sha"])
    gpu.set_option("force_generic", 1 if path == "generic" else 0)
    gpu.set_option("fused_version", 0)
    gpu.set_option("force_precise", 0)
    gpu.set_workload(wl)
    gpu.resize(len(X), wl["d"])
    gpu.upload(X=X, indices=idx, flg=np.ones(len(X), np.int16))
    r = gpu.step_given(0, len(X), 1.0, 1, 0)
    got = gpu.download(X=False, indices=False)
    big = max(np.max(np.abs(g["log_rho_head"])), 1.0)
    assert r.nok == int(g["nok"]) and r.retry == int(g["retry"])
    assert_close(got["log_rho"][:512], g["log_rho_head"], RTOL, what="log_rho")
    assert_close(got["weights"][:512], g["w_norm_head"], RTOL, scale=np.max(g["w_norm_head"]) * 1e-3, what="normalised weights")
    assert_close(r.logSum, float(g["logSum"]), RTOL, scale=big, what="logSum")
    assert_close(r.perplexity, float(g["perp"]), RTOL, what="perplexity")
    assert_close(r.ess, float(g["ess"]), RTOL, what="ess")
    assert np.array_equal(r.wght > 0, g["o_wght"] > 0)
    assert_close(r.wght, g["o_wght"], RTOL, scale=np.max(g["o_wght"]), what="weights of the updated mixture")
    K, d = wl["K"], wl["d"]
    kk, aa, bb = sample_entries(K, d, len(g["o_std_sel"]))
    em = max(relerr(r.mean[k], g["o_mean"][k], scale=max(g["o_std_max"][k], np.max(np.abs(g["o_mean"][k])))) for k in range(K))
    ed = max(relerr(np.diag(r.L[k]), g["o_std_diag"][k], scale=g["o_std_max"][k]) for k in range(K))
    es = float(np.max(np.abs(r.L[kk, aa, bb] - g["o_std_sel"]) / g["o_std_max"][kk]))
    ef = relerr(np.sqrt((np.tril(r.L) ** 2).sum((1, 2))), g["o_std_fro"])
    edet = relerr(np.log(r.detL), np.log(g["o_detL"]), scale=1.0)
    print(f"c5 update vs reference digest [{path}] precise={r.used_precise_path}: mean {em:.2e} diag {ed:.2e} entries {es:.2e} fro {ef:.2e} logdet {edet:.2e}")
    assert em <= RTOL and ed <= 10 * RTOL and es <= 10 * RTOL and ef <= 10 * RTOL and edet <= 1e-10


def test_misc_golden(gpu):
    g = np.load(os.path.join(GOLD, "misc.npz"))
    wl = W.banana()
    gpu.set_workload(wl)
    assert np.array_equal(gpu.ranindx(g["z"]), g["ranindx"])
    gpu.set_box(g["faces"])
    assert np.array_equal(gpu.isinbox(g["box_x"]), g["box_in"])
