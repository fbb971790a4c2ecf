#!/usr/bin/env python
"""Generates the golden fixtures from the UNMODIFIED reference (oracle/_ref/libpmcref.so, i.e. /root/reference's own C
files compiled against the GSL shim).  Run where /root/reference exists:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own for the PMC path, so these files are the pin:
inputs (X, indices, proposal, target) and every output of one iteration as the reference computed it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_py as O  # noqa: E402
from pmclib_b200 import workloads as W  # noqa: E402

CASES = {
    "c1_banana_d10_K10": (W.banana, 1500, 101),
    "c2_gmm_d5_K20": (W.gmm5, 2500, 102),
    "c4_student_d30_K50": (W.student30, 600, 103),
    "c5_gauss_d128_K64": (W.gauss128, 250, 104),
    "shipped_d2_K9": (W.shipped_example, 4000, 105),
}


C5U = dict(N=3200, seed=204, nsel=4000)


def sample_entries(K, d, nsel):
    """the (k, a, b) entries of the updated factors the digest keeps besides the diagonals"""
    rng = np.random.default_rng(2024)
    return rng.integers(0, K, nsel), rng.integers(0, d, nsel), rng.integers(0, d, nsel)


def update_digest():
    """configs[4] shape (d=128, K=64) with enough samples that every component passes count > MINCOUNT, so the RB update
    of the compiled reference is pinned too (the 250-sample c5 fixture ends in pmc_negWeight).  The inputs are regenerated
    from the seed by the oracle's deterministic draw; the 8 MB of updated factors are kept as a digest."""
    import hashlib
    wl = W.gauss128()
    X, idx = O.orc_simulate(C5U["N"], wl, seed=C5U["seed"])
    r = O.ref_iteration(X, idx, wl, do_update=1)
    assert r["rc"] == 0 and np.all(r["o_wght"] > 0)
    kk, aa, bb = sample_entries(wl["K"], wl["d"], C5U["nsel"])
    S = r["o_std"]
    out = dict(N=C5U["N"], seed=C5U["seed"], x_sha=hashlib.sha256(X.tobytes()).hexdigest(), idx_sha=hashlib.sha256(idx.tobytes()).hexdigest(),
               rc=r["rc"], nok=r["nok"], maxW=r["maxW"], maxR=r["maxR"], logSum=r["logSum"], perp=r["perp"], ess=r["ess"],
               ln_evidence=r["ln_evidence"], retry=r["retry"], o_wght=r["o_wght"], o_mean=r["o_mean"], o_detL=r["o_detL"],
               o_std_diag=np.array([np.diag(S[k]) for k in range(wl["K"])]), o_std_sel=S[kk, aa, bb],
               o_std_fro=np.sqrt((np.tril(S) ** 2).sum((1, 2))), o_std_max=np.abs(S).max((1, 2)),
               w_norm_head=r["w_norm"][:512], log_rho_head=r["log_rho"][:512])
    np.savez_compressed(os.path.join(HERE, "c5_update_d128_K64.npz"), **out)
    print("c5_update", r["rc"], r["perp"], os.path.getsize(os.path.join(HERE, "c5_update_d128_K64.npz")))


def combine_golden():
    """Targets with Gaussian priors (combine_lkl behind add_gaussian_prior / add_gaussian_prior_2): values of the REAL
    reference on seeded points."""
    out = {}
    for name, f in (("banana_prior", W.banana_prior), ("gmm_prior", W.gmm_prior)):
        wl = f()
        X, _ = O.orc_simulate(600, wl, seed=301)
        out[name + "_X"] = X
        out[name + "_post"] = O.ref_prior_log_pdf(X, wl)
    np.savez_compressed(os.path.join(HERE, "combine.npz"), **out)
    print("combine", {k: v.shape for k, v in out.items()})


def updates_golden():
    """update_prop (hard-assignment EM, pmc.c:1194-1310) and the band-limited update_prop_rb (pmc.c:1444) of the compiled
    reference on seeded samples: inputs are regenerated from the seed, outputs kept."""
    out = {}
    for name, f, N, upd, band in (("hard_c1", W.banana, 4000, 2, None), ("hard_c2", W.gmm5, 6000, 2, None),
                                  ("band_c1", W.banana, 4000, 1, 3), ("hardband_c2", W.gmm5, 6000, 2, 2)):
        wl = f()
        if band is not None:
            wl["band_limit"] = np.full(wl["K"], band, np.uint64)
        X, idx = O.orc_simulate(N, wl, seed=401)
        r = O.ref_iteration(X, idx, wl, do_update=upd)
        assert r["rc"] == 0, (name, r["rc"])
        for k in ("o_wght", "o_mean", "o_std", "o_detL", "retry", "perp"):
            out[f"{name}_{k}"] = np.asarray(r[k])
    np.savez_compressed(os.path.join(HERE, "updates.npz"), **out)
    print("updates", sorted(out)[:4], os.path.getsize(os.path.join(HERE, "updates.npz")))


def dumps_golden():
    """The on-disk formats written by the reference's own mix_mvdens_dump / mvdens_dump / pmc_simu_dump / out_pmc_simu
    (oracle/ref_driver.c:ref_dump_all) for a small seeded object; the input is stored next to the files."""
    import ctypes as C
    rng = np.random.default_rng(77)
    N, d, K = 40, 3, 2
    X = rng.normal(0, 2, (N, d)); idx = rng.integers(0, K, N).astype(np.uint64)
    w = rng.random(N); w /= w.sum(); lr = rng.normal(-4, 1, N); flg = (rng.random(N) > 0.2).astype(np.int16)
    wght = np.array([0.7, 0.3]); mean = rng.normal(0, 1, (K, d))
    cov = np.array([np.eye(d) * (1.5 + k) + 0.25 for k in range(K)]); df = np.array([-1, 5], np.int32)
    logSum = -3.25
    L = O.ref()
    L.ref_set_band_limit(C.c_size_t(0))
    prefix = os.path.join(HERE, "dump_ref")
    rc = L.ref_dump_all(prefix.encode(), C.c_long(N), d, K, O.P(X), O.P(idx), O.P(w), O.P(lr), O.P(flg), C.c_double(logSum), 0,
                        O.P(wght), O.P(mean), O.P(cov), O.P(df))
    assert rc == 0
    with open(prefix + ".in", "wb") as f:
        f.write(np.array([N, d, K, 0], np.int64).tobytes()); f.write(np.array([logSum]).tobytes())
        for a in (X, idx, w, lr, flg, wght, mean, cov, df):
            f.write(np.ascontiguousarray(a).tobytes())
    print("dumps", [os.path.getsize(prefix + e) for e in (".mix", ".mvd", ".psim", ".out", ".in")])


def main():
    combine_golden()
    updates_golden()
    dumps_golden()
    assert O.ref() is not None, "the compiled reference is needed"
    for name, (f, N, seed) in CASES.items():
        wl = f()
        X, idx = O.ref_simulate(N, wl, seed=seed)
        upd = 0 if wl["importance_only"] else 1
        r = O.ref_iteration(X, idx, wl, do_update=upd)
        out = dict(X=X, indices=idx, rc=r["rc"], nok=r["nok"], maxW=r["maxW"], maxR=r["maxR"], logSum=r["logSum"],
                   perp=r["perp"], ess=r["ess"], ln_evidence=r["ln_evidence"], retry=r["retry"], log_rho=r["log_rho"],
                   logw=r["logw"], w_norm=r["w_norm"], flg=r["flg"], do_update=upd)
        if upd:
            out.update(o_wght=r["o_wght"], o_mean=r["o_mean"], o_std=r["o_std"], o_detL=r["o_detL"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "rc", r["rc"], "perp", r["perp"], "ess", r["ess"], os.path.getsize(os.path.join(HERE, name + ".npz")))
    update_digest()
    # ranindx / isinBox / Cholesky known answers
    rng = np.random.default_rng(7)
    wl = W.banana()
    cw = np.cumsum(wl["wght"])
    z = np.concatenate([rng.random(2000), cw, np.nextafter(cw, 0), np.nextafter(cw, 2)])
    idx = np.zeros(len(z), np.uint64)
    O.ref().ref_ranindx(O.C.c_long(len(z)), wl["K"], O.P(cw), O.P(z), O.P(idx))
    xb = rng.normal(0, 25, (2000, wl["d"]))
    xb[:40, 0] = np.repeat([40.0, np.nextafter(40.0, 50), np.nextafter(40.0, 0), -40.0], 10)
    ins = np.zeros(len(xb), np.int32)
    O.ref().ref_isinbox(O.C.c_long(len(xb)), wl["d"], len(wl["faces"]), O.P(wl["faces"]), O.P(xb), O.P(ins))
    a = rng.normal(0, 1, (12, 12))
    spd = a @ a.T + 0.5 * np.eye(12)
    L = spd.copy()
    det = O.C.c_double()
    rc = O.ref().ref_cholesky(12, O.P(L), O.C.byref(det))
    bad = spd.copy(); bad[3, 3] = -1.0
    Lb = bad.copy()
    rcb = O.ref().ref_cholesky(12, O.P(Lb), O.C.byref(O.C.c_double()))
    np.savez_compressed(os.path.join(HERE, "misc.npz"), cw=cw, z=z, ranindx=idx, box_x=xb, box_in=ins, faces=wl["faces"],
                        spd=spd, chol=L, chol_det=det.value, chol_rc=rc, bad=bad, bad_rc=rcb, bad_after=Lb)
    print("misc", rc, rcb)


if __name__ == "__main__":
    main()
