"""ctypes access to the two CPU oracles (TEST INFRASTRUCTURE): oracle/liboracle.so (restatement) and
oracle/_ref/libpmcref.so (the unmodified reference compiled against the GSL shim)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpmcref.so")

vp = C.c_void_p


def P(a):
    return None if a is None else a.ctypes.data_as(vp)


def build_oracles():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


class OrcMix(C.Structure):
    _fields_ = [("K", C.c_int), ("d", C.c_int), ("wght", vp), ("mean", vp), ("L", vp), ("detL", vp), ("df", vp),
                ("band_limit", vp)]


class OrcTarget(C.Structure):
    _fields_ = [("kind", C.c_int), ("d", C.c_int), ("b", C.c_double), ("sig1", C.c_double), ("mix", OrcMix),
                ("nterms", C.c_int), ("terms", vp), ("dims", vp)]


_orc = None
_ref = None


def orc():
    global _orc
    if _orc is None:
        if not os.path.exists(ORC_SO):
            build_oracles()
        L = C.CDLL(ORC_SO)
        L.orc_scalar_product.restype = C.c_double
        L.orc_mvdens_log_pdf.restype = C.c_double
        L.orc_mix_log_pdf.restype = C.c_double
        L.orc_bentgauss.restype = C.c_double
        L.orc_target_log_pdf.restype = C.c_double
        L.orc_ranindx.restype = C.c_size_t
        L.orc_ranindx.argtypes = [vp, C.c_size_t, C.c_double]
        L.orc_weights.restype = C.c_long
        L.orc_simulate.restype = C.c_long
        L.orc_filter_histogram.restype = C.c_long
        L.orc_mean_from_psim.restype = C.c_double
        _orc = L
    return _orc


def ref():
    """The compiled reference, or None where it has not been built (no /root/reference and no prebuilt .so)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference/pmclib/src"):
                build_oracles()
            else:
                return None
        _ref = C.CDLL(REF_SO)
    return _ref


class Mix:
    """Mixture on the oracle side: keeps numpy arrays alive and exposes the ctypes struct."""

    def __init__(self, wght, mean, cov, df=None, factorise=True, band_limit=None):
        self.K, self.d = mean.shape
        self.wght = np.ascontiguousarray(wght, np.float64).copy()
        self.mean = np.ascontiguousarray(mean, np.float64).copy()
        self.L = np.ascontiguousarray(cov, np.float64).copy()
        self.detL = np.zeros(self.K)
        self.df = np.full(self.K, -1, np.int32) if df is None else np.ascontiguousarray(df, np.int32).copy()
        if factorise:
            for k in range(self.K):
                det = C.c_double()
                rc = orc().orc_cholesky(self.d, P(self.L[k]), C.byref(det))
                if rc:
                    raise ValueError(f"component {k} not positive definite")
                self.detL[k] = det.value
        self.band = None if band_limit is None else np.ascontiguousarray(band_limit, np.uint64)
        self.c = OrcMix(self.K, self.d, P(self.wght), P(self.mean), P(self.L), P(self.detL), P(self.df), P(self.band))


def make_target(t):
    if t["kind"] == 1:
        return OrcTarget(1, t["d"], t["b"], t["sig1"], OrcMix()), None
    if t["kind"] == 5:  # combine_lkl / add_gaussian_prior: terms on index-mapped coordinates
        n = len(t["terms"])
        terms = (OrcTarget * n)()
        dims = (vp * n)()
        keep = []
        for i, tt in enumerate(t["terms"]):
            sub, k = make_target(tt)
            terms[i] = sub
            dm = np.ascontiguousarray(tt["dims"], np.int32)
            dims[i] = dm.ctypes.data
            keep += [sub, k, dm]
        return OrcTarget(5, t["d"], 0.0, 0.0, OrcMix(), n, C.cast(terms, vp), C.cast(dims, vp)), (terms, dims, keep)
    m = Mix(t["wght"], t["mean"], t["cov"], t["df"])
    return OrcTarget(t["kind"], t["d"], 0.0, 0.0, m.c), m


def orc_iteration(X, indices, wl, t=1.0, do_update=1, in_retry=0, filt=None):
    """Restatement of one iteration on given samples; mirrors ref_iteration's outputs."""
    L = orc()
    X = np.ascontiguousarray(X, np.float64)
    N, d = X.shape
    if wl.get("is_chol"):  # wl["cov"] already holds Cholesky factors (e.g. the proposal a previous update produced)
        m = Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"], factorise=False, band_limit=wl.get("band_limit"))
        for k in range(m.K):
            det = 1.0
            for a in range(d):
                det *= m.L[k, a, a]
            m.detL[k] = det
    else:
        m = Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"], band_limit=wl.get("band_limit"))
    tgt, keep = make_target(wl["target"])
    w = np.zeros(N); lr = np.zeros(N); flg = np.zeros(N, np.int16)
    mw, mr = C.c_double(), C.c_double()
    nok = L.orc_weights(C.c_long(N), d, P(X), C.byref(m.c), C.byref(tgt), C.c_double(t), P(w), P(lr), P(flg),
                        C.byref(mw), C.byref(mr))
    out = dict(rc=0, nok=nok, maxW=mw.value, maxR=mr.value, log_rho=lr.copy(), logw=w.copy())
    if nok < 0:
        out["rc"] = nok
        return out
    if filt is not None:  # the pmc_filter hook between weights and normalisation (pmc.c:203-206)
        L.orc_filter_histogram.restype = C.c_long
        out["n_filtered"] = L.orc_filter_histogram(C.c_long(N), P(w), P(flg), *[int(f) for f in filt])
    ls = C.c_double()
    rc = L.orc_normalize(C.c_long(N), P(w), P(flg), mw, C.byref(ls))
    out.update(rc=rc, logSum=ls.value, w_norm=w.copy(), flg=flg.copy())
    if rc:
        return out
    perp, ess, lne = C.c_double(), C.c_double(), C.c_double()
    rc = L.orc_perplexity_ess(C.c_long(N), P(w), P(flg), C.byref(perp), C.byref(ess))
    rc = rc or L.orc_evidence(C.c_long(N), P(flg), ls, C.byref(lne))
    out.update(rc=rc, perp=perp.value, ess=ess.value, ln_evidence=lne.value)
    if rc or not do_update:
        return out
    idx = np.ascontiguousarray(indices, np.uint64)
    retry = C.c_int(in_retry)
    if do_update == 1:
        rc = L.orc_update_rb(C.c_long(N), P(X), P(idx), P(w), P(lr), P(flg), mr, C.byref(m.c), C.byref(retry))
    else:
        rc = L.orc_update(C.c_long(N), P(X), P(idx), P(w), P(flg), C.byref(m.c), C.byref(retry))
    out.update(rc=rc, retry=retry.value, o_wght=m.wght.copy(), o_mean=m.mean.copy(), o_std=m.L.copy(),
               o_detL=m.detL.copy())
    return out


def ref_iteration(X, indices, wl, t=1.0, do_update=1, in_retry=0):
    L = ref()
    bl = wl.get("band_limit")
    L.ref_set_band_limit(C.c_size_t(0 if bl is None else int(bl[0])))  # the driver applies one limit to the whole mixture
    X = np.ascontiguousarray(X, np.float64)
    N, d = X.shape
    K = wl["K"]
    tg = wl["target"]
    tp = np.array([tg["b"], tg["sig1"]])
    idx = None if indices is None else np.ascontiguousarray(indices, np.uint64)
    lr = np.zeros(N); lw = np.zeros(N); wn = np.zeros(N); flg = np.zeros(N, np.int16); scal = np.zeros(8)
    ow = np.zeros(K); om = np.zeros((K, d)); os_ = np.zeros((K, d, d)); och = np.zeros(K, np.int32); od = np.zeros(K)
    rc = L.ref_iteration(C.c_long(N), d, K, P(X), P(idx), P(np.ascontiguousarray(wl["wght"])),
                         P(np.ascontiguousarray(wl["mean"])), P(np.ascontiguousarray(wl["cov"])),
                         P(np.ascontiguousarray(wl["df"], np.int32)),
                         tg["kind"], P(tp), tg["K"], P(tg["wght"]), P(tg["mean"]), P(tg["cov"]), P(tg["df"]),
                         C.c_double(t), do_update, in_retry, P(lr), P(lw), P(wn), P(flg), P(scal),
                         P(ow), P(om), P(os_), P(och), P(od))
    return dict(rc=rc, nok=int(scal[0]), maxW=scal[1], maxR=scal[2], logSum=scal[3], perp=scal[4], ess=scal[5],
                ln_evidence=scal[6], retry=int(scal[7]), log_rho=lr, logw=lw, w_norm=wn, flg=flg,
                o_wght=ow, o_mean=om, o_std=os_, o_chol=och, o_detL=od)


def ref_prior_log_pdf(X, wl):
    """The REAL reference: make_target + add_gaussian_prior[_2] + distribution_lkl (oracle/ref_driver.c:ref_prior_log_pdf)."""
    L = ref()
    t = wl["target"]
    base, pr = t["terms"][0], t["prior"]
    X = np.ascontiguousarray(X, np.float64)
    N, d = X.shape
    tp = np.array([base["b"], base["sig1"]])
    out = np.zeros(N)
    full = 1 if np.ndim(pr["var"]) == 2 else 0
    idim = np.ascontiguousarray(pr["idim"], np.int32)
    loc, var = np.ascontiguousarray(pr["loc"], float), np.ascontiguousarray(pr["var"], float)
    if base["kind"] == 1:
        rc = L.ref_prior_log_pdf(C.c_long(N), d, P(X), 1, P(tp), 0, None, None, None, None, len(idim), P(idim), P(loc), P(var), full, P(out))
    else:
        tw, tm, tc = (np.ascontiguousarray(base[k], float) for k in ("wght", "mean", "cov"))
        tdf = np.ascontiguousarray(base["df"], np.int32)
        rc = L.ref_prior_log_pdf(C.c_long(N), d, P(X), base["kind"], P(tp), len(tw), P(tw), P(tm), P(tc), P(tdf), len(idim), P(idim),
                                 P(loc), P(var), full, P(out))
    assert rc == 0, rc
    return out


def orc_target_log_pdf(X, wl):
    """The restatement's target (any kind, including combined) on the rows of X."""
    L = orc()
    tgt, keep = make_target(wl["target"])
    X = np.ascontiguousarray(X, np.float64)
    N, d = X.shape
    tmp = np.zeros(d); ld = np.zeros(64); rc = C.c_int(0)
    out = np.array([L.orc_target_log_pdf(C.byref(tgt), P(X[i]), P(tmp), P(ld), C.byref(rc)) for i in range(N)])
    assert rc.value == 0
    return out


def ref_simulate(N, wl, seed=1):
    L = ref()
    d, K = wl["d"], wl["K"]
    X = np.zeros((N, d)); idx = np.zeros(N, np.uint64); wa = np.zeros(K)
    faces = wl["faces"]
    nf = 0 if faces is None else len(faces)
    rc = L.ref_simulate(C.c_long(N), d, K, P(np.ascontiguousarray(wl["wght"])), P(np.ascontiguousarray(wl["mean"])),
                        P(np.ascontiguousarray(wl["cov"])), P(np.ascontiguousarray(wl["df"], np.int32)),
                        nf, P(faces), C.c_ulong(seed), P(X), P(idx), P(wa))
    if rc:
        raise RuntimeError(f"ref_simulate rc={rc}")
    return X, idx


def orc_simulate(N, wl, seed=1):
    L = orc()
    m = Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"])
    d = wl["d"]
    X = np.zeros((N, d)); idx = np.zeros(N, np.uint64); flg = np.zeros(N, np.int16); cw = np.zeros(wl["K"])
    faces = wl["faces"]
    nf = 0 if faces is None else len(faces)
    n = L.orc_simulate(C.c_long(N), C.byref(m.c), nf, P(faces), C.c_ulonglong(seed), P(X), P(idx), P(flg), P(cw))
    if n != N:
        raise RuntimeError(f"orc_simulate rc={n}")
    return X, idx
