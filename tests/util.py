"""Shared helpers for the parity tests."""
import numpy as np

RTOL = 1e-12  # BASELINE.json north_star: "within 1e-12 relative in double precision"
PTOL = 1e-12  # perplexity, ESS: the same bar (measured errors are printed by tools/ab_step.py: 1e-15 .. 4e-14)
# Normalised weights w_i = exp(lw_i - maxW + 500) / S.  An ABSOLUTE error delta in a log-weight is a RELATIVE error delta
# in the weight, and lw_i = t*post - log_rho is a difference of log-densities of size up to `big` (a few hundred at
# d = 128), each good to a few eps relative: delta <= ~8 eps big = 1e-12 at big = 560.  So the bar is 1e-12 relative for
# every weight above 1e-3 of the largest (and 1e-12 of that floor below it), on all the shapes the tests run.
WTOL = 1e-12


def relerr(a, b, scale=0.0):
    """max |a-b| / max(|a|, |b|, scale) elementwise; NaNs must coincide."""
    a = np.atleast_1d(np.asarray(a, float))
    b = np.atleast_1d(np.asarray(b, float))
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN pattern differs"
    a, b = a[~na], b[~nb]
    if a.size == 0:
        return 0.0
    inf = np.isinf(a) | np.isinf(b)
    assert np.array_equal(a[inf], b[inf]), "inf pattern differs"
    a, b = a[~inf], b[~inf]
    if a.size == 0:
        return 0.0
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    den[den == 0] = 1.0
    return float(np.max(np.abs(a - b) / den))


def assert_close(a, b, rtol=RTOL, scale=0.0, what=""):
    e = relerr(a, b, scale)
    assert e <= rtol, f"{what}: relative error {e:.3e} > {rtol:.1e}"
    return e


def cov_from_L(L):
    """Sigma = tril(L) tril(L)^T for a stack of factors stored the reference way (upper triangle mirrored)."""
    T = np.tril(L)
    return T @ np.swapaxes(T, -1, -2)


def compare_mixture(got_w, got_m, got_L, ref_w, ref_m, ref_L, rtol=RTOL, what=""):
    """Updated proposal: weights and means elementwise (scaled by the component's spread), covariances relative to the
    largest entry of each component's matrix; dead components must be dead in both (bit-exact pattern)."""
    assert np.array_equal(got_w > 0, ref_w > 0), f"{what}: live/dead pattern differs"
    errs = {}
    errs["wght"] = assert_close(got_w, ref_w, rtol, scale=np.max(ref_w), what=what + " wght")
    Cg, Cr = cov_from_L(got_L), cov_from_L(ref_L)
    em = ec = el = 0.0
    for k in range(len(ref_w)):
        if ref_w[k] <= 0:
            assert not got_m[k].any() and not got_L[k].any(), f"{what}: dead component {k} not zeroed"
            continue
        sig = np.sqrt(np.max(np.diag(Cr[k])))
        em = max(em, relerr(got_m[k], ref_m[k], scale=max(sig, np.max(np.abs(ref_m[k])))))
        ec = max(ec, relerr(Cg[k], Cr[k], scale=np.max(np.abs(Cr[k]))))
        el = max(el, relerr(np.tril(got_L[k]), np.tril(ref_L[k]), scale=np.max(np.abs(ref_L[k]))))
    errs.update(mean=em, cov=ec, L=el)
    assert em <= rtol, f"{what}: mean error {em:.3e}"
    assert ec <= rtol, f"{what}: covariance error {ec:.3e}"
    assert el <= 10 * rtol, f"{what}: Cholesky factor error {el:.3e}"
    return errs
