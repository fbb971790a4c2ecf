"""Resampling of the weighted sample on the device (pmc.c:1585-1626).  The reference consumes a sequential GSL stream
(alias method / conditional binomials); the device version is validated statistically, like the draw."""
import numpy as np
import pytest

from pmclib_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _load(gpu, w, d=2):
    N = len(w)
    gpu.set_workload(W.shipped_example())
    gpu.resize(N, d)
    gpu.upload(X=np.zeros((N, d)), weights=w, flg=np.ones(N, np.int16))


def test_sample_indices_follow_the_weights(gpu):
    rng = np.random.default_rng(5)
    N = 3000
    w = rng.gamma(0.7, 1.0, N)
    w[::7] = 0.0                      # zero-weight entries are never drawn
    w /= w.sum()
    _load(gpu, w)
    counts = np.zeros(N)
    reps = 300
    for r in range(reps):
        idx = gpu.sample_indices(seed=1234, stream=r)
        assert idx.shape == (N,) and idx.max() < N
        counts += np.bincount(idx.astype(np.int64), minlength=N)
    assert np.all(counts[w == 0] == 0)
    exp = reps * N * w
    big = exp > 20
    chi2 = np.sum((counts[big] - exp[big]) ** 2 / exp[big])
    dof = big.sum()
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)
    # same (seed, stream) -> same draw; another stream -> another draw
    a, b, c = gpu.sample_indices(9, 1), gpu.sample_indices(9, 1), gpu.sample_indices(9, 2)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_resample_residual_multiplicities(gpu):
    rng = np.random.default_rng(6)
    N = 4000
    w = rng.gamma(0.5, 1.0, N)
    w /= w.sum()
    _load(gpu, w)
    res = N * w - np.floor(N * w)     # pmc.c:1619
    tot = np.zeros(N)
    reps = 300
    for r in range(reps):
        n = gpu.resample_residual(seed=77, stream=r)
        assert n.dtype == np.uint32 and int(n.sum()) == N   # Ntrials = nsamples (pmc.c:1611)
        tot += n
    exp = reps * N * res / res.sum()
    big = exp > 20
    chi2 = np.sum((tot[big] - exp[big]) ** 2 / exp[big])
    dof = big.sum()
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)
    # weights that are exact multiples of 1/N have no residual mass
    _load(gpu, np.full(N, 1.0 / N))
    assert not gpu.resample_residual(1, 0).any()


def test_prefix_sum_at_scale(gpu):
    """2e7 weights: the cumulative table spans many scan blocks; drawn indices stay in range and hit heavy entries."""
    N = 20_000_000
    w = np.full(N, 0.5 / N)
    w[12_345_678] = 0.5
    _load(gpu, w)
    idx = gpu.sample_indices(3, 0)
    frac = np.mean(idx == 12_345_678)
    assert abs(frac - 0.5) < 0.001
    assert idx.max() < N
