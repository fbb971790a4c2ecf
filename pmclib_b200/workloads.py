"""Synthetic PMC workloads of the shapes BASELINE.json names, numpy only.

A workload is a plain dict: proposal mixture (wght[K], mean[K,d], cov[K,d,d], df[K]), target spec and box faces.
Used by bench.py, the tests and tests/golden/make_golden.py so that all three see identical inputs.
"""
from __future__ import annotations

import numpy as np

TGT_BANANA, TGT_MVDENS, TGT_MIX, TGT_COMBINE = 1, 2, 3, 5


def box_faces(d: int, lo: float, hi: float) -> np.ndarray:
    """Axis-aligned slabs in parabox layout (parabox.c:75-87): per coordinate a (-e_j, -lo) and a (+e_j, hi) face."""
    faces = np.zeros((2 * d, d + 1))
    for j in range(d):
        faces[2 * j, j] = -1.0
        faces[2 * j, d] = -lo
        faces[2 * j + 1, j] = 1.0
        faces[2 * j + 1, d] = hi
    return faces


def _target(kind, d, **kw):
    t = dict(kind=kind, d=d, b=0.0, sig1=1.0, K=0, wght=None, mean=None, cov=None, df=None)
    t.update(kw)
    return t


def banana(d=10, K=10, seed=1234, mean_sigma=3.0, var=10.0, box=40.0, b=0.1, sig1=10.0):
    """configs[0] / configs[2]: d-D bentGauss (examples/pmc/test_pmc.lua:4-5 values), K Gaussian components."""
    rng = np.random.default_rng(seed)
    return dict(
        name=f"banana_d{d}_K{K}", d=d, K=K,
        wght=np.full(K, 1.0 / K), mean=rng.normal(0.0, mean_sigma, (K, d)),
        cov=np.tile(var * np.eye(d), (K, 1, 1)), df=np.full(K, -1, np.int32),
        target=_target(TGT_BANANA, d, b=b, sig1=sig1), faces=box_faces(d, -box, box), importance_only=False)


def gmm5(d=5, K=20, seed=4321):
    """configs[1]: 5-D 4-mode Gaussian-mixture target, 20-component Gaussian proposal."""
    rng = np.random.default_rng(seed)
    tm = np.zeros((4, d))
    tm[0, 0], tm[1, 0], tm[2, 1], tm[3, 1] = 3.0, -3.0, 3.0, -3.0
    return dict(
        name=f"gmm_d{d}_K{K}", d=d, K=K,
        wght=np.full(K, 1.0 / K), mean=rng.normal(0.0, 4.0, (K, d)),
        cov=np.tile(4.0 * np.eye(d), (K, 1, 1)), df=np.full(K, -1, np.int32),
        target=_target(TGT_MIX, d, K=4, wght=np.full(4, 0.25), mean=tm, cov=np.tile(np.eye(d), (4, 1, 1)),
                       df=np.full(4, -1, np.int32)),
        faces=box_faces(d, -30.0, 30.0), importance_only=False)


def student30(d=30, K=50, seed=99, nu=9):
    """configs[3]: fixed 50-component Student-t proposal (nu=9), 30-D standard-normal target, importance only."""
    rng = np.random.default_rng(seed)
    return dict(
        name=f"student_d{d}_K{K}", d=d, K=K,
        wght=np.full(K, 1.0 / K), mean=rng.normal(0.0, 0.5, (K, d)),
        cov=np.tile(1.5 * np.eye(d), (K, 1, 1)), df=np.full(K, nu, np.int32),
        target=_target(TGT_MVDENS, d, K=1, wght=np.ones(1), mean=np.zeros((1, d)), cov=np.eye(d)[None],
                       df=np.full(1, -1, np.int32)),
        faces=None, importance_only=True)


def gauss128(d=128, K=64, seed=7):
    """configs[4]: 128-D Gaussian target, 64-component proposal with Sigma_k = 1.2 I + 0.1 (low rank)."""
    rng = np.random.default_rng(seed)
    cov = np.empty((K, d, d))
    for k in range(K):
        u = rng.normal(0.0, 1.0, (d, 4)) / np.sqrt(d)
        cov[k] = 1.2 * np.eye(d) + 0.1 * (u @ u.T) * d / 4
    return dict(
        name=f"gauss_d{d}_K{K}", d=d, K=K,
        wght=np.full(K, 1.0 / K), mean=rng.normal(0.0, 0.3, (K, d)), cov=cov, df=np.full(K, -1, np.int32),
        target=_target(TGT_MVDENS, d, K=1, wght=np.ones(1), mean=np.zeros((1, d)), cov=np.eye(d)[None],
                       df=np.full(1, -1, np.int32)),
        faces=None, importance_only=False)


def shipped_example(seed=5):
    """examples/pmc/test_pmc.lua verbatim shape: 4-D bentGauss with dims 2,3 pinned -> 2 free dims, 9 components,
    box +-20; the pinned dims only add a constant to the target, so the free 2-D banana is used directly."""
    w = banana(d=2, K=9, seed=seed, mean_sigma=np.sqrt(10.0), var=10.0, box=20.0)
    w["name"] = "shipped_test_pmc"
    return w


def with_gaussian_prior(wl, idim, loc, var):
    """The workload's target with a Gaussian prior on the coordinates idim, as add_gaussian_prior (var 1-D: diagonal) /
    add_gaussian_prior_2 (var 2-D: full covariance) build it (distribution.c:548-606): combine_lkl of [target, mvdens]."""
    idim = np.asarray(idim, np.int32)
    var = np.asarray(var, float)
    cov = np.diag(var) if var.ndim == 1 else var
    n = len(idim)
    base = dict(wl["target"], dims=np.arange(wl["d"], dtype=np.int32))
    prior = _target(TGT_MVDENS, n, K=1, wght=np.ones(1), mean=np.asarray(loc, float)[None], cov=cov[None],
                    df=np.full(1, -1, np.int32), dims=idim)
    out = dict(wl)
    out["name"] = wl["name"] + "_prior"
    out["target"] = dict(kind=TGT_COMBINE, d=wl["d"], terms=[base, prior], prior=dict(idim=idim, loc=np.asarray(loc, float), var=var))
    return out


def banana_prior(**kw):
    """10-D banana with a Gaussian prior on coordinates 0, 3, 7 (the shape add_gaussian_prior produces for pmc_rc targets)."""
    return with_gaussian_prior(banana(**kw), [0, 3, 7], [0.5, -1.0, 2.0], [25.0, 4.0, 9.0])


def gmm_prior(**kw):
    """configs[1] target with a full-covariance prior on coordinates 1 and 4 (add_gaussian_prior_2)."""
    return with_gaussian_prior(gmm5(**kw), [1, 4], [0.3, -0.2], [[9.0, 1.5], [1.5, 4.0]])


CONFIGS = {"c1": banana, "c2": gmm5, "c3": banana, "c4": student30, "c5": gauss128}
