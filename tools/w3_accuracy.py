#!/usr/bin/env python
"""Measured errors of the third-generation weights kernel against the CPU oracle, exact-subtraction solve (w3_variant 0)
beside the row-scaled solve (w3_variant 3), on workloads chosen to stress the row-scaled form: components far from the
origin relative to their width (|mu_j / L_jj| large before the common shift), narrow and wide components mixed, non-uniform
weights, correlated covariances.  Prints max errors of log_rho, log-weights, normalised weights and perplexity.

    python tools/w3_accuracy.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_py as O
from pmclib_b200 import PmcGpu, workloads as W
from util import relerr


def offset(wl, off, scale=1.0, seed=3):
    """the same mixture moved to `off` in every coordinate (target moved along), covariances scaled and correlated"""
    rng = np.random.default_rng(seed)
    wl = dict(wl)
    K, d = wl["K"], wl["d"]
    wl["mean"] = wl["mean"] * scale + off
    A = rng.normal(size=(K, d, d)) * 0.3
    cov = np.einsum("kij,klj->kil", A, A) + np.eye(d)
    wl["cov"] = cov * (scale ** 2) * rng.uniform(0.3, 3.0, (K, 1, 1))
    w = rng.uniform(0.2, 1.0, K)
    wl["wght"] = w / w.sum()
    wl["faces"] = W.box_faces(d, off - 60.0 * scale - 40.0, off + 60.0 * scale + 40.0)
    return wl


CASES = [
    ("banana d10 K10", W.banana(), 40000),
    ("banana d10 K10 correlated, weights non-uniform", offset(W.banana(), 0.0), 40000),
    ("banana d10 K10 narrow (scale 0.02) at offset 30", offset(W.banana(), 30.0, 0.02), 40000),
    ("banana d10 K20", W.banana(K=20, seed=10, mean_sigma=2.0, var=6.0), 30000),
    ("gmm d5 K20", W.gmm5(), 40000),
    ("banana d5 K10", W.banana(d=5, K=10, seed=12, mean_sigma=2.0, var=6.0, box=30.0), 20000),
    ("shipped d2 K9", W.shipped_example(), 30000),
]

g = PmcGpu(0)
for name, wl, N in CASES:
    X, idx = O.orc_simulate(N, wl, seed=5)
    o = O.orc_iteration(X, idx, wl, do_update=0)
    big = max(float(np.max(np.abs(o["log_rho"]))), 1.0)
    for v in (0, 3):
        g.set_option("w3_variant", v)
        g.set_option("fused_version", 3)
        g.set_workload(wl); g.resize(N, wl["d"])
        g.upload(X=X, indices=idx, flg=np.ones(N, np.int16))
        nok, maxW, maxR = g.weights(1.0)
        got = g.download(X=False, indices=False)
        e_rho = relerr(got["log_rho"], o["log_rho"])
        e_lw = relerr(got["weights"], o["logw"], scale=big)
        logSum, cnt = g.normalize(maxW)
        got = g.download(X=False, indices=False, log_rho=False)
        e_w = relerr(got["weights"], o["w_norm"], scale=float(np.max(o["w_norm"])) * 1e-3)
        perp, ess, nfl = g.perplexity_ess()
        e_p = abs(perp - o["perp"]) / o["perp"]
        path = f"rowscaled={int(g.get_info('w3_rowscaled'))} ratio={g.get_info('w3_shift_ratio'):.1f}"
        print(f"{name:52s} variant {v}: log_rho {e_rho:.2e}  log-w {e_lw:.2e}  w_norm {e_w:.2e}  perplexity {e_p:.2e}  "
              f"flags {'same' if np.array_equal(got['flg'], o['flg']) else 'DIFFER'}  nok {nok}/{o['nok']}  path {path}", flush=True)
g.close()
