#!/usr/bin/env python
"""A/B of the small-d iteration on one GPU: per-kernel CUDA-event times of pmcgpu_pmc_step for a list of option sets.

    python tools/ab_step.py [--N 100000000] [--steps 5] [--wl banana|gmm5] "gen3=0" "gen3=1" ...

Each positional argument is a comma-separated list of name=value options applied to a fresh context.  Also prints the
parity of one small iteration against the CPU oracle for every option set (log_rho / weights / perplexity / update).
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pmclib_b200 import PmcGpu, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=100_000_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--wl", default="banana")
ap.add_argument("--no-parity", action="store_true")
ap.add_argument("sets", nargs="*", default=["gen3=0", "gen3=1"])
a = ap.parse_args()
wl = {"banana": W.banana, "gmm5": W.gmm5}[a.wl]()


def parity(opts):
    import oracle_py as O
    from util import relerr
    g = PmcGpu(0)
    for k, v in opts: g.set_option(k, v)
    g.set_workload(wl); g.resize(20000)
    r = g.pmc_step(seed=7, it=0, first_index=0, N_total=20000, do_update=1, retry=0)
    got = g.download()
    o = O.orc_iteration(got["X"], got["indices"], wl, do_update=1)
    big = float(np.max(np.abs(o["log_rho"])))
    e_rho = float(np.max(np.abs(got["log_rho"] - o["log_rho"]))) / big
    wmax = float(np.max(o["w_norm"]))
    m = o["w_norm"] > 0
    e_w = float(np.max(np.abs(got["weights"][m] - o["w_norm"][m]) / o["w_norm"][m]))
    e_p = abs(r.perplexity - o["perp"]) / o["perp"]
    e_mu = float(np.max(np.abs(r.mean - o["o_mean"]))) / float(np.max(np.abs(o["o_mean"])))
    e_a = float(np.max(np.abs(r.wght - o["o_wght"])))
    flg = bool(np.array_equal(got["flg"], o["flg"])) if "flg" in o else None
    g.close()
    return dict(rho=e_rho, w_rel=e_w, perp=e_p, mean=e_mu, alpha=e_a, flags=flg, precise=r.used_precise_path, wmax=wmax)


for spec in a.sets:
    opts = [(kv.split("=")[0], float(kv.split("=")[1])) for kv in spec.split(",") if kv]
    if not a.no_parity:
        try:
            print(f"[{spec}] parity vs oracle (N=20000): {parity(opts)}", flush=True)
        except Exception as e:  # keep going: the timing is still informative
            print(f"[{spec}] parity check failed: {e!r}", flush=True)
    g = PmcGpu(0)
    for k, v in opts: g.set_option(k, v)
    g.set_workload(wl); g.resize(a.N)
    retry = 0
    for it in range(2):
        r = g.pmc_step(seed=11, it=it, first_index=0, N_total=a.N, do_update=1, retry=retry); retry = r.retry
    g.set_option("timing", 1)
    t0 = time.perf_counter()
    for it in range(2, 2 + a.steps):
        r = g.pmc_step(seed=11, it=it, first_index=0, N_total=a.N, do_update=1, retry=retry); retry = r.retry
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    ms, nl, _ = g.get_timing()
    print(f"[{spec}] N={a.N} wall {wall:8.3f} ms/step -> {a.N / wall * 1e3:.4g} samples/s | draw {g.get_draw_timing() / a.steps:7.3f} "
          f"weights {ms / max(nl, 1):7.3f} stats(+norm) {g.get_stats_timing() / a.steps:7.3f} ms | perp {r.perplexity:.6f} "
          f"alive {r.n_alive} precise {r.used_precise_path}", flush=True)
    g.close()
