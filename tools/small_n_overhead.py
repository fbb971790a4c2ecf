#!/usr/bin/env python
"""Fixed cost of one iteration: wall-clock per call of the phases at small N (configs[0] shape), where the kernels take a few
microseconds each and what is left is launches, small copies, the synchronisation and the host finalisation.

    python tools/small_n_overhead.py [N ...]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pmclib_b200 import PmcGpu, workloads as W

Ns = [int(a) for a in sys.argv[1:]] or [10_000, 100_000, 1_000_000]
for wl in (W.banana(), W.gmm5()):
    g = PmcGpu(0)
    g.set_workload(wl)
    for N in Ns:
        g.resize(N, wl["d"])
        def tm(f, n=300):
            for _ in range(20): f()
            t0 = time.perf_counter()
            for _ in range(n): r = f()
            return (time.perf_counter() - t0) / n * 1e6
        it = [0]
        prop = g.get_proposal()
        def step(upd):  # every step starts from the same proposal (its upload is timed separately below)
            it[0] += 1
            if upd: g.set_proposal(prop["wght"], prop["mean"], prop["L"], prop["detL"], None)
            return g.pmc_step(1, it[0], 0, N, upd, 0)
        t_step = tm(lambda: step(1))
        g.set_workload(wl)
        t_imp = tm(lambda: step(0))
        t_draw = tm(lambda: g.draw(1, 0))
        t_w = tm(lambda: g.weights())
        t_set = tm(lambda: g.set_proposal(prop["wght"], prop["mean"], prop["L"], prop["detL"], None))
        print(f"{wl['name']:18s} N={N:8d}  step(update) {t_step:7.1f} us  step(no update) {t_imp:7.1f}  draw {t_draw:7.1f}  weights {t_w:7.1f}  set_proposal {t_set:6.1f}", flush=True)
    g.close()
