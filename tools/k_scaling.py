#!/usr/bin/env python
"""Weights / statistics kernel time against the number of components (d = 10): is the cost per component constant?

    python tools/k_scaling.py [--N 50000000]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pmclib_b200 import PmcGpu, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=50_000_000)
a = ap.parse_args()
for K in (5, 10, 20):
    wl = W.banana(d=10, K=K, seed=1234, mean_sigma=3.0, var=10.0, box=40.0)
    g = PmcGpu(0)
    g.set_workload(wl); g.resize(a.N)
    retry = 0
    for it in range(2):
        r = g.pmc_step(seed=11, it=it, first_index=0, N_total=a.N, do_update=1, retry=retry); retry = r.retry
    g.set_option("timing", 1)
    for it in range(2, 6):
        r = g.pmc_step(seed=11, it=it, first_index=0, N_total=a.N, do_update=1, retry=retry); retry = r.retry
    ms, n, _ = g.get_timing()
    print(f"K={K:2d} N={a.N} weights {ms / n:7.3f} ms = {ms / n / a.N * 1e9 / K:6.3f} ps per sample-component | stats {g.get_stats_timing() / n:7.3f} ms | draw {g.get_draw_timing() / n:7.3f} ms | alive {r.n_alive}", flush=True)
    g.close()
