#!/usr/bin/env python
"""Wall-clock throughput of one iteration for the other BASELINE.json config shapes (device-resident, after warm-up)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmclib_b200 import PmcGpu, workloads as W
cases = [("c2 gmm d5 K20", W.gmm5(), 20_000_000, 1), ("shipped d2 K9", W.shipped_example(), 20_000_000, 1),
         ("c4 student d30 K50 (importance only)", W.student30(), 2_000_000, 0), ("c5 gauss d128 K64", W.gauss128(), 200_000, 1)]
g = PmcGpu(0)
for name, wl, N, upd in cases:
    g.set_workload(wl); g.resize(N, wl["d"])
    r = g.pmc_step(1, 0, do_update=0)      # warm-up without touching the proposal
    t0 = time.perf_counter(); n = 3
    for it in range(n):                     # same proposal every time: the update result is computed, then discarded
        g.set_workload(wl)
        r = g.pmc_step(1, 10 + it, do_update=upd, retry=0)
    dt = (time.perf_counter() - t0) / n
    print(f"{name:40s} N={N:9d}  {dt*1e3:9.2f} ms/iter  {N/dt:.3e} samples/s  perp={r.perplexity:.4f} alive={r.n_alive} precise={r.used_precise_path}")
