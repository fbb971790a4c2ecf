#!/usr/bin/env python
"""Mid-size dimensions without a fused instantiation (e.g. d = 7, 12, 16): literal kernels vs the tensor path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmclib_b200 import PmcGpu, workloads as W
N = int(os.environ.get("N", 4_000_000))
for d, K in [(7, 3), (12, 10), (16, 8)]:
    wl = W.banana(d=d, K=K, seed=9, mean_sigma=2.0, var=6.0, box=40.0)
    for mind in (17, 2):
        g = PmcGpu(0)
        g.set_option("mma_min_d", mind)
        g.set_workload(wl); g.resize(N, d)
        r = g.pmc_step(1, 0, do_update=0)
        t0 = time.perf_counter(); n = 3
        for it in range(n):
            g.set_workload(wl)
            r = g.pmc_step(1, 10 + it, do_update=1, retry=0)
        dt = (time.perf_counter() - t0) / n
        print(f"d={d:3d} K={K:3d} mma_min_d={mind:2d}: {dt*1e3:8.2f} ms/iter  {N/dt:.3e} samples/s  perp={r.perplexity:.4f} precise={r.used_precise_path}")
        g.close()
