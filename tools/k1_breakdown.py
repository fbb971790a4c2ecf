#!/usr/bin/env python
"""Times the draw/weights kernel with and without the draw phase, and the statistics kernel (CUDA events inside the library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmclib_b200 import PmcGpu, workloads as W
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
g = PmcGpu(0)
for kv in sys.argv[2:]:
    k, v = kv.split("="); g.set_option(k, float(v))
wl = W.banana()
g.set_workload(wl); g.resize(N)
r = None
for it in range(3):
    r = g.pmc_step(1, it, retry=0 if r is None else r.retry)
prop = g.get_proposal()
def run(fn, reps=4):
    g.set_option("timing", 1)
    for _ in range(reps):
        g.set_proposal(prop["wght"], prop["mean"], prop["L"], prop["detL"]); fn()
    ms, n, _ = g.get_timing(); st = g.get_stats_timing(); g.set_option("timing", 0)
    return ms / n, st / n
print("draw+weights+stats (pmc_step)      k1 %.3f ms  k2 %.3f ms" % run(lambda: g.pmc_step(1, 7, retry=0)))
print("weights+stats on given X           k1 %.3f ms  k2 %.3f ms" % run(lambda: g.step_given(0, N, 1.0, 1, 0)))
print("draw+weights, no update            k1 %.3f ms  k2 %.3f ms" % run(lambda: g.pmc_step(1, 7, do_update=0)))
print("weights only on given X            k1 %.3f ms  k2 %.3f ms" % run(lambda: g.step_given(0, N, 1.0, 0, 0)))
