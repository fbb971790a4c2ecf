"""A few importance-only iterations at the configs[3] shape (N4 samples): the command profiled by ncu for
profiles/r01_mma_c4_ncu.txt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmclib_b200 import PmcGpu, workloads as W
N = int(os.environ.get("N4", 2_000_000))
g = PmcGpu(0); wl = W.student30(); g.set_workload(wl); g.resize(N, wl["d"])
for it in range(int(os.environ.get("ITERS", 2))):
    r = g.pmc_step(1, it, do_update=0)
    print("iter", it, r.perplexity, file=sys.stderr)
