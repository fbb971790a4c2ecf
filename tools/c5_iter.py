"""A few full iterations at the configs[4] shape (N5 samples, ITERS iterations): the command profiled by ncu for
profiles/r01_mma_ncu.txt; PMCB200_TRACE=1 prints the wall-clock of the phases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmclib_b200 import PmcGpu, workloads as W
N = int(os.environ.get("N5", 200000))
g = PmcGpu(0); wl = W.gauss128(); g.set_workload(wl); g.resize(N, wl["d"])
for it in range(int(os.environ.get("ITERS", 2))):
    r = g.pmc_step(1, it)
    print("iter", it, r.perplexity, r.n_alive, r.used_precise_path, file=sys.stderr)
