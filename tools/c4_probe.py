"""configs[3] shape: wall-clock of draw and weights separately over 45 iterations (shows the steady state and the power-cap
episodes of the 100 %-duty DMMA whitening)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmclib_b200 import PmcGpu, workloads as W
g = PmcGpu(0); wl = W.student30(); g.set_workload(wl); g.resize(10_000_000, wl["d"])
td=[]; tw=[]
for it in range(45):
    t0=time.perf_counter(); g.draw(20261020, it); t1=time.perf_counter(); g.weights(); t2=time.perf_counter()
    td.append((t1-t0)*1e3); tw.append((t2-t1)*1e3)
print("draw   ", " ".join(f"{x:.1f}" for x in td))
print("weights", " ".join(f"{x:.1f}" for x in tw))
