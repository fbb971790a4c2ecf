#!/usr/bin/env python
"""Wall-clock time of the separate phases (draw / weights / normalise / update) for the large-d config shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmclib_b200 import PmcGpu, workloads as W
cases = [("c4 student d30 K50", W.student30(), int(os.environ.get("N4", 2_000_000)), 0),
         ("c5 gauss d128 K64", W.gauss128(), int(os.environ.get("N5", 200_000)), 1)]
g = PmcGpu(0)
for o in sys.argv[1:]:
    k, v = o.split("="); g.set_option(k, float(v))
def tm(f, n=2):
    f()
    t0 = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t0) / n * 1e3, r
for name, wl, N, upd in cases:
    g.set_workload(wl); g.resize(N, wl["d"])
    td, _ = tm(lambda: g.draw(1, 0))
    tw, r = tm(lambda: g.weights())
    nok, maxW, maxR = r
    tn, _ = tm(lambda: (g.weights(), g.normalize(maxW)))
    tn -= tw
    line = f"{name:24s} N={N:9d} draw {td:8.2f} ms  weights {tw:8.2f} ms  normalise {tn:6.2f} ms"
    if upd:
        def up():
            g.set_workload(wl); g.weights(); g.normalize(maxW); return g.update_rb(maxR)
        tu, _ = tm(up, 1)
        line += f"  update(two-pass) {tu - tw - tn:8.2f} ms"
        def st():
            g.set_workload(wl); return g.step_given(do_update=1)
        g.set_option("timing", 1)
        ts, rr = tm(st, 1)
        ms, nl, _ = g.get_timing(); sms = g.get_stats_timing()
        g.set_option("timing", 0)
        tsw, _ = tm(lambda: g.set_workload(wl), 1)
        line += f"  step_given {ts - tsw:8.2f} ms (+set_workload {tsw:.1f}) precise={rr.used_precise_path}  [whitening {ms / max(nl, 1):.2f} ms, resp+stats {sms / 2:.2f} ms]"
    print(line)
