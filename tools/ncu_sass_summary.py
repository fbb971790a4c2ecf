#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: per-opcode instruction counts, shared-memory wavefronts (actual vs ideal),
stall samples, and the hottest instructions.  Usage: ncu_sass_summary.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, name):
    try: return float(r[col[name]])
    except: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in body)
tot_samp = sum(f(r, "# Samples") for r in body)
tot_wf = sum(f(r, "L1 Wavefronts Shared") for r in body)
tot_ideal = sum(f(r, "L1 Wavefronts Shared Ideal") for r in body)
print(f"instructions executed {tot_inst:.4g}  samples {tot_samp:.0f}  smem wavefronts {tot_wf:.4g} (ideal {tot_ideal:.4g})")
by = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in body:
    src = r[col["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    b = by[op]
    b[0] += f(r, "Instructions Executed"); b[1] += f(r, "# Samples"); b[2] += f(r, "L1 Wavefronts Shared"); b[3] += f(r, "L1 Wavefronts Shared Ideal"); b[4] += 1
print(f"{'opcode':28s} {'static':>6s} {'inst%':>7s} {'samp%':>7s} {'smem_wf':>11s} {'ideal':>11s}")
for op, b in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{op:28s} {b[4]:6d} {100*b[0]/tot_inst:7.2f} {100*b[1]/max(tot_samp,1):7.2f} {b[2]:11.4g} {b[3]:11.4g}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in body) for s in stalls}
print("stall samples:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
print("hottest instructions by samples:")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
    st = {s[6:]: int(f(r, s)) for s in stalls if f(r, s) > 0}
    print(f"  {r[col['Address']][-5:]} {r[col['Source']].strip()[:70]:70s} samp={int(f(r,'# Samples')):5d} exec={f(r,'Instructions Executed'):.3g} wf={f(r,'L1 Wavefronts Shared'):.3g}/{f(r,'L1 Wavefronts Shared Ideal'):.3g} {st}")
