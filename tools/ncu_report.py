#!/usr/bin/env python
"""Text summary of an ncu report (.ncu-rep) of the three kernels of the configs[2] iteration: per-kernel time, DRAM bytes,
pipe activity, shared-memory wavefronts, issue rate, warp-stall reasons and the opcodes that collect the stall samples.

    python tools/ncu_report.py gpurun_out/r02b_full.ncu-rep > profiles/r02_ncu.txt
"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr = raw[0]
col = {h: i for i, h in enumerate(hdr)}
M = [("time ms", "gpu__time_duration.sum", 1), ("DRAM rd GB", "dram__bytes_read.sum", 1), ("DRAM wr GB", "dram__bytes_write.sum", 1),
     ("FP64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1),
     ("DMMA pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
     ("smem wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
     ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
     ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
     ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("regs", "launch__registers_per_thread", 1),
     ("warp instructions", "smsp__inst_executed.sum", 1)]
print(f"{'kernel':44s} " + " ".join(f"{n:>18s}" for n, _, _ in M))
for r in raw[2:]:
    name = r[col["Kernel Name"]]
    short = re.sub(r"\(.*", "", name.replace("void ", "").replace("pmcb200::", ""))
    vals = []
    for n, key, _ in M:
        v = r[col[key]] if key in col else ""
        try:
            f = float(v)
            vals.append(f"{f:18.4g}")
        except ValueError:
            vals.append(f"{v:>18s}")
    print(f"{short:44s} " + " ".join(vals))
print("(FP64 pipe = sm__pipe_fp64_cycles_active; DMMA pipe = sm__pipe_tensor_cycles_active, 100 % = one DMMA.8x8x4 per 16 cycles and scheduler;\n"
      " warps active = share of the 64-warp maximum; units as ncu prints them: ms, Gbyte, %)\n")

src = ncu("--page", "source", "--csv")
blocks = re.split(r'(?m)^"Kernel Name",', src)[1:]
seen = set()
for b in blocks:
    lines = list(csv.reader(io.StringIO('"Kernel Name",' + b)))
    kname = lines[0][1]
    if kname in seen:  # ncu repeats the listing once per source view
        continue
    seen.add(kname)
    h = lines[1]
    c = {x: i for i, x in enumerate(h)}
    body = [r for r in lines[2:] if len(r) == len(h)]
    def f(r, n):
        try: return float(r[c[n]])
        except Exception: return 0.0
    tot_i = sum(f(r, "Instructions Executed") for r in body)
    tot_s = sum(f(r, "# Samples") for r in body)
    nwarp = nsamp / 32
    print(f"--- {kname}\n    {tot_i / nwarp:.0f} warp instructions per 32 samples, {tot_s:.0f} stall samples")
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    agg = sorted(((sum(f(r, s) for r in body), s) for s in stalls), reverse=True)
    tot_st = sum(a for a, _ in agg) or 1
    print("    stalls: " + ", ".join(f"{s[6:]} {100 * a / tot_st:.1f} %" for a, s in agg[:7]))
    by = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    for r in body:
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[c["Source"]].strip())
        op = m.group(2) if m else "?"
        by[op][0] += f(r, "Instructions Executed"); by[op][1] += f(r, "# Samples")
        by[op][2] += f(r, "L1 Wavefronts Shared"); by[op][3] += f(r, "L1 Wavefronts Shared Ideal")
    print(f"    {'opcode':24s} {'per 32 samples':>14s} {'stall samples %':>16s} {'smem wavefronts':>16s} {'ideal':>8s}")
    for op, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"    {op:24s} {v[0] / nwarp:14.1f} {100 * v[1] / max(tot_s, 1):16.1f} {v[2] / nwarp:16.1f} {v[3] / nwarp:8.1f}")
    print()
