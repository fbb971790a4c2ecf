#!/usr/bin/env python
"""bench.py — PMC samples/s per full iteration (draw + weight + RB update) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples S]

A "step" is one full PMC iteration over one batch of S proposal samples per GPU: configs[2] of BASELINE.json
(10-D banana target, 10-component Gaussian mixture proposal, 1e8 samples per iteration; at N GPUs every rank draws
and weighs its own 1e8-sample shard -> weak scaling, and only the sufficient statistics cross NVLink).
For N > 1 launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libpmcref.so = the reference's own C files compiled
against the GSL shim) on the host cores, same metric/config, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pmclib_b200 import workloads as W  # noqa: E402

METRIC = "PMC samples/s per iteration (draw+weight+RB update)"
UNIT = "samples/s"
FLOP_PER_SAMPLE = 2987.0   # algorithmic flops per sample at d=10, K=10 (banana target)
FLOP_DRAW = 140.0          # of which the draw kernel
FLOP_K1 = 1407.0           # the weights kernel (dominant): mixture log-density 1371 + target 28 + weight 8
FLOP_K2 = 1440.0           # and the statistics kernel: RB update 10 x 144
BYTES_PER_SAMPLE = 106.0   # compulsory HBM bytes per sample: X, weights, log_rho, indices, flg = 8d + 26
FP64_PEAK_TFLOPS = 36.9    # measured DFMA peak on this pool's B200, profiles/r01_fp64_peak.md (MEASURED_PEAKS.json has no FP64 entry)


def measured_hbm_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def workload():
    return W.banana(d=10, K=10, seed=1234, mean_sigma=3.0, var=10.0, box=40.0)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "power_w_max": float(np.max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(N, niter, nworkers, seed=42):
    """Times the compiled reference (vanilla_pmc loop; nworkers>1 = vanilla_pmc_mpi emulation) on the host."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import oracle_py as O
    L = O.ref()
    kind = "reference"
    wl = workload()
    K, d = wl["K"], wl["d"]
    if L is None:
        return None
    tg = wl["target"]
    tp = np.array([tg["b"], tg["sig1"]])
    times = np.zeros(5 * niter); stats = np.zeros(3 * niter)
    faces = wl["faces"]
    t0 = time.time()
    rc = L.ref_pmc_run(C.c_long(N), d, K, O.P(wl["wght"]), O.P(np.ascontiguousarray(wl["mean"])),
                       O.P(np.ascontiguousarray(wl["cov"])), O.P(wl["df"]), tg["kind"], O.P(tp), 0, None, None, None, None,
                       len(faces), O.P(faces), C.c_ulong(seed), niter, 1, nworkers, O.P(times), O.P(stats), None, None, None)
    wall = time.time() - t0
    if rc != 0:
        raise RuntimeError(f"reference run failed rc={rc}")
    per_iter = times.reshape(niter, 5).sum(1)
    return dict(kind=kind, per_iter_s=per_iter, wall=wall, phase_s=times.reshape(niter, 5).sum(0), stats=stats.reshape(niter, 3))


def parity_check(g, wl, rank, world, shard_range, torch, dist, n_par=200_003):
    """Outside the timed region: one sharded iteration over n_par samples in total (ragged shards) must give every rank the
    identical proposal, and - on rank 0, with the shards gathered - must equal the CPU oracle's iteration over the same
    global sample (log_rho, normalised weights, perplexity, logSum, updated mixture) to 1e-12.  The oracle is the checker
    only (tests/oracle_py.py -> oracle/liboracle.so); it is not in any timed region."""
    d, K = wl["d"], wl["K"]
    first, n = shard_range(n_par, rank, world)
    g.set_workload(wl)
    g.resize(n)
    r = g.pmc_step(777, 0, first, n_par, 1, 0)
    got = g.download()
    same = True
    shards = [got]
    if world > 1:
        vec = torch.from_numpy(np.concatenate([r.wght, r.mean.ravel(), r.L.ravel(), [r.perplexity, r.logSum]])).cuda()
        vmax, vmin = vec.clone(), vec.clone()
        dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(vmax, vmin))
        nmax = (n_par + world - 1) // world + 1
        buf = torch.zeros((nmax, d + 4), dtype=torch.float64, device="cuda")
        pack = np.concatenate([got["X"], got["indices"].astype(np.float64)[:, None], got["weights"][:, None],
                               got["log_rho"][:, None], got["flg"].astype(np.float64)[:, None]], axis=1)
        buf[:n] = torch.from_numpy(pack).cuda()
        allb = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf)
        shards = []
        for q in range(world):
            _, nq = shard_range(n_par, q, world)
            a = allb[q][:nq].cpu().numpy()
            shards.append(dict(X=np.ascontiguousarray(a[:, :d]), indices=a[:, d].astype(np.uint64), weights=a[:, d + 1].copy(),
                               log_rho=a[:, d + 2].copy(), flg=a[:, d + 3].astype(np.int16)))
    out = None
    if rank == 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_py as O
            from util import relerr, cov_from_L
            X = np.concatenate([s_["X"] for s_ in shards]); idx = np.concatenate([s_["indices"] for s_ in shards])
            w = np.concatenate([s_["weights"] for s_ in shards]); lr = np.concatenate([s_["log_rho"] for s_ in shards])
            fl = np.concatenate([s_["flg"] for s_ in shards])
            o = O.orc_iteration(X, idx, wl, do_update=1)
            big = max(float(np.max(np.abs(o["log_rho"]))), 1.0)
            errs = {"log_rho": relerr(lr, o["log_rho"]),
                    "weights": relerr(w, o["w_norm"], scale=float(np.max(o["w_norm"])) * 1e-3),
                    "perplexity": abs(r.perplexity - o["perp"]) / o["perp"],
                    "logSum": abs(r.logSum - o["logSum"]) / big,
                    "mixture_weights": relerr(r.wght, o["o_wght"], scale=float(np.max(o["o_wght"]))),
                    "mixture_means": relerr(r.mean, o["o_mean"], scale=float(np.max(np.abs(o["o_mean"])))),
                    "mixture_covariances": max(relerr(cov_from_L(r.L[k:k + 1]), cov_from_L(o["o_std"][k:k + 1]),
                                                      scale=float(np.max(np.abs(cov_from_L(o["o_std"][k:k + 1])))))
                                               for k in range(K) if o["o_wght"][k] > 0)}
            flags_equal = bool(np.array_equal(fl, o["flg"])) and bool(np.array_equal(r.wght > 0, o["o_wght"] > 0))
            worst = max(errs.values())
            out = {"ok": bool(same and flags_equal and worst <= 1e-12), "tolerance": 1e-12, "samples": n_par, "ranks": world,
                   "identical_proposal_on_all_ranks": same, "flags_and_live_components_bit_exact": flags_equal,
                   "max_rel_err_vs_oracle": {k_: float(v_) for k_, v_ in errs.items()},
                   "checker": "CPU oracle (oracle/pmc_oracle.c, pinned bit-identical to the compiled reference) on the gathered shards, outside the timed region"}
        except Exception as ex:  # pragma: no cover
            out = {"ok": False, "error": repr(ex)}
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    N = args.ref_samples
    K, Wm = args.steps, args.warmup
    r = cpu_reference_run(N, K + Wm, cores)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpmcref.so not built and /root/reference absent"}))
        return
    per = r["per_iter_s"][Wm:]
    total = float(np.sum(per))
    val = N * K / total
    line = {
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
        "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": "configs[2]: 10-D banana (b=0.1, sig1^2=10), 10-component Gaussian mixture proposal, box +-40",
                   "samples_per_step": N, "note": "bounded sample of the 1e8-sample iteration; throughput is per sample"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{N} samples/iteration x {K} iterations; vanilla_pmc_mpi emulation: weight loop forked over {cores} "
                                   "workers on send_simulation's slices, draw/normalise/update on rank 0 as pmc_mpi.c does"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def other_configs(g, stream, torch):
    """Short runs of the other BASELINE.json configurations on the same GPU (not the bench metric: reported beside it so that
    the driver's record holds them too).  Per configuration: full iterations (importance-only for configs[3]) timed with CUDA
    events on the library's stream after 3 warm-up iterations."""
    from pmclib_b200.workloads import banana, gmm5, student30, gauss128
    plan = [("configs[0]", banana, 10_000, 1, 20), ("configs[1]", gmm5, 1_000_000, 1, 20),
            ("configs[3]", student30, 10_000_000, 0, 4), ("configs[4]", gauss128, 1_250_000, 1, 4)]
    out = {}
    for name, fac, n, upd, steps in plan:
        try:
            wl = fac()
            g.set_workload(wl)
            g.resize(n, wl["d"])
            it, retry = 0, 0
            for _ in range(3):
                r = g.pmc_step(777, it, 0, n, upd, retry); retry = r.retry; it += 1
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                r = g.pmc_step(777, it, 0, n, upd, retry); retry = r.retry; it += 1
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"workload": wl["name"], "samples_per_step": n, "update": bool(upd), "ms_per_step": ms,
                         "samples_per_s": n / (ms * 1e-3), "perplexity_last": float(r.perplexity), "steps": steps}
        except Exception as ex:  # pragma: no cover
            out[name] = {"error": str(ex)}
    return out


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=100_000_000, help="proposal samples per GPU per iteration")
    ap.add_argument("--ref-samples", type=int, default=1_000_000, help="samples per iteration for the CPU reference arm")
    ap.add_argument("--cpu-baseline-samples", type=int, default=1_000_000)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --samples per GPU (default); strong: --samples in total, sharded over the GPUs (configs[2] as written)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of configs[0], [1], [3], [4]")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-iteration-vs-oracle check made before the warm-up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spl", type=int, default=None)
    ap.add_argument("--opts", default="", help="comma separated pmcgpu_set_option pairs, e.g. split_stats=0,split_draw=0")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from pmclib_b200 import PmcGpu
    from pmclib_b200.api import nccl_unique_id, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = workload()
    g = PmcGpu(local_rank)
    if args.spl:
        g.set_option("spl", args.spl)
    for kv in filter(None, args.opts.split(",")):
        k_, v_ = kv.split("=")
        g.set_option(k_, float(v_))
    g.set_workload(wl)
    N_total = args.samples * world if args.scaling == "weak" else args.samples
    first, n_local = shard_range(N_total, rank, world)
    g.resize(n_local)
    if world > 1:
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g.comm_init_nccl(ids[0], rank, world)
    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local_rank))
    seed = 20261019
    parity = None
    if not args.no_parity:
        parity = parity_check(g, wl, rank, world, shard_range, torch, dist)
        g.set_workload(wl)
        g.resize(n_local)

    def reset():
        g.set_workload(wl)

    # ---- device-resident throughput -------------------------------------------------------------------------------
    it = 0
    retry = 0
    for _ in range(args.warmup):
        r = g.pmc_step(seed, it, first, N_total, 1, retry); retry = r.retry; it += 1
    g.set_option("timing", 1)
    _, _, launches0 = g.get_timing()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    perps = []
    precise = 0
    for _ in range(args.steps):
        r = g.pmc_step(seed, it, first, N_total, 1, retry); retry = r.retry; it += 1
        perps.append(r.perplexity); precise += r.used_precise_path
    e1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    ev_ms = e0.elapsed_time(e1)
    fused_ms, fused_n, launches1 = g.get_timing()
    stats_ms = g.get_stats_timing()
    draw_ms = g.get_draw_timing()
    g.set_option("timing", 0)
    t = torch.tensor([ev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ev_ms, wall_ms = float(t[0]), float(t[1])
    value = N_total * args.steps / (ev_ms * 1e-3)
    n_alive = r.n_alive

    # ---- end to end: host buffers, H2D of the proposal and D2H of the whole sample inside the timed region ----------
    e2e = None
    if not args.no_e2e:
        d, K = wl["d"], wl["K"]
        host = dict(X=torch.empty((n_local, d), dtype=torch.float64).pin_memory().numpy(),
                    indices=torch.empty(n_local, dtype=torch.int64).pin_memory().numpy().view(np.uint64),
                    weights=torch.empty(n_local, dtype=torch.float64).pin_memory().numpy(),
                    log_rho=torch.empty(n_local, dtype=torch.float64).pin_memory().numpy(),
                    flg=torch.empty(n_local, dtype=torch.int16).pin_memory().numpy())
        from pmclib_b200.api import _p
        prop = g.get_proposal()

        # the caller's pmc_simu arrays are attached as the host sink: X and indices cross PCIe while the FP64 kernels
        # run, log_rho after the weight kernel, weights and flags after the normalisation (second stream)
        g._ck(g.lib.pmcgpu_set_host_sink(g.h, _p(host["X"]), _p(host["indices"]), _p(host["weights"]), _p(host["log_rho"]), _p(host["flg"])))

        def e2e_step(it_, retry_):
            # host -> device: the proposal the caller holds (reference layout), then the iteration, which delivers the
            # whole pmc_simu payload into host memory, as pmc_simu_pmc_step's contract promises
            g.set_proposal(prop["wght"], prop["mean"], prop["L"], prop["detL"], None)
            rr = g.pmc_step(seed, it_, first, N_total, 1, retry_)
            prop.update(wght=rr.wght, mean=rr.mean, L=rr.L, detL=rr.detL)
            return rr
        rr = e2e_step(it, retry); it += 1
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        nst = max(1, args.steps)
        for _ in range(nst):
            rr = e2e_step(it, rr.retry); it += 1
        e1.record(stream)
        barrier()
        e_ms = 1e3 * (time.perf_counter() - t0)
        t = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = K * (1 + d + d * d + 1) * 8 + K * 4
        d2h = g.get_info("sink_bytes")  # counted by the library from the copies of the last step
        e2e = {"value": N_total * nst / (float(t[0]) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "host_bytes_delivered_per_step": int(n_local * (8 * d + 8 + 8 + 8 + 2)), "steps": nst,
               "path": "pmcgpu_set_proposal (host mixture) -> pmcgpu_pmc_step with the host sink attached: X, indices, weights, log_rho, flg land in pinned host buffers (D2H on a second stream, overlapped with the kernels; the component indices cross PCIe as bytes and are widened to size_t on the host)"}
        g._ck(g.lib.pmcgpu_set_host_sink(g.h, None, None, None, None, None))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------
    roof = None
    if fused_n > 0:
        ms = fused_ms / fused_n
        ms2 = stats_ms / fused_n
        split = ms2 > 0
        msd = draw_ms / fused_n
        flop = (FLOP_K1 + (0.0 if msd > 0 else FLOP_DRAW)) if split else FLOP_PER_SAMPLE
        tf = flop * n_local / (ms * 1e-3) * 1e-12
        hbm_peak, which = measured_hbm_gbs()
        # compulsory bytes of the dominant kernel: the five pmc_simu arrays (+ the responsibilities it hands to the
        # statistics kernel in the split pipeline)
        bytes_k1 = BYTES_PER_SAMPLE + (8.0 * wl["K"] if split else 0.0)
        if split and msd > 0:
            bytes_k1 = 8.0 * wl["d"] + 8 + 18 + 8.0 * wl["K"]   # reads X and indices, writes weights, log_rho, flg, responsibilities
        roof = {"bound": "tensor", "bound_detail": "FP64 pipe: DFMA and the FP64 tensor instruction DMMA share it on B200 (36.9 / 37.15 TFLOP/s measured, not additive)", "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_PEAK_TFLOPS,
                "traffic": None, "kernel": ("k_weights3 (mixture log-density + target + log-weights + running normaliser)" if msd > 0 else "k_fused2 (draw + mixture log-density + target + log-weights)") if split else "k_fused2 (single pass)",
                "kernel_ms": ms, "kernel_share_of_step": fused_ms / ev_ms, "algorithmic_flop_per_sample": flop,
                "peak_source": "measured DFMA peak, profiles/r01_fp64_peak.md (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm": {"achieved": bytes_k1 * n_local / (ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_k1 * n_local / (ms * 1e-3) * 1e-9 / hbm_peak, "peak_source": which + " MEASURED_PEAKS.json",
                        "algorithmic_bytes_per_sample": bytes_k1}}
        if split:
            tf2 = FLOP_K2 * n_local / (ms2 * 1e-3) * 1e-12
            tfp = FLOP_PER_SAMPLE * n_local / ((ms + ms2 + msd) * 1e-3) * 1e-12
            if msd > 0:
                roof["draw_kernel"] = {"kernel": "k_fused2<MODE 3> (Philox + ziggurat + L g + box)", "kernel_ms": msd,
                                       "kernel_share_of_step": draw_ms / ev_ms,
                                       "hbm_gbs": (8.0 * wl["d"] + 8) * n_local / (msd * 1e-3) * 1e-9}
            roof["stats_kernel"] = {"kernel": "k_stats3 (normalisation + perplexity sums + DMMA sufficient statistics)", "kernel_ms": ms2, "achieved": tf2,
                                    "frac": tf2 / FP64_PEAK_TFLOPS, "algorithmic_flop_per_sample": FLOP_K2,
                                    "kernel_share_of_step": stats_ms / ev_ms}
            roof["pipeline"] = {"achieved": tfp, "frac": tfp / FP64_PEAK_TFLOPS, "algorithmic_flop_per_sample": FLOP_PER_SAMPLE,
                                "note": "all three kernels of the iteration together (draw + weights + statistics)"}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("samples") == n_local:
                roof["traffic"] = tr["dram_bytes_per_launch"]
        except Exception:
            pass

    cpu = None
    if not args.no_cpu_baseline:
        try:
            Nc, itc = args.cpu_baseline_samples, 4
            rr = cpu_reference_run(Nc, itc, 1)
            if rr is not None:
                per = rr["per_iter_s"][1:]
                cpu = {"value": Nc * len(per) / float(np.sum(per)), "unit": UNIT, "cores": 1, "kind": "reference",
                       "sample": f"{Nc} samples/iteration x {len(per)} iterations of the compiled reference (vanilla_pmc loop), same workload",
                       "phase_split_s": {k: float(v) for k, v in zip(["draw", "weight", "normalise", "update", "perplexity"], rr["phase_s"])}}
        except Exception as ex:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {ex}"}

    others = None
    if not args.no_other_configs and world == 1:
        others = other_configs(g, stream, torch)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2]: 10-D banana (b=0.1, sig1^2=10), 10-component Gaussian mixture proposal, box +-40",
                   "samples_per_gpu_per_step": n_local, "samples_per_step": N_total, "parallelism": f"sample shards x{world}",
                   "cache": "inputs larger than L2 (10.6 GB of sample arrays per GPU per step)",
                   "components_alive_at_end": int(n_alive), "perplexity_last": float(perps[-1]),
                   "precise_path_steps": int(precise), "wall_ms_per_step": wall_ms / args.steps},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(launches1 - launches0),
        "parity": parity,
        "other_configs": others,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
