#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (bench.py measures configs[2], the headline).

    python tools/bench_configs.py --config c5 [--samples N_per_gpu] [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_configs.py --config c5 --samples 1250000

One JSON line per run (rank 0): whole-job samples/s of a full iteration (draw + weights + normalise + perplexity
[+ RB update]), timed on the device stream between barriers (max over ranks), plus the share and the FP64 roofline
fraction of the tensor-path kernels (CUDA events inside the library; algorithmic flops of SURVEY.md section 8d).
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pmclib_b200 import workloads as W  # noqa: E402

FP64_PEAK_TFLOPS = 36.9  # measured, profiles/r01_fp64_peak.md

CFG = {
    # name: (factory, default samples per GPU per step, update?, description)
    "c1": (W.banana, 10_000, 1, "configs[0]: 10-D banana, 10-component Gaussian mixture proposal, 10k samples/iteration (the reference's CPU-runnable case)"),
    "c2": (W.gmm5, 1_000_000, 1, "configs[1]: 5-D 4-mode Gaussian-mixture target, 20-component Gaussian proposal, 1M samples/iteration"),
    "c4": (W.student30, 10_000_000, 0, "configs[3]: 30-D Gaussian target, fixed 50-component Student-t (nu=9) proposal, importance only, 1e9 samples streamed in steps"),
    "c5": (W.gauss128, 1_250_000, 1, "configs[4]: 128-D Gaussian target, 64-component Gaussian mixture proposal, 1e7 samples/iteration over 8 GPUs"),
}


def flops(wl, upd):
    """F(d,K,T) split into draw / proposal density / target / weight+norm / RB statistics."""
    d, K = wl["d"], wl["K"]
    t = wl["target"]
    st = 3 if np.any(np.asarray(wl["df"]) != -1) else 0
    comp = d * d + 3 * d + 4
    if t["kind"] == W.TGT_BANANA:
        T = 2 * d + 8
    elif t["kind"] == W.TGT_MVDENS:
        T = comp
    else:
        T = t["K"] * comp + 3 * t["K"] + 1
    return dict(draw=d * d + 4 * d, density=K * (comp + st) + 3 * K + 1, target=T, weight=8,
                stats=(K * (d * d + 4 * d + 4)) if upd else 0)


CPU_SAMPLE = {"c1": (10_000, 11), "c2": (1_000_000, 3), "c4": (100_000, 3), "c5": (2_000, 1)}  # (samples/iteration, iterations); c5: the reference's update runs at ~80 samples/s (denormals)


def cpu_reference(cfg, wl, upd, nworkers=1):
    """The compiled reference (oracle/_ref, vanilla_pmc loop) on the host cores, bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import oracle_py as O
    L = O.ref()
    if L is None:
        return None
    N, niter = CPU_SAMPLE[cfg]
    K, d, tg = wl["K"], wl["d"], wl["target"]
    tp = np.array([tg["b"], tg["sig1"]])
    times = np.zeros(5 * niter); stats = np.zeros(3 * niter)
    faces = wl["faces"]
    ca = lambda a, t=np.float64: None if a is None else np.ascontiguousarray(a, t)
    tw, tm, tc, tdf = ca(tg["wght"]), ca(tg["mean"]), ca(tg["cov"]), ca(tg["df"], np.int32)
    rc = L.ref_pmc_run(C.c_long(N), d, K, O.P(ca(wl["wght"])), O.P(ca(wl["mean"])), O.P(ca(wl["cov"])), O.P(ca(wl["df"], np.int32)),
                       tg["kind"], O.P(tp), tg["K"], O.P(tw), O.P(tm), O.P(tc), O.P(tdf),
                       0 if faces is None else len(faces), O.P(ca(faces)), C.c_ulong(42), niter, upd, nworkers,
                       O.P(times), O.P(stats), None, None, None)
    per = times.reshape(niter, 5).sum(1)
    per = per[1:] if niter > 1 else per
    return {"value": N * len(per) / float(np.sum(per)), "unit": "samples/s", "cores": nworkers, "kind": "reference", "rc": int(rc),
            "sample": f"{N} samples/iteration x {len(per)} timed iteration(s) of the compiled reference, same workload",
            "phase_split_s": {k: float(v) for k, v in zip(["draw", "weight", "normalise", "update", "perplexity"], times.reshape(niter, 5).sum(0))}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=list(CFG))
    ap.add_argument("--samples", type=int, default=0, help="samples per GPU per step")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--opts", default="")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pmclib_b200 import PmcGpu
    from pmclib_b200.api import nccl_unique_id, shard_range
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fac, n_def, upd, desc = CFG[args.config]
    wl = fac()
    n_gpu = args.samples or n_def
    g = PmcGpu(local)
    for kv in filter(None, args.opts.split(",")):
        k_, v_ = kv.split("=")
        g.set_option(k_, float(v_))
    g.set_workload(wl)
    N_total = n_gpu * world
    first, n_local = shard_range(N_total, rank, world)
    g.resize(n_local, wl["d"])
    if world > 1:
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g.comm_init_nccl(ids[0], rank, world)
    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local))
    seed, it, retry = 20261020, 0, 0
    for _ in range(args.warmup):
        r = g.pmc_step(seed, it, first, N_total, upd, retry); retry = r.retry; it += 1
    g.set_option("timing", 1)
    _, _, l0 = g.get_timing()
    from bench import ClockSampler  # nvidia-smi clocks / throttle reasons during the timed region
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    perps, precise = [], 0
    for _ in range(args.steps):
        r = g.pmc_step(seed, it, first, N_total, upd, retry); retry = r.retry; it += 1
        perps.append(r.perplexity); precise += r.used_precise_path
    e1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    ev_ms = e0.elapsed_time(e1)
    q_ms, q_n, l1 = g.get_timing()
    s_ms, d_ms = g.get_stats_timing(), g.get_draw_timing()
    t = torch.tensor([ev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ev_ms, wall_ms = float(t[0]), float(t[1])
    if rank == 0:
        F = flops(wl, upd)
        Ftot = sum(F.values())
        step_ms = wall_ms / args.steps  # host finalisation is part of an iteration: wall clock between barriers
        value = N_total * args.steps / (wall_ms * 1e-3)
        kern = {}
        if q_n > 0 and wl["d"] <= 16:   # fused small-d kernels (pmc_fused2.cu): same timers, other kernels
            kern["weights kernel k_fused2"] = {"ms": q_ms / q_n, "share_of_step": q_ms / q_n / step_ms}
            if s_ms > 0: kern["statistics kernel k_stats2"] = {"ms": s_ms / q_n, "share_of_step": s_ms / q_n / step_ms}
            if d_ms > 0: kern["draw kernel k_fused2<MODE 3>"] = {"ms": d_ms / q_n, "share_of_step": d_ms / q_n / step_ms}
        elif q_n > 0:
            ms = q_ms / q_n
            tf = (F["density"] + (F["target"] if wl["target"]["kind"] != W.TGT_BANANA else 0)) * n_local / (ms * 1e-3) * 1e-12
            kern["whitening k_q_mma (proposal + target densities)"] = {"ms": ms, "share_of_step": ms / step_ms, "algorithmic_tflops": tf, "frac_of_fp64_peak": tf / FP64_PEAK_TFLOPS}
            if s_ms > 0:
                ms2 = s_ms / q_n
                tf2 = F["stats"] * n_local / (ms2 * 1e-3) * 1e-12
                kern["statistics k_resp_q + k_centre + k_stats_mma"] = {"ms": ms2, "share_of_step": ms2 / step_ms, "algorithmic_tflops": tf2, "frac_of_fp64_peak": tf2 / FP64_PEAK_TFLOPS}
            if d_ms > 0:
                kern["draw k_draw_pick/perm/mma"] = {"ms": d_ms / q_n, "share_of_step": d_ms / q_n / step_ms}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu = cpu_reference(args.config, wl, upd)
            except Exception as ex:  # pragma: no cover
                cpu = {"value": None, "sample": f"failed: {ex}"}
        print(json.dumps({
            "metric": "PMC samples/s per iteration (draw+weight" + ("+RB update)" if upd else ", importance only)"),
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "device_ms_per_step": ev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "samples_per_gpu_per_step": n_local, "samples_per_step": N_total,
                       "perplexity": [float(p) for p in perps], "components_alive_at_end": int(r.n_alive),
                       "precise_path_steps": int(precise)},
            "roofline": {"bound": "tensor", "bound_detail": "FP64 pipe: DFMA and the FP64 tensor instruction DMMA share it on B200 (36.9 / 37.15 TFLOP/s measured, not additive)", "algorithmic_flop_per_sample": Ftot, "achieved": value / world * Ftot * 1e-12,
                         "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": value / world * Ftot * 1e-12 / FP64_PEAK_TFLOPS,
                         "kernels": kern},
            "cpu_baseline": cpu, "clocks": clocks, "gpu_launches": int(l1 - l0)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
