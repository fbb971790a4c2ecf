"""GPU: pmclib's own C API (include/pmclib_abi.h) implemented by libpmcb200.so, driven by a C program that is written like
a pmclib user (tests/c/dropin_driver.c).  Every iteration the driver records is re-computed by the oracle from the
recorded samples and the proposal that was in force, and must agree."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O
from pmclib_b200 import workloads as W
from pmclib_b200.lib import LIB_PATH
from util import PTOL, RTOL, WTOL, assert_close, compare_mixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("drv") / "dropin_driver")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "dropin_driver.c"),
                    "-o", exe, "-L" + os.path.dirname(LIB_PATH), "-lpmcb200", "-Wl,-rpath," + os.path.dirname(LIB_PATH), "-lm"], check=True)
    return exe


def _write_cfg(cfg, wl, N, niter, seed):
    K, d = wl["K"], wl["d"]
    faces = wl["faces"]
    with open(cfg, "wb") as f:
        f.write(np.array([K, d, N, niter, 0 if faces is None else len(faces), seed], np.int64).tobytes())
        f.write(np.array([wl["target"]["b"], wl["target"]["sig1"]]).tobytes())
        f.write(np.ascontiguousarray(wl["wght"]).tobytes()); f.write(np.ascontiguousarray(wl["mean"]).tobytes())
        f.write(np.ascontiguousarray(wl["cov"]).tobytes())
        if faces is not None:
            f.write(np.ascontiguousarray(faces).tobytes())


def _run_raw(driver, tmp_path, wl, N, niter, mode, env=None, seed=4242):
    """a client rank: no output file to parse, stdout only"""
    cfg, out = tmp_path / f"in{mode}.bin", tmp_path / f"out{mode}.bin"
    _write_cfg(cfg, wl, N, niter, seed)
    r = subprocess.run([driver, str(cfg), str(out), str(mode)], capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def _run(driver, tmp_path, wl, N, niter, mode, env=None, seed=4242):
    K, d = wl["K"], wl["d"]
    faces = wl["faces"]
    cfg, out = tmp_path / f"in{mode}.bin", tmp_path / f"out{mode}.bin"
    with open(cfg, "wb") as f:
        f.write(np.array([K, d, N, niter, 0 if faces is None else len(faces), seed], np.int64).tobytes())
        f.write(np.array([wl["target"]["b"], wl["target"]["sig1"]]).tobytes())
        f.write(np.ascontiguousarray(wl["wght"]).tobytes()); f.write(np.ascontiguousarray(wl["mean"]).tobytes())
        f.write(np.ascontiguousarray(wl["cov"]).tobytes())
        if faces is not None:
            f.write(np.ascontiguousarray(faces).tobytes())
    r = subprocess.run([driver, str(cfg), str(out), str(mode)], capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, np.uint8)
    rec = []
    pos = 0

    def take(dt, n):
        nonlocal pos
        a = raw[pos:pos + n * np.dtype(dt).itemsize].view(dt).copy()
        pos += n * np.dtype(dt).itemsize
        return a
    for _ in range(niter):
        sc = take(np.float64, 8)
        rec.append(dict(perp=sc[0], ess=sc[1], lne=sc[2], logSum=sc[3], maxW=sc[4], maxR=sc[5], retry=int(sc[6]), isLog=int(sc[7]),
                        X=take(np.float64, N * d).reshape(N, d), idx=take(np.uint64, N), w=take(np.float64, N),
                        lr=take(np.float64, N), flg=take(np.int16, N), wght=take(np.float64, K),
                        mean=take(np.float64, K * d).reshape(K, d), L=take(np.float64, K * d * d).reshape(K, d, d)))
    assert pos == len(raw)
    return rec, r.stdout


def _check_against_oracle(rec, wl, what, filt=None):
    cur = dict(wl)
    for it, r in enumerate(rec):
        o = O.orc_iteration(r["X"], r["idx"], cur, do_update=1, in_retry=0 if it == 0 else rec[it - 1]["retry"], filt=filt)
        assert o["rc"] == 0
        big = max(np.max(np.abs(o["log_rho"])), 1.0)
        assert r["isLog"] == 0 and np.array_equal(r["flg"], o["flg"])
        assert_close(r["lr"], o["log_rho"], RTOL, what=f"{what} it{it} log_rho")
        assert_close(r["w"], o["w_norm"], WTOL, scale=np.max(o["w_norm"]) * 1e-3, what=f"{what} it{it} weights")
        assert_close(r["perp"], o["perp"], PTOL, what="perplexity")
        assert_close(r["ess"], o["ess"], PTOL, what="ess")
        assert_close(r["logSum"], o["logSum"], RTOL, scale=big, what="logSum")
        assert_close(r["maxW"], o["maxW"], RTOL, scale=big, what="maxW")
        assert_close(r["maxR"], o["maxR"], RTOL, what="maxR")
        assert r["retry"] == o["retry"]
        errs = compare_mixture(r["wght"], r["mean"], r["L"], o["o_wght"], o["o_mean"], o["o_std"], what=f"{what} it{it}")
        print(f"{what} iteration {it}: perp {r['perp']:.4f} update errors {errs}")
        cur = dict(wl, wght=r["wght"], mean=r["mean"], cov=r["L"], is_chol=True)


def test_c_driver_target_with_gaussian_prior(driver, tmp_path):
    """add_gaussian_prior(init_bentGauss_distribution(...)) -> combine_lkl target: the drop-in layer recognises the terms and
    evaluates the sum on the device (no per-sample host callbacks); every iteration matches the oracle's combined target,
    which is bit-identical to the real reference's (tests/test_cpu_oracle.py)."""
    base = W.banana(d=10, K=10, seed=321, mean_sigma=2.0, var=6.0, box=40.0)
    wl = W.with_gaussian_prior(base, [0, 3, 7], [0.5, -1.0, 2.0], [25.0, 4.0, 9.0])
    rec, stdout = _run(driver, tmp_path, base, 40000, 3, 6)
    _check_against_oracle(rec, wl, "prior")


@pytest.mark.parametrize("mode", [7, 8, 9])
def test_c_driver_libpmc_mpi_entry_points_one_rank(driver, tmp_path, mode):
    """pmc_simu_init_mpi / pmc_simu_init_classic_pmc_mpi / pmc_simu_pmc_step_mpi (pmc_mpi.h:52-101) with a world of one rank:
    mode 7 device target, mode 8 opaque host callback target, mode 9 a hand-written master loop on send_simulation /
    receive_importance_weight / distribution_broadcast (pmc_mpi.c:334-520)."""
    wl = W.banana(d=10, K=10, seed=321, mean_sigma=2.0, var=6.0, box=40.0)
    N, niter = 30000, 3
    env = {k: "" for k in ()}
    rec, stdout = _run(driver, tmp_path, wl, N, niter, mode, env=env)
    _check_against_oracle(rec, wl, f"mpi mode {mode}")
    if mode == 8:
        assert f"user callback calls {N * niter}" in stdout


@pytest.mark.parametrize("mode", [7, 8, 9])
def test_c_driver_libpmc_mpi_two_ranks(driver, tmp_path, mode):
    """Two processes, one GPU each, rank/size from the environment, NCCL id through the rendezvous file: rank 0 ends every
    iteration with the WHOLE sample in its pmc_simu (as the reference's master does) and it matches the oracle; rank 1's
    proposal is identical (digest)."""
    from pmclib_b200.lib import load_library
    if load_library().pmcgpu_device_count() < 2:
        pytest.skip("needs two GPUs")
    import threading
    wl = W.banana(d=10, K=10, seed=321, mean_sigma=2.0, var=6.0, box=40.0)
    N, niter = 30001, 3
    outs = [None, None]

    def go(rank):
        env = {"PMCB200_RANK": str(rank), "PMCB200_WORLD_SIZE": "2", "PMCB200_LOCAL_RANK": str(rank),
               "PMCB200_RENDEZVOUS": str(tmp_path / "rdv")}
        sub = tmp_path / f"r{rank}"
        sub.mkdir()
        outs[rank] = _run(driver, sub, wl, N, niter, mode, env=env) if rank == 0 else _run_raw(driver, sub, wl, N, niter, mode, env)
    th = [threading.Thread(target=go, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    rec, so0 = outs[0]
    so1 = outs[1]
    _check_against_oracle(rec, wl, f"2-rank mpi mode {mode}")
    d0 = [l.split()[-1] for l in so0.splitlines() if "proposal digest" in l]
    d1 = [l.split()[-1] for l in so1.splitlines() if "proposal digest" in l]
    assert len(d0) == niter and d0 == d1
    if mode == 9:  # the reference's slice layout (pmc_mpi.c:344-352): master N/2, client the rest; every sample valid here
        assert f"rank 1 slice {N - N // 2} nok {N - N // 2}" in so1 and f"nok {N} of {N}" in so0
    if mode == 8:
        calls = [int(l.split()[-1]) for l in (so0 + so1).splitlines() if l.startswith("user callback calls")]
        assert sum(calls) == N * niter and min(calls) > 0


@pytest.mark.parametrize("mode", [0, 1, 2, 5])
def test_c_driver_iterations_match_oracle(driver, tmp_path, mode):
    wl = W.banana(d=10, K=10, seed=321, mean_sigma=2.0, var=6.0, box=40.0)
    N, niter = 40000, 3
    rec, stdout = _run(driver, tmp_path, wl, N, niter, mode)
    cur = dict(wl)
    for it, r in enumerate(rec):
        assert np.all(np.abs(r["X"]) <= 40.0)
        o = O.orc_iteration(r["X"], r["idx"], cur, do_update=1, in_retry=0 if it == 0 else rec[it - 1]["retry"])
        assert o["rc"] == 0
        big = max(np.max(np.abs(o["log_rho"])), 1.0)
        assert r["isLog"] == 0 and np.array_equal(r["flg"], o["flg"])
        assert_close(r["lr"], o["log_rho"], RTOL, what=f"it{it} log_rho")
        assert_close(r["w"], o["w_norm"], WTOL, scale=np.max(o["w_norm"]) * 1e-3, what=f"it{it} weights")
        assert_close(r["perp"], o["perp"], PTOL, what="perplexity")
        assert_close(r["ess"], o["ess"], PTOL, what="ess")
        assert_close(r["lne"], o["ln_evidence"], RTOL, scale=big, what="ln evidence")
        assert_close(r["logSum"], o["logSum"], RTOL, scale=big, what="logSum")
        assert_close(r["maxW"], o["maxW"], RTOL, scale=big, what="maxW")
        assert_close(r["maxR"], o["maxR"], RTOL, what="maxR")
        assert r["retry"] == o["retry"]
        errs = compare_mixture(r["wght"], r["mean"], r["L"], o["o_wght"], o["o_mean"], o["o_std"], what=f"mode{mode} it{it}")
        print(f"mode {mode} iteration {it}: perp {r['perp']:.4f} update errors {errs}")
        # the next iteration starts from the proposal the library produced (factors)
        cur = dict(wl, wght=r["wght"], mean=r["mean"], cov=r["L"], is_chol=True)
    if mode == 5:  # prepare_update as its own entry point: histogram of the ok samples, normalised state, snapshot taken
        lines = [l.split() for l in stdout.splitlines() if l.startswith("prepare_update")]
        assert len(lines) == niter
        for l in lines:
            assert l[2] == l[4] and l[6] == "0" and l[8] == "1"
    if mode == 2:
        assert f"user callback calls {N * niter}" in stdout  # the opaque target really ran on the host, once per sample


def test_c_driver_with_histogram_filter(driver, tmp_path):
    """pmc_simu_init_classic_pmc(..., filter_histogram_init(20,1,2), &filter_histogram): the filter runs on the device inside
    pmc_simu_pmc_step; afterwards out_pmc_simu (byte format of pmc.c:306-325) and mean_from_psim."""
    wl = W.banana(d=10, K=10, seed=321, mean_sigma=2.0, var=6.0, box=40.0)
    N, niter = 30000, 2
    rec, stdout = _run(driver, tmp_path, wl, N, niter, 4)
    cur = dict(wl)
    for it, r in enumerate(rec):
        o = O.orc_iteration(r["X"], r["idx"], cur, do_update=1, in_retry=0 if it == 0 else rec[it - 1]["retry"], filt=(20, 1, 2))
        assert o["rc"] == 0 and o["n_filtered"] > 0
        assert np.array_equal(r["flg"], o["flg"])
        assert_close(r["w"], o["w_norm"], WTOL, scale=np.max(o["w_norm"]) * 1e-3, what=f"it{it} weights")
        assert_close(r["perp"], o["perp"], PTOL, what="perplexity")
        assert r["retry"] == o["retry"]
        compare_mixture(r["wght"], r["mean"], r["L"], o["o_wght"], o["o_mean"], o["o_std"], what=f"filter it{it}")
        cur = dict(wl, wght=r["wght"], mean=r["mean"], cov=r["L"], is_chol=True)
    last = rec[-1]
    ok = last["flg"] != 0
    means = {int(l.split()[1]): float(l.split()[2]) for l in stdout.splitlines() if l.startswith("mean ")}
    for a in range(wl["d"]):
        assert_close(means[a], float(np.sum(last["w"][ok] * last["X"][ok, a])), 1e-11, scale=1.0, what=f"mean_from_psim {a}")
    lines = open(str(tmp_path / "out4.bin") + ".txt").read().splitlines()
    assert len(lines) == int(ok.sum())
    first = np.array(lines[0].split(), float)
    i0 = int(np.argmax(ok))
    want = np.array([float("% .5e" % v) for v in list(last["X"][i0]) + [last["w"][i0]]])
    assert np.array_equal(first, want)


@pytest.mark.parametrize("mode", [0, 4])
def test_c_driver_single_process_multi_gpu(driver, tmp_path, mode):
    """PMCB200_GPUS=2: the unmodified driver's pmc_simu_pmc_step shards the sample over two GPUs inside one process
    (host threads + rendezvous all-reduce); every iteration still matches the oracle on the samples it produced."""
    from pmclib_b200.lib import load_library
    if load_library().pmcgpu_device_count() < 2:
        pytest.skip("needs two GPUs")
    wl = W.banana(d=10, K=10, seed=321, mean_sigma=2.0, var=6.0, box=40.0)
    N, niter = 40001, 3
    rec, stdout = _run(driver, tmp_path, wl, N, niter, mode, env={"PMCB200_GPUS": "2"})
    ref, _ = _run(driver, tmp_path, wl, N, niter, mode, env={"PMCB200_GPUS": "1"})
    cur = dict(wl)
    filt = (20, 1, 2) if mode == 4 else None
    for it, (r, q) in enumerate(zip(rec, ref)):
        # Philox is counted by the global sample index: the sharded run draws the very same sample as one GPU
        assert np.array_equal(r["idx"], q["idx"]) and np.array_equal(r["flg"], q["flg"])
        if it == 0:
            assert np.array_equal(r["X"], q["X"])
        o = O.orc_iteration(r["X"], r["idx"], cur, do_update=1, in_retry=0 if it == 0 else rec[it - 1]["retry"], filt=filt)
        assert o["rc"] == 0
        big = max(np.max(np.abs(o["log_rho"])), 1.0)
        assert np.array_equal(r["flg"], o["flg"])
        assert_close(r["lr"], o["log_rho"], RTOL, what=f"it{it} log_rho")
        assert_close(r["w"], o["w_norm"], WTOL, scale=np.max(o["w_norm"]) * 1e-3, what=f"it{it} weights")
        assert_close(r["perp"], o["perp"], PTOL, what="perplexity")
        assert_close(r["logSum"], o["logSum"], RTOL, scale=big, what="logSum")
        assert r["retry"] == o["retry"]
        compare_mixture(r["wght"], r["mean"], r["L"], o["o_wght"], o["o_mean"], o["o_std"], what=f"2 GPUs it{it}")
        cur = dict(wl, wght=r["wght"], mean=r["mean"], cov=r["L"], is_chol=True)
