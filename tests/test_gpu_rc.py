"""GPU: the config / driver layer (pmc_rc.cpp: Lua-free readConf + pmc_rc).  The programs under test are the REFERENCE's own
command-line drivers, pmcexec/vanilla_pmc.c and vanilla_importance.c, compiled unmodified (oracle/Makefile `drivers`, from
the sources where they lie under /root/reference, into oracle/_ref/) and linked against libpmcb200.so instead of
libpmc + libmvdens + liberrorio + liblua.  They read pmclib-format parameter files, run on the GPU and write the reference's
files; the files are parsed back and every iteration is re-computed by the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O
from pmclib_b200 import workloads as W
from util import PTOL, RTOL, WTOL, assert_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def read_psim(path):
    """PMC_SIMD dump (pmc.c:342-371)."""
    raw = open(path, "rb").read()
    assert raw[:8] == b"PMC_SIMD"
    N = int(np.frombuffer(raw, np.int64, 1, 8)[0])
    d, nded, isLog = (int(v) for v in np.frombuffer(raw, np.int32, 3, 16))
    logSum = float(np.frombuffer(raw, np.float64, 1, 28)[0])
    pos = 36
    flg = np.frombuffer(raw, np.int16, N, pos); pos += 2 * N
    w = np.frombuffer(raw, np.float64, N, pos); pos += 8 * N
    lr = np.frombuffer(raw, np.float64, N, pos); pos += 8 * N
    X = np.frombuffer(raw, np.float64, N * d, pos).reshape(N, d); pos += 8 * N * d
    assert pos + 8 * N * nded == len(raw)
    return dict(N=N, d=d, isLog=isLog, logSum=logSum, flg=flg, w=w, lr=lr, X=X)


def read_mix(path):
    """mix_mvdens_dump text (mvdens.c:513-528, 114-134)."""
    tok = open(path).read().split()
    K, d = int(tok[0]), int(tok[1])
    pos = 2
    wght, mean, std, chol, df = np.zeros(K), np.zeros((K, d)), np.zeros((K, d, d)), np.zeros(K, int), np.zeros(K, np.int32)
    for k in range(K):
        wght[k] = float(tok[pos]); pos += 1
        nd, df[k], bdl, chol[k] = int(tok[pos]), int(tok[pos + 1]), int(tok[pos + 2]), int(tok[pos + 3]); pos += 4
        assert nd == d
        mean[k] = [float(t) for t in tok[pos:pos + d]]; pos += d
        std[k] = np.array([float(t) for t in tok[pos:pos + d * d]]).reshape(d, d); pos += d * d
    assert pos == len(tok)
    return dict(K=K, d=d, wght=wght, mean=mean, std=std, chol=chol, df=df)


def _driver(name):
    exe = os.path.join(REFDIR, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    return exe


def test_reference_vanilla_pmc_runs_on_the_gpu(tmp_path):
    exe = _driver("vanilla_pmc")
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "c", "test_pmc.par")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "----- end" in r.stdout
    _check_pmc_run(tmp_path, 4, 30000)


@pytest.mark.parametrize("nranks", [1, 2])
def test_reference_vanilla_pmc_mpi_runs_on_the_gpu(tmp_path, nranks):
    """pmcexec/vanilla_pmc_mpi.c unmodified (MPI_Init / MPI_Comm_rank from the test-only mpi.h stub, everything else through
    libpmc_mpi = libpmcb200.so over NCCL): one process per GPU; rank 0 builds the proposal from the parameter file and
    broadcasts it, every rank draws and weighs its shard, rank 0 writes the reference's files with the WHOLE sample."""
    exe = _driver("vanilla_pmc_mpi")
    from pmclib_b200.lib import load_library
    if load_library().pmcgpu_device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    import threading
    res = [None] * nranks

    def go(rank):
        env = dict(os.environ, PMCB200_RANK=str(rank), PMCB200_WORLD_SIZE=str(nranks), PMCB200_LOCAL_RANK=str(rank),
                   PMCB200_RENDEZVOUS=str(tmp_path / "rdv"))
        res[rank] = subprocess.run([exe, os.path.join(ROOT, "tests", "c", "test_pmc_mpi.par")], cwd=tmp_path, capture_output=True, text=True, env=env)
    th = [threading.Thread(target=go, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    for r in res:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "----- end" in res[0].stdout
    _check_pmc_run(tmp_path, 4, 30001)


def _check_pmc_run(tmp_path, niter, N):
    perp_lines = [l.split() for l in open(tmp_path / "pmc_exploration.perp").read().splitlines()]
    assert len(perp_lines) == niter
    prev = read_mix(tmp_path / "proposal.mix_mvdens")         # the initial proposal (draw_from), saved by rc_mix_mvdens
    # mvdens_ran factorises component 0 in place before the first draw (mvdens.c:287), and the flag and the factor are copied
    # to the other components (pmc_rc.c:548-553): the saved initial proposal already holds Cholesky factors
    assert prev["K"] == 9 and prev["d"] == 2 and np.allclose(prev["wght"], 1 / 9) and np.all(prev["chol"] == 1)
    assert np.allclose(prev["std"][:, 0, 0], np.sqrt(10.0)) and np.allclose(prev["std"][:, 1, 0], 0.0)
    assert np.all(prev["mean"][0] == 0.0) and np.all(np.abs(prev["mean"][1:]) < 20) and len(np.unique(prev["mean"][:, 0])) == 9
    tgt4 = W._target(W.TGT_BANANA, 4, b=0.1, sig1=10.0)
    pins = np.array([0.5, -1.5])
    perps = []
    for it in range(niter):
        ps = read_psim(tmp_path / f"pmc_exploration_{it}.pmc")
        assert ps["N"] == N and ps["d"] == 2 and ps["isLog"] == 0
        assert np.all(np.abs(ps["X"]) <= 20.0)
        # the samples of this iteration against the proposal that was in force (text dump: means to 1e-10, factors to 1e-15)
        cov = prev["std"] if prev["chol"][0] == 0 else None
        wl = dict(K=9, d=2, wght=prev["wght"], mean=prev["mean"], cov=prev["std"], df=prev["df"], target=W._target(W.TGT_BANANA, 2, b=0.1, sig1=10.0),
                  faces=None, is_chol=bool(prev["chol"][0]))
        o = O.orc_iteration(ps["X"], np.zeros(N, np.uint64), wl, do_update=0)
        assert_close(ps["lr"], o["log_rho"], 1e-8, what=f"it{it} log_rho from the dumped proposal")
        full = np.concatenate([ps["X"], np.tile(pins, (N, 1))], axis=1)
        post = O.orc_target_log_pdf(full, dict(target=tgt4))
        lw = post - ps["lr"]
        e = np.exp(lw - lw.max())
        assert_close(ps["w"], e / e.sum(), 1e-10, scale=np.max(e / e.sum()) * 1e-3, what=f"it{it} normalised weights")
        wn = ps["w"][ps["w"] > 0]
        perp = np.exp(-np.sum(wn * np.log(wn))) / N
        assert_close(float(perp_lines[it][0]), perp, 1e-5, what="perplexity column (%g format)")
        perps.append(perp)
        prev = read_mix(tmp_path / f"pmc_exploration_{it}.mix_mvdens")
        assert np.all(prev["chol"][prev["wght"] > 0] == 1)
    assert perps[-1] > perps[0] and perps[-1] > 0.5, perps   # the proposal adapts to the banana
    print("vanilla_pmc perplexities", perps)


def test_reference_vanilla_importance_runs_on_the_gpu(tmp_path):
    """Explicit Student-t components, weights given un-normalised, a Gaussian prior on the target (device-evaluated
    combine), 'var' as scalar / diagonal / full matrix (rc_get_real_assquarematrix)."""
    exe = _driver("vanilla_importance")
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "c", "test_importance.par")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ps = read_psim(tmp_path / "importance.pmc")
    N, d = 200000, 3
    assert ps["N"] == N and ps["d"] == d
    cov = np.array([3.0 * np.eye(3), np.diag([2.0, 2.5, 1.5]), [[2.0, 0.3, 0.0], [0.3, 2.0, 0.1], [0.0, 0.1, 1.0]]])
    base = W._target(W.TGT_MIX, 3, K=1, wght=np.ones(1), mean=np.array([[0.2, -0.1, 0.3]]), cov=np.diag([1.0, 2.0, 0.5])[None], df=np.full(1, -1, np.int32))
    wl = dict(name="imp", K=3, d=3, wght=np.array([0.5, 0.25, 0.25]), mean=np.array([[0, 0, 0], [1, -1, 0.5], [-1, 0.5, 0]], float), cov=cov,
              df=np.full(3, 5, np.int32), target=base, faces=None)
    wl = W.with_gaussian_prior(wl, [0, 2], [0.0, 0.5], [4.0, 9.0])
    o = O.orc_iteration(ps["X"], np.zeros(N, np.uint64), wl, do_update=0)
    assert_close(ps["lr"], o["log_rho"], RTOL, what="log_rho")
    # vanilla_importance dumps BEFORE normalising: the file holds log-weights
    assert ps["isLog"] == 1
    big = max(np.max(np.abs(o["log_rho"])), 1.0)
    assert_close(ps["w"], o["logw"], RTOL, scale=big, what="log-weights")
    line = [l for l in r.stdout.splitlines() if l.startswith("perplexity")][0].split()
    assert_close(float(line[1]), o["perp"], 1e-5, what="perplexity (%g)")
    assert_close(float(line[3]), o["ess"], 1e-5, what="ESS (%g)")
    assert_close(float(line[5]), np.exp(o["ln_evidence"]), 1e-5, what="evidence (%g)")
    # Student-t proposal, Gaussian target with prior: the evidence is the integral of the un-normalised target = Z of the product
    print("vanilla_importance:", line)
