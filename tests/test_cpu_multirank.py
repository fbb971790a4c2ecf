"""CPU, world_size 2 over gloo: the multi-rank host logic.  Each rank holds a shard (pmcgpu_shard_range), forms its
partial sufficient statistics (numpy, test infrastructure), the statistics are all-reduced exactly as the library does
over NCCL (max of the maxima, rescale, sum) and every rank finalises with pmcgpu_finalize_update: all ranks must obtain
the same proposal, equal to the single-process oracle update of the whole sample."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    from pmclib_b200 import workloads as W
    from pmclib_b200.api import shard_range
    from pmclib_b200.lib import load_library
    from test_cpu_abi import numpy_stats
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = load_library()
    wl = W.gmm5()
    N = 30001
    X, idx = O.orc_simulate(N, wl, seed=77)          # every rank regenerates the same global sample ...
    full = O.orc_iteration(X, idx, wl, do_update=1)
    first, n = shard_range(N, rank, world)            # ... and works on its own shard only
    sl = slice(first, first + n)
    K, d = wl["K"], wl["d"]
    # shard-local log-weights relative to the LOCAL maximum, as a rank would hold them
    lw = full["logw"][sl]
    ok = full["flg"][sl] == 1
    M_loc = lw[ok].max()
    mx = torch.tensor([M_loc, full["log_rho"][sl].max()], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    M = float(mx[0])
    w_loc = np.where(ok, np.exp(lw - M_loc), 0.0)
    pivot = (wl["wght"][:, None] * wl["mean"]).sum(0)
    # statistics relative to M_loc, rescaled to the global maximum before the sum (what step_impl does)
    W_, s1, S2 = numpy_stats(X[sl], w_loc, full["log_rho"][sl], full["flg"][sl], wl, float(mx[1]), pivot)
    sc = np.exp(M_loc - M)
    buf = torch.from_numpy(np.concatenate([W_ * sc, (s1 * sc).ravel(), (S2 * sc).ravel(),
                                           np.bincount(idx[sl][ok].astype(np.int64), minlength=K).astype(np.float64)]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    b = buf.numpy()
    W_, s1, S2 = b[:K].copy(), b[K:K + K * d].copy(), b[K + K * d:K + K * d + K * d * d].copy()
    count = b[-K:].astype(np.uint64)
    m = O.Mix(wl["wght"], wl["mean"], wl["cov"], wl["df"])
    wght, mean, L, detL = m.wght.copy(), m.mean.copy(), m.L.copy(), m.detL.copy()
    alive = np.zeros(K, np.int32); retry = C.c_int(0)
    rc = lib.pmcgpu_finalize_update(K, d, O.P(W_), O.P(W_), O.P(s1), O.P(S2), O.P(pivot), 0, O.P(count), N, None,
                                    C.byref(retry), O.P(wght), O.P(mean), O.P(L), O.P(detL), O.P(alive))
    q.put((rank, rc, wght, mean, L, full["o_wght"], full["o_mean"], full["o_std"], int(count.sum()), full["nok"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_statistics_allreduce_and_finalise():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import compare_mixture
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, rc0, w0, m0, L0, ow, om, oL, c0, nok), (r1, rc1, w1, m1, L1, *_rest) = res
    assert rc0 == 0 and rc1 == 0 and c0 == nok
    # identical on every rank (the reduced statistics are bit-identical, the host finalisation is deterministic)
    assert np.array_equal(w0, w1) and np.array_equal(m0, m1) and np.array_equal(L0, L1)
    compare_mixture(w0, m0, L0, ow, om, oL, rtol=1e-11, what="2-rank update")
