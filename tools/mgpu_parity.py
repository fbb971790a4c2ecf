#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):
the sharded iteration (NCCL all-reduce of the sufficient statistics) must give every rank the same proposal, equal to
the single-GPU iteration over the same N_total samples (Philox is counted by the global sample index)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pmclib_b200 import PmcGpu, workloads as W  # noqa: E402
from pmclib_b200.api import nccl_unique_id, shard_range  # noqa: E402
from util import compare_mixture  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
cfg = os.environ.get("MGPU_CONFIG", "c3")   # c3: fused small-d kernels; c5: large-d tensor path (pmc_mma.cu)
wl = W.gauss128() if cfg == "c5" else W.banana()
N = int(os.environ.get("MGPU_N", 400_003 if cfg == "c5" else 4_000_003))
g = PmcGpu(lrank)
g.set_workload(wl)
first, n = shard_range(N, rank, world)
g.resize(n, wl["d"])
ids = [nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
g.comm_init_nccl(ids[0], rank, world)
res = []
retry = 0
for it in range(3):
    r = g.pmc_step(seed=99, it=it, first_index=first, N_total=N, do_update=1, retry=retry)
    retry = r.retry
    res.append(r)
# every rank holds the same proposal
t = torch.from_numpy(np.concatenate([res[-1].wght, res[-1].mean.ravel(), res[-1].L.ravel()])).cuda()
tmax, tmin = t.clone(), t.clone()
dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
same = bool(torch.equal(tmax, tmin))
ok = same
if rank == 0:
    s = PmcGpu(lrank)
    s.set_workload(wl)
    s.resize(N, wl["d"])
    retry = 0
    for it in range(3):
        q = s.pmc_step(seed=99, it=it, first_index=0, N_total=N, do_update=1, retry=retry)
        retry = q.retry
        m = res[it]
        errs = compare_mixture(m.wght, m.mean, m.L, q.wght, q.mean, q.L, rtol=1e-11, what=f"iteration {it}")
        assert m.nok == q.nok and m.n_reject == q.n_reject
        assert abs(m.perplexity - q.perplexity) <= 1e-10 * q.perplexity and abs(m.logSum - q.logSum) <= 1e-11 * abs(q.logSum)
        print(f"it {it}: {world} GPUs vs 1 GPU  perp {m.perplexity:.6f} / {q.perplexity:.6f}  errors {errs}")
    print("MGPU PARITY", "OK" if ok else "FAILED: ranks disagree", "identical proposal on all ranks:", same)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
