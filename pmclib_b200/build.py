"""Builds pmclib_b200/libpmcb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pmclib_b200.build [--force]

Also builds the test oracles (oracle/Makefile).  Nothing here needs torch.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpmcb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-I/usr/include"]
SOURCES = ["pmc_iter3.cu", "pmc_fused2.cu", "pmc_fused.cu", "pmc_mma.cu", "pmc_generic.cu", "pmc_post.cu", "pmc_api.cu", "pmc_host.cpp", "pmc_dropin.cpp", "pmc_dropin2.cpp", "pmc_dropin_mpi.cpp", "pmc_rc.cpp", "pmc_hostio.cpp"]


def _deps():
    out = []
    for root in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh", ".cu", ".cpp")):
                out.append(os.path.join(root, f))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [NVCC] + CUFLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force or _stale(LIB, _deps()):
        if verbose:
            print(f"[pmclib_b200.build] nvcc {' '.join(ARCH)} -> {LIB}", file=sys.stderr)
        with cf.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
            objs = list(ex.map(_compile, srcs))
        r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_oracles() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    build_oracles()
    print(LIB)
