"""Build libumab.so in-tree with nvcc for sm_100a (no torch involved: the library is plain C ABI)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["engine.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tc2.cu", "neighbors.cu", "geometry.cu", "radial.cu",
           "edge_ops.cu", "node_ops.cu", "hessian_ops.cu"]
LIB = os.path.join(HERE, "libumab.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    hdrs = [os.path.join(HERE, "common.cuh"), os.path.join(HERE, "..", "..", "include", "umab.h")]
    hdrs += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        obj = os.path.join(HERE, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for log in logs:
            sys.stderr.write(log)
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
