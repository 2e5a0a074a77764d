"""Build libcmf_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library is a plain C-ABI shared object loaded with ctypes)."""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcmf_b200.so")
SOURCES = ["k_stream.cu", "k_gram.cu", "k_eigen.cu", "k_loo.cu", "k_screen.cu", "k_screen5.cu", "k_wide.cu", "k_gram8.cu", "k_modes.cu", "k_cluster.cu", "k_products.cu", "cmf_api.cu", "microbench.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(BUILD, src[:-3] + ".o")
    hdrs = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "cmf_b200.h"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + hdrs):
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("CMF_NVCC_EXTRA", "").split() + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    return obj, res.stderr


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        results = list(pool.map(_compile, SOURCES))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        with open(os.path.join(BUILD, "ptxas.log"), "w") as fh:
            fh.write(log)
    if _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
