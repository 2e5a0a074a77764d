"""Build the matched-filter library in-tree with nvcc for sm_100a (no torch extension machinery needed: plain
C-ABI shared objects loaded with ctypes).

  libcmf_b200.so        the product (include/cmf_b200.h): ignores the environment, one shape per kernel
  libcmf_b200_tools.so  the same sources with -DCMF_TUNING_HOOKS plus microbench.cu (include/cmf_b200_tools.h):
                        environment tuning hooks, kernel variants for the sweeps, the Jacobi cross-check solver and
                        the micro-benchmarks behind the roofline denominators.  Used by tools/, bench.py's peak
                        measurement and the cross-check tests; never by the product path.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
BUILD_TOOLS = os.path.join(HERE, "build", "tools")
LIB = os.path.join(HERE, "libcmf_b200.so")
LIB_TOOLS = os.path.join(HERE, "libcmf_b200_tools.so")
SOURCES = ["k_stream.cu", "k_gram.cu", "k_eigen.cu", "k_loo.cu", "k_screen.cu", "k_screen5.cu", "k_wide.cu",
           "k_gram8.cu", "k_modes.cu", "k_cluster.cu", "k_products.cu", "cmf_api.cu"]
TOOLS_SOURCES = SOURCES + ["microbench.cu"]
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


def _compile(job):
    src, outdir, extra = job
    obj = os.path.join(outdir, src[:-3] + ".o")
    hdrs = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    hdrs += [os.path.join(inc, h) for h in os.listdir(inc) if h.endswith(".h")]
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + hdrs):
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + extra + os.environ.get("CMF_NVCC_EXTRA", "").split() + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    return obj, res.stderr


def _link(lib, objs):
    if _stale(lib, objs):
        cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))


def build(force=False, verbose=False, tools=True):
    os.makedirs(BUILD_TOOLS, exist_ok=True)
    if force:
        for d in (BUILD_TOOLS, BUILD):
            for f in os.listdir(d):
                if f.endswith((".o", ".log")):
                    os.remove(os.path.join(d, f))
    jobs = [(s, BUILD, []) for s in SOURCES]
    if tools:
        jobs += [(s, BUILD_TOOLS, ["-DCMF_TUNING_HOOKS"]) for s in TOOLS_SOURCES]
    with cf.ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 8)) as pool:
        results = list(pool.map(_compile, jobs))
    log = "".join(l for _, l in results[:len(SOURCES)])
    if log:
        with open(os.path.join(BUILD, "ptxas.log"), "w") as fh:
            fh.write(log)
    _link(LIB, [o for o, _ in results[:len(SOURCES)]])
    if tools:
        _link(LIB_TOOLS, [o for o, _ in results[len(SOURCES):]])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
