"""C-ABI surface: the shared library loads, exports every symbol include/cmf_b200.h declares, and
refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cmf_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    from srcfinder_b200 import _lib
    return _lib.load()


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmf_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from srcfinder_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(cmf_[a-z_0-9]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared but not exported: %s" % missing
    assert sorted(_lib.SYMBOLS) == declared          # the python binding covers the whole header
    for s in declared:
        assert hasattr(lib, s)


def test_kernel_table(lib):
    n = lib.cmf_kernel_count()
    names = [lib.cmf_kernel_name(i).decode() for i in range(n)]
    assert names == ["repack", "mean", "gram", "eigen", "tables", "screen", "select", "loo", "finalize", "score",
                     "colstats"]
    assert b"sm_100a" in lib.cmf_version()


def test_native_code_is_blackwell_native(lib):
    """The shipped SASS is Blackwell-native: 5th-generation tensor-core MMAs with TMEM operands and accumulators
    (tcgen05: UTCHMMA = TF32 screen, UTCIMMA = integer Gram; LDTM / STTM = tcgen05.ld / .st), bulk async copies on the
    TMA engine (UBLKCP) tracked by mbarriers (SYNCS), FP64 tensor MMA (DMMA), built for sm_100a.  The per-kernel
    opcode census is committed under profiles/ (tools/sass_census.py)."""
    from srcfinder_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for op in ("UTCHMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "DMMA"):
        assert op in sass, op
    # per kernel: the screen and the integer Gram must be the tcgen05 kernels
    fn = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fn[cur] = set()
        elif cur:
            for op in ("UTCHMMA", "UTCIMMA", "LDTM", "STTM", "UBLKCP", "DMMA"):
                if op in line:
                    fn[cur].add(op)
    screen = [k for k in fn if "loo_screen5_kernel" in k]
    gram8 = [k for k in fn if "wide_gram8_kernel" in k]
    assert screen and all({"UTCHMMA", "LDTM", "STTM", "UBLKCP"} <= fn[k] for k in screen)
    assert gram8 and all({"UTCIMMA", "LDTM", "UBLKCP"} <= fn[k] for k in gram8)


def test_product_library_ignores_the_environment(lib):
    """Environment tuning hooks, the micro-benchmarks and the cross-check solver live in the tools build only."""
    from srcfinder_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "cmf_microbench" not in out
    und = subprocess.run(["nm", "-D", "--undefined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "getenv" not in und
    tools = subprocess.run(["nm", "-D", "--defined-only", _lib.TOOLS_LIB_PATH], capture_output=True, text=True).stdout
    assert "cmf_microbench" in tools


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    rc = lib.cmf_create(C.byref(ctx), 0)
    assert rc == -2 and not ctx
    assert b"no CPU path" in lib.cmf_last_error(None)
    from srcfinder_b200 import ColumnwiseMF, CmfError
    import numpy as np
    with pytest.raises(CmfError):
        ColumnwiseMF(16, 425, 4, [351, 422], np.zeros(72))


def test_product_never_imports_oracle():
    """The product package must not reach into oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "srcfinder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
