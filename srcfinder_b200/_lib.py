"""ctypes binding of libcmf_b200.so (include/cmf_b200.h).  No fallback: if the library is missing or
no CUDA device is present the caller gets an exception."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CMF_B200_LIB: development hook for other builds of the same library (tools/ and the cross-check tests load
# libcmf_b200_tools.so through it in a subprocess)
LIB_PATH = os.environ.get("CMF_B200_LIB") or os.path.join(_HERE, "libcmf_b200.so")
TOOLS_LIB_PATH = os.path.join(_HERE, "libcmf_b200_tools.so")

# every symbol include/cmf_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "cmf_create", "cmf_destroy", "cmf_last_error", "cmf_version", "cmf_set_stream", "cmf_set_problem",
    "cmf_upload_bil", "cmf_upload_lines", "cmf_upload_wait", "cmf_bind_device_slab", "cmf_set_labels", "cmf_set_clustering", "cmf_set_regfull", "cmf_set_exclusion", "cmf_run", "cmf_sync", "cmf_run_host", "cmf_download",
    "cmf_device_ptr", "cmf_output_bytes", "cmf_kernel_count", "cmf_kernel_name", "cmf_kernel_times",
    "cmf_launch_count", "cmf_screen_kernel", "cmf_host_alloc", "cmf_host_free", "cmf_host_register", "cmf_host_unregister",
    "cmf_pixel_flags", "cmf_column_profile", "cmf_column_profile_image",
    "cmf_detection_prefilter", "cmf_cnn_input", "cmf_looshrinkage", "cmf_set_screen_margin",
]

OUT_MF, OUT_MASK, OUT_COLSTATS, OUT_ALPHA_INDEX, OUT_NLL, OUT_MU, OUT_WEIGHTS, OUT_STATUS, OUT_NVALID, \
    OUT_EIGVALS, OUT_SWEEPS, OUT_NCAND, OUT_SCREEN_TOL, OUT_CLUSTER_ID, OUT_ALPHA_IMAGE, OUT_MODE_LIST, OUT_LABELS, OUT_PCA, \
    OUT_KMEANS_ITERS, OUT_FLAGS, OUT_SCREEN_CHECK = range(21)
RUN_TIMING = 1
RUN_EXACT = 2
RUN_ASYNC = 4
MODEL_LOOSHRINKAGE, MODEL_EMPIRICAL = 0, 1
FLAG_SATURATED, FLAG_SPECULAR, FLAG_DARK, FLAG_CLOUD = 1, 2, 4, 8
COL_EMPTY, COL_DEGENERATE, COL_SINGULAR, COL_NOCONVERGE, COL_ALLINF, COL_RECHECKED = 1, 2, 4, 8, 16, 32


class Problem(C.Structure):
    _fields_ = [
        ("lines", C.c_int32), ("bands", C.c_int32), ("samples", C.c_int32), ("interleave", C.c_int32),
        ("band_lo", C.c_int32), ("band_hi", C.c_int32), ("reflectance", C.c_int32), ("model", C.c_int32),
        ("num_alphas", C.c_int32), ("reserved", C.c_int32), ("nodata", C.c_double),
        ("alphas", C.POINTER(C.c_double)), ("abscf", C.POINTER(C.c_double)),
    ]


class FlagSpec(C.Structure):
    _fields_ = [
        ("sat_lo", C.c_int32), ("sat_hi", C.c_int32), ("spec_band", C.c_int32), ("dark_band", C.c_int32),
        ("cloud_a", C.c_int32), ("cloud_b", C.c_int32), ("sat_thresh", C.c_float), ("spec_thresh", C.c_float),
        ("dark_thresh", C.c_float), ("cloud_thresh", C.c_float), ("cloud_dwl", C.c_float),
    ]


_lib = None


def load():
    """Load the shared library and declare every prototype.  Raises OSError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("libcmf_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(srcfinder_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, i64, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64, C.c_size_t
    sig = {
        "cmf_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "cmf_destroy": (None, [vp]),
        "cmf_last_error": (C.c_char_p, [vp]),
        "cmf_version": (C.c_char_p, []),
        "cmf_set_stream": (C.c_int, [vp, vp]),
        "cmf_set_problem": (C.c_int, [vp, C.POINTER(Problem)]),
        "cmf_upload_bil": (C.c_int, [vp, vp]),
        "cmf_upload_lines": (C.c_int, [vp, vp, i32, i32, i32, i32]),
        "cmf_upload_wait": (C.c_int, [vp, vp]),
        "cmf_bind_device_slab": (C.c_int, [vp, vp, i64, i32]),
        "cmf_set_labels": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "cmf_set_clustering": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "cmf_set_regfull": (C.c_int, [vp, C.c_int]),
        "cmf_set_exclusion": (C.c_int, [vp, vp]),
        "cmf_set_screen_margin": (C.c_int, [vp, C.c_double, C.c_int]),
        "cmf_run": (C.c_int, [vp, u32]),
        "cmf_sync": (C.c_int, [vp]),
        "cmf_run_host": (C.c_int, [vp, vp, vp, vp, vp, u32]),
        "cmf_download": (C.c_int, [vp, C.c_int, vp, sz]),
        "cmf_device_ptr": (vp, [vp, C.c_int]),
        "cmf_output_bytes": (sz, [vp, C.c_int]),
        "cmf_kernel_count": (C.c_int, []),
        "cmf_kernel_name": (C.c_char_p, [C.c_int]),
        "cmf_kernel_times": (C.c_int, [vp, C.POINTER(C.c_float), C.c_int]),
        "cmf_launch_count": (C.c_int, [vp]),
        "cmf_screen_kernel": (C.c_char_p, [vp]),
        "cmf_host_alloc": (vp, [sz]),
        "cmf_host_free": (None, [vp]),
        "cmf_host_register": (C.c_int, [vp, sz]),
        "cmf_host_unregister": (C.c_int, [vp]),
        "cmf_pixel_flags": (C.c_int, [vp, vp, C.c_int, i32, i32, i32, C.POINTER(FlagSpec), vp]),
        "cmf_column_profile": (C.c_int, [vp, C.c_int, C.c_double, vp]),
        "cmf_column_profile_image": (C.c_int, [vp, vp, i32, i32, C.c_double, C.c_int, C.c_double, vp]),
        "cmf_detection_prefilter": (C.c_int, [vp, vp, i32, i32, vp, i32, C.c_double, C.c_double, vp, vp, vp]),
        "cmf_looshrinkage": (C.c_int, [vp, vp, i32, i32, vp, i32, i32, vp, i32, vp, vp, vp]),
        "cmf_cnn_input": (C.c_int, [vp, vp, i32, i32, C.c_float, C.c_float, C.c_float, C.c_float, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


_tools = None


def load_tools():
    """The tools build (include/cmf_b200_tools.h): only its micro-benchmark entry is declared here."""
    global _tools
    if _tools is None:
        if not os.path.exists(TOOLS_LIB_PATH):
            raise OSError("libcmf_b200_tools.so is not built: run `python -m srcfinder_b200.build`")
        _tools = C.CDLL(TOOLS_LIB_PATH)
        _tools.cmf_microbench.restype = C.c_double
        _tools.cmf_microbench.argtypes = [C.c_int, C.c_int, C.c_int]
    return _tools
