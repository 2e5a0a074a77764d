"""B200-native columnwise matched filter (drop-in for dsmbgu8/srcfinder cmf/robust_mf.py).

The compute path is libcmf_b200.so (hand-written sm_100a CUDA behind a C ABI, include/cmf_b200.h).
There is no CPU fallback: importing works anywhere, running needs a CUDA device.
"""
from .cmf import ColumnwiseMF, CmfError, alpha_grid, cmf_cube, looshrinkage  # noqa: F401

__all__ = ["ColumnwiseMF", "CmfError", "alpha_grid", "cmf_cube", "looshrinkage"]
