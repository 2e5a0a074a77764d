"""Phase cycles of eigen_ql_kernel from an instrumented build (CMF_NVCC_EXTRA=-DCMF_EIGEN_PROF, loaded through
CMF_B200_LIB): tred2 / accumulate / tql2 / the serial rotation recurrence inside tql2, per column."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = 20000, 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run(); eng.run(timing=True)
    w = eng.sweeps().astype(np.int64)
    kt = eng.kernel_times()
pro, red, tql, epi = (w & 255) * 8192, ((w >> 8) & 255) * 16384, ((w >> 16) & 255) * 32768, ((w >> 24) & 255) * 8192
for name, v in (("prologue (Gram sum, scaling)", pro), ("tred2 + accumulate", red), ("tql2", tql), ("epilogue (P, lam)", epi)):
    print("%-30s median %8.0f kcycles   max %8.0f" % (name, np.median(v) / 1e3, v.max() / 1e3))
print("note: the instrumented build reports phase cycles instead of the iteration count")
print("eigen kernel %.3f ms = %.0f kcycles at 1.965 GHz" % (kt["eigen"], kt["eigen"] * 1.965e3))
