"""Development probe run on the GPU box: micro-benchmarks + per-kernel timings at flightline size.
Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from srcfinder_b200 import ColumnwiseMF, _lib, synth

out = {}
lib = _lib.load()
names = {0: "dmma_tflops_32w", 8: "dmma_tflops_8w_ilp24", 1: "dfma_tflops", 6: "cvt_f32_f64_gops",
         7: "logdiv_gops", 9: "mma_sync_tf32_tflops", 10: "mma_sync_bf16_tflops", 11: "ffma_tflops", 2: "hbm_read_8B_gbs", 3: "hbm_read_16B_gbs", 4: "hbm_copy_gbs", 5: "hbm_bulk_read_gbs"}
if os.environ.get("PROBE_TC5", "1") == "1":
    # kind 21 (LBO/SBO swapped on purpose) faults and poisons the context: development use only
    for kind, name in {20: "tc5_selftest_n80", 22: "tc5_selftest_n96_rowoff112",
                       23: "tc5_selftest_n112", 24: "tc5_selftest_n72", 25: "tc5_selftest_n256_k8"}.items():
        out[name] = _lib.load_tools().cmf_microbench(0, kind, 1)
        print(name, out[name], flush=True)
if os.environ.get("PROBE_MICRO", "1") != "1":
    names = {}
for kind, name in names.items():
    out[name] = _lib.load_tools().cmf_microbench(0, kind, 5)
    print(name, out[name], flush=True)

L = int(os.environ.get("PROBE_L", "20000"))
S, active = 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
t0 = time.time()
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
torch.cuda.synchronize()
print("slab generated in %.1f s" % (time.time() - t0), flush=True)
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    for i in range(3):
        eng.run(timing=True)
        kt = eng.kernel_times()
        print(i, {k: round(v, 3) for k, v in kt.items()}, "total %.3f ms" % sum(kt.values()), flush=True)
    # screening quality: exact FP64 search vs screened search on the same flightline
    ai_s, nll_s, ncand = eng.alpha_index(), eng.nll(), eng.ncand()
    eng.run(timing=True, exact=True)
    kte = eng.kernel_times()
    ai_e, nll_e = eng.alpha_index(), eng.nll()
    fin = np.isfinite(nll_e) & np.isfinite(nll_s)
    err = np.abs(nll_e - nll_s)
    d_e, d_s = np.diff(nll_e, axis=1), np.diff(nll_s, axis=1)
    out["screen"] = {"index_mismatch": int((ai_s != ai_e).sum()), "max_abs_nll_err": float(err[fin].max()),
                     "max_abs_adjacent_diff_err": float(np.nanmax(np.abs(d_e - d_s))),
                     "ncand_hist": np.bincount(np.minimum(ncand, 10)).tolist(),
                     "refined_columns": int((ncand > 1).sum()), "exact_kernel_ms": kte}
    print("screen", out["screen"], flush=True)
    out["kernel_ms"] = kt
    out["total_ms"] = sum(kt.values())
    out["mpixel_s"] = L * S / (sum(kt.values()) * 1e-3) / 1e6
    out["sweeps_max"] = int(eng.sweeps().max())
    out["alpha_index_range"] = [int(eng.alpha_index().min()), int(eng.alpha_index().max())]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out))
