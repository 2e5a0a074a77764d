"""Device-resident time of the wide-window (-R, bands 5..420) path at flightline size, per kernel (GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth

L = int(os.environ.get("TW_L", 20000)); S = int(os.environ.get("TW_S", 598))
active = [5, 420]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=3)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
out = {"lines": L, "samples": S, "active": active}
with ColumnwiseMF(L, 425, S, active, ab, reflectance=True, stream=st) as eng:
    eng.bind_device(slab.data_ptr())
    for mode, exact in (("int8_gram", False), ("fp64_gram", True)):
        if exact and os.environ.get("TW_SKIP_EXACT"):
            continue
        eng.run(exact=exact)
        eng.run(timing=True, exact=exact)
        kt = eng.kernel_times()
        tot = sum(kt.values())
        out[mode] = {"ms": tot, "mpixel_s": L * S / tot / 1e3, "kernels_ms": kt}
        print(mode, json.dumps(out[mode]), flush=True)
    ai = eng.alpha_index(); stt = eng.status(); cs = eng.colstats()
    out["alpha_index_hist"] = np.bincount(ai[ai >= 0], minlength=201).tolist()
    out["status_nonzero"] = int((stt != 0).sum())
    out["ql_iters_mean"] = float(eng.sweeps().mean())
    out["colavg_over_colstd_max"] = float(np.max(np.abs(cs[1] / cs[2])))
    w = eng.weights(); mu = eng.mu()
    out["w_dot_t_minus_1_max"] = float(np.max(np.abs(np.sum(w * (ab[None, :] - mu), axis=1) - 1.0)))
    out["gpu_mem_gb"] = torch.cuda.mem_get_info()[1] / 2**30 - torch.cuda.mem_get_info()[0] / 2**30
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_wide.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "alpha_index_hist"}, indent=1))
