"""Screen margin sweep on the C2 flightline (GPU box): refined columns, LOO time, certificate measurements, mismatches
against the all-FP64 search."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = int(os.environ.get("MS_L", 20000)), 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
out = []
for seed in (2, 7):
    slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=seed)
    torch.cuda.synchronize()
    with ColumnwiseMF(L, 425, S, active, ab) as eng:
        eng.bind_device(slab.data_ptr())
        eng.run(exact=True)
        ex = eng.alpha_index()
        for m in (2e-5, 1e-5, 5e-6, 2.5e-6, 1.25e-6):
            eng.set_screen_margin(m, certify=True)
            eng.run(); eng.run(timing=True)
            kt = eng.kernel_times()
            chk = eng.screen_check(); nc = eng.ncand(); st = eng.status()
            meas = chk[chk > 0]
            rec = {"seed": seed, "margin": m, "refined_cols": int((nc > 1).sum()), "loo_ms": round(kt["loo"], 3),
                   "total_ms": round(sum(kt.values()), 3), "check_max": float(chk.max()),
                   "check_median": float(np.median(meas)) if len(meas) else 0.0, "rechecked": int(((st & 32) != 0).sum()),
                   "wrong": int((eng.alpha_index() != ex).sum())}
            out.append(rec); print(json.dumps(rec), flush=True)
    del slab
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/margin_sweep.json", "w"), indent=1)
