"""Tuning sweep of the repack pass (GPU box): one subprocess per CMF_REPACK_VARIANT ("CG,LT,NS:nsplit" also forces the number of line ranges;
CMF_REPACK_PAIR=0 in the environment keeps the 4-byte copy kernel).  Every variant must give the same masks / alpha indices, and scores equal to ~1e-9 sigma."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = 20000, 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
slab[17, :, 5] = float("nan"); slab[123, 7, 200] = -1.0; slab[19999, 71, 597] = float("inf")
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run(timing=False)
    for _ in range(5):
        eng.run(timing=True)
    kt = eng.kernel_times()
    mask = eng.mask(); ai = eng.alpha_index(); cs = eng.colstats()
    print(json.dumps({"repack_ms": kt["repack"], "gram_ms": kt["gram"], "total_ms": sum(kt.values()),
                      "mask_sha": hashlib.sha1(mask.tobytes()).hexdigest()[:12], "masked": int((mask == 0).sum()),
                      "alpha_sha": hashlib.sha1(ai.tobytes()).hexdigest()[:12],
                      "colstd_sum": float(cs[2].sum()), "colavg_absmax": float(np.abs(cs[1]).max())}))
''' % ROOT
out = {}
VARIANTS = ["32,4,4", "32,8,2", "32,4,3", "32,2,8", "16,4,2"]
for v in sys.argv[1:] or VARIANTS:
    env = dict(os.environ, CMF_REPACK_VARIANT=v.split(":")[0])
    if ":" in v:
        env["CMF_REPACK_NSPLIT"] = v.split(":")[1]
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:]
    print(v, line, flush=True)
    out[v] = line
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_repack.json"), "w"), indent=1)
