"""Tuning sweep of the tiled scoring kernel (GPU box): one subprocess per CMF_SCORE_VARIANT."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = 20000, 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run(timing=False)
    for _ in range(5):
        eng.run(timing=True)
    kt = eng.kernel_times()
    print(json.dumps({"score_ms": kt["score"], "colstats_ms": kt["colstats"], "sum": float(eng.colstats()[2].sum())}))
''' % ROOT
out = {}
for v in ["2,18,2", "2,12,2", "4,8,2", "2,8,3"]:
    env = dict(os.environ, CMF_SCORE_VARIANT=v)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    print(v, line, flush=True)
    out[v] = line
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_score.json"), "w"), indent=1)
