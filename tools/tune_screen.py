"""Tuning sweep of the screening kernel variants (GPU box): one subprocess per CMF_SCREEN_VARIANT.
Reports kernel time, screening error against the all-FP64 search and the refinement load."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = int(os.environ.get('PROBE_L', '20000')), 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], 'cuda', seed=2)
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run(exact=True)
    ai_e, nll_e = eng.alpha_index(), eng.nll()
    eng.run()
    for _ in range(3):
        eng.run(timing=True)
    kt = eng.kernel_times()
    ai_s, nll_s, nc, tol = eng.alpha_index(), eng.nll(), eng.ncand(), eng.screen_tol()
    fin = np.isfinite(nll_e) & np.isfinite(nll_s)
    print(json.dumps({'screen_ms': kt['screen'], 'loo_ms': kt['loo'], 'total_ms': sum(kt.values()),
                      'mismatch': int((ai_e != ai_s).sum()), 'max_err': float(np.abs(nll_e - nll_s)[fin].max()),
                      'refined': int((nc > 1).sum()), 'ncand_max': int(nc.max()),
                      'max_err_over_tol': float((np.where(fin, np.abs(nll_e - nll_s), 0).max(axis=1) / tol).max())}))
""" % ROOT
out = {}
for v in sys.argv[1:] or ["0,3", "0,2", "1,3", "1,2"]:
    env = dict(os.environ, CMF_SCREEN_VARIANT=v)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(v, line, flush=True)
    out[v] = line
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_screen.json"), "w"), indent=1)
