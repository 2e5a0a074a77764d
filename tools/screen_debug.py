"""Debug: screened vs exact alpha search on one synthetic flightline, with the CPU oracle on a few columns."""
import os, sys, json
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth
from oracle import cmf_oracle as orc
seed = int(os.environ.get("SEED", "7")); L = int(os.environ.get("LINES", "20000")); S = 598; active = [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=seed)
print("slab stats", float(slab.min()), float(slab.max()), float(slab.mean()), bool(torch.isfinite(slab).all()))
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run()
    nll_s, tol, ncand, ai_s, st_s = eng.nll(), eng.screen_tol(), eng.ncand(), eng.alpha_index(), eng.status()
    mf_s = eng.mf()
    eng.run(exact=True)
    nll_e, ai_e, st_e = eng.nll(), eng.alpha_index(), eng.status()
    mf_e = eng.mf()
    lam = eng.eigvals()
bad = np.flatnonzero(ai_s != ai_e)
print("mismatching columns", len(bad), bad[:20].tolist())
print("status screened", np.unique(st_s, return_counts=True), "exact", np.unique(st_e, return_counts=True))
print("ncand hist", np.bincount(np.minimum(ncand, 10)).tolist())
print("nonfinite nll_s rows", int((~np.isfinite(nll_s)).any(axis=1).sum()), "nll_e rows", int((~np.isfinite(nll_e)).any(axis=1).sum()))
for c in bad[:3]:
    i_s, i_e = ai_s[c], ai_e[c]
    lo, hi = max(0, min(i_s, i_e) - 2), min(200, max(i_s, i_e) + 3)
    print("col", c, "screened idx", i_s, "exact idx", i_e, "ncand", ncand[c], "tol", tol[c], "lam min/max", lam[c].min(), lam[c].max())
    print("   nll_s", nll_s[c, lo:hi].tolist())
    print("   nll_e", nll_e[c, lo:hi].tolist())
    x = slab[:, :, c].cpu().numpy()
    cube = np.zeros((L, 425, 1), np.float32); cube[:, active[0] - 1:active[1], 0] = x
    ref = orc.cmf_cube(cube, ab, active, keep_nll=True)
    print("   oracle idx", ref["alpha_index"][0], "nll", ref["nll"][0, lo:hi].tolist())
    sd = ref["colstd"][0]
    print("   max|mf_s - ref|/sd", float(np.abs(mf_s[:, c] - ref["mf"][:, 0]).max() / sd), " max|mf_e - ref|/sd", float(np.abs(mf_e[:, c] - ref["mf"][:, 0]).max() / sd))
