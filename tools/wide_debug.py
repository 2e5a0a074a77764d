"""GPU diagnostic of the wide-window kernel set (tools only): every intermediate against numpy / the oracle."""
import json
import os
import sys
import time

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from oracle import cmf_oracle as orc
from srcfinder_b200 import ColumnwiseMF, synth


def run(cube, active, reflectance, exact, model="looshrinkage"):
    L, B, S = cube.shape
    ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
    with ColumnwiseMF(L, B, S, active, ab, reflectance=reflectance, model=model) as eng:
        eng.upload(cube)
        t0 = time.time()
        eng.run(exact=exact)
        dt = time.time() - t0
        res = eng.results()
        res["eig"] = eng.eigvals()
        res["nll"] = eng.nll() if model == "looshrinkage" else None
        res["sweeps"] = eng.sweeps()
        res["dt"] = dt
    return res


def main():
    out = {}
    active = [5, 420]
    L, S = int(os.environ.get("WD_L", 900)), int(os.environ.get("WD_S", 3))
    cube = synth.make_cube(L, S, seed=41, bad_pixels=True)
    ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
    ref = orc.cmf_cube(cube, ab, active, reflectance=True, keep_nll=True)
    for name, exact in (("fp64", True), ("int8", False)):
        try:
            got = run(cube, active, True, exact)
        except Exception as exc:       # noqa: BLE001
            out[name] = {"error": repr(exc)}
            continue
        rec = {"dt": got["dt"], "mask_equal": bool(np.array_equal(got["mask"], ref["mask"])),
               "status": got["status"].tolist(), "ql_iters": got["sweeps"].tolist(),
               "alpha_got": got["alpha_index"].tolist(), "alpha_ref": ref["alpha_index"].tolist()}
        rec["mu_err"] = float(np.max(np.abs(got["mu"] - ref["mu"])))
        cols = []
        for c in range(S):
            ok = ref["mask"][:, c]
            x = np.float64(cube[ok, active[0] - 1:active[1], c])
            cov = np.cov(x.T, ddof=1)
            dinv = 1.0 / np.sqrt(np.diag(cov))
            lam = np.sort(np.linalg.eigvalsh(cov * dinv[:, None] * dinv[None, :]))
            g = np.sort(got["eig"][c])
            nll_r, nll_g = ref["nll"][c], got["nll"][c]
            fin = np.isfinite(nll_r) & np.isfinite(nll_g)
            cols.append({
                "eig_relerr": float(np.max(np.abs(g - lam) / np.maximum(np.abs(lam), 1e-300))),
                "eig_min": float(lam[0]), "eig_max": float(lam[-1]),
                "nll_finite_ref": int(np.isfinite(nll_r).sum()), "nll_finite_got": int(np.isfinite(nll_g).sum()),
                "nll_inf_pattern_equal": bool(np.array_equal(np.isfinite(nll_r), np.isfinite(nll_g))),
                "nll_maxabs": float(np.max(np.abs(nll_r[fin] - nll_g[fin]))) if fin.any() else None,
                "w_relerr": float(np.max(np.abs(got["weights"][c] - ref["weights"][c])) /
                                  np.max(np.abs(ref["weights"][c]))),
                "mf_sigma": float(np.max(np.abs(got["mf"][ok, c] - ref["mf"][ok, c])) / ref["colstd"][c]),
            })
        rec["cols"] = cols
        out[name] = rec
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "wide_debug.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
