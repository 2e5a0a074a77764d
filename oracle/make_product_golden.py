"""TEST INFRASTRUCTURE -- pins the steps either side of the matched filter to the
reference's OWN code, executed here.

The reference functions cannot be imported: ``spectrometer_masks/masks_sds.py`` parses ``sys.argv`` and opens files at
import time and nests its mask functions inside the per-file loop (:133-233); ``srcfinder_util.py`` imports GDAL;
``triage/cmf_profile.py`` keeps its statistics inside ``summarize`` between file I/O (:110-130); the CNN pipeline
imports rasterio.  So this script parses the reference sources with ``ast``, cuts out exactly

  * get_saturation_mask / get_spec_mask / get_dark_mask / get_cloud_mask   spectrometer_masks/masks_sds.py:133-233
  * extrema, kde                                                           srcfinder_util.py:647-658, 1383-1387
  * the statistics statements of summarize()                               triage/cmf_profile.py:110-133
  * the per-pixel head of filtdet() (KDE weighting, clip, candidate mask)  srcfinder_util.py:1428-1436
  * ClampCH4 and the (mean, std) of every transforms.Normalize             cnn/cnn_pred_pipeline.py:19-30, 126-157

compiles the UNMODIFIED source text of those nodes, runs them on seeded inputs and writes
``tests/golden/{flags,profile,filtdet,cnnnorm}_*.npz`` (inputs + the reference's outputs).  Run once in the build
container (needs ``/root/reference``):   python oracle/make_product_golden.py
"""
from __future__ import annotations

import argparse
import ast
import os
import sys
import textwrap
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from srcfinder_b200 import synth  # noqa: E402

REF = "/root/reference"
MASKS = os.path.join(REF, "spectrometer_masks", "masks_sds.py")
UTIL = os.path.join(REF, "srcfinder_util.py")
PROFILE = os.path.join(REF, "triage", "cmf_profile.py")
CNN = os.path.join(REF, "cnn", "cnn_pred_pipeline.py")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _tree(path):
    with open(path, "r") as fh:
        src = fh.read()
    return src, ast.parse(src)


def extract_defs(path, names):
    """Source text (dedented, unmodified) of the function / class definitions called ``names``, wherever they nest."""
    src, tree = _tree(path)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names and node.name not in out:
            out[node.name] = textwrap.dedent("\n".join(src.splitlines()[node.lineno - 1:node.end_lineno]))
    missing = set(names) - set(out)
    if missing:
        raise RuntimeError("%s: %s not found" % (path, sorted(missing)))
    return out


def module_constants(path, names):
    _, tree = _tree(path)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign):
            for tgt in node.targets:
                if isinstance(tgt, ast.Name) and tgt.id in names:
                    out[tgt.id] = ast.literal_eval(node.value)
                elif isinstance(tgt, ast.Tuple) and isinstance(node.value, ast.Tuple):
                    for t, v in zip(tgt.elts, node.value.elts):
                        if isinstance(t, ast.Name) and t.id in names:
                            out[t.id] = ast.literal_eval(v)
    return out


def argparse_defaults(path):
    """{dest: default} of every parser.add_argument(...) with a literal default."""
    _, tree = _tree(path)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            opts = [a.value for a in node.args if isinstance(a, ast.Constant) and isinstance(a.value, str)]
            dflt = [k.value for k in node.keywords if k.arg == "default"]
            if opts and dflt:
                longest = max(opts, key=len).lstrip("-").replace("-", "_")
                try:
                    out[longest] = ast.literal_eval(dflt[0])
                except ValueError:
                    pass
    return out


def statements_between(path, func, first, last):
    """Unmodified source of the statements of ``func`` (possibly nested) whose lines lie in [first, last]."""
    src, tree = _tree(path)
    lines = src.splitlines()
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func:
            picked = [st for st in node.body if st.lineno >= first and st.end_lineno <= last]
            return textwrap.dedent("\n".join("\n".join(lines[st.lineno - 1:st.end_lineno]) for st in picked))
    raise RuntimeError("%s: %s not found" % (path, func))


# ---------------------------------------------------------------------------------------------- flags
def reference_flag_functions(wavelengths):
    from typing import Optional, Tuple
    defs = extract_defs(MASKS, ["get_saturation_mask", "get_spec_mask", "get_dark_mask", "get_cloud_mask"])
    consts = module_constants(MASKS, ["SAT_THRESH_DEFAULT", "SAT_THRESH_CLD"])
    dflt = argparse_defaults(MASKS)
    args = types.SimpleNamespace(visible_mask_growing_threshold=dflt["visible_mask_growing_threshold"],
                                 dark_threshold=dflt["dark_threshold"])
    ns = {"np": np, "Optional": Optional, "Tuple": Tuple, "args": args, "wavelengths": wavelengths}
    ns.update(consts)
    for name, text in defs.items():
        exec(compile(text, "%s:%s" % (MASKS, name), "exec"), ns)
    return ns, dict(consts, **vars(args))


def flag_cube(L, S, seed):
    """Seeded radiance cube in which every flag fires somewhere (also combined, and at the thresholds)."""
    rng = np.random.default_rng(seed)
    cube = synth.make_cube(L, S, seed=seed, bad_pixels=True)
    n = L * S

    def pick(k):
        idx = rng.choice(n, size=k, replace=False)
        return idx // S, idx % S

    l, s = pick(12); cube[l, 330 + rng.integers(0, 90, 12), s] = 6.0 + rng.uniform(0.001, 2.0, 12)    # saturated
    l, s = pick(6); cube[l, 400, s] = 7.5; cube[l, 25, s] = 9.0 + rng.uniform(0.001, 3.0, 6)          # + specular
    l, s = pick(3); cube[l, 400, s] = 6.0                                                            # == threshold: not saturated
    l, s = pick(10); cube[l, 352, s] = rng.uniform(0.0, 0.104, 10).astype(np.float32)                 # dark
    l, s = pick(3); cube[l, 352, s] = np.float32(0.104)                                               # == threshold: not dark
    l, s = pick(4); cube[l, 352, s] = -9999.0                                                        # no-data is not dark
    l, s = pick(10); cube[l, 15, s] = 15.0 + rng.uniform(0.01, 10.0, 10); cube[l, 60, s] = 9.0        # cloud: bright, falling
    l, s = pick(6); cube[l, 15, s] = 18.0; cube[l, 60, s] = 25.0                                      # bright, rising: no cloud
    l, s = pick(4); cube[l, 15, s] = 18.0; cube[l, 60, s] = 9.0; cube[l, 175, s] = 40.0               # b->c rising (ignored, :231)
    l, s = pick(3); cube[l, 15, s] = np.nan
    return cube


def make_flags():
    wave = synth.load_ch4_library()[:, 1]
    ns, params = reference_flag_functions(wave)
    for name, L, S, seed in (("flags_48x20", 48, 20, 61), ("flags_33x7", 33, 7, 62)):
        cube = flag_cube(L, S, seed)
        data = np.ascontiguousarray(np.transpose(cube, (0, 2, 1)))           # the reference reads BIP blocks (:293-300)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sat = ns["get_saturation_mask"](data, wave)
            spec = ns["get_spec_mask"](data, sat)
            dark = ns["get_dark_mask"](data)
            cloud = ns["get_cloud_mask"](data, wave)
        keep = sorted(set(np.flatnonzero((wave >= 1945) & (wave <= 2485)).tolist()) | {25, 352, 15, 60, 175})
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), cube_kept=cube[:, keep, :], kept_bands=np.array(keep),
                            shape=np.array(cube.shape), wave=wave, saturated=sat, specular=spec, dark=dark, cloud=cloud,
                            sat_thresh=params["SAT_THRESH_DEFAULT"], cloud_thresh=params["SAT_THRESH_CLD"][0],
                            spec_thresh=params["visible_mask_growing_threshold"], dark_thresh=params["dark_threshold"])
        print("%-16s saturated %d specular %d dark %d cloud %d" % (name, sat.sum(), spec.sum(), dark.sum(), cloud.sum()))


# ---------------------------------------------------------------------------------------------- column profiles
def make_profiles():
    ext = extract_defs(UTIL, ["extrema"])
    stats_src = statements_between(PROFILE, "summarize", 110, 133)
    for name, L, S, seed in (("profile_700x9", 700, 9, 71), ("profile_64x5", 64, 5, 72)):
        rng = np.random.default_rng(seed)
        mf = rng.normal(0.0, 350.0, (L, S))
        mf[rng.random((L, S)) < 0.06] = -9999.0
        mf[rng.random((L, S)) < 0.01] = np.nan
        mf[:, 2] = -9999.0                                   # a column without any valid pixel
        if S > 4:
            mf[:, 4] = -np.abs(mf[:, 4])                      # a column without any positive pixel
            mf[5, 3] = 1234.5                                 # a column with a single positive pixel
            mf[np.arange(L) != 5, 3] = -1.0
        prod = np.zeros((L, S, 4))
        prod[..., -1] = mf
        rec = dict(mf=mf)
        for robust in (False, True):
            ns = {"np": np, "cmfmm": prod, "use_robust_stats": robust, "cmflid": name,
                  "cmfimg": types.SimpleNamespace(metadata={"data ignore value": "-9999", "band names": ["a b"] * 4})}
            exec(compile(ext["extrema"], UTIL + ":extrema", "exec"), ns)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                exec(compile(stats_src, PROFILE + ":110-133", "exec"), ns)
            tag = "robust" if robust else "plain"
            for key in ("colnum", "colavg", "colstd", "colmin", "colmax"):
                rec["%s_%s" % (tag, key)] = np.asarray(ns[key])
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **rec)
        print("%-16s plain avg[0] %.4f robust med[0] %.4f" % (name, rec["plain_colavg"][0], rec["robust_colavg"][0]))


# ---------------------------------------------------------------------------------------------- detection pre-filter
def make_filtdet():
    defs = extract_defs(UTIL, ["kde"])
    consts = module_constants(UTIL, ["kernel", "mfmin", "mfmax"])
    head = statements_between(UTIL, "filtdet", 1428, 1436)
    # keep the arithmetic statements only (the prints, the optional file write and the labelling need skimage / GDAL)
    keep = []
    for st in ast.parse(head).body:
        text = ast.get_source_segment(head, st)
        if isinstance(st, ast.Assign) or (isinstance(st, ast.If) and "kde(" in text):
            keep.append(text)
    head = "\n".join(keep)
    for name, L, S, seed in (("filtdet_300x180", 300, 180, 81), ("filtdet_90x40", 90, 40, 82)):
        rng = np.random.default_rng(seed)
        mf = rng.normal(0.0, 300.0, (L, S))
        ll, ss = np.meshgrid(np.arange(L), np.arange(S), indexing="ij")
        for _ in range(3):
            cl, cs, amp = rng.uniform(0.1, 0.9) * L, rng.uniform(0.1, 0.9) * S, rng.uniform(1500, 4000)
            mf += amp * np.exp(-0.5 * (((ll - cl) / 6.0) ** 2 + ((ss - cs) / 4.0) ** 2))
        ns = {"np": np, "ch4mf": mf.copy(), "use_abs": False, "skip_kde": False, "k": consts["kernel"],
              "mfmin": consts["mfmin"], "mfmax": consts["mfmax"], "kde_outf": None}
        exec(compile(defs["kde"], UTIL + ":kde", "exec"), ns)
        exec(compile(head, UTIL + ":filtdet", "exec"), ns)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), mf=mf, detkde=ns["detkde"], ch4min=ns["ch4min"],
                            detmask=ns["detmask"], k=consts["kernel"], mfmin=consts["mfmin"], mfmax=consts["mfmax"])
        print("%-16s candidates %d of %d" % (name, ns["detmask"].sum(), L * S))


# ---------------------------------------------------------------------------------------------- CNN input
def make_cnnnorm():
    import torch
    from torchvision import transforms
    defs = extract_defs(CNN, ["ClampCH4"])
    ns = {"torch": torch}
    exec(compile(defs["ClampCH4"], CNN + ":ClampCH4", "exec"), ns)
    # every transforms.Compose([ClampCH4(...), transforms.Normalize(mean=[..], std=[..])]) of the pipeline, with its model key
    _, tree = _tree(CNN)
    models = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare):
            key = [c.value for c in ast.walk(node.test) if isinstance(c, ast.Constant) and isinstance(c.value, str)]
            for call in ast.walk(ast.Module(body=node.body, type_ignores=[])):
                if isinstance(call, ast.Call) and getattr(call.func, "attr", "") == "Normalize" and key and key[0] != "__main__":
                    kw = {k.arg: ast.literal_eval(k.value) for k in call.keywords}
                    clamp = [c for c in ast.walk(ast.Module(body=node.body, type_ignores=[]))
                             if isinstance(c, ast.Call) and getattr(c.func, "id", "") == "ClampCH4"][0]
                    ck = {k.arg: ast.literal_eval(k.value) for k in clamp.keywords}
                    models[key[0]] = (ck["vmin"], ck["vmax"], kw["mean"][0], kw["std"][0])
    rng = np.random.default_rng(91)
    x = rng.normal(100.0, 900.0, (1, 70, 33)).astype(np.float32)
    x[0, rng.random((70, 33)) < 0.05] = -9999.0
    x[0, 0, :4] = [0.0, 4000.0, 4000.5, -0.5]
    rec = dict(x=x[0], names=np.array(sorted(models)))
    for name in sorted(models):
        vmin, vmax, mean, std = models[name]
        tf = transforms.Compose([ns["ClampCH4"](vmin=vmin, vmax=vmax), transforms.Normalize(mean=[mean], std=[std])])
        rec["out_" + name] = tf(torch.as_tensor(x, dtype=torch.float)).numpy()[0]
        rec["par_" + name] = np.array([vmin, vmax, mean, std], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "cnnnorm_70x33.npz"), **rec)
    print("cnnnorm_70x33    models %s" % sorted(models))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["flags", "profiles", "filtdet", "cnnnorm"])
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    for w in a.what:
        {"flags": make_flags, "profiles": make_profiles, "filtdet": make_filtdet, "cnnnorm": make_cnnnorm}[w]()


if __name__ == "__main__":
    main()
