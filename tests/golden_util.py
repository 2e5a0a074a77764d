"""Loader for tests/golden/*.npz (made by oracle/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# fixtures of other steps (oracle/make_product_golden.py, make_golden.py looshrinkage) live beside the CLI runs
_OTHER = ("flags_", "profile_", "filtdet_", "cnnnorm_", "looshrinkage_")


def case_names():
    """Golden runs of the reference CLI (oracle/make_golden.py CASES)."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if not n.startswith(_OTHER)]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    shape = tuple(int(v) for v in z["shape"])
    cube = np.zeros(shape, dtype=np.float32)
    cube[:, z["kept_bands"], :] = z["cube_kept"]
    case = dict(name=name, cube=cube, flags=json.loads(str(z["flags"])), libname=str(z["libname"]),
                active=[int(z["active"][0]), int(z["active"][1])], product=z["product"],
                header=json.loads(str(z["header"])), stdout_avg=z["stdout_avg"],
                stdout_std=z["stdout_std"])
    if "bgmeta" in z.files:
        case["bgmeta"] = z["bgmeta"]
        case["bgmeta_header"] = json.loads(str(z["bgmeta_header"]))
    flags = case["flags"]
    case["model"] = flags[flags.index("-M") + 1] if "-M" in flags else "looshrinkage"
    case["reflectance"] = "-R" in flags
    case["regfull"] = "-f" in flags
    case["kmodes"] = int(flags[flags.index("-k") + 1]) if "-k" in flags else 1
    case["reject_min"] = int((case["active"][1] - case["active"][0]) * 1.2) if "-r" in flags else 0   # :200
    if case["kmodes"] > 1:
        # the partition the (seeded) reference run used: _bgmeta band 0 holds l, or -l for rejected clusters
        # (:321-327); rejected pixels keep nodata in the score band, so validity needs both bands
        case["labels"] = np.abs(case["bgmeta"][..., 0]).astype(np.int32)
    return case
