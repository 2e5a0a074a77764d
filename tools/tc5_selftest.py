"""tcgen05 self tests of the screening kernel's building blocks, one subprocess per case (a trap in one
case must not take the CUDA context of the others with it).  Writes gpurun_out/tc5_selftest.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {20: "n80_k72", 21: "n80_k72_swapped_lbo_sbo", 22: "n96_k72_rowoff112", 23: "n112_k72", 24: "n72_k72",
         25: "n256_k8"}
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    from srcfinder_b200 import _lib
    print("RESULT", _lib.load_tools().cmf_microbench(0, int(sys.argv[1]), 1))
    sys.exit(0)
out = {}
for kind, name in CASES.items():
    try:
        r = subprocess.run([sys.executable, __file__, str(kind)], capture_output=True, text=True, timeout=120)
        val = [l.split()[1] for l in r.stdout.splitlines() if l.startswith("RESULT")]
        out[name] = {"rc": r.returncode, "max_abs_err": float(val[0]) if val else None, "stderr": r.stderr[-300:]}
    except subprocess.TimeoutExpired:
        out[name] = {"rc": "timeout"}
    print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tc5_selftest.json"), "w"), indent=1)
