"""Cycles per TS-form tcgen05 TF32 MMA as a function of N (tools build, GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srcfinder_b200 import _lib
lib = _lib.load_tools()
out = {}
for n16 in (2, 3, 4, 5, 6, 7, 8, 10, 12, 13, 14, 16):
    k = 40 if n16 == 16 else 40 + n16
    lib.cmf_microbench(0, k, 10)
    out[16 * n16] = round(lib.cmf_microbench(0, k, 200), 2)
print(json.dumps(out))
