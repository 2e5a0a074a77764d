"""Issue interval of FFMA and of the packed FFMA2 per scheduler (tools build, GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srcfinder_b200 import _lib
lib = _lib.load_tools()
names = {60: "FFMA, 1 warp/scheduler", 61: "FFMA, 4 warps/scheduler", 62: "FFMA2, 1 warp/scheduler", 63: "FFMA2, 4 warps/scheduler"}
print(json.dumps({names[k]: round(lib.cmf_microbench(0, k, 1), 3) for k in names}))
