"""Where the cycles of one 128-pixel tile of loo_screen5_kernel go: SM clocks at every hand-off of four consecutive
tiles of one CTA (tools build, GPU box).  Prints the events relative to GEMM1-done of the first tile."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["CMF_B200_LIB"] = os.path.join(ROOT, "srcfinder_b200", "libcmf_b200_tools.so")
from srcfinder_b200 import _lib
import numpy as np
import torch
from srcfinder_b200 import ColumnwiseMF, synth
lib = _lib.load_tools()
L, S, active = 20000, 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run(); eng.run(); eng.sync()
    tl = np.zeros((4, 32), dtype=np.int64)
    lib.cmf_tools_s5_timeline.argtypes = [C.c_void_p]
    rc = lib.cmf_tools_s5_timeline(C.c_void_p(tl.ctypes.data))
names = {0: "sq: wait G1", 1: "sq: G1 seen", 2: "sq: done (ZREADY)", 3: "cv: start", 4: "cv: x tile landed",
         5: "cv: done (XREADY)", 8: "mma: wait XREADY", 9: "mma: XREADY seen", 10: "mma: GEMM1 issued",
         11: "mma: ZREADY seen", 12: "mma: REMPTY0 seen", 13: "mma: GEMM2a issued", 14: "mma: REMPTY1 seen",
         15: "mma: GEMM2b issued", 16: "epi0: wait RFULL0", 17: "epi0: RFULL0 seen", 18: "epi0: done",
         20: "epi1: wait RFULL1", 21: "epi1: RFULL1 seen", 22: "epi1: done"}
t0 = int(tl[0, 1])
rows = []
for t in range(4):
    for ev in sorted(names):
        if tl[t, ev]:
            rows.append((int(tl[t, ev]) - t0, "tile %d  %s" % (6 + t, names[ev])))
rows.sort()
for c, n in rows:
    print("%8d  %s" % (c, n))
out = {"rc": rc, "period_cycles": [int(tl[t + 1, 1] - tl[t, 1]) for t in range(3)],
       "events": {"%d:%d" % (6 + t, ev): int(tl[t, ev]) - t0 for t in range(4) for ev in names if tl[t, ev]}}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/s5_timeline.json", "w"), indent=1)
cta = {k: int(tl[3, e]) for e, k in ((24, "entry"), (25, "setup done"), (26, "tables landed"), (28, "epilogue tiles done"),
                                       (29, "reduction done"), (30, "all warps done"), (31, "exit")) if tl[3, e]}
e0 = cta.get("entry", 0)
out["cta"] = {k: v - e0 for k, v in cta.items()}
out["cta"]["first G1 (tile 6) seen"] = t0 - e0
json.dump(out, open("gpurun_out/s5_timeline.json", "w"), indent=1)
print(json.dumps(out["cta"]))
print(json.dumps(out["period_cycles"]))
