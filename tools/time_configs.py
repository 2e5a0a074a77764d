"""Device-resident wall time of the BASELINE.json configurations other than the bench line (GPU box):
C1 2000 lines, C3 20000 lines with bad pixels (unimodal and k = 3 modes with rejection, partition on the device),
C5 EMIT shape 1242 x 285 x 1280.  CUDA-event time of run() after two warm-up runs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth


def timed(eng, n=3):
    eng.run(); eng.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream()
    e0.record(s)
    for _ in range(n):
        eng.run(sync=False)
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
active = [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
st = torch.cuda.current_stream().cuda_stream
for name, L, bad, k in (("C1_2000_lines", 2000, False, 1), ("C2_20000_lines", 20000, False, 1),
                        ("C3_20000_lines_bad_pixels", 20000, True, 1), ("C3_20000_lines_bad_pixels_k3_reject", 20000, True, 3)):
    slab = synth.make_slab_torch(L, 598, active[0], active[1], "cuda", seed=3)
    if bad:
        synth.inject_bad_pixels_torch(slab, 3)
    torch.cuda.synchronize()
    with ColumnwiseMF(L, 425, 598, active, ab, stream=st) as eng:
        eng.bind_device(slab.data_ptr())
        if k > 1:
            eng.set_clustering(k, pcadim=6, reject_min=85)
        ms = timed(eng)
    out[name] = {"ms": ms, "mpixel_s": L * 598 / ms / 1e3}
    print(name, out[name], flush=True)
    del slab
# C5: EMIT shape
wl = np.linspace(381.0, 2493.0, 285)
lib = synth.resample_library(wl)
inside = np.where((wl >= 2129.0) & (wl <= 2485.0))[0]
act = [int(inside[0]) + 1, int(inside[-1]) + 1]
slab = synth.make_slab_torch(1280, 1242, act[0], act[1], "cuda", seed=5, lib=lib)
torch.cuda.synchronize()
with ColumnwiseMF(1280, 285, 1242, act, lib[act[0] - 1:act[1], 2], stream=st) as eng:
    eng.bind_device(slab.data_ptr())
    ms = timed(eng)
out["C5_emit_1242x285x1280"] = {"ms": ms, "mpixel_s": 1280 * 1242 / ms / 1e3, "active": act}
print("C5", out["C5_emit_1242x285x1280"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_configs.json", "w"), indent=1)
