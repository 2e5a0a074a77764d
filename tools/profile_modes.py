"""ncu target: one k = 3 (-r) run on a 20000-line flightline with bad pixels (BASELINE configs[2])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = int(os.environ.get("PROBE_L", "20000")), 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=3)
synth.inject_bad_pixels_torch(slab, 3)
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.set_clustering(3, pcadim=6, reject_min=85)
    eng.run()
    print("kmeans iters max", int(eng.kmeans_iters().max()))
print("done")
