"""Target for ncu: two full-size runs of the pipeline on a device-resident synthetic flightline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcfinder_b200 import ColumnwiseMF, synth
L = int(os.environ.get("PROBE_L", "20000")); S = 598; active = [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    for _ in range(int(os.environ.get("PROBE_RUNS", "2"))):
        eng.run()
print("done")
