"""Device-resident per-kernel times of the CO2 window (bands 309..391, D = 83) at flightline size (GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = 20000, 598, [309, 391]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run(); eng.run()
    for _ in range(3):
        eng.run(timing=True)
    kt = eng.kernel_times()
    out = {"window": active, "screen_kernel": eng.screen_kernel(), "ms": sum(kt.values()),
           "mpixel_s": L * S / sum(kt.values()) / 1e3, "kernel_ms": kt, "status_nonzero": int((eng.status() != 0).sum())}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_co2.json", "w"), indent=1)
