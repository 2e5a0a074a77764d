"""Per-kernel times of one flightline with the library named by CMF_B200_LIB (variant A/B measurements, GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcfinder_b200 import ColumnwiseMF, synth
L, S, active = int(os.environ.get("TV_L", 20000)), 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
torch.cuda.synchronize()
with ColumnwiseMF(L, 425, S, active, ab) as eng:
    eng.bind_device(slab.data_ptr())
    eng.run(); eng.run()
    for _ in range(5):
        eng.run(timing=True)
    kt = eng.kernel_times()
    ai = eng.alpha_index()
print(json.dumps({"lib": os.path.basename(os.environ.get("CMF_B200_LIB", "default")), "total": sum(kt.values()),
                  "screen": kt["screen"], "loo": kt["loo"], "alpha_sum": int(ai.sum())}))
