"""How much do two (or three) contexts on separate streams gain over one (GPU box)?  Same flightline, device resident."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srcfinder_b200 import ColumnwiseMF, synth

L, S, active = int(os.environ.get("BP_L", 20000)), 598, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
full = os.environ.get("BP_FULL", "1") == "1"
if full:
    cube = torch.zeros((L, 425, S), dtype=torch.float32, device="cuda")
    synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2, out=None)
    cube[:, active[0] - 1:active[1], :] = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
    ptr = cube[:, active[0] - 1:, :].data_ptr()
    pitch = (425 * S, S)
else:
    slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
    ptr, pitch = slab.data_ptr(), (None, None)
torch.cuda.synchronize()
out = {}
for nctx in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(nctx)]
    engs = [ColumnwiseMF(L, 425, S, active, ab, stream=s.cuda_stream) for s in streams]
    for e in engs:
        e.bind_device(ptr, *pitch)
    for i in range(2 * nctx):
        engs[i % nctx].run(sync=False)
    torch.cuda.synchronize()
    n = 12
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(nctx)]
    e0.record(streams[0])
    for i in range(n):
        engs[i % nctx].run(sync=False)
    for s, e in zip(streams, ends):
        e.record(s)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e) for e in ends) / n
    out["contexts_%d" % nctx] = {"ms_per_flightline": ms, "mpixel_s": L * S / ms / 1e3}
    print(nctx, out["contexts_%d" % nctx], flush=True)
    if nctx == 1:
        engs[0].run(timing=True); out["kernel_ms"] = engs[0].kernel_times(); print(out["kernel_ms"])
    for e in engs:
        e.close()
out["full_cube_bound"] = full
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/batch_probe.json", "w"), indent=1)
