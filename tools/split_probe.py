"""Does one flightline run faster as 2-4 column ranges on separate streams (latency-bound kernels of one range under
the throughput-bound kernels of another)?  Device-resident 425-band cube, GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcfinder_b200 import ColumnwiseMF, synth
from srcfinder_b200.dist import column_shard_even

L, S, B, active = int(os.environ.get("BP_L", 20000)), 598, 425, [351, 422]
ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
cube = torch.zeros((L, B, S), dtype=torch.float32, device="cuda")
cube[:, active[0] - 1:active[1], :] = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=2)
base = cube[:, active[0] - 1:, :].data_ptr()
torch.cuda.synchronize()
out = {}
for nsplit in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    engs = []
    for r, st in enumerate(streams):
        s0, s1 = column_shard_even(S, nsplit, r)
        e = ColumnwiseMF(L, B, s1 - s0, active, ab, stream=st.cuda_stream)
        e.bind_device(base + 4 * s0, line_pitch=B * S, band_pitch=S)
        engs.append(e)
    for _ in range(3):
        for e in engs:
            e.run(sync=False)
    torch.cuda.synchronize()
    n = 10
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e0.record()
    for st in streams:
        st.wait_event(e0)
    for _ in range(n):
        for e in engs:
            e.run(sync=False)
    for st, ev in zip(streams, ends):
        ev.record(st)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in ends) / n
    out["ranges_%d" % nsplit] = {"ms_per_flightline": ms, "mpixel_s": L * S / ms / 1e3}
    print(nsplit, out["ranges_%d" % nsplit], flush=True)
    for e in engs:
        e.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/split_probe.json", "w"), indent=1)
