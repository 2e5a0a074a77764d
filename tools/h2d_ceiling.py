"""Plain host-to-device copy ceiling of the box at N ranks (torchrun): every rank copies the active window of its own
pinned flightline to its GPU, no compute.  Two host layouts: the full (L, 425, S) cube (strided: one 172 KB run per
line, what cmf_run_host reads) and a contiguous (L, 72, S) window.  Prints one JSON line from rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import bench
if world > 1:
    try:
        bench.bind_to_gpu_numa_node(torch.cuda.get_device_properties(dev).uuid, local)
    except Exception:
        pass
L, B, S, D, lo = 20000, 425, 598, 72, 350
out = {}
dst = torch.empty((L, D, S), dtype=torch.float32, device=dev)
for name in ("window", "full"):
    try:
        host = torch.zeros((L, D, S) if name == "window" else (L, B, S), dtype=torch.float32, pin_memory=True)
    except Exception as exc:
        out[name] = {"error": str(exc)[:100]}
        continue
    src = host if name == "window" else host[:, lo:lo + D, :]
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = 6
    t0 = time.perf_counter()
    for _ in range(n):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = L * D * S * 4 * n / dt / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        lo_t = t.clone(); dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out[name] = {"aggregate_gbs": float(t.item()), "slowest_rank_gbs": float(lo_t.item())}
    else:
        out[name] = {"aggregate_gbs": gbs, "slowest_rank_gbs": gbs}
    del host, src
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d": out}))
if world > 1:
    dist.destroy_process_group()
