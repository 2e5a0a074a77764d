#!/usr/bin/env python
"""Benchmark of the columnwise matched filter (BASELINE.json metric: CMF Mpixel/s on a 425-channel
AVIRIS-NG cube).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's NumPy path on host cores

A "step" is one pass of the whole hot path (repack/mask, statistics, eigen-decomposition, tensor-core screen
of the LOO alpha search + exact FP64 refinement, weights, scoring, column statistics) over one synthetic flightline: BASELINE configs[1], 598 columns x 425 channels
x 20 000 lines, active window 351..422, the whole 425-band cube resident in HBM (the kernels read the active
window in place, line pitch 425 x 598).  With N > 1 every rank processes its own flightline (flightline
sharding, no data-path collective) and the score tiles are gathered over NCCL inside the timed region, the
root rotating from flightline to flightline so that no single GPU receives every tile.  One JSON line is
printed by rank 0.

`value`  device-resident throughput (inputs already in HBM), CUDA events on the launching stream,
         max over ranks.
`e2e`    the same metric through the C-ABI host call (cmf_run_host): pinned host cube -> H2D of the
         active window -> all kernels -> D2H of scores, column statistics and alpha indices, every step.
`roofline`      the dominant kernel (screening pass of the alpha search, TF32 tensor bound) against the dense
                TF32 peak (half the measured bf16 figure of MEASURED_PEAKS.json); `roofline_fp64` the Gram
                kernel against the FP64 DMMA rate measured in this run (MEASURED_PEAKS.json has no FP64 figure).
`roofline_hbm`  the scoring pass against the HBM copy peak of MEASURED_PEAKS.json.
`cpu_baseline`  the oracle port of the reference (same NumPy/SciPy calls) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

METRIC = "cmf_mpixel_per_s"
UNIT = "Mpixel/s"
ACTIVE = [351, 422]
BANDS = 425
FALLBACK_HBM_GBS = 6650.0        # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


# ------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": FALLBACK_HBM_GBS}, "fallback"


class ClockSampler(object):
    """SM clock and throttle reasons sampled through NVML every 10 ms while the timed region runs (an in-process
    thread: nvidia-smi takes longer to start than a 10-step region lasts); `nvidia-smi -lms` is the fallback."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20),
            ("hw_thermal_slowdown", 0x40))

    def __init__(self, index=0, uuid=None):
        self.index, self.uuid = index, uuid
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = self.thread = self.handle = self.nvml = None
        self.stop = threading.Event()
        self.source = None

    def _open_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid:
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(self.uuid))
            except Exception:
                h = None
        if h is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            toks = [t for t in vis.split(",") if t.strip()]
            if toks and self.index < len(toks) and toks[self.index].strip().isdigit():
                idx = int(toks[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.nvml, self.handle = pynvml, h
        self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
        self.mx.append(self.max_sm)
        try:
            bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        for name, bit in self.BITS:
            if bits & bit:
                self.reasons.add(name)

    def _loop_nvml(self):
        while not self.stop.is_set():
            try:
                self._sample_nvml()
            except Exception:
                break
            self.stop.wait(0.01)

    def _pump_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [t.strip() for t in line.split(",")]
            try:
                self.sm.append(float(r[0])); self.mx.append(float(r[1]))
            except Exception:
                continue
            for name, flag in zip(names, r[2:6]):
                if flag.lower().startswith("active"):
                    self.reasons.add(name)

    def __enter__(self):
        try:
            self._open_nvml()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._loop_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.handle = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.handle is not None:
            try:
                self._sample_nvml()          # at least one sample taken before the region closes
            except Exception:
                pass
            self.stop.set()
            if self.thread is not None:
                self.thread.join(timeout=1)
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)),
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def abscf_window():
    from srcfinder_b200 import synth
    return synth.load_ch4_library()[ACTIVE[0] - 1:ACTIVE[1], 2]


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One host core: the oracle port on `ncols` columns of a synthetic flightline (active bands only)."""
    seed, lines, ncols = args
    for v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    from oracle import cmf_oracle as orc
    from srcfinder_b200 import synth
    lib = synth.load_ch4_library()[ACTIVE[0] - 1:ACTIVE[1]]
    cube = synth.make_cube(lines, ncols, bands=lib.shape[0], seed=seed, lib=lib)
    t0 = time.perf_counter()
    res = orc.cmf_cube(cube, lib[:, 2], [1, lib.shape[0]])
    dt = time.perf_counter() - t0
    return dt, int(res["mask"].sum())


def cpu_sample(lines, cols_per_core, cores, seed0=1000):
    """Run the oracle port on `cores` processes at once; returns (Mpixel/s, wall s, pixels)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    jobs = [(seed0 + i, lines, cols_per_core) for i in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(processes=cores) as pool:
        out = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    compute = max(dt for dt, _ in out)          # the pool start-up (imports) is not part of the metric
    pixels = lines * cols_per_core * cores
    return pixels / compute / 1e6, wall, pixels, compute


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_leg(lines, budget_s=20.0):
    cores = host_cores()
    # one column of a 20 000-line flightline costs ~2-3 s on one core; size the sample to the budget
    per_col = 2.5 * lines / 20000.0
    cols = max(1, int(budget_s / max(per_col, 1e-3)))
    cols = min(cols, 8)
    val, wall, pixels, compute = cpu_sample(lines, cols, cores)
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d columns x %d lines on each of %d processes (1 BLAS thread each), %.1f s"
                      % (cols, lines, cores, compute)}


def run_reference(args):
    """--impl reference: the reference's NumPy/SciPy path (oracle port, cmf/robust_mf.py restated with the
    same LAPACK calls) on the box's host cores; each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    lines = args.lines
    cores = host_cores()
    cols = max(1, min(4, int(10.0 / max(2.5 * lines / 20000.0, 1e-3))))
    vals, secs = [], []
    for i in range(args.warmup):
        cpu_sample(lines, 1, cores, seed0=500 + 10 * i)
    for i in range(args.steps):
        v, wall, pixels, compute = cpu_sample(lines, cols, cores, seed0=2000 + 100 * i)
        vals.append(v); secs.append(compute)
    value = float(np.mean(vals)) if vals else 0.0
    sample = "%d columns x %d lines per process, %d processes, 1 BLAS thread each" % (cols, lines, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(secs) * 1e3) if secs else None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, n):
    return {"workload": "configs[1]: full AVIRIS-NG flightline %d cols x %d ch x %d lines, single-pass "
                        "unimodal looshrinkage CMF, active bands %d..%d, 201 alphas"
                        % (args.samples, BANDS, args.lines, ACTIVE[0], ACTIVE[1]),
            "lines": args.lines, "samples": args.samples, "bands": BANDS, "active_bands": ACTIVE,
            "alphas": 201, "flightlines_per_gpu": 1 if args.shard == "flightline" else "1/%d (column range)" % n,
            "sharding": ("flightline per GPU; NCCL gather of the score tiles, root = flightline index mod N, overlapped "
                         "with the next flightline (two contexts per GPU)") if args.shard == "flightline" else
                        "one flightline per step, contiguous even-aligned column ranges per GPU, NCCL gather of the tiles",
            "input": "425-band cube resident in HBM, active window read in place",
            "parallelism": "dp%d" % n, "timing": "inputs (3.4 GB/flightline) far larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from srcfinder_b200 import ColumnwiseMF, _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback "
                         "(use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:
        gpu_uuid = torch.cuda.get_device_properties(dev).uuid
    except Exception:
        gpu_uuid = None
    cpus_bound = bind_to_gpu_numa_node(gpu_uuid, local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L, S = args.lines, args.samples
    D = ACTIVE[1] - ACTIVE[0] + 1
    ab = abscf_window()
    lib = _lib.load()

    # FP64 tensor peak for the roofline (rank 0, before the timed region)
    dmma_peak = _lib.load_tools().cmf_microbench(local, 0, 3) if rank == 0 else 0.0
    mma_peak = _lib.load_tools().cmf_microbench(local, 9, 3) if rank == 0 else 0.0

    # the whole 425-band cube resident in HBM (20.3 GB); only the active window carries data, as only it is read
    cube = torch.zeros((L, BANDS, S), dtype=torch.float32, device=dev)
    slab = cube[:, ACTIVE[0] - 1:ACTIVE[1], :]
    synth.make_slab_torch(L, S, ACTIVE[0], ACTIVE[1], dev, seed=2 + rank, out=slab)
    stream = torch.cuda.current_stream()
    # N > 1: two contexts used alternately, so that the score tile of flightline i is gathered straight out of its
    # context while flightline i+1 runs in the other one (no staging copy)
    nctx = 2 if world > 1 else 1
    by_columns = args.shard == "columns" and world > 1
    s0, s1 = (0, S)
    if by_columns:
        # ONE flightline per step for the whole job: every rank filters its own column range of the same cube in
        # place (columns are independent problems, cmf/robust_mf.py:297) and the tiles are gathered
        from srcfinder_b200.dist import column_shard_even
        s0, s1 = column_shard_even(S, world, rank)
        wmax = max(column_shard_even(S, world, r)[1] - column_shard_even(S, world, r)[0] for r in range(world))
    Sg = s1 - s0
    engines = [ColumnwiseMF(L, BANDS, Sg, ACTIVE, ab, device=local, stream=stream.cuda_stream) for _ in range(nctx)]
    for e_ in engines:
        e_.bind_device(slab.data_ptr() + 4 * s0, line_pitch=BANDS * S, band_pitch=S)
    eng = engines[0]
    mf_dev = [None] * nctx
    gathered = None
    pad = None
    if world > 1:
        mf_dev = [_as_tensor(torch, e_.device_ptr(_lib.OUT_MF), (L, Sg), torch.float64, dev) for e_ in engines]
        gw = wmax if by_columns else S
        gathered = [torch.empty((L, gw), dtype=torch.float64, device=dev) for _ in range(world)]   # every rank is a root in turn
        if by_columns and Sg != wmax:
            pad = [torch.zeros((L, wmax), dtype=torch.float64, device=dev) for _ in range(nctx)]
    pending = [None] * nctx
    counter = [0]

    def step(timing):
        i = counter[0]
        counter[0] += 1
        k = i % nctx
        if pending[k] is not None:
            pending[k].wait()                 # the gather that read this context's scores two flightlines ago
            pending[k] = None
        engines[k].run(timing=timing, sync=False)
        if world > 1:
            root = i % world                  # the tiles of flightline i land on GPU i mod N
            send = mf_dev[k]
            if pad is not None:                # shards narrower than the widest one are padded (12 MB copy)
                pad[k][:, :Sg].copy_(mf_dev[k], non_blocking=True)
                send = pad[k]
            pending[k] = dist.gather(send, gathered if rank == root else None, dst=root, async_op=True)

    def drain():
        for k in range(nctx):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def barrier():
        if world > 1:
            drain()
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: at least 3 steps, and with N > 1 at least two steps per gather root, so that NCCL has set up the
    # peer connections of EVERY root before the timed region (they are made lazily on first use: measured 700 ms
    # per step at N = 8 when roots were first used inside the timed region)
    nwarm = max(args.warmup, 3, 2 * world if world > 1 else 0)
    for _ in range(nwarm):
        step(False)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, gpu_uuid) as clocks:
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step(True)
        drain()                                   # the last gathers are inside the timed region as well
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # mean per-kernel ms over the timed steps (CUDA events of every context, same stream)
    kts = [e_.kernel_times() for e_ in engines[:min(nctx, args.steps)]]
    kt = {k: float(np.mean([t[k] for t in kts])) for k in kts[0]}
    launches = eng.launch_count() * args.steps        # this library's kernels only (per rank); NCCL / torch copies not counted
    jobs = 1 if by_columns else world             # flightlines finished per step by the whole job
    value = jobs * L * S * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the host API (rank-local; every rank does the same work)
    e2e = None
    if not args.no_e2e and not by_columns:
        e2e = measure_e2e(torch, engines, slab, L, S, D, args, world, rank, dev, barrier)
    skern = eng.screen_kernel() or "loo_screen_kernel"
    wide = None
    if rank == 0 and world == 1 and not args.no_wide:
        for e_ in engines:
            e_.close()
        engines = []
        del cube, slab
        torch.cuda.empty_cache()
        wide = measure_wide(torch, L, S, dev, stream)
    out = None
    if rank == 0:
        peaks, src = measured_peaks()
        # dominant kernel: the tensor-core screening pass of the alpha search.  Algorithmic work per pixel
        # (SURVEY 8(d)): the projection 2 D^2 plus the alpha contraction 2 D A; the kernel executes three
        # TF32 products per FP64-equivalent product (hi/lo splits), which is not counted as useful work.
        flops = (2.0 * D * D + 2.0 * D * 201) * L * Sg
        screen_ms = kt.get("screen", float("nan"))
        achieved = flops / (screen_ms * 1e-3) / 1e12
        tf32_peak = 0.5 * float(peaks.get("bf16_tflops", 1590.0))
        score_bytes = (4.0 * D + 8.0 + 1.0) * L * Sg            # one read of the slab + f64 score + mask byte
        score_ms = kt.get("score", float("nan"))
        traffic = load_traffic()
        tc5 = skern == "loo_screen5_kernel"
        DP = (D + 7) // 8 * 8
        n1, na = ((DP + 15) // 16 * 16, 208) if tc5 else (DP, 208)
        executed = 3.0 * (2.0 * DP * n1 + 2.0 * DP * na) * L * Sg / (screen_ms * 1e-3) / 1e12
        pipe = ("tcgen05.mma kind::tf32 (SASS UTCMMA), accumulators in TMEM" if tc5 else
                "mma.sync (SASS HMMA), whose measured ceiling in this run is %.0f TFLOP/s" % mma_peak)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": nwarm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if by_columns else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "kernel_ms": {k: round(v, 4) for k, v in kt.items()},
            "roofline": {"kernel": skern, "bound": "tensor", "achieved": achieved, "peak": tf32_peak,
                         "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": traffic.get(skern),
                         "peak_source": "%s: dense TF32 = half of MEASURED_PEAKS.json bf16_tflops (%s, burst; sustained "
                                        "%s); the kernel issues TF32 through %s and executes 3 split products (hi*hi, "
                                        "lo*hi, hi*lo) per algorithmic product, padded to N=%d/%d"
                                        % (src, peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained"), pipe,
                                           n1, na),
                         "flops_per_launch": flops, "executed_tflops": executed,
                         "executed_frac": executed / tf32_peak, "mma_sync_tf32_peak": mma_peak,
                         "share_of_step": screen_ms / max(sum(kt.values()), 1e-9)},
            "roofline_fp64": {"kernel": "gram_kernel", "bound": "tensor", "achieved": D * (D + 8.0) * L * Sg
                              / (kt.get("gram", float("nan")) * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
                              "peak_source": "FP64 DMMA.8x8x4 rate measured in this run (cmf_microbench kind 0)",
                              "note": "lower-triangle 8x8 tiles only: D(D+8) flop per pixel"},
            "roofline_hbm": {"kernel": "score_tiled_kernel", "bound": "hbm",
                             "achieved": score_bytes / (score_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": score_bytes / (score_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "traffic": traffic.get("score_tiled_kernel"),
                             "peak_source": src + " (MEASURED_PEAKS.json hbm_gbs)",
                             "bytes_per_launch": score_bytes},
            "clocks": clocks.summary(), "gpu_launches": launches,
        }
        if e2e is not None:
            out["e2e"] = e2e
        if wide is not None:
            out["wide_window"] = wide
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu and world == 1:
            out["cpu_baseline"] = cpu_baseline_leg(L, budget_s=args.cpu_seconds)
        emit(out)
    for e_ in engines:
        e_.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _as_tensor(torch, ptr, shape, dtype, dev):
    """Zero-copy torch view of a device buffer owned by the C library."""
    n = int(np.prod(shape))
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Arr(object):
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8" if itemsize == 8 else "<f4",
                                    "data": (int(ptr), False), "version": 3, "strides": None}
    return torch.as_tensor(_Arr(), device=dev).view(shape)


def h2d_ceiling(torch, host_window, dev, world, barrier, copies=6):
    """Plain pinned-host -> device copies of one flightline's active window on every rank at once, nothing else
    running: what the box's PCIe fabric gives N GPUs (GB/s per rank, slowest rank)."""
    import torch.distributed as dist
    dst = torch.empty(host_window.shape, dtype=torch.float32, device=dev)
    dst.copy_(host_window, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(copies):
        dst.copy_(host_window, non_blocking=True)
    torch.cuda.synchronize()
    gbs = host_window.numel() * 4 * copies / (time.perf_counter() - t0) / 1e9
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        gbs = float(t.item())
    del dst
    return gbs


def measure_wide(torch, L, S, dev, stream):
    """The reference's other window: `-R` with the CH4 library selects bands 5..420 (cmf/robust_mf.py:186-187,
    D = 416).  One device-resident flightline through the wide-window kernel set, per-kernel CUDA events."""
    from srcfinder_b200 import ColumnwiseMF, synth
    try:
        act = [5, 420]
        Dw = act[1] - act[0] + 1
        ab = synth.load_ch4_library()[act[0] - 1:act[1], 2]
        slab = synth.make_slab_torch(L, S, act[0], act[1], dev, seed=3)
        torch.cuda.synchronize()
        with ColumnwiseMF(L, BANDS, S, act, ab, reflectance=True, device=dev.index, stream=stream.cuda_stream) as eng:
            eng.bind_device(slab.data_ptr())
            eng.run()
            eng.run(timing=True)
            kt = eng.kernel_times()
            launches = eng.launch_count()
            ok = int((eng.status() == 0).sum())
        ms = float(sum(kt.values()))
        lines64 = (L + 63) // 64 * 64
        nrb = (Dw + 31) // 32
        ops = 2.0 * (nrb * (nrb + 1) / 2) * 128 * 128 * lines64 * S       # executed int8 MACs x 2 of the Gram pass
        return {"workload": "-R window: %d cols x %d ch x %d lines, active bands %d..%d (D = %d), unimodal looshrinkage"
                            % (S, BANDS, L, act[0], act[1], Dw),
                "ms_per_flightline": ms, "value": L * S / ms / 1e3, "unit": UNIT,
                "kernel_ms": {k: round(v, 3) for k, v in kt.items()}, "gpu_launches": launches,
                "columns_ok": ok,
                "roofline_int8": {"kernel": "wide_gram8_kernel", "bound": "tensor", "unit": "TOP/s",
                                  "achieved": ops / (kt["gram"] * 1e-3) / 1e12, "peak": 4500.0,
                                  "frac": ops / (kt["gram"] * 1e-3) / 1e12 / 4500.0,
                                  "peak_source": "nominal dense int8 rate of B200 (4.5 POP/s; MEASURED_PEAKS.json has "
                                                 "no integer figure)",
                                  "fp64_equivalent_tflops": 2.0 * Dw * Dw * L * S / (kt["gram"] * 1e-3) / 1e12}}
    except Exception as exc:       # noqa: BLE001 -- the main line must survive
        return {"error": str(exc)[:200]}


def measure_e2e(torch, engines_in, slab, L, S, D, args, world, rank, dev, barrier):
    """cmf_run_host on a pinned host cube shaped like the reference's input file (L, 425, S) float32.

    Headline `value`: flightlines streamed through two contexts used alternately (CMF_RUN_ASYNC), i.e. the
    upload of step i+1 is already on the PCIe link while step i is being factorised and scored; every step's
    H2D copy and D2H read of its scores are inside the timed region.  `sync_call` is the same call made
    synchronously, one flightline at a time (latency of a single call)."""
    import torch.distributed as dist
    from srcfinder_b200 import ColumnwiseMF
    nbytes = L * BANDS * S * 4
    host_note = "pinned float32 BIL cube (L,425,S); only the active window is copied"
    own = []
    try:
        host = torch.zeros((L, BANDS, S), dtype=torch.float32, pin_memory=True)
        host[:, ACTIVE[0] - 1:ACTIVE[1], :].copy_(slab)         # only the bands the path reads carry data
        torch.cuda.synchronize()
        # two contexts with their OWN streams (the device-resident contexts above share torch's stream): the upload of
        # flightline i+1 must not queue behind the kernels of flightline i
        engines = [ColumnwiseMF(L, BANDS, S, ACTIVE, abscf_window(), device=dev.index) for _ in range(2)]
        own = list(engines)
    except Exception as exc:
        # not enough lockable memory for the full 425-band cube on this box: the same bytes cross PCIe from a
        # pinned cube that holds the active window only (declared as a D-band cube)
        try:
            host = torch.zeros((L, D, S), dtype=torch.float32, pin_memory=True)
            host.copy_(slab)
            torch.cuda.synchronize()
            engines = [ColumnwiseMF(L, D, S, [1, D], abscf_window(), device=dev.index) for _ in range(2)]
            own = list(engines)
            host_note = ("pinned float32 BIL cube of the active window only (L,%d,S): the full %d-byte cube could "
                         "not be pinned (%s); identical bytes cross PCIe" % (D, nbytes, str(exc)[:80]))
        except Exception as exc2:
            return {"value": None, "unit": UNIT, "error": "pinned host cube failed: %s" % (exc2,)}
    mf = [torch.empty((L, S), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    cs = [torch.empty((3, S), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    ai = [torch.empty((S,), dtype=torch.int32, pin_memory=True) for _ in range(2)]
    steps = max(1, min(args.steps, args.e2e_steps))

    def submit(i, wait):
        k = i % 2
        engines[k].run_host(host.data_ptr(), mf[k].data_ptr(), cs[k].data_ptr(), ai[k].data_ptr(), wait=wait)

    for i in range(4):
        submit(i, True)
    # ---- one synchronous call per flightline
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        submit(0, True)
    dt_sync = time.perf_counter() - t0
    # ---- two flightlines in flight
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        engines[i % 2].sync()                                   # the result of step i-2 has landed
        submit(i, False)
    engines[0].sync(); engines[1].sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, dt_sync], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_sync = float(t[0].item()), float(t[1].item())
    checksum = float(cs[0][2].sum())                            # the result really came back
    same = bool(torch.equal(mf[0], mf[1])) if steps > 1 else None
    for e_ in own:
        e_.close()
    # what the PCIe fabric of this box gives the same bytes with nothing else running (all ranks at once)
    try:
        win = torch.empty((L, D, S), dtype=torch.float32, pin_memory=True)
        ceiling = h2d_ceiling(torch, win, dev, world, barrier)
        del win
    except Exception:
        ceiling = None
    achieved_gbs = L * D * S * 4 / (dt / steps) / 1e9
    return {"value": world * L * S * steps / dt / 1e6, "unit": UNIT, "ms_per_step": dt / steps * 1e3,
            "h2d_ceiling_gbs": ceiling,
            "frac_of_h2d_ceiling": (achieved_gbs / ceiling) if ceiling else None,
            "h2d_ceiling_note": "plain pinned-host -> device copy of one flightline's active window (3.44 GB) on every rank "
                                "at once, slowest rank, GB/s per rank; measured per N on this pool: 55.6 / 55.6 / 28.8 / "
                                "23.2 at N = 1 / 2 / 4 / 8 (profiles/r02_h2d_ceiling.jsonl): pairs of GPUs share ~58 GB/s",
            "h2d_bytes_per_step": L * D * S * 4, "d2h_bytes_per_step": L * S * 8 + 3 * S * 8 + S * 4,
            "steps": steps, "mode": "two contexts used alternately, cmf_run_host(CMF_RUN_ASYNC): the next "
                                    "flightline uploads while the previous one is scored",
            "sync_call": {"value": world * L * S * steps / dt_sync / 1e6, "unit": UNIT,
                          "ms_per_step": dt_sync / steps * 1e3},
            "host_buffer": host_note, "h2d_gbs": achieved_gbs,
            "colstd_checksum": checksum, "both_contexts_identical": same}


def load_traffic():
    """dram bytes per launch from the committed ncu capture (profiles/traffic.json), if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return {}


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner under NCCL_DEBUG, worker
    processes) also write to fd 1, so everything else is pointed at stderr and the result line goes to a
    private duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def bind_to_gpu_numa_node(uuid, index):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) before the pinned host cube is
    allocated, so the flightline a rank streams over PCIe sits in the memory of the socket the GPU hangs off."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(uuid)) if uuid else None
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lines", type=int, default=20000)
    ap.add_argument("--samples", type=int, default=598)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-wide", action="store_true", help="skip the extra -R (416-band) flightline timing")
    ap.add_argument("--shard", default="flightline", choices=["flightline", "columns"],
                    help="N > 1: one flightline per rank and step (weak scaling, the contract's default) or ONE "
                         "flightline per step split into column ranges (strong scaling, SURVEY 8(e)(i))")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
