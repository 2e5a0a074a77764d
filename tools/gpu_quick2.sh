#!/bin/bash
# parity tests + per-kernel timing probe (no micro-benchmarks)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
PROBE_TC5=0 PROBE_MICRO=0 timeout 300 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
tail -4 gpurun_out/probe.log | cut -c1-1500
if [ -n "$1" ]; then timeout 900 bash -c "$1" > gpurun_out/extra.log 2>&1; tail -30 gpurun_out/extra.log; fi
