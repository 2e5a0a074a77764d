#!/bin/bash
# tcgen05 bring-up: self tests, then parity tests and the per-kernel probe with the tcgen05 screening kernel.
mkdir -p gpurun_out
timeout 600 python tools/tc5_selftest.py > gpurun_out/tc5_selftest.log 2>&1
cat gpurun_out/tc5_selftest.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
PROBE_TC5=0 PROBE_MICRO=0 timeout 300 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
tail -6 gpurun_out/probe.log
