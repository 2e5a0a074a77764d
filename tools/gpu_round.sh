#!/bin/bash
# One GPU-box session: parity tests, probe, bench (both arms), ncu launch list + full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
PROBE_TC5=0 timeout 300 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
tail -3 gpurun_out/probe.log | cut -c1-1200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-900 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-400 gpurun_out/bench_ref.json
PROBE_RUNS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py > gpurun_out/ncu_launch.log 2>&1
PROBE_RUNS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loo_screen5_kernel|loo_kernel|score_tiled_kernel|gram_kernel|repack_pair_kernel|eigen_ql_kernel' -s 6 -c 6 -o gpurun_out/prof_r01 -f python tools/profile_target.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
