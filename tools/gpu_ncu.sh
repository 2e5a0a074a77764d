#!/bin/bash
# ncu --set full capture of selected kernels of one pipeline run: tools/gpu_ncu.sh <regex> <skip> <count> <name>
mkdir -p gpurun_out
PROBE_RUNS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-0} -c ${3:-2} -o gpurun_out/${4:-prof} -f python tools/profile_target.py > gpurun_out/ncu_${4:-prof}.log 2>&1
tail -3 gpurun_out/ncu_${4:-prof}.log
