#!/bin/bash
# Round-end measurement session on one GPU: parity suite, bench (both arms), other configs, launch list, --set full capture.
# The .ncu-rep stays on the box (too large for gpurun_out/); its raw page comes back as CSV.
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-600 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_ref.err
cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 300 python tools/time_configs.py > gpurun_out/${TAG}_time_configs.log 2>&1; cp gpurun_out/time_configs.json gpurun_out/${TAG}_time_configs.json
timeout 300 python tools/time_co2.py > gpurun_out/${TAG}_time_co2.log 2>&1; tail -1 gpurun_out/${TAG}_time_co2.log | cut -c1-400
timeout 300 python tools/s5_timeline.py > gpurun_out/${TAG}_s5_timeline.txt 2>&1; cp gpurun_out/s5_timeline.json gpurun_out/${TAG}_s5_timeline.json
PROBE_RUNS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv python tools/profile_target.py > gpurun_out/ncu_launch.log 2>&1
PROBE_RUNS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loo_screen5_kernel|loo_kernel|score_tiled_kernel|gram_kernel|repack_pair_kernel|eigen_ql_kernel' -s 6 -c 6 -o /tmp/${TAG}_full -f python tools/profile_target.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>/dev/null
ls -la gpurun_out | tail -15
