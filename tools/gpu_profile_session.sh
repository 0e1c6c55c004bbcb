#!/bin/bash
# Profile session for profiles/ (1 GPU): plain run first, then the ncu launch list, then one --set full capture of the
# dominant kernels of a step, plus the Euler line, the convergence study print-out and the default bench line.
mkdir -p gpurun_out
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
timeout 300 $P > gpurun_out/r02p_plain.json 2> gpurun_out/r02p_plain.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02p_launches.csv $P > gpurun_out/r02p_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k "regex:k_tile_t16|k_bicg_persist|k_tile_assemble|k_epsilon_t16|k_stats" --launch-skip 135 --launch-count 36 -f -o gpurun_out/r02p_full $P > gpurun_out/r02p_ncu2.log 2>&1; echo "full capture rc=$?"
timeout 600 python -m pytest tests/test_gpu_convergence.py -q -s > gpurun_out/r02p_convergence.log 2>&1; grep "L2 errors" gpurun_out/r02p_convergence.log; tail -1 gpurun_out/r02p_convergence.log
timeout 600 python bench.py --workload euler --steps 10 --warmup 3 > gpurun_out/r02p_euler.json 2> gpurun_out/r02p_euler.err; echo "euler rc=$?"; cut -c1-600 gpurun_out/r02p_euler.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_default.json 2> gpurun_out/r02p_default.err; echo "default rc=$?"; cut -c1-300 gpurun_out/r02p_default.json
ls -la gpurun_out/r02p_*
