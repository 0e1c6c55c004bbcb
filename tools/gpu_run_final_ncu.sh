#!/bin/bash
# final ncu evidence of the committed build: plain run, launch list, one full capture of the dominant kernel
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
timeout 120 $P > gpurun_out/r02g_plain.json 2> gpurun_out/r02g_plain.err; echo "plain rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02g_launches.csv $P > gpurun_out/r02g_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 120 ncu --set full --clock-control none --cache-control none --import-source on -k "regex:k_bicg_persist" --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02g_solver $P > gpurun_out/r02g_ncu2.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/r02g_*
