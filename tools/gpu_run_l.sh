#!/bin/bash
# r02 session L: what the distributed (GHOST) variant of the Chebyshev T16 kernel costs on one GPU -- ncu of both variants
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
timeout 300 $P > gpurun_out/r02l_plain.json 2> gpurun_out/r02l_plain.err; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k "regex:k_tile_t16" --launch-skip 200 --launch-count 4 -f -o gpurun_out/r02l_t16_plain $P > gpurun_out/r02l_ncu1.log 2>&1; echo "ncu plain rc=$?"
CFEM_FORCE_GHOST=1 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k "regex:k_tile_t16" --launch-skip 200 --launch-count 4 -f -o gpurun_out/r02l_t16_ghost $P > gpurun_out/r02l_ncu2.log 2>&1; echo "ncu ghost rc=$?"
ls -la gpurun_out/r02l_*
