#!/bin/bash
# final validation of the round: GPU test suite, smoke, the default bench line (with cpu_baseline) on the committed build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02f_smoke.log
timeout 600 python bench.py > gpurun_out/r02f_default.json 2> gpurun_out/r02f_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02f_default.json
