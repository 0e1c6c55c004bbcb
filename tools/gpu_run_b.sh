#!/bin/bash
# GPU session B (1 GPU): device L2 limits, parity suite with the left-preconditioned solvers, A/B of the two SpMV
# families, ncu --set full captures of the SpMV-type kernels in both families and of the vector / assembly kernels.
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'conservation-fem_b200')
from cfem_b200 import _lib as L; print(L.device_limits(0))" > gpurun_out/r02b_limits.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity"
$B > gpurun_out/r02b_t16.json 2> gpurun_out/r02b_t16.err
CFEM_SPMV=stream $B > gpurun_out/r02b_stream.json 2> gpurun_out/r02b_stream.err
CFEM_SPMV=stream CFEM_L2PERSIST=0 $B > gpurun_out/r02b_stream_nopersist.json 2> gpurun_out/r02b_stream_nopersist.err
N="ncu --set full --clock-control none --cache-control none --import-source on"
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
timeout 600 $N -k regex:k_tile_t16 --launch-skip 200 --launch-count 5 -f -o gpurun_out/r02b_t16 $P > gpurun_out/r02b_ncu_t16.log 2>&1
CFEM_SPMV=stream timeout 600 $N -k "regex:k_cheb_stream|k_spmv_stream" --launch-skip 200 --launch-count 5 -f -o gpurun_out/r02b_stream $P > gpurun_out/r02b_ncu_stream.log 2>&1
timeout 600 $N -k "regex:k_bl_|k_tile_assemble|k_epsilon|k_stats" --launch-skip 60 --launch-count 14 -f -o gpurun_out/r02b_other $P > gpurun_out/r02b_ncu_other.log 2>&1
cat gpurun_out/r02b_limits.txt; tail -5 gpurun_out/r02b_pytest.log
for f in gpurun_out/r02b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print("ms/step %.3f  e2e %.3f  its n/k/m %s/%s/%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["newton_its_per_step"], d["config"]["krylov_its_per_step"], d["config"]["mass_pcg_its_per_step"]))
    print("  breakdown", {k: round(v,3) for k,v in r["breakdown_ms_per_step"].items()})
    print("  per launch", {k:(round(v["avg_launch_ms"]*1e3,2), round(v["frac"],3)) for k,v in r["per_kernel"].items()})
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
ls -la gpurun_out/*.ncu-rep
