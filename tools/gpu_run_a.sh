#!/bin/bash
# GPU session A (1 GPU): parity suite, then the SpMV-kernel / L2-residency A/B matrix on the default workload.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r02a_gpu.txt 2>&1
python -c "import torch;p=torch.cuda.get_device_properties(0);print(p)" >> gpurun_out/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
CFEM_SPMV=stream CFEM_L2PERSIST=0 $B --no-parity > gpurun_out/r02a_stream_nopersist.json 2> gpurun_out/r02a_stream_nopersist.err
CFEM_SPMV=stream $B --no-parity > gpurun_out/r02a_stream_persist.json 2> gpurun_out/r02a_stream_persist.err
CFEM_L2PERSIST=0 $B --no-parity > gpurun_out/r02a_t16_nopersist.json 2> gpurun_out/r02a_t16_nopersist.err
$B > gpurun_out/r02a_t16_persist.json 2> gpurun_out/r02a_t16_persist.err
CFEM_LIB=$PWD/conservation-fem_b200/cfem_b200/libcfem_b200_minb8.so $B --no-parity > gpurun_out/r02a_t16_persist_minb8.json 2> gpurun_out/r02a_t16_persist_minb8.err
CFEM_L2_SETASIDE_MB=40 $B --no-parity > gpurun_out/r02a_t16_persist40.json 2> gpurun_out/r02a_t16_persist40.err
CFEM_SPMV=stream CFEM_L2PERSIST=0 $B --no-parity --workload kpp --steps 10 > gpurun_out/r02a_kpp_stream_nopersist.json 2> gpurun_out/r02a_kpp_stream_nopersist.err
$B --no-parity --workload kpp --steps 10 > gpurun_out/r02a_kpp_t16_persist.json 2> gpurun_out/r02a_kpp_t16_persist.err
tail -3 gpurun_out/r02a_pytest.log
for f in gpurun_out/r02a_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print("ms/step %.3f  e2e %.3f  its n/k/m %s/%s/%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["newton_its_per_step"], d["config"]["krylov_its_per_step"], d["config"]["mass_pcg_its_per_step"]))
    print("  breakdown", {k: round(v,3) for k,v in r["breakdown_ms_per_step"].items()})
    print("  per launch", {k:(round(v["avg_launch_ms"]*1e3,2), round(v["frac"],3)) for k,v in r["per_kernel"].items()}, "parity", d.get("parity_rel_l2"))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
