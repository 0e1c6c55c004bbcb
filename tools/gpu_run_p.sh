#!/bin/bash
# r02 session P (one shot): staged assembly -- GPU suite, then A/B against the committed build
D=/root/repo/conservation-fem_b200/cfem_b200
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/r02p2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p2_pytest.log
timeout 45 python bench.py --no-cpu-baseline --no-parity --steps 20 --warmup 3 > gpurun_out/r02p2_new.json 2> gpurun_out/r02p2_new.err; echo "new rc=$?"
CFEM_LIB=$D/libcfem_b200_head.so timeout 45 python bench.py --no-cpu-baseline --no-parity --steps 20 --warmup 3 > gpurun_out/r02p2_head.json 2> gpurun_out/r02p2_head.err; echo "head rc=$?"
python - <<'PY'
import json
for n in ("new","head"):
    try:
        d=json.loads(open(f"gpurun_out/r02p2_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["roofline"]["breakdown_ms_per_step"].items()})
    except Exception as e:
        print(n, "FAILED", e)
PY
