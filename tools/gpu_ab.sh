#!/bin/bash
# Generic GPU A/B session: tools/gpu_ab.sh TAG [--pytest] "name|ENV=.. ENV=..|bench args" ...
# Each entry runs `bench.py --no-cpu-baseline <args>` with the given environment; JSON lines land in
# gpurun_out/TAG_<name>.json and a one-screen summary is printed at the end.
tag=$1; shift
mkdir -p gpurun_out
if [ "$1" = "--pytest" ]; then
  shift
  timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -4 gpurun_out/${tag}_pytest.log
fi
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; args=${rest#*|}
  env $envs timeout 600 python bench.py --no-cpu-baseline $args > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
done
for spec in "$@"; do
  name=${spec%%|*}
  echo "== $name"
  python - "gpurun_out/${tag}_${name}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print("ms/step %.3f  e2e %.3f  its n/k/m %s/%s/%s launches/step %.0f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["newton_its_per_step"], d["config"]["krylov_its_per_step"], d["config"]["mass_pcg_its_per_step"], d["gpu_launches"]/d["steps"]))
    print("  breakdown", {k: round(v,3) for k,v in r["breakdown_ms_per_step"].items()})
    print("  per launch", {k:(round(v["avg_launch_ms"]*1e3,2), round(v["frac"],3)) for k,v in r["per_kernel"].items()}, "parity", d.get("parity_rel_l2"))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
