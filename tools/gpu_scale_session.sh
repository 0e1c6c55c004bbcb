#!/bin/bash
# Scaling session at N GPUs (one box): tools/gpu_scale_session.sh N TAG [what...]
#   burgers   default bench line at N GPUs (weak scaling, 1024^2 cells per GPU), with in-run parity and wait accounting
#   metis     the same in sweep mode with the METIS partition
#   kstrong   KPP strong scaling: 4000^2 x 2 = 32 M cells over N GPUs
#   kweak     KPP weak scaling: 2828^2 x 2 = 16 M cells per GPU (128 M cells at 8)
#   dist      tests/dist_gpu_check.py (oracle parity of all four loops on N ranks)
N=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
run() { name=$1; shift
  if [ "$N" = 1 ]; then timeout 420 python bench.py "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  else timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err; fi
  python - "gpurun_out/${TAG}_${name}" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f+".json").read().strip().splitlines()[-1]); c=d["config"]
    its=(c.get("krylov_its_per_step"), c.get("mass_its_per_step") or c.get("mass_pcg_its_per_step"))
    print(f"{f.split('/')[-1]:28s} N={d['n_gpus']} {d['config']['workload'][:40]} cells {c['cells']} ms/step {d['ms_per_step']:.3f} value {d['value']/1e6:.1f}M its {its} setup {c.get('context_setup_s_rank0')} mesh {c.get('mesh_generation_s')}")
    print("    parity", d.get("parity_rel_l2"), "\n    wait", c.get("comm_wait"), "\n    solves", c.get("solve_timing"))
except Exception as e:
    print(f, "FAILED", e); print(open(f+".err").read()[-1200:])
PY
}
for what in "$@"; do
  case $what in
    burgers) run burgers --steps 20 --warmup 3 --no-cpu-baseline ;;
    metis)   run burgers_metis --steps 20 --warmup 3 --sweep --no-parity --partition metis ;;
    kstrong) run kpp_strong --workload kpp --strong --size 4000 --steps 10 --warmup 3 --sweep --no-parity ;;
    kstrong_metis) run kpp_strong_metis --workload kpp --strong --size 4000 --steps 10 --warmup 3 --sweep --no-parity --partition metis ;;
    kweak)   run kpp_weak --workload kpp --size 2828 --steps 10 --warmup 3 --sweep --no-parity ;;
    dist)    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_gpu_check.py > gpurun_out/${TAG}_dist.log 2>&1; grep "dist x\|DIST_\|Error" gpurun_out/${TAG}_dist.log | grep -v "comm:" | sort | uniq | head -8 ;;
  esac
done
