#!/bin/bash
# 2-GPU diagnosis of the Chebyshev chain: GHOST kernel variants on one GPU (no peers), LL vs flag halo on two
mkdir -p gpurun_out
sw() { tag=$1; shift; envs=$1; shift; n=$1; shift
  if [ "$n" = 1 ]; then env $envs python bench.py --sweep --no-parity "$@" > gpurun_out/r02h_$tag.json 2> gpurun_out/r02h_$tag.err
  else env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --sweep --no-parity "$@" > gpurun_out/r02h_$tag.json 2> gpurun_out/r02h_$tag.err; fi
  python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r02h_{t}.json").read().strip().splitlines()[-1]); c=d["config"]
    print(f"{t:22s} N={d['n_gpus']} ms/step {d['ms_per_step']:.3f}  its k/m {c['krylov_its_per_step']:.1f}/{c['mass_its_per_step']:.1f} setup {c['context_setup_s_rank0']:.1f}s wait {c['comm_wait']}")
except Exception as e:
    print(t, "FAILED", e); print(open(f"gpurun_out/r02h_{t}.err").read()[-800:])
PY
}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; tail -3 gpurun_out/r02h_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_gpu_check.py > gpurun_out/r02h_dist.log 2>&1; grep "DIST_" gpurun_out/r02h_dist.log
sw n1 "A=1" 1 --steps 20 --warmup 3
sw n1_ghost "CFEM_FORCE_GHOST=1" 1 --steps 20 --warmup 3
sw n2 "A=1" 2 --steps 20 --warmup 3
sw n2_metis "A=1" 2 --steps 20 --warmup 3 --partition metis
sw n2_kpp "A=1" 2 --steps 10 --warmup 3 --workload kpp
