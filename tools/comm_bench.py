"""torchrun diagnostic: latency of the communication primitives at the bench's per-GPU size.
usage: python -m torch.distributed.run --nproc-per-node N ... tools/comm_bench.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "conservation-fem_b200"))
import numpy as np, torch, torch.distributed as dist
from cfem_b200 import Context, meshes, distributed as D, _lib as L
rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank)); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
a, b = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
x, c = meshes.rectangle(a * n, b * n, (0.0, 0.0), (float(a), float(b)))
ctx = Context((x, c), device=local, comm=D.make_comm(dist))
u = np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
ctx.state_set(uh=u, u_n=u, u_old=u, u_oo=u, RH=0 * u, h=ctx.nodal_h(), t=0.0)
out = {}
for name, k in (("allreduce3", 6), ("halo", 7), ("spmv+push", 0)):
    dist.barrier(); torch.cuda.synchronize()
    ms, _ = ctx.time_kernel(k, "burgers", reps=500)
    out[name] = round(1e3 * ms, 2)
print(f"[r{rank}] us per op: {out}  env: ALLREDUCE={os.environ.get('CFEM_ALLREDUCE')} PUSH={os.environ.get('CFEM_PUSH')} ghosts={ctx.n_ghosts}", flush=True)
dist.destroy_process_group()
