"""torchrun diagnostic: distributed SpMV (halo only) and mass solve (halo + all-reduce) vs scipy."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "conservation-fem_b200"))
import numpy as np, torch, torch.distributed as dist
from cfem_b200 import Context, meshes, distributed as D, _lib as L
from oracle import p1
rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = D.make_comm(dist)
x, c = meshes.rectangle(48, 40)
ctx = Context((x, c), device=local, comm=comm)
M = p1.mass_matrix(x, c)
v = np.random.default_rng(0).normal(size=x.shape[0])
own = ctx.owned_dofs()
for rep in range(3):
    y = ctx.spmv(L.MAT_MASS, v)
    err = np.abs(y[own] - (M @ v)[own]).max()
    print(f"[r{rank}] spmv rep{rep} err {err:.2e} owned {own.size} ghosts {ctx.n_ghosts}", flush=True)
try:
    z = ctx.solve(L.MAT_MASS, M @ v, solver="chebyshev", rtol=1e-12)
    print(f"[r{rank}] cheb its {ctx.last_iterations} relres {ctx.last_relres:.2e} err {np.abs(z[own]-v[own]).max():.2e}", flush=True)
except Exception as e:
    print(f"[r{rank}] cheb failed: {e}", flush=True)
try:
    z = ctx.solve(L.MAT_MASS, M @ v, solver="pcg", rtol=1e-12)
    print(f"[r{rank}] pcg its {ctx.last_iterations} relres {ctx.last_relres:.2e} err {np.abs(z[own]-v[own]).max():.2e}", flush=True)
except Exception as e:
    print(f"[r{rank}] pcg failed: {e}", flush=True)
dist.destroy_process_group()
