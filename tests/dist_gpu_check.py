"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/dist_gpu_check.py

Runs Burgers, KPP (unstructured) and advection for a few steps on a mesh partitioned over
the ranks and compares the merged field with the CPU oracle (<= 1e-10 relative L2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "conservation-fem_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cfem_b200 import Context, meshes, step_params, distributed as D, solvers as GS  # noqa: E402
from oracle import solvers as S  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = D.make_comm(dist)
    world = comm[1]
    ok = True

    def report(name, got, ref, tol=1e-10):
        nonlocal ok
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        if rank == 0:
            print(f"[dist x{world}] {name}: rel L2 vs oracle {err:.2e}", flush=True)
        ok = ok and err < tol

    # Burgers, structured
    x, c = meshes.rectangle(48, 40)
    ctx = Context((x, c), device=local, comm=comm)
    assert ctx.n_owned + 0 <= ctx.n and (world == 1 or ctx.n_ghosts > 0)
    dt, n = 0.5 / 48, 10
    uh, st = GS.solve_burgers(ctx, dt=dt, num_steps=n, return_stats=True)
    full = D.allgather_field(ctx, uh.x.array, dist)
    ref, _, _ = S.run_burgers(x, c, dt, n)
    report("burgers 48x40, 10 steps", full, ref.uh)
    assert st["newton_iterations"] == sum(ref.newton_its)
    cs = ctx.comm_stats()
    if rank == 0:
        print(f"[dist x{world}] comm: {cs}", flush=True)
    # owned-layout state calls: what comes back is the owned slice of the global field, and a state sent
    # that way (ghosts refreshed by the halo exchange inside) steps exactly like the global-array one
    own = ctx.owned_dofs()
    g = ctx.state_get(("uh", "u_n", "u_old", "u_oo", "RH"))
    o = ctx.state_get_owned(("uh", "u_n", "u_old", "u_oo", "RH"))
    for k in ("uh", "u_n", "u_old", "u_oo", "RH"):
        assert np.array_equal(o[k], g[k][own]), k
    p = step_params("burgers", dt, 0.5, 10.0, scheme="bdf2", bc_kind="burgers_exact")
    ctx.step_scalar(p, 1)
    a = ctx.state_get_owned(("uh",))["uh"]
    ctx.state_update_owned(uh=o["uh"], u_n=o["u_n"], u_old=o["u_old"], u_oo=o["u_oo"], RH=o["RH"], t=o["t"])
    ctx.step_scalar(p, 1)
    b = ctx.state_get_owned(("uh",))["uh"]
    assert np.linalg.norm(a - b) <= 1e-12 * np.linalg.norm(a), np.abs(a - b).max()   # iteration predictions differ
    ctx.close()

    # KPP, unstructured + permuted numbering
    x, c = meshes.jittered(40, 40, (-2, -2), (2, 2))
    ctx = Context((x, c), device=local, comm=comm)
    dt, n = 0.64 * 4 / 40, 8
    uh = GS.solve_kpp(ctx, dt=dt, num_steps=n)
    full = D.allgather_field(ctx, uh.x.array, dist)
    ref, _, _ = S.run_kpp(x, c, dt, n)
    report("kpp jittered 40x40, 8 steps", full, ref.uh)
    ctx.close()

    # linear advection
    x, c = meshes.rectangle(40, 36)
    ctx = Context((x, c), device=local, comm=comm)
    dt = S.advection_dt(S.advection_velocity(x), 1 / 40)
    uh = GS.solve_advection(ctx, dt=dt, num_steps=10)
    full = D.allgather_field(ctx, uh.x.array, dist)
    ref_uh, _, _, _ = S.run_advection(x, c, dt, 10)
    report("advection 40x36, 10 steps", full, ref_uh)
    ctx.close()

    # Euler system (4 components, halo width 4)
    from oracle import euler as E

    x, c = meshes.rectangle(40, 20, (0, 0), (2, 1))
    ctx = Context((x, c), device=local, comm=comm)
    dt = 0.2 * 2 / 40
    U = GS.solve_euler(ctx, dt=dt, num_steps=6)
    full = np.stack([D.allgather_field(ctx, np.ascontiguousarray(U[:, k]), dist) for k in range(4)], axis=1)
    ref, _ = E.run_euler(x, c, dt, 6)
    report("euler 40x20 sod, 6 steps", full, ref["Uh"])
    ctx.close()

    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_OK" if t.item() == 1 else "DIST_CHECK_FAILED", flush=True)
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
