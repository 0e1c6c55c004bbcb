"""Single-GPU feasibility run at BASELINE.json configs[4] size: KPP on a 32M-cell mesh (4000 x 4000 x 2)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "conservation-fem_b200"))
import numpy as np
from cfem_b200 import Context, meshes, solvers as GS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
t0 = time.time(); x, c = meshes.jittered(n, n, (-2, -2), (2, 2)); t1 = time.time()
ctx = Context((x, c)); t2 = time.time()
h = ctx.nodal_h(); t3 = time.time()
uh, st = GS.solve_kpp(ctx, dt=0.64 * 4 / n, num_steps=3, h=h, return_stats=True); t4 = time.time()
u = uh.x.array
print(f"n={n} cells={c.shape[0]} nodes={x.shape[0]} mesh {t1-t0:.1f}s context {t2-t1:.1f}s nodal_h {t3-t2:.1f}s (its {ctx.nodal_h_iterations}) "
      f"3 steps {t4-t3:.1f}s device {st['device_ms']:.0f} ms newton {st['newton_iterations']} krylov {st['krylov_iterations']} "
      f"mass {st['mass_iterations']} dev mem {ctx.device_bytes/2**30:.1f} GiB min {u.min():.3f} max {u.max():.3f} finite {np.isfinite(u).all()}")
print("DoF-updates/s", x.shape[0] * 3 / (st["device_ms"] * 1e-3))
