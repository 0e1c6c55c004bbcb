"""Driver hooks: build() compiles every native piece; smoke() runs one tiny hot-path call on cuda:0."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "conservation-fem_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def build() -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (csrc/Makefile) -> libcfem_b200.so, in-tree."""
    subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-j8"])
    import cfem_b200  # noqa: F401
    from cfem_b200 import _lib

    _lib.load()  # binds every symbol declared in include/cfem_b200.h


def smoke() -> None:
    """Three Burgers RV steps on a 32x32 mesh on cuda:0, checked against the CPU oracle."""
    import numpy as np

    from cfem_b200 import meshes, solvers as GS
    from oracle import solvers as S

    x, c = meshes.rectangle(32, 32)
    dt, n = 0.5 / 32, 3
    uh, stats = GS.solve_burgers((x, c), dt=dt, num_steps=n, return_stats=True)
    st, _, _ = S.run_burgers(x, c, dt, n)
    err = np.linalg.norm(uh.x.array - st.uh) / np.linalg.norm(st.uh)
    assert err < 1e-10, f"smoke parity failed: rel L2 {err:.3e}"
    assert stats["kernel_launches"] > 0
    print(f"smoke ok: rel L2 vs oracle {err:.2e}, kernels launched {stats['kernel_launches']}")


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
