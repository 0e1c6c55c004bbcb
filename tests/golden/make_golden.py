"""Regenerates the fixtures in tests/golden/ (run in the authoring container only).

1. Meshes lifted from the reference's own data files (read with cfem_b200.io.H5File, a reader for
   the HDF5 subset dolfinx writes; h5py is not available):
     Code/Linear_advection/Data/RV/RV_node.h5  -> rv_node_mesh.npz  (1,919 tris / 1,011 nodes, unit disk)
     Data/KPP_RV.h5                            -> kpp_rv_mesh.npz   (9,514 tris / 4,886 nodes, [-2,2]^2)
   plus the first <Time> stamp of RV_node.xdmf, which pins the reference's dt formula.
2. Frames of the three dolfinx time series the reference stores on that disk mesh
     Code/Linear_advection/Data/RV/RV_node.h5   (tests/eps_func.py)                 -> ref_series_eps_func.npz
     Code/Linear_advection/Data/RV/RV_cell.h5   (Code/Linear_advection/RV_cell.py)  -> ref_series_rv_cell.npz
     Code/Linear_advection/Data/SI/smoothness.h5 (smoothness_old_convergence.py loop) -> ref_series_si_old.npz
   (read through cfem_b200.io.read_xdmf; 13 of the 285 frames are kept, the generator checks all 285
   against the oracle).  These ARE dolfinx output: they pin the oracle.
3. Oracle outputs on small seeded cases (fields after N steps), so GPU parity can also be
   checked against committed vectors.  These come from oracle/ (CPU restatement): the reference
   itself cannot run in this image (no dolfinx).
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "conservation-fem_b200"))
REF = "/root/reference"


KEEP = [0, 1, 2, 3, 5, 10, 20, 50, 100, 150, 200, 250, 284]
SERIES = {"eps_func": "Code/Linear_advection/Data/RV/RV_node.h5",
          "rv_cell": "Code/Linear_advection/Data/RV/RV_cell.h5",
          "si_old": "Code/Linear_advection/Data/SI/smoothness.h5"}


def stored_series(S, x, c):
    from cfem_b200 import io
    for variant, rel in SERIES.items():
        d = io.read_xdmf(f"{REF}/{rel}".replace(".h5", ".xdmf"))
        assert np.array_equal(d["x"], x) and np.array_equal(d["cells"], c)
        times, F = d["series"]["uh"]
        assert F.shape == (285, len(x)) and len(times) == 285
        U, _, dt = S.run_advection_stored(x, c, variant)
        err = np.linalg.norm(U - F, axis=1) / np.linalg.norm(F, axis=1)
        print(f"{variant}: oracle vs all 285 stored dolfinx frames, max rel L2 error {err.max():.2e}")
        assert err.max() < 1e-12
        np.savez_compressed(f"{HERE}/ref_series_{variant}.npz", frames=F[KEEP], index=np.array(KEEP),
                            times=times[KEEP], dt=np.array(dt))


def main():
    from oracle import p1, solvers as S
    from cfem_b200 import io, meshes

    x, c = io.read_mesh(f"{REF}/Code/Linear_advection/Data/RV/RV_node.h5")
    area, _ = p1.cell_geometry(x, c)
    assert abs(area.sum() - np.pi) < 5e-3 and np.all(area > 0)
    xdmf = open(f"{REF}/Code/Linear_advection/Data/RV/RV_node.xdmf").read()
    t0 = re.search(r'<Time Value="([0-9.eE+-]+)"', xdmf).group(1)
    np.savez_compressed(f"{HERE}/rv_node_mesh.npz", x=x, cells=c, first_time_stamp=np.array(float(t0)),
                        first_time_stamp_text=np.array(t0))
    x, c = io.read_mesh(f"{REF}/Data/KPP_RV.h5")
    area, _ = p1.cell_geometry(x, c)
    assert abs(area.sum() - 16.0) < 1e-9 and np.all(area > 0)
    np.savez_compressed(f"{HERE}/kpp_rv_mesh.npz", x=x, cells=c)
    d = np.load(f"{HERE}/rv_node_mesh.npz")
    stored_series(S, d["x"], d["cells"])

    # oracle outputs
    xb, cb = meshes.rectangle(24, 24)
    st, m, h = S.run_burgers(xb, cb, 0.5 / 24, 8)
    np.savez_compressed(f"{HERE}/burgers_24x24_8steps.npz", uh=st.uh, eps=st.eps, RH=st.RH, h=h,
                        newton_its=np.array(st.newton_its), dt=np.array(0.5 / 24))
    xk, ck = np.load(f"{HERE}/kpp_rv_mesh.npz")["x"], np.load(f"{HERE}/kpp_rv_mesh.npz")["cells"]
    st, m, h = S.run_kpp(xk, ck, 0.04, 5)   # dt = 0.64 h with h = 1/16 on this mesh (KPP_exact.py:38,75)
    np.savez_compressed(f"{HERE}/kpp_refmesh_5steps.npz", uh=st.uh, eps=st.eps, RH=st.RH, h=h,
                        newton_its=np.array(st.newton_its), dt=np.array(0.04))
    xa, ca = np.load(f"{HERE}/rv_node_mesh.npz")["x"], np.load(f"{HERE}/rv_node_mesh.npz")["cells"]
    dt = S.advection_dt(S.advection_velocity(xa), 1 / 16)
    uh, eps, m, h = S.run_advection(xa, ca, dt, 10)
    np.savez_compressed(f"{HERE}/advection_refmesh_10steps.npz", uh=uh, eps=eps, h=h, dt=np.array(dt))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
