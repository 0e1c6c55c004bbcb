"""Regenerates the fixtures in tests/golden/ (run in the authoring container only).

1. Meshes lifted from the reference's own data files (contiguous HDF5 datasets read by
   byte offset, h5py is not available; offsets documented in SURVEY.md section 4):
     Code/Linear_advection/Data/RV/RV_node.h5  -> rv_node_mesh.npz  (1,919 tris / 1,011 nodes, unit disk)
     Data/KPP_RV.h5                            -> kpp_rv_mesh.npz   (9,514 tris / 4,886 nodes, [-2,2]^2)
   plus the first <Time> stamp of RV_node.xdmf, which pins the reference's dt formula.
2. Oracle outputs on small seeded cases (fields after N steps), so GPU parity can also be
   checked against committed vectors.  These come from oracle/ (CPU restatement), not from
   dolfinx: the reference cannot run in this image ("parity unpinned", see oracle/__init__.py).
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


def read_mesh(path, n_cells, n_nodes, topo_off, geom_off):
    raw = open(path, "rb").read()
    cells = np.frombuffer(raw, dtype="<i8", count=3 * n_cells, offset=topo_off).reshape(-1, 3)
    if geom_off < 0:
        geom_off = len(raw) + geom_off
    x = np.frombuffer(raw, dtype="<f8", count=2 * n_nodes, offset=geom_off).reshape(-1, 2)
    assert cells.min() == 0 and cells.max() == n_nodes - 1
    return x.copy(), cells.astype(np.int32)


def main():
    from oracle import p1, solvers as S
    from cfem_b200 import meshes

    x, c = read_mesh(f"{REF}/Code/Linear_advection/Data/RV/RV_node.h5", 1919, 1011, 3464, 51568)
    area, _ = p1.cell_geometry(x, c)
    assert abs(area.sum() - np.pi) < 5e-3 and np.all(area > 0)
    xdmf = open(f"{REF}/Code/Linear_advection/Data/RV/RV_node.xdmf").read()
    t0 = re.search(r'<Time Value="([0-9.eE+-]+)"', xdmf).group(1)
    np.savez_compressed(f"{HERE}/rv_node_mesh.npz", x=x, cells=c, first_time_stamp=np.array(float(t0)),
                        first_time_stamp_text=np.array(t0))
    x, c = read_mesh(f"{REF}/Data/KPP_RV.h5", 9514, 4886, 3464, -78176)
    area, _ = p1.cell_geometry(x, c)
    assert abs(area.sum() - 16.0) < 1e-9 and np.all(area > 0)
    np.savez_compressed(f"{HERE}/kpp_rv_mesh.npz", x=x, cells=c)

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
