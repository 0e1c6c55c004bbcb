"""CPU: mesh / time-series I/O: the XDMF writer round-trips through the reader; the
HDF5 reader opens the reference's own dolfinx files when they are present (authoring container)."""
import os

import numpy as np
import pytest

from cfem_b200 import io, meshes
from cfem_b200.solvers import NodalFunction

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference"


def test_xdmf_writer_roundtrip(tmp_path):
    x, c = meshes.jittered(7, 5)
    rng = np.random.default_rng(0)
    path = str(tmp_path / "solution.xdmf")
    frames, vec = [], rng.normal(size=(x.shape[0], 2))
    with io.XdmfWriter(path, x, c) as w:                      # xdmf.write_mesh(domain)
        t = 0.0
        for k in range(4):
            t += 0.1
            uh = NodalFunction(rng.normal(size=x.shape[0]), "uh")
            frames.append(uh.x.array.copy())
            w.write_function(uh, t)                            # xdmf.write_function(uh, t)
            if k == 1:                                         # readable while the run is still going
                d = io.read_xdmf(path)
                assert d["series"]["uh"][1].shape == (2, x.shape[0])
        w.write_function(vec, 0.4, name="w")
    d = io.read_xdmf(path)
    assert np.array_equal(d["x"], x) and np.array_equal(d["cells"], c)
    t, F = d["series"]["uh"]
    acc, s = [], 0.0
    for _ in range(4):
        s += 0.1
        acc.append(s)
    assert np.array_equal(t, np.array(acc)) and np.array_equal(F, np.array(frames))   # times bit-exact via repr
    assert np.array_equal(d["series"]["w"][1][0], vec)
    xm, cm = io.read_mesh(path)
    assert np.array_equal(xm, x) and np.array_equal(cm, c)
    assert os.path.getsize(str(tmp_path / "solution.bin")) == c.size * 4 + x.size * 8 + 4 * x.shape[0] * 8 + vec.size * 8


def test_h5_reader_rejects_garbage(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(io.H5Error):
        io.H5File(str(p))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reader_opens_the_reference_files():
    g = np.load(os.path.join(GOLD, "kpp_rv_mesh.npz"))
    f = io.H5File(f"{REF}/Data/KPP_RV.h5")
    assert f.datasets["/Mesh/mesh/topology"][0] == (9514, 3) and f.datasets["/Mesh/mesh/geometry"][0] == (4886, 2)
    for path in (f"{REF}/Data/KPP_RV.h5", f"{REF}/Data/KPP_RV.xdmf", f"{REF}/Code/KPP/Data/KPP_RV.h5"):
        x, c = io.read_mesh(path)
        assert np.array_equal(x, g["x"]) and np.array_equal(c, g["cells"])
    for variant, rel in (("eps_func", "RV/RV_node"), ("rv_cell", "RV/RV_cell"), ("si_old", "SI/smoothness")):
        d = io.read_xdmf(f"{REF}/Code/Linear_advection/Data/{rel}.xdmf")
        r = np.load(os.path.join(GOLD, f"ref_series_{variant}.npz"))
        t, F = d["series"]["uh"]
        assert F.shape == (285, 1011)
        assert np.array_equal(F[r["index"]], r["frames"]) and np.array_equal(t[r["index"]], r["times"])
