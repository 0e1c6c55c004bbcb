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
                w.flush()                                      # (files are brought up to date every flush_every frames / on flush)
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


# ---------------------------------------------------------------------------------------- HDF5 writer
import struct  # noqa: E402


def h5_structure(path):
    """Independent walk of an old-style HDF5 file that checks the invariants libhdf5 relies on:
    B-tree keys bracket their children, names strictly increasing, sibling links consistent, node fill within
    [1, 2K], heap strings inside the data segment.  Returns {dataset path: {msg type: body bytes}}."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0
    leaf_k, int_k = struct.unpack_from("<HH", b, 16)
    eof = struct.unpack_from("<Q", b, 40)[0]
    assert eof == len(b)
    out = {}

    def messages(a):
        ver, _, nmsg, ref, size = struct.unpack_from("<BBHII", b, a)
        assert ver == 1 and ref == 1
        blocks, res = [(a + 16, size)], []
        while blocks and len(res) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(res) < nmsg:
                t, s, _f = struct.unpack_from("<HHB", b, p)
                assert s % 8 == 0
                if t == 0x10:
                    blocks.append(struct.unpack_from("<QQ", b, p + 8))
                res.append((t, b[p + 8:p + 8 + s]))
                p += 8 + s
        assert len(res) == nmsg
        return res

    def group(prefix, bt, hp):
        assert b[hp:hp + 4] == b"HEAP"
        dsize, _free, daddr = struct.unpack_from("<QQQ", b, hp + 8)

        def name(off):
            assert off < dsize
            s = daddr + off
            e = b.index(b"\0", s)
            assert e < daddr + dsize
            return b[s:e].decode()

        seen = []

        def node(a, lo_key, hi_key, is_root):
            if b[a:a + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, a + 6)[0]
                assert 1 <= n <= 2 * leaf_k
                for k in range(n):
                    e = a + 8 + 40 * k
                    off, hdr, ctype, _ = struct.unpack_from("<QQII", b, e)
                    nm = name(off)
                    assert (not seen or seen[-1] < nm) and name(lo_key) < nm <= name(hi_key) or (name(lo_key) == "" and nm <= name(hi_key))
                    seen.append(nm)
                    msgs = messages(hdr)
                    stab = [m for m in msgs if m[0] == 0x11]
                    if stab:
                        sbt, shp = struct.unpack("<QQ", stab[0][1])
                        if ctype == 1:
                            assert b[e + 24:e + 40] == stab[0][1]
                        group(f"{prefix}/{nm}", sbt, shp)
                    else:
                        out[f"{prefix}/{nm}"] = {t: body for t, body in msgs}
                return
            assert b[a:a + 4] == b"TREE" and b[a + 4] == 0
            level, used = b[a + 5], struct.unpack_from("<H", b, a + 6)[0]
            assert 1 <= used <= 2 * int_k
            keys = [struct.unpack_from("<Q", b, a + 24 + 16 * k)[0] for k in range(used + 1)]
            kids = [struct.unpack_from("<Q", b, a + 32 + 16 * k)[0] for k in range(used)]
            assert keys[0] == lo_key and keys[-1] == hi_key
            for k, child in enumerate(kids):
                if level > 0:
                    assert b[child:child + 4] == b"TREE" and b[child + 5] == level - 1
                    ls, rs = struct.unpack_from("<QQ", b, child + 8)
                    assert ls == (kids[k - 1] if k else ls) and rs == (kids[k + 1] if k + 1 < used else rs)
                node(child, keys[k], keys[k + 1], False)

        used = struct.unpack_from("<H", b, bt + 6)[0]
        first = struct.unpack_from("<Q", b, bt + 24)[0]
        last = struct.unpack_from("<Q", b, bt + 24 + 16 * used)[0]
        assert name(first) == ""
        node(bt, first, last, True)

    _, root, cache, _ = struct.unpack_from("<QQII", b, 56)
    bt, hp = struct.unpack_from("<QQ", b, 80)
    assert cache == 1 and [m for m in messages(root) if m[0] == 0x11][0][1] == b[80:96]
    group("", bt, hp)
    return out


def test_h5_writer_roundtrip_and_structure(tmp_path):
    rng = np.random.default_rng(0)
    x, c = meshes.jittered(9, 7)
    p = str(tmp_path / "out.h5")
    w = io.H5Writer(p)
    w.write("/Mesh/mesh/topology", c.astype(np.int64))
    w.write("/Mesh/mesh/geometry", x)
    frames = {}
    for k in range(300):                      # > 256 entries in one group: a two-level B-tree
        nm = "/Function/uh/" + repr(0.0035 * (k + 1)).replace(".", "_")
        frames[nm] = rng.normal(size=(x.shape[0], 1))
        w.write(nm, frames[nm])
        if k == 40:                           # complete file after every flush, more data may follow
            w.flush()
            assert len(io.H5File(p).datasets) == 43 and len(h5_structure(p)) == 43
    w.close()
    f = io.H5File(p)
    assert np.array_equal(f.read("/Mesh/mesh/topology"), c) and f.read("/Mesh/mesh/topology").dtype == np.int64
    assert np.array_equal(f.read("/Mesh/mesh/geometry"), x)
    assert all(np.array_equal(f.read(k), v) for k, v in frames.items())
    s = h5_structure(p)
    assert set(s) == set(f.datasets) and len(s) == 302
    with pytest.raises(io.H5Error):
        io.H5Writer(str(tmp_path / "bad.h5")).write("/a", np.array(["x"]))


def test_xdmf_writer_hdf5_backend(tmp_path):
    x, c = meshes.jittered(6, 5)
    path = str(tmp_path / "series.xdmf")
    rng = np.random.default_rng(1)
    fr = [rng.normal(size=x.shape[0]) for _ in range(3)]
    with io.XdmfWriter(path, x, c, heavy="hdf5") as w:
        for k, a in enumerate(fr):
            w.write_function(a, 0.25 * (k + 1), name="uh")
    d = io.read_xdmf(path)
    assert np.array_equal(d["x"], x) and np.array_equal(d["cells"], c)
    assert np.array_equal(d["series"]["uh"][0], [0.25, 0.5, 0.75]) and np.array_equal(d["series"]["uh"][1], np.array(fr))
    names = sorted(io.H5File(str(tmp_path / "series.h5")).datasets)
    assert names == ["/Function/uh/0_25", "/Function/uh/0_5", "/Function/uh/0_75", "/Mesh/mesh/geometry", "/Mesh/mesh/topology"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_h5_writer_matches_libhdf5_encodings(tmp_path):
    """The structure checker accepts the files libhdf5 wrote for the reference, and this writer's dataset
    messages (dataspace, datatype, fill value, layout) are byte-identical to libhdf5's for the same data."""
    ref_mesh = h5_structure(f"{REF}/Data/KPP_RV.h5")
    ref_series = h5_structure(f"{REF}/Code/Linear_advection/Data/RV/RV_node.h5")
    assert len(ref_mesh) == 2 and len(ref_series) == 287
    x, c = io.read_mesh(f"{REF}/Data/KPP_RV.h5")
    p = str(tmp_path / "kpp.h5")
    w = io.H5Writer(p)
    w.write("/Mesh/mesh/topology", c.astype(np.int64))
    w.write("/Mesh/mesh/geometry", x)
    w.write("/Function/uh/0_5", x[:, :1])
    w.close()
    mine = h5_structure(p)
    for name in ("/Mesh/mesh/topology", "/Mesh/mesh/geometry"):
        for t in (0x0001, 0x0003, 0x0005):
            assert mine[name][t] == ref_mesh[name][t], (name, hex(t))
        a, r = mine[name][0x0008], ref_mesh[name][0x0008]
        assert a[:2] == r[:2] and a[10:] == r[10:]        # same class and size; the address differs
    frame = next(v for k, v in ref_series.items() if k.startswith("/Function/uh/"))
    for t in (0x0001, 0x0003, 0x0005):
        assert mine["/Function/uh/0_5"][t][:8] == frame[t][:8]
