"""CPU: the C ABI loads and exports everything include/cfem_b200.h declares; the host-side mesh
analysis (ordering, patches, boundary, tiles, packed codes) is right; nothing computes without a GPU;
the product never touches oracle/."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from cfem_b200 import _lib as L, meshes
from oracle import p1

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "cfem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(cfem_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 35
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/cfem_b200.h but not exported"
    assert names == set(L.SIGNATURES), names ^ set(L.SIGNATURES)
    assert lib.cfem_version() >= 100


def test_struct_layouts_match_header(lib):
    import ctypes as C

    assert C.sizeof(L.StepParams) == lib.cfem_struct_size(0) == 96
    assert C.sizeof(L.StepStats) == lib.cfem_struct_size(1) == 80


def test_no_cpu_fallback(lib):
    if lib.cfem_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from cfem_b200 import Context, CfemError

    x, c = meshes.rectangle(3, 3)
    with pytest.raises(CfemError, match="no CPU fallback"):
        Context((x, c))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "conservation-fem_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def emulate_tile_mass(x, hm):
    """numpy emulation of the tile kernel's gather (MassOp) from the packed codes."""
    n2u = hm["n2u"]
    c = hm["cells"].reshape(-1, 3)
    area, _ = p1.cell_geometry(x[n2u], c)
    Me = p1.mass_elements(area)
    rowptr, colidx = hm["rowptr"], hm["colidx"]
    vals = np.zeros(colidx.size)
    tn, tcp, tc, vp, code = hm["tile_node"], hm["tile_cellptr"], hm["tile_cells"], hm["v2c_ptr"], hm["v2c_code"]
    for t in range(tn.size - 1):
        cl_list = tc[tcp[t]:tcp[t + 1]]
        assert np.all(np.diff(cl_list) > 0)
        for i in range(tn[t], tn[t + 1]):
            for e in range(vp[i], vp[i + 1]):
                cd = int(code[e])
                cl, k = cd & 8191, (cd >> 13) & 3
                cell = cl_list[cl]
                assert c[cell, k] == i
                for j in range(3):
                    pos = (cd >> (15 + 5 * j)) & 31
                    assert colidx[rowptr[i] + pos] == c[cell, j]
                    vals[rowptr[i] + pos] += Me[cell, k, j]
    return vals


CASES = {
    "right": lambda: meshes.rectangle(13, 9),
    "crossed": lambda: meshes.rectangle(5, 6, diagonal="crossed"),
    "jittered_permuted": lambda: meshes.jittered(30, 22),
    "delaunay": lambda: meshes.delaunay(900),
    "single": lambda: (np.array([[0.0, 0, 0], [1.0, 0, 0], [0.0, 1, 0]]), np.array([[0, 1, 2]], dtype=np.int64)),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("order", [L.ORDER_HILBERT, L.ORDER_NATURAL])
def test_host_analysis(name, order):
    x, c = CASES[name]()
    hm = L.host_analyse(x, c, order)
    nn = x.shape[0]
    n2u = hm["n2u"]
    assert sorted(n2u.tolist()) == list(range(nn))           # a permutation
    if order == L.ORDER_NATURAL:
        assert np.array_equal(n2u, np.arange(nn))
    rp, ci = p1.patch_csr(c, nn)
    G = sp.csr_matrix((np.ones(hm["colidx"].size), hm["colidx"], hm["rowptr"]), shape=(nn, nn)).tocoo()
    Gu = sp.coo_matrix((G.data, (n2u[G.row], n2u[G.col])), shape=(nn, nn)).tocsr()
    Gu.sort_indices()
    assert np.array_equal(Gu.indptr, rp) and np.array_equal(Gu.indices, ci)      # patches, bit-exact
    assert np.array_equal(hm["bnd_user"], p1.boundary_nodes(c, nn))               # boundary dofs, bit-exact
    assert np.array_equal(np.flatnonzero(hm["is_bnd"]), np.sort(np.argsort(n2u)[hm["bnd_user"]]))
    # every cell appears in the tile lists of all its vertices' tiles; tiles respect their capacities
    assert np.diff(hm["tile_node"]).max() <= 256 and np.diff(hm["tile_cellptr"]).max() <= 768
    vals = emulate_tile_mass(np.asarray(x)[:, :2], hm)
    M = sp.csr_matrix((vals, hm["colidx"], hm["rowptr"]), shape=(nn, nn)).tocoo()
    Mu = sp.coo_matrix((M.data, (n2u[M.row], n2u[M.col])), shape=(nn, nn)).tocsr()
    ref = p1.mass_matrix(x, c)
    assert abs(Mu - ref).max() <= 1e-15 * abs(ref).max()


@pytest.mark.parametrize("name", ["right", "jittered_permuted", "delaunay"])
@pytest.mark.parametrize("order", [L.ORDER_HILBERT, L.ORDER_NATURAL])
@pytest.mark.parametrize("world", [1, 3])
def test_t16_tile_format_decodes_to_the_csr_columns(name, order, world):
    """The 16-bit tile-local columns of the SpMV-type kernels: index < 256 is a row of the same tile, otherwise
    256 + position in the tile's ascending external-column list; decoding gives colidx back bit for bit, external
    lists hold no own row and no duplicate, and ghost columns (>= n_owned) sort last."""
    x, c = CASES[name]() if name != "right" else meshes.rectangle(40, 37)
    for rank in range(world):
        hm = L.host_analyse(x, c, order, rank=rank, world=world)
        rp, ci, lc = hm["rowptr"], hm["colidx"], hm["lc16"].astype(np.int64)
        tn, ep, ext = hm["tile_node"], hm["tile_extptr"], hm["tile_ext"]
        assert lc.size == ci.size == hm["nnz"] and ep.size == tn.size and ep[-1] == ext.size
        for t in range(tn.size - 1):
            a, b = tn[t], tn[t + 1]
            e = ext[ep[t]:ep[t + 1]]
            assert np.all(np.diff(e) > 0) and not np.any((e >= a) & (e < b))
            p0, p1_ = rp[a], rp[b]
            loc = lc[p0:p1_]
            own = loc < 256
            dec = np.where(own, a + loc, e[np.clip(loc - 256, 0, max(e.size - 1, 0))] if e.size else a + loc)
            assert np.array_equal(dec, ci[p0:p1_])
            assert np.all(loc[~own] - 256 < e.size)
        order_t = hm["tile_order"]
        assert sorted(order_t.tolist()) == list(range(tn.size - 1))


@pytest.mark.parametrize("world", [1, 3])
def test_last_cell_is_highest_numbered_incident_cell(world):
    """RV_cell.py:190-192 writes each cell's viscosity to its dofs in cell order: a node keeps the value
    of its highest-numbered cell.  last_cell names that cell (by its vertex set; cells are renumbered)."""
    x, c = meshes.permuted(*meshes.jittered(14, 11), np.random.default_rng(3))
    want = {}
    for k, tri in enumerate(c):
        for v in tri:
            want[int(v)] = frozenset(int(t) for t in tri)   # later cells overwrite
    for rank in range(world):
        hm = L.host_analyse(x, c, rank=rank, world=world)
        n2u, cl = hm["n2u"], hm["cells"].reshape(-1, 3)
        lc = hm["last_cell"]
        assert lc.size == (x.shape[0] if world == 1 else lc.size) and lc.min() >= 0
        for i, k in enumerate(lc):
            assert frozenset(int(n2u[v]) for v in cl[k]) == want[int(n2u[i])]


def test_hilbert_order_is_local():
    """Consecutive internal ids are spatial neighbours: the tile halo stays small."""
    x, c = meshes.jittered(96, 96)
    hm = L.host_analyse(x, c)
    staged = hm["tile_cells"].size / c.shape[0]
    assert staged < 1.35, staged   # each cell staged ~1.2 times on average (vs ~3 for a random order)
    hn = L.host_analyse(x, c, L.ORDER_NATURAL)
    assert hn["tile_cells"].size / c.shape[0] > 2.5


def test_bad_meshes_rejected():
    x = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [5.0, 5.0]])
    with pytest.raises(L.CfemError, match="belongs to no cell"):
        L.host_analyse(x, np.array([[0, 1, 2]], dtype=np.int32))
    with pytest.raises(L.CfemError, match="out-of-range"):
        L.host_analyse(x[:3], np.array([[0, 1, 7]], dtype=np.int32))
    with pytest.raises(L.CfemError, match="repeated"):
        L.host_analyse(x[:3], np.array([[0, 1, 1]], dtype=np.int32))
    # valence above the 5-bit row-position encoding: a fan of 40 triangles around one node
    k = 40
    th = np.linspace(0, 2 * np.pi, k, endpoint=False)
    xf = np.concatenate([[[0.0, 0.0]], np.stack([np.cos(th), np.sin(th)], 1)])
    cf = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.int32)
    with pytest.raises(L.CfemError, match="neighbours"):
        L.host_analyse(xf, cf)
