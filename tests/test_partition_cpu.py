"""CPU: the partition / halo layer of the multi-GPU path (cfem_create_distributed's host analysis).

(1) single process: the lists of all ranks are mutually consistent for several world sizes;
(2) world_size 2 over torch.distributed (gloo): a real halo exchange + owned-row SpMV reproduces
    the global operator (arithmetic by the oracle / numpy: this is host logic, no GPU)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from cfem_b200 import _lib as L, meshes
from oracle import p1


def local_mass_rows(x, hm):
    """Mass-matrix rows of the owned nodes, assembled from the rank's local cells only."""
    xy = x[hm["n2u"]]
    c = hm["cells"].reshape(-1, 3)
    area, _ = p1.cell_geometry(xy, c)
    Me = p1.mass_elements(area)
    no, nl = hm["n_owned"], hm["n_local"]
    rows = np.repeat(c, 3, axis=1).ravel()
    cols = np.tile(c, (1, 3)).ravel()
    keep = rows < no
    A = sp.coo_matrix((Me.reshape(-1)[keep], (rows[keep], cols[keep])), shape=(no, nl)).tocsr()
    # the CSR pattern the library built for the owned rows must be exactly this pattern
    G = sp.csr_matrix((np.ones(hm["colidx"].size), hm["colidx"], hm["rowptr"]), shape=(no, nl))
    G.sort_indices()
    A.sort_indices()
    assert np.array_equal(G.indptr, A.indptr) and np.array_equal(G.indices, A.indices)
    return A


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("mesh", ["rect", "jittered"])
def test_partition_consistency(world, mesh):
    x, c = meshes.rectangle(24, 17) if mesh == "rect" else meshes.jittered(21, 19)
    check_partition(x, c, world)


def test_partition_consistency_random_meshes():
    """Property test: unstructured Delaunay meshes with random node / cell numbering, random world sizes
    (ragged partitions, ranks whose range is a handful of nodes)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=12, deadline=None, derandomize=True)
    @given(st.integers(40, 400), st.integers(1, 9), st.integers(0, 10 ** 6))
    def run(npts, world, seed):
        x, c = meshes.delaunay(npts, seed=seed % 1000)
        x, c = meshes.permuted(x, c, np.random.default_rng(seed))
        check_partition(x, c, world)

    run()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("mesh", ["rect", "jittered"])
def test_partition_consistency_metis(world, mesh):
    """The same consistency contract with the METIS k-way partition of the nodal graph (node_part from
    cfem_host_partition): parts become contiguous ranges of the internal order, everything downstream is shared."""
    x, c = meshes.rectangle(24, 17) if mesh == "rect" else meshes.jittered(21, 19)
    part = L.host_partition(x, c, world, "metis")
    assert part.min() == 0 and part.max() == world - 1
    assert np.array_equal(part, L.host_partition(x, c, world, "metis"))   # deterministic (fixed seed)
    parts = check_partition(x, c, world, node_part=part)
    for r, p in enumerate(parts):                                          # rank r owns exactly METIS part r
        assert np.array_equal(np.sort(p["n2u"][: p["n_owned"]]), np.flatnonzero(part == r))
    # the Hilbert-range partition written as a node_part array gives the default analysis back, array for array
    hp = L.host_partition(x, c, world, "hilbert")
    for r in range(world):
        a, b = L.host_analyse(x, c, rank=r, world=world), L.host_analyse(x, c, rank=r, world=world, node_part=hp)
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_metis_halo_is_smaller_than_hilbert_ranges_on_the_kpp_mesh():
    """What the graph partition buys on the jittered KPP mesh (numbers for DESIGN.md section 5 come from the same
    code at 1448^2): fewer ghost values per rank and no more neighbours than the curve ranges."""
    x, c = meshes.jittered(160, 160, (-2.0, -2.0), (2.0, 2.0))
    world = 8
    part = L.host_partition(x, c, world, "metis")
    ghosts = {"hilbert": [], "metis": []}
    for name, npart in (("hilbert", None), ("metis", part)):
        for r in range(world):
            p = L.host_analyse(x, c, rank=r, world=world, node_part=npart)
            ghosts[name].append(p["n_local"] - p["n_owned"])
    assert sum(ghosts["metis"]) < sum(ghosts["hilbert"])
    counts = np.bincount(part, minlength=world)
    assert counts.max() <= 1.04 * counts.mean()                            # METIS load balance (ufactor 1.03)


def check_partition(x, c, world, node_part=None):
    nn = x.shape[0]
    parts = [L.host_analyse(x, c, rank=r, world=world, node_part=node_part) for r in range(world)]
    owned = [p["n2u"][: p["n_owned"]] for p in parts]
    allowned = np.concatenate(owned)
    assert np.array_equal(np.sort(allowned), np.arange(nn))            # a partition of the dofs
    if node_part is None:
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1  # balanced
    ncell_sum = sum(p["n_cells"] for p in parts)
    assert ncell_sum >= c.shape[0] and (world > 1 or ncell_sum == c.shape[0])
    M = p1.mass_matrix(x, c)
    v = np.random.default_rng(0).normal(size=nn)
    for r, p in enumerate(parts):
        no = p["n_owned"]
        ghosts_user = p["n2u"][no:]
        assert not set(ghosts_user.tolist()) & set(owned[r].tolist())
        # what I receive from q is exactly what q sends to me, in the same order
        for k, q in enumerate(p["peer_rank"]):
            mine = p["n2u"][p["recv_off"][k]: p["recv_off"][k] + p["recv_cnt"][k]]
            pq = parts[q]
            kq = list(pq["peer_rank"]).index(r)
            theirs = pq["n2u"][pq["send_idx"][pq["send_ptr"][kq]: pq["send_ptr"][kq + 1]]]
            assert np.array_equal(mine, theirs)
        assert sum(p["recv_cnt"]) == p["n_local"] - no                  # every ghost has a source
        # owned rows assembled from local cells == global rows
        A = local_mass_rows(x, p)
        assert np.allclose(A @ v[p["n2u"]], (M @ v)[owned[r]], rtol=0, atol=1e-15)
        # boundary flags of owned AND ghost nodes come from the global mesh
        bnd = np.zeros(nn, bool)
        bnd[p1.boundary_nodes(c, nn)] = True
        assert np.array_equal(p["is_bnd"].astype(bool), bnd[p["n2u"]])
        assert np.array_equal(p["bnd_user"], p1.boundary_nodes(c, nn))
    return parts


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, c = meshes.jittered(26, 23)
        nn = x.shape[0]
        p = L.host_analyse(x, c, rank=rank, world=world)
        no, nl = p["n_owned"], p["n_local"]
        A = local_mass_rows(x, p)
        rng = np.random.default_rng(7)
        v_global = rng.normal(size=nn)
        v = np.full(nl, np.nan)
        v[:no] = v_global[p["n2u"][:no]]                     # ghosts unknown until the exchange
        reqs, bufs = [], []
        for k, peer in enumerate(p["peer_rank"]):
            send = torch.from_numpy(v[p["send_idx"][p["send_ptr"][k]: p["send_ptr"][k + 1]]].copy())
            recv = torch.empty(int(p["recv_cnt"][k]), dtype=torch.float64)
            if send.numel():
                reqs.append(dist.isend(send, int(peer)))
            if recv.numel():
                reqs.append(dist.irecv(recv, int(peer)))
            bufs.append((k, recv, send))
        for r in reqs:
            r.wait()
        for k, recv, _ in bufs:
            v[p["recv_off"][k]: p["recv_off"][k] + p["recv_cnt"][k]] = recv.numpy()
        assert not np.isnan(v).any()
        assert np.array_equal(v, v_global[p["n2u"]])          # forward halo exchange is exact
        y = A @ v
        # dot products: owned partial sums + all-reduce == global dot
        t = torch.tensor([float(v[:no] @ y)], dtype=torch.float64)
        dist.all_reduce(t)
        M = p1.mass_matrix(x, c)
        ok = np.allclose(y, (M @ v_global)[p["n2u"][:no]], rtol=0, atol=1e-15) and \
            abs(t.item() - v_global @ (M @ v_global)) < 1e-12
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}
