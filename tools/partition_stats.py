"""Halo size / neighbour count of the two partitions on a jittered KPP mesh (host only, no GPU).
  python tools/partition_stats.py [n=1448] [world=8]"""
import sys, time
sys.path.insert(0, "conservation-fem_b200")
import numpy as np
from cfem_b200 import _lib as L, meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1448
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
x, c = meshes.jittered(n, n, (-2.0, -2.0), (2.0, 2.0))
t = time.time(); part = L.host_partition(x, c, world, "metis"); tm = time.time() - t
print(f"mesh {n}x{n} jittered: {x.shape[0]} nodes, {c.shape[0]} cells; METIS k-way {world} parts in {tm:.1f} s")
for name, npart in (("hilbert ranges", None), ("metis k-way", part)):
    gh, snd, peers, own, cells, ta = [], [], [], [], [], 0.0
    for r in range(world):
        t = time.time(); p = L.host_analyse(x, c, rank=r, world=world, node_part=npart); ta += time.time() - t
        gh.append(p["n_local"] - p["n_owned"]); snd.append(p["send_idx"].size); peers.append(p["peer_rank"].size)
        own.append(p["n_owned"]); cells.append(p["n_cells"])
    print(f"{name:15s} owned min/max {min(own)}/{max(own)}  ghosts mean {np.mean(gh):.0f} max {max(gh)}  "
          f"sent values mean {np.mean(snd):.0f} max {max(snd)}  neighbours mean {np.mean(peers):.1f} max {max(peers)}  "
          f"local cells max {max(cells)}  analysis {ta / world:.2f} s per rank")
