// Graph partition of the mesh for the multi-GPU path: METIS k-way on the nodal graph (one vertex per P1 dof, one
// edge per mesh edge), as the reference's tool chain does for its MPI runs (dolfinx partitions with the graph
// partitioners of its environment: metis=5.1.0 / parmetis=4.0.3, Environment/fenicsx-env.yml:171,192).
//
// The result is one part id per node in CALLER numbering; cfem_create_partitioned / cfem_host_analyse_partitioned
// order the nodes by (part, Hilbert key), so a part is a contiguous range of the internal order exactly like the
// default equal-range partition of the Hilbert curve -- everything downstream (ghost layer, halo lists, tiles) is
// shared.  METIS itself comes from the CUDA toolkit's libmetis_static.a (64-bit idx_t, 32-bit real_t; there is no
// header in the toolkit, so the two entry points are declared here).
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "internal.h"

extern "C" {
typedef int64_t metis_idx_t;
typedef float metis_real_t;
int METIS_SetDefaultOptions(metis_idx_t* options);
int METIS_PartGraphKway(metis_idx_t* nvtxs, metis_idx_t* ncon, metis_idx_t* xadj, metis_idx_t* adjncy, metis_idx_t* vwgt,
                        metis_idx_t* vsize, metis_idx_t* adjwgt, metis_idx_t* nparts, metis_real_t* tpwgts,
                        metis_real_t* ubvec, metis_idx_t* options, metis_idx_t* objval, metis_idx_t* part);
}

namespace cfem {

template <class I>
static void nodal_graph(int64_t nn, int64_t nc, const I* cells, std::vector<metis_idx_t>& xadj, std::vector<metis_idx_t>& adj) {
  std::vector<int64_t> deg(nn + 1, 0);
  for (int64_t c = 0; c < nc; ++c)
    for (int k = 0; k < 3; ++k) {
      const int64_t v = (int64_t)cells[3 * c + k];
      if (v < 0 || v >= nn) CFEM_THROW(-1, "cell connectivity has an out-of-range vertex");
      deg[v + 1] += 2;
    }
  for (int64_t i = 0; i < nn; ++i) deg[i + 1] += deg[i];
  std::vector<int32_t> raw(deg[nn]);
  {
    std::vector<int64_t> fill(deg.begin(), deg.end() - 1);
    for (int64_t c = 0; c < nc; ++c) {
      const int64_t v[3] = {(int64_t)cells[3 * c], (int64_t)cells[3 * c + 1], (int64_t)cells[3 * c + 2]};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
          if (a != b) raw[fill[v[a]]++] = (int32_t)v[b];
    }
  }
  xadj.assign(nn + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    int32_t* b = raw.data() + deg[i];
    int32_t* e = raw.data() + deg[i + 1];
    std::sort(b, e);
    xadj[i + 1] = (metis_idx_t)(std::unique(b, e) - b);
  }
  for (int64_t i = 0; i < nn; ++i) xadj[i + 1] += xadj[i];
  adj.resize(xadj[nn]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    const int64_t len = xadj[i + 1] - xadj[i];
    for (int64_t k = 0; k < len; ++k) adj[xadj[i] + k] = raw[deg[i] + k];
  }
}

void metis_partition(int world, int64_t nn, int64_t nc, const void* cells, int idx_bytes, int32_t* part_out) {
  if (world < 1 || nn <= 0 || nc <= 0 || !cells || !part_out) CFEM_THROW(-1, "partition: bad argument");
  if (idx_bytes != 4 && idx_bytes != 8) CFEM_THROW(-1, "cell_index_bytes must be 4 or 8");
  if (world == 1) { std::fill(part_out, part_out + nn, 0); return; }
  std::vector<metis_idx_t> xadj, adj;
  if (idx_bytes == 4) nodal_graph(nn, nc, (const int32_t*)cells, xadj, adj);
  else nodal_graph(nn, nc, (const int64_t*)cells, xadj, adj);
  metis_idx_t nv = nn, ncon = 1, nparts = world, objval = 0;
  metis_idx_t options[40];
  METIS_SetDefaultOptions(options);
  options[8] = 20241118;   // METIS_OPTION_SEED: fixed, every rank that calls this gets the same partition
  options[16] = 1;         // METIS_OPTION_CONTIG: connected parts
  std::vector<metis_idx_t> part(nn);
  const int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &nparts, nullptr,
                                     nullptr, options, &objval, part.data());
  if (rc != 1) CFEM_THROW(-6, "METIS_PartGraphKway failed with code " + std::to_string(rc));
  for (int64_t i = 0; i < nn; ++i) part_out[i] = (int32_t)part[i];
}

}  // namespace cfem
