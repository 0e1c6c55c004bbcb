// Host-side, once-per-mesh analysis: Hilbert ordering, vertex->cell adjacency,
// P1 CSR pattern (== node patches, reference Code/Utils/SI.py:12-28), boundary
// dofs (reference Code/KPP/KPP_exact.py:85-89), the partition of the Hilbert
// curve into one contiguous range of nodes per rank with its ghost layer and
// halo-exchange lists, and the assembly tiles with their packed (node, cell)
// codes.  Everything is O(N) apart from one sort.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

#include "internal.h"

namespace cfem {

static inline uint64_t hilbert_key(uint32_t x, uint32_t y, int bits) {
  const uint32_t n = 1u << bits;
  uint64_t d = 0;
  for (uint32_t s = n >> 1; s > 0; s >>= 1) {
    const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
    d += (uint64_t)s * (uint64_t)s * ((3u * rx) ^ ry);
    if (ry == 0) {
      if (rx == 1) { x = n - 1 - x; y = n - 1 - y; }
      std::swap(x, y);
    }
  }
  return d;
}

template <class It>
static void sort_pairs(It b, It e) {
#ifdef _OPENMP
  __gnu_parallel::sort(b, e);
#else
  std::sort(b, e);
#endif
}

// ---- stage A: what every rank computes for the WHOLE mesh (light: a few passes and one sort) ------------------
struct GlobalOrder {
  int64_t nn = 0, nc = 0;
  std::vector<int32_t> n2u, u2n;       // global internal <-> caller numbering
  std::vector<int64_t> part_off;       // world+1 offsets into the internal order
  std::vector<uint8_t> is_bnd_user;    // boundary flag per caller node
};

static inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser: a bijection, so distinct nodes get distinct keys
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

template <class I>
static inline int64_t cell_vertex(const void* cells, int64_t c, int k) { return (int64_t)((const I*)cells)[3 * c + k]; }

// Node ordering (Hilbert curve, or caller order), the partition of that order into one contiguous range per rank
// (equal ranges, or -- node_part given -- the caller's parts, e.g. from METIS, each kept in Hilbert order), and the
// boundary nodes of the mesh.
static void global_order(GlobalOrder& g, int64_t nn, int64_t nc, const double* x, int xdim, const void* cells_in,
                         int idx_bytes, int order, int world, const int32_t* node_part) {
  if (nn <= 0 || nc <= 0) CFEM_THROW(-1, "empty mesh");
  if (nn >= (int64_t)1 << 31 || 3 * nc >= (int64_t)1 << 31) CFEM_THROW(-1, "mesh too large for int32 indices");
  if (xdim != 2 && xdim != 3) CFEM_THROW(-1, "xdim must be 2 or 3");
  if (idx_bytes != 4 && idx_bytes != 8) CFEM_THROW(-1, "cell_index_bytes must be 4 or 8");
  g.nn = nn;
  g.nc = nc;
  // ---- 1. node ordering ----------------------------------------------------
  g.n2u.resize(nn);
  g.u2n.resize(nn);
  g.part_off.assign(world + 1, 0);
  if (node_part) {
    for (int64_t i = 0; i < nn; ++i) {
      if (node_part[i] < 0 || node_part[i] >= world) CFEM_THROW(-1, "node_part entry out of range");
      g.part_off[node_part[i] + 1]++;
    }
    for (int r = 0; r < world; ++r) {
      if (g.part_off[r + 1] == 0) CFEM_THROW(-1, "a rank owns no node: empty part in node_part");
      g.part_off[r + 1] += g.part_off[r];
    }
  } else {
    for (int r = 0; r <= world; ++r) g.part_off[r] = (int64_t)((__int128)nn * r / world);
    for (int r = 0; r < world; ++r)
      if (g.part_off[r + 1] <= g.part_off[r]) CFEM_THROW(-1, "a rank owns no node: mesh too small for this many ranks");
  }
  if (order == CFEM_ORDER_NATURAL && !node_part) {
    std::iota(g.n2u.begin(), g.n2u.end(), 0);
  } else {
    double xmin = x[0], xmax = x[0], ymin = x[1], ymax = x[1];
    for (int64_t i = 0; i < nn; ++i) {
      const double a = x[i * xdim], b = x[i * xdim + 1];
      if (!(std::isfinite(a) && std::isfinite(b))) CFEM_THROW(-1, "non-finite node coordinate");
      xmin = std::min(xmin, a); xmax = std::max(xmax, a);
      ymin = std::min(ymin, b); ymax = std::max(ymax, b);
    }
    const int bits = 24;
    const double span = std::max(std::max(xmax - xmin, ymax - ymin), 1e-300);
    const double scale = (double)((1u << bits) - 1) / span;
    const bool hilbert = order != CFEM_ORDER_NATURAL;
    std::vector<std::pair<uint64_t, int32_t>> keys(nn);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) {
      uint64_t k;
      if (hilbert) {
        const uint32_t qx = (uint32_t)((x[i * xdim] - xmin) * scale);
        const uint32_t qy = (uint32_t)((x[i * xdim + 1] - ymin) * scale);
        k = hilbert_key(qx, qy, bits);   // < 2^48
      } else {
        k = (uint64_t)i;
      }
      if (node_part) k |= (uint64_t)node_part[i] << 48;   // parts first, each in curve order
      keys[i] = {k, (int32_t)i};
    }
    sort_pairs(keys.begin(), keys.end());
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) g.n2u[i] = keys[i].second;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) g.u2n[g.n2u[i]] = (int32_t)i;

  // ---- 2. connectivity checks + boundary nodes, one pass over the cells, no relation built ------------------
  // Around a node every neighbour appears once per cell that contains the edge to it: twice for an interior edge,
  // once for a boundary edge.  XOR-ing a bijective 64-bit key of the two other vertices of each incident cell
  // therefore leaves 0 exactly for nodes without a boundary edge.
  std::vector<uint64_t> acc(nn, 0);
  std::vector<uint8_t> touched(nn, 0);
  bool range_bad = false, rep_bad = false;
#pragma omp parallel for schedule(static) reduction(|| : range_bad) reduction(|| : rep_bad)
  for (int64_t c = 0; c < nc; ++c) {
    int64_t v[3];
    for (int k = 0; k < 3; ++k) {
      v[k] = idx_bytes == 4 ? cell_vertex<int32_t>(cells_in, c, k) : cell_vertex<int64_t>(cells_in, c, k);
      if (v[k] < 0 || v[k] >= nn) { range_bad = true; v[k] = 0; }
    }
    if (v[0] == v[1] || v[0] == v[2] || v[1] == v[2]) rep_bad = true;
    const uint64_t h0 = mix64((uint64_t)v[0]), h1 = mix64((uint64_t)v[1]), h2 = mix64((uint64_t)v[2]);
    const uint64_t a0 = h1 ^ h2, a1 = h0 ^ h2, a2 = h0 ^ h1;
    // XOR commutes: the result does not depend on the order in which the threads get here
#pragma omp atomic
    acc[v[0]] ^= a0;
#pragma omp atomic
    acc[v[1]] ^= a1;
#pragma omp atomic
    acc[v[2]] ^= a2;
    touched[v[0]] = touched[v[1]] = touched[v[2]] = 1;   // benign same-value races
  }
  if (range_bad) CFEM_THROW(-1, "cell connectivity has an out-of-range vertex");
  if (rep_bad) CFEM_THROW(-1, "cell connectivity has a repeated vertex");
  g.is_bnd_user.resize(nn);
  for (int64_t i = 0; i < nn; ++i) {
    if (!touched[i]) CFEM_THROW(-1, "mesh has a node that belongs to no cell (node " + std::to_string(i) + ")");
    g.is_bnd_user[i] = acc[i] != 0;
  }
}

// ---- stage B: rank's part (contiguous range [lo, hi) of the internal order) plus one ghost layer -----------
// Everything heavy (cell list, adjacency, CSR pattern) is built for the part only: O(N / world) memory and time on top
// of the light global stage.
static void build_part(const GlobalOrder& g, const double* x, int xdim, const void* cells_in, int idx_bytes, int rank,
                       int world, HostMesh& hm, std::vector<int32_t>& lv2c, std::vector<uint8_t>& lv2k) {
  const int64_t nn = g.nn, ncg = g.nc;
  hm.rank = rank;
  hm.world = world;
  hm.nn_global = nn;
  hm.part_off = g.part_off;
  const int64_t lo = hm.part_off[rank], hi = hm.part_off[rank + 1];
  const int64_t no = hi - lo;
  hm.n_owned = no;
  auto owner_of = [&](int32_t node) {
    return (int)(std::upper_bound(hm.part_off.begin(), hm.part_off.end(), (int64_t)node) - hm.part_off.begin()) - 1;
  };

  // local cells: every cell touching an owned node, ordered by (smallest internal vertex, caller index)
  struct LCell { int32_t vmin, user, v[3]; };
  std::vector<LCell> lc;
  {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    std::vector<std::vector<LCell>> found(nthreads);   // per thread; the sort below fixes the order
#pragma omp parallel
    {
      int tid = 0;
#ifdef _OPENMP
      tid = omp_get_thread_num();
#endif
      std::vector<LCell>& mine_cells = found[tid];
#pragma omp for schedule(static)
      for (int64_t c = 0; c < ncg; ++c) {
        LCell e;
        bool mine = false;
        for (int k = 0; k < 3; ++k) {
          const int64_t u = idx_bytes == 4 ? cell_vertex<int32_t>(cells_in, c, k) : cell_vertex<int64_t>(cells_in, c, k);
          e.v[k] = g.u2n[u];
          mine = mine || (e.v[k] >= lo && e.v[k] < hi);
        }
        if (!mine) continue;
        e.vmin = std::min(e.v[0], std::min(e.v[1], e.v[2]));
        e.user = (int32_t)c;
        mine_cells.push_back(e);
      }
    }
    size_t total = 0;
    for (auto& f : found) total += f.size();
    lc.reserve(total);
    for (auto& f : found) { lc.insert(lc.end(), f.begin(), f.end()); std::vector<LCell>().swap(f); }
  }
  auto cell_less = [](const LCell& a, const LCell& b) { return a.vmin != b.vmin ? a.vmin < b.vmin : a.user < b.user; };
#ifdef _OPENMP
  __gnu_parallel::sort(lc.begin(), lc.end(), cell_less);
#else
  std::sort(lc.begin(), lc.end(), cell_less);
#endif
  const int64_t nc = (int64_t)lc.size();
  hm.nc = nc;

  // ghosts: vertices of local cells outside [lo, hi), ascending internal id (=> grouped by owner)
  std::vector<int32_t> ghosts;
  if (world > 1) {
    for (int64_t c = 0; c < nc; ++c)
      for (int k = 0; k < 3; ++k)
        if (lc[c].v[k] < lo || lc[c].v[k] >= hi) ghosts.push_back(lc[c].v[k]);
    sort_pairs(ghosts.begin(), ghosts.end());
    ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  }
  const int64_t ng = (int64_t)ghosts.size();
  hm.ghost_global = ghosts;
  const int64_t nl = no + ng;
  hm.nn = nl;
  auto to_local = [&](int32_t v) -> int32_t {
    if (v >= lo && v < hi) return (int32_t)(v - lo);
    return (int32_t)(no + (std::lower_bound(ghosts.begin(), ghosts.end(), v) - ghosts.begin()));
  };
  auto to_global = [&](int64_t l) -> int32_t { return l < no ? (int32_t)(lo + l) : ghosts[l - no]; };

  hm.n2u.resize(nl);
  hm.xy.resize(2 * nl);
  hm.is_bnd.resize(nl);
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < nl; ++l) {
    const int64_t u = g.n2u[to_global(l)];
    hm.n2u[l] = (int32_t)u;
    hm.xy[2 * l] = x[u * xdim];
    hm.xy[2 * l + 1] = x[u * xdim + 1];
    hm.is_bnd[l] = g.is_bnd_user[u];
  }
  hm.u2n = g.u2n;  // user -> GLOBAL internal id (only meaningful together with part_off)
  hm.cells.resize(3 * nc);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c)
    for (int k = 0; k < 3; ++k) hm.cells[3 * c + k] = to_local(lc[c].v[k]);

  hm.cell_user.resize(nc);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) hm.cell_user[c] = (lc[c].vmin >= lo && lc[c].vmin < hi) ? lc[c].user : ~lc[c].user;

  // vertex -> (cell, k) of the owned nodes, cells ascending
  hm.v2c_ptr.assign(no + 1, 0);
  for (int64_t e = 0; e < 3 * nc; ++e)
    if (hm.cells[e] < no) hm.v2c_ptr[hm.cells[e] + 1]++;
  for (int64_t i = 0; i < no; ++i) hm.v2c_ptr[i + 1] += hm.v2c_ptr[i];
  const int64_t ne = hm.v2c_ptr[no];
  lv2c.resize(ne);
  lv2k.resize(ne);
  {
    std::vector<int32_t> fill(hm.v2c_ptr.begin(), hm.v2c_ptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
      for (int k = 0; k < 3; ++k) {
        const int32_t v = hm.cells[3 * c + k];
        if (v >= no) continue;
        const int64_t p = fill[v]++;
        lv2c[p] = (int32_t)c;
        lv2k[p] = (uint8_t)k;
      }
  }

  // owned rows of the P1 pattern (== node patches), columns in ascending GLOBAL order so that every rank count
  // sums a row in the same order
  hm.rowptr.assign(no + 1, 0);
  int max_row = 0;
  bool row_overflow = false;
#pragma omp parallel for schedule(static) reduction(max : max_row) reduction(|| : row_overflow)
  for (int64_t i = 0; i < no; ++i) {
    int32_t buf[3 * 64];
    const int deg = hm.v2c_ptr[i + 1] - hm.v2c_ptr[i];
    if (deg > 64) { row_overflow = true; continue; }
    int m = 0;
    for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e)
      for (int k = 0; k < 3; ++k) buf[m++] = lc[lv2c[e]].v[k];
    std::sort(buf, buf + m);
    const int len = (int)(std::unique(buf, buf + m) - buf);
    hm.rowptr[i + 1] = len;
    max_row = std::max(max_row, len);
  }
  if (row_overflow || max_row > kMaxRow)
    CFEM_THROW(-1, "a node has more than " + std::to_string(kMaxRow - 1) + " neighbours; unsupported mesh");
  for (int64_t i = 0; i < no; ++i) hm.rowptr[i + 1] += hm.rowptr[i];
  hm.max_row = max_row;
  hm.nnz = hm.rowptr[no];
  hm.colidx.resize(hm.nnz);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < no; ++i) {
    int32_t buf[3 * 64];
    int m = 0;
    for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e)
      for (int k = 0; k < 3; ++k) buf[m++] = lc[lv2c[e]].v[k];
    std::sort(buf, buf + m);
    const int len = (int)(std::unique(buf, buf + m) - buf);
    int32_t* row = &hm.colidx[hm.rowptr[i]];
    for (int a = 0; a < len; ++a) row[a] = to_local(buf[a]);
  }

  // per owned node the incident cell with the highest caller index: a per-cell loop that writes a cell
  // value to its three dofs (Code/Linear_advection/RV_cell.py:190-192) leaves exactly that cell's value
  hm.last_cell.resize(no);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < no; ++i) {
    int32_t best = -1, best_user = -1;
    for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e) {
      const int32_t cu = lc[lv2c[e]].user;
      if (cu > best_user) { best_user = cu; best = lv2c[e]; }
    }
    hm.last_cell[i] = best;
  }

  // boundary dofs in caller numbering: global list (every rank reports the same set)
  hm.bnd_user_sorted.clear();
  for (int64_t u = 0; u < nn; ++u)
    if (g.is_bnd_user[u]) hm.bnd_user_sorted.push_back((int32_t)u);

  // halo exchange lists
  hm.peer_rank.clear();
  hm.send_ptr.assign(1, 0);
  hm.send_idx.clear();
  hm.recv_off.clear();
  hm.recv_cnt.clear();
  if (world > 1) {
    std::vector<std::vector<int32_t>> send(world);
    for (int64_t c = 0; c < nc; ++c) {
      int own[3];
      for (int k = 0; k < 3; ++k) own[k] = owner_of(lc[c].v[k]);
      for (int a = 0; a < 3; ++a)
        if (own[a] == rank)
          for (int b = 0; b < 3; ++b)
            if (own[b] != rank) send[own[b]].push_back(lc[c].v[a]);
    }
    std::vector<int64_t> rcnt(world, 0), roff(world, 0);
    for (int64_t k = 0; k < ng; ++k) rcnt[owner_of(ghosts[k])]++;
    int64_t accum = 0;
    for (int r = 0; r < world; ++r) { roff[r] = accum; accum += rcnt[r]; }
    for (int r = 0; r < world; ++r) {
      auto& sl = send[r];
      std::sort(sl.begin(), sl.end());
      sl.erase(std::unique(sl.begin(), sl.end()), sl.end());
      if (sl.empty() && rcnt[r] == 0) continue;
      hm.peer_rank.push_back(r);
      for (int32_t v : sl) hm.send_idx.push_back((int32_t)(v - lo));
      hm.send_ptr.push_back((int32_t)hm.send_idx.size());
      hm.recv_off.push_back((int32_t)(no + roff[r]));
      hm.recv_cnt.push_back((int32_t)rcnt[r]);
    }
  }
}

// Tiles over the owned nodes + packed codes.
static void build_tiles(HostMesh& hm, const std::vector<int32_t>& v2c, const std::vector<uint8_t>& v2k) {
  const int64_t no = hm.n_owned, nc = hm.nc;
  hm.v2c_code.resize(v2c.size());
  hm.tile_node.assign(1, 0);
  hm.tile_cellptr.assign(1, 0);
  hm.tile_cells.clear();
  hm.max_tile_cells = hm.max_tile_nnz = 0;
  std::vector<int32_t> stamp(nc, -1), loc(nc, 0);
  std::vector<int32_t> cur;
  cur.reserve(kTileCellCap);
  int64_t tile_begin = 0;
  int tile_nnz = 0, tile_id = 0;
  auto close_tile = [&](int64_t end_node) {
    std::sort(cur.begin(), cur.end());
    for (size_t a = 0; a < cur.size(); ++a) loc[cur[a]] = (int32_t)a;
    for (int64_t i = tile_begin; i < end_node; ++i) {
      const int32_t* row = &hm.colidx[hm.rowptr[i]];
      const int len = hm.rowptr[i + 1] - hm.rowptr[i];
      for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e) {
        const int64_t c = v2c[e];
        uint32_t code = (uint32_t)loc[c] | ((uint32_t)v2k[e] << kCodeCellBits);
        for (int j = 0; j < 3; ++j) {
          const int32_t col = hm.cells[3 * c + j];
          int pos = 0;
          while (pos < len && row[pos] != col) ++pos;  // rows are short; order may not be sorted locally
          code |= (uint32_t)pos << (kCodeCellBits + 2 + 5 * j);
        }
        hm.v2c_code[e] = code;
      }
    }
    hm.max_tile_cells = std::max(hm.max_tile_cells, (int)cur.size());
    hm.max_tile_nnz = std::max(hm.max_tile_nnz, tile_nnz);
    hm.tile_cells.insert(hm.tile_cells.end(), cur.begin(), cur.end());
    hm.tile_node.push_back((int32_t)end_node);
    hm.tile_cellptr.push_back((int32_t)hm.tile_cells.size());
    cur.clear();
    tile_begin = end_node;
    tile_nnz = 0;
    ++tile_id;
  };
  for (int64_t i = 0; i < no; ++i) {
    const int len = hm.rowptr[i + 1] - hm.rowptr[i];
    int fresh = 0;
    for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e)
      if (stamp[v2c[e]] != tile_id) ++fresh;
    const bool full = (i - tile_begin) >= kTileNodes || (int)cur.size() + fresh > kTileCellCap ||
                      tile_nnz + len > kTileNnzCap;
    if (full && i > tile_begin) close_tile(i);
    for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e)
      if (stamp[v2c[e]] != tile_id) { stamp[v2c[e]] = tile_id; cur.push_back(v2c[e]); }
    tile_nnz += len;
    if ((int)cur.size() > kTileCellCap) CFEM_THROW(-1, "node valence exceeds tile capacity");
  }
  close_tile(no);
  // T16 format: per tile the distinct external columns (ascending) and 16-bit tile-local column indices
  {
    const int ntl = (int)hm.tile_node.size() - 1;
    hm.lc16.resize(hm.nnz);
    hm.tile_extptr.assign(ntl + 1, 0);
    std::vector<std::vector<int32_t>> ext(ntl);
    int max_ext = 0;
    bool overflow = false;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : max_ext) reduction(|| : overflow)
    for (int t = 0; t < ntl; ++t) {
      const int32_t a = hm.tile_node[t], b = hm.tile_node[t + 1];
      const int32_t p0 = hm.rowptr[a], p1 = hm.rowptr[b];
      std::vector<int32_t>& e = ext[t];
      for (int32_t p = p0; p < p1; ++p) {
        const int32_t col = hm.colidx[p];
        if (col < a || col >= b) e.push_back(col);
      }
      std::sort(e.begin(), e.end());
      e.erase(std::unique(e.begin(), e.end()), e.end());
      if ((int)e.size() + kTileNodes > 65535) overflow = true;
      max_ext = std::max(max_ext, (int)e.size());
      for (int32_t p = p0; p < p1; ++p) {
        const int32_t col = hm.colidx[p];
        if (col >= a && col < b) hm.lc16[p] = (uint16_t)(col - a);
        else hm.lc16[p] = (uint16_t)(kTileNodes + (std::lower_bound(e.begin(), e.end(), col) - e.begin()));
      }
    }
    if (overflow) CFEM_THROW(-1, "tile has too many external columns for 16-bit local indices");
    hm.max_tile_ext = max_ext;
    for (int t = 0; t < ntl; ++t) hm.tile_extptr[t + 1] = hm.tile_extptr[t] + (int32_t)ext[t].size();
    hm.tile_ext.resize(hm.tile_extptr[ntl]);
    for (int t = 0; t < ntl; ++t) std::copy(ext[t].begin(), ext[t].end(), hm.tile_ext.begin() + hm.tile_extptr[t]);
  }
  // interior tiles (no ghost column in any row) first: SpMV-type kernels start on them while the
  // neighbours' halo values are still in flight
  const int nt = (int)hm.tile_node.size() - 1;
  std::vector<int32_t> interior, boundary;
  for (int t = 0; t < nt; ++t) {
    bool needs_ghost = false;
    for (int p = hm.rowptr[hm.tile_node[t]]; p < hm.rowptr[hm.tile_node[t + 1]] && !needs_ghost; ++p)
      needs_ghost = hm.colidx[p] >= no;
    (needs_ghost ? boundary : interior).push_back(t);
  }
  hm.n_interior_tiles = (int)interior.size();
  hm.tile_order = interior;
  hm.tile_order.insert(hm.tile_order.end(), boundary.begin(), boundary.end());
}

int32_t user_to_local(const HostMesh& hm, int64_t user_dof) {
  if (user_dof < 0 || user_dof >= hm.nn_global) return -1;
  const int64_t g = hm.u2n[user_dof];
  const int64_t lo = hm.part_off[hm.rank], hi = hm.part_off[hm.rank + 1];
  if (g >= lo && g < hi) return (int32_t)(g - lo);
  auto it = std::lower_bound(hm.ghost_global.begin(), hm.ghost_global.end(), (int32_t)g);
  if (it != hm.ghost_global.end() && *it == (int32_t)g) return (int32_t)(hm.n_owned + (it - hm.ghost_global.begin()));
  return -1;
}

void analyse_mesh(HostMesh& hm, int64_t nn, int64_t nc, const double* x, int xdim, const void* cells,
                  int idx_bytes, int order, int rank, int world, const int32_t* node_part) {
  if (world < 1 || rank < 0 || rank >= world) CFEM_THROW(-1, "bad rank / world size");
  std::vector<int32_t> lv2c;
  std::vector<uint8_t> lv2k;
  {
    GlobalOrder g;
    global_order(g, nn, nc, x, xdim, cells, idx_bytes, order, world, node_part);
    build_part(g, x, xdim, cells, idx_bytes, rank, world, hm, lv2c, lv2k);
  }  // the global order is released before tiling (u2n lives on in hm)
  build_tiles(hm, lv2c, lv2k);
}

}  // namespace cfem
