// Host-side, once-per-mesh analysis: Hilbert ordering, vertex->cell adjacency,
// P1 CSR pattern (== node patches, reference Code/Utils/SI.py:12-28), boundary
// dofs (reference Code/KPP/KPP_exact.py:85-89), and the assembly tiles with
// their packed (node, cell) codes.  Everything is O(N) apart from one sort.
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

void analyse_mesh(HostMesh& hm, int64_t nn, int64_t nc, const double* x, int xdim,
                  const void* cells_in, int idx_bytes, int order) {
  if (nn <= 0 || nc <= 0) CFEM_THROW(-1, "empty mesh");
  if (nn >= (int64_t)1 << 31 || 3 * nc >= (int64_t)1 << 31) CFEM_THROW(-1, "mesh too large for int32 indices");
  if (xdim != 2 && xdim != 3) CFEM_THROW(-1, "xdim must be 2 or 3");
  if (idx_bytes != 4 && idx_bytes != 8) CFEM_THROW(-1, "cell_index_bytes must be 4 or 8");
  hm.nn = nn;
  hm.nc = nc;

  // ---- 1. node ordering ----------------------------------------------------
  hm.n2u.resize(nn);
  hm.u2n.resize(nn);
  if (order == CFEM_ORDER_NATURAL) {
    std::iota(hm.n2u.begin(), hm.n2u.end(), 0);
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
    std::vector<std::pair<uint64_t, int32_t>> keys(nn);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) {
      const uint32_t qx = (uint32_t)((x[i * xdim] - xmin) * scale);
      const uint32_t qy = (uint32_t)((x[i * xdim + 1] - ymin) * scale);
      keys[i] = {hilbert_key(qx, qy, bits), (int32_t)i};
    }
    sort_pairs(keys.begin(), keys.end());
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) hm.n2u[i] = keys[i].second;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) hm.u2n[hm.n2u[i]] = (int32_t)i;
  hm.xy.resize(2 * nn);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    const int64_t u = hm.n2u[i];
    hm.xy[2 * i] = x[u * xdim];
    hm.xy[2 * i + 1] = x[u * xdim + 1];
  }

  // ---- 2. cells -> internal ids, ordered by their smallest vertex -----------
  std::vector<int32_t> ctmp(3 * nc);
  {
    bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
    for (int64_t c = 0; c < nc; ++c) {
      for (int k = 0; k < 3; ++k) {
        int64_t v = idx_bytes == 4 ? (int64_t)((const int32_t*)cells_in)[3 * c + k]
                                   : (int64_t)((const int64_t*)cells_in)[3 * c + k];
        if (v < 0 || v >= nn) { bad = true; v = 0; }
        ctmp[3 * c + k] = hm.u2n[v];
      }
      if (ctmp[3 * c] == ctmp[3 * c + 1] || ctmp[3 * c] == ctmp[3 * c + 2] || ctmp[3 * c + 1] == ctmp[3 * c + 2])
        bad = true;
    }
    if (bad) CFEM_THROW(-1, "cell connectivity has an out-of-range or repeated vertex");
  }
  hm.cells.resize(3 * nc);
  {
    // counting sort by min vertex (stable -> deterministic)
    std::vector<int32_t> cnt(nn + 1, 0);
    for (int64_t c = 0; c < nc; ++c)
      cnt[std::min(ctmp[3 * c], std::min(ctmp[3 * c + 1], ctmp[3 * c + 2])) + 1]++;
    for (int64_t i = 0; i < nn; ++i) cnt[i + 1] += cnt[i];
    for (int64_t c = 0; c < nc; ++c) {
      const int32_t m = std::min(ctmp[3 * c], std::min(ctmp[3 * c + 1], ctmp[3 * c + 2]));
      const int64_t p = cnt[m]++;
      hm.cells[3 * p] = ctmp[3 * c];
      hm.cells[3 * p + 1] = ctmp[3 * c + 1];
      hm.cells[3 * p + 2] = ctmp[3 * c + 2];
    }
  }
  ctmp.clear();
  ctmp.shrink_to_fit();

  // ---- 3. vertex -> (cell, k), cells ascending ------------------------------
  hm.v2c_ptr.assign(nn + 1, 0);
  for (int64_t e = 0; e < 3 * nc; ++e) hm.v2c_ptr[hm.cells[e] + 1]++;
  for (int64_t i = 0; i < nn; ++i) {
    if (hm.v2c_ptr[i + 1] == 0) CFEM_THROW(-1, "mesh has a node that belongs to no cell (node " + std::to_string(hm.n2u[i]) + ")");
    hm.v2c_ptr[i + 1] += hm.v2c_ptr[i];
  }
  std::vector<int32_t> v2c(3 * nc);  // entry = 4*cell... stored as cell*4+k would overflow; keep two arrays
  std::vector<uint8_t> v2k(3 * nc);
  {
    std::vector<int32_t> fill(hm.v2c_ptr.begin(), hm.v2c_ptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
      for (int k = 0; k < 3; ++k) {
        const int64_t p = fill[hm.cells[3 * c + k]]++;
        v2c[p] = (int32_t)c;
        v2k[p] = (uint8_t)k;
      }
  }

  // ---- 4. CSR pattern (sorted rows) + boundary flags ------------------------
  hm.rowptr.assign(nn + 1, 0);
  int max_row = 0;
  bool row_overflow = false;
#pragma omp parallel for schedule(static) reduction(max : max_row) reduction(|| : row_overflow)
  for (int64_t i = 0; i < nn; ++i) {
    int32_t buf[3 * 64];
    const int deg = hm.v2c_ptr[i + 1] - hm.v2c_ptr[i];
    if (deg > 64) { row_overflow = true; continue; }
    int m = 0;
    for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e)
      for (int k = 0; k < 3; ++k) buf[m++] = hm.cells[3 * (int64_t)v2c[e] + k];
    std::sort(buf, buf + m);
    const int len = (int)(std::unique(buf, buf + m) - buf);
    hm.rowptr[i + 1] = len;
    max_row = std::max(max_row, len);
  }
  if (row_overflow || max_row > kMaxRow)
    CFEM_THROW(-1, "a node has more than " + std::to_string(kMaxRow - 1) + " neighbours; unsupported mesh");
  for (int64_t i = 0; i < nn; ++i) hm.rowptr[i + 1] += hm.rowptr[i];
  hm.nnz = hm.rowptr[nn];
  hm.max_row = max_row;
  hm.colidx.resize(hm.nnz);
  hm.is_bnd.assign(nn, 0);
  std::vector<uint8_t> bnd_edge_flag(hm.nnz, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) {
    int32_t buf[3 * 64];
    int m = 0;
    for (int e = hm.v2c_ptr[i]; e < hm.v2c_ptr[i + 1]; ++e)
      for (int k = 0; k < 3; ++k) buf[m++] = hm.cells[3 * (int64_t)v2c[e] + k];
    std::sort(buf, buf + m);
    // run lengths: neighbour j shares (count) cells with i; 1 => boundary edge
    int32_t* row = &hm.colidx[hm.rowptr[i]];
    int len = 0;
    for (int a = 0; a < m;) {
      int b = a;
      while (b < m && buf[b] == buf[a]) ++b;
      row[len] = buf[a];
      if (buf[a] != (int32_t)i && (b - a) == 1) bnd_edge_flag[hm.rowptr[i] + len] = 1;
      ++len;
      a = b;
    }
  }
  for (int64_t i = 0; i < nn; ++i)
    for (int p = hm.rowptr[i]; p < hm.rowptr[i + 1]; ++p)
      if (bnd_edge_flag[p]) { hm.is_bnd[i] = 1; hm.is_bnd[hm.colidx[p]] = 1; }
  bnd_edge_flag.clear();
  hm.bnd_user_sorted.clear();
  for (int64_t i = 0; i < nn; ++i)
    if (hm.is_bnd[i]) hm.bnd_user_sorted.push_back(hm.n2u[i]);
  std::sort(hm.bnd_user_sorted.begin(), hm.bnd_user_sorted.end());

  // ---- 5. tiles + packed codes -----------------------------------------------
  hm.v2c_code.resize(3 * nc);
  hm.tile_node.clear();
  hm.tile_cellptr.clear();
  hm.tile_cells.clear();
  hm.tile_node.push_back(0);
  hm.tile_cellptr.push_back(0);
  std::vector<int32_t> stamp(nc, -1), loc(nc, 0);
  std::vector<int32_t> cur;  // cells of the open tile
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
          const int pos = (int)(std::lower_bound(row, row + len, col) - row);
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
  for (int64_t i = 0; i < nn; ++i) {
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
  close_tile(nn);
}

}  // namespace cfem
