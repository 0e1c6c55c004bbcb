// extern "C" surface (include/cfem_b200.h): context life cycle, host/device
// argument staging with the internal<->caller permutation, the single-operator
// entry points, and the time loops.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "device_utils.cuh"
#include "launch.h"

using namespace cfem;

namespace cfem {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

ProfScope::ProfScope(cfem_ctx* c_, int cat, int launches) : c(c_), active(false) {
  Profiler& p = c->prof;
  if (!p.on) return;
  if (p.depth++ > 0) return;  // nested scopes are charged to the outermost category
  if (p.used + 2 > p.ev.size()) return;
  active = true;
  idx = p.used;
  p.used += 2;
  p.cat[idx / 2] = cat;
  p.weight[idx / 2] = launches;
  cudaEventRecord(p.ev[idx], c->stream);
}
ProfScope::~ProfScope() {
  Profiler& p = c->prof;
  if (p.on && p.depth > 0) --p.depth;
  if (!active) return;
  cudaEventRecord(p.ev[idx + 1], c->stream);
}
}  // namespace cfem

#define API_BEGIN try {
#define API_END                                                   \
  return 0;                                                       \
  }                                                               \
  catch (const cfem::Error& e) { cfem::set_error(e.msg); return e.code; } \
  catch (const std::exception& e) { cfem::set_error(e.what()); return -9; }

namespace {

template <class T>
T* dalloc(cfem_ctx* c, int64_t count) {
  void* p = nullptr;
  const size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(T);
  CUDA_OK(cudaMalloc(&p, bytes));
  c->allocs.push_back(p);
  c->bytes += (int64_t)bytes;
  return (T*)p;
}
template <class T>
T* upload(cfem_ctx* c, const std::vector<T>& v) {
  T* d = dalloc<T>(c, (int64_t)v.size());
  if (!v.empty()) CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// caller array (host or device, caller numbering, GLOBAL size) -> local-order device vector
// (owned + ghost entries, so imported fields arrive with valid ghosts)
void import_vec(cfem_ctx* c, const double* user, double* dst, int stage_slot = 0) {
  const int64_t n = c->dm.nn;
  if (is_device_ptr(user)) { launch_gather(c, user, c->d_n2u, dst, n); return; }
  if (c->world > 1) {
    // distributed: gather this rank's entries on the host, ship only those
    CUDA_OK(cudaStreamSynchronize(c->stream));  // h_stage may still feed an earlier copy
    const int32_t* n2u = c->hm.n2u.data();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) c->h_stage[i] = user[n2u[i]];
    CUDA_OK(cudaMemcpyAsync(dst, c->h_stage, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return;
  }
  CUDA_OK(cudaMemcpyAsync(c->stage[stage_slot], user, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_gather(c, c->stage[stage_slot], c->d_n2u, dst, n);
}
void import_vec2(cfem_ctx* c, const double* user, double2* dst) {
  const int64_t n = c->dm.nn;
  if (is_device_ptr(user)) { launch_gather2(c, (const double2*)user, c->d_n2u, dst, n); return; }
  if (c->world > 1) {
    CUDA_OK(cudaStreamSynchronize(c->stream));
    const int32_t* n2u = c->hm.n2u.data();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      c->h_stage[2 * i] = user[2 * (int64_t)n2u[i]];
      c->h_stage[2 * i + 1] = user[2 * (int64_t)n2u[i] + 1];
    }
    CUDA_OK(cudaMemcpyAsync(dst, c->h_stage, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return;
  }
  // stage[2] and stage[3] are contiguous (allocated as one block)
  CUDA_OK(cudaMemcpyAsync(c->stage[2], user, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_gather2(c, (const double2*)c->stage[2], c->d_n2u, dst, n);
}
// local-order device vector -> caller array (synchronous for host destinations).
// Distributed contexts write the OWNED entries only (the caller's array is global-sized).
void export_vec(cfem_ctx* c, const double* internal, double* user) {
  if (c->world > 1) {
    const int64_t no = c->dm.no;
    if (is_device_ptr(user)) { launch_scatter(c, internal, c->d_n2u, user, no); return; }
    CUDA_OK(cudaMemcpyAsync(c->h_stage, internal, no * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    const int32_t* n2u = c->hm.n2u.data();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < no; ++i) user[n2u[i]] = c->h_stage[i];
    return;
  }
  const int64_t n = c->dm.nn;
  if (is_device_ptr(user)) {
    launch_gather(c, internal, c->d_u2n, user, n);
  } else {
    launch_gather(c, internal, c->d_u2n, c->stage[1], n);
    CUDA_OK(cudaMemcpyAsync(user, c->stage[1], n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
}

// Dirichlet values in caller's bc order -> nodal g (internal), zero if null
void import_bc(cfem_ctx* c, const double* bc_values) {
  if (!bc_values) { launch_bc_values(c, CFEM_BC_CONSTANT, 0.0, 0.0, nullptr, c->g); return; }
  const double* src = bc_values;
  if (!is_device_ptr(bc_values)) {
    CUDA_OK(cudaMemcpyAsync(c->stage[0], bc_values, c->nbc_user * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    src = c->stage[0];
  }
  launch_bc_values(c, CFEM_BC_USER, 0.0, 0.0, src, c->g);
}

void apply_dirichlet(cfem_ctx* c, const int32_t* dofs, int64_t n) {
  const int64_t nl = c->dm.nn;
  std::vector<uint8_t> flag(nl, 0);
  std::vector<int32_t> nodes, pos;
  c->bc_user.assign(dofs, dofs + n);
  for (int64_t j = 0; j < n; ++j) {
    if (dofs[j] < 0 || dofs[j] >= c->hm.nn_global) CFEM_THROW(-1, "Dirichlet dof out of range");
    const int32_t l = user_to_local(c->hm, dofs[j]);
    if (l < 0) continue;  // held by another rank
    nodes.push_back(l);
    pos.push_back((int32_t)j);
    flag[l] = 1;
  }
  if (n > 2 * nl && c->world == 1) CFEM_THROW(-1, "Dirichlet list longer than the mesh");
  CUDA_OK(cudaMemcpy(c->d_is_bc, flag.data(), nl, cudaMemcpyHostToDevice));
  if (!nodes.empty()) {
    CUDA_OK(cudaMemcpy(c->d_bc_nodes, nodes.data(), nodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->d_bc_pos, pos.data(), pos.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  c->nbc = (int64_t)nodes.size();
  c->nbc_user = n;
  launch_mass(c, c->mat[CFEM_MAT_MASS_BC], true);
  launch_fill(c, c->g, 0.0, nl);
}

Matrix& get_matrix(cfem_ctx* c, int which) {
  if (which < 0 || which > 3) CFEM_THROW(-1, "unknown matrix id");
  if (!c->mat[which].valid) CFEM_THROW(-1, "matrix has not been assembled yet");
  return c->mat[which];
}

void build_user_csr(cfem_ctx* c) {
  if (c->world > 1) CFEM_THROW(-1, "the global CSR pattern / matrix export is not available in a distributed context");
  if (!c->u_rowptr.empty()) return;
  const HostMesh& hm = c->hm;
  const int64_t nn = hm.nn;
  c->u_rowptr.assign(nn + 1, 0);
  for (int64_t u = 0; u < nn; ++u) {
    const int32_t i = hm.u2n[u];  // world == 1: global == local
    c->u_rowptr[u + 1] = c->u_rowptr[u] + (hm.rowptr[i + 1] - hm.rowptr[i]);
  }
  c->u_colidx.resize(hm.nnz);
  c->u_slot.resize(hm.nnz);
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < nn; ++u) {
    const int32_t i = hm.u2n[u];
    const int len = hm.rowptr[i + 1] - hm.rowptr[i];
    std::pair<int32_t, int32_t> tmp[kMaxRow];
    for (int k = 0; k < len; ++k) tmp[k] = {hm.n2u[hm.colidx[hm.rowptr[i] + k]], hm.rowptr[i] + k};
    std::sort(tmp, tmp + len);
    for (int k = 0; k < len; ++k) {
      c->u_colidx[c->u_rowptr[u] + k] = tmp[k].first;
      c->u_slot[c->u_rowptr[u] + k] = tmp[k].second;
    }
  }
}

// mass-matrix solves: Chebyshev unless the caller asks for PCG (mass_solver == 100 + CFEM_SOLVER_PCG)
SolveResult mass_solve(cfem_ctx* c, int mass_solver, const Matrix& M, const double* b, double* x, double rtol,
                       int max_it, int* predict) {
  if (mass_solver == 100 + CFEM_SOLVER_PCG) return pcg(c, M, b, x, rtol, 0.0, max_it, predict);
  return chebyshev_mass(c, M, b, x, rtol, max_it, predict);
}

SolveResult run_solver(cfem_ctx* c, int solver, const Matrix& A, const double* b, double* x, double rtol,
                       double atol, int max_it, int* predict) {
  switch (solver) {
    case CFEM_SOLVER_PCG: return pcg(c, A, b, x, rtol, atol, max_it, predict);
    case CFEM_SOLVER_BICGSTAB: return bicgstab(c, A, b, x, rtol, atol, max_it, predict);
    case CFEM_SOLVER_GMRES: return gmres(c, A, b, x, rtol, atol, max_it, predict);
    case CFEM_SOLVER_CHEBYSHEV: return chebyshev_mass(c, A, b, x, rtol, max_it, predict);
    default: CFEM_THROW(-1, "unknown solver id");
  }
}

}  // namespace

// =============================================================================
extern "C" {

const char* cfem_last_error(void) { return g_error.c_str(); }
int cfem_version(void) { return 100; }
int cfem_struct_size(int which) { return which == 0 ? (int)sizeof(cfem_step_params) : (which == 1 ? (int)sizeof(cfem_step_stats) : -1); }
int cfem_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static int create_impl(cfem_ctx** out, int device, int rank, int world, const void* nccl_id, int64_t n_nodes,
                       int64_t n_cells, const double* x, int xdim, const void* cells, int cell_index_bytes,
                       int order, const int32_t* node_part = nullptr) {
  cfem_ctx* c = nullptr;
  API_BEGIN
  if (!out || !x || !cells) CFEM_THROW(-1, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    CFEM_THROW(-2, "no CUDA device: cfem_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) CFEM_THROW(-1, "device index out of range");
  c = new cfem_ctx();
  c->device = device;
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  if (world > 1 && !nccl_id) CFEM_THROW(-1, "distributed context needs the NCCL unique id");
  analyse_mesh(c->hm, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, order, rank, world, node_part);
  comm_init(c, rank, world, nccl_id);
  HostMesh& hm = c->hm;
  const int64_t nn = hm.nn;
  DevMesh& dm = c->dm;
  dm.nn = nn; dm.nc = hm.nc; dm.nnz = hm.nnz; dm.no = hm.n_owned; dm.nn_global = hm.nn_global;
  dm.ntiles = (int)hm.tile_node.size() - 1;
  dm.xy = (const double2*)upload(c, hm.xy);
  dm.cells = upload(c, hm.cells);
  { // hot block: MASS_BC values | rowptr | lc16 | extptr | ext | SYSTEM values | colidx (one allocation: see cfem_ctx::hot_base)
    std::vector<int32_t> rp(hm.rowptr), ci(hm.colidx);  // padded copies for the aligned bulk-copy windows
    rp.resize(rp.size() + 8, rp.back()); ci.resize(ci.size() + 8, 0);
    std::vector<uint16_t> lc(hm.lc16);
    lc.resize(lc.size() + 16, 0);
    auto up256 = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t vbytes = up256((size_t)(hm.nnz + 8) * sizeof(double));
    size_t* off = c->hot_off;
    off[0] = 0;
    off[1] = vbytes;
    off[2] = off[1] + up256(rp.size() * sizeof(int32_t));
    off[3] = off[2] + up256(lc.size() * sizeof(uint16_t));
    off[4] = off[3] + up256(hm.tile_extptr.size() * sizeof(int32_t));
    off[5] = off[4] + up256((hm.tile_ext.size() + 8) * sizeof(int32_t));
    off[6] = off[5] + vbytes;
    off[7] = off[6] + up256(ci.size() * sizeof(int32_t));
    c->hot_base = (char*)dalloc<char>(c, (int64_t)off[7]);
    auto put = [&](size_t o, const void* src, size_t bytes) {
      if (bytes) CUDA_OK(cudaMemcpy(c->hot_base + o, src, bytes, cudaMemcpyHostToDevice));
    };
    put(off[1], rp.data(), rp.size() * sizeof(int32_t));
    put(off[2], lc.data(), lc.size() * sizeof(uint16_t));
    put(off[3], hm.tile_extptr.data(), hm.tile_extptr.size() * sizeof(int32_t));
    put(off[4], hm.tile_ext.data(), hm.tile_ext.size() * sizeof(int32_t));
    put(off[6], ci.data(), ci.size() * sizeof(int32_t));
    dm.rowptr = (const int32_t*)(c->hot_base + off[1]);
    dm.lc16 = (const uint16_t*)(c->hot_base + off[2]);
    dm.tile_extptr = (const int32_t*)(c->hot_base + off[3]);
    dm.tile_ext = (const int32_t*)(c->hot_base + off[4]);
    dm.colidx = (const int32_t*)(c->hot_base + off[6]);
    dm.ext_cap = (hm.max_tile_ext + 31) / 32 * 32;
    std::vector<uint16_t>().swap(hm.lc16);
    std::vector<int32_t>().swap(hm.tile_ext);
    // persisting-L2 set-aside: what the device allows unless CFEM_L2_SETASIDE_MB says otherwise.  It is a DEVICE-wide
    // limit that shrinks the L2 every other access sees, so it is only switched on while a solve whose matrix fits is
    // running (linalg.cu: l2_prefer / l2_release) -- a context whose matrices are too large, or the Euler path, must
    // not pay for it (Euler 8 M cells: 81 -> 60 ms per step).
    const char* l2off = getenv("CFEM_L2PERSIST");
    if (!(l2off && std::string(l2off) == "0") && prop.persistingL2CacheMaxSize > 0) {
      size_t want = (size_t)prop.persistingL2CacheMaxSize;
      if (const char* mb = getenv("CFEM_L2_SETASIDE_MB")) want = std::min(want, (size_t)atol(mb) << 20);
      c->l2_setaside = want;
      c->l2_max_window = (size_t)prop.accessPolicyMaxWindowSize;
    }
  }
  dm.v2c_ptr = upload(c, hm.v2c_ptr);
  dm.v2c_code = upload(c, hm.v2c_code);
  dm.tile_node = upload(c, hm.tile_node);
  dm.tile_cellptr = upload(c, hm.tile_cellptr);
  dm.tile_cells = upload(c, hm.tile_cells);
  dm.tile_order = upload(c, hm.tile_order);
  dm.n_interior = hm.n_interior_tiles;
  dm.last_cell = upload(c, hm.last_cell);
  dm.cell_user = upload(c, hm.cell_user);
  c->d_n2u = upload(c, hm.n2u);
  c->d_u2n = world == 1 ? upload(c, hm.u2n) : nullptr;  // user -> local is only a permutation on one GPU
  c->d_send_idx = upload(c, hm.send_idx);
  c->d_sendbuf = dalloc<double>(c, 4 * (int64_t)hm.send_idx.size());  // widest exchange: 4 components
  c->d_is_bnd = upload(c, hm.is_bnd);
  c->d_is_bc = dalloc<uint8_t>(c, nn);
  dm.is_bc = c->d_is_bc;
  c->d_bc_nodes = dalloc<int32_t>(c, nn);
  c->d_bc_pos = dalloc<int32_t>(c, nn);
  // host-side copies that are only needed on the device from here on
  std::vector<double>().swap(hm.xy);
  std::vector<int32_t>().swap(hm.cells);
  std::vector<uint32_t>().swap(hm.v2c_code);
  std::vector<int32_t>().swap(hm.v2c_ptr);
  std::vector<int32_t>().swap(hm.tile_cells);
  if (world > 1) CUDA_OK(cudaMallocHost((void**)&c->h_stage, 4 * nn * sizeof(double)));
  for (int k = 0; k < 4; ++k) {
    // +8: 16-byte aligned bulk-copy windows may overrun the last tile
    if (k == CFEM_MAT_MASS_BC) c->mat[k].vals = (double*)(c->hot_base + c->hot_off[0]);
    else if (k == CFEM_MAT_SYSTEM) c->mat[k].vals = (double*)(c->hot_base + c->hot_off[5]);
    else c->mat[k].vals = dalloc<double>(c, hm.nnz + 8);
    c->mat[k].dinv = dalloc<double>(c, nn);
  }
  double** state[] = {&c->uh, &c->u_n, &c->u_old, &c->u_oo, &c->RH, &c->eps, &c->h, &c->g, &c->fluxn};
  for (double** s : state) { *s = dalloc<double>(c, nn); CUDA_OK(cudaMemsetAsync(*s, 0, nn * sizeof(double), c->stream)); }
  c->w = (double2*)dalloc<double>(c, 2 * nn);
  CUDA_OK(cudaMemsetAsync(c->w, 0, 2 * nn * sizeof(double), c->stream));
  for (int k = 0; k < 10; ++k) c->wk[k] = dalloc<double>(c, nn);
  c->stage[0] = dalloc<double>(c, nn);
  c->stage[1] = dalloc<double>(c, nn);
  c->stage[2] = dalloc<double>(c, 2 * nn);
  c->stage[3] = c->stage[2] + nn;
  c->partials = dalloc<double>(c, 12 * (int64_t)kMaxPartials);   // slots 0..7 named in linalg.cu, 5..8 = the four dots of Ep16BiT
  c->partials12 = dalloc<double>(c, 12 * (int64_t)kMaxPartials);
  c->scalars = dalloc<double>(c, 32);
  c->status = dalloc<int32_t>(c, 8);
  CUDA_OK(cudaMemsetAsync(c->status, 0, 8 * sizeof(int32_t), c->stream));
  CUDA_OK(cudaMemsetAsync(c->scalars, 0, 32 * sizeof(double), c->stream));
  CUDA_OK(cudaMallocHost((void**)&c->h_pinned, (64 + kMaxPartials) * sizeof(double)));
  CUDA_OK(cudaMallocHost((void**)&c->h_status, 8 * sizeof(int32_t)));
  if (hm.max_tile_cells > kTileCellCap || hm.max_tile_nnz > kTileNnzCap) CFEM_THROW(-1, "tile capacity exceeded");
  comm_setup_exchange(c);
  launch_mass(c, c->mat[CFEM_MAT_MASS], false);
  // every boundary dof is a Dirichlet dof by default (reference loops: KPP_exact.py:85-89)
  apply_dirichlet(c, hm.bnd_user_sorted.data(), (int64_t)hm.bnd_user_sorted.size());
  CUDA_OK(cudaStreamSynchronize(c->stream));
  *out = c;
  return 0;
  }
  catch (const cfem::Error& e) { cfem::set_error(e.msg); if (c) cfem_destroy(c); return e.code; }
  catch (const std::exception& e) { cfem::set_error(e.what()); if (c) cfem_destroy(c); return -9; }
}

int cfem_create(cfem_ctx** out, int device, int64_t n_nodes, int64_t n_cells, const double* x, int xdim,
                const void* cells, int cell_index_bytes, int order) {
  return create_impl(out, device, 0, 1, nullptr, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, order);
}

int cfem_nccl_unique_id(void* out128) {
  API_BEGIN
  if (!out128) CFEM_THROW(-1, "null argument");
  comm_unique_id(out128);
  API_END
}

int cfem_create_distributed(cfem_ctx** out, int device, int rank, int world, const void* nccl_id128,
                            int64_t n_nodes, int64_t n_cells, const double* x, int xdim, const void* cells,
                            int cell_index_bytes, int order) {
  return create_impl(out, device, rank, world, nccl_id128, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, order);
}

int cfem_create_partitioned(cfem_ctx** out, int device, int rank, int world, const void* nccl_id128, int64_t n_nodes,
                            int64_t n_cells, const double* x, int xdim, const void* cells, int cell_index_bytes,
                            int order, const int32_t* node_part) {
  return create_impl(out, device, rank, world, nccl_id128, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, order, node_part);
}

int cfem_host_partition(int method, int world, int64_t n_nodes, int64_t n_cells, const double* x, int xdim,
                        const void* cells, int cell_index_bytes, int32_t* part_out) {
  API_BEGIN
  if (!x || !cells || !part_out) CFEM_THROW(-1, "null argument");
  if (method == CFEM_PART_METIS) {
    metis_partition(world, n_nodes, n_cells, cells, cell_index_bytes, part_out);
  } else if (method == CFEM_PART_HILBERT) {
    HostMesh hm;   // rank 0's analysis carries the global order and the range offsets
    analyse_mesh(hm, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, CFEM_ORDER_HILBERT, 0, world);
    for (int64_t u = 0; u < n_nodes; ++u) {
      const int64_t g = hm.u2n[u];
      part_out[u] = (int32_t)(std::upper_bound(hm.part_off.begin(), hm.part_off.end(), g) - hm.part_off.begin()) - 1;
    }
  } else {
    CFEM_THROW(-1, "unknown partition method");
  }
  API_END
}

int64_t cfem_num_owned(const cfem_ctx* c) { return c->dm.no; }
int64_t cfem_num_ghosts(const cfem_ctx* c) { return c->dm.nn - c->dm.no; }
int cfem_comm_stats(const cfem_ctx* c, int64_t* halo_exchanges, int64_t* allreduces, int64_t* halo_doubles_sent) {
  if (halo_exchanges) *halo_exchanges = c->halo_exchanges;
  if (allreduces) *allreduces = c->allreduces;
  if (halo_doubles_sent) *halo_doubles_sent = (int64_t)c->hm.send_idx.size();
  return 0;
}

int cfem_comm_timers(cfem_ctx* c, double out[12], int reset) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  comm_timers(c, out, reset != 0);
  API_END
}

void cfem_destroy(cfem_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  l2_release(c);
  comm_destroy(c);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  for (void* p : c->allocs) cudaFree(p);
  for (cudaEvent_t e : c->prof.ev) cudaEventDestroy(e);
  euler_free(c);
  smooth_plan_free(c);
  persist_plan_free(c);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int cfem_synchronize(cfem_ctx* c) {
  API_BEGIN
  CUDA_OK(cudaStreamSynchronize(c->stream));
  API_END
}

int64_t cfem_num_nodes(const cfem_ctx* c) { return c->dm.nn_global; }
int64_t cfem_num_cells(const cfem_ctx* c) { return c->dm.nc; }
int64_t cfem_num_nonzeros(const cfem_ctx* c) { return c->dm.nnz; }
int64_t cfem_num_boundary(const cfem_ctx* c) { return (int64_t)c->hm.bnd_user_sorted.size(); }
int64_t cfem_num_dirichlet(const cfem_ctx* c) { return c->nbc_user; }
int64_t cfem_num_tiles(const cfem_ctx* c) { return c->dm.ntiles; }
int64_t cfem_device_bytes(const cfem_ctx* c) { return c->bytes; }
int cfem_device_limits(int device, int64_t out[4]) {
  API_BEGIN
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  out[0] = prop.l2CacheSize;
  out[1] = prop.persistingL2CacheMaxSize;
  out[2] = prop.accessPolicyMaxWindowSize;
  out[3] = prop.multiProcessorCount;
  API_END
}

int cfem_get_csr_pattern(cfem_ctx* c, int32_t* rowptr, int32_t* colidx) {
  API_BEGIN
  build_user_csr(c);
  if (rowptr) std::memcpy(rowptr, c->u_rowptr.data(), c->u_rowptr.size() * sizeof(int32_t));
  if (colidx) std::memcpy(colidx, c->u_colidx.data(), c->u_colidx.size() * sizeof(int32_t));
  API_END
}

int cfem_get_boundary_dofs(cfem_ctx* c, int32_t* dofs) {
  API_BEGIN
  std::memcpy(dofs, c->hm.bnd_user_sorted.data(), c->hm.bnd_user_sorted.size() * sizeof(int32_t));
  API_END
}

int cfem_set_dirichlet(cfem_ctx* c, const int32_t* dofs, int64_t n) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (n < 0 || (n > 0 && !dofs)) CFEM_THROW(-1, "bad Dirichlet set");
  apply_dirichlet(c, dofs, n);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  API_END
}

int cfem_get_ordering(cfem_ctx* c, int32_t* n2u) {
  API_BEGIN
  // distributed: the first cfem_num_owned entries are the owned dofs, then the ghosts
  std::memcpy(n2u, c->hm.n2u.data(), c->hm.n2u.size() * sizeof(int32_t));
  API_END
}

int cfem_nodal_h(cfem_ctx* c, double* h_out, double rtol, int max_it, int* iters) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  double* b = c->wk[8];
  launch_nodal_h_rhs(c, b);
  launch_fill(c, c->h, 0.0, c->dm.nn);
  int predict = 30;
  SolveResult r = mass_solve(c, 0, c->mat[CFEM_MAT_MASS], b, c->h, rtol, max_it, &predict);
  if (iters) *iters = r.iters;
  if (h_out) export_vec(c, c->h, h_out);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  if (!r.converged) CFEM_THROW(-3, "nodal_h: PCG did not converge");
  API_END
}

int cfem_rv_residual(cfem_ctx* c, int flux, int scheme, double dt, const double* u_n, const double* u_old,
                     const double* u_oo, const double* w, int use_bc, double* R_io, double rtol, int max_it,
                     int* iters) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!u_n || !u_old || !R_io) CFEM_THROW(-1, "rv_residual: null argument");
  if (scheme != CFEM_BDF1 && scheme != CFEM_BDF2) CFEM_THROW(-1, "rv_residual: unknown scheme");
  import_vec(c, u_n, c->u_n);
  import_vec(c, u_old, c->u_old);
  if (u_oo) import_vec(c, u_oo, c->u_oo);
  if (w) import_vec2(c, w, c->w);
  import_vec(c, R_io, c->RH);
  double* b = c->wk[8];
  launch_rv_rhs(c, flux, scheme, dt, c->u_n, c->u_old, u_oo ? c->u_oo : nullptr, w ? c->w : nullptr, use_bc != 0, b,
                nullptr);
  SolveResult r = mass_solve(c, 0, c->mat[use_bc ? CFEM_MAT_MASS_BC : CFEM_MAT_MASS], b, c->RH, rtol, max_it, &c->pcg_predict);
  if (iters) *iters = r.iters;
  export_vec(c, c->RH, R_io);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  if (!r.converged) CFEM_THROW(-3, "rv_residual: PCG did not converge");
  API_END
}

int cfem_rv_epsilon(cfem_ctx* c, int variant, int flux, double Cvel, double Crv, const double* uh,
                    const double* u_n, double* Rh, const double* h, const double* w, double* eps_out) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!eps_out) CFEM_THROW(-1, "rv_epsilon: null output");
  if (uh) import_vec(c, uh, c->uh);
  if (u_n) import_vec(c, u_n, c->u_n);
  if (Rh) import_vec(c, Rh, c->RH);
  if (h) import_vec(c, h, c->h);
  if (w) import_vec2(c, w, c->w);
  launch_epsilon(c, variant, flux, Cvel, Crv, uh ? c->uh : nullptr, u_n ? c->u_n : nullptr, Rh ? c->RH : nullptr,
                 h ? c->h : nullptr, w ? c->w : nullptr, c->eps);
  export_vec(c, c->eps, eps_out);
  if (variant == CFEM_EPS_LINEAR_SIMPLE && Rh) export_vec(c, c->RH, Rh);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_si_epsilon(cfem_ctx* c, int flux, double Cm, double floor_, int use_bc, const double* u_n, const double* h,
                    const double* w, double* psi_out, double* eps_out) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!u_n || !h || !eps_out) CFEM_THROW(-1, "si_epsilon: null argument");
  if (!c->unit_stiffness.vals) {
    c->unit_stiffness.vals = dalloc<double>(c, c->dm.nnz + 8);
    c->unit_stiffness.dinv = dalloc<double>(c, c->dm.nn);
  }
  if (!c->unit_stiffness.valid) launch_stiffness(c, c->unit_stiffness, nullptr);
  import_vec(c, u_n, c->u_n);
  import_vec(c, h, c->h);
  if (w) import_vec2(c, w, c->w);
  launch_si_epsilon(c, flux, Cm, floor_, use_bc != 0, c->unit_stiffness, c->u_n, c->h, w ? c->w : nullptr,
                    psi_out ? c->wk[9] : nullptr, c->eps);
  export_vec(c, c->eps, eps_out);
  if (psi_out) export_vec(c, c->wk[9], psi_out);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_assemble_advection(cfem_ctx* c, double dt, const double* w, const double* eps, const double* u_n,
                            const double* bc_values, double* b_out) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!w || !u_n) CFEM_THROW(-1, "assemble_advection: null argument");
  import_vec2(c, w, c->w);
  if (eps) import_vec(c, eps, c->eps);
  import_vec(c, u_n, c->u_n);
  import_bc(c, bc_values);
  launch_adv_system(c, dt, c->w, eps ? c->eps : nullptr, c->u_n, c->g, c->mat[CFEM_MAT_SYSTEM], c->wk[8]);
  if (b_out) export_vec(c, c->wk[8], b_out);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_assemble_cn_residual(cfem_ctx* c, int flux, double dt, const double* uh, const double* u_n,
                              const double* eps, const double* bc_values, double* F_out) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!uh || !u_n || !eps || !F_out) CFEM_THROW(-1, "assemble_cn_residual: null argument");
  import_vec(c, uh, c->uh);
  import_vec(c, u_n, c->u_n);
  import_vec(c, eps, c->eps);
  import_bc(c, bc_values);
  launch_cn_residual(c, flux, dt, c->uh, c->u_n, c->eps, c->g, nullptr, c->wk[8], nullptr);
  export_vec(c, c->wk[8], F_out);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_assemble_cn_jacobian(cfem_ctx* c, int flux, double dt, const double* uh, const double* eps) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!uh || !eps) CFEM_THROW(-1, "assemble_cn_jacobian: null argument");
  import_vec(c, uh, c->uh);
  import_vec(c, eps, c->eps);
  launch_cn_jacobian(c, flux, dt, c->uh, c->eps, c->mat[CFEM_MAT_SYSTEM]);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_assemble_stiffness(cfem_ctx* c, const double* eps) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (eps) import_vec(c, eps, c->eps);
  launch_stiffness(c, c->mat[CFEM_MAT_STIFFNESS], eps ? c->eps : nullptr);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_matrix_values(cfem_ctx* c, int which, double* vals_out) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  Matrix& A = get_matrix(c, which);
  build_user_csr(c);
  std::vector<double> tmp(c->hm.nnz);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaMemcpy(tmp.data(), A.vals, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (is_device_ptr(vals_out)) CFEM_THROW(-1, "matrix_values: output must be host memory");
  for (int64_t s = 0; s < c->hm.nnz; ++s) vals_out[s] = tmp[c->u_slot[s]];
  API_END
}

int cfem_spmv(cfem_ctx* c, int which, const double* x, double* y) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  Matrix& A = get_matrix(c, which);
  import_vec(c, x, c->wk[8]);
  launch_spmv(c, A, c->wk[8], c->wk[9]);
  export_vec(c, c->wk[9], y);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_solve(cfem_ctx* c, int which, int solver, const double* b, double* x_io, double rtol, double atol,
               int max_it, int* iters, double* relres) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  Matrix& A = get_matrix(c, which);
  import_vec(c, b, c->wk[8]);
  import_vec(c, x_io, c->wk[9]);
  int predict = 8;
  SolveResult r = run_solver(c, solver, A, c->wk[8], c->wk[9], rtol, atol, max_it, &predict);
  if (iters) *iters = r.iters;
  if (relres) *relres = r.relres;
  export_vec(c, c->wk[9], x_io);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  if (!r.converged) CFEM_THROW(-3, "solve: Krylov solver did not converge");
  API_END
}

static void state_load(cfem_ctx* c, const double* uh, const double* u_n, const double* u_old, const double* u_oo,
                       const double* RH, const double* h, const double* w, double t) {
  CUDA_OK(cudaSetDevice(c->device));
  if (uh) import_vec(c, uh, c->uh);
  if (u_n) import_vec(c, u_n, c->u_n);
  if (u_old) import_vec(c, u_old, c->u_old);
  if (u_oo) import_vec(c, u_oo, c->u_oo);
  if (RH) import_vec(c, RH, c->RH);
  if (h) import_vec(c, h, c->h);
  if (w) import_vec2(c, w, c->w);
  c->t = t;
}

int cfem_state_set(cfem_ctx* c, const double* uh, const double* u_n, const double* u_old, const double* u_oo,
                   const double* RH, const double* h, const double* w, double t) {
  API_BEGIN
  state_load(c, uh, u_n, u_old, u_oo, RH, h, w, t);
  // iteration-count predictions restart with the state, so a run is a pure function of its inputs
  c->pcg_predict = 28;
  c->krylov_predict = 8;
  c->dx_guess_valid = false;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_state_update(cfem_ctx* c, const double* uh, const double* u_n, const double* u_old, const double* u_oo,
                      const double* RH, const double* h, const double* w, double t) {
  API_BEGIN
  state_load(c, uh, u_n, u_old, u_oo, RH, h, w, t);  // stream-ordered: no host sync needed before the next call
  API_END
}

int cfem_state_get(cfem_ctx* c, double* uh, double* u_n, double* u_old, double* u_oo, double* RH, double* eps,
                   double* t) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (uh) export_vec(c, c->uh, uh);
  if (u_n) export_vec(c, c->u_n, u_n);
  if (u_old) export_vec(c, c->u_old, u_old);
  if (u_oo) export_vec(c, c->u_oo, u_oo);
  if (RH) export_vec(c, c->RH, RH);
  if (eps) export_vec(c, c->eps, eps);
  if (t) *t = c->t;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_state_update_owned(cfem_ctx* c, const double* uh, const double* u_n, const double* u_old,
                            const double* u_oo, const double* RH, double t) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  const double* src[5] = {uh, u_n, u_old, u_oo, RH};
  double* dst[5] = {c->uh, c->u_n, c->u_old, c->u_oo, c->RH};
  for (int k = 0; k < 5; ++k)
    if (src[k]) CUDA_OK(cudaMemcpyAsync(dst[k], src[k], c->dm.no * sizeof(double), cudaMemcpyDefault, c->stream));
  for (int k = 0; k < 5; ++k)
    if (src[k]) halo_exchange(c, dst[k]);
  c->t = t;
  API_END
}

int cfem_state_get_owned(cfem_ctx* c, double* uh, double* u_n, double* u_old, double* u_oo, double* RH, double* eps,
                         double* t) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  double* dst[6] = {uh, u_n, u_old, u_oo, RH, eps};
  const double* src[6] = {c->uh, c->u_n, c->u_old, c->u_oo, c->RH, c->eps};
  for (int k = 0; k < 6; ++k)
    if (dst[k]) CUDA_OK(cudaMemcpyAsync(dst[k], src[k], c->dm.no * sizeof(double), cudaMemcpyDefault, c->stream));
  if (t) *t = c->t;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

// ||F||_2 of the last residual assembly on the host (one sync).  With in-kernel reductions the assembly kernel has
// already finished the sum (over its CTAs and, distributed, the ranks) into scalars[24]; otherwise (NCCL fallback) the
// per-CTA partials are all-reduced and added up here.
static double partials_norm(cfem_ctx* c, double* part, int npart) {
  double* tmp = c->h_pinned + 64;
  if (fin_available(c)) {
    CUDA_OK(cudaMemcpyAsync(tmp, c->scalars + 24, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return sqrt(tmp[0]);
  }
  npart = allreduce_sum1(c, part, npart);
  CUDA_OK(cudaMemcpyAsync(tmp, part, npart * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  double s = 0.0;
  for (int i = 0; i < npart; ++i) s += tmp[i];
  return sqrt(s);
}

// CUDA event pair that is destroyed on every exit path (a throwing solver must not leak events)
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() { CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b)); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  EventPair(const EventPair&) = delete;
  EventPair& operator=(const EventPair&) = delete;
};
// restores the simulation time when a stepper leaves by an exception: the fields of the failed step are undefined
// (documented in the header), but a caller that reloads the state and retries starts from a consistent clock
struct TimeGuard {
  cfem_ctx* c; double t0; bool armed = true;
  explicit TimeGuard(cfem_ctx* c_) : c(c_), t0(c_->t) {}
  ~TimeGuard() { if (armed) c->t = t0; }
};

// smoothness-indicator variant of the scalar stepper (Exact_Burger_SI.py:159-197): SI viscosity instead of the
// residual projection + RV formula, optional smooth_vector post-filter after the Newton solve
struct SiOptions { double Cm, floor; double smooth_l; const int32_t* smooth_order; };

static int step_scalar_impl(cfem_ctx* c, const cfem_step_params* p, int n_steps, const double* bc_values,
                            cfem_step_stats* stats, const SiOptions* si) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!p) CFEM_THROW(-1, "step_scalar: null params");
  if (p->flux != CFEM_FLUX_BURGERS && p->flux != CFEM_FLUX_KPP) CFEM_THROW(-1, "step_scalar: flux must be BURGERS or KPP");
  if (!(p->dt > 0.0)) CFEM_THROW(-1, "step_scalar: dt must be positive");
  const int64_t nn = c->dm.nn;
  const Launches l0 = c->launches;
  cfem_step_stats st{};
  EventPair ev;
  TimeGuard tguard(c);
  cudaEvent_t ev0 = ev.a, ev1 = ev.b;
  CUDA_OK(cudaEventRecord(ev0, c->stream));
  const double* d_bc_user = nullptr;
  if (p->bc_kind == CFEM_BC_USER) {
    if (!bc_values) CFEM_THROW(-1, "step_scalar: CFEM_BC_USER needs bc_values");
    if (is_device_ptr(bc_values)) d_bc_user = bc_values;
    else {
      if ((int64_t)n_steps * c->nbc_user > 2 * nn) CFEM_THROW(-1, "step_scalar: too many user bc values for one call");
      CUDA_OK(cudaMemcpyAsync(c->stage[2], bc_values, (size_t)n_steps * c->nbc_user * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      d_bc_user = c->stage[2];
    }
  }
  const double mass_rtol = p->mass_rtol > 0.0 ? p->mass_rtol : p->lin_rtol;
  double* b = c->wk[8];
  double* F = c->wk[8];
  double* dx = c->wk[9];
  double* normpart = c->partials + 7 * kMaxPartials;
  Matrix& J = c->mat[CFEM_MAT_SYSTEM];
  for (int s = 0; s < n_steps; ++s) {
    c->t += p->dt;
    launch_bc_values(c, p->bc_kind, p->bc_value, c->t, d_bc_user ? d_bc_user + (int64_t)s * c->nbc_user : nullptr, c->g);
    const double* fluxn = nullptr;
    if (!si) {
      // (a-3) residual projection  M_bc RH = b
      launch_rv_rhs(c, p->flux, p->scheme, p->dt, c->u_n, c->u_old, c->u_oo, nullptr, true, b, c->fluxn);
      fluxn = c->fluxn;
      SolveResult rm = mass_solve(c, p->mass_solver, c->mat[CFEM_MAT_MASS_BC], b, c->RH, mass_rtol, p->lin_max_it, &c->pcg_predict);
      if (!rm.converged) CFEM_THROW(-3, "step_scalar: residual PCG did not converge");
      st.mass_iterations += rm.iters;
      // (a-4) nodal viscosity
      launch_epsilon(c, CFEM_EPS_NONLINEAR, p->flux, p->Cvel, p->Crv, c->uh, c->u_n, c->RH, c->h, nullptr, c->eps);
    } else {
      // (f-1) smoothness-indicator viscosity from the bc'd unit stiffness matrix (Exact_Burger_SI.py:169-174)
      if (!c->unit_stiffness.vals) {
        c->unit_stiffness.vals = dalloc<double>(c, c->dm.nnz + 8);
        c->unit_stiffness.dinv = dalloc<double>(c, c->dm.nn);
      }
      if (!c->unit_stiffness.valid) launch_stiffness(c, c->unit_stiffness, nullptr);
      launch_si_epsilon(c, p->flux, si->Cm, si->floor, true, c->unit_stiffness, c->u_n, c->h, nullptr, nullptr, c->eps);
    }
    // (a-8) Newton on the Crank-Nicolson residual, dolfinx NewtonSolver 'residual' criterion
    // the first evaluation assembles F and J together (one cell pass); CFEM_FUSED_FJ=0 keeps them apart
    static const bool fused_fj = !(getenv("CFEM_FUSED_FJ") && std::string(getenv("CFEM_FUSED_FJ")) == "0");
    int np = fused_fj ? launch_cn_residual_jacobian(c, p->flux, p->dt, c->uh, c->u_n, c->eps, c->g, fluxn, F, normpart, J)
                      : launch_cn_residual(c, p->flux, p->dt, c->uh, c->u_n, c->eps, c->g, fluxn, F, normpart);
    // ||F(u_0)||: the host needs it only to normalise the LATER residuals (and for the never-taken "already below
    // atol" exit), so with in-kernel finished norms its copy is queued here and read after the first linear solve --
    // one host round trip per step less (each one drains the stream and, distributed, exposes every rank's host
    // jitter to all the others).  CFEM_SYNC_RES0=1 waits here as round 1 did.
    static const bool sync_res0 = getenv("CFEM_SYNC_RES0") && std::string(getenv("CFEM_SYNC_RES0")) == "1";
    const bool defer0 = fin_available(c) && !sync_res0;
    double res = 0.0, res0 = 0.0;
    if (defer0) {
      CUDA_OK(cudaMemcpyAsync(c->h_pinned + 32, c->scalars + 24, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    } else {
      res = res0 = partials_norm(c, normpart, np);
    }
    bool converged = !defer0 && res < p->newton_atol;
    // CFEM_STEP_SYNC=1: poll the linear solve before queuing what follows it (two round trips per Newton iteration)
    static const bool step_sync = getenv("CFEM_STEP_SYNC") && std::string(getenv("CFEM_STEP_SYNC")) == "1";
    const bool async_solve = defer0 && !step_sync && p->solver == CFEM_SOLVER_BICGSTAB && bicgstab_async_available(c);
    int it = 0;
    while (!converged && it < p->newton_max_it) {
      if (it > 0 || !fused_fj) launch_cn_jacobian(c, p->flux, p->dt, c->uh, c->eps, J);
      // The first Newton update of a step is close to the previous step's (the solution moves by nearly the
      // same amount): start the Krylov solve from it.  The solve still runs to lin_rtol, so only the
      // iteration count changes.  Later Newton iterations (tiny corrections) start from zero.
      static const bool use_guess = !(getenv("CFEM_DXGUESS") && std::string(getenv("CFEM_DXGUESS")) == "0");
      const bool guess = use_guess && it == 0 && c->dx_guess_valid;
      if (guess) launch_copy(c, dx, c->dx_guess, nn);
      else launch_fill(c, dx, 0.0, nn);
      // Dirichlet rows of J are identity rows and their columns are lifted into F: dx = F there, exactly.  Starting
      // from it keeps those rows out of the iteration (zero residual, decoupled), so uh - dx lands on g to the bit
      // like the reference's LU does.
      launch_copy_indexed(c, dx, F, c->d_bc_nodes, c->nbc);
      if (async_solve) {
        // ONE host round trip per Newton iteration: the solve, the update, the ghost refresh and the next residual are
        // queued back to back; the verdict of the solve and ||F|| are read together.  (A failed solve has then already
        // touched uh -- the step throws, and the fields of a failed step are undefined anyway.)
        bicgstab_persist_begin(c, J, F, dx, p->lin_rtol, 0.0, p->lin_max_it);
        if (use_guess && it == 0) {
          if (!c->dx_guess) c->dx_guess = dalloc<double>(c, nn);
          launch_copy(c, c->dx_guess, dx, nn);
        }
        // dolfinx takes no Newton step when the initial residual already meets atol: the device skips the update then
        if (it == 0) launch_sub_unless_below(c, c->uh, dx, c->dm.no, c->scalars + 24, p->newton_atol * p->newton_atol);
        else launch_sub(c, c->uh, dx, c->dm.no);
        halo_exchange(c, c->uh);
        np = launch_cn_residual(c, p->flux, p->dt, c->uh, c->u_n, c->eps, c->g, fluxn, F, normpart);
        res = partials_norm(c, normpart, np);          // the round trip
        const SolveResult rk = bicgstab_persist_end(c);
        if (it == 0) {
          res0 = sqrt(c->h_pinned[32]);
          if (res0 < p->newton_atol) { res = res0; converged = true; break; }
        }
        if (!rk.converged) CFEM_THROW(-3, "step_scalar: Krylov solve did not converge (relres " + std::to_string(rk.relres) + ")");
        st.krylov_iterations += rk.iters;
        if (use_guess && it == 0) c->dx_guess_valid = true;
        ++it;
        converged = (res / res0 < p->newton_rtol) || (res < p->newton_atol);
        continue;
      }
      SolveResult rk = run_solver(c, p->solver, J, F, dx, p->lin_rtol, 0.0, p->lin_max_it, &c->krylov_predict);
      if (defer0 && it == 0) {
        CUDA_OK(cudaStreamSynchronize(c->stream));   // the solver's own poll has already drained the stream
        res = res0 = sqrt(c->h_pinned[32]);
        // dolfinx takes no Newton step when the initial residual already meets atol: drop the update just computed
        if (res0 < p->newton_atol) { converged = true; break; }
      }
      if (!rk.converged) CFEM_THROW(-3, "step_scalar: Krylov solve did not converge (relres " + std::to_string(rk.relres) + ")");
      st.krylov_iterations += rk.iters;
      if (use_guess && it == 0) {
        if (!c->dx_guess) c->dx_guess = dalloc<double>(c, nn);
        launch_copy(c, c->dx_guess, dx, nn);
        c->dx_guess_valid = true;
      }
      launch_sub(c, c->uh, dx, c->dm.no);
      halo_exchange(c, c->uh);
      ++it;
      np = launch_cn_residual(c, p->flux, p->dt, c->uh, c->u_n, c->eps, c->g, fluxn, F, normpart);
      res = partials_norm(c, normpart, np);
      converged = (res / res0 < p->newton_rtol) || (res < p->newton_atol);
    }
    st.newton_iterations += it;
    st.last_newton_residual = res;
    if (!converged) CFEM_THROW(-3, "Newton solver did not converge in " + std::to_string(it) + " iterations");
    if (si && si->smooth_l > 0.0) launch_smooth_vector(c, c->uh, si->smooth_order, si->smooth_l);  // helpers.py:40-50
    // rotate  u_oo <- u_old <- u_n <- uh   (KPP_exact.py:159-161)
    double* tmp = c->u_oo;
    c->u_oo = c->u_old;
    c->u_old = c->u_n;
    c->u_n = tmp;
    launch_copy(c, c->u_n, c->uh, nn);
    st.steps++;
  }
  CUDA_OK(cudaEventRecord(ev1, c->stream));
  CUDA_OK(cudaEventSynchronize(ev1));
  comm_check(c);
  { float ms = 0.f; CUDA_OK(cudaEventElapsedTime(&ms, ev0, ev1)); st.device_ms = ms; }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  tguard.armed = false;
  st.time = c->t;
  st.kernel_launches = c->launches.total - l0.total;
  st.spmv_launches = c->launches.spmv - l0.spmv;
  st.assembly_launches = c->launches.assembly - l0.assembly;
  if (stats) *stats = st;
  API_END
}

int cfem_step_scalar(cfem_ctx* c, const cfem_step_params* p, int n_steps, const double* bc_values,
                     cfem_step_stats* stats) {
  return step_scalar_impl(c, p, n_steps, bc_values, stats, nullptr);
}

int cfem_step_scalar_si(cfem_ctx* c, const cfem_step_params* p, double Cm, double floor_, double smooth_l,
                        const int32_t* smooth_order, int n_steps, const double* bc_values, cfem_step_stats* stats) {
  const SiOptions si{Cm, floor_, smooth_l, smooth_order};
  return step_scalar_impl(c, p, n_steps, bc_values, stats, &si);
}

int cfem_smooth_vector(cfem_ctx* c, double* u_io, const int32_t* order, double l) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!u_io) CFEM_THROW(-1, "smooth_vector: null field");
  double* u = c->wk[9];
  import_vec(c, u_io, u);
  launch_smooth_vector(c, u, order, l);
  export_vec(c, u, u_io);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  comm_check(c);   // a timed-out peer exchange must not return garbage with status 0
  API_END
}

int cfem_step_advection(cfem_ctx* c, const cfem_step_params* p, int n_steps, int first_gfem,
                        cfem_step_stats* stats) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!p) CFEM_THROW(-1, "step_advection: null params");
  if (!(p->dt > 0.0)) CFEM_THROW(-1, "step_advection: dt must be positive");
  const int64_t nn = c->dm.nn;
  const Launches l0 = c->launches;
  cfem_step_stats st{};
  EventPair ev;
  TimeGuard tguard(c);
  cudaEvent_t ev0 = ev.a, ev1 = ev.b;
  CUDA_OK(cudaEventRecord(ev0, c->stream));
  const double mass_rtol = p->mass_rtol > 0.0 ? p->mass_rtol : p->lin_rtol;
  double* b = c->wk[8];
  Matrix& A = c->mat[CFEM_MAT_SYSTEM];
  launch_bc_values(c, CFEM_BC_CONSTANT, p->bc_kind == CFEM_BC_CONSTANT ? p->bc_value : 0.0, 0.0, nullptr, c->g);
  for (int s = 0; s < n_steps; ++s) {
    c->t += p->dt;
    const bool gfem = first_gfem && s == 0;
    if (!gfem) {
      // (a-3) BDF1 residual projection, RV_node.py:209-214
      launch_rv_rhs(c, CFEM_FLUX_ADVECTION, CFEM_BDF1, p->dt, c->u_n, c->u_old, nullptr, c->w, p->residual_bc != 0, b, nullptr);
      SolveResult rm = mass_solve(c, p->mass_solver, c->mat[p->residual_bc ? CFEM_MAT_MASS_BC : CFEM_MAT_MASS], b, c->RH,
                                  mass_rtol, p->lin_max_it, &c->pcg_predict);
      if (!rm.converged) CFEM_THROW(-3, "step_advection: residual PCG did not converge");
      st.mass_iterations += rm.iters;
      // (a-5) nodal viscosity
      launch_epsilon(c, CFEM_EPS_LINEAR, CFEM_FLUX_ADVECTION, p->Cvel, p->Crv, c->uh, c->u_n, c->RH, c->h, c->w, c->eps);
    }
    // (a-7) CN system + rhs, RV_node.py:220-242
    launch_adv_system(c, p->dt, c->w, gfem ? nullptr : c->eps, c->u_n, c->g, A, b);
    SolveResult rk = run_solver(c, p->solver, A, b, c->uh, p->lin_rtol, 0.0, p->lin_max_it, &c->krylov_predict);
    if (!rk.converged) CFEM_THROW(-3, "step_advection: Krylov solve did not converge");
    halo_exchange(c, c->uh);   // the new solution feeds the next step's cell loops: ghosts needed
    st.krylov_iterations += rk.iters;
    if (gfem) {
      launch_copy(c, c->u_n, c->uh, nn);  // u_old keeps the initial condition (RV_node.py:157)
    } else {
      double* tmp = c->u_old;
      c->u_old = c->u_n;
      c->u_n = tmp;
      launch_copy(c, c->u_n, c->uh, nn);
    }
    st.steps++;
  }
  CUDA_OK(cudaEventRecord(ev1, c->stream));
  CUDA_OK(cudaEventSynchronize(ev1));
  comm_check(c);
  { float ms = 0.f; CUDA_OK(cudaEventElapsedTime(&ms, ev0, ev1)); st.device_ms = ms; }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  tguard.armed = false;
  st.time = c->t;
  st.kernel_launches = c->launches.total - l0.total;
  st.spmv_launches = c->launches.spmv - l0.spmv;
  st.assembly_launches = c->launches.assembly - l0.assembly;
  if (stats) *stats = st;
  API_END
}

int cfem_profile_begin(cfem_ctx* c, int max_launches) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  Profiler& p = c->prof;
  if (max_launches < 16) max_launches = 16;
  while (p.ev.size() < (size_t)2 * max_launches) {
    cudaEvent_t e;
    CUDA_OK(cudaEventCreate(&e));
    p.ev.push_back(e);
  }
  p.cat.assign(p.ev.size() / 2, 0);
  p.weight.assign(p.ev.size() / 2, 1);
  p.used = 0;
  p.depth = 0;
  p.on = true;
  API_END
}

int cfem_profile_end(cfem_ctx* c, double* ms_per_category, int64_t* launches_per_category) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  Profiler& p = c->prof;
  p.on = false;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < PROF_NCAT; ++k) { ms_per_category[k] = 0.0; launches_per_category[k] = 0; }
  for (size_t e = 0; e + 1 < p.used; e += 2) {
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, p.ev[e], p.ev[e + 1]));
    ms_per_category[p.cat[e / 2]] += ms;
    launches_per_category[p.cat[e / 2]] += p.weight[e / 2];
  }
  p.used = 0;
  API_END
}

// Idle time of the stream BETWEEN the profiled scopes: gap_ms_after_category[k] = sum over consecutive scopes (a, b)
// with a in category k of max(0, start(b) - end(a)).  Call before cfem_profile_end (which releases the records).
int cfem_profile_gaps(cfem_ctx* c, double* gap_ms_after_category) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  Profiler& p = c->prof;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < PROF_NCAT; ++k) gap_ms_after_category[k] = 0.0;
  for (size_t e = 0; e + 3 < p.used; e += 2) {
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, p.ev[e + 1], p.ev[e + 2]));
    if (ms > 0.f) gap_ms_after_category[p.cat[e / 2]] += ms;
  }
  API_END
}

// ---- Euler system ---------------------------------------------
// (Nn,4) caller arrays <-> local AoS device vectors; host staging through pageable copies.
// One GPU: the caller's array goes to the device as it is (one copy at PCIe speed from pinned memory) and is permuted
// there; distributed contexts gather their own entries on the host and ship only those.
static double* stage4(cfem_ctx* c) {
  if (!c->stage4) c->stage4 = dalloc<double>(c, 4 * c->dm.nn);
  return c->stage4;
}
static void import_vec4(cfem_ctx* c, const double* user, double* dst) {
  const int64_t nl = c->dm.nn;
  if (c->world == 1) {
    const double* src = user;
    if (!is_device_ptr(user)) {
      CUDA_OK(cudaMemcpyAsync(stage4(c), user, 4 * (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      src = stage4(c);
    }
    launch_gather4(c, (const double4*)src, c->d_n2u, (double4*)dst, nl);
    return;
  }
  std::vector<double> tmp;
  const double* src = user;
  std::vector<double> host_copy;
  if (is_device_ptr(user)) {
    host_copy.resize(4 * (size_t)c->dm.nn_global);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(host_copy.data(), user, host_copy.size() * sizeof(double), cudaMemcpyDeviceToHost));
    src = host_copy.data();
  }
  tmp.resize(4 * (size_t)nl);
  const int32_t* n2u = c->hm.n2u.data();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nl; ++i)
    for (int k = 0; k < 4; ++k) tmp[4 * i + k] = src[4 * (int64_t)n2u[i] + k];
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaMemcpy(dst, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
}
static void export_vec4(cfem_ctx* c, const double* internal, double* user) {
  if (is_device_ptr(user)) CFEM_THROW(-1, "euler_state_get: outputs must be host arrays");
  const int64_t no = c->dm.no;
  if (c->world == 1) {
    launch_gather4(c, (const double4*)internal, c->d_u2n, (double4*)stage4(c), no);
    CUDA_OK(cudaMemcpyAsync(user, stage4(c), 4 * (size_t)no * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return;
  }
  std::vector<double> tmp(4 * (size_t)no);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaMemcpy(tmp.data(), internal, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const int32_t* n2u = c->hm.n2u.data();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < no; ++i)
    for (int k = 0; k < 4; ++k) user[4 * (int64_t)n2u[i] + k] = tmp[4 * i + k];
}

int cfem_euler_state_set(cfem_ctx* c, const double* Uh, const double* Un, const double* Uold, const double* Uoo,
                         const double* bc_state, const double* h, double t) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  double *dUh, *dUn, *dUold, *dUoo, *dG;
  euler_state_ptrs(c, &dUh, &dUn, &dUold, &dUoo, &dG, nullptr);
  if (Uh) import_vec4(c, Uh, dUh);
  if (Un) import_vec4(c, Un, dUn);
  if (Uold) import_vec4(c, Uold, dUold);
  if (Uoo) import_vec4(c, Uoo, dUoo);
  if (bc_state) import_vec4(c, bc_state, dG);
  if (h) import_vec(c, h, c->h);
  c->t = t;
  euler_reset_predictions(c);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  API_END
}

int cfem_euler_state_get(cfem_ctx* c, double* Uh, double* R, double* eps, double* t) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  double *dUh, *dR;
  euler_state_ptrs(c, &dUh, nullptr, nullptr, nullptr, nullptr, &dR);
  if (Uh) export_vec4(c, dUh, Uh);
  if (R) export_vec4(c, dR, R);
  if (eps) export_vec(c, c->eps, eps);
  if (t) *t = c->t;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  API_END
}

int cfem_step_euler(cfem_ctx* c, const cfem_step_params* p, int n_steps, cfem_step_stats* stats) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!p) CFEM_THROW(-1, "step_euler: null params");
  if (!(p->dt > 0.0)) CFEM_THROW(-1, "step_euler: dt must be positive");
  const Launches l0 = c->launches;
  cfem_step_stats st{};
  EventPair ev;
  TimeGuard tguard(c);
  cudaEvent_t ev0 = ev.a, ev1 = ev.b;
  l2_release(c);   // the Euler kernels stream far more than the L2 holds: give them all of it
  CUDA_OK(cudaEventRecord(ev0, c->stream));
  euler_steps(c, p, n_steps, &st);
  CUDA_OK(cudaEventRecord(ev1, c->stream));
  CUDA_OK(cudaEventSynchronize(ev1));
  comm_check(c);
  { float ms = 0.f; CUDA_OK(cudaEventElapsedTime(&ms, ev0, ev1)); st.device_ms = ms; }
  tguard.armed = false;
  st.time = c->t;
  st.kernel_launches = c->launches.total - l0.total;
  st.spmv_launches = c->launches.spmv - l0.spmv;
  st.assembly_launches = c->launches.assembly - l0.assembly;
  if (stats) *stats = st;
  API_END
}

int cfem_l2_error_p3(cfem_ctx* c, const double* uh, const double* uex_cells, double* err_out) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (!uex_cells || !err_out) CFEM_THROW(-1, "l2_error_p3: null argument");
  if (uh) import_vec(c, uh, c->uh);
  const int64_t nc = c->dm.nc;
  // rows of the caller's (Nc_global, 10) table for this rank's cells, gathered on the host (or read in place on the device)
  const double* table = uex_cells;
  bool local_rows = false;
  double* dtab = nullptr;
  if (!is_device_ptr(uex_cells)) {
    std::vector<double> rows(10 * (size_t)nc);
    const int32_t* cu = c->hm.cell_user.data();
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nc; ++k) {
      const int64_t u = cu[k] >= 0 ? cu[k] : ~cu[k];
      for (int a = 0; a < 10; ++a) rows[10 * k + a] = uex_cells[10 * u + a];
    }
    CUDA_OK(cudaMalloc((void**)&dtab, rows.size() * sizeof(double)));
    CUDA_OK(cudaMemcpy(dtab, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
    table = dtab;
    local_rows = true;
  }
  const double sq = launch_l2_error_p3(c, c->uh, table, local_rows);
  if (dtab) cudaFree(dtab);
  comm_check(c);
  *err_out = sqrt(sq > 0.0 ? sq : 0.0);
  API_END
}

int cfem_time_kernel(cfem_ctx* c, int kernel, int flux, int reps, double* ms_per_launch, double* algorithmic_bytes) {
  API_BEGIN
  CUDA_OK(cudaSetDevice(c->device));
  if (reps < 1) reps = 1;
  const double nn = (double)c->dm.nn, nc = (double)c->dm.nc, nnz = (double)c->dm.nnz;
  const double dt = 1e-3;
  // adjacency + tiling metadata every assembly launch streams: tile cell lists (4 B per
  // tile-cell), connectivity (12 B per tile-cell), codes (4 B per (node,cell) = 3 Nc),
  // v2c_ptr + rowptr + tile tables; coordinates 16 B per node.
  const double tile_cells = (double)c->dm.nc * 1.0;  // lower bound: every cell staged once
  const double meta = 16.0 * tile_cells + 4.0 * 3.0 * nc + 8.0 * nn + 16.0 * nn;
  double bytes = 0.0;
  EventPair evp;
  cudaEvent_t e0 = evp.a, e1 = evp.b;
  Matrix& M = c->mat[CFEM_MAT_MASS_BC];
  Matrix& J = c->mat[CFEM_MAT_SYSTEM];
  auto body = [&]() {
    switch (kernel) {
      case CFEM_KERNEL_SPMV: l2_prefer(c, M); launch_spmv(c, M, c->u_n, c->wk[9]); bytes = 12.0 * nnz + 4.0 * (nn + 1) + 16.0 * nn; break;
      case CFEM_KERNEL_SPMV_SYSTEM:
        if (!J.valid) CFEM_THROW(-1, "time_kernel: no system matrix assembled yet");
        l2_prefer(c, J);
        launch_spmv_dots2(c, J, c->u_n, c->wk[9], c->wk[8], c->partials + 5 * kMaxPartials, c->partials + 6 * kMaxPartials);
        bytes = 12.0 * nnz + 4.0 * (nn + 1) + 24.0 * nn; break;
      case CFEM_KERNEL_ASM_RESIDUAL:
        launch_cn_residual(c, flux, dt, c->uh, c->u_n, c->eps, c->g, c->fluxn, c->wk[8], c->partials + 7 * kMaxPartials);
        bytes = meta + 8.0 * 5 * nn + 8.0 * nn; break;   // uh,u_n,eps,g,fluxn in; F out
      case CFEM_KERNEL_ASM_JACOBIAN:
        launch_cn_jacobian(c, flux, dt, c->uh, c->eps, J);
        bytes = meta + 8.0 * 2 * nn + 8.0 * nnz + 8.0 * nn; break;  // uh,eps in; vals + dinv out
      case CFEM_KERNEL_RV_EPSILON:
        launch_epsilon(c, CFEM_EPS_NONLINEAR, flux, 0.5, 4.0, c->uh, c->u_n, c->RH, c->h, nullptr, c->eps);
        bytes = 8.0 * 4 * nn + 16.0 * nn + 4.0 * nnz + 4.0 * nn + 8.0 * nn; break;  // uh,u_n,Rh,h; beta w+r; graph; eps
      case CFEM_KERNEL_ASM_RV_RHS:
        launch_rv_rhs(c, flux, CFEM_BDF2, dt, c->u_n, c->u_old, c->u_oo, nullptr, true, c->wk[8], c->fluxn);
        bytes = meta + 8.0 * 3 * nn + 16.0 * nn; break;
      case CFEM_KERNEL_COMM_ALLREDUCE: {
        double* sl[3] = {c->partials + 5 * kMaxPartials, c->partials + 6 * kMaxPartials, c->partials + 7 * kMaxPartials};
        const int op[3] = {0, 1, 2};
        allreduce_partials(c, 3, sl, op, 1184);
        bytes = 24.0 * c->world; break; }
      case CFEM_KERNEL_COMM_HALO: halo_exchange(c, c->u_n, 1); bytes = 8.0 * (double)(c->dm.nn - c->dm.no); break;
      case CFEM_KERNEL_CHEB_ITER: {
        // a whole mass solve of exactly 24 iterations (tolerance 0: no early exit; one norm check at the end), as the
        // residual projection of a step runs it -- collective in a distributed context.  Reported per iteration.
        int predict = 24;
        chebyshev_mass(c, M, c->u_n, c->wk[9], 0.0, 24, &predict);
        bytes = 24.0 * (12.0 * nnz + 4.0 * (nn + 1) + 48.0 * nn); break; }
      case CFEM_KERNEL_KRYLOV_ITER: {
        // a whole BiCGStab solve of exactly 16 iterations on the current system matrix (tolerances 0), collective
        if (!J.valid) CFEM_THROW(-1, "time_kernel: no system matrix assembled yet");
        launch_fill(c, c->wk[9], 0.0, c->dm.nn);
        bicgstab(c, J, c->u_n, c->wk[9], 0.0, 0.0, 16, nullptr);
        bytes = 16.0 * (24.0 * nnz + 8.0 * (nn + 1) + 136.0 * nn); break; }
      default: CFEM_THROW(-1, "time_kernel: unknown kernel id");
    }
  };
  body();  // warm-up (also sets function attributes)
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaEventRecord(e0, c->stream));
  for (int r = 0; r < reps; ++r) body();
  CUDA_OK(cudaEventRecord(e1, c->stream));
  CUDA_OK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_per_launch) *ms_per_launch = (double)ms / reps;
  if (algorithmic_bytes) *algorithmic_bytes = bytes;
  API_END
}

}  // extern "C"

// ---- host analysis without a device ------------------------------------------
struct cfem_host_mesh { cfem::HostMesh hm; };

template <class F>
static auto host_array(const cfem_host_mesh* h, int what, F&& f) {
  const HostMesh& m = h->hm;
  switch (what) {
    case CFEM_HM_N2U: return f(m.n2u.data(), m.n2u.size(), 4);
    case CFEM_HM_CELLS: return f(m.cells.data(), m.cells.size(), 4);
    case CFEM_HM_ROWPTR: return f(m.rowptr.data(), m.rowptr.size(), 4);
    case CFEM_HM_COLIDX: return f(m.colidx.data(), m.colidx.size(), 4);
    case CFEM_HM_V2C_PTR: return f(m.v2c_ptr.data(), m.v2c_ptr.size(), 4);
    case CFEM_HM_V2C_CODE: return f(m.v2c_code.data(), m.v2c_code.size(), 4);
    case CFEM_HM_TILE_NODE: return f(m.tile_node.data(), m.tile_node.size(), 4);
    case CFEM_HM_TILE_CELLPTR: return f(m.tile_cellptr.data(), m.tile_cellptr.size(), 4);
    case CFEM_HM_TILE_CELLS: return f(m.tile_cells.data(), m.tile_cells.size(), 4);
    case CFEM_HM_IS_BND: return f(m.is_bnd.data(), m.is_bnd.size(), 1);
    case CFEM_HM_BND_USER: return f(m.bnd_user_sorted.data(), m.bnd_user_sorted.size(), 4);
    case CFEM_HM_PEER_RANK: return f(m.peer_rank.data(), m.peer_rank.size(), 4);
    case CFEM_HM_SEND_PTR: return f(m.send_ptr.data(), m.send_ptr.size(), 4);
    case CFEM_HM_SEND_IDX: return f(m.send_idx.data(), m.send_idx.size(), 4);
    case CFEM_HM_RECV_OFF: return f(m.recv_off.data(), m.recv_off.size(), 4);
    case CFEM_HM_RECV_CNT: return f(m.recv_cnt.data(), m.recv_cnt.size(), 4);
    case CFEM_HM_LAST_CELL: return f(m.last_cell.data(), m.last_cell.size(), 4);
    case CFEM_HM_LC16: return f(m.lc16.data(), m.lc16.size(), 2);
    case CFEM_HM_TILE_EXTPTR: return f(m.tile_extptr.data(), m.tile_extptr.size(), 4);
    case CFEM_HM_TILE_EXT: return f(m.tile_ext.data(), m.tile_ext.size(), 4);
    case CFEM_HM_TILE_ORDER: return f(m.tile_order.data(), m.tile_order.size(), 4);
    default: return f(nullptr, (size_t)0, 0);
  }
}

extern "C" {

int cfem_host_analyse(cfem_host_mesh** out, int64_t n_nodes, int64_t n_cells, const double* x, int xdim,
                      const void* cells, int cell_index_bytes, int order) {
  cfem_host_mesh* h = nullptr;
  try {
    if (!out || !x || !cells) CFEM_THROW(-1, "null argument");
    h = new cfem_host_mesh();
    analyse_mesh(h->hm, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, order);
    *out = h;
    return 0;
  } catch (const cfem::Error& e) { cfem::set_error(e.msg); delete h; return e.code; }
  catch (const std::exception& e) { cfem::set_error(e.what()); delete h; return -9; }
}

int cfem_host_analyse_part(cfem_host_mesh** out, int rank, int world, int64_t n_nodes, int64_t n_cells,
                           const double* x, int xdim, const void* cells, int cell_index_bytes, int order) {
  cfem_host_mesh* h = nullptr;
  try {
    if (!out || !x || !cells) CFEM_THROW(-1, "null argument");
    h = new cfem_host_mesh();
    analyse_mesh(h->hm, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, order, rank, world);
    *out = h;
    return 0;
  } catch (const cfem::Error& e) { cfem::set_error(e.msg); delete h; return e.code; }
  catch (const std::exception& e) { cfem::set_error(e.what()); delete h; return -9; }
}
int cfem_host_analyse_partitioned(cfem_host_mesh** out, int rank, int world, int64_t n_nodes, int64_t n_cells,
                                  const double* x, int xdim, const void* cells, int cell_index_bytes, int order,
                                  const int32_t* node_part) {
  cfem_host_mesh* h = nullptr;
  try {
    if (!out || !x || !cells) CFEM_THROW(-1, "null argument");
    h = new cfem_host_mesh();
    analyse_mesh(h->hm, n_nodes, n_cells, x, xdim, cells, cell_index_bytes, order, rank, world, node_part);
    *out = h;
    return 0;
  } catch (const cfem::Error& e) { cfem::set_error(e.msg); delete h; return e.code; }
  catch (const std::exception& e) { cfem::set_error(e.what()); delete h; return -9; }
}
int64_t cfem_host_info(const cfem_host_mesh* h, int what) {
  const HostMesh& m = h->hm;
  switch (what) {
    case 0: return m.n_owned;
    case 1: return m.nn;
    case 2: return m.nn_global;
    case 3: return m.nc;
    case 4: return m.nnz;
    default: return -1;
  }
}
int64_t cfem_host_size(const cfem_host_mesh* h, int what) {
  return host_array(h, what, [](const void*, size_t n, int) { return (int64_t)n; });
}
int cfem_host_copy(const cfem_host_mesh* h, int what, void* dst) {
  int64_t n = host_array(h, what, [&](const void* p, size_t n, int es) { if (p && n) std::memcpy(dst, p, n * es); return (int64_t)n; });
  return n >= 0 ? 0 : -1;
}
void cfem_host_free(cfem_host_mesh* h) { delete h; }

}  // extern "C"
