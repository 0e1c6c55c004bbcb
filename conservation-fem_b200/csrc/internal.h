// Internal declarations shared by the cfem_b200 translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cfem_b200.h"

namespace cfem {

// ---- error plumbing -----------------------------------------------------
void set_error(const std::string& msg);
struct Error { std::string msg; int code; };
#define CFEM_THROW(code, msg) throw ::cfem::Error{std::string(msg), (code)}
#define CUDA_OK(call)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess)                                                             \
      CFEM_THROW(-2, std::string(#call) + ": " + cudaGetErrorString(_e) + " at " +     \
                         __FILE__ + ":" + std::to_string(__LINE__));                   \
  } while (0)

// ---- tile / adjacency encoding --------------------------------------------
// One 32-bit code per (node, incident cell):
//   bits  0..12  index of the cell in the tile's cell list (<= 8191)
//   bits 13..14  local vertex number k of the node inside that cell
//   bits 15..19 / 20..24 / 25..29  position, inside the node's CSR row, of the
//                columns of the cell's vertices 0 / 1 / 2
constexpr int kCodeCellBits = 13;
constexpr int kMaxRow = 32;          // 5-bit positions
constexpr int kTileNodes = 256;      // nodes per assembly tile (== threads per CTA)
constexpr int kTileCellCap = 768;    // cells staged in shared memory per tile
constexpr int kTileNnzCap = 2560;    // CSR entries of a tile's rows staged in shared memory

// Host-side result of the mesh analysis (setup.cpp). All ids internal unless
// suffixed _user.
// In a distributed context (world > 1) "internal" ids are LOCAL: owned nodes first (a contiguous
// range [part_off[rank], part_off[rank+1]) of the global Hilbert order), then the ghost layer in
// ascending global order (hence grouped by owner rank).
struct HostMesh {
  int64_t nn = 0, nc = 0, nnz = 0;        // local nodes (owned + ghosts), local cells, nnz of owned rows
  int64_t n_owned = 0, nn_global = 0;
  int rank = 0, world = 1;
  std::vector<int64_t> part_off;          // world+1 offsets into the global Hilbert order
  std::vector<int32_t> ghost_global;      // global internal ids of the ghosts, ascending
  std::vector<int32_t> peer_rank;         // ranks we exchange with
  std::vector<int32_t> send_ptr, send_idx;  // per peer: owned local ids whose values the peer ghosts
  std::vector<int32_t> recv_off, recv_cnt;  // per peer: where its values land in the ghost segment
  std::vector<int32_t> n2u, u2n;          // local->user ; user->GLOBAL internal
  std::vector<double> xy;                 // 2*nn, internal order
  std::vector<int32_t> cells;             // 3*nc, internal ids, internal cell order
  std::vector<int32_t> rowptr, colidx;    // P1 pattern == node patches
  std::vector<int32_t> v2c_ptr;           // nn+1
  std::vector<uint32_t> v2c_code;         // 3*nc
  std::vector<int32_t> tile_node;         // ntiles+1
  std::vector<int32_t> tile_cellptr;      // ntiles+1
  std::vector<int32_t> tile_cells;        // concatenated cell lists
  // T16 tile format of the SpMV-type kernels: a tile's CSR columns as 16-bit tile-local indices.  Index < kTileNodes:
  // the tile's own row (n0 + index); otherwise kTileNodes + position in the tile's EXTERNAL column list (distinct
  // columns outside the tile, ascending -> ghost columns last).  A kernel stages x for own rows (coalesced) and
  // externals (one gather each) in shared memory once per tile instead of gathering x per CSR entry.
  std::vector<uint16_t> lc16;             // nnz
  std::vector<int32_t> tile_extptr;       // ntiles+1
  std::vector<int32_t> tile_ext;          // concatenated external column lists (local node ids)
  int max_tile_ext = 0;
  std::vector<int32_t> tile_order;        // tiles whose rows touch no ghost column first, the others last
  int n_interior_tiles = 0;
  std::vector<uint8_t> is_bnd;            // nn
  std::vector<int32_t> last_cell;         // n_owned: incident local cell with the highest caller index
  std::vector<int32_t> cell_user;         // nc: caller index of each local cell; ~index (negative) for cells whose smallest
                                          // vertex another rank owns (so cell-wise functionals count every cell once)
  std::vector<int32_t> bnd_user_sorted;   // boundary dofs, user ids ascending
  int max_row = 0, max_tile_cells = 0, max_tile_nnz = 0;
};

// node_part (nullable): caller-numbered part id per node (e.g. from cfem_host_partition); null = equal ranges of the
// Hilbert order
void analyse_mesh(HostMesh& hm, int64_t nn, int64_t nc, const double* x, int xdim,
                  const void* cells, int idx_bytes, int order, int rank = 0, int world = 1,
                  const int32_t* node_part = nullptr);
// METIS k-way partition of the nodal graph (partition.cpp): one part id per caller node
void metis_partition(int world, int64_t nn, int64_t nc, const void* cells, int idx_bytes, int32_t* part_out);
// user dof -> local id (owned or ghost) or -1 if this rank does not hold it
int32_t user_to_local(const HostMesh& hm, int64_t user_dof);

// Device view handed to kernels by value.
struct DevMesh {
  int64_t nn, nc, nnz;     // local nodes (owned + ghosts), local cells, nnz of owned rows
  int64_t no;              // owned nodes == rows; == nn on a single GPU
  int64_t nn_global;
  int ntiles;
  const double2* xy;
  const int32_t* cells;
  const int32_t* rowptr;
  const int32_t* colidx;
  const int32_t* v2c_ptr;
  const uint32_t* v2c_code;
  const int32_t* tile_node;
  const int32_t* tile_cellptr;
  const int32_t* tile_cells;
  const uint16_t* lc16;        // T16 format, see HostMesh
  const int32_t* tile_extptr;
  const int32_t* tile_ext;
  int ext_cap;                 // max external columns of any tile, rounded up to a multiple of 32
  const int32_t* tile_order;   // interior tiles first (identity on one GPU)
  int n_interior;              // number of tiles that need no ghost value
  const uint8_t* is_bc;    // current Dirichlet flags
  const int32_t* last_cell;  // per owned node, see HostMesh
  const int32_t* cell_user;  // per local cell, see HostMesh
};

constexpr int kMaxPartials = 4096;   // >= any reduction grid

struct Matrix {
  double* vals = nullptr;   // nnz
  double* dinv = nullptr;   // nn, 1/diag
  bool valid = false;
};

struct Launches {  // counters of our own kernel launches
  int64_t total = 0, spmv = 0, assembly = 0;
};

// Optional per-launch CUDA-event bracketing (bench.py roofline leg).
enum { PROF_SPMV = 0, PROF_ASM_VEC = 1, PROF_ASM_MAT = 2, PROF_KRYLOV_VEC = 3, PROF_RV = 4, PROF_MISC = 5, PROF_CHEB = 6, PROF_COMM = 7, PROF_SOLVER = 8, PROF_NCAT = 9 };
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<int> cat;
  std::vector<int> weight;       // kernel launches a scope stands for (a chain timed as a whole)
  size_t used = 0;               // events used
  int depth = 0;                 // nesting depth of live scopes
};

}  // namespace cfem

struct cfem_ctx;
namespace cfem {
// RAII: brackets the launches issued inside its scope with two events when profiling is on.
struct ProfScope {
  cfem_ctx* c; bool active; size_t idx = 0;
  ProfScope(cfem_ctx* c, int cat, int launches = 1);
  ~ProfScope();
};
}  // namespace cfem

struct cfem_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cfem::HostMesh hm;
  cfem::DevMesh dm{};
  // device copies owned by the context
  std::vector<void*> allocs;
  int64_t bytes = 0;
  int32_t *d_n2u = nullptr, *d_u2n = nullptr;
  uint8_t *d_is_bc = nullptr, *d_is_bnd = nullptr;
  int32_t* d_bc_nodes = nullptr;  // local ids of the Dirichlet nodes this rank holds (owned or ghost)
  int32_t* d_bc_pos = nullptr;    // their position in the caller's Dirichlet list (CFEM_BC_USER values)
  int64_t nbc = 0;                // held locally
  int64_t nbc_user = 0;           // size of the caller's list
  std::vector<int32_t> bc_user;   // caller's Dirichlet set
  cfem::Matrix mat[4];
  cfem::Matrix unit_stiffness;      // int grad u . grad v, assembled on first use (SI viscosity)
  void* smooth_plan = nullptr;      // level schedule of the last smooth_vector order (smooth.cu)
  // state vectors (internal order)
  double *uh = nullptr, *u_n = nullptr, *u_old = nullptr, *u_oo = nullptr, *RH = nullptr,
         *eps = nullptr, *h = nullptr, *g = nullptr, *fluxn = nullptr;
  double2* w = nullptr;
  double t = 0.0;
  // work vectors
  double *wk[10] = {nullptr};
  double *stage[4] = {nullptr};   // staging for host<->device + permutation
  double* stage4 = nullptr;       // 4 * nn doubles, allocated on first use (Euler state import / export on one GPU)
  double *partials = nullptr;     // 8 * kMaxPartials doubles
  double *partials12 = nullptr;   // 12 more slots (Euler: sum/min/max of 4 components)
  double* gmres_V = nullptr;      // lazily allocated Krylov basis (31 vectors) and small dense block
  double* gmres_small = nullptr;
  void* euler = nullptr;          // lazily created Euler state (euler.cu)
  double *scalars = nullptr;      // small device scalar block
  int32_t* status = nullptr;      // device ints: [0]=done flag, [1]=iterations
  double* h_pinned = nullptr;     // pinned host scratch (small)
  int32_t* h_status = nullptr;    // pinned
  // user-order CSR export (lazy)
  std::vector<int32_t> u_rowptr, u_colidx, u_slot;
  cfem::Launches launches;
  // ---- distributed (world > 1): NCCL communicator + halo buffers
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  void* p2p = nullptr;              // peer-memory exchange state (comm.cu), null -> NCCL path
  int32_t* d_send_idx = nullptr;
  double* d_sendbuf = nullptr;      // 2 * total send count (double2 exchanges)
  double* h_stage = nullptr;        // pinned host staging, 2 * nn doubles
  int64_t halo_exchanges = 0, allreduces = 0;
  cfem::Profiler prof;
  int pcg_predict = 28, krylov_predict = 8;
  double* dx_guess = nullptr;   // first Newton update of the previous step (initial guess of the next Krylov solve)
  bool dx_guess_valid = false;
  // ---- L2 residency of the matrix a solve streams repeatedly (linalg.cu: l2_prefer).  MASS_BC values, the T16
  // pattern tables and SYSTEM values are carved from ONE allocation in that order, so either matrix together with
  // the shared pattern is one contiguous access-policy window.
  char* hot_base = nullptr;
  size_t hot_off[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // MASS_BC vals | rowptr | lc16 | extptr | ext | SYSTEM vals | colidx | end
  size_t l2_setaside = 0;                // bytes of L2 set aside for persisting lines (0: feature off)
  size_t l2_max_window = 0;
  std::vector<std::pair<const void*, int>> asm_occ;   // occupancy of the assembly kernel instantiations on this context
  void* persist_plan = nullptr;          // launch plan of the persistent BiCGStab kernel (persist.cu)
  int t16_grid = 0;                      // grid of the T16 tile kernels (occupancy x SMs, <= tiles), 0 = not sized yet
  int l2_window = -1;                     // matrix id whose window is currently attached to the stream, -1 none
};
