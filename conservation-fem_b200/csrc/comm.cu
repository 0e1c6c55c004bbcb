// Multi-GPU plumbing: NCCL (resolved at run time from the already-loaded torch copy or the
// system libnccl.so.2 — the single-GPU path never touches it), forward halo exchange of
// ghost values before every SpMV / assembly, and all-reduce of the per-rank reduction scalars.
//
// The reference has no working parallel path of its own (its Utils loops are not MPI-safe,
// SURVEY.md section 1); this layer replaces dolfinx's scatter_forward / ghostUpdate calls
// (Code/Linear_advection/RV_node.py:241,246) and PETSc's parallel dot products.
#include <dlfcn.h>
#include <nccl.h>

#include <map>
#include <string>

#include "device_utils.cuh"
#include "launch.h"

namespace cfem {

#define LAUNCHED(c) do { CUDA_OK(cudaGetLastError()); (c)->launches.total++; } while (0)

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  if (api.ok) return api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // torch's copy, if imported
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) CFEM_THROW(-5, std::string("cannot load libnccl.so.2: ") + dlerror());
#define SYM(field, name)                                                   \
  api.field = (decltype(api.field))dlsym(h, name);                         \
  if (!api.field) CFEM_THROW(-5, std::string("libnccl lacks ") + name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  api.ok = true;
  return api;
}

#define NCCL_OK(call)                                                                          \
  do {                                                                                         \
    ncclResult_t _r = (call);                                                                  \
    if (_r != ncclSuccess)                                                                     \
      CFEM_THROW(-5, std::string(#call) + ": " + nccl().GetErrorString(_r));                   \
  } while (0)

void comm_unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCCL_OK(nccl().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

void comm_init(cfem_ctx* c, int rank, int world, const void* id128) {
  c->rank = rank;
  c->world = world;
  if (world == 1) return;
  // One communicator per unique id for the life of the process: contexts made with the same
  // id (several meshes in one run) share it; an id can only be used for one ncclCommInitRank.
  static std::map<std::string, ncclComm_t> cache;
  const std::string key((const char*)id128, sizeof(ncclUniqueId));
  auto it = cache.find(key);
  if (it == cache.end()) {
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    NCCL_OK(nccl().CommInitRank(&comm, world, id, rank));
    it = cache.emplace(key, comm).first;
  }
  c->nccl_comm = it->second;
}

void comm_destroy(cfem_ctx* c) { c->nccl_comm = nullptr; }  // communicators are process-lifetime

__global__ void k_pack(const double* __restrict__ v, const int32_t* __restrict__ idx, double* __restrict__ out,
                       int n, int width) {
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n * width; i += gridDim.x * kBlock) {
    const int node = i / width, k = i - node * width;
    out[i] = v[(size_t)idx[node] * width + k];
  }
}

// ghosts of v <- owners' values.  width 1 (double) or 2 (double2, interleaved).
void halo_exchange(cfem_ctx* c, double* v, int width) {
  if (c->world == 1) return;
  const HostMesh& hm = c->hm;
  const int npeer = (int)hm.peer_rank.size();
  if (npeer == 0) return;
  ProfScope ps(c, PROF_COMM);
  const int nsend = hm.send_ptr[npeer];
  if (nsend > 0) {
    const int g = (nsend * width + kBlock - 1) / kBlock;
    k_pack<<<g, kBlock, 0, c->stream>>>(v, c->d_send_idx, c->d_sendbuf, nsend, width);
    LAUNCHED(c);
  }
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  NCCL_OK(nccl().GroupStart());
  for (int k = 0; k < npeer; ++k) {
    const int cnt = hm.send_ptr[k + 1] - hm.send_ptr[k];
    if (cnt > 0)
      NCCL_OK(nccl().Send(c->d_sendbuf + (size_t)width * hm.send_ptr[k], (size_t)width * cnt, ncclDouble, hm.peer_rank[k], comm, c->stream));
    if (hm.recv_cnt[k] > 0)
      NCCL_OK(nccl().Recv(v + (size_t)width * hm.recv_off[k], (size_t)width * hm.recv_cnt[k], ncclDouble, hm.peer_rank[k], comm, c->stream));
  }
  NCCL_OK(nccl().GroupEnd());
  c->halo_exchanges++;
}

// Reduce each listed partial array (npart entries) to its element 0 locally, then all-reduce
// those scalars over the ranks.  Returns the partial count consumers must use afterwards (1).
// op: 0 sum, 1 min, 2 max.
struct SlotTable { double* p[8]; int op[8]; };

__global__ void __launch_bounds__(kBlock)
k_reduce_slots(const SlotTable t, int npart) {
  __shared__ double red[9];
  double* p = t.p[blockIdx.x];
  const int op = t.op[blockIdx.x];
  double s = op == 0 ? 0.0 : (op == 1 ? INFINITY : -INFINITY);
  for (int i = threadIdx.x; i < npart; i += kBlock) {
    const double x = p[i];
    s = op == 0 ? s + x : (op == 1 ? fmin(s, x) : fmax(s, x));
  }
  s = op == 0 ? block_sum(s, red) : (op == 1 ? block_min(s, red) : block_max(s, red));
  if (threadIdx.x == 0) p[0] = s;
}

int allreduce_partials(cfem_ctx* c, int nslots, double* const* slots, const int* ops, int npart) {
  if (c->world == 1) return npart;
  if (nslots > 8) CFEM_THROW(-1, "allreduce_partials: too many slots");
  ProfScope ps(c, PROF_COMM);
  SlotTable t;
  for (int k = 0; k < nslots; ++k) { t.p[k] = slots[k]; t.op[k] = ops[k]; }
  k_reduce_slots<<<nslots, kBlock, 0, c->stream>>>(t, npart); LAUNCHED(c);
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  NCCL_OK(nccl().GroupStart());
  for (int k = 0; k < nslots; ++k)
    NCCL_OK(nccl().AllReduce(slots[k], slots[k], 1, ncclDouble, ops[k] == 0 ? ncclSum : (ops[k] == 1 ? ncclMin : ncclMax), comm, c->stream));
  NCCL_OK(nccl().GroupEnd());
  c->allreduces++;
  return 1;
}

int allreduce_sum1(cfem_ctx* c, double* slot, int npart) {
  double* s[1] = {slot};
  const int op[1] = {0};
  return allreduce_partials(c, 1, s, op, npart);
}

}  // namespace cfem
