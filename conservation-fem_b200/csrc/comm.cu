// Multi-GPU plumbing: NCCL (resolved at run time from the already-loaded torch copy or the
// system libnccl.so.2 — the single-GPU path never touches it), forward halo exchange of
// ghost values before every SpMV / assembly, and all-reduce of the per-rank reduction scalars.
//
// The reference has no working parallel path of its own (its Utils loops are not MPI-safe,
// SURVEY.md section 1); this layer replaces dolfinx's scatter_forward / ghostUpdate calls
// (Code/Linear_advection/RV_node.py:241,246) and PETSc's parallel dot products.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "device_utils.cuh"
#include "launch.h"
#include "p2p.cuh"

namespace cfem {

#define LAUNCHED(c) do { CUDA_OK(cudaGetLastError()); (c)->launches.total++; } while (0)

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
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
  SYM(AllGather, "ncclAllGather");
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

void comm_destroy_p2p(cfem_ctx* c);
void comm_destroy(cfem_ctx* c) {
  comm_destroy_p2p(c);
  c->nccl_comm = nullptr;  // communicators are process-lifetime (cached by unique id)
}

// ---------------------------------------------------------------- peer-memory path (NVLink, CUDA IPC)
// NCCL's latency (~15 us per grouped send/recv or all-reduce) dominates at ~1M dofs per GPU, where
// a step issues ~100 such operations.  Here every rank owns a "mailbox" in device memory that its
// neighbours map through CUDA IPC; ONE kernel per exchange stores the halo values (or reduction
// scalars) directly into the neighbours' mailboxes over NVLink, publishes a sequence number, waits
// for the neighbours' numbers and copies the received values into place.  Two mailbox generations
// (sequence parity) make the protocol race free: a rank can only be one exchange ahead of a
// neighbour because each exchange waits for the neighbour's previous one.  Kernels on DIFFERENT
// GPUs wait on one another; nothing waits on another kernel of the same GPU.  Spins are bounded
// (~30 s) and raise an error flag instead of hanging.
// A cudaMalloc'ed mailbox may be a sub-allocation of a larger block: its IPC handle then names the BLOCK, so two
// contexts of one process can hold mailboxes with the same handle, and a handle can be opened only once per process.
// Opened blocks are therefore cached process-wide (reference counted) and every rank publishes its mailbox as
// (handle of the block, offset of the mailbox inside it).
struct IpcBlock { void* base; int refs; };
static std::map<std::string, IpcBlock>& ipc_cache() { static std::map<std::string, IpcBlock> m; return m; }
static void* ipc_open(const cudaIpcMemHandle_t& h) {
  const std::string key((const char*)&h, sizeof(h));
  auto it = ipc_cache().find(key);
  if (it != ipc_cache().end()) { it->second.refs++; return it->second.base; }
  void* ptr = nullptr;
  CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  ipc_cache()[key] = IpcBlock{ptr, 1};
  return ptr;
}
static void ipc_close(const std::string& key) {
  auto it = ipc_cache().find(key);
  if (it == ipc_cache().end()) return;
  if (--it->second.refs == 0) { cudaIpcCloseMemHandle(it->second.base); ipc_cache().erase(it); }
}

struct P2P {
  P2PDev d;
  std::vector<std::string> opened;   // keys of the peer blocks this context holds a reference to
  unsigned long long halo_seq = 0, red_seq = 0;
  int32_t* d_send_ptr = nullptr;
  int32_t* d_peer_rank = nullptr;
  P2PDev* d_dev = nullptr;   // device copy of d (for kernels that take it by pointer)
  int* h_error = nullptr;
};

// push only: the consumer (a SpMV-type kernel) does the waiting, and only in its boundary-tile CTAs
__global__ void __launch_bounds__(kBlock)
k_p2p_push(const P2PDev a, const double* __restrict__ v, const int width, const unsigned long long seq,
           const int32_t* __restrict__ status) {
  const int parity = (int)(seq & 1);
  if (!(status && status[0])) {
    for (int k = 0; k < a.npeer; ++k) {
      const int s0 = a.send_ptr[k], cnt = (a.send_ptr[k + 1] - s0) * width;
      double* dst = (double*)(a.peer_base[k] + a.halo_off + parity * a.halo_stride) + (size_t)a.dst_off[k] * width;
      for (int i = blockIdx.x * kBlock + threadIdx.x; i < cnt; i += gridDim.x * kBlock) {
        const int node = i / width, kk = i - node * width;
        dst[i] = v[(size_t)a.send_idx[s0 + node] * width + kk];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (gridDim.x == 1) {  // small halos: one CTA, no ticket
    if (threadIdx.x < a.npeer) *(volatile unsigned long long*)(a.peer_base[threadIdx.x] + 8 * a.rank) = seq;
    return;
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.counter, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < a.npeer) {
    __threadfence_system();
    *(volatile unsigned long long*)(a.peer_base[threadIdx.x] + 8 * a.rank) = seq;  // publish (even when gated: keeps the sequence aligned)
    if (threadIdx.x == 0) *a.counter = 0;
  }
}

__global__ void __launch_bounds__(kBlock)
k_p2p_halo(const P2PDev a, double* __restrict__ v, const int width, const unsigned long long seq) {
  const int parity = (int)(seq & 1);
  // ---- push my owned values into every neighbour's mailbox
  for (int k = 0; k < a.npeer; ++k) {
    const int s0 = a.send_ptr[k], cnt = (a.send_ptr[k + 1] - s0) * width;
    double* dst = (double*)(a.peer_base[k] + a.halo_off + parity * a.halo_stride) + (size_t)a.dst_off[k] * width;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < cnt; i += gridDim.x * kBlock) {
      const int node = i / width, kk = i - node * width;
      dst[i] = v[(size_t)a.send_idx[s0 + node] * width + kk];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.counter, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < a.npeer) {
    __threadfence_system();
    *(volatile unsigned long long*)(a.peer_base[threadIdx.x] + 8 * a.rank) = seq;  // publish
    if (threadIdx.x == 0) *a.counter = 0;
  }
  // ---- wait for the neighbours, then move their values into the ghost segment
  if (threadIdx.x < a.npeer)
    wait_flag((const volatile unsigned long long*)(a.local + 8 * a.peer_rank[threadIdx.x]), seq, a.error);
  __syncthreads();
  __threadfence_system();
  const volatile double* src = (const volatile double*)(a.local + a.halo_off + parity * a.halo_stride);
  double* ghost = v + (size_t)a.n_owned * width;
  const int64_t total = a.n_ghost * width;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < total; i += (int64_t)gridDim.x * kBlock) ghost[i] = src[i];
}

struct SlotTable { double* p[8]; int op[8]; };

// local reduction of each slot, exchange of the scalars with every rank, global reduction in rank order
__global__ void __launch_bounds__(kBlock)
k_p2p_allreduce(const P2PDev a, const SlotTable t, const int nslots, const int npart, const unsigned long long seq) {
  __shared__ double red[9];
  __shared__ bool last;
  const int parity = (int)(seq & 1);
  const int slot = blockIdx.x;
  double* p = t.p[slot];
  const int op = t.op[slot];
  double s = op == 0 ? 0.0 : (op == 1 ? INFINITY : -INFINITY);
  for (int i = threadIdx.x; i < npart; i += kBlock) {
    const double x = p[i];
    s = op == 0 ? s + x : (op == 1 ? fmin(s, x) : fmax(s, x));
  }
  s = op == 0 ? block_sum(s, red) : (op == 1 ? block_min(s, red) : block_max(s, red));
  if (threadIdx.x < a.world) {
    double* dst = (double*)(a.rank_base[threadIdx.x] + kFlagBytes) + ((size_t)parity * kMaxWorld + a.rank) * 8 + slot;
    *(volatile double*)dst = s;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.counter + 1, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < a.world) {
    __threadfence_system();
    *(volatile unsigned long long*)(a.rank_base[threadIdx.x] + 128 + 8 * a.rank) = seq;
    if (threadIdx.x == 0) a.counter[1] = 0;
  }
  if (threadIdx.x < a.world) wait_flag((const volatile unsigned long long*)(a.local + 128 + 8 * threadIdx.x), seq, a.error);
  __syncthreads();
  __threadfence_system();
  if (threadIdx.x == 0) {
    const volatile double* src = (const volatile double*)(a.local + kFlagBytes) + (size_t)parity * kMaxWorld * 8 + slot;
    double r = src[0];
    for (int q = 1; q < a.world; ++q) {
      const double x = src[(size_t)q * 8];
      r = op == 0 ? r + x : (op == 1 ? fmin(r, x) : fmax(r, x));
    }
    p[0] = r;
  }
}

// Low-latency variant: every scalar travels as two 8-byte words {32 data bits | 32-bit sequence tag}.
// An aligned 8-byte store is delivered atomically, so the tag doubles as the arrival flag: no
// __threadfence_system, no inter-CTA ticket, no separate flag store.  One CTA; warp w reduces slot w's
// partials (fixed lane-strided order + shuffle tree), lanes 0..world-1 store the two words into every
// rank's mailbox (own rank included), poll their own mailbox for the words of rank `lane`, and lane 0
// folds the `world` values in rank order -- bitwise identical on every rank.  (One CTA; thread (slot, q)
// handles slot `slot` to / from rank q.)
__global__ void __launch_bounds__(kBlock)
k_p2p_allreduce_ll(const P2PDev a, const SlotTable t, const int nslots, const int npart, const unsigned long long seq) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[9];
  __shared__ double local_sum[8];
  __shared__ double recv[8][kMaxWorld];
  const int parity = (int)(seq & 1);
  const unsigned int tag = (unsigned int)seq;
  for (int slot = 0; slot < nslots; ++slot) {  // fixed order: bitwise reproducible
    const double* p = t.p[slot];
    const int op = t.op[slot];
    double s = op == 0 ? 0.0 : (op == 1 ? INFINITY : -INFINITY);
    for (int i = threadIdx.x; i < npart; i += kBlock) {
      const double x = p[i];
      s = op == 0 ? s + x : (op == 1 ? fmin(s, x) : fmax(s, x));
    }
    s = op == 0 ? block_sum(s, red) : (op == 1 ? block_min(s, red) : block_max(s, red));
    if (threadIdx.x == 0) local_sum[slot] = s;
  }
  __syncthreads();
  const size_t ll_off = kFlagBytes + kRedBytes;
  if ((int)threadIdx.x < nslots * a.world) {
    const int slot = threadIdx.x / a.world, q = threadIdx.x - slot * a.world;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(local_sum[slot]);
    volatile unsigned long long* dst = (volatile unsigned long long*)(a.rank_base[q] + ll_off) +
                                       (((size_t)parity * kMaxWorld + a.rank) * 8 + slot) * 2;
    dst[0] = (bits & 0xffffffff00000000ull) | tag;
    dst[1] = (bits << 32) | tag;
    const volatile unsigned long long* src = (const volatile unsigned long long*)(a.local + ll_off) +
                                             (((size_t)parity * kMaxWorld + q) * 8 + slot) * 2;
    const long long t0 = clock64();
    unsigned long long hi = src[0], lo = src[1];
    while ((unsigned int)hi != tag || (unsigned int)lo != tag) {
      if (clock64() - t0 > 60000000000LL) { *a.error = 1; break; }
      hi = src[0]; lo = src[1];
    }
    recv[slot][q] = __longlong_as_double((long long)((hi & 0xffffffff00000000ull) | (lo >> 32)));
  }
  __syncthreads();
  if ((int)threadIdx.x < nslots) {
    const int op = t.op[threadIdx.x];
    double r = recv[threadIdx.x][0];
    for (int q = 1; q < a.world; ++q) {
      const double x = recv[threadIdx.x][q];
      r = op == 0 ? r + x : (op == 1 ? fmin(r, x) : fmax(r, x));
    }
    t.p[threadIdx.x][0] = r;
  }
}

static void p2p_setup(cfem_ctx* c) {
  const HostMesh& hm = c->hm;
  const int world = c->world, rank = c->rank, npeer = (int)hm.peer_rank.size();
  if (world > kMaxWorld) CFEM_THROW(-5, "peer-memory path supports at most 16 ranks");
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  P2P* pp = new P2P();
  P2PDev& d = pp->d;
  const int64_t ng = hm.nn - hm.n_owned;
  // The mailbox layout must be IDENTICAL on every rank (a sender computes addresses inside its
  // neighbours' mailboxes): size the two halo generations by the largest ghost count of any rank.
  int64_t* dng = nullptr;
  CUDA_OK(cudaMalloc((void**)&dng, sizeof(int64_t)));
  CUDA_OK(cudaMemcpy(dng, &ng, sizeof(int64_t), cudaMemcpyHostToDevice));
  NCCL_OK(nccl().AllReduce(dng, dng, 1, ncclInt64, ncclMax, comm, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  int64_t ng_max = 0;
  CUDA_OK(cudaMemcpy(&ng_max, dng, sizeof(int64_t), cudaMemcpyDeviceToHost));
  cudaFree(dng);
  d.halo_off = kFlagBytes + kRedBytes + kLLBytes;
  d.halo_stride = (((size_t)ng_max * 4 * sizeof(double)) + 255) / 256 * 256 + 256;
  d.ll_off = d.halo_off + 2 * d.halo_stride;
  d.ll_stride = (((size_t)ng_max * 2 * sizeof(unsigned long long)) + 255) / 256 * 256 + 256;
  const size_t bytes = d.ll_off + 2 * d.ll_stride;
  void* box = nullptr;
  CUDA_OK(cudaMalloc(&box, bytes));
  CUDA_OK(cudaMemset(box, 0, bytes));
  c->allocs.push_back(box);
  // ---- all-gather the IPC handles and the landing offsets (NCCL is the setup plumbing)
  cudaIpcMemHandle_t mine;
  CUDA_OK(cudaIpcGetMemHandle(&mine, box));
  int64_t box_offset = 0;   // of the mailbox inside the allocation the handle names
  {
    // driver entry point through the runtime: the library must load on machines without libcuda (CPU-only tests)
    typedef CUresult (*GetRange)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    CUdeviceptr base = 0;
    size_t size = 0;
    if (!fn || ((GetRange)fn)(&base, &size, (CUdeviceptr)box) != CUDA_SUCCESS) CFEM_THROW(-5, "cuMemGetAddressRange failed for the mailbox");
    box_offset = (int64_t)((CUdeviceptr)box - base);
  }
  std::vector<int32_t> land(world, -1);  // where rank q's values land in MY ghost segment
  for (int k = 0; k < npeer; ++k) land[hm.peer_rank[k]] = hm.recv_off[k] - (int32_t)hm.n_owned;
  const size_t rec = sizeof(cudaIpcMemHandle_t) + sizeof(int64_t) + world * sizeof(int32_t);
  std::vector<char> sendrec(rec), allrec(rec * world);
  memcpy(sendrec.data(), &mine, sizeof(mine));
  memcpy(sendrec.data() + sizeof(mine), &box_offset, sizeof(int64_t));
  memcpy(sendrec.data() + sizeof(mine) + sizeof(int64_t), land.data(), world * sizeof(int32_t));
  char *dsend = nullptr, *drecv = nullptr;
  CUDA_OK(cudaMalloc((void**)&dsend, rec));
  CUDA_OK(cudaMalloc((void**)&drecv, rec * world));
  CUDA_OK(cudaMemcpy(dsend, sendrec.data(), rec, cudaMemcpyHostToDevice));
  NCCL_OK(nccl().AllGather(dsend, drecv, rec, ncclChar, comm, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaMemcpy(allrec.data(), drecv, rec * world, cudaMemcpyDeviceToHost));
  cudaFree(dsend);
  cudaFree(drecv);
  d.local = (char*)box;
  d.world = world;
  d.rank = rank;
  d.npeer = npeer;
  d.n_owned = hm.n_owned;
  d.n_ghost = ng;
  for (int q = 0; q < world; ++q) {
    if (q == rank) { d.rank_base[q] = (char*)box; continue; }
    cudaIpcMemHandle_t h;
    int64_t off = 0;
    memcpy(&h, allrec.data() + q * rec, sizeof(h));
    memcpy(&off, allrec.data() + q * rec + sizeof(h), sizeof(int64_t));
    d.rank_base[q] = (char*)ipc_open(h) + off;
    pp->opened.emplace_back((const char*)&h, sizeof(h));
  }
  for (int k = 0; k < npeer; ++k) {
    const int q = hm.peer_rank[k];
    d.peer_base[k] = d.rank_base[q];
    d.peer_rank[k] = q;
    const int32_t* their_land = (const int32_t*)(allrec.data() + q * rec + sizeof(cudaIpcMemHandle_t) + sizeof(int64_t));
    d.dst_off[k] = their_land[rank];
    if (d.dst_off[k] < 0 && hm.send_ptr[k + 1] > hm.send_ptr[k]) CFEM_THROW(-5, "inconsistent halo lists between ranks");
  }
  CUDA_OK(cudaMalloc((void**)&pp->d_send_ptr, (npeer + 1) * sizeof(int32_t)));
  CUDA_OK(cudaMemcpy(pp->d_send_ptr, hm.send_ptr.data(), (npeer + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  c->allocs.push_back(pp->d_send_ptr);
  CUDA_OK(cudaMalloc((void**)&pp->d_peer_rank, (npeer + 1) * sizeof(int32_t)));
  CUDA_OK(cudaMemcpy(pp->d_peer_rank, hm.peer_rank.data(), npeer * sizeof(int32_t), cudaMemcpyHostToDevice));
  c->allocs.push_back(pp->d_peer_rank);
  d.send_ptr = pp->d_send_ptr;
  d.send_idx = c->d_send_idx;
  CUDA_OK(cudaMalloc((void**)&d.counter, 2 * sizeof(unsigned int)));
  CUDA_OK(cudaMemset(d.counter, 0, 2 * sizeof(unsigned int)));
  c->allocs.push_back(d.counter);
  CUDA_OK(cudaMalloc((void**)&d.tim, 12 * sizeof(unsigned long long)));
  CUDA_OK(cudaMemset(d.tim, 0, 12 * sizeof(unsigned long long)));
  c->allocs.push_back(d.tim);
  CUDA_OK(cudaMallocHost((void**)&pp->h_error, sizeof(int)));
  *pp->h_error = 0;
  d.error = pp->h_error;
  CUDA_OK(cudaMalloc((void**)&pp->d_dev, sizeof(P2PDev)));
  CUDA_OK(cudaMemcpy(pp->d_dev, &d, sizeof(P2PDev), cudaMemcpyHostToDevice));
  c->allocs.push_back(pp->d_dev);
  // every rank must have its mailbox mapped everywhere before the first exchange: the agreement all-reduce in
  // comm_setup_exchange is that barrier (kept out of here so that a rank failing above still takes part in it)
  c->p2p = pp;
}

// Peer mappings are released (reference counted per process) and the pinned error flag freed; the mailbox itself
// is part of c->allocs.  The caller destroys contexts collectively (every rank leaves its time loop before any
// rank frees a mailbox the others may still map) -- cfem_destroy documents that.
void comm_destroy_p2p(cfem_ctx* c) {
  if (!c->p2p) return;
  P2P* pp = (P2P*)c->p2p;
  for (const std::string& key : pp->opened) ipc_close(key);
  if (pp->h_error) cudaFreeHost(pp->h_error);
  delete pp;
  c->p2p = nullptr;
}

void comm_setup_exchange(cfem_ctx* c) {
  if (c->world == 1) return;
  const char* e = getenv("CFEM_COMM");
  if (e && std::string(e) == "nccl") return;
  // Peer memory needs CUDA IPC between all ranks (one node, peer access).  Where a rank cannot set it up, ALL ranks
  // must fall back to the NCCL data plane together: the verdict is agreed with a MIN all-reduce.
  int ok = 1;
  std::string why;
  try {
    p2p_setup(c);
  } catch (const Error& err) {
    ok = 0;
    why = err.msg;
    cudaGetLastError();
  }
  int* dflag = nullptr;
  CUDA_OK(cudaMalloc((void**)&dflag, sizeof(int)));
  CUDA_OK(cudaMemcpy(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice));
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  NCCL_OK(nccl().AllReduce(dflag, dflag, 1, ncclInt, ncclMin, comm, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  int all_ok = 0;
  CUDA_OK(cudaMemcpy(&all_ok, dflag, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(dflag);
  if (!all_ok) {
    comm_destroy_p2p(c);   // a rank that did succeed lets go of its mappings again
    if (!ok) fprintf(stderr, "cfem_b200: rank %d: peer-memory exchange unavailable (%s); all ranks use the NCCL path\n", c->rank, why.c_str());
  }
}

void comm_check(cfem_ctx* c) {
  if (c->p2p && *((P2P*)c->p2p)->h_error) {
    *((P2P*)c->p2p)->h_error = 0;
    CFEM_THROW(-5, "peer-memory exchange timed out waiting for a neighbour rank");
  }
}

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
  if (c->p2p) {
    P2P* pp = (P2P*)c->p2p;
    const int64_t work = std::max<int64_t>((int64_t)hm.send_idx.size(), hm.nn - hm.n_owned) * width;
    int g = (int)((work + kBlock - 1) / kBlock);
    if (g < 1) g = 1;
    if (g > 32) g = 32;  // all CTAs must be co-resident: they wait for the neighbours
    k_p2p_halo<<<g, kBlock, 0, c->stream>>>(pp->d, v, width, ++pp->halo_seq);
    LAUNCHED(c);
    c->halo_exchanges++;
    return;
  }
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

// Producer half of a halo exchange for a vector that a SpMV-type kernel is about to read: pushes
// the owned boundary values to the neighbours and returns where the consumer finds its ghosts.
// (NCCL mode: does the whole exchange and returns an empty GhostSrc.)
GhostSrc halo_push(cfem_ctx* c, double* v, bool gated, bool in_consumer) {
  GhostSrc g;
  if (c->world == 1) {
    // measurement hook: CFEM_FORCE_GHOST=1 runs the ghost-aware kernel variants on one GPU (no ghosts, no waiting)
    static const bool force = getenv("CFEM_FORCE_GHOST") != nullptr;
    if (force) { g.mbox = c->stage[0]; g.flags = (const char*)c->status; g.seq = 0; g.npeer = 0; }
    return g;
  }
  static const bool fused = !(getenv("CFEM_HALO") && std::string(getenv("CFEM_HALO")) == "exchange");
  if (!c->p2p || !fused) { halo_exchange(c, v, 1); return g; }
  const HostMesh& hm = c->hm;
  const int npeer = (int)hm.peer_rank.size();
  if (npeer == 0) return g;
  ProfScope ps(c, PROF_COMM);
  P2P* pp = (P2P*)c->p2p;
  const unsigned long long seq = ++pp->halo_seq;
  // CFEM_PUSH=kernel: always a separate push kernel; CFEM_PUSH=multi: that kernel with the multi-CTA ticket
  static const std::string push_mode = getenv("CFEM_PUSH") ? getenv("CFEM_PUSH") : "";
  const bool fuse = in_consumer && push_mode.empty() && hm.send_idx.size() <= 16384;
  if (!fuse) {
    int grid = (int)((hm.send_idx.size() + kBlock - 1) / kBlock);
    if (grid > 32) grid = 32;
    if (grid < 1 || (hm.send_idx.size() <= 8192 && push_mode != "multi")) grid = 1;
    k_p2p_push<<<grid, kBlock, 0, c->stream>>>(pp->d, v, 1, seq, gated ? c->status : nullptr);
    LAUNCHED(c);
  } else {
    g.pushdev = pp->d_dev;
    // low-latency words instead of values + fence + flag (CFEM_HALO_LL=0: the flag protocol)
    static const bool ll = !(getenv("CFEM_HALO_LL") && std::string(getenv("CFEM_HALO_LL")) == "0");
    if (ll) g.ll = (const unsigned long long*)(pp->d.local + pp->d.ll_off + (seq & 1) * pp->d.ll_stride);
  }
  c->halo_exchanges++;
  g.mbox = (const double*)(pp->d.local + pp->d.halo_off + (seq & 1) * pp->d.halo_stride);
  g.flags = pp->d.local;
  g.seq = seq;
  g.npeer = npeer;
  g.peer_rank = pp->d_peer_rank;
  g.error = pp->d.error;
  g.tim = pp->d.tim;
  return g;
}

// Reduce each listed partial array (npart entries) to its element 0 locally, then all-reduce
// those scalars over the ranks.  Returns the partial count consumers must use afterwards (1).
// op: 0 sum, 1 min, 2 max.
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
  if (c->p2p) {
    P2P* pp = (P2P*)c->p2p;
    static const bool ticket = getenv("CFEM_ALLREDUCE") && std::string(getenv("CFEM_ALLREDUCE")) == "ticket";
    if (ticket) k_p2p_allreduce<<<nslots, kBlock, 0, c->stream>>>(pp->d, t, nslots, npart, ++pp->red_seq);
    else launch_pdl(k_p2p_allreduce_ll, 1, kBlock, 0, c->stream, pp->d, t, nslots, npart, ++pp->red_seq);
    LAUNCHED(c);
    c->allreduces++;
    return 1;
  }
  k_reduce_slots<<<nslots, kBlock, 0, c->stream>>>(t, npart); LAUNCHED(c);
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  NCCL_OK(nccl().GroupStart());
  for (int k = 0; k < nslots; ++k)
    NCCL_OK(nccl().AllReduce(slots[k], slots[k], 1, ncclDouble, ops[k] == 0 ? ncclSum : (ops[k] == 1 ? ncclMin : ncclMax), comm, c->stream));
  NCCL_OK(nccl().GroupEnd());
  c->allreduces++;
  return 1;
}

void persist_comm_args(cfem_ctx* c, const P2PDev** dev, const char** mailbox, size_t* halo_off, size_t* halo_stride,
                       const int32_t** peer_rank, int* npeer, int** error, unsigned long long* halo_seq,
                       unsigned long long* red_seq, unsigned long long** tim) {
  *dev = nullptr; *mailbox = nullptr; *halo_off = 0; *halo_stride = 0; *peer_rank = nullptr; *npeer = 0; *error = nullptr;
  *halo_seq = 0; *red_seq = 0; *tim = nullptr;
  if (c->world == 1 || !c->p2p) return;
  P2P* pp = (P2P*)c->p2p;
  *dev = pp->d_dev;
  *mailbox = pp->d.local;
  *halo_off = pp->d.ll_off;       // the persistent kernels use the low-latency halo words only
  *halo_stride = pp->d.ll_stride;
  *peer_rank = pp->d_peer_rank;
  *npeer = pp->d.npeer;
  *error = pp->d.error;
  *halo_seq = pp->halo_seq;
  *red_seq = pp->red_seq;
  *tim = pp->d.tim;
}

void persist_comm_advance(cfem_ctx* c, int64_t halo_exchanges, int64_t allreduces) {
  if (c->world == 1 || !c->p2p) return;
  P2P* pp = (P2P*)c->p2p;
  pp->halo_seq += (unsigned long long)halo_exchanges;
  pp->red_seq += (unsigned long long)allreduces;
  c->halo_exchanges += halo_exchanges;
  c->allreduces += allreduces;
}

void persist_seq_reserve(cfem_ctx* c, int64_t halo_exchanges, int64_t allreduces) {
  if (c->world == 1 || !c->p2p) return;
  P2P* pp = (P2P*)c->p2p;
  pp->halo_seq += (unsigned long long)halo_exchanges;
  pp->red_seq += (unsigned long long)allreduces;
}

void persist_comm_count(cfem_ctx* c, int64_t halo_exchanges, int64_t allreduces) {
  if (c->world == 1 || !c->p2p) return;
  c->halo_exchanges += halo_exchanges;
  c->allreduces += allreduces;
}

// cycles -> microseconds with the device's nominal SM clock; reset: start a new accounting interval
void comm_timers(cfem_ctx* c, double* out12, bool reset) {
  for (int k = 0; k < 12; ++k) out12[k] = 0.0;
  if (c->world == 1 || !c->p2p) return;
  P2P* pp = (P2P*)c->p2p;
  unsigned long long h[12];
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaMemcpy(h, pp->d.tim, sizeof(h), cudaMemcpyDeviceToHost));
  if (reset) CUDA_OK(cudaMemset(pp->d.tim, 0, sizeof(h)));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
  const double us = khz > 0 ? 1e3 / (double)khz : 0.0;
  for (int k = 0; k < 12; ++k) out12[k] = (k == 1 || k == 4 || k == 7 || k == 9) ? (double)h[k] : (double)h[k] * us;
}

bool fin_available(const cfem_ctx* c) { return c->world == 1 || c->p2p != nullptr; }

Fin make_fin(cfem_ctx* c) {
  Fin f;
  f.counter = (unsigned int*)(c->status + 4);   // zeroed at creation, reset by the last CTA of every launch
  if (c->world > 1 && c->p2p) {
    P2P* pp = (P2P*)c->p2p;
    f.dev = pp->d_dev;
    f.seq = ++pp->red_seq;
    c->allreduces++;
  }
  return f;
}

int allreduce_sum1(cfem_ctx* c, double* slot, int npart) {
  double* s[1] = {slot};
  const int op[1] = {0};
  return allreduce_partials(c, 1, s, op, npart);
}

}  // namespace cfem
