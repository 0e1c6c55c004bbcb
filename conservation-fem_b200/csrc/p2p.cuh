// Device-side view of the peer-memory exchange (comm.cu) shared with the SpMV-type kernels.
#pragma once
#include "internal.h"

namespace cfem {

constexpr int kMaxWorld = 16;
constexpr size_t kFlagBytes = 256;                                 // halo_flag[16] | red_flag[16]  (uint64)
constexpr size_t kRedBytes = 2 * kMaxWorld * 8 * sizeof(double);    // [parity][src rank][8 slots]  (ticket all-reduce)
constexpr size_t kLLBytes = 2 * kMaxWorld * 8 * 2 * sizeof(unsigned long long);  // [parity][src rank][8 slots][hi|lo word]

struct P2PDev {  // passed to kernels by value
  char* peer_base[kMaxWorld];   // by peer INDEX (halo) ...
  char* rank_base[kMaxWorld];   // ... and by RANK (reductions; own rank -> local mailbox)
  char* local;
  int32_t dst_off[kMaxWorld];   // by peer index: where my values land in that peer's ghost segment (nodes)
  int32_t peer_rank[kMaxWorld];
  int npeer, world, rank;
  size_t halo_off, halo_stride;  // bytes
  const int32_t *send_ptr, *send_idx;
  unsigned int* counter;         // 2 block counters
  int* error;                    // pinned host flag
  int64_t n_owned, n_ghost;
};

// Where a SpMV-type kernel finds the ghost entries of its input vector.
//  mbox == nullptr : in the vector itself (one GPU, or after a full halo_exchange)
//  mbox != nullptr : in this rank's mailbox, valid once every neighbour has published `seq`;
//                    only CTAs that reach a boundary tile (index >= n_interior in tile_order) wait.
struct GhostSrc {
  const double* mbox = nullptr;
  const char* flags = nullptr;          // local mailbox base (halo flags by rank, uint64 each)
  unsigned long long seq = 0;
  const int32_t* peer_rank = nullptr;   // device array (no by-value arrays: they would force the struct onto the stack)
  int npeer = 0;
  int* error = nullptr;
  // producer half fused into the consumer: when set, CTA 0 of the SpMV-type kernel stores this rank's
  // boundary values into the neighbours' mailboxes (and publishes seq) while the other CTAs already
  // work on interior tiles; the kernel is then launched with one extra CTA.
  const P2PDev* pushdev = nullptr;      // device copy of the exchange tables
};

__device__ __forceinline__ bool wait_flag(const volatile unsigned long long* f, unsigned long long seq, int* error) {
  const long long t0 = clock64();
  while (*f < seq) {
    if (clock64() - t0 > 60000000000LL) { *error = 1; return false; }
  }
  return true;
}

// The whole producer half of a halo exchange, run by ONE CTA: owned boundary values of v -> the
// neighbours' mailboxes (generation seq & 1), then the sequence number.  `gate` set: skip the values
// (the consumers are gated off too) but still publish, so the sequence stays aligned between ranks.
__device__ __forceinline__ void push_cta(const P2PDev* __restrict__ a, const double* __restrict__ v,
                                         const unsigned long long seq, const bool gate) {
  const int npeer = a->npeer;
  if (!gate) {
    const size_t gen = a->halo_off + (size_t)(seq & 1) * a->halo_stride;
    for (int k = 0; k < npeer; ++k) {
      const int s0 = a->send_ptr[k], cnt = a->send_ptr[k + 1] - s0;
      double* dst = (double*)(a->peer_base[k] + gen) + a->dst_off[k];
      const int32_t* __restrict__ idx = a->send_idx + s0;
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = v[idx[i]];
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < npeer) *(volatile unsigned long long*)(a->peer_base[threadIdx.x] + 8 * a->rank) = seq;
}

// called by all threads of a CTA before it touches a boundary tile
__device__ __forceinline__ void ghost_wait(const GhostSrc& g) {
  if ((int)threadIdx.x < g.npeer)
    wait_flag((const volatile unsigned long long*)(g.flags + 8 * g.peer_rank[threadIdx.x]), g.seq, g.error);
  __syncthreads();
  __threadfence_system();
}

}  // namespace cfem
