// Device-side view of the peer-memory exchange (comm.cu) shared with the SpMV-type kernels.
#pragma once
#include "device_utils.cuh"

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
  size_t ll_off, ll_stride;      // low-latency halo area: two generations of two 64-bit words per ghost value
  const int32_t *send_ptr, *send_idx;
  unsigned int* counter;         // 2 block counters
  int* error;                    // pinned host flag
  int64_t n_owned, n_ghost;
  // wait accounting (SM cycles), always on: [0] sum [1] count [2] max of ghost_wait per waiting CTA;
  // [3] sum [4] count [5] max of the cross-rank part of an in-kernel all-reduce; [6] sum [7] count of the time the
  // first worker CTA of the persistent solver spends in grid barriers (local wait + reduction + all-reduce)
  unsigned long long* tim;
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
  unsigned long long* tim = nullptr;    // wait accounting, see P2PDev::tim
  // Low-latency halo (default for the fused push): ghost value g of this exchange sits in ll[2g], ll[2g+1] as two
  // self-validating words {32 data bits | 32-bit sequence tag}; the reader polls the words it needs, nobody waits for
  // a flag and the producer needs no system-scope fence (see ll_store / ll_load).
  const unsigned long long* ll = nullptr;
};

__device__ __forceinline__ bool wait_flag(const volatile unsigned long long* f, unsigned long long seq, int* error) {
  const long long t0 = clock64();
  while (*f < seq) {
    if (clock64() - t0 > 60000000000LL) { *error = 1; return false; }
  }
  return true;
}

// ---- low-latency halo words ------------------------------------------------------------------------------------------
// A double travels as two aligned 8-byte words, each carrying 32 of its bits and the 32-bit sequence tag of the
// exchange.  An aligned 8-byte store is delivered atomically, so a word whose tag matches is complete: the tag is the
// arrival flag.  Compared with values + __threadfence_system + flag this takes one NVLink store latency (~2 us)
// instead of ~9 us from the producer's first load to the consumer's first use.  Tags start at 1; the area is zeroed.
__device__ __forceinline__ void ll_store(unsigned long long* dst2, const double v, const unsigned int tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  ((volatile unsigned long long*)dst2)[0] = (bits & 0xffffffff00000000ull) | tag;
  ((volatile unsigned long long*)dst2)[1] = (bits << 32) | tag;
}
__device__ __forceinline__ double ll_load(const unsigned long long* src2, const unsigned int tag, int* error) {
  const volatile unsigned long long* s = (const volatile unsigned long long*)src2;
  unsigned long long hi = s[0], lo = s[1];
  if ((unsigned int)hi != tag || (unsigned int)lo != tag) {
    const long long t0 = clock64();
    do {
      if (clock64() - t0 > 40000000000LL) { if (error) *error = 1; break; }
      hi = s[0]; lo = s[1];
    } while ((unsigned int)hi != tag || (unsigned int)lo != tag);
  }
  return __longlong_as_double((long long)((hi & 0xffffffff00000000ull) | (lo >> 32)));
}

// low-latency producer half, run by ONE CTA: f(node) for every owned boundary node -> the neighbours' LL areas
template <class F>
__device__ __forceinline__ void push_ll(const P2PDev* __restrict__ a, const unsigned long long seq, const F f) {
  const int npeer = a->npeer;
  const size_t gen = a->ll_off + (size_t)(seq & 1) * a->ll_stride;
  const unsigned int tag = (unsigned int)seq;
  for (int k = 0; k < npeer; ++k) {
    const int s0 = a->send_ptr[k], cnt = a->send_ptr[k + 1] - s0;
    unsigned long long* dst = (unsigned long long*)(a->peer_base[k] + gen) + 2 * (size_t)a->dst_off[k];
    const int32_t* __restrict__ idx = a->send_idx + s0;
    // four nodes per thread in flight: the index -> value chains are independent, and a loop of one chain at a time
    // (two dependent L2 loads per node) made the push of a 4 K-node boundary take ~10 us of a ~24 us kernel
    int i = threadIdx.x;
    const int step = blockDim.x;
    for (; i + 3 * step < cnt; i += 4 * step) {
      const int n0 = idx[i], n1 = idx[i + step], n2 = idx[i + 2 * step], n3 = idx[i + 3 * step];
      const double v0 = f(n0), v1 = f(n1), v2 = f(n2), v3 = f(n3);
      ll_store(dst + 2 * (size_t)i, v0, tag);
      ll_store(dst + 2 * (size_t)(i + step), v1, tag);
      ll_store(dst + 2 * (size_t)(i + 2 * step), v2, tag);
      ll_store(dst + 2 * (size_t)(i + 3 * step), v3, tag);
    }
    for (; i < cnt; i += step) ll_store(dst + 2 * (size_t)i, f(idx[i]), tag);
  }
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
  const long long t0 = clock64();
  if ((int)threadIdx.x < g.npeer) {
    wait_flag((const volatile unsigned long long*)(g.flags + 8 * g.peer_rank[threadIdx.x]), g.seq, g.error);
    __threadfence_system();   // by the observing threads only; the barrier below extends the order to the CTA
  }                           // (a system-scope fence in all 256 threads cost ~4 us per wait)
  __syncthreads();
  if (g.tim && threadIdx.x == 0) {
    const unsigned long long dt = (unsigned long long)(clock64() - t0);
    atomicAdd(g.tim + 0, dt);
    atomicAdd(g.tim + 1, 1ull);
    atomicMax(g.tim + 2, dt);
  }
}

// ---- in-kernel finalisation of reductions -------------------------------------------------------------------------
// A kernel whose CTAs each produce partial sums hands them to fin_reduce: the LAST CTA to arrive (ticket counter)
// adds the partials up in a fixed order and -- in a distributed context -- exchanges the totals with every rank
// through the tagged-word slots of the peer mailboxes (same protocol as k_p2p_allreduce_ll) and folds them in rank
// order, so all ranks end up with bitwise identical scalars.  The consumer kernel then reads ready scalars: no
// finalise launch, no all-reduce launch, no per-CTA re-reduction of ~1000 partials in the consumer's prologue.
struct Fin {
  unsigned int* counter = nullptr;   // device ticket counter, zero between launches (null: finalisation off)
  const P2PDev* dev = nullptr;       // non-null: all-reduce over the ranks
  unsigned long long seq = 0;        // all-reduce sequence number of this launch (same on every rank)
};
template <int NS>
struct Slots { double* p[NS]; };

// sums[k] (shared memory) <- sum over ranks, rank order; called by all threads of ONE CTA (blockDim.x >= NS * world)
template <int NS>
__device__ __forceinline__ void cta_allreduce(const P2PDev* __restrict__ a, const unsigned long long seq, double* sums) {
  __shared__ double recv[NS][kMaxWorld];
  const long long t_in = clock64();
  const int world = a->world, rank = a->rank;
  const int parity = (int)(seq & 1);
  const unsigned int tag = (unsigned int)seq;
  const size_t ll_off = kFlagBytes + kRedBytes;
  if ((int)threadIdx.x < NS * world) {
    const int slot = threadIdx.x / world, q = threadIdx.x - slot * world;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(sums[slot]);
    volatile unsigned long long* dst = (volatile unsigned long long*)(a->rank_base[q] + ll_off) +
                                       (((size_t)parity * kMaxWorld + rank) * 8 + slot) * 2;
    dst[0] = (bits & 0xffffffff00000000ull) | tag;
    dst[1] = (bits << 32) | tag;
    const volatile unsigned long long* src = (const volatile unsigned long long*)(a->local + ll_off) +
                                             (((size_t)parity * kMaxWorld + q) * 8 + slot) * 2;
    const long long t0 = clock64();
    unsigned long long hi = src[0], lo = src[1];
    while ((unsigned int)hi != tag || (unsigned int)lo != tag) {
      if (clock64() - t0 > 60000000000LL) { *a->error = 1; break; }
      hi = src[0]; lo = src[1];
    }
    recv[slot][q] = __longlong_as_double((long long)((hi & 0xffffffff00000000ull) | (lo >> 32)));
  }
  __syncthreads();
  if ((int)threadIdx.x < NS) {
    double r = recv[threadIdx.x][0];
    for (int q = 1; q < world; ++q) r += recv[threadIdx.x][q];
    sums[threadIdx.x] = r;
  }
  __syncthreads();
  if (a->tim && threadIdx.x == 0) {
    const unsigned long long dt = (unsigned long long)(clock64() - t_in);
    atomicAdd(a->tim + 3, dt);
    atomicAdd(a->tim + 4, 1ull);
    atomicMax(a->tim + 5, dt);
  }
}

// Every thread of every participating CTA calls this after thread 0 of the CTA has stored part.p[k][cta].
// Returns true in all threads of the last CTA, with sums[0..NS) (shared memory) holding the global totals.
template <int NS>
__device__ __forceinline__ bool fin_reduce(const Fin& fin, const Slots<NS>& part, const int ncta, double* red /*>= 9*/,
                                           double* sums /*>= NS, shared*/) {
  __shared__ bool fin_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                               // this CTA's partials are visible before its ticket is
    const unsigned int t = atomicAdd(fin.counter, 1u);
    fin_last = (t == (unsigned int)ncta - 1u);
  }
  __syncthreads();
  if (!fin_last) return false;
  __threadfence();
  // all NS slots in one pass: the loads of the slots are independent, the NS tree reductions share their barriers.
  // Per slot the order is the one reduce_partials uses (thread-strided partial sums, shuffle tree, warp order).
  double s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.0;
  for (int i = threadIdx.x; i < ncta; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] += __ldcg(part.p[k] + i);
  }
  {
    __shared__ double wsum[NS][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const double w = warp_sum(s[k]);
      if (lane == 0) wsum[k][wid] = w;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        double t = lane < (kBlock / 32) ? wsum[k][lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sums[k] = t;
      }
    }
  }
  __syncthreads();
  if (fin.dev) cta_allreduce<NS>(fin.dev, fin.seq, sums);
  if (threadIdx.x == 0) *fin.counter = 0;          // ready for the next launch (stream ordered)
  return true;
}

}  // namespace cfem
