// Persistent BiCGStab: ONE cooperative launch runs every iteration of a Crank-Nicolson / Jacobian solve.
//
// Replaces the linear solves the reference does by LU (Code/KPP/KPP_exact.py:147-154, dolfinx NewtonSolver's
// KSP preonly + PC lu; Code/Linear_advection/RV_node.py:131-134).  The launch-per-phase forms in linalg.cu spend
// a third of a ~100 us iteration on kernel boundaries (drain, ramp-up, reduction tails, consumer prologues); here a
// grid of resident CTAs walks the phases of an iteration separated by grid-wide barriers that double as the
// reduction points:
//
//   P1  v = D^-1 A p                                   partial (rhat,v)            --- barrier + reduce: alpha
//   P3  s = r - alpha v formed while STAGING (own rows and the tile's external columns, so s is never stored and
//       no barrier separates the update from the product);  t = D^-1 A s
//                                                      partials (t,s) (t,t) (rhat,t) (rhat,s)
//                                                                                  --- barrier + reduce: omega, rho', beta
//   P4  x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v)
//                                                      partial ||r||^2             --- barrier + reduce: verdict
//
// A CTA owns the same tiles (rows) in every phase, so a row's own-entries are always produced and consumed by the
// same thread; values of OTHER CTAs' rows (the external columns of a tile) are read with ld.global.cg after the
// barrier that follows their production -- L1 is not coherent across SMs and, unlike at a kernel boundary, is not
// invalidated between phases.  The last CTA to arrive at a barrier adds the per-CTA partials in a fixed order and,
// in a distributed context, exchanges the totals with the other ranks through the peer mailboxes (cta_allreduce)
// before it releases the grid: reductions stay bitwise reproducible and identical on every rank.  Halo values of p
// (P1) and s (P3) are pushed by a dedicated communication CTA at the start of the phase and awaited only by CTAs that
// reach a boundary tile (tile_order keeps interior tiles first).  Every wait is bounded and raises an error flag.
#include <cstdlib>
#include <string>

#include "device_utils.cuh"
#include "launch.h"
#include "p2p.cuh"

namespace cfem {

#ifndef CFEM_PERSIST_MINB
#define CFEM_PERSIST_MINB 4   // 64 registers: no spills; measured 2.58 vs 2.76 (5 CTAs, 48 regs, spills) vs 2.85 ms per step (6 CTAs)
#endif

// device scalars / partial slots shared with linalg.cu
enum { PS_BB = 3, PS_RELRES = 7, PS_RR = 8, PS_D0 = 16, PS_RHO0 = 21 };
enum { PP_PQ = 0, PP_RR = 3, PP_A = 5 };

struct BicgArgs {
  int64_t no;
  int ntiles, n_interior, ext_cap;
  const int32_t *tile_order, *tile_node, *rowptr, *tile_extptr, *tile_ext;
  const uint16_t* lc16;
  const double *vals, *dinv, *rhat;
  double *x, *r, *p, *v, *t;
  double *part, *scalars;
  int32_t* status;          // [0] verdict, [1] iterations, [3] barrier time-out flag
  unsigned int* bar;        // [0] arrivals, [1] generation (both zero at launch)
  double rtol2, atol2;
  int max_it;
  // distributed (dev == nullptr on one GPU)
  const P2PDev* dev;
  const char* mailbox;      // local mailbox base
  size_t halo_off, halo_stride;
  const int32_t* peer_rank;
  int npeer;
  int* error;
  unsigned long long halo_seq0, red_seq0;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier that also finishes NS reductions.  Every thread of every CTA calls it after thread 0 of each
// worker CTA stored part.p[k][worker].  On return sums[0..NS) (shared) hold the totals in every CTA.
template <int NS>
__device__ __forceinline__ void grid_reduce(const BicgArgs& a, const int nblk, const int nwork, const Slots<NS>& part,
                                            const unsigned long long rseq, double* out, double* sums, unsigned int& gen) {
  __shared__ bool s_last;
  ++gen;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(a.bar, 1u);
    s_last = (t == (unsigned int)nblk - 1u);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double s[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] = 0.0;
    for (int i = threadIdx.x; i < nwork; i += kBlock) {
#pragma unroll
      for (int k = 0; k < NS; ++k) s[k] += __ldcg(part.p[k] + i);
    }
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
    __syncthreads();
    if (a.dev) cta_allreduce<NS>(a.dev, rseq, sums);
    if ((int)threadIdx.x < NS) out[threadIdx.x] = sums[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      a.bar[0] = 0;
      __threadfence();
      atomicExch(a.bar + 1, gen);   // release
    }
  } else {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while ((int)(ld_acquire_u32(a.bar + 1) - gen) < 0) {
        // bounded (~10 s): raise the flag and fall through; once it is up every later barrier falls through at once
        if (clock64() - t0 > 20000000000LL || *(volatile int32_t*)(a.status + 3)) {
          a.status[3] = 1;
          if (a.error) *a.error = 1;
          break;
        }
      }
      __threadfence();
    }
    __syncthreads();
    if ((int)threadIdx.x < NS) sums[threadIdx.x] = __ldcg(out + threadIdx.x);
  }
  __syncthreads();
}

// owned boundary values f(node) -> the neighbours' mailboxes (generation seq & 1), then the sequence number
template <class F>
__device__ __forceinline__ void push_values(const P2PDev* __restrict__ a, const unsigned long long seq, const F f) {
  const int npeer = a->npeer;
  const size_t gen = a->halo_off + (size_t)(seq & 1) * a->halo_stride;
  for (int k = 0; k < npeer; ++k) {
    const int s0 = a->send_ptr[k], cnt = a->send_ptr[k + 1] - s0;
    double* dst = (double*)(a->peer_base[k] + gen) + a->dst_off[k];
    const int32_t* __restrict__ idx = a->send_idx + s0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = f(idx[i]);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < npeer) *(volatile unsigned long long*)(a->peer_base[threadIdx.x] + 8 * a->rank) = seq;
}

// One SpMV-type phase over this CTA's tiles.  MODE 0: x = p, y = v, acc[0] += rhat.y.  MODE 1: x = r - alpha v,
// y = t, acc = {(t,s), (t,t), (rhat,t), (rhat,s)}.
template <int MODE, bool GHOST>
__device__ __forceinline__ void spmv_phase(const BicgArgs& a, const int wid, const int nwork, const double alpha,
                                           const unsigned long long hseq, double* prod, double* xs, int32_t* rp,
                                           double* acc) {
  const int tid = threadIdx.x;
  const int64_t no = a.no;
  const double* const mbox_shifted =
      GHOST ? (const double*)(a.mailbox + a.halo_off + (size_t)(hseq & 1) * a.halo_stride) - no : nullptr;
  bool waited = false;
  for (int t = wid; t < a.ntiles; t += nwork) {
    const int tile = GHOST ? a.tile_order[t] : t;
    const int n0 = a.tile_node[tile], nrows = a.tile_node[tile + 1] - n0;
    const int e0 = a.tile_extptr[tile], ne = a.tile_extptr[tile + 1] - e0;
    const int start = a.rowptr[n0], cnt = a.rowptr[n0 + nrows] - start;
    for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = a.rowptr[n0 + i] - start;
    const double* __restrict__ v = a.vals + start;
    const uint16_t* __restrict__ lc = a.lc16 + start;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    if (tid < cnt) { v0 = v[tid]; l0 = lc[tid]; }
    if (tid + kTileNodes < cnt) { v1 = v[tid + kTileNodes]; l1 = lc[tid + kTileNodes]; }
    if (tid + 2 * kTileNodes < cnt) { v2 = v[tid + 2 * kTileNodes]; l2 = lc[tid + 2 * kTileNodes]; }
    if (tid + 3 * kTileNodes < cnt) { v3 = v[tid + 3 * kTileNodes]; l3 = lc[tid + 3 * kTileNodes]; }
    if (GHOST && t >= a.n_interior && !waited) {
      GhostSrc g;
      g.flags = a.mailbox; g.seq = hseq; g.peer_rank = a.peer_rank; g.npeer = a.npeer; g.error = a.error;
      ghost_wait(g);
      waited = true;
    }
    // ---- stage x: own rows, then the external columns (other CTAs' rows: L2 loads, see the file header)
    double xown = 0.0, rh = 0.0, di = 0.0;
    if (tid < nrows) {
      const int row = n0 + tid;
      xown = MODE == 0 ? __ldcg(a.p + row) : __ldcg(a.r + row) - alpha * __ldcg(a.v + row);
      xs[tid] = xown;
      rh = a.rhat[row];
      di = a.dinv[row];
    }
    for (int e = tid; e < ne; e += kTileNodes) {
      const int cc = a.tile_ext[e0 + e];
      double val;
      if (GHOST && cc >= no) val = mbox_shifted[cc];
      else val = MODE == 0 ? __ldcg(a.p + cc) : __ldcg(a.r + cc) - alpha * __ldcg(a.v + cc);
      xs[kTileNodes + e] = val;
    }
    __syncthreads();
    if (tid < cnt) prod[tid] = v0 * xs[l0];
    if (tid + kTileNodes < cnt) prod[tid + kTileNodes] = v1 * xs[l1];
    if (tid + 2 * kTileNodes < cnt) prod[tid + 2 * kTileNodes] = v2 * xs[l2];
    if (tid + 3 * kTileNodes < cnt) prod[tid + 3 * kTileNodes] = v3 * xs[l3];
    for (int p = tid + 4 * kTileNodes; p < cnt; p += kTileNodes) prod[p] = v[p] * xs[lc[p]];
    __syncthreads();
    if (tid < nrows) {
      const int row = n0 + tid;
      double s = 0.0;
      for (int k = rp[tid]; k < rp[tid + 1]; ++k) s += prod[k];
      s *= di;
      if (MODE == 0) {
        a.v[row] = s;
        acc[0] += rh * s;
      } else {
        a.t[row] = s;
        acc[0] += s * xown;
        acc[1] += s * s;
        acc[2] += rh * s;
        acc[3] += rh * xown;
      }
    }
    __syncthreads();
  }
}

template <bool GHOST>
__global__ void __launch_bounds__(kTileNodes, CFEM_PERSIST_MINB)
k_bicg_persist(const BicgArgs a) {
  extern __shared__ double ps_smem[];
  double* const prod = ps_smem;               // [kTileNnzCap]
  double* const xs = ps_smem + kTileNnzCap;   // [kTileNodes + ext_cap]
  __shared__ int32_t rp[kTileNodes + 1];
  __shared__ double red[9];
  __shared__ double sums[4];
  const int tid = threadIdx.x;
  const int nblk = gridDim.x;
  const bool comm_cta = GHOST && blockIdx.x == 0;       // pushes halo values, owns no tile
  const int nwork = GHOST ? nblk - 1 : nblk;
  const int wid = GHOST ? (int)blockIdx.x - 1 : (int)blockIdx.x;
  unsigned int gen = 0;
  unsigned long long hseq = a.halo_seq0, rseq = a.red_seq0;
  if (__ldcg(a.status) != 0) return;                    // the initial residual already met the tolerance (uniform)
  double rho = __ldcg(a.scalars + PS_RHO0);
  const double bb = __ldcg(a.scalars + PS_BB);
  double* const out = a.scalars + PS_D0;
  double* const pPQ = a.part + (size_t)PP_PQ * kMaxPartials;
  double* const pA = a.part + (size_t)PP_A * kMaxPartials;
  double* const pRR = a.part + (size_t)PP_RR * kMaxPartials;

  for (int it = 0; it < a.max_it; ++it) {
    // ------------------------------------------------ P1: v = D^-1 A p
    ++hseq;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (comm_cta) {
      const double* p = a.p;
      push_values(a.dev, hseq, [p](int node) { return __ldcg(p + node); });
    } else {
      spmv_phase<0, GHOST>(a, wid, nwork, 0.0, hseq, prod, xs, rp, acc);
      const double s0 = block_sum(acc[0], red);
      if (tid == 0) pPQ[wid] = s0;
    }
    {
      Slots<1> sl;
      sl.p[0] = pPQ;
      grid_reduce<1>(a, nblk, nwork, sl, ++rseq, out, sums, gen);
    }
    const double rv = sums[0];
    const double alpha = rv != 0.0 ? rho / rv : 0.0;
    // ------------------------------------------------ P3: t = D^-1 A (r - alpha v)
    ++hseq;
    if (comm_cta) {
      const double *r = a.r, *v = a.v;
      push_values(a.dev, hseq, [r, v, alpha](int node) { return __ldcg(r + node) - alpha * __ldcg(v + node); });
    } else {
      acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
      spmv_phase<1, GHOST>(a, wid, nwork, alpha, hseq, prod, xs, rp, acc);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double sk = block_sum(acc[k], red);
        if (tid == 0) pA[(size_t)k * kMaxPartials + wid] = sk;
      }
    }
    {
      Slots<4> sl;
#pragma unroll
      for (int k = 0; k < 4; ++k) sl.p[k] = pA + (size_t)k * kMaxPartials;
      grid_reduce<4>(a, nblk, nwork, sl, ++rseq, out, sums, gen);
    }
    const double ts = sums[0], tt = sums[1], rt = sums[2], rs = sums[3];
    const double omega = tt > 0.0 ? ts / tt : 0.0;
    const double rho_new = rs - omega * rt;   // (rhat, s - omega t)
    // omega == 0 only when s vanished (the alpha half-step solved the system): then r = s = 0 below and the verdict
    // is "converged"; beta must not turn that into 0 * inf
    const double beta = (omega != 0.0 && rho != 0.0) ? (rho_new / rho) * (alpha / omega) : 0.0;
    // ------------------------------------------------ P4: x, r, p on this CTA's rows
    double rr = 0.0;
    if (!comm_cta) {
      for (int t = wid; t < a.ntiles; t += nwork) {
        const int tile = GHOST ? a.tile_order[t] : t;
        const int n0 = a.tile_node[tile], nrows = a.tile_node[tile + 1] - n0;
        if (tid < nrows) {
          const int row = n0 + tid;
          const double vi = __ldcg(a.v + row), pi = __ldcg(a.p + row), ti = __ldcg(a.t + row);
          const double si = __ldcg(a.r + row) - alpha * vi;
          a.x[row] = __ldcg(a.x + row) + alpha * pi + omega * si;
          const double ri = si - omega * ti;
          a.r[row] = ri;
          a.p[row] = ri + beta * (pi - omega * vi);
          rr += ri * ri;
        }
      }
      rr = block_sum(rr, red);
      if (tid == 0) pRR[wid] = rr;
    }
    {
      Slots<1> sl;
      sl.p[0] = pRR;
      grid_reduce<1>(a, nblk, nwork, sl, ++rseq, out, sums, gen);
    }
    const double grr = sums[0];
    int verdict = 0;
    if (!(grr == grr)) verdict = 2;
    else if (grr <= a.rtol2 * bb || grr <= a.atol2) verdict = 1;
    else if (!(beta == beta)) verdict = 2;   // breakdown: the next direction is not finite
    if (blockIdx.x == 0 && tid == 0) {
      a.scalars[PS_RR] = grr;
      a.scalars[PS_RELRES] = bb > 0.0 ? sqrt(grr / bb) : sqrt(grr);
      a.status[1] = it + 1;
      a.status[0] = verdict;
    }
    rho = rho_new;
    if (verdict != 0 || *(volatile int32_t*)(a.status + 3)) break;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
struct PersistPlan { int grid = 0; size_t smem = 0; bool ok = false, tried = false; };

template <bool GHOST>
static bool plan_one(cfem_ctx* c, PersistPlan& pl) {
  pl.smem = sizeof(double) * ((size_t)kTileNnzCap + kTileNodes + c->dm.ext_cap);
  if (cudaFuncSetAttribute(k_bicg_persist<GHOST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) { cudaGetLastError(); return false; }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bicg_persist<GHOST>, kTileNodes, pl.smem) != cudaSuccess || occ < 1) { cudaGetLastError(); return false; }
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
  if (!coop) return false;
  int64_t grid = (int64_t)occ * c->sm_count;
  const int64_t want = c->dm.ntiles + (GHOST ? 1 : 0);
  if (grid > want) grid = want;
  if (grid > kMaxPartials) grid = kMaxPartials;
  if (grid < (GHOST ? 2 : 1)) return false;
  pl.grid = (int)grid;
  return true;
}

bool bicgstab_persist_available(cfem_ctx* c) {
  static const bool off = getenv("CFEM_BICGSTAB") && std::string(getenv("CFEM_BICGSTAB")) != "persist";
  if (off) return false;
  PersistPlan* pl = (PersistPlan*)c->persist_plan;
  if (!pl) { pl = new PersistPlan(); c->persist_plan = pl; }
  if (!pl->tried) {
    pl->tried = true;
    pl->ok = c->world > 1 ? plan_one<true>(c, *pl) : plan_one<false>(c, *pl);
  }
  return pl->ok;
}

void persist_plan_free(cfem_ctx* c) {
  delete (PersistPlan*)c->persist_plan;
  c->persist_plan = nullptr;
}

// The loop part of a BiCGStab solve after r, rhat, p, rho_0, ||b|| and the verdict on x_0 are in place (k_bm_init).
// Returns after the launch; the caller polls the device flag (one host sync per solve).
void launch_bicg_persist(cfem_ctx* c, const Matrix& A, const double* rhat, double* x, double* r, double* p, double* v,
                         double* t, double rtol2, double atol2, int max_it) {
  PersistPlan* pl = (PersistPlan*)c->persist_plan;
  BicgArgs a{};
  const DevMesh& m = c->dm;
  a.no = m.no; a.ntiles = m.ntiles; a.n_interior = m.n_interior; a.ext_cap = m.ext_cap;
  a.tile_order = m.tile_order; a.tile_node = m.tile_node; a.rowptr = m.rowptr; a.tile_extptr = m.tile_extptr;
  a.tile_ext = m.tile_ext; a.lc16 = m.lc16;
  a.vals = A.vals; a.dinv = A.dinv; a.rhat = rhat;
  a.x = x; a.r = r; a.p = p; a.v = v; a.t = t;
  a.part = c->partials; a.scalars = c->scalars; a.status = c->status;
  a.bar = (unsigned int*)(c->status + 5);
  a.rtol2 = rtol2; a.atol2 = atol2; a.max_it = max_it;
  CUDA_OK(cudaMemsetAsync(c->status + 3, 0, 4 * sizeof(int32_t), c->stream));   // time-out flag, fin ticket, barrier words
  persist_comm_args(c, &a.dev, &a.mailbox, &a.halo_off, &a.halo_stride, &a.peer_rank, &a.npeer, &a.error, &a.halo_seq0, &a.red_seq0);
  void* args[] = {(void*)&a};
  if (c->world > 1)
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_bicg_persist<true>, dim3(pl->grid), dim3(kTileNodes), args, pl->smem, c->stream));
  else
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_bicg_persist<false>, dim3(pl->grid), dim3(kTileNodes), args, pl->smem, c->stream));
  c->launches.total++;
  c->launches.spmv++;
}

}  // namespace cfem
