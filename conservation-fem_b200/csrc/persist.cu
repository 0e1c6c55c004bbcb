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
//                                                      partials (t,s) (t,t) (rhat,t) (rhat,s) (s,s)
//                                                                                  --- barrier + reduce: omega, rho', beta, ||r||^2
//   P4  x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v)   --- barrier (no reduction)
//
// ||r||^2 = (s,s) - 2 omega (t,s) + omega^2 (t,t) comes out of the second reduction, so the verdict on an iteration
// is known BEFORE its update: the last iteration updates x only, and the third barrier carries no reduction -- two
// cross-rank all-reduces per iteration in a distributed context instead of three.  (The recurrence loses digits only
// when ||r|| << ||s||, i.e. relative error ~ 1e-16 ||s||^2 / ||r||^2; r = s - omega t takes a factor 2-5 off s.)
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

// The two words of the grid barrier live 256 bytes apart (different L2 slices: the address hash starts at bit 8): the
// arrival atomics of the late CTAs -- the critical path -- do not queue behind the polls of the early ones.
#ifndef CFEM_BAR_GEN
#define CFEM_BAR_GEN 64      // index (in 32-bit words) of the generation word; the arrival counter is word 0
#endif
#ifndef CFEM_BAR_BACKOFF
#define CFEM_BAR_BACKOFF 0   // ns of __nanosleep between polls of the generation word (0: none)
#endif

#ifndef CFEM_PERSIST_MINB
#define CFEM_PERSIST_MINB 4   // 64 registers: no spills; measured 2.58 vs 2.76 (5 CTAs, 48 regs, spills) vs 2.85 ms per step (6 CTAs)
#endif

// device scalars / partial slots shared with linalg.cu
enum { PS_BB = 3, PS_RELRES = 7, PS_RR = 8, PS_D0 = 16, PS_RHO0 = 21 };
enum { PP_PQ = 0, PP_RR = 3, PP_BB = 4, PP_A = 5 };

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
  size_t halo_off, halo_stride;   // of the low-latency halo area (P2PDev::ll_off / ll_stride)
  const int32_t* peer_rank;
  int npeer;
  int* error;
  unsigned long long halo_seq0, red_seq0;
  unsigned long long* tim;  // wait accounting (P2PDev::tim), null on one GPU
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
  const long long t_in = clock64();
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
      atomicExch(a.bar + CFEM_BAR_GEN, gen);   // release
    }
  } else {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while ((int)(ld_acquire_u32(a.bar + CFEM_BAR_GEN) - gen) < 0) {
        if (CFEM_BAR_BACKOFF) __nanosleep(CFEM_BAR_BACKOFF);
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
  if (a.tim && threadIdx.x == 0 && blockIdx.x == 1) {   // first worker CTA of a distributed run
    atomicAdd(a.tim + 6, (unsigned long long)(clock64() - t_in));
    atomicAdd(a.tim + 7, 1ull);
  }
}

// One SpMV-type phase over this CTA's tiles.  MODE 0: x = p, y = v, acc[0] += rhat.y.  MODE 1: x = r - alpha v,
// y = t, acc = {(t,s), (t,t), (rhat,t), (rhat,s), (s,s)}.
__device__ __forceinline__ void grid_sync(unsigned int* bar, int32_t* status, int* error, const int nblk, unsigned int& gen);
template <int MODE, bool GHOST>
__device__ __forceinline__ void spmv_phase(const BicgArgs& a, const int wid, const int nwork, const double alpha,
                                           const unsigned long long hseq, double* prod, double* xs, int32_t* rp,
                                           const TileMeta* smeta, double* acc) {
  const int tid = threadIdx.x;
  const int64_t no = a.no;
  // low-latency halo words of this exchange (p2p.cuh): ghost g in ll[2g], ll[2g+1], polled by the reader
  const unsigned long long* const ll =
      GHOST ? (const unsigned long long*)(a.mailbox + a.halo_off + (size_t)(hseq & 1) * a.halo_stride) : nullptr;
  const unsigned int tag = (unsigned int)hseq;
  const int nrounds = (a.ntiles + nwork - 1) / nwork;
  for (int kk = 0; kk < nrounds; ++kk) {
    const TileMeta tm = nrounds <= kMetaRounds ? smeta[kk]
                                               : tile_meta_of<GHOST>(kk, nrounds, wid, nwork, a.ntiles, a.tile_order, a.tile_node, a.tile_extptr, a.rowptr);
    if (tm.t < 0) continue;
    const int n0 = tm.n0, nrows = tm.nrows, e0 = tm.e0, ne = tm.ne, start = tm.start, cnt = tm.cnt;
    for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = a.rowptr[n0 + i] - start;
    const double* __restrict__ v = a.vals + start;
    const uint16_t* __restrict__ lc = a.lc16 + start;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    if (tid < cnt) { v0 = v[tid]; l0 = lc[tid]; }
    if (tid + kTileNodes < cnt) { v1 = v[tid + kTileNodes]; l1 = lc[tid + kTileNodes]; }
    if (tid + 2 * kTileNodes < cnt) { v2 = v[tid + 2 * kTileNodes]; l2 = lc[tid + 2 * kTileNodes]; }
    if (tid + 3 * kTileNodes < cnt) { v3 = v[tid + 3 * kTileNodes]; l3 = lc[tid + 3 * kTileNodes]; }
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
      if (GHOST && cc >= no) {
        const long long t0 = clock64();
        val = ll_load(ll + 2 * (size_t)(cc - no), tag, a.error);
        if (a.tim && e == ne - 1) {   // one sample per boundary tile (its last external column is a ghost)
          const unsigned long long dt = (unsigned long long)(clock64() - t0);
          atomicAdd(a.tim + 0, dt);
          atomicAdd(a.tim + 1, 1ull);
          atomicMax(a.tim + 2, dt);
        }
      } else {
        val = MODE == 0 ? __ldcg(a.p + cc) : __ldcg(a.r + cc) - alpha * __ldcg(a.v + cc);
      }
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
        acc[4] += xown * xown;
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
  __shared__ double sums[5];
  const int tid = threadIdx.x;
  const int nblk = gridDim.x;
  const bool comm_cta = GHOST && blockIdx.x == 0;       // pushes halo values, owns no tile
  const int nwork = GHOST ? nblk - 1 : nblk;
  const int wid = GHOST ? (int)blockIdx.x - 1 : (int)blockIdx.x;
  unsigned int gen = 0;
  unsigned long long hseq = a.halo_seq0, rseq = a.red_seq0;
  if (__ldcg(a.status) != 0) return;                    // the initial residual already met the tolerance (uniform)
  // this CTA's tile schedule: resolved once per solve, used by every phase of every iteration
  __shared__ TileMeta smeta[kMetaRounds];
  const int nrounds = comm_cta ? 0 : (a.ntiles + nwork - 1) / nwork;
  fetch_tile_meta<GHOST>(smeta, nrounds, wid, nwork, a.ntiles, a.tile_order, a.tile_node, a.tile_extptr, a.rowptr);
  __syncthreads();
  double rho = __ldcg(a.scalars + PS_RHO0);
  const double bb = __ldcg(a.scalars + PS_BB);
  double* const out = a.scalars + PS_D0;
  double* const pPQ = a.part + (size_t)PP_PQ * kMaxPartials;
  double* const pA = a.part + (size_t)PP_A * kMaxPartials;

  for (int it = 0; it < a.max_it; ++it) {
    // ------------------------------------------------ P1: v = D^-1 A p
    ++hseq;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (comm_cta) {
      const double* p = a.p;
      if (a.dev) push_ll(a.dev, hseq, [p](int node) { return __ldcg(p + node); });
    } else {
      spmv_phase<0, GHOST>(a, wid, nwork, 0.0, hseq, prod, xs, rp, smeta, acc);
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
      if (a.dev) push_ll(a.dev, hseq, [r, v, alpha](int node) { return __ldcg(r + node) - alpha * __ldcg(v + node); });
    } else {
      acc[0] = acc[1] = acc[2] = acc[3] = acc[4] = 0.0;
      spmv_phase<1, GHOST>(a, wid, nwork, alpha, hseq, prod, xs, rp, smeta, acc);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double sk = block_sum(acc[k], red);
        if (tid == 0) pA[(size_t)k * kMaxPartials + wid] = sk;
      }
    }
    {
      Slots<5> sl;
#pragma unroll
      for (int k = 0; k < 5; ++k) sl.p[k] = pA + (size_t)k * kMaxPartials;
      grid_reduce<5>(a, nblk, nwork, sl, ++rseq, out, sums, gen);
    }
    const double ts = sums[0], tt = sums[1], rt = sums[2], rs = sums[3], ss = sums[4];
    const double omega = tt > 0.0 ? ts / tt : 0.0;
    const double rho_new = rs - omega * rt;   // (rhat, s - omega t)
    // omega == 0 only when s vanished (the alpha half-step solved the system): then r = s = 0 below and the verdict
    // is "converged"; beta must not turn that into 0 * inf
    const double beta = (omega != 0.0 && rho != 0.0) ? (rho_new / rho) * (alpha / omega) : 0.0;
    // ||s - omega t||^2 from the dots at hand (never negative in exact arithmetic; clamp the rounding)
    double grr = ss - 2.0 * omega * ts + omega * omega * tt;
    if (grr < 0.0) grr = 0.0;
    int verdict = 0;
    if (!(grr == grr)) verdict = 2;
    else if (grr <= a.rtol2 * bb || grr <= a.atol2) verdict = 1;
    else if (!(beta == beta)) verdict = 2;   // breakdown: the next direction is not finite
    if (it + 1 >= a.max_it && verdict == 0) verdict = -1;   // out of iterations: finish this update, report "not converged"
    // ------------------------------------------------ P4: x (always); r, p unless this was the last iteration
    if (!comm_cta) {
      for (int kk = 0; kk < nrounds; ++kk) {
        const TileMeta tm = nrounds <= kMetaRounds ? smeta[kk]
                                                   : tile_meta_of<GHOST>(kk, nrounds, wid, nwork, a.ntiles, a.tile_order, a.tile_node, a.tile_extptr, a.rowptr);
        if (tm.t < 0) continue;
        const int n0 = tm.n0, nrows = tm.nrows;
        if (tid < nrows) {
          const int row = n0 + tid;
          const double vi = __ldcg(a.v + row), pi = __ldcg(a.p + row);
          const double si = __ldcg(a.r + row) - alpha * vi;
          a.x[row] = __ldcg(a.x + row) + alpha * pi + omega * si;
          if (verdict == 0) {
            const double ri = si - omega * __ldcg(a.t + row);
            a.r[row] = ri;
            a.p[row] = ri + beta * (pi - omega * vi);
          }
        }
      }
    }
    if (blockIdx.x == 0 && tid == 0) {
      a.scalars[PS_RR] = grr;
      a.scalars[PS_RELRES] = bb > 0.0 ? sqrt(grr / bb) : sqrt(grr);
      a.status[1] = it + 1;
      a.status[0] = verdict < 0 ? 0 : verdict;
    }
    rho = rho_new;
    if (verdict != 0 || *(volatile int32_t*)(a.status + 3)) break;
    // next product reads p of other CTAs' rows: a plain grid barrier (nothing to reduce, nothing crosses the ranks)
    grid_sync(a.bar, a.status, a.error, nblk, gen);
  }
  // Two cross-rank reductions and two halo exchanges per iteration: the exchange sequence numbers a solve uses are
  // even whatever its iteration count, so the host can reserve an even block of them and queue the kernels that
  // follow the solve without first learning the count (linalg.cu: bicgstab_persist_begin) -- the word slots of the
  // protocols alternate with the parity of the sequence number.
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent Chebyshev mass solve: all iterations of  r = b - M x, z = D^-1 r, d = c1 d + c2 z, x+ = x + d  in one
// cooperative launch (x ping-pongs between two buffers; a plain grid barrier separates the iterations; the
// row-equilibrated norms ||D^-1 r||, ||D^-1 b|| are reduced once, at the end).  Replaces the LU solve of the residual
// projection (Code/KPP/KPP_exact.py:128-137, Code/Utils/helpers.py:35).  In a distributed context the iterations of
// neighbouring ranks are coupled only through the halo flags (no global synchronisation).
struct ChebArgs {
  int64_t no;
  int ntiles, n_interior, ext_cap;
  const int32_t *tile_order, *tile_node, *rowptr, *tile_extptr, *tile_ext;
  const uint16_t* lc16;
  const double *vals, *dinv, *b;
  double *x0, *x1, *d;      // iteration k reads (k & 1 ? x1 : x0) and writes the other one
  double *part, *scalars;
  int32_t* status;
  unsigned int* bar;
  int first;                // this launch starts the solve (d undefined, ||D^-1 b|| wanted)
  int iters;                // iterations in this launch
  double rho0, sigma1, theta, delta;   // recurrence state at entry
  const P2PDev* dev;
  const char* mailbox;
  size_t halo_off, halo_stride;
  const int32_t* peer_rank;
  int npeer;
  int* error;
  unsigned long long halo_seq0, red_seq0;
  unsigned long long* tim;
};

__device__ __forceinline__ void grid_sync(unsigned int* bar, int32_t* status, int* error, const int nblk, unsigned int& gen) {
  ++gen;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(bar, 1u);
    if (t == (unsigned int)nblk - 1u) {
      bar[0] = 0;
      __threadfence();
      atomicExch(bar + CFEM_BAR_GEN, gen);
    } else {
      const long long t0 = clock64();
      while ((int)(ld_acquire_u32(bar + CFEM_BAR_GEN) - gen) < 0) {
        if (CFEM_BAR_BACKOFF) __nanosleep(CFEM_BAR_BACKOFF);
        if (clock64() - t0 > 20000000000LL || *(volatile int32_t*)(status + 3)) {
          status[3] = 1;
          if (error) *error = 1;
          break;
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

template <bool GHOST>
__global__ void __launch_bounds__(kTileNodes, CFEM_PERSIST_MINB)
k_cheb_persist(const ChebArgs a) {
  extern __shared__ double ps_smem[];
  double* const prod = ps_smem;
  double* const xs = ps_smem + kTileNnzCap;
  __shared__ int32_t rp[kTileNodes + 1];
  __shared__ double red[9];
  __shared__ double sums[2];
  const int tid = threadIdx.x;
  const int nblk = gridDim.x;
  const bool comm_cta = GHOST && blockIdx.x == 0;
  const int nwork = GHOST ? nblk - 1 : nblk;
  const int wid = GHOST ? (int)blockIdx.x - 1 : (int)blockIdx.x;
  const int64_t no = a.no;
  unsigned int gen = 0;
  unsigned long long hseq = a.halo_seq0;
  double rho = a.rho0;
  double zz = 0.0, bb = 0.0;
  for (int it = 0; it < a.iters; ++it) {
    const bool first = a.first && it == 0, last = it == a.iters - 1;
    const double* __restrict__ xin = (it & 1) ? a.x1 : a.x0;
    double* __restrict__ xout = (it & 1) ? a.x0 : a.x1;
    double c1 = 0.0, c2 = 1.0 / a.theta;
    if (!first) {
      const double rho_new = 1.0 / (2.0 * a.sigma1 - rho);
      c1 = rho_new * rho;
      c2 = 2.0 * rho_new / a.delta;
      rho = rho_new;
    }
    ++hseq;
    if (comm_cta) {
      if (a.dev) push_ll(a.dev, hseq, [xin](int node) { return __ldcg(xin + node); });
    } else {
      const unsigned long long* const ll =
          GHOST ? (const unsigned long long*)(a.mailbox + a.halo_off + (size_t)(hseq & 1) * a.halo_stride) : nullptr;
      const unsigned int tag = (unsigned int)hseq;
      const int nrounds = (a.ntiles + nwork - 1) / nwork;
      for (int kk = 0; kk < nrounds; ++kk) {
        const TileMeta tm = tile_meta_of<GHOST>(kk, nrounds, wid, nwork, a.ntiles, a.tile_order, a.tile_node, a.tile_extptr, a.rowptr);
        if (tm.t < 0) continue;
        const int n0 = tm.n0, nrows = tm.nrows, e0 = tm.e0, ne = tm.ne, start = tm.start, cnt = tm.cnt;
        for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = a.rowptr[n0 + i] - start;
        const double* __restrict__ v = a.vals + start;
        const uint16_t* __restrict__ lc = a.lc16 + start;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
        if (tid < cnt) { v0 = v[tid]; l0 = lc[tid]; }
        if (tid + kTileNodes < cnt) { v1 = v[tid + kTileNodes]; l1 = lc[tid + kTileNodes]; }
        if (tid + 2 * kTileNodes < cnt) { v2 = v[tid + 2 * kTileNodes]; l2 = lc[tid + 2 * kTileNodes]; }
        if (tid + 3 * kTileNodes < cnt) { v3 = v[tid + 3 * kTileNodes]; l3 = lc[tid + 3 * kTileNodes]; }
        double xown = 0.0, bi = 0.0, di = 0.0, dprev = 0.0;
        if (tid < nrows) {
          const int row = n0 + tid;
          xown = __ldcg(xin + row);
          xs[tid] = xown;
          bi = a.b[row];
          di = a.dinv[row];
          if (!first) dprev = a.d[row];
        }
        for (int e = tid; e < ne; e += kTileNodes) {
          const int cc = a.tile_ext[e0 + e];
          xs[kTileNodes + e] = (GHOST && cc >= no) ? ll_load(ll + 2 * (size_t)(cc - no), tag, a.error) : __ldcg(xin + cc);
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
          const double r = bi - s, z = di * r;
          const double dk = first ? c2 * z : c1 * dprev + c2 * z;
          a.d[row] = dk;
          xout[row] = xown + dk;
          if (last) zz += z * z;
          if (first) bb += (di * bi) * (di * bi);
        }
        __syncthreads();
      }
    }
    if (!last) grid_sync(a.bar, a.status, a.error, nblk, gen);
  }
  // ---- the norms: ||D^-1 (b - M x_{last input})||^2 and, when this launch started the solve, ||D^-1 b||^2
  if (!comm_cta) {
    zz = block_sum(zz, red);
    bb = block_sum(bb, red);
    if (tid == 0) {
      a.part[(size_t)PP_RR * kMaxPartials + wid] = zz;
      a.part[(size_t)PP_BB * kMaxPartials + wid] = bb;
    }
  }
  BicgArgs ba{};   // grid_reduce reads only these fields
  ba.bar = a.bar; ba.status = a.status; ba.error = a.error; ba.dev = a.dev; ba.tim = a.tim;
  Slots<2> sl;
  sl.p[0] = a.part + (size_t)PP_RR * kMaxPartials;
  sl.p[1] = a.part + (size_t)PP_BB * kMaxPartials;
  grid_reduce<2>(ba, nblk, nwork, sl, a.red_seq0 + 1, a.scalars + PS_D0, sums, gen);
  if (blockIdx.x == 0 && tid == 0) {
    const double gbb = a.first ? sums[1] : a.scalars[PS_BB];
    if (a.first) a.scalars[PS_BB] = gbb;
    a.scalars[PS_RR] = sums[0];
    a.scalars[PS_RELRES] = gbb > 0.0 ? sqrt(sums[0] / gbb) : sqrt(sums[0]);
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
constexpr size_t kBarBytes = 512;
struct PersistPlan { int grid = 0, grid_cheb = 0; size_t smem = 0; bool ok = false, tried = false, ghost = false; unsigned int* bar = nullptr; };

template <class K>
static int plan_kernel(cfem_ctx* c, K kern, size_t smem, bool ghost) {
  // fixed ceiling, not this mesh's need: the attribute belongs to the function and is shared by all contexts
  if (smem > kDynSmemCeiling) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemCeiling) != cudaSuccess) { cudaGetLastError(); return 0; }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTileNodes, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); return 0; }
  int64_t grid = (int64_t)occ * c->sm_count;
  const int64_t want = c->dm.ntiles + (ghost ? 1 : 0);
  if (grid > want) grid = want;
  if (grid > kMaxPartials) grid = kMaxPartials;
  if (grid < (ghost ? 2 : 1)) return 0;
  return (int)grid;
}

static bool persist_plan(cfem_ctx* c) {
  PersistPlan* pl = (PersistPlan*)c->persist_plan;
  if (!pl) { pl = new PersistPlan(); c->persist_plan = pl; }
  if (!pl->tried) {
    pl->tried = true;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
    // CFEM_FORCE_GHOST=1 runs the distributed kernel variants on one GPU (no peers): isolates their own cost
    pl->ghost = c->world > 1 || getenv("CFEM_FORCE_GHOST") != nullptr;
    pl->smem = sizeof(double) * ((size_t)kTileNnzCap + kTileNodes + c->dm.ext_cap);
    if (coop) {
      pl->grid = pl->ghost ? plan_kernel(c, k_bicg_persist<true>, pl->smem, true) : plan_kernel(c, k_bicg_persist<false>, pl->smem, false);
      pl->grid_cheb = pl->ghost ? plan_kernel(c, k_cheb_persist<true>, pl->smem, true) : plan_kernel(c, k_cheb_persist<false>, pl->smem, false);
    }
    pl->ok = pl->grid > 0 && pl->grid_cheb > 0 && (c->world == 1 || c->p2p != nullptr);
    if (pl->ok && cudaMalloc((void**)&pl->bar, kBarBytes) != cudaSuccess) { cudaGetLastError(); pl->bar = nullptr; pl->ok = false; }
  }
  return pl->ok;
}

bool bicgstab_persist_available(cfem_ctx* c) {
  static const bool off = getenv("CFEM_BICGSTAB") && std::string(getenv("CFEM_BICGSTAB")) != "persist";
  return !off && persist_plan(c);
}

// opt-in (CFEM_CHEB=persist): measured SLOWER than the chain of T16 launches (33 vs 26 us per iteration at 1 M rows on
// one GPU, 0.97 vs 0.79 ms per step on two) -- the iteration has no reduction to fold into the barrier, the chain
// already overlaps its launches programmatically, and the cooperative kernel runs at 4 instead of 6 CTAs per SM
bool cheb_persist_available(cfem_ctx* c) {
  static const bool on = getenv("CFEM_CHEB") && std::string(getenv("CFEM_CHEB")) == "persist";
  return on && persist_plan(c);
}

void persist_plan_free(cfem_ctx* c) {
  if (c->persist_plan && ((PersistPlan*)c->persist_plan)->bar) cudaFree(((PersistPlan*)c->persist_plan)->bar);
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
  a.bar = pl->bar;
  a.rtol2 = rtol2; a.atol2 = atol2; a.max_it = max_it;
  CUDA_OK(cudaMemsetAsync(c->status + 3, 0, 2 * sizeof(int32_t), c->stream));   // time-out flag, fin ticket
  CUDA_OK(cudaMemsetAsync(pl->bar, 0, kBarBytes, c->stream));                    // barrier words
  persist_comm_args(c, &a.dev, &a.mailbox, &a.halo_off, &a.halo_stride, &a.peer_rank, &a.npeer, &a.error, &a.halo_seq0, &a.red_seq0, &a.tim);
  void* args[] = {(void*)&a};
  if (pl->ghost)
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_bicg_persist<true>, dim3(pl->grid), dim3(kTileNodes), args, pl->smem, c->stream));
  else
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_bicg_persist<false>, dim3(pl->grid), dim3(kTileNodes), args, pl->smem, c->stream));
  c->launches.total++;
  c->launches.spmv++;
}

// `iters` Chebyshev iterations starting from x_in (result in x_in when iters is even, else in x_other); the
// row-equilibrated relative residual of the LAST iteration's input lands in scalars[S_RELRES].
void launch_cheb_persist(cfem_ctx* c, const Matrix& A, const double* b, double* x_in, double* x_other, double* d,
                         bool first, int iters, double rho0, double sigma1, double theta, double delta) {
  PersistPlan* pl = (PersistPlan*)c->persist_plan;
  ChebArgs a{};
  const DevMesh& m = c->dm;
  a.no = m.no; a.ntiles = m.ntiles; a.n_interior = m.n_interior; a.ext_cap = m.ext_cap;
  a.tile_order = m.tile_order; a.tile_node = m.tile_node; a.rowptr = m.rowptr; a.tile_extptr = m.tile_extptr;
  a.tile_ext = m.tile_ext; a.lc16 = m.lc16;
  a.vals = A.vals; a.dinv = A.dinv; a.b = b;
  a.x0 = x_in; a.x1 = x_other; a.d = d;
  a.part = c->partials; a.scalars = c->scalars; a.status = c->status;
  a.bar = pl->bar;
  a.first = first ? 1 : 0; a.iters = iters;
  a.rho0 = rho0; a.sigma1 = sigma1; a.theta = theta; a.delta = delta;
  CUDA_OK(cudaMemsetAsync(c->status + 3, 0, 2 * sizeof(int32_t), c->stream));
  CUDA_OK(cudaMemsetAsync(pl->bar, 0, kBarBytes, c->stream));
  persist_comm_args(c, &a.dev, &a.mailbox, &a.halo_off, &a.halo_stride, &a.peer_rank, &a.npeer, &a.error, &a.halo_seq0, &a.red_seq0, &a.tim);
  void* args[] = {(void*)&a};
  if (pl->ghost)
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_cheb_persist<true>, dim3(pl->grid_cheb), dim3(kTileNodes), args, pl->smem, c->stream));
  else
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_cheb_persist<false>, dim3(pl->grid_cheb), dim3(kTileNodes), args, pl->smem, c->stream));
  persist_comm_advance(c, iters, 1);
  c->launches.total++;
  c->launches.spmv++;
}

}  // namespace cfem
