// fp64 CSR SpMV (sub-warp per row, shuffle reduction) and the Jacobi-
// preconditioned Krylov solvers built on it.  They replace the reference's
// sparse direct solves: KSP PREONLY + PC LU (Code/Linear_advection/RV_node.py:131-134,
// Code/Utils/helpers.py:35, dolfinx NewtonSolver default used by Code/KPP/KPP_exact.py:128-154).
//
// Design: all Krylov scalars stay on the device.  Dot products are reduced in
// two fixed-order stages: each CTA writes one partial, every CTA of the NEXT
// kernel re-reduces the (<= kMaxPartials) partials from L2.  No atomics, so
// results are bitwise reproducible.  A device-side `done` flag turns the
// remaining launches of a chunk into no-ops; the host polls it every few
// iterations (after a predicted iteration count) instead of every iteration.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "device_utils.cuh"
#include "launch.h"
#include "p2p.cuh"

namespace cfem {

static inline int vec_grid(const cfem_ctx* c, int64_t n) {
  int64_t b = (n + kBlock - 1) / kBlock;
  const int64_t cap = (int64_t)c->sm_count * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

#define LAUNCHED(c) do { CUDA_OK(cudaGetLastError()); (c)->launches.total++; } while (0)

// partial slots inside ctx->partials (each kMaxPartials doubles)
enum { P_PQ = 0, P_RZ0 = 1, P_RZ1 = 2, P_RR = 3, P_BB = 4, P_A = 5, P_B = 6, P_C = 7 };
// device scalars
enum { S_SUM = 0, S_MIN = 1, S_MAX = 2, S_BB = 3, S_ALPHA = 4, S_OMEGA = 5, S_RHO = 6, S_RELRES = 7, S_RR = 8,
       S_D0 = 16 /* .. S_D0+3: finalised dot products of the SpMV-type kernels */, S_RHO0 = 21, S_RHO1 = 22 };

// SpMV-type kernel family (A/B switch CFEM_SPMV):
//   t16     staged tile kernels over the 16-bit tile-local column format (k_tile_t16)
//   stream  CSR-stream tile kernels, one x gather per entry (k_spmv_stream / k_cheb_stream; round-1 default)
// (the sub-warp-per-row and TMA-staged variants of round 1 were measured slower and are gone: DESIGN.md section 4a)
static int g_spmv_t16 = -1;
static inline bool use_t16() {
  if (g_spmv_t16 < 0) {
    const char* e = getenv("CFEM_SPMV");
    g_spmv_t16 = (e && std::string(e) == "stream") ? 0 : 1;
  }
  return g_spmv_t16 == 1;
}

// ---------------------------------------------------------------- L2 residency of the solve's matrix
// A Krylov / Chebyshev solve streams the same matrix 20-60 times; at ~1 M rows its values + pattern (88 MB) fit
// the persisting part of the 126 MB L2.  The window covers [values | rowptr | colidx] (MASS_BC) or
// [rowptr | colidx | values] (SYSTEM) of the hot block; when it is larger than the set-aside, hitRatio keeps a
// fixed random subset of its lines persisting instead of letting them thrash.  Vector traffic misses as
// "streaming" lines, which are the first to be evicted.
// device-wide persisting set-aside currently in force, per device (the limit belongs to the device, not to a context)
static size_t g_l2_limit[64] = {0};
static void l2_set_limit(cfem_ctx* c, size_t bytes) {
  const int d = c->device < 64 ? c->device : 63;
  if (g_l2_limit[d] == bytes) return;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) { cudaGetLastError(); return; }
  if (bytes == 0) cudaCtxResetPersistingL2Cache();   // lines that were persisting go back to normal replacement
  g_l2_limit[d] = bytes;
}

void l2_prefer(cfem_ctx* c, const Matrix& A) {
  if (!c->l2_setaside) return;
  int which = -1;
  if (A.vals == c->mat[CFEM_MAT_MASS_BC].vals) which = CFEM_MAT_MASS_BC;
  else if (A.vals == c->mat[CFEM_MAT_SYSTEM].vals) which = CFEM_MAT_SYSTEM;
  size_t lo = 0, bytes = 0;
  if (which >= 0) {
    // the legacy CSR-stream kernels (CFEM_SPMV=stream) read colidx, which lies behind the SYSTEM values
    lo = which == CFEM_MAT_MASS_BC ? c->hot_off[0] : c->hot_off[1];
    const size_t hi = which == CFEM_MAT_MASS_BC ? c->hot_off[5] : (use_t16() ? c->hot_off[6] : c->hot_off[7]);
    bytes = hi - lo;
    // only when the whole window fits the set-aside: a partly resident matrix (large meshes) saves little DRAM
    // traffic and costs the vectors the L2 capacity they were using (4 M-cell KPP: 15.7 -> 22 ms per step)
    if (bytes > c->l2_setaside || (c->l2_max_window && bytes > c->l2_max_window)) which = -1;
  }
  l2_set_limit(c, which >= 0 ? c->l2_setaside : 0);
  if (which == c->l2_window) return;
  cudaStreamAttrValue attr{};
  if (which >= 0) {
    attr.accessPolicyWindow.base_ptr = c->hot_base + lo;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;  // detach
  }
  CUDA_OK(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  c->l2_window = which;
}

void l2_release(cfem_ctx* c) {
  if (!c->l2_setaside) return;
  if (c->l2_window >= 0) {
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    c->l2_window = -1;
  }
  l2_set_limit(c, 0);
}

// ---------------------------------------------------------------- basic vector kernels
__global__ void k_gather(const double* __restrict__ src, const int32_t* __restrict__ idx, double* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) dst[i] = src[idx[i]];
}
__global__ void k_gather2(const double2* __restrict__ src, const int32_t* __restrict__ idx, double2* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) dst[i] = src[idx[i]];
}
__global__ void k_gather4(const double4* __restrict__ src, const int32_t* __restrict__ idx, double4* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) dst[i] = src[idx[i]];
}
__global__ void k_fill(double* __restrict__ dst, double v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) dst[i] = v;
}
__global__ void k_sub(double* __restrict__ x, const double* __restrict__ dx, int64_t n) {
  pdl_wait();
  pdl_launch();
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) x[i] -= dx[i];
}
__global__ void k_norm2(const double* __restrict__ v, int64_t n, double* __restrict__ partials) {
  __shared__ double red[9];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) s += v[i] * v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void k_sum_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double red[9];
  const double s = reduce_partials(partials, n, red);
  if (threadIdx.x == 0) *out = s;
}

__global__ void k_scatter(const double* __restrict__ src, const int32_t* __restrict__ idx, double* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) dst[idx[i]] = src[i];
}
void launch_scatter(cfem_ctx* c, const double* src, const int32_t* idx, double* dst, int64_t n) {
  ProfScope ps(c, PROF_MISC);
  k_scatter<<<vec_grid(c, n), kBlock, 0, c->stream>>>(src, idx, dst, n); LAUNCHED(c);
}
void launch_gather(cfem_ctx* c, const double* src, const int32_t* idx, double* dst, int64_t n) {
  ProfScope ps(c, PROF_MISC);
  k_gather<<<vec_grid(c, n), kBlock, 0, c->stream>>>(src, idx, dst, n); LAUNCHED(c);
}
void launch_gather2(cfem_ctx* c, const double2* src, const int32_t* idx, double2* dst, int64_t n) {
  ProfScope ps(c, PROF_MISC);
  k_gather2<<<vec_grid(c, n), kBlock, 0, c->stream>>>(src, idx, dst, n); LAUNCHED(c);
}
void launch_gather4(cfem_ctx* c, const double4* src, const int32_t* idx, double4* dst, int64_t n) {
  ProfScope ps(c, PROF_MISC);
  k_gather4<<<vec_grid(c, n), kBlock, 0, c->stream>>>(src, idx, dst, n); LAUNCHED(c);
}
void launch_fill(cfem_ctx* c, double* dst, double v, int64_t n) {
  ProfScope ps(c, PROF_MISC);
  k_fill<<<vec_grid(c, n), kBlock, 0, c->stream>>>(dst, v, n); LAUNCHED(c);
}
void launch_copy(cfem_ctx* c, double* dst, const double* src, int64_t n) {
  CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
}
__global__ void k_copy_indexed(double* __restrict__ dst, const double* __restrict__ src, const int32_t* __restrict__ idx, int64_t n) {
  for (int64_t k = blockIdx.x * (int64_t)kBlock + threadIdx.x; k < n; k += (int64_t)gridDim.x * kBlock) {
    const int32_t i = idx[k];
    dst[i] = src[i];
  }
}
void launch_copy_indexed(cfem_ctx* c, double* dst, const double* src, const int32_t* idx, int64_t n) {
  if (n <= 0) return;
  ProfScope ps(c, PROF_MISC);
  k_copy_indexed<<<vec_grid(c, n), kBlock, 0, c->stream>>>(dst, src, idx, n); LAUNCHED(c);
}
// x -= dx unless *norm2 < thresh2 (a device-side "the Newton iteration was not needed": the host learns it later)
__global__ void k_sub_unless(double* __restrict__ x, const double* __restrict__ dx, int64_t n,
                             const double* __restrict__ norm2, double thresh2) {
  pdl_wait();
  pdl_launch();
  if (*norm2 < thresh2) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) x[i] -= dx[i];
}
void launch_sub_unless_below(cfem_ctx* c, double* x, const double* dx, int64_t n, const double* norm2, double thresh2) {
  ProfScope ps(c, PROF_MISC);
  launch_pdl(k_sub_unless, vec_grid(c, n), kBlock, 0, c->stream, x, dx, n, norm2, thresh2); LAUNCHED(c);
}
void launch_sub(cfem_ctx* c, double* x, const double* dx, int64_t n) {
  ProfScope ps(c, PROF_MISC);
  launch_pdl(k_sub, vec_grid(c, n), kBlock, 0, c->stream, x, dx, n); LAUNCHED(c);
}

double norm2(cfem_ctx* c, const double* v, int64_t n) {
  const int g = vec_grid(c, n);
  k_norm2<<<g, kBlock, 0, c->stream>>>(v, n, c->partials + P_C * kMaxPartials); LAUNCHED(c);
  k_sum_partials<<<1, kBlock, 0, c->stream>>>(c->partials + P_C * kMaxPartials, g, c->scalars + 15); LAUNCHED(c);
  CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + 15, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return sqrt(c->h_pinned[0]);
}

// ---------------------------------------------------------------- SpMV
// ghost entries of the input vector come from the mailbox when the exchange was push-only
// (a select on the base pointer, then ONE load: predicated twin loads cost ~30% of the kernel)
#define XG(vec, col) ((GHOST && (col) >= no) ? ghost_value(gsrc, mbox_shifted, (col), no) : (vec)[col])
// (not inlined: a handful of ghost entries per boundary tile take this path, and the polling loop inlined into the
// staging code cost the distributed kernel variants 32 bytes of spills in their hot loop)
__device__ __noinline__ double ghost_value_ll(const unsigned long long* ll2, const unsigned int tag, int* error) {
  return ll_load(ll2, tag, error);
}
__device__ __forceinline__ double ghost_value(const GhostSrc& g, const double* mbox_shifted, const int col, const int64_t no) {
  if (g.ll) return ghost_value_ll(g.ll + 2 * (size_t)(col - no), (unsigned int)g.seq, g.error);
  return mbox_shifted[col];
}
// producer half of the fused halo exchange, run by CTA 0 of a SpMV-type kernel
__device__ __forceinline__ void ghost_push(const GhostSrc& g, const double* __restrict__ v, const bool gate) {
  if (g.ll) {
    if (!gate) push_ll(g.pushdev, g.seq, [v](int node) { return v[node]; });
  } else {
    push_cta(g.pushdev, v, g.seq, gate);
  }
}

// CSR-stream SpMV: a CTA takes one assembly tile (<= kTileNodes consecutive rows,
// <= kTileNnzCap entries).  Every thread streams entries p, p+256, ... of the
// tile's contiguous CSR segment (vals/colidx fully coalesced, loads independent
// -> deep memory-level parallelism), multiplies by the gathered x[col] and parks
// the product in shared memory; then thread r sums row r's products in column
// order.  Fixed order, no atomics.
template <int NDOT, bool GHOST>
__global__ void __launch_bounds__(kTileNodes)
k_spmv_stream(const GhostSrc gsrc, const int64_t no, const int32_t* __restrict__ tile_order, const int n_interior,
              const int ntiles, const int32_t* __restrict__ tile_node, const int32_t* __restrict__ rowptr,
              const int32_t* __restrict__ colidx, const double* __restrict__ vals, const double* __restrict__ x,
              double* __restrict__ y, const double* __restrict__ d0, const double* __restrict__ d1,
              double* __restrict__ part0, double* __restrict__ part1, const int32_t* __restrict__ status,
              const double* __restrict__ scale /* nullable: y = scale .* (A x), the left Jacobi preconditioner */) {
  // Everything up to pdl_wait() reads mesh tables only, so under a programmatic launch it overlaps the
  // previous kernel's drain; x, status, the mailbox and the partials are touched after it.
  int bid = blockIdx.x, nblk = gridDim.x;
  if (GHOST && gsrc.pushdev) {  // CTA 0 is the producer half of the halo exchange (p2p.cuh)
    if (bid == 0) { pdl_wait(); pdl_launch(); ghost_push(gsrc, x, status && status[0]); return; }
    --bid; --nblk;
  }
  __shared__ double prod[kTileNnzCap];
  __shared__ int32_t rp[kTileNodes + 1];
  __shared__ double red[9];
  const int tid = threadIdx.x;
  double acc0 = 0.0, acc1 = 0.0;
  bool waited = false, synced = false;
  const double* const mbox_shifted = GHOST ? gsrc.mbox - no : nullptr;  // mbox_shifted[col] == mailbox[col - no]
  if (bid >= ntiles) { pdl_wait(); pdl_launch(); }
  for (int t = bid; t < ntiles; t += nblk) {
    const int tile = GHOST ? tile_order[t] : t;
    const int n0 = tile_node[tile], nrows = tile_node[tile + 1] - n0;
    for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = rowptr[n0 + i];
    __syncthreads();
    if (!synced) {
      pdl_wait();
      pdl_launch();
      synced = true;
      if (status && status[0]) return;
    }
    if (GHOST && t >= n_interior && !waited) { if (!gsrc.ll) ghost_wait(gsrc); waited = true; }
    const int start = rp[0], cnt = rp[nrows] - start;
    const double* __restrict__ v = vals + start;
    const int32_t* __restrict__ ci = colidx + start;
    int p = tid;
    for (; p + 3 * kTileNodes < cnt; p += 4 * kTileNodes) {
      const int c0 = ci[p], c1 = ci[p + kTileNodes], c2 = ci[p + 2 * kTileNodes], c3 = ci[p + 3 * kTileNodes];
      const double v0 = v[p], v1 = v[p + kTileNodes], v2 = v[p + 2 * kTileNodes], v3 = v[p + 3 * kTileNodes];
      const double x0 = XG(x, c0), x1 = XG(x, c1), x2 = XG(x, c2), x3 = XG(x, c3);
      prod[p] = v0 * x0;
      prod[p + kTileNodes] = v1 * x1;
      prod[p + 2 * kTileNodes] = v2 * x2;
      prod[p + 3 * kTileNodes] = v3 * x3;
    }
    for (; p < cnt; p += kTileNodes) { const int cc = ci[p]; prod[p] = v[p] * XG(x, cc); }
    __syncthreads();
    if (tid < nrows) {
      const int a = rp[tid] - start, b = rp[tid + 1] - start;
      const int row = n0 + tid;
      const double sc = scale ? scale[row] : 1.0;
      double s = 0.0;
      for (int k = a; k < b; ++k) s += prod[k];
      s *= sc;
      y[row] = s;
      if (NDOT >= 1) acc0 += s * (d0 == y ? s : d0[row]);
      if (NDOT >= 2) acc1 += s * (d1 == y ? s : d1[row]);
    }
    __syncthreads();
  }
  if (NDOT >= 1) {
    acc0 = block_sum(acc0, red);
    if (tid == 0) part0[bid] = acc0;
  }
  if (NDOT >= 2) {
    acc1 = block_sum(acc1, red);
    if (tid == 0) part1[bid] = acc1;
  }
}

// ---------------------------------------------------------------- T16 staged tile kernels (default)
// Same row arithmetic and summation order as k_spmv_stream / k_cheb_stream, but x is staged ONCE per tile:
// own rows with one coalesced load, the tile's external columns (~70-130 of them for a 256-row Hilbert tile) with
// one gather each, ghosts straight from the mailbox.  The CSR stream then is values (8 B) + 16-bit tile-local
// column (2 B) per entry and every x lookup hits shared memory.  The round-1 kernels gathered x per entry:
// ~7.3 M 32-byte sectors per launch at 1 M rows, which kept the L2->SM fabric near its ~6300 B/clk cap and is the
// reason an L2-resident matrix alone would not have helped.  Per launch at 1 M rows: L2->SM traffic ~300 MB -> ~150 MB.
#ifndef CFEM_T16_MINB
#define CFEM_T16_MINB 6   // resident CTAs per SM the register budget is held to (<= 42 registers)
#endif
#ifndef CFEM_T16_MINB_GHOST
#define CFEM_T16_MINB_GHOST CFEM_T16_MINB   // the distributed variants (ghost columns, halo words)
#endif
struct EpPre { double a, b, c; };

template <int NDOT>
struct Ep16Spmv {  // y = scale .* (A x) with NDOT fused dot products; a dot operand equal to x uses the staged own entry
  static constexpr int NACC = NDOT;
  double* y;
  const double *d0, *d1;
  double *p0, *p1;
  const double* scale;  // nullable
  const double* x;      // the input vector (to recognise d0 == x)
  double* out;          // finalised dots (Fin), nullable
  __device__ __forceinline__ EpPre pre(int row) const {
    EpPre q{0.0, 0.0, 1.0};
    if (NDOT >= 1 && d0 != y && d0 != x) q.a = d0[row];
    if (NDOT >= 2 && d1 != y && d1 != x) q.b = d1[row];
    if (scale) q.c = scale[row];
    return q;
  }
  __device__ __forceinline__ void row(int row, double s, double xown, const EpPre& q, double* acc) const {
    s *= q.c;
    y[row] = s;
    if (NDOT >= 1) acc[0] += s * (d0 == y ? s : (d0 == x ? xown : q.a));
    if (NDOT >= 2) acc[1] += s * (d1 == y ? s : (d1 == x ? xown : q.b));
  }
  __device__ __forceinline__ Slots<(NDOT > 0 ? NDOT : 1)> parts() const {
    Slots<(NDOT > 0 ? NDOT : 1)> sl;
    sl.p[0] = p0;
    if (NDOT >= 2) sl.p[NDOT >= 2 ? 1 : 0] = p1;
    return sl;
  }
  __device__ __forceinline__ void finish(const double* sums) const {
    if (out && (int)threadIdx.x < NDOT) out[threadIdx.x] = sums[threadIdx.x];
  }
};

// Second SpMV of a BiCGStab iteration: t = scale .* (A s) with the four inner products the merged update needs,
// (t,s) (t,t) (rhat,t) (rhat,s); s is the input vector, so its own entry comes from the staging.
struct Ep16BiT {
  static constexpr int NACC = 4;
  double* t;
  const double *rhat, *scale;
  double* part;   // 4 consecutive partial arrays of kMaxPartials doubles
  double* out;    // 4 finalised scalars
  __device__ __forceinline__ EpPre pre(int row) const { return EpPre{rhat[row], 0.0, scale[row]}; }
  __device__ __forceinline__ void row(int row, double s, double xown, const EpPre& q, double* acc) const {
    s *= q.c;
    t[row] = s;
    acc[0] += s * xown;
    acc[1] += s * s;
    acc[2] += q.a * s;
    acc[3] += q.a * xown;
  }
  __device__ __forceinline__ Slots<4> parts() const {
    Slots<4> sl;
#pragma unroll
    for (int k = 0; k < 4; ++k) sl.p[k] = part + (size_t)k * kMaxPartials;
    return sl;
  }
  __device__ __forceinline__ void finish(const double* sums) const {
    if (threadIdx.x < 4) out[threadIdx.x] = sums[threadIdx.x];
  }
};

template <bool FIRST>
struct Ep16Cheb {  // one Chebyshev iteration of the mass solve: r = b - M x, z = D^-1 r, d = c1 d + c2 z, x+ = x + d
  static constexpr int NACC = FIRST ? 2 : 1;
  const double *dinv, *b;
  double *xn, *d;
  double c1, c2;
  double *p0, *p1;  // ||D^-1 r||^2 , ||D^-1 b||^2 partials
  __device__ __forceinline__ EpPre pre(int row) const { return EpPre{b[row], dinv[row], FIRST ? 0.0 : d[row]}; }
  __device__ __forceinline__ void row(int row, double s, double xown, const EpPre& q, double* acc) const {
    const double r = q.a - s, z = q.b * r;
    const double dk = FIRST ? c2 * z : c1 * q.c + c2 * z;
    d[row] = dk;
    xn[row] = xown + dk;
    acc[0] += z * z;                               // ||D^-1 r||^2: the row-equilibrated residual (see chebyshev_mass)
    if (FIRST) acc[1] += (q.b * q.a) * (q.b * q.a);  // ||D^-1 b||^2
  }
  __device__ __forceinline__ Slots<NACC> parts() const {
    Slots<NACC> sl;
    sl.p[0] = p0;
    if (FIRST) sl.p[NACC - 1] = p1;
    return sl;
  }
  // finished norms (Fin): scalars[S_D0] = ||D^-1 r||^2 of this launch, scalars[S_D0 + 1] = ||D^-1 b||^2 (first launch)
  double* out;
  __device__ __forceinline__ void finish(const double* sums) const {
    if (out && threadIdx.x == 0) {
      out[0] = sums[0];
      if (FIRST) out[1] = sums[NACC - 1];
    }
  }
};

// One tile of a staged tile kernel.  GH: the tile has ghost columns (a boundary tile of a distributed context); the
// interior-tile instance carries none of the mailbox code, so the distributed kernels run their interior tiles --
// 95 % of them at 1 M rows per GPU -- through the same instruction stream as the one-GPU kernels.  Returns false when
// the launch is gated off (a predecessor raised the status flag).
template <class EP, bool GH>
__device__ __forceinline__ bool t16_tile(const TileMeta& tm, const GhostSrc& gsrc, const int64_t no,
                                         const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ lc16,
                                         const int32_t* __restrict__ ext, const double* __restrict__ vals,
                                         const double* __restrict__ x, const EP& ep, const int32_t* __restrict__ status,
                                         double* const prod, double* const xs, int32_t* const rp, double* acc,
                                         bool& synced, bool& waited) {
  const int tid = threadIdx.x;
  const int n0 = tm.n0, nrows = tm.nrows, e0 = tm.e0, ne = tm.ne, start = tm.start, cnt = tm.cnt;
  for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = rowptr[n0 + i] - start;
  const double* __restrict__ v = vals + start;
  const uint16_t* __restrict__ lc = lc16 + start;
  // mesh tables only up to here (nothing a predecessor kernel writes), so under a programmatic launch these
  // requests overlap the previous kernel's drain; the matrix values may come straight from an assembly kernel
  const int ecol = tid < ne ? ext[e0 + tid] : 0;
  int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  if (tid < cnt) l0 = lc[tid];
  if (tid + kTileNodes < cnt) l1 = lc[tid + kTileNodes];
  if (tid + 2 * kTileNodes < cnt) l2 = lc[tid + 2 * kTileNodes];
  if (tid + 3 * kTileNodes < cnt) l3 = lc[tid + 3 * kTileNodes];
  if (!synced) {
    pdl_wait();
    pdl_launch();
    synced = true;
    if (status && status[0]) return false;
  }
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  if (tid < cnt) v0 = v[tid];
  if (tid + kTileNodes < cnt) v1 = v[tid + kTileNodes];
  if (tid + 2 * kTileNodes < cnt) v2 = v[tid + 2 * kTileNodes];
  if (tid + 3 * kTileNodes < cnt) v3 = v[tid + 3 * kTileNodes];
  if (GH && !waited) { if (!gsrc.ll) ghost_wait(gsrc); waited = true; }
  // ---- stage x: own rows, then the external columns
  double xown = 0.0;
  if (tid < nrows) { xown = x[n0 + tid]; xs[tid] = xown; }
  if (!GH) {
    if (tid < ne) xs[kTileNodes + tid] = x[ecol];
    for (int e = tid + kTileNodes; e < ne; e += kTileNodes) xs[kTileNodes + e] = x[ext[e0 + e]];
  } else {
    const double* const mbox_shifted = gsrc.mbox - no;  // mbox_shifted[col] == mailbox[col - no]
    const long long t0 = clock64();
    if (tid < ne) xs[kTileNodes + tid] = ecol >= no ? ghost_value(gsrc, mbox_shifted, ecol, no) : x[ecol];
    for (int e = tid + kTileNodes; e < ne; e += kTileNodes) {
      const int cc = ext[e0 + e];
      xs[kTileNodes + e] = cc >= no ? ghost_value(gsrc, mbox_shifted, cc, no) : x[cc];
    }
    if (gsrc.tim && tid == (ne - 1) % kTileNodes) {   // one sample per boundary tile: the poll of its last (ghost) column
      const unsigned long long dt = (unsigned long long)(clock64() - t0);
      atomicAdd(gsrc.tim + 8, dt);
      atomicAdd(gsrc.tim + 9, 1ull);
      atomicMax(gsrc.tim + 10, dt);
    }
  }
  EpPre q{0.0, 0.0, 0.0};
  if (tid < nrows) q = ep.pre(n0 + tid);   // epilogue operands requested before the barrier
  __syncthreads();
  // ---- products
  {
    const int p = tid;
    if (p < cnt) prod[p] = v0 * xs[l0];
    if (p + kTileNodes < cnt) prod[p + kTileNodes] = v1 * xs[l1];
    if (p + 2 * kTileNodes < cnt) prod[p + 2 * kTileNodes] = v2 * xs[l2];
    if (p + 3 * kTileNodes < cnt) prod[p + 3 * kTileNodes] = v3 * xs[l3];
  }
  for (int p = tid + 4 * kTileNodes; p < cnt; p += kTileNodes) prod[p] = v[p] * xs[lc[p]];
  __syncthreads();
  if (tid < nrows) {
    const int a = rp[tid], b = rp[tid + 1];
    double s = 0.0;
    for (int k = a; k < b; ++k) s += prod[k];
    ep.row(n0 + tid, s, xown, q, acc);
  }
  __syncthreads();
  return true;
}

template <class EP, bool GHOST>
__global__ void __launch_bounds__(kTileNodes, GHOST ? CFEM_T16_MINB_GHOST : CFEM_T16_MINB)
k_tile_t16(const GhostSrc gsrc, const int64_t no, const int32_t* __restrict__ tile_order, const int n_interior,
           const int ntiles, const int32_t* __restrict__ tile_node, const int32_t* __restrict__ rowptr,
           const uint16_t* __restrict__ lc16, const int32_t* __restrict__ extptr, const int32_t* __restrict__ ext,
           const double* __restrict__ vals, const double* __restrict__ x, const EP ep,
           const int32_t* __restrict__ status, const Fin fin) {
  int bid = blockIdx.x, nblk = gridDim.x;
  if (GHOST && gsrc.pushdev) {  // CTA 0 is the producer half of the halo exchange (p2p.cuh)
    if (bid == 0) {
      pdl_wait();
      pdl_launch();
      // a capacity-limited grid gives one worker's slot to this CTA (launch_t16): the consumers of the per-CTA
      // partials still read `grid` of them, so the slot of the missing worker holds the neutral element
      if (EP::NACC >= 1 && nblk - 1 < ntiles) {
        const Slots<(EP::NACC > 0 ? EP::NACC : 1)> sl = ep.parts();
        if ((int)threadIdx.x < EP::NACC) sl.p[threadIdx.x][nblk - 1] = 0.0;
      }
      ghost_push(gsrc, x, status && status[0]);
      return;
    }
    --bid; --nblk;
  }
  extern __shared__ double t16_smem[];
  double* const prod = t16_smem;               // [kTileNnzCap]
  double* const xs = t16_smem + kTileNnzCap;   // [kTileNodes + ext_cap]
  __shared__ int32_t rp[kTileNodes + 1];
  __shared__ double red[9];
  const int tid = threadIdx.x;
  constexpr int NA = EP::NACC > 0 ? EP::NACC : 1;
  double acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.0;
  bool waited = false, synced = false;
  if (bid >= ntiles) { pdl_wait(); pdl_launch(); }
  // the CTA's tile schedule, resolved once (mesh tables only: overlaps the previous kernel's drain).  Rounds in
  // tile_order: the interior tiles first, a CTA's boundary tiles (if any) in its last rounds -- by then the
  // neighbours' halo values, pushed at the start of the kernel, have arrived.
  __shared__ TileMeta smeta[kMetaRounds];
  const int nrounds = (ntiles + nblk - 1) / nblk;
  fetch_tile_meta<GHOST, false>(smeta, nrounds, bid, nblk, ntiles, tile_order, tile_node, extptr, rowptr);
  __syncthreads();
  int kk = 0;
  for (; kk < nrounds; ++kk) {
    const TileMeta tm = nrounds <= kMetaRounds ? smeta[kk]
                                               : tile_meta_of<GHOST, false>(kk, nrounds, bid, nblk, ntiles, tile_order, tile_node, extptr, rowptr);
    if (tm.t < 0) continue;
    if (GHOST && tm.t >= n_interior) break;
    if (!t16_tile<EP, false>(tm, gsrc, no, rowptr, lc16, ext, vals, x, ep, status, prod, xs, rp, acc, synced, waited)) return;
  }
  if (GHOST) {
    for (; kk < nrounds; ++kk) {
      const TileMeta tm = nrounds <= kMetaRounds ? smeta[kk]
                                                 : tile_meta_of<GHOST, false>(kk, nrounds, bid, nblk, ntiles, tile_order, tile_node, extptr, rowptr);
      if (tm.t < 0) continue;
      if (!t16_tile<EP, true>(tm, gsrc, no, rowptr, lc16, ext, vals, x, ep, status, prod, xs, rp, acc, synced, waited)) return;
    }
  }
  if (EP::NACC >= 1) {
    const Slots<NA> sl = ep.parts();
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      const double a = block_sum(acc[k], red);
      if (tid == 0) sl.p[k][bid] = a;
    }
    if (fin.counter) {
      __shared__ double sums[NA];
      if (fin_reduce<NA>(fin, sl, nblk, red, sums)) ep.finish(sums);
    }
  }
}

static size_t t16_smem_bytes(const cfem_ctx* c) { return sizeof(double) * ((size_t)kTileNnzCap + kTileNodes + c->dm.ext_cap); }

template <class EP, bool GHOST>
static int t16_prepare_one(cfem_ctx* c) {
  const size_t smem = t16_smem_bytes(c);
  // The opt-in is a property of the FUNCTION, shared by every context of the process: it is raised to a fixed
  // ceiling, never set to what one mesh needs (a later context with smaller tiles would lower it under the feet
  // of an earlier one -- "invalid argument" at the next launch of the earlier context).
  if (smem > kDynSmemCeiling) CFEM_THROW(-2, "T16 tile kernel: a tile has too many external columns for shared memory");
  CUDA_OK(cudaFuncSetAttribute(k_tile_t16<EP, GHOST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemCeiling));
  int occ = 0;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tile_t16<EP, GHOST>, kTileNodes, smem));
  if (occ < 1) CFEM_THROW(-2, "T16 tile kernel does not fit on an SM (too many external columns per tile)");
  return occ;
}

// once per context: opt the instantiations in to the dynamic shared memory this mesh needs and size the grid by
// the smallest occupancy among them (kept in the context, not in a function static: two contexts may differ)
static int t16_grid(cfem_ctx* c) {
  if (c->t16_grid == 0) {
    int occ = 8;
    occ = std::min(occ, t16_prepare_one<Ep16Spmv<0>, false>(c));
    occ = std::min(occ, t16_prepare_one<Ep16Spmv<1>, false>(c));
    occ = std::min(occ, t16_prepare_one<Ep16Spmv<2>, false>(c));
    occ = std::min(occ, t16_prepare_one<Ep16BiT, false>(c));
    occ = std::min(occ, t16_prepare_one<Ep16Cheb<true>, false>(c));
    occ = std::min(occ, t16_prepare_one<Ep16Cheb<false>, false>(c));
    if (c->world > 1 || getenv("CFEM_FORCE_GHOST")) {
      occ = std::min(occ, t16_prepare_one<Ep16Spmv<0>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Spmv<1>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Spmv<2>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16BiT, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Cheb<true>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Cheb<false>, true>(c));
    }
    const int64_t cap = (int64_t)c->sm_count * occ;
    c->t16_grid = (int)(c->dm.ntiles < cap ? c->dm.ntiles : cap);
  }
  return c->t16_grid;
}

template <class EP>
static void launch_t16(cfem_ctx* c, const GhostSrc& gsrc, const Matrix& A, const double* x, const EP& ep, bool gated,
                       const Fin& fin = Fin()) {
  const int grid = t16_grid(c);
  const size_t smem = t16_smem_bytes(c);
  const int32_t* st = gated ? c->status : nullptr;
  const DevMesh& m = c->dm;
  // With the producer CTA the grid must still be co-resident: one CTA more than the SMs hold would start only when
  // the first worker retires and then run its whole schedule alone -- a tail as long as the push itself (measured:
  // +5 us per launch on two GPUs, +11 us on eight).  A capacity-limited grid therefore gives up one worker instead.
  const int total = gsrc.pushdev ? (grid < m.ntiles ? grid : grid + 1) : grid;
  if (gsrc.mbox)
    launch_pdl(k_tile_t16<EP, true>, total, kTileNodes, smem, c->stream, gsrc, m.no, m.tile_order,
               m.n_interior, m.ntiles, m.tile_node, m.rowptr, m.lc16, m.tile_extptr, m.tile_ext, A.vals, x, ep, st, fin);
  else
    launch_pdl(k_tile_t16<EP, false>, grid, kTileNodes, smem, c->stream, gsrc, m.no, m.tile_order, m.n_interior, m.ntiles,
               m.tile_node, m.rowptr, m.lc16, m.tile_extptr, m.tile_ext, A.vals, x, ep, st, fin);
}

static inline int spmv_grid(cfem_ctx* c) {
  if (use_t16()) return t16_grid(c);
  const int64_t cap = (int64_t)c->sm_count * 8;
  return (int)(c->dm.ntiles < cap ? c->dm.ntiles : cap);
}

// y = scale .* (A x) (scale nullable) with NDOT fused dot products of y with d0 / d1 (either may be y or x itself)
template <int NDOT>
static void spmv_dots(cfem_ctx* c, const Matrix& A, const double* x, double* y, const double* d0,
                      const double* d1, double* p0, double* p1, bool gated, const double* scale = nullptr) {
  const GhostSrc gsrc = halo_push(c, const_cast<double*>(x), gated, true);  // producer half; the kernel waits in its boundary CTAs
  ProfScope ps(c, PROF_SPMV);
  if (use_t16())
    launch_t16(c, gsrc, A, x, Ep16Spmv<NDOT>{y, d0, d1, p0, p1, scale, x, nullptr}, gated);
  else if (gsrc.mbox)
    launch_pdl(k_spmv_stream<NDOT, true>, spmv_grid(c) + (gsrc.pushdev ? 1 : 0), kTileNodes, 0, c->stream, gsrc, c->dm.no,
               c->dm.tile_order, c->dm.n_interior, c->dm.ntiles, c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, x, y,
               d0, d1, p0, p1, gated ? c->status : nullptr, scale);
  else
    launch_pdl(k_spmv_stream<NDOT, false>, spmv_grid(c), kTileNodes, 0, c->stream, gsrc, c->dm.no, c->dm.tile_order,
               c->dm.n_interior, c->dm.ntiles, c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, x, y, d0, d1, p0, p1,
               gated ? c->status : nullptr, scale);
  LAUNCHED(c);
  c->launches.spmv++;
  if (c->world > 1 && NDOT >= 1) {
    double* sl[2] = {p0, p1};
    const int op[2] = {0, 0};
    allreduce_partials(c, NDOT, sl, op, spmv_grid(c));
  }
}

void launch_spmv(cfem_ctx* c, const Matrix& A, const double* x, double* y) {
  spmv_dots<0>(c, A, x, y, nullptr, nullptr, nullptr, nullptr, false);
}

void launch_spmv_dots2(cfem_ctx* c, const Matrix& A, const double* x, double* y, const double* d0, double* p0, double* p1) {
  spmv_dots<2>(c, A, x, y, d0, y, p0, p1, false, A.dinv);
}

// ---------------------------------------------------------------- Chebyshev (mass matrix)
// For P1 triangles every eigenvalue of D^-1 M lies in [1/2, 2] (element-wise bound,
// inherited by the Dirichlet-reduced matrix), so the Chebyshev semi-iteration needs
// no inner products: one fused kernel per iteration (SpMV + residual + update), no
// reductions, no host round trips until the final check.
//   r_k = b - M x_k ; z_k = D^-1 r_k ; d_k = c1 d_{k-1} + c2 z_k ; x_{k+1} = x_k + d_k
// Convergence (here and in every solver below) is tested on the ROW-EQUILIBRATED residual ||D^-1 r|| / ||D^-1 b||:
// the rows of these matrices scale with the local cell area and Dirichlet rows are identity rows, so the plain
// 2-norm is dominated by whichever rows happen to be large (a moving Dirichlet jump puts O(1) entries into b next
// to O(h^2) interior ones) and 1e-13 of it says little about the small rows.  The reference solves by LU, which is
// accurate row by row; the equilibrated test is what makes the iterative answer LU-equivalent on fine meshes
// (1024^2 Burgers: field error vs the oracle 2.4e-10 with the plain norm).
template <bool FIRST, bool GHOST>
__global__ void __launch_bounds__(kTileNodes)
k_cheb_stream(const GhostSrc gsrc, const int64_t no, const int32_t* __restrict__ tile_order, const int n_interior,
              const int ntiles, const int32_t* __restrict__ tile_node, const int32_t* __restrict__ rowptr,
              const int32_t* __restrict__ colidx, const double* __restrict__ vals, const double* __restrict__ dinv,
              const double* __restrict__ b, const double* __restrict__ xk, double* __restrict__ xn,
              double* __restrict__ d, const double c1, const double c2, double* __restrict__ part_rr,
              double* __restrict__ part_bb) {
  int bid = blockIdx.x, nblk = gridDim.x;
  if (GHOST && gsrc.pushdev) {  // CTA 0 is the producer half of the halo exchange (p2p.cuh)
    if (bid == 0) { pdl_wait(); pdl_launch(); ghost_push(gsrc, xk, false); return; }
    --bid; --nblk;
  }
  __shared__ double prod[kTileNnzCap];
  __shared__ int32_t rp[kTileNodes + 1];
  __shared__ double red[9];
  const int tid = threadIdx.x;
  double rr = 0.0, bb = 0.0;
  bool waited = false, synced = false;
  const double* const mbox_shifted = GHOST ? gsrc.mbox - no : nullptr;  // mbox_shifted[col] == mailbox[col - no]
  if (bid >= ntiles) { pdl_wait(); pdl_launch(); }
  for (int t = bid; t < ntiles; t += nblk) {
    const int tile = GHOST ? tile_order[t] : t;
    const int n0 = tile_node[tile], nrows = tile_node[tile + 1] - n0;
    for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = rowptr[n0 + i];
    __syncthreads();
    if (!synced) { pdl_wait(); pdl_launch(); synced = true; }   // mesh tables only above (see k_spmv_stream)
    if (GHOST && t >= n_interior && !waited) { if (!gsrc.ll) ghost_wait(gsrc); waited = true; }
    const int start = rp[0], cnt = rp[nrows] - start;
    const double* __restrict__ v = vals + start;
    const int32_t* __restrict__ ci = colidx + start;
    int p = tid;
    for (; p + 3 * kTileNodes < cnt; p += 4 * kTileNodes) {
      const int c0 = ci[p], c1i = ci[p + kTileNodes], c2i = ci[p + 2 * kTileNodes], c3 = ci[p + 3 * kTileNodes];
      const double v0 = v[p], v1 = v[p + kTileNodes], v2 = v[p + 2 * kTileNodes], v3 = v[p + 3 * kTileNodes];
      const double x0 = XG(xk, c0), x1 = XG(xk, c1i), x2 = XG(xk, c2i), x3 = XG(xk, c3);
      prod[p] = v0 * x0;
      prod[p + kTileNodes] = v1 * x1;
      prod[p + 2 * kTileNodes] = v2 * x2;
      prod[p + 3 * kTileNodes] = v3 * x3;
    }
    for (; p < cnt; p += kTileNodes) { const int cc = ci[p]; prod[p] = v[p] * XG(xk, cc); }
    __syncthreads();
    if (tid < nrows) {
      const int row = n0 + tid;
      // operands of the update are requested before the row sum so their latency overlaps it
      const double bi = b[row], di = dinv[row], xo = xk[row];
      const double dprev = FIRST ? 0.0 : d[row];
      const int a = rp[tid] - start, e = rp[tid + 1] - start;
      double s = 0.0;
      for (int k = a; k < e; ++k) s += prod[k];
      const double r = bi - s, z = di * r;
      const double dk = FIRST ? c2 * z : c1 * dprev + c2 * z;
      d[row] = dk;
      xn[row] = xo + dk;
      rr += z * z;                             // row-equilibrated norms, see chebyshev_mass
      if (FIRST) bb += (di * bi) * (di * bi);
    }
    __syncthreads();
  }
  rr = block_sum(rr, red);
  if (tid == 0) part_rr[bid] = rr;
  if (FIRST) {
    bb = block_sum(bb, red);
    if (tid == 0) part_bb[bid] = bb;
  }
}

// writes sqrt(rr/bb) to scalars[S_RELRES] from the partials
__global__ void __launch_bounds__(kBlock)
k_relres(const double* __restrict__ part, int npart_rr, int npart_bb, double* __restrict__ scalars) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[9];
  const double rr = reduce_partials(part + P_RR * kMaxPartials, npart_rr, red);
  const double bb = reduce_partials(part + P_BB * kMaxPartials, npart_bb, red);
  if (threadIdx.x == 0) { scalars[S_RR] = rr; scalars[S_BB] = bb; scalars[S_RELRES] = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr); }
}

SolveResult chebyshev_mass(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, int max_it,
                           int* predict) {
  const int64_t n = c->dm.nn;  // copies move ghosts too
  l2_prefer(c, A);
  int np_bb = 0;
  double *xa = x, *xb = c->wk[0], *d = c->wk[1];
  double* part = c->partials;
  const int gs = spmv_grid(c);
  const double lmin = 0.5, lmax = 2.0;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma1 = theta / delta;
  double rho = 1.0 / sigma1;
  SolveResult res{0, 0.0, false};
  int it = 0;
  int target = predict ? *predict : 28;
  if (target < 2) target = 2;
  if (target > max_it) target = max_it;
  const bool persist = use_t16() && cheb_persist_available(c);
  const bool fin_norms = use_t16() && fin_available(c) && !persist;   // norms finished inside the first / last launch
  bool first_chunk = true;
  double bb2 = 0.0;
  while (true) {
    if (persist) {
      // all iterations up to the next check in ONE cooperative launch (persist.cu); the scope stands for
      // (target - it) iterations so the breakdown keeps reporting time per iteration
      const int n_it = target - it;
      {
        ProfScope chain(c, PROF_CHEB, n_it);
        launch_cheb_persist(c, A, b, xa, xb, d, it == 0, n_it, rho, sigma1, theta, delta);
      }
      for (int k = 0; k < n_it; ++k)
        if (it + k > 0) rho = 1.0 / (2.0 * sigma1 - rho);
      if (n_it & 1) std::swap(xa, xb);
      it = target;
      CUDA_OK(cudaMemcpyAsync(c->h_status, c->status, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + S_RELRES, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CUDA_OK(cudaStreamSynchronize(c->stream));
      if (c->h_status[3]) CFEM_THROW(-5, "a grid barrier of the persistent Chebyshev solver timed out");
    } else {
    // the iterations up to the next check are timed as ONE scope of (target - it) launches: the per-launch
    // figure then is the in-situ one (back-to-back launches, programmatic dependent launch active)
    {
    ProfScope chain(c, PROF_CHEB, target - it);
    for (; it < target; ++it) {
      const GhostSrc gsrc = halo_push(c, xa, false, true);
      ProfScope ps(c, PROF_CHEB);
#define CHEB_LAUNCH(FIRST, GHOST, C1, C2, PBB)                                                                   \
  launch_pdl(k_cheb_stream<FIRST, GHOST>, gs + (GHOST && gsrc.pushdev ? 1 : 0), kTileNodes, 0, c->stream, gsrc, c->dm.no, \
             c->dm.tile_order, c->dm.n_interior, c->dm.ntiles, c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, A.dinv, \
             b, xa, xb, d, C1, C2, part + P_RR * kMaxPartials, PBB)
      if (use_t16() && it == 0) {
        // the first launch finishes ||D^-1 b||^2, the last one of the chunk ||D^-1 r||^2 (in-kernel, all ranks)
        launch_t16(c, gsrc, A, xa, Ep16Cheb<true>{A.dinv, b, xb, d, 0.0, 1.0 / theta, part + P_RR * kMaxPartials, part + P_BB * kMaxPartials, c->scalars + S_D0},
                   false, fin_norms ? make_fin(c) : Fin());
      } else if (use_t16()) {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        launch_t16(c, gsrc, A, xa, Ep16Cheb<false>{A.dinv, b, xb, d, rho_new * rho, 2.0 * rho_new / delta, part + P_RR * kMaxPartials, nullptr, c->scalars + S_D0},
                   false, (fin_norms && it == target - 1) ? make_fin(c) : Fin());
        rho = rho_new;
      } else if (it == 0) {
        if (gsrc.mbox) CHEB_LAUNCH(true, true, 0.0, 1.0 / theta, part + P_BB * kMaxPartials);
        else CHEB_LAUNCH(true, false, 0.0, 1.0 / theta, part + P_BB * kMaxPartials);
      } else {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        if (gsrc.mbox) CHEB_LAUNCH(false, true, rho_new * rho, 2.0 * rho_new / delta, nullptr);
        else CHEB_LAUNCH(false, false, rho_new * rho, 2.0 * rho_new / delta, nullptr);
        rho = rho_new;
      }
#undef CHEB_LAUNCH
      LAUNCHED(c);
      c->launches.spmv++;
      if (it == 0 && !fin_norms) np_bb = allreduce_sum1(c, part + P_BB * kMaxPartials, gs);
      std::swap(xa, xb);
    }
    }
    // the last kernel measured ||b - M x_{it-1}||; x_it is one update further on
    if (fin_norms) {
      // scalars[S_D0] = ||D^-1 r||^2 (last launch), scalars[S_D0 + 1] = ||D^-1 b||^2 (first launch of the solve; kept in bb2)
      CUDA_OK(cudaMemcpyAsync(c->h_pinned + 8, c->scalars + S_D0, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CUDA_OK(cudaStreamSynchronize(c->stream));
      if (first_chunk) bb2 = c->h_pinned[9];
      first_chunk = false;
      c->h_pinned[0] = bb2 > 0.0 ? sqrt(c->h_pinned[8] / bb2) : sqrt(c->h_pinned[8]);
    } else {
    const int np_rr = allreduce_sum1(c, part + P_RR * kMaxPartials, gs);
    { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_relres, 1, kBlock, 0, c->stream, part, np_rr, np_bb, c->scalars); LAUNCHED(c); }
    CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + S_RELRES, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    }
    res.iters = it;
    res.relres = c->h_pinned[0];
    if (!(res.relres == res.relres)) break;
    if (res.relres <= rtol) { res.converged = true; break; }
    if (it >= max_it) break;
    // error contracts by ~1/3 per iteration: extend by what the bound asks for, at least 2
    int more = (int)ceil(log(res.relres / rtol) / log(3.0));
    if (more < 2) more = 2;
    target = it + more;
    if (target > max_it) target = max_it;
  }
  if (xa != x) launch_copy(c, x, xa, n);
  halo_exchange(c, x);  // the solution leaves with valid ghosts
  if (predict) {
    // next solve: drop the iterations the achieved residual shows were not needed (keep one spare)
    int spare = (res.converged && res.relres > 0.0) ? (int)floor(log(rtol / res.relres) / log(3.0)) - 1 : 0;
    if (spare < 0) spare = 0;
    *predict = res.iters - spare > 2 ? res.iters - spare : 2;
  }
  return res;
}

// ---------------------------------------------------------------- PCG
// r = b - q, z = dinv r, p = z ; partials: rz -> RZ0, rr -> RR, bb -> BB
__global__ void __launch_bounds__(kBlock)
k_pcg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
           double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ part,
           int32_t* __restrict__ status) {
  __shared__ double red[9];
  double rz = 0.0, rr = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double di = dinv[i], bi = b[i], ri = bi - q[i], zi = di * ri;
    r[i] = ri; z[i] = zi; p[i] = zi;
    rz += ri * zi; rr += zi * zi; bb += (di * bi) * (di * bi);
  }
  rz = block_sum(rz, red); rr = block_sum(rr, red); bb = block_sum(bb, red);
  if (threadIdx.x == 0) {
    part[P_RZ0 * kMaxPartials + blockIdx.x] = rz;
    part[P_RR * kMaxPartials + blockIdx.x] = rr;
    part[P_BB * kMaxPartials + blockIdx.x] = bb;
    if (blockIdx.x == 0) { status[0] = 0; status[1] = 0; }
  }
}

// decides convergence from the partials of the previous kernel (all CTAs agree)
__global__ void __launch_bounds__(kBlock)
k_check(const double* __restrict__ part, int npart, double* __restrict__ scalars, int32_t* __restrict__ status,
        double rtol2, double atol2, int first, int iters_if_stop) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[9];
  if (status[0]) return;
  const double rr = reduce_partials(part + P_RR * kMaxPartials, npart, red);
  double bb = scalars[S_BB];
  if (first) bb = reduce_partials(part + P_BB * kMaxPartials, npart, red);
  if (threadIdx.x == 0) {
    if (first) scalars[S_BB] = bb;
    scalars[S_RR] = rr;
    scalars[S_RELRES] = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    if (rr <= rtol2 * bb || rr <= atol2 || !(rr == rr)) {
      status[0] = (rr == rr) ? 1 : 2;
      if (iters_if_stop >= 0) status[1] = iters_if_stop;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
k_pcg_update(int64_t n, const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dinv,
             double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double* __restrict__ part,
             int npart_spmv, int npart_vec, int cur, const int32_t* __restrict__ status) {
  if (status[0]) return;
  __shared__ double red[9];
  const double pq = reduce_partials(part + P_PQ * kMaxPartials, npart_spmv, red);
  const double rz = reduce_partials(part + (cur ? P_RZ1 : P_RZ0) * kMaxPartials, npart_vec, red);
  const double alpha = rz / pq;
  double nrz = 0.0, rr = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i], zi = dinv[i] * ri;
    r[i] = ri; z[i] = zi;
    nrz += ri * zi; rr += zi * zi;
  }
  nrz = block_sum(nrz, red); rr = block_sum(rr, red);
  if (threadIdx.x == 0) {
    part[(cur ? P_RZ0 : P_RZ1) * kMaxPartials + blockIdx.x] = nrz;
    part[P_RR * kMaxPartials + blockIdx.x] = rr;
  }
}

// p = z + beta p ; CTA 0 also records the iteration count and tests convergence
__global__ void __launch_bounds__(kBlock)
k_pcg_p(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* __restrict__ part,
        int npart_vec, int cur, double* __restrict__ scalars, int32_t* __restrict__ status, double rtol2, double atol2) {
  if (status[0]) return;
  __shared__ double red[9];
  const double rz_old = reduce_partials(part + (cur ? P_RZ1 : P_RZ0) * kMaxPartials, npart_vec, red);
  const double rz_new = reduce_partials(part + (cur ? P_RZ0 : P_RZ1) * kMaxPartials, npart_vec, red);
  const double rr = reduce_partials(part + P_RR * kMaxPartials, npart_vec, red);
  const double bb = scalars[S_BB];
  const double beta = rz_new / rz_old;
  const bool stop = rr <= rtol2 * bb || rr <= atol2 || !(rr == rr);
  if (!stop)
    for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
      p[i] = z[i] + beta * p[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    status[1] += 1;
    scalars[S_RR] = rr;
    scalars[S_RELRES] = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    if (stop) status[0] = (rr == rr) ? 1 : 2;
  }
}

static bool poll_done(cfem_ctx* c, SolveResult& res) {
  CUDA_OK(cudaMemcpyAsync(c->h_status, c->status, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + S_RELRES, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (c->h_status[3]) CFEM_THROW(-5, "a grid barrier of the persistent solver timed out");
  res.iters = c->h_status[1];
  res.relres = c->h_pinned[0];
  res.converged = c->h_status[0] == 1;
  return c->h_status[0] != 0;
}

SolveResult pcg(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol,
                int max_it, int* predict) {
  const int64_t n = c->dm.no;
  const bool dist = c->world > 1;
  l2_prefer(c, A);
  double *r = c->wk[0], *z = c->wk[1], *p = c->wk[2], *q = c->wk[3];
  double* part = c->partials;
  const int gv = vec_grid(c, n), gs = spmv_grid(c);
  const int npv = dist ? 1 : gv, nps = dist ? 1 : gs;  // partial counts seen by consumers
  const int sum2[3] = {0, 0, 0};
  const double rtol2 = rtol * rtol, atol2 = atol * atol;
  launch_spmv(c, A, x, q);
  { ProfScope ps(c, PROF_KRYLOV_VEC); k_pcg_init<<<gv, kBlock, 0, c->stream>>>(n, b, q, A.dinv, r, z, p, part, c->status); LAUNCHED(c); }
  { double* sl[3] = {part + P_RZ0 * kMaxPartials, part + P_RR * kMaxPartials, part + P_BB * kMaxPartials}; allreduce_partials(c, 3, sl, sum2, gv); }
  { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_check, 1, kBlock, 0, c->stream, part, npv, c->scalars, c->status, rtol2, atol2, 1, 0); LAUNCHED(c); }
  SolveResult res{0, 0.0, false};
  int it = 0, next_poll = predict ? (*predict > 2 ? *predict - 1 : 1) : 4;
  while (it < max_it) {
    const int cur = it & 1;
    spmv_dots<1>(c, A, p, q, p, nullptr, part + P_PQ * kMaxPartials, nullptr, true);
    { ProfScope ps(c, PROF_KRYLOV_VEC); k_pcg_update<<<gv, kBlock, 0, c->stream>>>(n, p, q, A.dinv, x, r, z, part, nps, npv, cur, c->status); LAUNCHED(c); }
    { double* sl[2] = {part + (cur ? P_RZ0 : P_RZ1) * kMaxPartials, part + P_RR * kMaxPartials}; allreduce_partials(c, 2, sl, sum2, gv); }
    { ProfScope ps(c, PROF_KRYLOV_VEC); k_pcg_p<<<gv, kBlock, 0, c->stream>>>(n, z, p, part, npv, cur, c->scalars, c->status, rtol2, atol2); LAUNCHED(c); }
    ++it;
    if (it >= next_poll || it == max_it) {
      if (poll_done(c, res)) break;
      next_poll = it + 3;
    }
  }
  if (!res.converged) poll_done(c, res);
  halo_exchange(c, x);
  if (predict) *predict = res.iters > 0 ? res.iters : 1;
  return res;
}

// ---------------------------------------------------------------- BiCGStab (right Jacobi)
// init: r = b - q ; rhat = r ; p = r ; y = dinv p ; rho = rr
__global__ void __launch_bounds__(kBlock)
k_bi_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ rhat, double* __restrict__ p, double* __restrict__ y,
          double* __restrict__ part, int32_t* __restrict__ status) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[9];
  double rr = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double bi = b[i], ri = bi - q[i];
    r[i] = ri; rhat[i] = ri; p[i] = ri; y[i] = dinv[i] * ri;
    rr += ri * ri; bb += bi * bi;
  }
  rr = block_sum(rr, red); bb = block_sum(bb, red);
  if (threadIdx.x == 0) {
    part[P_RR * kMaxPartials + blockIdx.x] = rr;
    part[P_RZ0 * kMaxPartials + blockIdx.x] = rr;  // rho_0 = (rhat, r)
    part[P_BB * kMaxPartials + blockIdx.x] = bb;
    if (blockIdx.x == 0) { status[0] = 0; status[1] = 0; }
  }
}

// iteration k >= 1: beta = (rho_new/rho_old)(alpha/omega); p = r + beta (p - omega v); y = dinv p
__global__ void __launch_bounds__(kBlock)
k_bi_p(int64_t n, const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ dinv,
       double* __restrict__ p, double* __restrict__ y, const double* __restrict__ part, int npart, int cur,
       double* __restrict__ scalars, int32_t* __restrict__ status, double rtol2, double atol2) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  __shared__ double red[9];
  const double rho_old = reduce_partials(part + (cur ? P_RZ0 : P_RZ1) * kMaxPartials, npart, red);
  const double rho_new = reduce_partials(part + (cur ? P_RZ1 : P_RZ0) * kMaxPartials, npart, red);
  const double rr = reduce_partials(part + P_RR * kMaxPartials, npart, red);
  const double bb = scalars[S_BB], alpha = scalars[S_ALPHA], omega = scalars[S_OMEGA];
  const bool stop = rr <= rtol2 * bb || rr <= atol2 || !(rr == rr);
  if (!stop) {
    const double beta = (rho_new / rho_old) * (alpha / omega);
    for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
      const double pi = r[i] + beta * (p[i] - omega * v[i]);
      p[i] = pi; y[i] = dinv[i] * pi;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    status[1] += 1;
    scalars[S_RR] = rr;
    scalars[S_RELRES] = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    if (stop) status[0] = (rr == rr) ? 1 : 2;
  }
}

// alpha = rho / (rhat, v) ; s = r - alpha v ; z = dinv s
__global__ void __launch_bounds__(kBlock)
k_bi_s(int64_t n, const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ dinv,
       double* __restrict__ s, double* __restrict__ z, const double* __restrict__ part, int npart_vec,
       int npart_spmv, int cur, double* __restrict__ scalars, const int32_t* __restrict__ status) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  __shared__ double red[9];
  const double rho = reduce_partials(part + (cur ? P_RZ1 : P_RZ0) * kMaxPartials, npart_vec, red);
  const double rv = reduce_partials(part + P_PQ * kMaxPartials, npart_spmv, red);
  const double alpha = rv != 0.0 ? rho / rv : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double si = r[i] - alpha * v[i];
    s[i] = si; z[i] = dinv[i] * si;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) scalars[S_ALPHA] = alpha;
}

// omega = (t,s)/(t,t) ; x += alpha y + omega z ; r = s - omega t ; rho' = (rhat, r), rr
__global__ void __launch_bounds__(kBlock)
k_bi_x(int64_t n, const double* __restrict__ y, const double* __restrict__ z, const double* __restrict__ s,
       const double* __restrict__ t, const double* __restrict__ rhat, double* __restrict__ x, double* __restrict__ r,
       double* __restrict__ part, int npart_spmv, int cur, double* __restrict__ scalars,
       const int32_t* __restrict__ status) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  __shared__ double red[9];
  const double ts = reduce_partials(part + P_A * kMaxPartials, npart_spmv, red);
  const double tt = reduce_partials(part + P_B * kMaxPartials, npart_spmv, red);
  const double omega = tt > 0.0 ? ts / tt : 0.0;
  const double alpha = scalars[S_ALPHA];
  double rho = 0.0, rr = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    x[i] += alpha * y[i] + omega * z[i];
    const double ri = s[i] - omega * t[i];
    r[i] = ri;
    rho += rhat[i] * ri; rr += ri * ri;
  }
  rho = block_sum(rho, red); rr = block_sum(rr, red);
  if (threadIdx.x == 0) {
    part[(cur ? P_RZ0 : P_RZ1) * kMaxPartials + blockIdx.x] = rho;
    part[P_RR * kMaxPartials + blockIdx.x] = rr;
    if (blockIdx.x == 0) scalars[S_OMEGA] = omega;
  }
}

SolveResult bicgstab_generic(cfem_ctx* c, int64_t n, int halo_width, const double* dinv, const LinApply& apply,
                             double* const* work, const double* b, double* x, double rtol, double atol, int max_it,
                             int* predict) {
  const bool dist = c->world > 1;
  double *r = work[0], *rhat = work[1], *p = work[2], *v = work[3], *s = work[4], *t = work[5], *y = work[6],
         *z = work[7];
  double* part = c->partials;
  const int gv = vec_grid(c, n);
  const double rtol2 = rtol * rtol, atol2 = atol * atol;
  const int npv = dist ? 1 : gv;  // partial counts seen by consumers of vector-kernel partials
  const int sum3[3] = {0, 0, 0};
  apply(x, v, 0, nullptr, nullptr, nullptr, nullptr, false);
  { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_bi_init, gv, kBlock, 0, c->stream, n, b, v, dinv, r, rhat, p, y, part, c->status); LAUNCHED(c); }
  { double* sl[3] = {part + P_RR * kMaxPartials, part + P_RZ0 * kMaxPartials, part + P_BB * kMaxPartials}; allreduce_partials(c, 3, sl, sum3, gv); }
  { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_check, 1, kBlock, 0, c->stream, part, npv, c->scalars, c->status, rtol2, atol2, 1, 0); LAUNCHED(c); }
  SolveResult res{0, 0.0, false};
  int it = 0, next_poll = predict ? (*predict > 2 ? *predict - 1 : 1) : 4;
  while (it < max_it) {
    const int cur = it & 1;  // rho of this iteration lives in RZ[cur]
    if (it > 0) {
      ProfScope ps(c, PROF_KRYLOV_VEC);
      launch_pdl(k_bi_p, gv, kBlock, 0, c->stream, n, r, v, dinv, p, y, part, npv, cur, c->scalars, c->status, rtol2, atol2);
      LAUNCHED(c);
    }
    const int nps1 = apply(y, v, 1, rhat, nullptr, part + P_PQ * kMaxPartials, nullptr, true);
    { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_bi_s, gv, kBlock, 0, c->stream, n, r, v, dinv, s, z, part, npv, nps1, cur, c->scalars, c->status); LAUNCHED(c); }
    const int nps2 = apply(z, t, 2, s, t, part + P_A * kMaxPartials, part + P_B * kMaxPartials, true);
    { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_bi_x, gv, kBlock, 0, c->stream, n, y, z, s, t, rhat, x, r, part, nps2, cur, c->scalars, c->status); LAUNCHED(c); }
    { double* sl[2] = {part + (cur ? P_RZ0 : P_RZ1) * kMaxPartials, part + P_RR * kMaxPartials}; allreduce_partials(c, 2, sl, sum3, gv); }
    ++it;
    if (it >= next_poll || it == max_it) {
      // the convergence test for iteration `it` runs inside the next k_bi_p; issue a stand-alone check
      { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_check, 1, kBlock, 0, c->stream, part, npv, c->scalars, c->status, rtol2, atol2, 0, it); LAUNCHED(c); }
      if (poll_done(c, res)) break;
      next_poll = it + 2;
    }
  }
  if (!res.converged) { poll_done(c, res); }
  halo_exchange(c, x, halo_width);
  if (predict) *predict = res.iters > 0 ? res.iters : 1;
  return res;
}

// ---------------------------------------------------------------- BiCGStab, LEFT Jacobi (scalar CSR systems)
// The iteration runs on D^-1 A x = D^-1 b: the SpMV-type kernel scales its row sums by 1/diag on the way out, so no
// preconditioned copies (y, z) of the direction vectors exist, the vector kernels move 14 instead of 19 vectors per
// iteration, and the residual the recurrence carries IS the row-equilibrated one the convergence test wants.
//   v = D^-1 A p ; alpha = rho/(rhat,v) ; s = r - alpha v ; t = D^-1 A s ; omega = (t,s)/(t,t)
//   x += alpha p + omega s ; r = s - omega t ; rho' = (rhat,r) ; beta = (rho'/rho)(alpha/omega) ; p = r + beta (p - omega v)
__global__ void __launch_bounds__(kBlock)
k_bl_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ rhat, double* __restrict__ p, double* __restrict__ part,
          int32_t* __restrict__ status) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[9];
  double rr = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double di = dinv[i], bi = di * b[i], ri = bi - di * q[i];
    r[i] = ri; rhat[i] = ri; p[i] = ri;
    rr += ri * ri; bb += bi * bi;
  }
  rr = block_sum(rr, red); bb = block_sum(bb, red);
  if (threadIdx.x == 0) {
    part[P_RR * kMaxPartials + blockIdx.x] = rr;
    part[P_RZ0 * kMaxPartials + blockIdx.x] = rr;  // rho_0 = (rhat, r)
    part[P_BB * kMaxPartials + blockIdx.x] = bb;
    if (blockIdx.x == 0) { status[0] = 0; status[1] = 0; }
  }
}

__global__ void __launch_bounds__(kBlock)
k_bl_p(int64_t n, const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ p,
       const double* __restrict__ part, int npart, int cur, double* __restrict__ scalars, int32_t* __restrict__ status,
       double rtol2, double atol2) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  __shared__ double red[9];
  const double rho_old = reduce_partials(part + (cur ? P_RZ0 : P_RZ1) * kMaxPartials, npart, red);
  const double rho_new = reduce_partials(part + (cur ? P_RZ1 : P_RZ0) * kMaxPartials, npart, red);
  const double rr = reduce_partials(part + P_RR * kMaxPartials, npart, red);
  const double bb = scalars[S_BB], alpha = scalars[S_ALPHA], omega = scalars[S_OMEGA];
  const bool stop = rr <= rtol2 * bb || rr <= atol2 || !(rr == rr);
  if (!stop) {
    const double beta = (rho_new / rho_old) * (alpha / omega);
    for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
      p[i] = r[i] + beta * (p[i] - omega * v[i]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    status[1] += 1;
    scalars[S_RR] = rr;
    scalars[S_RELRES] = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    if (stop) status[0] = (rr == rr) ? 1 : 2;
  }
}

__global__ void __launch_bounds__(kBlock)
k_bl_s(int64_t n, const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ s,
       const double* __restrict__ part, int npart_vec, int npart_spmv, int cur, double* __restrict__ scalars,
       const int32_t* __restrict__ status) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  __shared__ double red[9];
  const double rho = reduce_partials(part + (cur ? P_RZ1 : P_RZ0) * kMaxPartials, npart_vec, red);
  const double rv = reduce_partials(part + P_PQ * kMaxPartials, npart_spmv, red);
  const double alpha = rv != 0.0 ? rho / rv : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    s[i] = r[i] - alpha * v[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) scalars[S_ALPHA] = alpha;
}

__global__ void __launch_bounds__(kBlock)
k_bl_x(int64_t n, const double* __restrict__ p, const double* __restrict__ s, const double* __restrict__ t,
       const double* __restrict__ rhat, double* __restrict__ x, double* __restrict__ r, double* __restrict__ part,
       int npart_spmv, int cur, double* __restrict__ scalars, const int32_t* __restrict__ status) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  __shared__ double red[9];
  const double ts = reduce_partials(part + P_A * kMaxPartials, npart_spmv, red);
  const double tt = reduce_partials(part + P_B * kMaxPartials, npart_spmv, red);
  const double omega = tt > 0.0 ? ts / tt : 0.0;
  const double alpha = scalars[S_ALPHA];
  double rho = 0.0, rr = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double si = s[i];
    x[i] += alpha * p[i] + omega * si;
    const double ri = si - omega * t[i];
    r[i] = ri;
    rho += rhat[i] * ri; rr += ri * ri;
  }
  rho = block_sum(rho, red); rr = block_sum(rr, red);
  if (threadIdx.x == 0) {
    part[(cur ? P_RZ0 : P_RZ1) * kMaxPartials + blockIdx.x] = rho;
    part[P_RR * kMaxPartials + blockIdx.x] = rr;
    if (blockIdx.x == 0) scalars[S_OMEGA] = omega;
  }
}

static SolveResult bicgstab_5k(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol,
                               int max_it, int* predict) {
  l2_prefer(c, A);
  const int64_t n = c->dm.no;
  const bool dist = c->world > 1;
  double *r = c->wk[0], *rhat = c->wk[1], *p = c->wk[2], *v = c->wk[3], *s = c->wk[4], *t = c->wk[5];
  double* part = c->partials;
  const int gv = vec_grid(c, n);
  const double rtol2 = rtol * rtol, atol2 = atol * atol;
  const int npv = dist ? 1 : gv;              // partial counts seen by the consumers of vector-kernel partials
  const int nps = dist ? 1 : spmv_grid(c);    // ... and of SpMV partials
  const int sum3[3] = {0, 0, 0};
  spmv_dots<0>(c, A, x, v, nullptr, nullptr, nullptr, nullptr, false);
  { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_bl_init, gv, kBlock, 0, c->stream, n, b, v, A.dinv, r, rhat, p, part, c->status); LAUNCHED(c); }
  { double* sl[3] = {part + P_RR * kMaxPartials, part + P_RZ0 * kMaxPartials, part + P_BB * kMaxPartials}; allreduce_partials(c, 3, sl, sum3, gv); }
  { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_check, 1, kBlock, 0, c->stream, part, npv, c->scalars, c->status, rtol2, atol2, 1, 0); LAUNCHED(c); }
  SolveResult res{0, 0.0, false};
  int it = 0, next_poll = predict ? (*predict > 2 ? *predict - 1 : 1) : 4;
  while (it < max_it) {
    const int cur = it & 1;  // rho of this iteration lives in RZ[cur]
    if (it > 0) {
      ProfScope ps(c, PROF_KRYLOV_VEC);
      launch_pdl(k_bl_p, gv, kBlock, 0, c->stream, n, r, v, p, part, npv, cur, c->scalars, c->status, rtol2, atol2);
      LAUNCHED(c);
    }
    spmv_dots<1>(c, A, p, v, rhat, nullptr, part + P_PQ * kMaxPartials, nullptr, true, A.dinv);
    { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_bl_s, gv, kBlock, 0, c->stream, n, r, v, s, part, npv, nps, cur, c->scalars, c->status); LAUNCHED(c); }
    spmv_dots<2>(c, A, s, t, s, t, part + P_A * kMaxPartials, part + P_B * kMaxPartials, true, A.dinv);
    { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_bl_x, gv, kBlock, 0, c->stream, n, p, s, t, rhat, x, r, part, nps, cur, c->scalars, c->status); LAUNCHED(c); }
    { double* sl[2] = {part + (cur ? P_RZ0 : P_RZ1) * kMaxPartials, part + P_RR * kMaxPartials}; allreduce_partials(c, 2, sl, sum3, gv); }
    ++it;
    if (it >= next_poll || it == max_it) {
      // the convergence test for iteration `it` runs inside the next k_bl_p; issue a stand-alone check
      { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_check, 1, kBlock, 0, c->stream, part, npv, c->scalars, c->status, rtol2, atol2, 0, it); LAUNCHED(c); }
      if (poll_done(c, res)) break;
      next_poll = it + 2;
    }
  }
  if (!res.converged) { poll_done(c, res); }
  // the ghost entries of x are NOT refreshed: the callers use the owned part (a Newton update, an exported result)
  // or exchange the vector they form from it
  if (predict) *predict = res.iters > 0 ? res.iters : 1;
  return res;
}

// ---------------------------------------------------------------- BiCGStab, merged form (default)
// Same recurrence as above, regrouped so an iteration is FOUR launches with TWO reduction points instead of five / three:
//   K1  v = D^-1 A p                      dot (rhat,v)                                   [SpMV-type kernel]
//   K2  alpha = rho/(rhat,v) ; s = r - alpha v
//   K3  t = D^-1 A s                      dots (t,s) (t,t) (rhat,t) (rhat,s)             [SpMV-type kernel]
//   K4  omega = (t,s)/(t,t) ; rho' = (rhat,s) - omega (rhat,t)   [= (rhat, s - omega t): no pass over r needed]
//       beta = (rho'/rho)(alpha/omega) ; x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v)
//       ||r||^2 -> convergence verdict (device flag)
// Every reduction is finalised inside the kernel that produces it (fin_reduce: last CTA, plus the exchange with the
// other ranks through the peer mailboxes in a distributed context), so there is no finalise or all-reduce launch
// and the consumers read ready scalars.
__global__ void __launch_bounds__(kBlock)
k_bm_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ rhat, double* __restrict__ p, double* __restrict__ part,
          double* __restrict__ scalars, int32_t* __restrict__ status, double rtol2, double atol2, const Fin fin) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[9];
  __shared__ double sums[2];
  double rr = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double di = dinv[i], bi = di * b[i], ri = bi - di * q[i];
    r[i] = ri; rhat[i] = ri; p[i] = ri;
    rr += ri * ri; bb += bi * bi;
  }
  rr = block_sum(rr, red); bb = block_sum(bb, red);
  if (threadIdx.x == 0) {
    part[P_RR * kMaxPartials + blockIdx.x] = rr;
    part[P_BB * kMaxPartials + blockIdx.x] = bb;
  }
  Slots<2> sl;
  sl.p[0] = part + P_RR * kMaxPartials;
  sl.p[1] = part + P_BB * kMaxPartials;
  if (fin_reduce<2>(fin, sl, gridDim.x, red, sums) && threadIdx.x == 0) {
    const double grr = sums[0], gbb = sums[1];
    scalars[S_RR] = grr;
    scalars[S_BB] = gbb;
    scalars[S_RHO0] = grr;   // rho_0 = (rhat, r)
    scalars[S_RELRES] = gbb > 0.0 ? sqrt(grr / gbb) : sqrt(grr);
    status[1] = 0;
    status[0] = !(grr == grr) ? 2 : ((grr <= rtol2 * gbb || grr <= atol2) ? 1 : 0);
  }
}

__global__ void __launch_bounds__(kBlock)
k_bm_s(int64_t n, const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ s, int cur,
       double* __restrict__ scalars, const int32_t* __restrict__ status) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  const double rho = scalars[cur ? S_RHO1 : S_RHO0], rv = scalars[S_D0];
  const double alpha = rv != 0.0 ? rho / rv : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    s[i] = r[i] - alpha * v[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) scalars[S_ALPHA] = alpha;
}

__global__ void __launch_bounds__(kBlock)
k_bm_xrp(int64_t n, const double* __restrict__ s, const double* __restrict__ t, const double* __restrict__ v,
         double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, double* __restrict__ part, int cur,
         double* __restrict__ scalars, int32_t* __restrict__ status, double rtol2, double atol2, const Fin fin) {
  pdl_wait();
  pdl_launch();
  if (status[0]) return;
  __shared__ double red[9];
  __shared__ double sums[1];
  const double ts = scalars[S_D0], tt = scalars[S_D0 + 1], rt = scalars[S_D0 + 2], rs = scalars[S_D0 + 3];
  const double alpha = scalars[S_ALPHA], rho = scalars[cur ? S_RHO1 : S_RHO0];
  const double omega = tt > 0.0 ? ts / tt : 0.0;
  const double rho_new = rs - omega * rt;
  // omega == 0 only when s vanished (the alpha half-step already solved the system): r = s = 0 below and the
  // verdict is "converged"; beta must not turn that into 0 * inf
  const double beta = (omega != 0.0 && rho != 0.0) ? (rho_new / rho) * (alpha / omega) : 0.0;
  double rr = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double si = s[i], pi = p[i];
    x[i] += alpha * pi + omega * si;
    const double ri = si - omega * t[i];
    r[i] = ri;
    p[i] = ri + beta * (pi - omega * v[i]);
    rr += ri * ri;
  }
  rr = block_sum(rr, red);
  if (threadIdx.x == 0) part[P_RR * kMaxPartials + blockIdx.x] = rr;
  Slots<1> sl;
  sl.p[0] = part + P_RR * kMaxPartials;
  if (fin_reduce<1>(fin, sl, gridDim.x, red, sums) && threadIdx.x == 0) {
    const double grr = sums[0], bb = scalars[S_BB];
    scalars[S_RR] = grr;
    scalars[S_RELRES] = bb > 0.0 ? sqrt(grr / bb) : sqrt(grr);
    scalars[S_OMEGA] = omega;
    scalars[cur ? S_RHO0 : S_RHO1] = rho_new;
    status[1] += 1;
    if (!(grr == grr)) status[0] = 2;
    else if (grr <= rtol2 * bb || grr <= atol2) status[0] = 1;
    else if (!(beta == beta)) status[0] = 2;   // breakdown: the next direction is not finite
  }
}

static SolveResult bicgstab_merged(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol,
                                   int max_it, int* predict) {
  l2_prefer(c, A);
  const int64_t n = c->dm.no;
  double *r = c->wk[0], *rhat = c->wk[1], *p = c->wk[2], *v = c->wk[3], *s = c->wk[4], *t = c->wk[5];
  double* part = c->partials;
  double* dots = c->scalars + S_D0;
  const int gv = vec_grid(c, n);
  const double rtol2 = rtol * rtol, atol2 = atol * atol;
  spmv_dots<0>(c, A, x, v, nullptr, nullptr, nullptr, nullptr, false);
  { ProfScope ps(c, PROF_KRYLOV_VEC);
    launch_pdl(k_bm_init, gv, kBlock, 0, c->stream, n, b, v, A.dinv, r, rhat, p, part, c->scalars, c->status, rtol2, atol2, make_fin(c));
    LAUNCHED(c); }
  SolveResult res{0, 0.0, false};
  int it = 0, next_poll = predict ? (*predict > 1 ? *predict : 1) : 4;
  while (it < max_it) {
    const int cur = it & 1;  // rho of this iteration lives in S_RHO[cur]
    { const GhostSrc gsrc = halo_push(c, p, true, true);
      ProfScope ps(c, PROF_SPMV);
      launch_t16(c, gsrc, A, p, Ep16Spmv<1>{v, rhat, nullptr, part + P_PQ * kMaxPartials, nullptr, A.dinv, p, dots}, true, make_fin(c));
      LAUNCHED(c); c->launches.spmv++; }
    { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_bm_s, gv, kBlock, 0, c->stream, n, r, v, s, cur, c->scalars, c->status); LAUNCHED(c); }
    { const GhostSrc gsrc = halo_push(c, s, true, true);
      ProfScope ps(c, PROF_SPMV);
      launch_t16(c, gsrc, A, s, Ep16BiT{t, rhat, A.dinv, part + P_A * kMaxPartials, dots}, true, make_fin(c));
      LAUNCHED(c); c->launches.spmv++; }
    { ProfScope ps(c, PROF_KRYLOV_VEC);
      launch_pdl(k_bm_xrp, gv, kBlock, 0, c->stream, n, s, t, v, x, r, p, part, cur, c->scalars, c->status, rtol2, atol2, make_fin(c));
      LAUNCHED(c); }
    ++it;
    if (it >= next_poll || it == max_it) {
      if (poll_done(c, res)) break;
      next_poll = it + 2;
    }
  }
  if (!res.converged) poll_done(c, res);
  // the ghost entries of x are NOT refreshed: the callers use the owned part (a Newton update, an exported result)
  // or exchange the vector they form from it
  if (predict) *predict = res.iters > 0 ? res.iters : 1;
  return res;
}

// whole iteration loop in one cooperative launch (persist.cu); the host syncs once, to learn the verdict
static SolveResult bicgstab_persist(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol,
                                    int max_it, int* predict) {
  l2_prefer(c, A);
  const int64_t n = c->dm.no;
  double *r = c->wk[0], *rhat = c->wk[1], *p = c->wk[2], *v = c->wk[3], *t = c->wk[5];
  const double rtol2 = rtol * rtol, atol2 = atol * atol;
  spmv_dots<0>(c, A, x, v, nullptr, nullptr, nullptr, nullptr, false);
  { ProfScope ps(c, PROF_KRYLOV_VEC);
    launch_pdl(k_bm_init, vec_grid(c, n), kBlock, 0, c->stream, n, b, v, A.dinv, r, rhat, p, c->partials, c->scalars, c->status, rtol2, atol2, make_fin(c));
    LAUNCHED(c); }
  SolveResult res{0, 0.0, false};
  { ProfScope ps(c, PROF_SOLVER);   // one launch = the whole BiCGStab loop (its own category of the breakdown)
    launch_bicg_persist(c, A, rhat, x, r, p, v, t, rtol2, atol2, max_it); }
  poll_done(c, res);
  persist_comm_advance(c, 2 * (int64_t)res.iters, 2 * (int64_t)res.iters);   // per iteration: p and s halos, two reductions
  // the ghost entries of x are NOT refreshed: the callers use the owned part (a Newton update, an exported result)
  // or exchange the vector they form from it
  if (predict) *predict = res.iters > 0 ? res.iters : 1;
  return res;
}

// ---- asynchronous form: launch the solve and return; the caller queues the kernels that consume x, synchronises
// the stream ONCE (for whatever it needs next) and then collects the verdict.  One host round trip per Newton
// iteration instead of two -- each one drains the stream and, distributed, exposes the host latency of the slowest
// rank to all of them.  The exchange sequence numbers the solve may use are reserved up front (even blocks, so the
// slot parity of the next exchange is the one it would have had; see the end of k_bicg_persist).
bool bicgstab_async_available(cfem_ctx* c) {
  static const bool want5 = getenv("CFEM_BICGSTAB") && std::string(getenv("CFEM_BICGSTAB")) == "5k";
  return !want5 && use_t16() && fin_available(c) && bicgstab_persist_available(c);
}

void bicgstab_persist_begin(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol, int max_it) {
  l2_prefer(c, A);
  const int64_t n = c->dm.no;
  double *r = c->wk[0], *rhat = c->wk[1], *p = c->wk[2], *v = c->wk[3], *t = c->wk[5];
  const double rtol2 = rtol * rtol, atol2 = atol * atol;
  spmv_dots<0>(c, A, x, v, nullptr, nullptr, nullptr, nullptr, false);
  { ProfScope ps(c, PROF_KRYLOV_VEC);
    launch_pdl(k_bm_init, vec_grid(c, n), kBlock, 0, c->stream, n, b, v, A.dinv, r, rhat, p, c->partials, c->scalars, c->status, rtol2, atol2, make_fin(c));
    LAUNCHED(c); }
  { ProfScope ps(c, PROF_SOLVER);
    launch_bicg_persist(c, A, rhat, x, r, p, v, t, rtol2, atol2, max_it); }
  CUDA_OK(cudaMemcpyAsync(c->h_status, c->status, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + S_RELRES, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  persist_seq_reserve(c, 2 * (int64_t)max_it, 2 * (int64_t)max_it);
}

SolveResult bicgstab_persist_end(cfem_ctx* c) {
  SolveResult res{0, 0.0, false};
  if (c->h_status[3]) CFEM_THROW(-5, "a grid barrier of the persistent solver timed out");
  res.iters = c->h_status[1];
  res.relres = c->h_pinned[0];
  res.converged = c->h_status[0] == 1;
  persist_comm_count(c, 2 * (int64_t)res.iters, 2 * (int64_t)res.iters);
  return res;
}

SolveResult bicgstab(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol,
                     int max_it, int* predict) {
  // merged form: needs the staged tile kernels and in-kernel all-reduces (one GPU, or the peer-memory path);
  // CFEM_BICGSTAB=5k, CFEM_SPMV=stream and the NCCL fallback use the five-launch form
  static const bool want5 = getenv("CFEM_BICGSTAB") && std::string(getenv("CFEM_BICGSTAB")) == "5k";
  if (want5 || !use_t16() || !fin_available(c)) return bicgstab_5k(c, A, b, x, rtol, atol, max_it, predict);
  // default: the persistent kernel; CFEM_BICGSTAB=merged keeps one launch per phase (four per iteration)
  if (bicgstab_persist_available(c)) return bicgstab_persist(c, A, b, x, rtol, atol, max_it, predict);
  return bicgstab_merged(c, A, b, x, rtol, atol, max_it, predict);
}

// ---------------------------------------------------------------- restarted GMRES(30), left Jacobi
// Arnoldi with classical Gram-Schmidt (all inner products of a step in one pass over the basis), Givens
// rotations and the small triangular solve on the device; the host only polls the done flag.
//   per step:  w = D^-1 A v_j ; h_i = (w, v_i), i <= j ; w -= sum h_i v_i ; v_{j+1} = w / ||w||
// (left preconditioning: residuals are the row-equilibrated ones, see chebyshev_mass)
constexpr int kGmresM = 30;
// layout of the small device block (doubles): H (31 x 30, column major) | cs[30] | sn[30] | g[31] | y[30] | misc
constexpr int kGmH = 0, kGmCs = 31 * 30, kGmSn = kGmCs + 30, kGmG = kGmSn + 30, kGmY = kGmG + 31, kGmMisc = kGmY + 30,
              kGmSmall = kGmMisc + 8;

__global__ void __launch_bounds__(kBlock)
k_gm_residual(int64_t n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
              double* __restrict__ w, double* __restrict__ part_rr, double* __restrict__ part_bb) {
  __shared__ double red[9];
  double rr = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double di = dinv[i], bi = di * b[i], ri = bi - di * q[i];
    w[i] = ri;
    rr += ri * ri; bb += bi * bi;
  }
  rr = block_sum(rr, red); bb = block_sum(bb, red);
  if (threadIdx.x == 0) { part_rr[blockIdx.x] = rr; part_bb[blockIdx.x] = bb; }
}

// start of a cycle: beta = ||r||, g = beta e_1, convergence test on the TRUE residual
__global__ void k_gm_start(const double* __restrict__ part_rr, const double* __restrict__ part_bb, int npart,
                           double* __restrict__ sm, double* __restrict__ scalars, int32_t* __restrict__ status,
                           double rtol2, double atol2, int first_cycle) {
  __shared__ double red[9];
  const double rr = reduce_partials(part_rr, npart, red);
  const double bbn = reduce_partials(part_bb, npart, red);
  if (threadIdx.x == 0) {
    if (first_cycle) { scalars[S_BB] = bbn; status[1] = 0; }
    const double bb = first_cycle ? bbn : scalars[S_BB];
    scalars[S_RR] = rr;
    scalars[S_RELRES] = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    sm[kGmG] = sqrt(rr);
    sm[kGmMisc] = 0.0;      // Arnoldi steps completed in this cycle
    sm[kGmMisc + 1] = sqrt(rr);  // norm used to normalise v_0
    status[0] = !(rr == rr) ? 2 : ((rr <= rtol2 * bb || rr <= atol2) ? 1 : 0);
  }
}

// v = w / nrm     (nrm read from the small block: slot kGmMisc+1)
__global__ void __launch_bounds__(kBlock)
k_gm_normalize(int64_t n, const double* __restrict__ w, const double* __restrict__ sm, double* __restrict__ v,
               const int32_t* __restrict__ status) {
  if (status[0]) return;
  const double inv = 1.0 / sm[kGmMisc + 1];
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) v[i] = w[i] * inv;
}

// partial (w, v_i) for i = 0..j: one pass over w and the basis
__global__ void __launch_bounds__(kBlock)
k_gm_dots(int64_t n, int64_t stride, int j, const double* __restrict__ w, const double* __restrict__ V,
          double* __restrict__ gp, const int32_t* __restrict__ status) {
  if (status[0]) return;
  __shared__ double red[9];
  double acc[kGmresM + 1];
#pragma unroll
  for (int i = 0; i <= kGmresM; ++i) acc[i] = 0.0;
  for (int64_t e = blockIdx.x * (int64_t)kBlock + threadIdx.x; e < n; e += (int64_t)gridDim.x * kBlock) {
    const double we = w[e];
#pragma unroll
    for (int i = 0; i <= kGmresM; ++i)
      if (i <= j) acc[i] += we * V[(size_t)i * stride + e];
  }
#pragma unroll
  for (int i = 0; i <= kGmresM; ++i)
    if (i <= j) {
      const double a = block_sum(acc[i], red);
      if (threadIdx.x == 0) gp[(size_t)i * kMaxPartials + blockIdx.x] = a;
    }
}

// slot[k][0] = sum of its partials (one CTA per slot) — single-GPU counterpart of allreduce_partials
__global__ void __launch_bounds__(kBlock)
k_gm_reduce(double* __restrict__ gp, int npart) {
  __shared__ double red[9];
  double* p = gp + (size_t)blockIdx.x * kMaxPartials;
  const double s = reduce_partials(p, npart, red);
  if (threadIdx.x == 0) p[0] = s;
}

// w -= sum_i h_i v_i (h_i = gp[i][0]) ; partial ||w||^2 -> gp[m+1]
__global__ void __launch_bounds__(kBlock)
k_gm_update(int64_t n, int64_t stride, int j, double* __restrict__ w, const double* __restrict__ V,
            double* __restrict__ gp, double* __restrict__ sm, const int32_t* __restrict__ status) {
  if (status[0]) return;
  __shared__ double red[9];
  __shared__ double h[kGmresM + 1];
  if (threadIdx.x <= j) h[threadIdx.x] = gp[(size_t)threadIdx.x * kMaxPartials];
  __syncthreads();
  double nn = 0.0;
  for (int64_t e = blockIdx.x * (int64_t)kBlock + threadIdx.x; e < n; e += (int64_t)gridDim.x * kBlock) {
    double we = w[e];
    for (int i = 0; i <= j; ++i) we -= h[i] * V[(size_t)i * stride + e];
    w[e] = we;
    nn += we * we;
  }
  nn = block_sum(nn, red);
  if (threadIdx.x == 0) gp[(size_t)(kGmresM + 1) * kMaxPartials + blockIdx.x] = nn;
  if (blockIdx.x == 0 && threadIdx.x <= j) sm[kGmH + j * (kGmresM + 1) + threadIdx.x] = h[threadIdx.x];
}

// column j of the Hessenberg matrix: previous rotations, new rotation, residual estimate |g_{j+1}|
__global__ void k_gm_givens(int j, const double* __restrict__ gp, int npart, double* __restrict__ sm,
                            double* __restrict__ scalars, int32_t* __restrict__ status, double rtol2, double atol2) {
  __shared__ double red[9];
  if (status[0]) return;
  const double nn = reduce_partials(gp + (size_t)(kGmresM + 1) * kMaxPartials, npart, red);
  if (threadIdx.x == 0) {
    double* H = sm + kGmH + j * (kGmresM + 1);
    double *cs = sm + kGmCs, *sn = sm + kGmSn, *g = sm + kGmG;
    const double hn = sqrt(nn);
    H[j + 1] = hn;
    sm[kGmMisc + 1] = hn;  // normalises v_{j+1}
    for (int i = 0; i < j; ++i) {
      const double t = cs[i] * H[i] + sn[i] * H[i + 1];
      H[i + 1] = -sn[i] * H[i] + cs[i] * H[i + 1];
      H[i] = t;
    }
    const double d = sqrt(H[j] * H[j] + hn * hn);
    cs[j] = d > 0.0 ? H[j] / d : 1.0;
    sn[j] = d > 0.0 ? hn / d : 0.0;
    H[j] = d;
    H[j + 1] = 0.0;
    g[j + 1] = -sn[j] * g[j];
    g[j] = cs[j] * g[j];
    sm[kGmMisc] = (double)(j + 1);
    status[1] += 1;
    const double res = g[j + 1] * g[j + 1], bb = scalars[S_BB];
    scalars[S_RR] = res;
    scalars[S_RELRES] = bb > 0.0 ? sqrt(res / bb) : sqrt(res);
    if (!(res == res)) status[0] = 2;
    else if (res <= 0.25 * rtol2 * bb || res <= 0.25 * atol2 || hn == 0.0) status[0] = 3;  // cycle ends: verify on the true residual
  }
}

// y = H^-1 g (k x k upper triangular, k = steps completed in this cycle)
__global__ void k_gm_solve(double* __restrict__ sm) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int k = (int)sm[kGmMisc];
  double* y = sm + kGmY;
  for (int i = k - 1; i >= 0; --i) {
    double s = sm[kGmG + i];
    for (int l = i + 1; l < k; ++l) s -= sm[kGmH + l * (kGmresM + 1) + i] * y[l];
    y[i] = s / sm[kGmH + i * (kGmresM + 1) + i];
  }
}

// x += sum_i y_i v_i
__global__ void __launch_bounds__(kBlock)
k_gm_xupdate(int64_t n, int64_t stride, const double* __restrict__ V, const double* __restrict__ sm,
             double* __restrict__ x) {
  __shared__ double y[kGmresM];
  const int k = (int)sm[kGmMisc];
  if (threadIdx.x < k) y[threadIdx.x] = sm[kGmY + threadIdx.x];
  __syncthreads();
  for (int64_t e = blockIdx.x * (int64_t)kBlock + threadIdx.x; e < n; e += (int64_t)gridDim.x * kBlock) {
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += y[i] * V[(size_t)i * stride + e];
    x[e] += s;
  }
}

SolveResult gmres(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol, int max_it,
                  int* predict) {
  const int64_t n = c->dm.no, nl = c->dm.nn;
  l2_prefer(c, A);
  if (!c->gmres_V) {
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, (size_t)(kGmresM + 1) * nl * sizeof(double)));
    c->allocs.push_back(p);
    c->bytes += (int64_t)(kGmresM + 1) * nl * sizeof(double);
    c->gmres_V = (double*)p;
    CUDA_OK(cudaMalloc(&p, ((size_t)(kGmresM + 2) * kMaxPartials + kGmSmall) * sizeof(double)));
    c->allocs.push_back(p);
    c->gmres_small = (double*)p;
  }
  double* V = c->gmres_V;
  double* gp = c->gmres_small;                                       // (m+2) partial arrays
  double* sm = c->gmres_small + (size_t)(kGmresM + 2) * kMaxPartials;  // small dense block
  double *w = c->wk[0], *q = c->wk[2];
  double* part = c->partials;
  const int gv = vec_grid(c, n);
  const double rtol2 = rtol * rtol, atol2 = atol * atol;
  const bool dist = c->world > 1;
  SolveResult res{0, 0.0, false};
  int total = 0;
  int next_poll = predict && *predict > 1 ? *predict : 8;
  for (int cycle = 0;; ++cycle) {
    // true residual r = b - A x: starts a cycle and is the convergence verdict on the previous one
    launch_spmv(c, A, x, q);
    { ProfScope ps(c, PROF_KRYLOV_VEC);
      k_gm_residual<<<gv, kBlock, 0, c->stream>>>(n, b, q, A.dinv, w, part + P_RR * kMaxPartials, part + P_BB * kMaxPartials); LAUNCHED(c); }
    int np = gv;
    if (dist) { double* sl[2] = {part + P_RR * kMaxPartials, part + P_BB * kMaxPartials}; const int op[2] = {0, 0}; np = allreduce_partials(c, 2, sl, op, gv); }
    { ProfScope ps(c, PROF_KRYLOV_VEC);
      k_gm_start<<<1, kBlock, 0, c->stream>>>(part + P_RR * kMaxPartials, part + P_BB * kMaxPartials, np, sm, c->scalars, c->status, rtol2, atol2, cycle == 0); LAUNCHED(c); }
    if (cycle > 0 || total >= max_it) {  // the first cycle defers this poll to the inner loop (x0 is rarely converged)
      if (poll_done(c, res) || total >= max_it) break;
    }
    { ProfScope ps(c, PROF_KRYLOV_VEC); k_gm_normalize<<<gv, kBlock, 0, c->stream>>>(n, w, sm, V, c->status); LAUNCHED(c); }
    for (int j = 0; j < kGmresM && total < max_it; ++j) {
      spmv_dots<0>(c, A, V + (size_t)j * nl, w, nullptr, nullptr, nullptr, nullptr, true, A.dinv);
      { ProfScope ps(c, PROF_KRYLOV_VEC); k_gm_dots<<<gv, kBlock, 0, c->stream>>>(n, nl, j, w, V, gp, c->status); LAUNCHED(c); }
      if (dist) {
        for (int i0 = 0; i0 <= j; i0 += 8) {
          double* sl[8]; const int op[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          const int cnt = std::min(8, j + 1 - i0);
          for (int k = 0; k < cnt; ++k) sl[k] = gp + (size_t)(i0 + k) * kMaxPartials;
          allreduce_partials(c, cnt, sl, op, gv);
        }
      } else {
        ProfScope ps(c, PROF_KRYLOV_VEC); k_gm_reduce<<<j + 1, kBlock, 0, c->stream>>>(gp, gv); LAUNCHED(c);
      }
      { ProfScope ps(c, PROF_KRYLOV_VEC); k_gm_update<<<gv, kBlock, 0, c->stream>>>(n, nl, j, w, V, gp, sm, c->status); LAUNCHED(c); }
      const int npn = allreduce_sum1(c, gp + (size_t)(kGmresM + 1) * kMaxPartials, gv);
      { ProfScope ps(c, PROF_KRYLOV_VEC);
        k_gm_givens<<<1, kBlock, 0, c->stream>>>(j, gp, npn, sm, c->scalars, c->status, rtol2, atol2); LAUNCHED(c);
        k_gm_normalize<<<gv, kBlock, 0, c->stream>>>(n, w, sm, V + (size_t)(j + 1) * nl, c->status); LAUNCHED(c); }
      ++total;
      if (total >= next_poll || total == max_it) {
        CUDA_OK(cudaMemcpyAsync(c->h_status, c->status, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        if (c->h_status[0] != 0) break;  // estimate met the tolerance (3), or NaN (2): close the cycle
        next_poll = total + 2;
      }
    }
    // close the cycle: x += D^-1 V y (k_gm_solve uses the number of steps the device actually completed)
    { ProfScope ps(c, PROF_KRYLOV_VEC);
      k_gm_solve<<<1, 32, 0, c->stream>>>(sm); LAUNCHED(c);
      k_gm_xupdate<<<gv, kBlock, 0, c->stream>>>(n, nl, V, sm, x); LAUNCHED(c); }
  }
  if (!res.converged) poll_done(c, res);
  halo_exchange(c, x);
  if (predict) *predict = res.iters > 0 ? res.iters : 1;
  return res;
}

}  // namespace cfem
