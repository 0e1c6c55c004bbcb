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
enum { S_SUM = 0, S_MIN = 1, S_MAX = 2, S_BB = 3, S_ALPHA = 4, S_OMEGA = 5, S_RHO = 6, S_RELRES = 7, S_RR = 8 };

// SpMV-type kernel family (A/B switch CFEM_SPMV, default t16):
//   t16     staged tile kernels over the 16-bit tile-local column format (k_tile_t16)
//   stream  CSR-stream tile kernels, one x gather per entry (k_spmv_stream / k_cheb_stream; round-1 default)
//   tma     TMA-staged variant of the stream kernels (measured slower)
//   subwarp sub-warp per row
static int g_spmv_mode = -1;  // 0 = tile kernels, 1 = sub-warp per row
static int g_spmv_tma = 0;
static int g_spmv_t16 = 1;
static inline int spmv_mode() {
  if (g_spmv_mode < 0) {
    const char* e = getenv("CFEM_SPMV");
    const std::string m = e ? e : "";
    g_spmv_mode = m == "subwarp" ? 1 : 0;
    g_spmv_tma = m == "tma" ? 1 : 0;
    g_spmv_t16 = (m == "stream" || m == "tma" || m == "subwarp") ? 0 : 1;
  }
  return g_spmv_mode;
}

// ---------------------------------------------------------------- L2 residency of the solve's matrix
// A Krylov / Chebyshev solve streams the same matrix 20-60 times; at ~1 M rows its values + pattern (88 MB) fit
// the persisting part of the 126 MB L2.  The window covers [values | rowptr | colidx] (MASS_BC) or
// [rowptr | colidx | values] (SYSTEM) of the hot block; when it is larger than the set-aside, hitRatio keeps a
// fixed random subset of its lines persisting instead of letting them thrash.  Vector traffic misses as
// "streaming" lines, which are the first to be evicted.
void l2_prefer(cfem_ctx* c, const Matrix& A) {
  if (!c->l2_setaside) return;
  int which = -1;
  if (A.vals == c->mat[CFEM_MAT_MASS_BC].vals) which = CFEM_MAT_MASS_BC;
  else if (A.vals == c->mat[CFEM_MAT_SYSTEM].vals) which = CFEM_MAT_SYSTEM;
  if (which == c->l2_window) return;
  cudaStreamAttrValue attr{};
  if (which >= 0) {
    // the legacy CSR-stream kernels (CFEM_SPMV=stream) read colidx, which lies behind the SYSTEM values
    const size_t lo = which == CFEM_MAT_MASS_BC ? c->hot_off[0] : c->hot_off[1];
    const size_t hi = which == CFEM_MAT_MASS_BC ? c->hot_off[5] : ((spmv_mode(), g_spmv_t16) ? c->hot_off[6] : c->hot_off[7]);
    size_t bytes = hi - lo;
    if (c->l2_max_window && bytes > c->l2_max_window) bytes = c->l2_max_window;
    attr.accessPolicyWindow.base_ptr = c->hot_base + lo;
    attr.accessPolicyWindow.num_bytes = bytes;
    const double ratio = (double)c->l2_setaside / (double)bytes;
    attr.accessPolicyWindow.hitRatio = ratio < 1.0 ? (float)ratio : 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;  // detach
  }
  CUDA_OK(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  c->l2_window = which;
}

// ---------------------------------------------------------------- basic vector kernels
__global__ void k_gather(const double* __restrict__ src, const int32_t* __restrict__ idx, double* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) dst[i] = src[idx[i]];
}
__global__ void k_gather2(const double2* __restrict__ src, const int32_t* __restrict__ idx, double2* __restrict__ dst, int64_t n) {
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
void launch_fill(cfem_ctx* c, double* dst, double v, int64_t n) {
  ProfScope ps(c, PROF_MISC);
  k_fill<<<vec_grid(c, n), kBlock, 0, c->stream>>>(dst, v, n); LAUNCHED(c);
}
void launch_copy(cfem_ctx* c, double* dst, const double* src, int64_t n) {
  CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
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
// LANES lanes cooperate on one row (P1 rows hold ~7 entries, so a warp covers
// 32/LANES consecutive rows whose CSR entries are contiguous -> coalesced).
// NDOT fused dot products of y with up to two vectors (d0, d1 == y allowed).
template <int LANES, int NDOT>
__global__ void __launch_bounds__(kBlock)
k_spmv(const int64_t nn, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
       const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
       const double* __restrict__ d0, const double* __restrict__ d1, double* __restrict__ part0,
       double* __restrict__ part1, const int32_t* __restrict__ status) {
  if (status && status[0]) return;
  constexpr int RPW = 32 / LANES;  // rows per warp
  __shared__ double red[9];
  const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
  const int64_t warp = (blockIdx.x * (int64_t)kBlock + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kBlock) >> 5;
  double acc0 = 0.0, acc1 = 0.0;
  for (int64_t base = warp * RPW; base < nn; base += nwarps * RPW) {
    const int64_t row = base + sub;
    double s = 0.0;
    if (row < nn) {
      const int p1 = rowptr[row + 1];
      for (int p = rowptr[row] + sl; p < p1; p += LANES) s += vals[p] * x[colidx[p]];
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (sl == 0 && row < nn) {
      y[row] = s;
      if (NDOT >= 1) acc0 += s * (d0 == y ? s : d0[row]);
      if (NDOT >= 2) acc1 += s * (d1 == y ? s : d1[row]);
    }
  }
  if (NDOT >= 1) {
    acc0 = block_sum(acc0, red);
    if (threadIdx.x == 0) part0[blockIdx.x] = acc0;
  }
  if (NDOT >= 2) {
    acc1 = block_sum(acc1, red);
    if (threadIdx.x == 0) part1[blockIdx.x] = acc1;
  }
}

// ghost entries of the input vector come from the mailbox when the exchange was push-only
// (a select on the base pointer, then ONE load: predicated twin loads cost ~30% of the kernel)
#define XG(vec, col) ((GHOST ? (((col) >= no) ? mbox_shifted : (vec)) : (vec))[col])

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
              double* __restrict__ part0, double* __restrict__ part1, const int32_t* __restrict__ status) {
  // Everything up to pdl_wait() reads mesh tables only, so under a programmatic launch it overlaps the
  // previous kernel's drain; x, status, the mailbox and the partials are touched after it.
  int bid = blockIdx.x, nblk = gridDim.x;
  if (GHOST && gsrc.pushdev) {  // CTA 0 is the producer half of the halo exchange (p2p.cuh)
    if (bid == 0) { pdl_wait(); pdl_launch(); push_cta(gsrc.pushdev, x, gsrc.seq, status && status[0]); return; }
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
    if (GHOST && t >= n_interior && !waited) { ghost_wait(gsrc); waited = true; }
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
      double s = 0.0;
      for (int k = a; k < b; ++k) s += prod[k];
      const int row = n0 + tid;
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

// ---- row epilogues shared by the tile kernels
template <int NDOT>
struct EpSpmv {  // y = A x with NDOT fused dot products
  static constexpr int NACC = NDOT;
  double* y;
  const double *d0, *d1;
  double *p0, *p1;
  __device__ __forceinline__ void row(int row, double s, double& a0, double& a1) const {
    y[row] = s;
    if (NDOT >= 1) a0 += s * (d0 == y ? s : d0[row]);
    if (NDOT >= 2) a1 += s * (d1 == y ? s : d1[row]);
  }
};

template <bool FIRST>
struct EpCheb {  // one Chebyshev iteration of the mass solve (see k_cheb_stream)
  static constexpr int NACC = FIRST ? 2 : 1;
  const double *dinv, *b, *xk;
  double *xn, *d;
  double c1, c2;
  double *p0, *p1;  // ||r||^2 , ||b||^2 partials
  __device__ __forceinline__ void row(int row, double s, double& a0, double& a1) const {
    const double bi = b[row], r = bi - s, z = dinv[row] * r;
    const double dk = FIRST ? c2 * z : c1 * d[row] + c2 * z;
    d[row] = dk;
    xn[row] = xk[row] + dk;
    a0 += r * r;
    if (FIRST) a1 += bi * bi;
  }
};

// ---------------------------------------------------------------- TMA-staged tile SpMV
// Same arithmetic as k_spmv_stream / k_cheb_stream, with the DRAM stream moved off the warps:
// one elected thread issues three 1-D bulk copies per tile (cp.async.bulk -> UBLKCP: vals, colidx,
// rowptr slice; 16-byte aligned windows around the tile's CSR segment) into one of two
// shared-memory buffers and an mbarrier counts the bytes in.  While the CTA gathers x, multiplies
// and row-sums tile t out of buffer (t&1), the copy engine is already filling the other buffer
// with tile t+1, and the metadata of tile t+2 is in flight in registers.
constexpr int kTmaVals = (kTileNnzCap + 2) * 8;                 // window may start one entry early
constexpr int kTmaCols = ((kTileNnzCap + 4) * 4 + 15) / 16 * 16;
constexpr int kTmaRows = ((kTileNodes + 1 + 4) * 4 + 15) / 16 * 16;
constexpr int kTmaBuf = kTmaVals + kTmaCols + kTmaRows;
static_assert(kTmaVals % 16 == 0 && kTmaBuf % 16 == 0, "bulk copies need 16-byte aligned destinations");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// bounded wait: returns false (and the kernel flags an error) instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (int spin = 0; spin < (1 << 26); ++spin) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

template <class EP, bool GHOST>
__global__ void __launch_bounds__(kTileNodes)
k_tile_spmv_tma(const GhostSrc gsrc, const int64_t no, const int32_t* __restrict__ tile_order, const int n_interior,
                const int ntiles, const int32_t* __restrict__ tile_node, const int32_t* __restrict__ rowptr,
                const int32_t* __restrict__ colidx, const double* __restrict__ vals, const double* __restrict__ x,
                const EP ep, const int32_t* __restrict__ status, int* __restrict__ error) {
  if (status && status[0]) return;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ double red[9];
  const int tid = threadIdx.x;
  const double* const mbox_shifted = GHOST ? gsrc.mbox - no : nullptr;
  double acc0 = 0.0, acc1 = 0.0;
  bool waited = false;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  struct Meta { int n0, nrows, start, cnt; };
  auto load_meta = [&](int t) {
    Meta m{0, 0, 0, 0};
    if (t < ntiles) {
      const int tile = GHOST ? tile_order[t] : t;
      m.n0 = tile_node[tile];
      const int n1 = tile_node[tile + 1];
      m.nrows = n1 - m.n0;
      m.start = rowptr[m.n0];
      m.cnt = rowptr[n1] - m.start;
    }
    return m;
  };
  auto issue = [&](int b, const Meta& m) {  // thread 0 only
    unsigned char* base = tma_smem + b * kTmaBuf;
    const uint32_t bv = (uint32_t)(((m.cnt + (m.start & 1)) * 8 + 15) & ~15);
    const uint32_t bc = (uint32_t)(((m.cnt + (m.start & 3)) * 4 + 15) & ~15);
    const uint32_t br = (uint32_t)(((m.nrows + 1 + (m.n0 & 3)) * 4 + 15) & ~15);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to this buffer are done
    mbar_expect_tx(&bars[b], bv + bc + br);
    bulk_g2s(base, vals + (m.start & ~1), bv, &bars[b]);
    bulk_g2s(base + kTmaVals, colidx + (m.start & ~3), bc, &bars[b]);
    bulk_g2s(base + kTmaVals + kTmaCols, rowptr + (m.n0 & ~3), br, &bars[b]);
  };

  int t = blockIdx.x;
  Meta cur = load_meta(t);
  if (t < ntiles && tid == 0) issue(0, cur);
  Meta nxt = load_meta(t + gridDim.x);
  int buf = 0;
  uint32_t phase[2] = {0, 0};
  for (; t < ntiles; t += gridDim.x) {
    const bool has_next = t + (int)gridDim.x < ntiles;
    if (has_next && tid == 0) issue(buf ^ 1, nxt);
    const Meta nxt2 = load_meta(t + 2 * (int)gridDim.x);  // in flight during this tile's arithmetic
    if (!mbar_wait(&bars[buf], phase[buf])) { if (tid == 0 && error) *error = 2; return; }
    phase[buf] ^= 1;
    if (GHOST && t >= n_interior && !waited) { ghost_wait(gsrc); waited = true; }
    unsigned char* base = tma_smem + buf * kTmaBuf;
    double* sv = (double*)base + (cur.start & 1);
    const int32_t* sc = (const int32_t*)(base + kTmaVals) + (cur.start & 3);
    const int32_t* sr = (const int32_t*)(base + kTmaVals + kTmaCols) + (cur.n0 & 3);
    const int cnt = cur.cnt;
    int p = tid;
    for (; p + 3 * kTileNodes < cnt; p += 4 * kTileNodes) {
      const int c0 = sc[p], c1 = sc[p + kTileNodes], c2 = sc[p + 2 * kTileNodes], c3 = sc[p + 3 * kTileNodes];
      const double x0 = XG(x, c0), x1 = XG(x, c1), x2 = XG(x, c2), x3 = XG(x, c3);
      sv[p] *= x0;
      sv[p + kTileNodes] *= x1;
      sv[p + 2 * kTileNodes] *= x2;
      sv[p + 3 * kTileNodes] *= x3;
    }
    for (; p < cnt; p += kTileNodes) { const int cc = sc[p]; sv[p] *= XG(x, cc); }
    __syncthreads();
    if (tid < cur.nrows) {
      const int a = sr[tid] - cur.start, e = sr[tid + 1] - cur.start;
      double s = 0.0;
      for (int k = a; k < e; ++k) s += sv[k];
      ep.row(cur.n0 + tid, s, acc0, acc1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my generic writes precede the next bulk copy into this buffer
    __syncthreads();
    buf ^= 1;
    cur = nxt;
    nxt = nxt2;
  }
  if (EP::NACC >= 1) {
    acc0 = block_sum(acc0, red);
    if (tid == 0) ep.p0[blockIdx.x] = acc0;
  }
  if (EP::NACC >= 2) {
    acc1 = block_sum(acc1, red);
    if (tid == 0) ep.p1[blockIdx.x] = acc1;
  }
}

static int tma_grid(const cfem_ctx* c) {  // persistent: SMs x CTAs that fit (2 x 31 KB of shared memory each)
  const int64_t cap = (int64_t)c->sm_count * 3;
  return (int)(c->dm.ntiles < cap ? c->dm.ntiles : cap);
}

template <class EP>
static void launch_tile_spmv(cfem_ctx* c, const GhostSrc& gsrc, const Matrix& A, const double* x, const EP& ep, bool gated) {
  static bool configured[2] = {false, false};
  const size_t smem = 2 * (size_t)kTmaBuf;
  const int32_t* st = gated ? c->status : nullptr;
  int* err = (int*)(c->h_status + 6);  // pinned, host-visible
  if (gsrc.mbox) {
    if (!configured[1]) { CUDA_OK(cudaFuncSetAttribute(k_tile_spmv_tma<EP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured[1] = true; }
    k_tile_spmv_tma<EP, true><<<tma_grid(c), kTileNodes, smem, c->stream>>>(gsrc, c->dm.no, c->dm.tile_order, c->dm.n_interior, c->dm.ntiles,
                                                                           c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, x, ep, st, err);
  } else {
    if (!configured[0]) { CUDA_OK(cudaFuncSetAttribute(k_tile_spmv_tma<EP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured[0] = true; }
    k_tile_spmv_tma<EP, false><<<tma_grid(c), kTileNodes, smem, c->stream>>>(gsrc, c->dm.no, c->dm.tile_order, c->dm.n_interior, c->dm.ntiles,
                                                                            c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, x, ep, st, err);
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
struct EpPre { double a, b, c; };

template <int NDOT>
struct Ep16Spmv {  // y = A x with NDOT fused dot products
  static constexpr int NACC = NDOT;
  double* y;
  const double *d0, *d1;
  double *p0, *p1;
  __device__ __forceinline__ EpPre pre(int row) const {
    EpPre q{0.0, 0.0, 0.0};
    if (NDOT >= 1 && d0 != y) q.a = d0[row];
    if (NDOT >= 2 && d1 != y) q.b = d1[row];
    return q;
  }
  __device__ __forceinline__ void row(int row, double s, double, const EpPre& q, double& a0, double& a1) const {
    y[row] = s;
    if (NDOT >= 1) a0 += s * (d0 == y ? s : q.a);
    if (NDOT >= 2) a1 += s * (d1 == y ? s : q.b);
  }
};

template <bool FIRST>
struct Ep16Cheb {  // one Chebyshev iteration of the mass solve: r = b - M x, z = D^-1 r, d = c1 d + c2 z, x+ = x + d
  static constexpr int NACC = FIRST ? 2 : 1;
  const double *dinv, *b;
  double *xn, *d;
  double c1, c2;
  double *p0, *p1;  // ||r||^2 , ||b||^2 partials
  __device__ __forceinline__ EpPre pre(int row) const { return EpPre{b[row], dinv[row], FIRST ? 0.0 : d[row]}; }
  __device__ __forceinline__ void row(int row, double s, double xown, const EpPre& q, double& a0, double& a1) const {
    const double r = q.a - s, z = q.b * r;
    const double dk = FIRST ? c2 * z : c1 * q.c + c2 * z;
    d[row] = dk;
    xn[row] = xown + dk;
    a0 += r * r;
    if (FIRST) a1 += q.a * q.a;
  }
};

template <class EP, bool GHOST>
__global__ void __launch_bounds__(kTileNodes, CFEM_T16_MINB)
k_tile_t16(const GhostSrc gsrc, const int64_t no, const int32_t* __restrict__ tile_order, const int n_interior,
           const int ntiles, const int32_t* __restrict__ tile_node, const int32_t* __restrict__ rowptr,
           const uint16_t* __restrict__ lc16, const int32_t* __restrict__ extptr, const int32_t* __restrict__ ext,
           const double* __restrict__ vals, const double* __restrict__ x, const EP ep,
           const int32_t* __restrict__ status) {
  int bid = blockIdx.x, nblk = gridDim.x;
  if (GHOST && gsrc.pushdev) {  // CTA 0 is the producer half of the halo exchange (p2p.cuh)
    if (bid == 0) { pdl_wait(); pdl_launch(); push_cta(gsrc.pushdev, x, gsrc.seq, status && status[0]); return; }
    --bid; --nblk;
  }
  extern __shared__ double t16_smem[];
  double* const prod = t16_smem;               // [kTileNnzCap]
  double* const xs = t16_smem + kTileNnzCap;   // [kTileNodes + ext_cap]
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
    const int e0 = extptr[tile], ne = extptr[tile + 1] - e0;
    const int start = rowptr[n0], cnt = rowptr[n0 + nrows] - start;
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
      if (status && status[0]) return;
    }
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    if (tid < cnt) v0 = v[tid];
    if (tid + kTileNodes < cnt) v1 = v[tid + kTileNodes];
    if (tid + 2 * kTileNodes < cnt) v2 = v[tid + 2 * kTileNodes];
    if (tid + 3 * kTileNodes < cnt) v3 = v[tid + 3 * kTileNodes];
    if (GHOST && t >= n_interior && !waited) { ghost_wait(gsrc); waited = true; }
    // ---- stage x: own rows, then the external columns
    double xown = 0.0;
    if (tid < nrows) { xown = x[n0 + tid]; xs[tid] = xown; }
    if (tid < ne) xs[kTileNodes + tid] = XG(x, ecol);
    for (int e = tid + kTileNodes; e < ne; e += kTileNodes) { const int cc = ext[e0 + e]; xs[kTileNodes + e] = XG(x, cc); }
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
      ep.row(n0 + tid, s, xown, q, acc0, acc1);
    }
    __syncthreads();
  }
  if (EP::NACC >= 1) {
    acc0 = block_sum(acc0, red);
    if (tid == 0) ep.p0[bid] = acc0;
  }
  if (EP::NACC >= 2) {
    acc1 = block_sum(acc1, red);
    if (tid == 0) ep.p1[bid] = acc1;
  }
}

static size_t t16_smem_bytes(const cfem_ctx* c) { return sizeof(double) * ((size_t)kTileNnzCap + kTileNodes + c->dm.ext_cap); }

template <class EP, bool GHOST>
static int t16_prepare_one(cfem_ctx* c) {
  const size_t smem = t16_smem_bytes(c);
  CUDA_OK(cudaFuncSetAttribute(k_tile_t16<EP, GHOST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    occ = std::min(occ, t16_prepare_one<Ep16Cheb<true>, false>(c));
    occ = std::min(occ, t16_prepare_one<Ep16Cheb<false>, false>(c));
    if (c->world > 1 || getenv("CFEM_FORCE_GHOST")) {
      occ = std::min(occ, t16_prepare_one<Ep16Spmv<0>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Spmv<1>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Spmv<2>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Cheb<true>, true>(c));
      occ = std::min(occ, t16_prepare_one<Ep16Cheb<false>, true>(c));
    }
    const int64_t cap = (int64_t)c->sm_count * occ;
    c->t16_grid = (int)(c->dm.ntiles < cap ? c->dm.ntiles : cap);
  }
  return c->t16_grid;
}

template <class EP>
static void launch_t16(cfem_ctx* c, const GhostSrc& gsrc, const Matrix& A, const double* x, const EP& ep, bool gated) {
  const int grid = t16_grid(c);
  const size_t smem = t16_smem_bytes(c);
  const int32_t* st = gated ? c->status : nullptr;
  const DevMesh& m = c->dm;
  if (gsrc.mbox)
    launch_pdl(k_tile_t16<EP, true>, grid + (gsrc.pushdev ? 1 : 0), kTileNodes, smem, c->stream, gsrc, m.no, m.tile_order,
               m.n_interior, m.ntiles, m.tile_node, m.rowptr, m.lc16, m.tile_extptr, m.tile_ext, A.vals, x, ep, st);
  else
    launch_pdl(k_tile_t16<EP, false>, grid, kTileNodes, smem, c->stream, gsrc, m.no, m.tile_order, m.n_interior, m.ntiles,
               m.tile_node, m.rowptr, m.lc16, m.tile_extptr, m.tile_ext, A.vals, x, ep, st);
}

static inline int spmv_grid(cfem_ctx* c) {
  const int64_t cap = (int64_t)c->sm_count * 8;
  if (spmv_mode() == 0 && g_spmv_t16) return t16_grid(c);
  if (spmv_mode() == 0 && g_spmv_tma) return tma_grid(c);
  if (spmv_mode() == 0) return (int)(c->dm.ntiles < cap ? c->dm.ntiles : cap);
  const int64_t rows_per_block = (kBlock / 32) * 4;
  int64_t b = (c->dm.no + rows_per_block - 1) / rows_per_block;
  return (int)(b < cap ? b : cap);
}

template <int NDOT>
static void spmv_dots(cfem_ctx* c, const Matrix& A, const double* x, double* y, const double* d0,
                      const double* d1, double* p0, double* p1, bool gated) {
  GhostSrc gsrc;
  if (spmv_mode() == 0) gsrc = halo_push(c, const_cast<double*>(x), gated, !g_spmv_tma);  // producer half; the kernel waits in its boundary CTAs
  else halo_exchange(c, const_cast<double*>(x));
  ProfScope ps(c, PROF_SPMV);
  if (spmv_mode() == 0 && g_spmv_t16)
    launch_t16(c, gsrc, A, x, Ep16Spmv<NDOT>{y, d0, d1, p0, p1}, gated);
  else if (spmv_mode() == 0 && g_spmv_tma)
    launch_tile_spmv(c, gsrc, A, x, EpSpmv<NDOT>{y, d0, d1, p0, p1}, gated);
  else if (spmv_mode() == 0 && gsrc.mbox)
    launch_pdl(k_spmv_stream<NDOT, true>, spmv_grid(c) + (gsrc.pushdev ? 1 : 0), kTileNodes, 0, c->stream, gsrc, c->dm.no,
               c->dm.tile_order, c->dm.n_interior, c->dm.ntiles, c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, x, y,
               d0, d1, p0, p1, gated ? c->status : nullptr);
  else if (spmv_mode() == 0)
    launch_pdl(k_spmv_stream<NDOT, false>, spmv_grid(c), kTileNodes, 0, c->stream, gsrc, c->dm.no, c->dm.tile_order,
               c->dm.n_interior, c->dm.ntiles, c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, x, y, d0, d1, p0, p1,
               gated ? c->status : nullptr);
  else
    k_spmv<8, NDOT><<<spmv_grid(c), kBlock, 0, c->stream>>>(c->dm.no, c->dm.rowptr, c->dm.colidx, A.vals, x, y, d0,
                                                            d1, p0, p1, gated ? c->status : nullptr);
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
  spmv_dots<2>(c, A, x, y, d0, y, p0, p1, false);
}

// ---------------------------------------------------------------- Chebyshev (mass matrix)
// For P1 triangles every eigenvalue of D^-1 M lies in [1/2, 2] (element-wise bound,
// inherited by the Dirichlet-reduced matrix), so the Chebyshev semi-iteration needs
// no inner products: one fused kernel per iteration (SpMV + residual + update), no
// reductions, no host round trips until the final check.
//   r_k = b - M x_k ; z_k = D^-1 r_k ; d_k = c1 d_{k-1} + c2 z_k ; x_{k+1} = x_k + d_k
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
    if (bid == 0) { pdl_wait(); pdl_launch(); push_cta(gsrc.pushdev, xk, gsrc.seq, false); return; }
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
    if (GHOST && t >= n_interior && !waited) { ghost_wait(gsrc); waited = true; }
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
      rr += r * r;
      if (FIRST) bb += bi * bi;
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
  while (true) {
    // the iterations up to the next check are timed as ONE scope of (target - it) launches: the per-launch
    // figure then is the in-situ one (back-to-back launches, programmatic dependent launch active)
    {
    ProfScope chain(c, PROF_CHEB, target - it);
    for (; it < target; ++it) {
      const GhostSrc gsrc = halo_push(c, xa, false, !g_spmv_tma);
      ProfScope ps(c, PROF_CHEB);
#define CHEB_LAUNCH(FIRST, GHOST, C1, C2, PBB)                                                                   \
  launch_pdl(k_cheb_stream<FIRST, GHOST>, gs + (GHOST && gsrc.pushdev ? 1 : 0), kTileNodes, 0, c->stream, gsrc, c->dm.no, \
             c->dm.tile_order, c->dm.n_interior, c->dm.ntiles, c->dm.tile_node, c->dm.rowptr, c->dm.colidx, A.vals, A.dinv, \
             b, xa, xb, d, C1, C2, part + P_RR * kMaxPartials, PBB)
      if (g_spmv_t16 && it == 0) {
        launch_t16(c, gsrc, A, xa, Ep16Cheb<true>{A.dinv, b, xb, d, 0.0, 1.0 / theta, part + P_RR * kMaxPartials, part + P_BB * kMaxPartials}, false);
      } else if (g_spmv_t16) {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        launch_t16(c, gsrc, A, xa, Ep16Cheb<false>{A.dinv, b, xb, d, rho_new * rho, 2.0 * rho_new / delta, part + P_RR * kMaxPartials, nullptr}, false);
        rho = rho_new;
      } else if (g_spmv_tma && it == 0) {
        launch_tile_spmv(c, gsrc, A, xa, EpCheb<true>{A.dinv, b, xa, xb, d, 0.0, 1.0 / theta, part + P_RR * kMaxPartials, part + P_BB * kMaxPartials}, false);
      } else if (g_spmv_tma) {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        launch_tile_spmv(c, gsrc, A, xa, EpCheb<false>{A.dinv, b, xa, xb, d, rho_new * rho, 2.0 * rho_new / delta, part + P_RR * kMaxPartials, nullptr}, false);
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
      if (it == 0) np_bb = allreduce_sum1(c, part + P_BB * kMaxPartials, gs);
      std::swap(xa, xb);
    }
    }
    // the last kernel measured ||b - M x_{it-1}||; x_it is one update further on
    const int np_rr = allreduce_sum1(c, part + P_RR * kMaxPartials, gs);
    { ProfScope ps(c, PROF_KRYLOV_VEC); launch_pdl(k_relres, 1, kBlock, 0, c->stream, part, np_rr, np_bb, c->scalars); LAUNCHED(c); }
    CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + S_RELRES, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
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
    const double bi = b[i], ri = bi - q[i], zi = dinv[i] * ri;
    r[i] = ri; z[i] = zi; p[i] = zi;
    rz += ri * zi; rr += ri * ri; bb += bi * bi;
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
    nrz += ri * zi; rr += ri * ri;
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
  CUDA_OK(cudaMemcpyAsync(c->h_status, c->status, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + S_RELRES, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
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

SolveResult bicgstab(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol,
                     int max_it, int* predict) {
  l2_prefer(c, A);
  LinApply op = [&](const double* xin, double* yout, int ndot, const double* d0, const double* d1, double* p0,
                    double* p1, bool gated) {
    if (ndot == 0) spmv_dots<0>(c, A, xin, yout, nullptr, nullptr, nullptr, nullptr, gated);
    else if (ndot == 1) spmv_dots<1>(c, A, xin, yout, d0, nullptr, p0, nullptr, gated);
    else spmv_dots<2>(c, A, xin, yout, d0, d1, p0, p1, gated);
    return c->world > 1 ? 1 : spmv_grid(c);
  };
  return bicgstab_generic(c, c->dm.no, 1, A.dinv, op, c->wk, b, x, rtol, atol, max_it, predict);
}

// ---------------------------------------------------------------- restarted GMRES(30), right Jacobi
// Arnoldi with classical Gram-Schmidt (all inner products of a step in one pass over the basis), Givens
// rotations and the small triangular solve on the device; the host only polls the done flag.
//   per step:  w = A (D^-1 v_j) ; h_i = (w, v_i), i <= j ; w -= sum h_i v_i ; v_{j+1} = w / ||w||
constexpr int kGmresM = 30;
// layout of the small device block (doubles): H (31 x 30, column major) | cs[30] | sn[30] | g[31] | y[30] | misc
constexpr int kGmH = 0, kGmCs = 31 * 30, kGmSn = kGmCs + 30, kGmG = kGmSn + 30, kGmY = kGmG + 31, kGmMisc = kGmY + 30,
              kGmSmall = kGmMisc + 8;

__global__ void __launch_bounds__(kBlock)
k_gm_residual(int64_t n, const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ w,
              double* __restrict__ part_rr, double* __restrict__ part_bb) {
  __shared__ double red[9];
  double rr = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double bi = b[i], ri = bi - q[i];
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

// v = w / nrm ; z = D^-1 v     (nrm read from the small block: slot kGmMisc+1)
__global__ void __launch_bounds__(kBlock)
k_gm_normalize(int64_t n, const double* __restrict__ w, const double* __restrict__ dinv, const double* __restrict__ sm,
               double* __restrict__ v, double* __restrict__ z, const int32_t* __restrict__ status) {
  if (status[0]) return;
  const double inv = 1.0 / sm[kGmMisc + 1];
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double vi = w[i] * inv;
    v[i] = vi; z[i] = dinv[i] * vi;
  }
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

// x += D^-1 sum_i y_i v_i
__global__ void __launch_bounds__(kBlock)
k_gm_xupdate(int64_t n, int64_t stride, const double* __restrict__ V, const double* __restrict__ dinv,
             const double* __restrict__ sm, double* __restrict__ x) {
  __shared__ double y[kGmresM];
  const int k = (int)sm[kGmMisc];
  if (threadIdx.x < k) y[threadIdx.x] = sm[kGmY + threadIdx.x];
  __syncthreads();
  for (int64_t e = blockIdx.x * (int64_t)kBlock + threadIdx.x; e < n; e += (int64_t)gridDim.x * kBlock) {
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += y[i] * V[(size_t)i * stride + e];
    x[e] += dinv[e] * s;
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
  double *w = c->wk[0], *z = c->wk[1], *q = c->wk[2];
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
      k_gm_residual<<<gv, kBlock, 0, c->stream>>>(n, b, q, w, part + P_RR * kMaxPartials, part + P_BB * kMaxPartials); LAUNCHED(c); }
    int np = gv;
    if (dist) { double* sl[2] = {part + P_RR * kMaxPartials, part + P_BB * kMaxPartials}; const int op[2] = {0, 0}; np = allreduce_partials(c, 2, sl, op, gv); }
    { ProfScope ps(c, PROF_KRYLOV_VEC);
      k_gm_start<<<1, kBlock, 0, c->stream>>>(part + P_RR * kMaxPartials, part + P_BB * kMaxPartials, np, sm, c->scalars, c->status, rtol2, atol2, cycle == 0); LAUNCHED(c); }
    if (cycle > 0 || total >= max_it) {  // the first cycle defers this poll to the inner loop (x0 is rarely converged)
      if (poll_done(c, res) || total >= max_it) break;
    }
    { ProfScope ps(c, PROF_KRYLOV_VEC); k_gm_normalize<<<gv, kBlock, 0, c->stream>>>(n, w, A.dinv, sm, V, z, c->status); LAUNCHED(c); }
    for (int j = 0; j < kGmresM && total < max_it; ++j) {
      spmv_dots<0>(c, A, z, w, nullptr, nullptr, nullptr, nullptr, true);
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
        k_gm_normalize<<<gv, kBlock, 0, c->stream>>>(n, w, A.dinv, sm, V + (size_t)(j + 1) * nl, z, c->status); LAUNCHED(c); }
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
      k_gm_xupdate<<<gv, kBlock, 0, c->stream>>>(n, nl, V, A.dinv, sm, x); LAUNCHED(c); }
  }
  if (!res.converged) poll_done(c, res);
  halo_exchange(c, x);
  if (predict) *predict = res.iters > 0 ? res.iters : 1;
  return res;
}

}  // namespace cfem
