// Device helpers: deterministic block reductions, geometry, quadrature tables.
#pragma once
#include "internal.h"

namespace cfem {

constexpr int kBlock = 256;  // threads per CTA for every kernel in the library
constexpr size_t kDynSmemCeiling = 200 * 1024;  // dynamic shared memory every tile kernel is opted in to (sm_100: 227 KB per CTA)

// Programmatic dependent launch (griddepcontrol): a kernel launched with launch_pdl() may become resident while
// the previous kernel of the stream is still draining.  pdl_wait() blocks until that kernel has completed and
// its writes are visible (a no-op for ordinary launches) -- nothing that depends on the stream order may be
// touched before it; pdl_launch() lets the NEXT kernel's CTAs take the slots this grid frees as it drains.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Host side: same as kern<<<grid, block, smem, stream>>>(args...) with programmatic stream serialisation allowed.
// Only for kernels whose first statement is pdl_wait().  CFEM_PDL=0 falls back to plain launches.
inline bool pdl_enabled() {
  static const bool on = !(getenv("CFEM_PDL") && std::string(getenv("CFEM_PDL")) == "0");
  return on;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum over the CTA (blockDim.x == kBlock), result valid in every thread.
// Fixed shuffle tree + fixed warp order -> bitwise reproducible.
__device__ __forceinline__ double block_sum(double v, double* sh /* >= 9 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < (kBlock / 32) ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[8] = t;
  }
  __syncthreads();
  return sh[8];
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < (kBlock / 32) ? sh[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) sh[8] = t;
  }
  __syncthreads();
  return sh[8];
}
__device__ __forceinline__ double block_min(double v, double* sh) { return -block_max(-v, sh); }

// Every CTA re-reduces the per-CTA partials a previous kernel wrote (n small,
// L2-resident): removes separate "finalise" launches, keeps a fixed order.
__device__ __forceinline__ double reduce_partials(const double* __restrict__ p, int n, double* sh) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += kBlock) s += p[i];
  return block_sum(s, sh);
}

// ---- per-CTA tile schedule of the SpMV-type kernels -------------------------------------------------------------
// A CTA visits tiles  wid, wid + nwork, wid + 2 nwork, ... ("rounds") of tile_order.  What a tile needs before its first
// data load -- tile id, row range, external-column range, CSR range -- is a chain of three dependent table loads;
// fetched per tile it costs an L2 latency or two at every tile start (and one more in the distributed variants, which
// go through tile_order).  fetch_tile_meta() resolves the chain for ALL rounds of the CTA at once (one thread per round)
// into shared memory: one latency chain per launch -- or, in the persistent solvers, per solve.
constexpr int kMetaRounds = 16;   // schedules longer than this fall back to per-tile fetches
struct TileMeta { int t, n0, nrows, e0, ne, start, cnt; };

// Order in which a CTA visits its rounds.  tile_order keeps the tiles with ghost columns at the end, i.e. in the last
// round; the distributed variants visit that round in the MIDDLE, so the neighbours' halo values (pushed at the start
// of the kernel / phase) have time to arrive and whatever wait remains is followed by more work of the same CTA.
template <bool GHOST>
__device__ __forceinline__ int round_of(const int kk, const int nrounds) {
  if (!GHOST || nrounds < 3) return kk;
  const int mid = nrounds / 2;
  return kk == mid ? nrounds - 1 : (kk > mid ? kk - 1 : kk);
}

template <bool GHOST, bool MID = GHOST>
__device__ __forceinline__ TileMeta tile_meta_of(const int kk, const int nrounds, const int wid, const int nwork,
                                                 const int ntiles, const int32_t* __restrict__ tile_order,
                                                 const int32_t* __restrict__ tile_node, const int32_t* __restrict__ extptr,
                                                 const int32_t* __restrict__ rowptr) {
  TileMeta m;
  m.t = wid + round_of<MID>(kk, nrounds) * nwork;
  if (m.t >= ntiles) { m.t = -1; m.n0 = m.nrows = m.e0 = m.ne = m.start = m.cnt = 0; return m; }
  const int tile = GHOST ? tile_order[m.t] : m.t;
  m.n0 = tile_node[tile];
  m.nrows = tile_node[tile + 1] - m.n0;
  m.e0 = extptr[tile];
  m.ne = extptr[tile + 1] - m.e0;
  m.start = rowptr[m.n0];
  m.cnt = rowptr[m.n0 + m.nrows] - m.start;
  return m;
}

// all threads of the CTA; smeta: kMetaRounds entries of shared memory; followed by a __syncthreads() of the caller
template <bool GHOST, bool MID = GHOST>
__device__ __forceinline__ void fetch_tile_meta(TileMeta* smeta, const int nrounds, const int wid, const int nwork,
                                                const int ntiles, const int32_t* __restrict__ tile_order,
                                                const int32_t* __restrict__ tile_node, const int32_t* __restrict__ extptr,
                                                const int32_t* __restrict__ rowptr) {
  if (nrounds <= kMetaRounds && (int)threadIdx.x < nrounds)
    smeta[threadIdx.x] = tile_meta_of<GHOST, MID>(threadIdx.x, nrounds, wid, nwork, ntiles, tile_order, tile_node, extptr, rowptr);
}

struct CellGeom {
  double gx[3], gy[3];  // gradients of the three P1 basis functions
  double area;
};

__device__ __forceinline__ CellGeom cell_geom(const double2 p0, const double2 p1, const double2 p2) {
  CellGeom g;
  const double e1x = p1.x - p0.x, e1y = p1.y - p0.y;
  const double e2x = p2.x - p0.x, e2y = p2.y - p0.y;
  const double det = e1x * e2y - e1y * e2x;
  const double inv = 1.0 / det;
  g.gx[1] = e2y * inv;
  g.gy[1] = -e2x * inv;
  g.gx[2] = -e1y * inv;
  g.gy[2] = e1x * inv;
  g.gx[0] = -g.gx[1] - g.gx[2];
  g.gy[0] = -g.gy[1] - g.gy[2];
  g.area = 0.5 * fabs(det);
  return g;
}

// Symmetric triangle rules (barycentric points, weights sum to 1):
// degree 4 / 6 points and degree 5 / 7 points — the rules FFCx selects for the
// KPP residual and Jacobian forms.
constexpr double kQ4a = 0.4459484909159648863183292538830519883991;
constexpr double kQ4b = 0.09157621350977074345957146340220150785433;
constexpr double kQ4wa = 0.2233815896780114656950070084331228043703;
constexpr double kQ4wb = 0.1099517436553218676383263249002105289631;
constexpr double kQ5a = 0.1012865073234563388009873619151238280556;
constexpr double kQ5b = 0.4701420641051150897704412095134476005159;
constexpr double kQ5wa = 0.1259391805448271525956839455001813336576;
constexpr double kQ5wb = 0.1323941527885061807376493878331519996757;
constexpr double kQ5w0 = 0.225;

}  // namespace cfem
