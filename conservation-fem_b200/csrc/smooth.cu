// smooth_vector post-filter of the reference (Code/Utils/helpers.py:40-50), with its exact semantics: an
// IN-PLACE sweep over the nodes in a given order,
//     u[i] <- (sum_{j in patch(i), j != i} u[j] + (l - 1) d_i u[i]) / (l d_i),   d_i = |patch(i)| - 1,
// so node i sees the already-smoothed values of the neighbours that come earlier in the order and the old
// values of the later ones (a Gauss-Seidel-like sweep, Exact_Burger_SI.py:193 calls it every step).
//
// Parallel form: level scheduling.  level(i) = 1 + max level of the neighbours that precede i in the sweep;
// two neighbours never share a level and the levels respect the order along every edge, so processing level
// after level -- all nodes of a level at once -- reproduces the sequential sweep exactly (up to the order of
// the <= ~8 additions inside one patch sum, which the reference leaves to Python's set iteration).
//  * many narrow levels (orderings that follow the mesh: ~nx + 2 ny levels on a structured mesh): ONE CTA of
//    1024 threads walks all levels with a __syncthreads() between them -- a level is narrower than the CTA
//    and the barrier costs tens of cycles instead of a kernel boundary;
//  * few wide levels (random orderings): one grid-wide launch per level.
#include <algorithm>
#include <numeric>

#include "device_utils.cuh"
#include "launch.h"

namespace cfem {

struct SmoothPlan {
  uint64_t key = 0;
  int nlevels = 0, max_width = 0;
  int32_t* d_nodes = nullptr;      // internal node ids, level after level
  int32_t* d_level_ptr = nullptr;  // nlevels + 1
  std::vector<int32_t> level_ptr;
};

static uint64_t order_key(const int32_t* order, int64_t n) {
  uint64_t h = 1469598103934665603ull ^ (uint64_t)n;
  if (!order) return h;
  for (int64_t i = 0; i < n; ++i) { h ^= (uint64_t)(uint32_t)order[i]; h *= 1099511628211ull; }
  return h;
}

// order: caller dof ids in sweep order (a permutation of 0..n-1), or nullptr for ascending caller ids
void build_smooth_levels(const HostMesh& hm, const int32_t* order, std::vector<int32_t>& nodes,
                         std::vector<int32_t>& level_ptr) {
  const int64_t n = hm.nn;
  std::vector<int32_t> level(n, 0);
  std::vector<uint8_t> seen(order ? n : 0, 0);
  int nlev = 0;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t user = order ? order[k] : k;
    if (user < 0 || user >= n) CFEM_THROW(-1, "smooth_vector: sweep order entry out of range");
    if (order) { if (seen[user]) CFEM_THROW(-1, "smooth_vector: sweep order is not a permutation"); seen[user] = 1; }
    const int32_t i = hm.u2n[user];
    int lv = 0;
    for (int p = hm.rowptr[i]; p < hm.rowptr[i + 1]; ++p) lv = std::max(lv, level[hm.colidx[p]]);  // unvisited = 0; own entry = 0
    level[i] = lv + 1;
    nlev = std::max(nlev, lv + 1);
  }
  level_ptr.assign(nlev + 1, 0);
  for (int64_t i = 0; i < n; ++i) level_ptr[level[i]]++;          // level l counted in slot l (1-based)
  for (int l = 0; l < nlev; ++l) level_ptr[l + 1] += level_ptr[l];
  nodes.resize(n);
  std::vector<int32_t> fill(level_ptr.begin(), level_ptr.end() - 1);
  for (int64_t i = 0; i < n; ++i) nodes[fill[level[i] - 1]++] = (int32_t)i;  // ascending internal id inside a level
}

__device__ __forceinline__ void smooth_node(const int i, const int32_t* __restrict__ rowptr,
                                            const int32_t* __restrict__ colidx, double* u, const double l) {
  const int p0 = rowptr[i], p1 = rowptr[i + 1];
  double s = 0.0;
  for (int p = p0; p < p1; ++p) {
    const int j = colidx[p];
    if (j != i) s += __ldcg(u + j);   // L2 view: values written by other SMs / earlier levels
  }
  const double d = (double)(p1 - p0 - 1);
  u[i] = (s + (l - 1.0) * d * __ldcg(u + i)) / (l * d);
}

__global__ void __launch_bounds__(1024)
k_smooth_walk(const int nlevels, const int32_t* __restrict__ level_ptr, const int32_t* __restrict__ nodes,
              const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, double* u, const double l) {
  for (int lv = 0; lv < nlevels; ++lv) {
    const int a = level_ptr[lv], b = level_ptr[lv + 1];
    for (int k = a + threadIdx.x; k < b; k += blockDim.x) smooth_node(nodes[k], rowptr, colidx, u, l);
    __threadfence_block();
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kBlock)
k_smooth_level(const int a, const int b, const int32_t* __restrict__ nodes, const int32_t* __restrict__ rowptr,
               const int32_t* __restrict__ colidx, double* u, const double l) {
  for (int k = a + blockIdx.x * kBlock + threadIdx.x; k < b; k += gridDim.x * kBlock)
    smooth_node(nodes[k], rowptr, colidx, u, l);
}

// u: device vector in internal numbering (all nn entries)
void launch_smooth_vector(cfem_ctx* c, double* u, const int32_t* order_host, double l) {
  if (c->world > 1) CFEM_THROW(-1, "smooth_vector: the sequential sweep is not partitioned; single-GPU contexts only");
  if (!(l > 0.0)) CFEM_THROW(-1, "smooth_vector: l must be positive");
  const int64_t n = c->dm.nn;
  const uint64_t key = order_key(order_host, n);
  SmoothPlan* plan = (SmoothPlan*)c->smooth_plan;
  if (!plan || plan->key != key) {
    if (!plan) { plan = new SmoothPlan(); c->smooth_plan = plan; }
    std::vector<int32_t> nodes;
    build_smooth_levels(c->hm, order_host, nodes, plan->level_ptr);
    plan->key = key;
    plan->nlevels = (int)plan->level_ptr.size() - 1;
    plan->max_width = 0;
    for (int lv = 0; lv < plan->nlevels; ++lv)
      plan->max_width = std::max(plan->max_width, plan->level_ptr[lv + 1] - plan->level_ptr[lv]);
    if (!plan->d_nodes) {
      CUDA_OK(cudaMalloc((void**)&plan->d_nodes, n * sizeof(int32_t)));
      CUDA_OK(cudaMalloc((void**)&plan->d_level_ptr, (n + 1) * sizeof(int32_t)));
      c->allocs.push_back(plan->d_nodes);
      c->allocs.push_back(plan->d_level_ptr);
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(plan->d_nodes, nodes.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(plan->d_level_ptr, plan->level_ptr.data(), plan->level_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  ProfScope ps(c, PROF_MISC);
  if (plan->nlevels > 64) {
    k_smooth_walk<<<1, 1024, 0, c->stream>>>(plan->nlevels, plan->d_level_ptr, plan->d_nodes, c->dm.rowptr, c->dm.colidx, u, l);
    CUDA_OK(cudaGetLastError()); c->launches.total++;
  } else {
    for (int lv = 0; lv < plan->nlevels; ++lv) {
      const int a = plan->level_ptr[lv], b = plan->level_ptr[lv + 1];
      int g = (b - a + kBlock - 1) / kBlock;
      g = std::max(1, std::min(g, c->sm_count * 8));
      k_smooth_level<<<g, kBlock, 0, c->stream>>>(a, b, plan->d_nodes, c->dm.rowptr, c->dm.colidx, u, l);
      CUDA_OK(cudaGetLastError()); c->launches.total++;
    }
  }
}

void smooth_plan_free(cfem_ctx* c) {
  delete (SmoothPlan*)c->smooth_plan;
  c->smooth_plan = nullptr;
}

}  // namespace cfem
