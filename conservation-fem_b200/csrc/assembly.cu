// Tile assembly: cell-local quadrature staged in shared memory (phase A),
// then a deterministic node-centric gather into vectors / CSR rows (phase B).
//
// A tile is a run of <= kTileNodes Hilbert-consecutive nodes plus the list of
// every cell touching them.  Phase A: one thread per cell gathers the three
// vertices' coordinates and nodal fields, evaluates the element vector(s) and
// 3x3 element matrix once, and stores them in shared memory.  Phase B: one
// thread per node walks its vertex->cell adjacency (cells ascending, fixed
// order, no atomics) and sums the staged element entries into its vector entry
// and its CSR row; the tile's rows are contiguous in the CSR value array and
// are written back coalesced.
//
// Replaces dolfinx assemble_vector / assemble_matrix / apply_lifting / set_bc
// as called from the reference loops (Code/Linear_advection/RV_node.py:224-242,
// Code/KPP/KPP_exact.py:123-154, Code/Utils/helpers.py:29-36).
#include "device_utils.cuh"
#include "launch.h"
#include "p2p.cuh"

#ifndef CFEM_RES_PREFETCH
#define CFEM_RES_PREFETCH true    // A/B switch of the residual op's phase-A prefetch (see k_tile_assemble)
#endif

namespace cfem {

// ------------------------------------------------------------------ fluxes
// flux_vec: out[a] = int f'(u).grad(u) phi_a ;  flux_jac: J[a*3+b] = d out[a] / d u_b
template <int FLUX>
__device__ __forceinline__ void flux_vec(const CellGeom& g, const double u[3], double out[3]);
template <int FLUX>
__device__ __forceinline__ void flux_jac(const CellGeom& g, const double u[3], double J[9]);

// Burgers f'(u) = (u,u)  (reference Code/Burgers_equation/Exact_Burger_RV.py:33-35)
template <>
__device__ __forceinline__ void flux_vec<CFEM_FLUX_BURGERS>(const CellGeom& g, const double u[3], double out[3]) {
  const double s = u[0] * (g.gx[0] + g.gy[0]) + u[1] * (g.gx[1] + g.gy[1]) + u[2] * (g.gx[2] + g.gy[2]);
  const double su = u[0] + u[1] + u[2];
  const double m = g.area * (1.0 / 12.0);
#pragma unroll
  for (int a = 0; a < 3; ++a) out[a] = s * (m * (su + u[a]));
}
template <>
__device__ __forceinline__ void flux_jac<CFEM_FLUX_BURGERS>(const CellGeom& g, const double u[3], double J[9]) {
  const double s = u[0] * (g.gx[0] + g.gy[0]) + u[1] * (g.gx[1] + g.gy[1]) + u[2] * (g.gx[2] + g.gy[2]);
  const double su = u[0] + u[1] + u[2];
  const double m = g.area * (1.0 / 12.0);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double Mu = m * (su + u[a]);
#pragma unroll
    for (int b = 0; b < 3; ++b) J[a * 3 + b] = (g.gx[b] + g.gy[b]) * Mu + s * m * (a == b ? 2.0 : 1.0);
  }
}

// KPP f'(u) = (cos u, -sin u)  (reference Code/KPP/KPP_exact.py:55-57)
__device__ __forceinline__ void kpp_vec_point(double l0, double l1, double l2, double w, const double u[3],
                                              double ux, double uy, double out[3]) {
  double s, c;
  sincos(l0 * u[0] + l1 * u[1] + l2 * u[2], &s, &c);
  const double v = w * (c * ux - s * uy);
  out[0] += v * l0;
  out[1] += v * l1;
  out[2] += v * l2;
}
template <>
__device__ __forceinline__ void flux_vec<CFEM_FLUX_KPP>(const CellGeom& g, const double u[3], double out[3]) {
  const double ux = u[0] * g.gx[0] + u[1] * g.gx[1] + u[2] * g.gx[2];
  const double uy = u[0] * g.gy[0] + u[1] * g.gy[1] + u[2] * g.gy[2];
  double o[3] = {0.0, 0.0, 0.0};
  constexpr double a = kQ4a, ac = 1.0 - 2.0 * kQ4a, b = kQ4b, bc = 1.0 - 2.0 * kQ4b;
  kpp_vec_point(ac, a, a, kQ4wa, u, ux, uy, o);
  kpp_vec_point(a, ac, a, kQ4wa, u, ux, uy, o);
  kpp_vec_point(a, a, ac, kQ4wa, u, ux, uy, o);
  kpp_vec_point(bc, b, b, kQ4wb, u, ux, uy, o);
  kpp_vec_point(b, bc, b, kQ4wb, u, ux, uy, o);
  kpp_vec_point(b, b, bc, kQ4wb, u, ux, uy, o);
#pragma unroll
  for (int k = 0; k < 3; ++k) out[k] = g.area * o[k];
}
__device__ __forceinline__ void kpp_jac_point(double l0, double l1, double l2, double w, const CellGeom& g,
                                              const double u[3], double ux, double uy, double J[9]) {
  double s, c;
  sincos(l0 * u[0] + l1 * u[1] + l2 * u[2], &s, &c);
  const double d = -s * ux - c * uy;
  const double l[3] = {l0, l1, l2};
  double col[3];
#pragma unroll
  for (int b = 0; b < 3; ++b) col[b] = d * l[b] + c * g.gx[b] - s * g.gy[b];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double wa = w * l[a];
#pragma unroll
    for (int b = 0; b < 3; ++b) J[a * 3 + b] += wa * col[b];
  }
}
template <>
__device__ __forceinline__ void flux_jac<CFEM_FLUX_KPP>(const CellGeom& g, const double u[3], double J[9]) {
  const double ux = u[0] * g.gx[0] + u[1] * g.gx[1] + u[2] * g.gx[2];
  const double uy = u[0] * g.gy[0] + u[1] * g.gy[1] + u[2] * g.gy[2];
  double o[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = 0.0;
  constexpr double a = kQ5a, ac = 1.0 - 2.0 * kQ5a, b = kQ5b, bc = 1.0 - 2.0 * kQ5b, t = 1.0 / 3.0;
  kpp_jac_point(t, t, t, kQ5w0, g, u, ux, uy, o);
  kpp_jac_point(ac, a, a, kQ5wa, g, u, ux, uy, o);
  kpp_jac_point(a, ac, a, kQ5wa, g, u, ux, uy, o);
  kpp_jac_point(a, a, ac, kQ5wa, g, u, ux, uy, o);
  kpp_jac_point(bc, b, b, kQ5wb, g, u, ux, uy, o);
  kpp_jac_point(b, bc, b, kQ5wb, g, u, ux, uy, o);
  kpp_jac_point(b, b, bc, kQ5wb, g, u, ux, uy, o);
#pragma unroll
  for (int k = 0; k < 9; ++k) J[k] = g.area * o[k];
}

// ------------------------------------------------------------------ element ops
// An Op provides
//   cell(v, g, bcmask, evec[NV*3], emat[9])  — element vector(s) / matrix
//   node(i, acc[NV]) -> double               — finalise + store; returned value is
//                                              block-summed into the partials
__device__ __forceinline__ void mass_elem(const CellGeom& g, double emat[9]) {
  const double m = g.area * (1.0 / 12.0);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) emat[a * 3 + b] = (a == b) ? 2.0 * m : m;
}

struct MassOp {
  static constexpr int NV = 0;
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = true;
  __device__ void cell(const int32_t*, const CellGeom& g, int, double*, double* emat) const { mass_elem(g, emat); }
  __device__ double node(int32_t, const double*) const { return 0.0; }
};

struct StiffnessOp {
  static constexpr int NV = 0;
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = true;
  const double* eps;  // nullable
  __device__ void cell(const int32_t* v, const CellGeom& g, int, double*, double* emat) const {
    const double e = eps ? (eps[v[0]] + eps[v[1]] + eps[v[2]]) * (1.0 / 3.0) : 1.0;
    const double f = g.area * e;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) emat[a * 3 + b] = f * (g.gx[a] * g.gx[b] + g.gy[a] * g.gy[b]);
  }
  __device__ double node(int32_t, const double*) const { return 0.0; }
};

// Euler path (csrc/euler.cu): Cd[a][b] = int phi_a d_d phi_b = |K|/3 grad_d phi_b
struct GradOp {
  static constexpr int NV = 0;
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = true;
  int d;
  __device__ void cell(const int32_t*, const CellGeom& g, int, double*, double* emat) const {
    const double f = g.area * (1.0 / 3.0);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) emat[a * 3 + b] = f * (d == 0 ? g.gx[b] : g.gy[b]);
  }
  __device__ double node(int32_t, const double*) const { return 0.0; }
};

// S = M + coef K_eps (no Dirichlet handling: the Euler apply kernels do it by row)
struct MassStiffOp {
  static constexpr int NV = 0;
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = true;
  const double* eps;
  double coef;
  __device__ void cell(const int32_t* v, const CellGeom& g, int, double*, double* emat) const {
    const double e = (eps[v[0]] + eps[v[1]] + eps[v[2]]) * (1.0 / 3.0);
    const double m = g.area * (1.0 / 12.0), kf = coef * e * g.area;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        emat[a * 3 + b] = (a == b ? 2.0 * m : m) + kf * (g.gx[a] * g.gx[b] + g.gy[a] * g.gy[b]);
  }
  __device__ double node(int32_t, const double*) const { return 0.0; }
};

// b_i = sum_K h_K |K| / 3, h_K = shortest edge  (reference Code/Utils/helpers.py:18-31)
struct NodalHOp {
  static constexpr int NV = 1;
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = false;
  const double2* xy;
  double* b;
  __device__ void cell(const int32_t* v, const CellGeom& g, int, double* evec, double*) const {
    const double2 p0 = xy[v[0]], p1 = xy[v[1]], p2 = xy[v[2]];
    const double e01 = sqrt((p0.x - p1.x) * (p0.x - p1.x) + (p0.y - p1.y) * (p0.y - p1.y));
    const double e02 = sqrt((p0.x - p2.x) * (p0.x - p2.x) + (p0.y - p2.y) * (p0.y - p2.y));
    const double e12 = sqrt((p1.x - p2.x) * (p1.x - p2.x) + (p1.y - p2.y) * (p1.y - p2.y));
    const double hk = fmin(fmin(e01, e02), e12);
    evec[0] = evec[1] = evec[2] = hk * g.area / 3.0;
  }
  __device__ double node(int32_t i, const double* acc) const { b[i] = acc[0]; return 0.0; }
};

// RV residual right-hand side (reference Code/KPP/KPP_exact.py:123-127,
// Exact_Burger_RV.py:187-191, Exact_Burger_RV_conv.py:186, RV_node.py:209-210)
template <int FLUX, int NVEC>
struct RvRhsOp {
  static constexpr int NV = NVEC;  // 1: b only; 2: b and nodal flux(u_n)
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = false;
  const double *u_n, *u_old, *u_oo;
  const double2* w;
  const uint8_t* is_bc;  // null -> no bc
  double c_n, c_old, c_oo;  // D_t u = c_n u_n + c_old u_old + c_oo u_oo
  double *b, *fluxn;
  __device__ void cell(const int32_t* v, const CellGeom& g, int, double* evec, double*) const {
    double u[3], D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      u[k] = u_n[v[k]];
      D[k] = c_n * u[k] + c_old * u_old[v[k]];
      if (u_oo) D[k] += c_oo * u_oo[v[k]];
    }
    double f[3];
    if constexpr (FLUX == CFEM_FLUX_ADVECTION) {
      // int (w . grad u_n) phi_a, w P1:  |K|/12 * (W + w_a) . grad u_n
      const double ux = u[0] * g.gx[0] + u[1] * g.gx[1] + u[2] * g.gx[2];
      const double uy = u[0] * g.gy[0] + u[1] * g.gy[1] + u[2] * g.gy[2];
      const double2 w0 = w[v[0]], w1 = w[v[1]], w2 = w[v[2]];
      const double Wx = w0.x + w1.x + w2.x, Wy = w0.y + w1.y + w2.y;
      const double m = g.area * (1.0 / 12.0);
      f[0] = m * ((Wx + w0.x) * ux + (Wy + w0.y) * uy);
      f[1] = m * ((Wx + w1.x) * ux + (Wy + w1.y) * uy);
      f[2] = m * ((Wx + w2.x) * ux + (Wy + w2.y) * uy);
    } else {
      flux_vec<FLUX>(g, u, f);
    }
    const double m = g.area * (1.0 / 12.0), sd = D[0] + D[1] + D[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      evec[a] = m * (sd + D[a]) + f[a];
      if (NVEC == 2) evec[3 + a] = f[a];
    }
  }
  __device__ double node(int32_t i, const double* acc) const {
    b[i] = (is_bc && is_bc[i]) ? 0.0 : acc[0];
    if (NVEC == 2) fluxn[i] = acc[1];
    return 0.0;
  }
};

// Crank-Nicolson residual F(uh) with dolfinx NonlinearProblem.F bc handling
// (reference Code/KPP/KPP_exact.py:141-147, Exact_Burger_RV.py:207-213).
template <int FLUX, bool HAVE_FLUXN>
struct CnResidualOp {
  static constexpr int NV = 1;
  static constexpr int MINB = 4;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = CFEM_RES_PREFETCH;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = false;
  const double *uh, *u_n, *eps, *g, *fluxn;
  const uint8_t* is_bc;
  double hdt;  // dt/2
  double* F;
  __device__ void cell(const int32_t* v, const CellGeom& cg, int bcmask, double* evec, double*) const {
    double u[3], un[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { u[k] = uh[v[k]]; un[k] = u_n[v[k]]; }
    const double eb = (eps[v[0]] + eps[v[1]] + eps[v[2]]) * (1.0 / 3.0);
    double f[3];
    flux_vec<FLUX>(cg, u, f);
    if (!HAVE_FLUXN) {
      double fn[3];
      flux_vec<FLUX>(cg, un, fn);
#pragma unroll
      for (int a = 0; a < 3; ++a) f[a] += fn[a];
    }
    const double m = cg.area * (1.0 / 12.0);
    const double d0 = u[0] - un[0], d1 = u[1] - un[1], d2 = u[2] - un[2], sd = d0 + d1 + d2;
    const double dd[3] = {d0, d1, d2};
    const double sx = (u[0] + un[0]) * cg.gx[0] + (u[1] + un[1]) * cg.gx[1] + (u[2] + un[2]) * cg.gx[2];
    const double sy = (u[0] + un[0]) * cg.gy[0] + (u[1] + un[1]) * cg.gy[1] + (u[2] + un[2]) * cg.gy[2];
    const double kf = hdt * eb * cg.area;
#pragma unroll
    for (int a = 0; a < 3; ++a) evec[a] = m * (sd + dd[a]) + hdt * f[a] + kf * (sx * cg.gx[a] + sy * cg.gy[a]);
    if (bcmask) {
      // lifting: F_a += J_ab (g_b - uh_b) over Dirichlet columns b (alpha = -1, x0 = uh)
      double J[9];
      flux_jac<FLUX>(cg, u, J);
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        if (bcmask & (1 << b)) {
          const double dg = g[v[b]] - u[b];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const double Jab = (a == b ? 2.0 * m : m) + hdt * J[a * 3 + b] +
                               kf * (cg.gx[a] * cg.gx[b] + cg.gy[a] * cg.gy[b]);
            evec[a] += Jab * dg;
          }
        }
      }
    }
  }
  __device__ double node(int32_t i, const double* acc) const {
    double f = acc[0];
    if (HAVE_FLUXN) f += hdt * fluxn[i];
    if (is_bc[i]) f = uh[i] - g[i];
    F[i] = f;
    return f * f;
  }
};

template <int FLUX>
struct CnJacobianOp {
  static constexpr int NV = 0;
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = true;
  const double *uh, *eps;
  double hdt;
  __device__ void cell(const int32_t* v, const CellGeom& cg, int, double*, double* emat) const {
    const double u[3] = {uh[v[0]], uh[v[1]], uh[v[2]]};
    const double eb = (eps[v[0]] + eps[v[1]] + eps[v[2]]) * (1.0 / 3.0);
    double J[9];
    flux_jac<FLUX>(cg, u, J);
    const double m = cg.area * (1.0 / 12.0), kf = hdt * eb * cg.area;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        emat[a * 3 + b] = (a == b ? 2.0 * m : m) + hdt * J[a * 3 + b] +
                          kf * (cg.gx[a] * cg.gx[b] + cg.gy[a] * cg.gy[b]);
  }
  __device__ double node(int32_t, const double*) const { return 0.0; }
};

// F(uh) AND J = dF/duh in one cell pass: what the first Newton iteration of a step needs (same uh, u_n, eps, geometry;
// the flux Jacobian the lifting of F uses is the one J is built from).  Saves a launch and a full gather pass per step.
template <int FLUX, bool HAVE_FLUXN>
struct CnResJacOp {
  static constexpr int NV = 1;
  static constexpr int MINB = 1;
  static constexpr bool PREFETCH = true;
  static constexpr bool MAT = true;
  const double *uh, *u_n, *eps, *g, *fluxn;
  const uint8_t* is_bc;
  double hdt;  // dt/2
  double* F;
  __device__ void cell(const int32_t* v, const CellGeom& cg, int bcmask, double* evec, double* emat) const {
    double u[3], un[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { u[k] = uh[v[k]]; un[k] = u_n[v[k]]; }
    const double eb = (eps[v[0]] + eps[v[1]] + eps[v[2]]) * (1.0 / 3.0);
    double f[3];
    flux_vec<FLUX>(cg, u, f);
    if (!HAVE_FLUXN) {
      double fn[3];
      flux_vec<FLUX>(cg, un, fn);
#pragma unroll
      for (int a = 0; a < 3; ++a) f[a] += fn[a];
    }
    double J[9];
    flux_jac<FLUX>(cg, u, J);
    const double m = cg.area * (1.0 / 12.0);
    const double d0 = u[0] - un[0], d1 = u[1] - un[1], d2 = u[2] - un[2], sd = d0 + d1 + d2;
    const double dd[3] = {d0, d1, d2};
    const double sx = (u[0] + un[0]) * cg.gx[0] + (u[1] + un[1]) * cg.gx[1] + (u[2] + un[2]) * cg.gx[2];
    const double sy = (u[0] + un[0]) * cg.gy[0] + (u[1] + un[1]) * cg.gy[1] + (u[2] + un[2]) * cg.gy[2];
    const double kf = hdt * eb * cg.area;
#pragma unroll
    for (int a = 0; a < 3; ++a) evec[a] = m * (sd + dd[a]) + hdt * f[a] + kf * (sx * cg.gx[a] + sy * cg.gy[a]);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        emat[a * 3 + b] = (a == b ? 2.0 * m : m) + hdt * J[a * 3 + b] + kf * (cg.gx[a] * cg.gx[b] + cg.gy[a] * cg.gy[b]);
    if (bcmask) {
      // lifting: F_a += J_ab (g_b - uh_b) over Dirichlet columns b (alpha = -1, x0 = uh); same expression order as
      // CnResidualOp so that the two evaluations of F agree to the bit
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        if (bcmask & (1 << b)) {
          const double dg = g[v[b]] - u[b];
#pragma unroll
          for (int a = 0; a < 3; ++a) evec[a] += emat[a * 3 + b] * dg;
        }
      }
    }
  }
  __device__ double node(int32_t i, const double* acc) const {
    double f = acc[0];
    if (HAVE_FLUXN) f += hdt * fluxn[i];
    if (is_bc[i]) f = uh[i] - g[i];
    F[i] = f;
    return f * f;
  }
};

// Linear advection Crank-Nicolson system (reference RV_node.py:220-242)
struct AdvSystemOp {
  static constexpr int NV = 1;
  static constexpr int MINB = 1;         // resident CTAs per SM the register budget is held to
  static constexpr bool PREFETCH = true;  // phase A requests all of a thread's cells up front
  static constexpr bool MAT = true;
  const double2* w;
  const double *eps, *u_n, *g;  // eps, g nullable
  const uint8_t* is_bc;
  double hdt;
  double* b;
  __device__ void cell(const int32_t* v, const CellGeom& cg, int bcmask, double* evec, double* emat) const {
    const double2 w0 = w[v[0]], w1 = w[v[1]], w2 = w[v[2]];
    const double Wx = w0.x + w1.x + w2.x, Wy = w0.y + w1.y + w2.y;
    const double wx[3] = {Wx + w0.x, Wx + w1.x, Wx + w2.x};
    const double wy[3] = {Wy + w0.y, Wy + w1.y, Wy + w2.y};
    const double m = cg.area * (1.0 / 12.0);
    const double eb = eps ? (eps[v[0]] + eps[v[1]] + eps[v[2]]) * (1.0 / 3.0) : 0.0;
    const double kf = hdt * eb * cg.area;
    const double un[3] = {u_n[v[0]], u_n[v[1]], u_n[v[2]]};
    double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const double Mab = (a == bb) ? 2.0 * m : m;
        const double S = hdt * m * (wx[a] * cg.gx[bb] + wy[a] * cg.gy[bb]) +
                         kf * (cg.gx[a] * cg.gx[bb] + cg.gy[a] * cg.gy[bb]);
        emat[a * 3 + bb] = Mab + S;
        r[a] += (Mab - S) * un[bb];
      }
    }
    if (bcmask && g) {
      // apply_lifting: b -= A[:, bc] g
#pragma unroll
      for (int bb = 0; bb < 3; ++bb)
        if (bcmask & (1 << bb)) {
          const double gv = g[v[bb]];
#pragma unroll
          for (int a = 0; a < 3; ++a) r[a] -= emat[a * 3 + bb] * gv;
        }
    }
    evec[0] = r[0]; evec[1] = r[1]; evec[2] = r[2];
  }
  __device__ double node(int32_t i, const double* acc) const {
    b[i] = is_bc[i] ? (g ? g[i] : 0.0) : acc[0];
    return 0.0;
  }
};

// ------------------------------------------------------------------ the tile kernel
static_assert(kTileCellCap <= 3 * kTileNodes, "phase A visits at most three cells per thread");
template <class Op>
__global__ void __launch_bounds__(kTileNodes, Op::MINB)
k_tile_assemble(const DevMesh m, const Op op, const bool bc, double* __restrict__ vals,
                double* __restrict__ dinv, double* __restrict__ partials, const int ccap, const int nnzcap,
                const Fin fin, double* __restrict__ fin_out) {
  pdl_wait();
  pdl_launch();
  constexpr int NV = Op::NV;
  constexpr bool MAT = Op::MAT;
  extern __shared__ double smem[];
  double* evec = smem;                                   // [NV][ccap][3]
  double* emat = evec + (size_t)NV * 3 * ccap;           // [ccap][9]
  double* rowbuf = emat + (MAT ? (size_t)9 * ccap : 0);  // [nnzcap]
  __shared__ double red[9];
  double local = 0.0;
  const int tid = threadIdx.x;

  for (int tile = blockIdx.x; tile < m.ntiles; tile += gridDim.x) {
    const int n0 = m.tile_node[tile], n1 = m.tile_node[tile + 1];
    const int c0 = m.tile_cellptr[tile], ncl = m.tile_cellptr[tile + 1] - c0;
    const int row0 = m.rowptr[n0], tnnz = m.rowptr[n1] - row0;
    if (MAT)
      for (int p = tid; p < tnnz; p += kTileNodes) rowbuf[p] = 0.0;

    // ---- phase A: cell-local quadrature -> shared memory
    // Phase A visits this thread's cells cl = tid, tid + 256, ... (<= kTileCellCap / kTileNodes = 3).  With
    // PREFETCH their ids and vertices are requested up front, so the dependent tile_cells -> cells -> xy/field
    // chains of the passes overlap; the residual op is also register bound and is held to 64 registers / 4 CTAs
    // per SM.  Measured at 1 M nodes (Burgers): residual 82 -> 57 us, Jacobian 120 -> 108 us, RV rhs 55 -> 54 us.
    int32_t vv[3][3];
    if (Op::PREFETCH) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int cl = tid + j * kTileNodes;
        if (cl < ncl) {
          const int c = m.tile_cells[c0 + cl];
          vv[j][0] = m.cells[3 * (int64_t)c]; vv[j][1] = m.cells[3 * (int64_t)c + 1]; vv[j][2] = m.cells[3 * (int64_t)c + 2];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int cl = tid + j * kTileNodes;
      if (cl >= ncl) break;
      int32_t v[3];
      if (Op::PREFETCH) {
        v[0] = vv[j][0]; v[1] = vv[j][1]; v[2] = vv[j][2];
      } else {
        const int c = m.tile_cells[c0 + cl];
        v[0] = m.cells[3 * (int64_t)c]; v[1] = m.cells[3 * (int64_t)c + 1]; v[2] = m.cells[3 * (int64_t)c + 2];
      }
      const CellGeom g = cell_geom(m.xy[v[0]], m.xy[v[1]], m.xy[v[2]]);
      int bcmask = 0;
      if (bc) bcmask = (int)m.is_bc[v[0]] | ((int)m.is_bc[v[1]] << 1) | ((int)m.is_bc[v[2]] << 2);
      double ev[NV > 0 ? NV * 3 : 1];
      double em[9];
      op.cell(v, g, bcmask, ev, em);
#pragma unroll
      for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int k = 0; k < 3; ++k) evec[((size_t)q * ccap + cl) * 3 + k] = ev[q * 3 + k];
      if (MAT) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const bool kill = (bcmask >> a | bcmask >> b) & 1;
            emat[(size_t)cl * 9 + a * 3 + b] = kill ? 0.0 : em[a * 3 + b];
          }
      }
    }
    __syncthreads();

    // ---- phase B: node-centric gather, fixed order
    const int i = n0 + tid;
    if (i < n1) {
      double acc[NV > 0 ? NV : 1];
#pragma unroll
      for (int q = 0; q < (NV > 0 ? NV : 1); ++q) acc[q] = 0.0;
      const int rbase = m.rowptr[i] - row0;
      int diag = 0;
      const int e1 = m.v2c_ptr[i + 1];
      for (int e = m.v2c_ptr[i]; e < e1; ++e) {
        const uint32_t code = m.v2c_code[e];
        const int cl = code & ((1u << kCodeCellBits) - 1);
        const int k = (code >> kCodeCellBits) & 3;
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] += evec[((size_t)q * ccap + cl) * 3 + k];
        if (MAT) {
          const int p0 = (code >> (kCodeCellBits + 2)) & 31, p1 = (code >> (kCodeCellBits + 7)) & 31,
                    p2 = (code >> (kCodeCellBits + 12)) & 31;
          const double* er = &emat[(size_t)cl * 9 + k * 3];
          rowbuf[rbase + p0] += er[0];
          rowbuf[rbase + p1] += er[1];
          rowbuf[rbase + p2] += er[2];
          diag = k == 0 ? p0 : (k == 1 ? p1 : p2);
        }
      }
      local += op.node(i, acc);
      if (MAT) {
        if (bc && m.is_bc[i]) rowbuf[rbase + diag] = 1.0;
        if (dinv) dinv[i] = 1.0 / rowbuf[rbase + diag];
      }
    }
    __syncthreads();
    if (MAT)
      for (int p = tid; p < tnnz; p += kTileNodes) vals[row0 + p] = rowbuf[p];
    __syncthreads();
  }
  if (partials) {
    const double s = block_sum(local, red);
    if (tid == 0) partials[blockIdx.x] = s;
    if (fin.counter) {   // total (over the CTAs and, distributed, the ranks) finished here: the host reads ONE scalar
      __shared__ double sums[1];
      Slots<1> sl;
      sl.p[0] = partials;
      if (fin_reduce<1>(fin, sl, gridDim.x, red, sums) && tid == 0) fin_out[0] = sums[0];
    }
  }
}

int assembly_grid(const cfem_ctx* c) {
  const int want = c->sm_count * 3;
  return c->dm.ntiles < want ? c->dm.ntiles : want;
}

// Launches one persistent CTA per resident slot (occupancy x SM count), each looping
// over tiles; returns the grid size (== number of per-CTA partials written).
template <class Op>
static int run_tiles(cfem_ctx* c, const Op& op, bool bc, double* vals, double* dinv, double* partials) {
  const int ccap = c->hm.max_tile_cells, nnzcap = c->hm.max_tile_nnz;
  const size_t smem = sizeof(double) * ((size_t)Op::NV * 3 * ccap + (Op::MAT ? (size_t)9 * ccap + nnzcap : 0));
  // occupancy per (instantiation, context): the answer depends on the tile capacities of THIS mesh and on the device
  // of THIS context, so it is cached in the context (keyed by the kernel's address), not in a function static
  int occ = 0;
  {
    const void* key = (const void*)k_tile_assemble<Op>;
    for (auto& e : c->asm_occ)
      if (e.first == key) { occ = e.second; break; }
    if (occ == 0) {
      CUDA_OK(cudaFuncSetAttribute(k_tile_assemble<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tile_assemble<Op>, kTileNodes, smem));
      if (occ < 1) CFEM_THROW(-2, "assembly kernel does not fit on an SM");
      c->asm_occ.emplace_back(key, occ);
    }
  }
  int grid = c->sm_count * occ;
  if (grid > c->dm.ntiles) grid = c->dm.ntiles;
  if (grid > kMaxPartials) grid = kMaxPartials;
  ProfScope ps(c, Op::MAT ? PROF_ASM_MAT : PROF_ASM_VEC);
  // a kernel that produces partials also finishes their sum when in-kernel reductions are available (scalars[24])
  const bool finish = partials != nullptr && fin_available(c);
  launch_pdl(k_tile_assemble<Op>, grid, kTileNodes, smem, c->stream, c->dm, op, bc, vals, dinv, partials, ccap, nnzcap,
             finish ? make_fin(c) : Fin(), c->scalars + 24);
  CUDA_OK(cudaGetLastError());
  c->launches.total++;
  c->launches.assembly++;
  return grid;
}

void launch_mass(cfem_ctx* c, Matrix& M, bool bc) {
  run_tiles(c, MassOp{}, bc, M.vals, M.dinv, nullptr);
  M.valid = true;
}

void launch_stiffness(cfem_ctx* c, Matrix& K, const double* eps) {
  run_tiles(c, StiffnessOp{eps}, false, K.vals, K.dinv, nullptr);
  K.valid = true;
}

void launch_grad_matrix(cfem_ctx* c, int d, Matrix& C) {
  run_tiles(c, GradOp{d}, false, C.vals, C.dinv, nullptr);
  C.valid = true;
}

void launch_mass_stiff(cfem_ctx* c, const double* eps, double coef, Matrix& S) {
  run_tiles(c, MassStiffOp{eps, coef}, false, S.vals, S.dinv, nullptr);
  S.valid = true;
}

void launch_nodal_h_rhs(cfem_ctx* c, double* b) {
  run_tiles(c, NodalHOp{c->dm.xy, b}, false, nullptr, nullptr, nullptr);
}

template <int FLUX>
static void rv_rhs_t(cfem_ctx* c, int scheme, double dt, const double* u_n, const double* u_old,
                     const double* u_oo, const double2* w, bool use_bc, double* b, double* fluxn) {
  double cn, co, coo;
  if (scheme == CFEM_BDF2) { cn = 3.0 / (2.0 * dt); co = -4.0 / (2.0 * dt); coo = 1.0 / (2.0 * dt); }
  else { cn = 1.0 / dt; co = -1.0 / dt; coo = 0.0; u_oo = nullptr; }
  const uint8_t* isbc = use_bc ? c->dm.is_bc : nullptr;
  if (fluxn)
    run_tiles(c, RvRhsOp<FLUX, 2>{u_n, u_old, u_oo, w, isbc, cn, co, coo, b, fluxn}, false, nullptr, nullptr, nullptr);
  else
    run_tiles(c, RvRhsOp<FLUX, 1>{u_n, u_old, u_oo, w, isbc, cn, co, coo, b, nullptr}, false, nullptr, nullptr, nullptr);
}

void launch_rv_rhs(cfem_ctx* c, int flux, int scheme, double dt, const double* u_n, const double* u_old,
                   const double* u_oo, const double2* w, bool use_bc, double* b, double* fluxn) {
  if (scheme == CFEM_BDF2 && !u_oo) CFEM_THROW(-1, "BDF2 residual needs u_old_old");
  switch (flux) {
    case CFEM_FLUX_ADVECTION:
      if (!w) CFEM_THROW(-1, "advection residual needs the velocity field w");
      rv_rhs_t<CFEM_FLUX_ADVECTION>(c, scheme, dt, u_n, u_old, u_oo, w, use_bc, b, fluxn); break;
    case CFEM_FLUX_BURGERS: rv_rhs_t<CFEM_FLUX_BURGERS>(c, scheme, dt, u_n, u_old, u_oo, w, use_bc, b, fluxn); break;
    case CFEM_FLUX_KPP: rv_rhs_t<CFEM_FLUX_KPP>(c, scheme, dt, u_n, u_old, u_oo, w, use_bc, b, fluxn); break;
    default: CFEM_THROW(-1, "unknown flux kind");
  }
}

template <int FLUX>
static int cn_residual_t(cfem_ctx* c, double dt, const double* uh, const double* u_n, const double* eps,
                         const double* g, const double* fluxn, double* F, double* partials) {
  if (fluxn)
    return run_tiles(c, CnResidualOp<FLUX, true>{uh, u_n, eps, g, fluxn, c->dm.is_bc, 0.5 * dt, F}, true, nullptr, nullptr, partials);
  return run_tiles(c, CnResidualOp<FLUX, false>{uh, u_n, eps, g, fluxn, c->dm.is_bc, 0.5 * dt, F}, true, nullptr, nullptr, partials);
}

int launch_cn_residual(cfem_ctx* c, int flux, double dt, const double* uh, const double* u_n,
                       const double* eps, const double* g, const double* fluxn, double* F,
                       double* partials) {
  switch (flux) {
    case CFEM_FLUX_BURGERS: return cn_residual_t<CFEM_FLUX_BURGERS>(c, dt, uh, u_n, eps, g, fluxn, F, partials);
    case CFEM_FLUX_KPP: return cn_residual_t<CFEM_FLUX_KPP>(c, dt, uh, u_n, eps, g, fluxn, F, partials);
    default: CFEM_THROW(-1, "cn_residual: flux must be BURGERS or KPP");
  }
}

template <int FLUX>
static int cn_resjac_t(cfem_ctx* c, double dt, const double* uh, const double* u_n, const double* eps, const double* g,
                       const double* fluxn, double* F, double* partials, Matrix& J) {
  if (fluxn)
    return run_tiles(c, CnResJacOp<FLUX, true>{uh, u_n, eps, g, fluxn, c->dm.is_bc, 0.5 * dt, F}, true, J.vals, J.dinv, partials);
  return run_tiles(c, CnResJacOp<FLUX, false>{uh, u_n, eps, g, fluxn, c->dm.is_bc, 0.5 * dt, F}, true, J.vals, J.dinv, partials);
}

int launch_cn_residual_jacobian(cfem_ctx* c, int flux, double dt, const double* uh, const double* u_n,
                                const double* eps, const double* g, const double* fluxn, double* F, double* partials,
                                Matrix& J) {
  int np = 0;
  switch (flux) {
    case CFEM_FLUX_BURGERS: np = cn_resjac_t<CFEM_FLUX_BURGERS>(c, dt, uh, u_n, eps, g, fluxn, F, partials, J); break;
    case CFEM_FLUX_KPP: np = cn_resjac_t<CFEM_FLUX_KPP>(c, dt, uh, u_n, eps, g, fluxn, F, partials, J); break;
    default: CFEM_THROW(-1, "cn_residual_jacobian: flux must be BURGERS or KPP");
  }
  J.valid = true;
  return np;
}

void launch_cn_jacobian(cfem_ctx* c, int flux, double dt, const double* uh, const double* eps, Matrix& J) {
  switch (flux) {
    case CFEM_FLUX_BURGERS: run_tiles(c, CnJacobianOp<CFEM_FLUX_BURGERS>{uh, eps, 0.5 * dt}, true, J.vals, J.dinv, nullptr); break;
    case CFEM_FLUX_KPP: run_tiles(c, CnJacobianOp<CFEM_FLUX_KPP>{uh, eps, 0.5 * dt}, true, J.vals, J.dinv, nullptr); break;
    default: CFEM_THROW(-1, "cn_jacobian: flux must be BURGERS or KPP");
  }
  J.valid = true;
}

void launch_adv_system(cfem_ctx* c, double dt, const double2* w, const double* eps, const double* u_n,
                       const double* g, Matrix& A, double* b) {
  run_tiles(c, AdvSystemOp{w, eps, u_n, g, c->dm.is_bc, 0.5 * dt, b}, true, A.vals, A.dinv, nullptr);
  A.valid = true;
}

}  // namespace cfem
