// cont_tile.cuh -- tiled, write-once ContinuityAdvElemKernel assembly (one dof per node) for structured
// Hex8 blocks.  What is computed is what AssembleElemSolverAlgorithm::execute +
// ContinuityAdvElemKernel::execute + TpetraLinearSystem::sumInto produce
// (src/AssembleElemSolverAlgorithm.C:74-100, src/kernel/ContinuityAdvElemKernel.C:86-148,
// src/TpetraLinearSystem.C:1457-1522); every CSR value and residual entry is written to HBM exactly once,
// already summed over the <= 8 elements that touch it -- no atomics, no zeroSystem, bit-reproducible.
//
// Work decomposition (different from the momentum kernel: a row is 27 doubles, so one thread owns it):
//   * a CTA (128 threads) owns a 15 x 7 column of NODES and sweeps z one ELEMENT LAYER at a time; the layer
//     of the tile plus its halo is 16 x 8 = 128 elements, one per thread in phase 1;
//   * phase 1, thread = element: the 12 SCS integration points of the element.  Per point a 4-double
//     record in shared memory: b_k = -(sum_i adj(J)[i][k] A_i)/det J (k = parametric axis) at the location
//     of the lhs grad-op, so that the reference's lhsfac(ic) = -dndx_lhs(ic,:).A is dN_ic/dxi . b (three
//     FMAs with compile-time dN/dxi in phase 2), and the residual flux s = mdot/projTimeScale with
//     mdot = (rho u + tau Gp)_ip . A - tau dpdx_ip . A (ContinuityAdvElemKernel.C:135-145; rho u + tau Gp
//     is interpolated as one nodal vector -- the shape-function interpolation is linear);
//   * phase 2, thread = node: the 4 elements of the layer around the node feed TWO rows held in registers,
//     the row of the node below the layer (its upper octants; complete after this layer -> written out)
//     and the row of the node above it (its lower octants; completed by the next layer).  All indices are
//     compile-time: 576 FMAs per thread and layer on 54 accumulators;
//   * nodal fields are staged with cp.async two planes ahead; records of elements outside the block are
//     zero, so phase 2 carries no boundary tests; absent neighbours are skipped through the slot map
//     (node_slot27 == 0xFF).
// Every phase is a plain function of the thread index so that tests/cpu_harness executes the kernel thread by
// thread on the CPU against the oracle.
#pragma once
#include "box_tile.cuh"

namespace nalu_b200 {
namespace cont {

using box::bit_x;
using box::bit_y;
using box::bit_z;
using box::ip_of;
using box::node_of_bits;

struct ContParams {
  int nx, ny, nz;    // elements of the local block
  int NX, NY, NZ;    // nodes
  int z_chunk, tiles_x, tiles_y;
  int z_first, z_last; // node planes [z_first, z_last) of this launch
  const double *coords, *vel, *rho, *p, *gp; // [nNodes][3], [3], [1], [1], [3]
  const int *node_row;
  const long long *browptr;
  const unsigned char *slot27;
  double *values;
  double *rhs;
  int *err;
  double tau;      // projTimeScale = dt / gamma1 (ContinuityAdvElemKernel.C:77-84)
  int options;     // NALU_OPT_SHIFT_MDOT | SHIFT_POISSON | REDUCED_SENS
  int overwrite;
};

struct Cfg {
  static constexpr int TBX = 15, TBY = 7;                 // nodes of the tile
  static constexpr int ELX = 16, ELY = 8, NEL = ELX * ELY; // elements of one layer incl. the halo
  static constexpr int NT = 128;
  static constexpr int RNX = ELX + 1, RNY = ELY + 1, RN = RNX * RNY; // nodes of those elements
  static constexpr int NF = 11;                           // x[3], p, u[3], rho, Gp[3]
  static constexpr int PLN = RN * NF;
  static constexpr int ES = 12 * 4 + 2;                   // element stride: 400 bytes = 16 mod 128
  static constexpr int OFF_REC = 0, OFF_RING = NEL * ES;
  static constexpr int SMEM_DOUBLES = OFF_RING + 3 * PLN;
};

// ---- staging of one node plane of the tile's neighbourhood (global -> shared, cp.async) -----------------
__host__ __device__ __forceinline__ void copy8_or_zero(double *dst, const double *src, bool valid)
{
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(valid ? 8 : 0)
               : "memory");
#else
  *dst = valid ? *src : 0.0;
#endif
}
__host__ __device__ __forceinline__ void stage_wait()
{
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
__host__ __device__ inline void stage_plane(const ContParams &P, double *ring_plane, int tid, int x0, int y0, int zp)
{
  using C = Cfg;
  const long long pn = (long long)P.NX * P.NY;
  const bool zin = zp >= 0 && zp < P.NZ;
  // one node per thread and round: the index arithmetic is paid once per node, not once per double
#pragma unroll 1
  for (int rn = tid; rn < C::RN; rn += C::NT) {
    const int x = x0 - 1 + rn % C::RNX, y = y0 - 1 + rn / C::RNX;
    const bool valid = zin && x >= 0 && y >= 0 && x < P.NX && y < P.NY;
    const long long n = valid ? x + (long long)P.NX * y + pn * zp : 0;
    double *d = ring_plane + rn * C::NF;
    const double *xc = P.coords + n * 3, *u = P.vel + n * 3, *g = P.gp + n * 3;
    copy8_or_zero(d + 0, xc + 0, valid);
    copy8_or_zero(d + 1, xc + 1, valid);
    copy8_or_zero(d + 2, xc + 2, valid);
    copy8_or_zero(d + 3, P.p + n, valid);
    copy8_or_zero(d + 4, u + 0, valid);
    copy8_or_zero(d + 5, u + 1, valid);
    copy8_or_zero(d + 6, u + 2, valid);
    copy8_or_zero(d + 7, P.rho + n, valid);
    copy8_or_zero(d + 8, g + 0, valid);
    copy8_or_zero(d + 9, g + 1, valid);
    copy8_or_zero(d + 10, g + 2, valid);
  }
}

// ---- phase 1: thread = element el of layer ez ---------------------------------------------------------------
// b at parametric location (s1,s2,s3): b[k] = (sum_i adj[i][k] A[i]) / det, so that dndx(n,:).A = dN_n/dxi . b
// (generic_grad_op, include/master_element/MasterElementFunctions.h:51-108).  Returns det.
__host__ __device__ __forceinline__ double grad_dot_area(const double (&c)[8][3], double s1, double s2, double s3,
                                                         const double (&A)[3], double (&b)[3])
{
  double j[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    double d[3];
    hex8::shape_deriv(n, s1, s2, s3, d);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) j[i][k] += d[k] * c[n][i];
  }
  double adj[3][3];
  adj[0][0] = j[1][1] * j[2][2] - j[2][1] * j[1][2];
  adj[0][1] = j[1][2] * j[2][0] - j[2][2] * j[1][0];
  adj[0][2] = j[1][0] * j[2][1] - j[2][0] * j[1][1];
  adj[1][0] = j[0][2] * j[2][1] - j[2][2] * j[0][1];
  adj[1][1] = j[0][0] * j[2][2] - j[2][0] * j[0][2];
  adj[1][2] = j[0][1] * j[2][0] - j[2][1] * j[0][0];
  adj[2][0] = j[0][1] * j[1][2] - j[1][1] * j[0][2];
  adj[2][1] = j[0][2] * j[1][0] - j[1][2] * j[0][0];
  adj[2][2] = j[0][0] * j[1][1] - j[1][0] * j[0][1];
  const double det = j[0][0] * adj[0][0] + j[1][0] * adj[1][0] + j[2][0] * adj[2][0];
  const double inv = 1.0 / det;
#pragma unroll
  for (int k = 0; k < 3; ++k) b[k] = (adj[0][k] * A[0] + adj[1][k] * A[1] + adj[2][k] * A[2]) * inv;
  return det;
}

__host__ __device__ inline void p1_element(const ContParams &P, double *sm, int el, int x0, int y0, int ez, const double *pl0,
                                           const double *pl1)
{
  using C = Cfg;
  const int elx = el % C::ELX, ely = el / C::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  double *rec = sm + C::OFF_REC + el * C::ES;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny || ez < 0 || ez >= P.nz) {
#pragma unroll
    for (int i = 0; i < 48; ++i) rec[i] = 0.0;
    return;
  }
  double c[8][3], pn[8], w[8][3];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const double *nd = (bit_z(n) ? pl1 : pl0) + ((elx + bit_x(n)) + C::RNX * (ely + bit_y(n))) * C::NF;
    c[n][0] = nd[0];
    c[n][1] = nd[1];
    c[n][2] = nd[2];
    pn[n] = nd[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) w[n][j] = nd[7] * nd[4 + j] + P.tau * nd[8 + j]; // rho u + tau Gp
  }
  const bool sp = (P.options & NALU_OPT_SHIFT_POISSON) != 0;
  const bool sl = sp || (P.options & NALU_OPT_REDUCED_SENS) != 0; // v_dndx_lhs, ContinuityAdvElemKernel.C:107-108
  const double ql = sl ? 0.5 : 0.25, qr = sp ? 0.5 : 0.25, qm = (P.options & NALU_OPT_SHIFT_MDOT) ? 0.5 : 0.25;
#pragma unroll 1
  for (int ip = 0; ip < 12; ++ip) {
    double A[3], bl[3], br[3], s1, s2, s3;
    hex8::scs_areav_ip(c, ip, A);
    hex8::scs_intg_loc(ip, ql, s1, s2, s3);
    const double det = grad_dot_area(c, s1, s2, s3, A, bl);
    if (!(det > 0.0)) *P.err = 1; // MasterElementFunctions.h:91-94
    if (sl != sp) {
      hex8::scs_intg_loc(ip, qr, s1, s2, s3);
      const double det2 = grad_dot_area(c, s1, s2, s3, A, br);
      if (!(det2 > 0.0)) *P.err = 1;
    } else {
      br[0] = bl[0];
      br[1] = bl[1];
      br[2] = bl[2];
    }
    // dpdx_ip . A with the rhs grad-op: sum_k (dp/dxi_k) br[k]
    hex8::scs_intg_loc(ip, qr, s1, s2, s3);
    double dp[3] = {0, 0, 0};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      double d[3];
      hex8::shape_deriv(n, s1, s2, s3, d);
      dp[0] += d[0] * pn[n];
      dp[1] += d[1] * pn[n];
      dp[2] += d[2] * pn[n];
    }
    const double dpdxA = dp[0] * br[0] + dp[1] * br[1] + dp[2] * br[2];
    hex8::scs_intg_loc(ip, qm, s1, s2, s3);
    double wip[3] = {0, 0, 0};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const double r = hex8::shape_fcn(n, s1, s2, s3);
      wip[0] += r * w[n][0];
      wip[1] += r * w[n][1];
      wip[2] += r * w[n][2];
    }
    const double mdot = wip[0] * A[0] + wip[1] * A[1] + wip[2] * A[2] - P.tau * dpdxA;
    rec[ip * 4 + 0] = -bl[0];
    rec[ip * 4 + 1] = -bl[1];
    rec[ip * 4 + 2] = -bl[2];
    rec[ip * 4 + 3] = mdot / P.tau;
  }
}

// ---- coefficient form of phase 1 ---------------------------------------------------------------------------------
// A trilinear field is f(xi) = sum_m F[m] prod_{a in m} xi_a / 2^(3-|m|) with the Walsh-Hadamard coefficients
// F[m] = sum_n (prod_{a in m} sigma_na) f_n (m: bit 0 = x, 1 = y, 2 = z).  At the SCS point of the edge along t
// (local axes (t,u,v) cyclic) on the sides (su, sv): xi_t = 0, xi_u = su q, xi_v = sv q, hence
//   f = F0/8 + Fu xi_u/4 + Fv xi_v/4 + Fuv xi_u xi_v/2
//   df/dxi_t = Ft/4 + Ftu xi_u/2 + Ftv xi_v/2 + Ftuv xi_u xi_v,  df/dxi_u = Fu/4 + Fuv xi_v/2,  df/dxi_v = Fv/4 + Fuv xi_u/2
// and the facet's area vector along +t is Xu x Xv/64 + Xuv x (sv Xv - su Xu)/128: the vector area of the closed
// quadrilateral (edge midpoint, two face centroids, element centroid) the reference triangulates
// (Hex8CVFEM.C:547-581, Hex8GeometryFunctions.h:30-76) -- 5 vector FMAs for the Jacobian and one cross product for
// the area instead of four 8-node interpolations and a 72-FMA Jacobian per point.
struct Coef {
  double X[3][8]; // coordinates
};
__host__ __device__ __forceinline__ void gather_bits(const double *pl0, const double *pl1, int n000, int f, double (&g)[8])
{
  using C = Cfg;
#pragma unroll
  for (int m = 0; m < 8; ++m)
    g[m] = ((m & 4) ? pl1 : pl0)[n000 + ((m & 1) + C::RNX * ((m >> 1) & 1)) * C::NF + f];
}
__host__ __device__ __forceinline__ void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3])
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
// geometry of the point (T, SU, SV): area vector along +t, and at location q the vector b~ with
// b~[k] = (d xi_k / d x) . A (k = 0 t, 1 u, 2 v).  Returns det J.
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ void area_plus_t(const double (&X)[3][8], double (&A)[3])
{
  constexpr int t = 1 << T, u = 1 << ((T + 1) % 3), v = 1 << ((T + 2) % 3);
  (void)t;
  constexpr double su = SU ? 1.0 : -1.0, sv = SV ? 1.0 : -1.0;
  double xu[3], xv[3], xuv[3], w[3], c1[3], c2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    xu[i] = X[i][u];
    xv[i] = X[i][v];
    xuv[i] = X[i][u | v];
    w[i] = sv * xv[i] - su * xu[i];
  }
  cross3(xu, xv, c1);
  cross3(xuv, w, c2);
#pragma unroll
  for (int i = 0; i < 3; ++i) A[i] = c1[i] * (1.0 / 64.0) + c2[i] * (1.0 / 128.0);
}
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ double grad_dot_area_c(const double (&X)[3][8], double q, const double (&A)[3], double (&b)[3])
{
  constexpr int t = 1 << T, u = 1 << ((T + 1) % 3), v = 1 << ((T + 2) % 3);
  const double xu_ = (SU ? q : -q), xv_ = (SV ? q : -q);
  double ct[3], cu[3], cv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    ct[i] = 0.25 * X[i][t] + (0.5 * xu_) * X[i][t | u] + (0.5 * xv_) * X[i][t | v] + (xu_ * xv_) * X[i][t | u | v];
    cu[i] = 0.25 * X[i][u] + (0.5 * xv_) * X[i][u | v];
    cv[i] = 0.25 * X[i][v] + (0.5 * xu_) * X[i][u | v];
  }
  double n_t[3], n_u[3], n_v[3];
  cross3(cu, cv, n_t); // det * grad xi_t
  cross3(cv, ct, n_u);
  cross3(ct, cu, n_v);
  const double det = ct[0] * n_t[0] + ct[1] * n_t[1] + ct[2] * n_t[2];
  const double inv = 1.0 / det;
  b[0] = (n_t[0] * A[0] + n_t[1] * A[1] + n_t[2] * A[2]) * inv;
  b[1] = (n_u[0] * A[0] + n_u[1] * A[1] + n_u[2] * A[2]) * inv;
  b[2] = (n_v[0] * A[0] + n_v[1] * A[1] + n_v[2] * A[2]) * inv;
  return det;
}
// value and parametric gradient (t,u,v order) of a field given by its coefficients, at location q of the point
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ double value_c(const double (&F)[8], double q)
{
  constexpr int u = 1 << ((T + 1) % 3), v = 1 << ((T + 2) % 3);
  const double xu_ = (SU ? q : -q), xv_ = (SV ? q : -q);
  return 0.125 * F[0] + (0.25 * xu_) * F[u] + (0.25 * xv_) * F[v] + (0.5 * xu_ * xv_) * F[u | v];
}
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ void grad_c(const double (&F)[8], double q, double (&d)[3])
{
  constexpr int t = 1 << T, u = 1 << ((T + 1) % 3), v = 1 << ((T + 2) % 3);
  const double xu_ = (SU ? q : -q), xv_ = (SV ? q : -q);
  d[0] = 0.25 * F[t] + (0.5 * xu_) * F[t | u] + (0.5 * xv_) * F[t | v] + (xu_ * xv_) * F[t | u | v];
  d[1] = 0.25 * F[u] + (0.5 * xv_) * F[u | v];
  d[2] = 0.25 * F[v] + (0.5 * xu_) * F[u | v];
}
// the integration point (its index, and the sign that turns the +t area vector into the reference's left -> right one)
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ int point_index(double &orient)
{
  constexpr int U = (T + 1) % 3, V = (T + 2) % 3;
  int bits[3];
  bits[T] = 0;
  bits[U] = SU;
  bits[V] = SV;
  const int k0 = node_of_bits(bits[0], bits[1], bits[2]); // the node of the edge with bit_t = 0
  const int ip = ip_of(k0, T);
  orient = hex8::scs_left(ip) == k0 ? 1.0 : -1.0;
  return ip;
}

struct ContElem {
  double X[3][8], Pc[8], W[3][8];
  double ql, qr, qm, tau, inv_tau;
  bool two; // the rhs grad-op sits at another location than the lhs one
};
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ void p1_point_c(const ContParams &P, const ContElem &E, double *rec)
{
  constexpr int U = (T + 1) % 3, V = (T + 2) % 3;
  double orient;
  const int ip = point_index<T, SU, SV>(orient);
  double A[3], bl[3], br[3];
  area_plus_t<T, SU, SV>(E.X, A);
  A[0] *= orient;
  A[1] *= orient;
  A[2] *= orient;
  const double det = grad_dot_area_c<T, SU, SV>(E.X, E.ql, A, bl);
  if (!(det > 0.0)) *P.err = 1; // MasterElementFunctions.h:91-94
  if (E.two) {
    const double det2 = grad_dot_area_c<T, SU, SV>(E.X, E.qr, A, br);
    if (!(det2 > 0.0)) *P.err = 1;
  } else {
    br[0] = bl[0];
    br[1] = bl[1];
    br[2] = bl[2];
  }
  double dp[3];
  grad_c<T, SU, SV>(E.Pc, E.qr, dp);
  const double dpdxA = dp[0] * br[0] + dp[1] * br[1] + dp[2] * br[2];
  const double w0 = value_c<T, SU, SV>(E.W[0], E.qm), w1 = value_c<T, SU, SV>(E.W[1], E.qm), w2 = value_c<T, SU, SV>(E.W[2], E.qm);
  const double mdot = w0 * A[0] + w1 * A[1] + w2 * A[2] - E.tau * dpdxA;
  rec[ip * 4 + T] = -bl[0];
  rec[ip * 4 + U] = -bl[1];
  rec[ip * 4 + V] = -bl[2];
  rec[ip * 4 + 3] = mdot * E.inv_tau; // mdot / projTimeScale (one division per element instead of twelve)
}
template <int T>
__host__ __device__ __forceinline__ void p1_edge_dir_c(const ContParams &P, const ContElem &E, double *rec)
{
  p1_point_c<T, 0, 0>(P, E, rec);
  p1_point_c<T, 1, 0>(P, E, rec);
  p1_point_c<T, 0, 1>(P, E, rec);
  p1_point_c<T, 1, 1>(P, E, rec);
}
__host__ __device__ inline void p1_element_c(const ContParams &P, double *sm, int el, int x0, int y0, int ez, const double *pl0,
                                             const double *pl1)
{
  using C = Cfg;
  const int elx = el % C::ELX, ely = el / C::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  double *rec = sm + C::OFF_REC + el * C::ES;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny || ez < 0 || ez >= P.nz) {
#pragma unroll
    for (int i = 0; i < 48; ++i) rec[i] = 0.0;
    return;
  }
  const int n000 = (elx + C::RNX * ely) * C::NF;
  ContElem E;
  double u[3][8], rho[8], gp[3][8];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gather_bits(pl0, pl1, n000, i, E.X[i]);
    gather_bits(pl0, pl1, n000, 4 + i, u[i]);
    gather_bits(pl0, pl1, n000, 8 + i, gp[i]);
  }
  gather_bits(pl0, pl1, n000, 3, E.Pc);
  gather_bits(pl0, pl1, n000, 7, rho);
  // coordinates relative to the element's first node: the coefficients used are differences, and differences of
  // partial sums of large coordinates would cost |x|/h ulps
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double o = E.X[i][0];
#pragma unroll
    for (int m = 0; m < 8; ++m) E.X[i][m] -= o;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int m = 0; m < 8; ++m) E.W[i][m] = rho[m] * u[i][m] + P.tau * gp[i][m]; // rho u + tau Gp
    box::butterfly8(E.X[i]);
    box::butterfly8(E.W[i]);
  }
  box::butterfly8(E.Pc);
  const bool sp = (P.options & NALU_OPT_SHIFT_POISSON) != 0;
  const bool sl = sp || (P.options & NALU_OPT_REDUCED_SENS) != 0; // v_dndx_lhs, ContinuityAdvElemKernel.C:107-108
  E.ql = sl ? 0.5 : 0.25;
  E.qr = sp ? 0.5 : 0.25;
  E.qm = (P.options & NALU_OPT_SHIFT_MDOT) ? 0.5 : 0.25;
  E.tau = P.tau;
  E.inv_tau = 1.0 / P.tau;
  E.two = sl != sp;
  p1_edge_dir_c<0>(P, E, rec);
  p1_edge_dir_c<1>(P, E, rec);
  p1_edge_dir_c<2>(P, E, rec);
}

// ---- phase 2: thread = node (lx, ly) of the tile -----------------------------------------------------------
// accA / rA: row of the node in plane ez (below the layer), accB / rB: row of the node in plane ez+1.
// acc index = (dz+1)*9 + (dy+1)*3 + (dx+1) of the column node relative to the row node (the order of
// node_slot27).  SH: the lhs grad-op is the shifted one (integration points at the edge midpoints).
template <bool SH, int SX, int SY, int DIR, int LVL>
__host__ __device__ __forceinline__ void p2_point(const double *er, double (&acc)[27], double &r)
{
  // the row node is corner (!SX, !SY, LVL) of the element that lies on the (SX, SY) side of it
  constexpr int bx = !SX, by = !SY, bz = LVL;
  const int k = node_of_bits(bx, by, bz);
  const int ip = ip_of(k, DIR);
  const double sgn = hex8::scs_left(ip) == k ? 1.0 : -1.0; // lhs(il,ic) += lhsfac, lhs(ir,ic) -= lhsfac
  double s1, s2, s3;
  hex8::scs_intg_loc(ip, SH ? 0.5 : 0.25, s1, s2, s3);
  const double b0 = er[ip * 4 + 0], b1 = er[ip * 4 + 1], b2 = er[ip * 4 + 2], s = er[ip * 4 + 3];
#pragma unroll
  for (int ic = 0; ic < 8; ++ic) {
    double d[3];
    hex8::shape_deriv(ic, s1, s2, s3, d);
    const int idx = (bit_z(ic) - bz + 1) * 9 + (bit_y(ic) - by + 1) * 3 + (bit_x(ic) - bx + 1);
    acc[idx] += (sgn * d[0]) * b0;
    acc[idx] += (sgn * d[1]) * b1;
    acc[idx] += (sgn * d[2]) * b2;
  }
  r -= sgn * s;
}
template <bool SH, int SX, int SY>
__host__ __device__ __forceinline__ void p2_element(const double *sm, int lx, int ly, double (&accA)[27], double &rA,
                                                    double (&accB)[27], double &rB)
{
  using C = Cfg;
  const double *er = sm + C::OFF_REC + ((lx + SX) + C::ELX * (ly + SY)) * C::ES;
  p2_point<SH, SX, SY, 0, 0>(er, accA, rA);
  p2_point<SH, SX, SY, 1, 0>(er, accA, rA);
  p2_point<SH, SX, SY, 2, 0>(er, accA, rA);
  p2_point<SH, SX, SY, 0, 1>(er, accB, rB);
  p2_point<SH, SX, SY, 1, 1>(er, accB, rB);
  p2_point<SH, SX, SY, 2, 1>(er, accB, rB);
}
template <bool SH>
__host__ __device__ __forceinline__ void p2_layer(const double *sm, int lx, int ly, double (&accA)[27], double &rA,
                                                  double (&accB)[27], double &rB)
{
  p2_element<SH, 0, 0>(sm, lx, ly, accA, rA, accB, rB);
  p2_element<SH, 1, 0>(sm, lx, ly, accA, rA, accB, rB);
  p2_element<SH, 0, 1>(sm, lx, ly, accA, rA, accB, rB);
  p2_element<SH, 1, 1>(sm, lx, ly, accA, rA, accB, rB);
}

// ---- the finished row of node (x, y, z) -> CSR ---------------------------------------------------------------
template <class Params>
__host__ __device__ inline void write_row(const Params &P, int x, int y, int z, const double (&acc)[27], double r)
{
  const long long n = x + (long long)P.NX * (y + (long long)P.NY * z);
  const int row = P.node_row[n];
  const long long b0 = P.browptr[row];
  const int nb = (int)(P.browptr[row + 1] - b0);
  const unsigned char *sl = P.slot27 + n * 27;
  unsigned s[27];
  int nvalid = 0;
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    s[k] = sl[k];
    nvalid += s[k] != 0xFF;
  }
  double *row_v = P.values + b0;
  if (P.overwrite) {
    // columns no local element touches (remote columns announced by the ranks sharing the node) read zero
    if (nvalid != nb)
      for (int q = 0; q < nb; ++q) row_v[q] = 0.0;
#pragma unroll
    for (int k = 0; k < 27; ++k)
      if (s[k] != 0xFF) row_v[s[k]] = acc[k];
    P.rhs[row] = r;
  } else {
#pragma unroll
    for (int k = 0; k < 27; ++k)
      if (s[k] != 0xFF) row_v[s[k]] += acc[k];
    P.rhs[row] += r;
  }
}

} // namespace cont
} // namespace nalu_b200
