// slist_tile.cuh -- tiled, write-once assembly of the scalar transport kernel LISTS on structured Hex8 blocks (one dof
// per node), e.g. the turbulence equations of BASELINE config 4:
//   tke: [ScalarMassElemKernel(lumped), ScalarUpwAdvDiffElemKernel, TurbKineticEnergySSTSrcElemKernel]
//   sdr: [ScalarMassElemKernel(lumped), ScalarUpwAdvDiffElemKernel, SpecificDissipationRateSSTSrcElemKernel]
// (src/TurbKineticEnergyEquationSystem.C:447-494, src/SpecificDissipationRateEquationSystem.C:334-361), evaluated in
// ONE launch: what AssembleElemSolverAlgorithm::execute sums into one scratch lhs/rhs per element
// (src/AssembleElemSolverAlgorithm.C:86-91) and TpetraLinearSystem::sumInto scatters (src/TpetraLinearSystem.C:1457-1522),
// every CSR value written exactly once, no atomics, no zeroSystem.  Kernels followed:
//   src/kernel/ScalarUpwAdvDiffElemKernel.C:110-256 (or ScalarAdvDiffElemKernel.C:69-127 / ScalarDiffElemKernel.C:64-110),
//   src/kernel/ScalarMassElemKernel.C:117-154 (lumped), src/kernel/TurbKineticEnergySSTSrcElemKernel.C:95-174,
//   src/kernel/SpecificDissipationRateSSTSrcElemKernel.C:100-204.
//
// Same decomposition as scalar_tile.cuh (a CTA owns a 15 x 7 column of nodes and sweeps z one element layer at a time;
// thread = element in phase 1, thread = node row with 54 register accumulators in phase 2), with two record passes per
// layer that share one shared-memory buffer:
//   SCV pass: per sub-control volume (= element corner) V (Grandy volume), f (source lhs factor x V), r (source rhs x V)
//             -> phase 2b adds them to the rows of the corner nodes and gathers the dual volume for the lumped mass;
//   SCS pass: per integration point b_k = -Gamma_ip (d xi_k/dx . A), mdot, the Peclet blending factor and the residual
//             flux -> phase 2a; the upwind operator's factors (ScalarUpwAdvDiffElemKernel.C:208-223) follow from
//             (mdot, pecfac) with three flops, so the record stays at 6 doubles.
// Phase 1 works on the Walsh-Hadamard coefficients of the trilinear fields (cont_tile.cuh): a parametric gradient is 3
// FMAs per axis, the Jacobian at a sub-control-volume point 27 FMAs.
// Every phase is a plain function of the thread index: tests/cpu_harness/slist_host_check.cpp runs the kernel thread by
// thread on the CPU against the oracle.
#pragma once
#include "box5_tile.cuh"
#include "scalar_tile.cuh"

namespace nalu_b200 {
namespace slist {

using box::bit_x;
using box::bit_y;
using box::bit_z;
using box::ip_of;
using box::node_of_bits;
using cont::copy8_or_zero;

struct SlistParams {
  int nx, ny, nz; // elements of the local block
  int NX, NY, NZ; // nodes
  int z_chunk, tiles_x, tiles_y;
  int z_first, z_last;
  // staged nodal fields: coordinates [3], scalarQ, diffFluxCoeff, velocity (rtm) [3], density, Gjq [3], the "other"
  // turbulence scalar (sdr for the tke list, tke for the sdr list), turbulent_viscosity, sst_f_one_blending
  const double *coords, *q, *gam, *vel, *rho, *gjq, *other, *tvisc, *f1;
  const double *mdot; // [nElem][12] mass_flow_rate_scs; null: ScalarDiffElemKernel
  // lumped ScalarMassElemKernel: old states, read once per row node
  const double *q_n, *q_nm1, *rho_n, *rho_nm1;
  const int *node_row;
  const long long *browptr;
  const unsigned char *slot27;
  double *values;
  double *rhs;
  int *err;
  int options;     // NALU_OPT_SHIFTED_GRAD_OP | NALU_OPT_SKEW_SYMMETRIC of the adv-diff kernel
  int upw;         // 1: ScalarUpwAdvDiffElemKernel; 0: ScalarAdvDiffElemKernel / ScalarDiffElemKernel
  int upw_options; // NALU_OPT_USE_LIMITER | NALU_OPT_PECLET_TANH
  double alpha, alpha_upw, ho_upwind, hybrid, tanh_trans, tanh_width;
  int mass;        // fused lumped ScalarMassElemKernel
  double g1_dt, g2_dt, g3_dt; // gamma / dt
  int src;         // 0: none, 1: TurbKineticEnergySSTSrcElemKernel, 2: SpecificDissipationRateSSTSrcElemKernel
  int src_lumped;  // the source kernel's lumpedMass flag (shifted SCV shape function)
  int src_shifted; // its shifted SCV grad-op
  double turb[10]; // NALU_TM_*
  int overwrite;
};

struct Cfg {
  static constexpr int TBX = 15, TBY = 7;
  static constexpr int ELX = 16, ELY = 8, NEL = ELX * ELY;
  static constexpr int NT = 128;
  static constexpr int RNX = ELX + 1, RNY = ELY + 1, RN = RNX * RNY;
  // x[3], q, Gamma, u[3], rho, Gjq[3], other, tvisc, f1: an odd stride, so the elements of a warp (consecutive nodes)
  // read their corner nodes conflict-free
  static constexpr int NF = 15;
  enum { FX = 0, FQ = 3, FG = 4, FU = 5, FR = 8, FGQ = 9, FO = 12, FT = 13, FF = 14 };
  static constexpr int PLN = RN * NF;
  static constexpr int RS = 6;  // SCS record: b[3], mdot, pecfac, s
  static constexpr int RV = 3;  // SCV record: V, f, r
  static constexpr int ES = 12 * RS + 1; // odd element stride: the rows of a warp read their records conflict-free
  static constexpr int OFF_REC = 0, OFF_RING = NEL * ES;
  static constexpr int SMEM_DOUBLES = OFF_RING + 2 * PLN; // two node planes: plane ez+2 replaces plane ez during phase 2a
};
static_assert(8 * Cfg::RV <= 12 * Cfg::RS, "the SCV records reuse the SCS record buffer");

__host__ __device__ inline void stage_plane(const SlistParams &P, double *ring_plane, int tid, int x0, int y0, int zp)
{
  using C = Cfg;
  const long long pn = (long long)P.NX * P.NY;
  const bool zin = zp >= 0 && zp < P.NZ;
#pragma unroll 1
  for (int rn = tid; rn < C::RN; rn += C::NT) {
    const int x = x0 - 1 + rn % C::RNX, y = y0 - 1 + rn / C::RNX;
    const bool valid = zin && x >= 0 && y >= 0 && x < P.NX && y < P.NY;
    const long long n = valid ? x + (long long)P.NX * y + pn * zp : 0;
    double *d = ring_plane + rn * C::NF;
    const double *xc = P.coords + n * 3;
    copy8_or_zero(d + C::FX + 0, xc + 0, valid);
    copy8_or_zero(d + C::FX + 1, xc + 1, valid);
    copy8_or_zero(d + C::FX + 2, xc + 2, valid);
    copy8_or_zero(d + C::FQ, P.q + n, valid);
    copy8_or_zero(d + C::FG, P.gam + n, valid);
    if (P.vel) {
      const double *u = P.vel + n * 3;
      copy8_or_zero(d + C::FU + 0, u + 0, valid);
      copy8_or_zero(d + C::FU + 1, u + 1, valid);
      copy8_or_zero(d + C::FU + 2, u + 2, valid);
    }
    if (P.rho) copy8_or_zero(d + C::FR, P.rho + n, valid);
    if (P.gjq) {
      const double *g = P.gjq + n * 3;
      copy8_or_zero(d + C::FGQ + 0, g + 0, valid);
      copy8_or_zero(d + C::FGQ + 1, g + 1, valid);
      copy8_or_zero(d + C::FGQ + 2, g + 2, valid);
    }
    if (P.other) copy8_or_zero(d + C::FO, P.other + n, valid);
    if (P.tvisc) copy8_or_zero(d + C::FT, P.tvisc + n, valid);
    if (P.f1) copy8_or_zero(d + C::FF, P.f1 + n, valid);
  }
}

// staged node `n` (Hex8 numbering) of the element whose first node sits at n000
__host__ __device__ __forceinline__ const double *node_ptr(const double *pl0, const double *pl1, int n000, int n)
{
  return (bit_z(n) ? pl1 : pl0) + n000 + (bit_x(n) + Cfg::RNX * bit_y(n)) * Cfg::NF;
}
// nodal values in bit order (bit 0 = x, 1 = y, 2 = z) -> Walsh-Hadamard coefficients
__host__ __device__ __forceinline__ void coef_of(const double *pl0, const double *pl1, int n000, int f, double (&F)[8])
{
  cont::gather_bits(pl0, pl1, n000, f, F); // cont::Cfg::RNX == Cfg::RNX, but its NF differs: gather here
}
__host__ __device__ __forceinline__ void gather_bits(const double *pl0, const double *pl1, int n000, int f, double (&g)[8])
{
  using C = Cfg;
#pragma unroll
  for (int m = 0; m < 8; ++m) g[m] = ((m & 4) ? pl1 : pl0)[n000 + ((m & 1) + C::RNX * ((m >> 1) & 1)) * C::NF + f];
}

// ---- ScalarUpwAdvDiffElemKernel helpers ---------------------------------------------------------------------------------
// ClassicPecletFunction (A = 5) / TanhFunction, src/PecletFunction.C:67-71,98-101
__host__ __device__ __forceinline__ double peclet_function(const SlistParams &P, double x)
{
  if (P.upw_options & NALU_OPT_PECLET_TANH) return 0.50 * (1.0 + tanh((x - P.tanh_trans) / P.tanh_width));
  const double m = P.hybrid * x;
  return m * m / (5.0 + m * m);
}
// ScalarUpwAdvDiffElemKernel.C:258-266
__host__ __device__ __forceinline__ double van_leer(double dqm, double dqp)
{
  return (2.0 * (dqm * dqp + fabs(dqm * dqp))) / ((dqm + dqp) * (dqm + dqp) + 1.0e-16);
}

// ---- phase 1a: thread = element, the 12 SCS points -------------------------------------------------------------------------
struct SElem {
  double X[3][8], Q[8], G[8];
  double o[3]; // coordinates of the element's first node (the coefficients are those of x - o)
  double qg, qa;
  const double *md;
};
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ void p1a_point(const SlistParams &P, const SElem &E, const double *pl0, const double *pl1, int n000,
                                                   double *rec)
{
  using C = Cfg;
  constexpr int U = (T + 1) % 3, V = (T + 2) % 3;
  double orient;
  const int ip = cont::point_index<T, SU, SV>(orient);
  double A[3], b[3];
  cont::area_plus_t<T, SU, SV>(E.X, A);
  A[0] *= orient;
  A[1] *= orient;
  A[2] *= orient;
  const double det = cont::grad_dot_area_c<T, SU, SV>(E.X, E.qg, A, b);
  if (!(det > 0.0)) *P.err = 1; // MasterElementFunctions.h:91-94
  double dq[3];
  cont::grad_c<T, SU, SV>(E.Q, E.qg, dq);
  const double gam = cont::value_c<T, SU, SV>(E.G, 0.25); // diffFluxCoeffIp: never shifted
  const double qip = cont::value_c<T, SU, SV>(E.Q, E.qa); // advection shape function
  const double tmdot = E.md ? E.md[ip] : 0.0;
  const double b0 = -gam * b[0], b1 = -gam * b[1], b2 = -gam * b[2];
  const double qDiff = b0 * dq[0] + b1 * dq[1] + b2 * dq[2];
  double pecfac = 0.0, aflux = tmdot * qip;
  if (P.upw) {
    // ScalarUpwAdvDiffElemKernel.C:128-206: Peclet factor along the edge, limited extrapolation from both ends
    const double *nl = node_ptr(pl0, pl1, n000, hex8::scs_left(ip)), *nr = node_ptr(pl0, pl1, n000, hex8::scs_right(ip));
    const double diffIp = 0.5 * (nl[C::FG] / nl[C::FR] + nr[C::FG] / nr[C::FR]);
    double udotx = 0.0, dqL = 0.0, dqR = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double xl = nl[C::FX + j] - E.o[j], xr = nr[C::FX + j] - E.o[j];
      const double cip = cont::value_c<T, SU, SV>(E.X[j], E.qa);
      udotx += 0.5 * (nl[C::FU + j] + nr[C::FU + j]) * (xr - xl);
      dqL += (cip - xl) * nl[C::FGQ + j];
      dqR += (xr - cip) * nr[C::FGQ + j];
    }
    pecfac = peclet_function(P, fabs(udotx) / (diffIp + 1.0e-16));
    const double om_pecfac = 1.0 - pecfac;
    double limitL = 1.0, limitR = 1.0;
    if (P.upw_options & NALU_OPT_USE_LIMITER) {
      const double dqe = nr[C::FQ] - nl[C::FQ];
      limitL = van_leer(2.0 * 2.0 * dqL - dqe, dqe);
      limitR = van_leer(2.0 * 2.0 * dqR - dqe, dqe);
    }
    const double qIpL = nl[C::FQ] + dqL * P.ho_upwind * limitL, qIpR = nr[C::FQ] - dqR * P.ho_upwind * limitR;
    const double om_alpha = 1.0 - P.alpha, om_alphaUpw = 1.0 - P.alpha_upw;
    const double qUpwind = tmdot > 0 ? P.alpha_upw * qIpL + om_alphaUpw * qip : P.alpha_upw * qIpR + om_alphaUpw * qip;
    const double qHatL = P.alpha * qIpL + om_alpha * qip, qHatR = P.alpha * qIpR + om_alpha * qip;
    aflux = tmdot * (pecfac * qUpwind + om_pecfac * (0.5 * (qHatL + qHatR)));
  }
  rec[ip * C::RS + T] = b0;
  rec[ip * C::RS + U] = b1;
  rec[ip * C::RS + V] = b2;
  rec[ip * C::RS + 3] = tmdot;
  rec[ip * C::RS + 4] = pecfac;
  rec[ip * C::RS + 5] = aflux + qDiff;
}
template <int T>
__host__ __device__ __forceinline__ void p1a_edge_dir(const SlistParams &P, const SElem &E, const double *pl0, const double *pl1, int n000,
                                                      double *rec)
{
  p1a_point<T, 0, 0>(P, E, pl0, pl1, n000, rec);
  p1a_point<T, 1, 0>(P, E, pl0, pl1, n000, rec);
  p1a_point<T, 0, 1>(P, E, pl0, pl1, n000, rec);
  p1a_point<T, 1, 1>(P, E, pl0, pl1, n000, rec);
}
// is element (elx, ely) of layer ez inside the block?
__host__ __device__ __forceinline__ bool elem_inside(const SlistParams &P, int el, int x0, int y0, int ez, int &ex, int &ey)
{
  using C = Cfg;
  ex = x0 - 1 + el % C::ELX;
  ey = y0 - 1 + el / C::ELX;
  return !(ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny || ez < 0 || ez >= P.nz);
}
__host__ __device__ inline void p1a_element(const SlistParams &P, double *sm, int el, int x0, int y0, int ez, const double *pl0,
                                            const double *pl1)
{
  using C = Cfg;
  int ex, ey;
  double *rec = sm + C::OFF_REC + el * C::ES;
  if (!elem_inside(P, el, x0, y0, ez, ex, ey)) {
#pragma unroll
    for (int i = 0; i < 12 * C::RS; ++i) rec[i] = 0.0;
    return;
  }
  const int n000 = ((el % C::ELX) + C::RNX * (el / C::ELX)) * C::NF;
  SElem E;
  E.md = P.mdot ? P.mdot + 12 * (ex + (long long)P.nx * (ey + (long long)P.ny * ez)) : nullptr;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gather_bits(pl0, pl1, n000, C::FX + i, E.X[i]);
    E.o[i] = E.X[i][0]; // relative to the element's first node (see cont_tile.cuh)
#pragma unroll
    for (int m = 0; m < 8; ++m) E.X[i][m] -= E.o[i];
    box::butterfly8(E.X[i]);
  }
  gather_bits(pl0, pl1, n000, C::FQ, E.Q);
  gather_bits(pl0, pl1, n000, C::FG, E.G);
  box::butterfly8(E.Q);
  box::butterfly8(E.G);
  E.qg = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  E.qa = (P.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25;
  p1a_edge_dir<0>(P, E, pl0, pl1, n000, rec);
  p1a_edge_dir<1>(P, E, pl0, pl1, n000, rec);
  p1a_edge_dir<2>(P, E, pl0, pl1, n000, rec);
}

// ---- phase 1b: thread = element, the 8 sub-control volumes -------------------------------------------------------------------
// A trilinear field with coefficients F at eta = 2 xi: value and parametric gradient (cont_tile.cuh's normalisation)
__host__ __device__ __forceinline__ double value_eta(const double (&F)[8], double ex, double ey, double ez)
{
  return 0.125 * (F[0] + F[1] * ex + F[2] * ey + F[3] * (ex * ey) + ez * (F[4] + F[5] * ex + F[6] * ey + F[7] * (ex * ey)));
}
__host__ __device__ __forceinline__ void pgrad_eta(const double (&F)[8], double ex, double ey, double ez, double (&d)[3])
{
  d[0] = 0.25 * (F[1] + F[3] * ey + F[5] * ez + F[7] * (ey * ez));
  d[1] = 0.25 * (F[2] + F[3] * ex + F[6] * ez + F[7] * (ex * ez));
  d[2] = 0.25 * (F[4] + F[5] * ex + F[6] * ey + F[7] * (ex * ey));
}
// d/dx_j = sum_a (d/dxi_a) n_a[j] / det  (n_a = det grad xi_a: the cross products of the Jacobian's columns)
__host__ __device__ __forceinline__ void to_physical(const double (&d)[3], const double (&n)[3][3], double inv, double (&g)[3])
{
#pragma unroll
  for (int j = 0; j < 3; ++j) g[j] = (d[0] * n[0][j] + d[1] * n[1][j] + d[2] * n[2][j]) * inv;
}
// nodal field f interpolated to the SCV point of corner n with the SCV shape function at xi = s * sigma
// (HexSCV::shape_fcn / shifted_shape_fcn, Hex8CVFEM.C:149-178: s = 1/4; shifted: the node itself)
__host__ __device__ __forceinline__ double scv_interp(const double *pl0, const double *pl1, int n000, int f, int n, bool lumped)
{
  if (lumped) return node_ptr(pl0, pl1, n000, n)[f];
  double v = 0.0;
#pragma unroll
  for (int ic = 0; ic < 8; ++ic) {
    const int same = (bit_x(ic) == bit_x(n)) + (bit_y(ic) == bit_y(n)) + (bit_z(ic) == bit_z(n));
    const double w = same == 3 ? 0.421875 : (same == 2 ? 0.140625 : (same == 1 ? 0.046875 : 0.015625)); // (3/4)^s (1/4)^(3-s)
    v += w * node_ptr(pl0, pl1, n000, ic)[f];
  }
  return v;
}
__host__ __device__ inline void p1b_element(const SlistParams &P, double *sm, int el, int x0, int y0, int ez, const double *pl0,
                                            const double *pl1)
{
  using C = Cfg;
  int ex, ey;
  double *rec = sm + C::OFF_REC + el * C::ES;
  if (!elem_inside(P, el, x0, y0, ez, ex, ey)) {
#pragma unroll
    for (int i = 0; i < 8 * C::RV; ++i) rec[i] = 0.0;
    return;
  }
  const int n000 = ((el % C::ELX) + C::RNX * (el / C::ELX)) * C::NF;
  box5::Coef Cx;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gather_bits(pl0, pl1, n000, C::FX + i, Cx.X[i]);
    const double o = Cx.X[i][0];
#pragma unroll
    for (int m = 0; m < 8; ++m) Cx.X[i][m] -= o;
    box::butterfly8(Cx.X[i]);
  }
  // HexSCV::determinant (Hex8CVFEM.C:149-178): the 8 Grandy volumes from the coefficient vectors (box5_tile.cuh)
  double vol[16];
  box5::scv_volumes_part<0>(Cx, vol);
  box5::scv_volumes_part<1>(Cx, vol);
  if (!P.src) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      rec[n * C::RV + 0] = vol[n] + vol[8 + n];
      rec[n * C::RV + 1] = 0.0;
      rec[n * C::RV + 2] = 0.0;
    }
    return;
  }
  double Uc[3][8], Kc[8], Wc[8];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gather_bits(pl0, pl1, n000, C::FU + i, Uc[i]);
    box::butterfly8(Uc[i]);
  }
  gather_bits(pl0, pl1, n000, P.src == 1 ? C::FQ : C::FO, Kc); // turbulent_ke
  gather_bits(pl0, pl1, n000, P.src == 1 ? C::FO : C::FQ, Wc); // specific_dissipation_rate
  box::butterfly8(Kc);
  box::butterfly8(Wc);
  const double eg = P.src_shifted ? 1.0 : 0.5, es = P.src_lumped ? 1.0 : 0.5;
  const double betaStar = P.turb[NALU_TM_BETA_STAR], prodLimit = P.turb[NALU_TM_TKE_PROD_LIMIT_RATIO];
#pragma unroll 1
  for (int n = 0; n < 8; ++n) {
    const double sx = bit_x(n) ? 1.0 : -1.0, sy = bit_y(n) ? 1.0 : -1.0, sz = bit_z(n) ? 1.0 : -1.0;
    const double gx = eg * sx, gy = eg * sy, gz = eg * sz;
    // generic_grad_op (MasterElementFunctions.h:51-108) at the SCV point: columns of the Jacobian, their cross products
    double c[3][3], nn[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double d[3];
      pgrad_eta(Cx.X[i], gx, gy, gz, d);
      c[0][i] = d[0];
      c[1][i] = d[1];
      c[2][i] = d[2];
    }
    cont::cross3(c[1], c[2], nn[0]);
    cont::cross3(c[2], c[0], nn[1]);
    cont::cross3(c[0], c[1], nn[2]);
    const double det = c[0][0] * nn[0][0] + c[0][1] * nn[0][1] + c[0][2] * nn[0][2];
    if (!(det > 0.0)) *P.err = 1;
    const double inv = 1.0 / det;
    double dudx[3][3], dkdx[3], dwdx[3], d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      pgrad_eta(Uc[i], gx, gy, gz, d);
      to_physical(d, nn, inv, dudx[i]);
    }
    pgrad_eta(Kc, gx, gy, gz, d);
    to_physical(d, nn, inv, dkdx);
    pgrad_eta(Wc, gx, gy, gz, d);
    to_physical(d, nn, inv, dwdx);
    const double tke = value_eta(Kc, es * sx, es * sy, es * sz), sdr = value_eta(Wc, es * sx, es * sy, es * sz);
    const double rho = scv_interp(pl0, pl1, n000, C::FR, n, P.src_lumped);
    const double tvisc = scv_interp(pl0, pl1, n000, C::FT, n, P.src_lumped);
    double Pk = 0.0, crossDiff = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      crossDiff += dkdx[i] * dwdx[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) Pk += dudx[i][j] * (dudx[i][j] + dudx[j][i]);
    }
    Pk *= tvisc;
    const double scV = vol[n] + vol[8 + n];
    double f, r;
    if (P.src == 1) {
      // TurbKineticEnergySSTSrcElemKernel.C:155-170
      const double tkeFac = betaStar * rho * sdr;
      const double Dk = tkeFac * tke;
      Pk = fmin(Pk, prodLimit * Dk);
      r = (Pk - Dk) * scV;
      f = tkeFac * scV;
    } else {
      // SpecificDissipationRateSSTSrcElemKernel.C:167-200
      const double fOneBlend = scv_interp(pl0, pl1, n000, C::FF, n, P.src_lumped);
      const double Dk = betaStar * rho * sdr * tke;
      Pk = fmin(Pk, prodLimit * Dk);
      const double om_fOneBlend = 1.0 - fOneBlend;
      const double beta = fOneBlend * P.turb[NALU_TM_BETA_ONE] + om_fOneBlend * P.turb[NALU_TM_BETA_TWO];
      const double gamma = fOneBlend * P.turb[NALU_TM_GAMMA_ONE] + om_fOneBlend * P.turb[NALU_TM_GAMMA_TWO];
      const double sigmaD = 2.0 * om_fOneBlend * P.turb[NALU_TM_SIGMA_W_TWO];
      const double Pw = gamma * rho * Pk / fmax(tvisc, 1.0e-16);
      const double Dw = beta * rho * sdr * sdr;
      const double Sw = sigmaD * rho * crossDiff / sdr;
      r = (Pw - Dw + Sw) * scV;
      f = (2.0 * beta * rho * sdr + fmax(Sw / sdr, 0.0)) * scV;
    }
    rec[n * C::RV + 0] = scV;
    rec[n * C::RV + 1] = f;
    rec[n * C::RV + 2] = r;
  }
}

// ---- phase 2a: thread = node (lx, ly) of the tile, SCS records (see cont_tile.cuh for the row bookkeeping) ----------------
// what phase 2 derives from (mdot, pecfac): ScalarUpwAdvDiffElemKernel.C:208-235.  For the central kernels the launcher
// sets alpha = alpha_upw = 0 and phase 1 leaves pecfac = 0: adv = mdot, aL = aR = 0.
struct UpwK {
  double alpha, alpha_upw;
};
template <bool SH, bool SK, int SX, int SY, int DIR, int LVL>
__host__ __device__ __forceinline__ void p2a_point(const UpwK &K, const double *er, double (&acc)[27], double &r)
{
  using C = Cfg;
  constexpr int bx = !SX, by = !SY, bz = LVL;
  const int k = node_of_bits(bx, by, bz);
  const int ip = ip_of(k, DIR);
  const int il = hex8::scs_left(ip), ir = hex8::scs_right(ip);
  const double sgn = il == k ? 1.0 : -1.0;
  double s1, s2, s3, a1, a2, a3;
  hex8::scs_intg_loc(ip, SH ? 0.5 : 0.25, s1, s2, s3);
  hex8::scs_intg_loc(ip, SK ? 0.5 : 0.25, a1, a2, a3);
  const double b0 = er[ip * C::RS + 0], b1 = er[ip * C::RS + 1], b2 = er[ip * C::RS + 2], md = er[ip * C::RS + 3],
               pf = er[ip * C::RS + 4], s = er[ip * C::RS + 5];
  const double om_pf = 1.0 - pf;
  const double adv = md * (pf * (1.0 - K.alpha_upw) + om_pf * (1.0 - K.alpha));
  const double cds = 0.5 * K.alpha * om_pf * md;
  const double aL = 0.5 * (md + fabs(md)) * pf * K.alpha_upw + cds, aR = 0.5 * (md - fabs(md)) * pf * K.alpha_upw + cds;
#pragma unroll
  for (int ic = 0; ic < 8; ++ic) {
    double d[3];
    hex8::shape_deriv(ic, s1, s2, s3, d);
    const int idx = (bit_z(ic) - bz + 1) * 9 + (bit_y(ic) - by + 1) * 3 + (bit_x(ic) - bx + 1);
    acc[idx] += (sgn * d[0]) * b0;
    acc[idx] += (sgn * d[1]) * b1;
    acc[idx] += (sgn * d[2]) * b2;
    acc[idx] += (sgn * hex8::shape_fcn(ic, a1, a2, a3)) * adv;
  }
  const int xl = (bit_z(il) - bz + 1) * 9 + (bit_y(il) - by + 1) * 3 + (bit_x(il) - bx + 1);
  const int xr = (bit_z(ir) - bz + 1) * 9 + (bit_y(ir) - by + 1) * 3 + (bit_x(ir) - bx + 1);
  acc[xl] += sgn * aL;
  acc[xr] += sgn * aR;
  r -= sgn * s;
}
template <bool SH, bool SK, int SX, int SY>
__host__ __device__ __forceinline__ void p2a_element(const UpwK &K, const double *sm, int lx, int ly, double (&accA)[27], double &rA,
                                                     double (&accB)[27], double &rB)
{
  using C = Cfg;
  const double *er = sm + C::OFF_REC + ((lx + SX) + C::ELX * (ly + SY)) * C::ES;
  p2a_point<SH, SK, SX, SY, 0, 0>(K, er, accA, rA);
  p2a_point<SH, SK, SX, SY, 1, 0>(K, er, accA, rA);
  p2a_point<SH, SK, SX, SY, 2, 0>(K, er, accA, rA);
  p2a_point<SH, SK, SX, SY, 0, 1>(K, er, accB, rB);
  p2a_point<SH, SK, SX, SY, 1, 1>(K, er, accB, rB);
  p2a_point<SH, SK, SX, SY, 2, 1>(K, er, accB, rB);
}
template <bool SH, bool SK>
__host__ __device__ __forceinline__ void p2a_layer(const UpwK &K, const double *sm, int lx, int ly, double (&accA)[27], double &rA,
                                                   double (&accB)[27], double &rB)
{
  p2a_element<SH, SK, 0, 0>(K, sm, lx, ly, accA, rA, accB, rB);
  p2a_element<SH, SK, 1, 0>(K, sm, lx, ly, accA, rA, accB, rB);
  p2a_element<SH, SK, 0, 1>(K, sm, lx, ly, accA, rA, accB, rB);
  p2a_element<SH, SK, 1, 1>(K, sm, lx, ly, accA, rA, accB, rB);
}

// ---- phase 2b: thread = node, SCV records: the corner of each of the 4 elements of the layer that is the row node -----------
template <int SX, int SY, int LVL>
__host__ __device__ __forceinline__ void p2b_corner(const double *sm, int lx, int ly, bool lumped, double (&acc)[27], double &r, double &vol)
{
  using C = Cfg;
  constexpr int bx = !SX, by = !SY, bz = LVL;
  const int k = node_of_bits(bx, by, bz);
  const double *er = sm + C::OFF_REC + ((lx + SX) + C::ELX * (ly + SY)) * C::ES + k * C::RV;
  const double V = er[0], f = er[1];
  vol += V;
  r += er[2];
  if (lumped) {
    acc[13] += f;
  } else {
    // lhs(nearestNode, ic) += N_ic(SCV point) f: N = prod over the axes of 3/4 (column on the corner's side) or 1/4
#pragma unroll
    for (int ic = 0; ic < 8; ++ic) {
      const int same = (bit_x(ic) == bx) + (bit_y(ic) == by) + (bit_z(ic) == bz);
      const double w = same == 3 ? 0.421875 : (same == 2 ? 0.140625 : (same == 1 ? 0.046875 : 0.015625));
      const int idx = (bit_z(ic) - bz + 1) * 9 + (bit_y(ic) - by + 1) * 3 + (bit_x(ic) - bx + 1);
      acc[idx] += w * f;
    }
  }
}
__host__ __device__ __forceinline__ void p2b_layer(const double *sm, int lx, int ly, bool lumped, double (&accA)[27], double &rA, double &volA,
                                                   double (&accB)[27], double &rB, double &volB)
{
  p2b_corner<0, 0, 0>(sm, lx, ly, lumped, accA, rA, volA);
  p2b_corner<1, 0, 0>(sm, lx, ly, lumped, accA, rA, volA);
  p2b_corner<0, 1, 0>(sm, lx, ly, lumped, accA, rA, volA);
  p2b_corner<1, 1, 0>(sm, lx, ly, lumped, accA, rA, volA);
  p2b_corner<0, 0, 1>(sm, lx, ly, lumped, accB, rB, volB);
  p2b_corner<1, 0, 1>(sm, lx, ly, lumped, accB, rB, volB);
  p2b_corner<0, 1, 1>(sm, lx, ly, lumped, accB, rB, volB);
  p2b_corner<1, 1, 1>(sm, lx, ly, lumped, accB, rB, volB);
}
// lumped ScalarMassElemKernel (src/kernel/ScalarMassElemKernel.C:117-154 with the shifted shape function): every quantity
// is the row node's own, times its dual volume
__host__ __device__ __forceinline__ void mass_apply(const SlistParams &P, int x, int y, int z, double vol, double &diag, double &r)
{
  const long long n = x + (long long)P.NX * (y + (long long)P.NY * z);
  const double rp = P.rho[n];
  diag += P.g1_dt * rp * vol;
  r -= (P.g1_dt * rp * P.q[n] + P.g2_dt * P.rho_n[n] * P.q_n[n] + P.g3_dt * P.rho_nm1[n] * P.q_nm1[n]) * vol;
}

} // namespace slist
} // namespace nalu_b200
