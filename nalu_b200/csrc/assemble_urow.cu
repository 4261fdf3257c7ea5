// assemble_urow.cu -- momentum assembly on ANY Hex8 connectivity (unstructured meshes, periodic aliasing, decomposed
// pieces) without the warp-per-element scratch matrix of the general kernel:
//   [MomentumMassElemKernel(lumped)]? + MomentumAdvDiffElemKernel  (src/LowMachEquationSystem.C:1304-1398, BASELINE config 3)
// through AssembleElemSolverAlgorithm::execute (src/AssembleElemSolverAlgorithm.C:74-100) + sumInto
// (src/TpetraLinearSystem.C:1457-1522).  Kernels followed: src/kernel/MomentumAdvDiffElemKernel.C:95-186,
// src/kernel/MomentumMassElemKernel.C:98-171 (lumped).
//
// Decomposition: a thread owns (element, row node r): the 3 x 24 block of rows the element adds to node r.  Every sub-control
// surface of the element that touches r is one of the three edges leaving r, so the thread needs the geometry of three
// integration points only: area vector, G~ = -mu_ip adj(J) / det J, signed mass flow -- 16 doubles per point, after which
// the coordinates are dead.  It then walks the 8 column nodes and finishes ONE 3 x 3 block at a time:
//     h = G~ dN_c/dxi (9 FMAs), block += h (x) A (9), D_c += N_c mdot + h . A (4), divU_ip += u_c . h (3)
// -- 10 accumulators live instead of the 72 a whole row block would take -- adds the block to the CSR through the element's
// slot map and folds it into the residual (rhs = -(row . u) + 2/3 includeDivU sum_ip (mu divU)_ip A_ip, the form the tiled
// kernel uses).  No shared memory, no warp barriers: the 6 350 warp instructions per element of the general kernel become
// about 600.  Elements are processed colour by colour (graph.cpp): within a launch one element touches a row, so the
// scatter is value-less reductions (red.global.add.f64), deterministic.
#include "box5_tile.cuh"
#include "hex8_dev.cuh"
#include "internal.h"

namespace nalu_b200 {
namespace urow {

struct UrowParams {
  long long n_elems;       // elements of this colour
  const int *elem_list;    // their ids
  long long n_rows;        // row-owner kernel: row nodes, and the elements around each as (element << 3 | local node)
  const long long *rowelem_ptr;
  const int *rowelem;
  const int *elem_nodes, *node_row;
  const long long *browptr;
  const unsigned char *elem_slot;
  double *values, *rhs;
  int *err;
  // MomentumAdvDiffElemKernel (vel) / ScalarAdvDiffElemKernel, ScalarDiffElemKernel (q; visc = diffFluxCoeff; mdot may be null)
  const double *coords, *vel, *visc, *mdot, *q;
  // ContinuityAdvElemKernel: vel (velocity_rtm), rho, pressure, dpdx; tau = dt / gamma1; NALU_OPT_SHIFT_MDOT / _SHIFT_POISSON /
  // _REDUCED_SENS in options
  const double *pressure;
  double tau;
  // ScalarUpwAdvDiffElemKernel (upw = 1): Gjq, velocity_rtm in vel, density in rho; constants as in nalu_b200_kernel_desc::upw
  int upw, upw_options;
  const double *gjq;
  double alpha, alpha_upw, ho_upwind, hybrid, tanh_trans, tanh_width;
  // lumped ScalarMassElemKernel (mass = 1 with ndof 1): q_n, q_nm1 beside rho, rho_n, rho_nm1 and g*_dt
  const double *q_n, *q_nm1;
  // TurbKineticEnergySSTSrcElemKernel (src = 1) / SpecificDissipationRateSSTSrcElemKernel (src = 2): turbulent_ke,
  // specific_dissipation_rate, turbulent_viscosity, sst_f_one_blending, velocity in src_vel, density in rho; NALU_TM_* constants
  int src, src_lumped, src_shifted;
  const double *tke, *sdr, *tvisc, *f1, *src_vel;
  double turb[10];
  int options;
  double include_divu;
  // MomentumUpwAdvDiffElemKernel (upw = 1; constants shared with the scalar upwind kernel below): dudx in gju, velocity_rtm
  // in vrtm (the kernel's own velocity stays in vel), density in rho.
  // MomentumBuoyancySrcElemKernel (buoy = 1) / MomentumBuoyancyBoussinesqSrcElemKernel (buoy = 2): bfield = density / temperature
  const double *gju, *vrtm, *bfield;
  int buoy;
  double gravity[3], rho_ref, t_ref, beta;
  // lumped MomentumMassElemKernel
  int mass;
  const double *vel_n, *vel_nm1, *rho, *rho_n, *rho_nm1, *dpdx;
  double g1_dt, g2_dt, g3_dt, dens_fac, om_dens_fac;
};

// No "memory" clobber: the matrix and the residual never alias the fields the kernel reads, and the loads of the next column
// must be free to move above the reductions of this one (the kernel is bound by gather latency, not by arithmetic).
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v)); }

// src/PecletFunction.C:67-71,98-101; ScalarUpwAdvDiffElemKernel.C:258-266 == MomentumUpwAdvDiffElemKernel.C:314-322
__device__ __forceinline__ double urow_peclet(const UrowParams &P, double x)
{
  if (P.upw_options & NALU_OPT_PECLET_TANH) return 0.50 * (1.0 + tanh((x - P.tanh_trans) / P.tanh_width));
  const double m = P.hybrid * x;
  return m * m / (5.0 + m * m);
}
__device__ __forceinline__ double urow_van_leer(double dqm, double dqp)
{
  return (2.0 * (dqm * dqp + fabs(dqm * dqp))) / ((dqm + dqp) * (dqm + dqp) + 1.0e-16);
}

// the 8 slot bytes of (element, row node) in one load: slot c of the row's block list holds the column block of element node c
struct Slots {
  unsigned long long w;
  __device__ __forceinline__ int operator[](int c) const { return (int)((w >> (8 * c)) & 0xFFull); }
};
__device__ __forceinline__ Slots load_slots(const unsigned char *elem_slot, long long e, int r)
{
  return Slots{*reinterpret_cast<const unsigned long long *>(elem_slot + e * 64 + r * 8)};
}

struct IpGeom {
  double As[3];    // area vector, signed so that it points OUT of node r's control volume
  double Gt[3][3]; // Gt[i][a] = -mu_ip adj(J)[i][a] / det J: -mu dN/dx_i = sum_a Gt[i][a] dN/dxi_a
  double tm;       // signed mass flow rate
  double s[3];     // parametric location of the gradient operator
};

// geometry of the integration point on the edge leaving node r in direction DIR
__device__ __forceinline__ void ip_geometry(const UrowParams &P, const double (&x)[8][3], const int (&nodes)[8], long long e, int r,
                                            int dir, IpGeom &g, double &mu_out)
{
  const int ip = box::ip_of(r, dir);
  const double sgn = hex8::scs_left(ip) == r ? 1.0 : -1.0;
  double A[3];
  hex8::scs_areav_ip(x, ip, A);
  hex8::scs_intg_loc(ip, (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25, g.s[0], g.s[1], g.s[2]);
  double jact[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double u1, u2, u3, mu = 0.0;
  hex8::scs_intg_loc(ip, 0.25, u1, u2, u3); // properties are never shifted
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    double d[3];
    hex8::shape_deriv(n, g.s[0], g.s[1], g.s[2], d);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) jact[i][j] += d[j] * x[n][i];
    mu += hex8::shape_fcn(n, u1, u2, u3) * P.visc[nodes[n]];
  }
  double adj[3][3];
  adj[0][0] = jact[1][1] * jact[2][2] - jact[2][1] * jact[1][2];
  adj[0][1] = jact[1][2] * jact[2][0] - jact[2][2] * jact[1][0];
  adj[0][2] = jact[1][0] * jact[2][1] - jact[2][0] * jact[1][1];
  adj[1][0] = jact[0][2] * jact[2][1] - jact[2][2] * jact[0][1];
  adj[1][1] = jact[0][0] * jact[2][2] - jact[2][0] * jact[0][2];
  adj[1][2] = jact[0][1] * jact[2][0] - jact[2][1] * jact[0][0];
  adj[2][0] = jact[0][1] * jact[1][2] - jact[1][1] * jact[0][2];
  adj[2][1] = jact[0][2] * jact[1][0] - jact[1][2] * jact[0][0];
  adj[2][2] = jact[0][0] * jact[1][1] - jact[1][0] * jact[0][1];
  const double det = jact[0][0] * adj[0][0] + jact[1][0] * adj[1][0] + jact[2][0] * adj[2][0];
  if (!(det > 0.0)) *P.err = 1; // MasterElementFunctions.h:91-94
  const double f = -mu / det;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g.As[i] = sgn * A[i];
#pragma unroll
    for (int a = 0; a < 3; ++a) g.Gt[i][a] = f * adj[i][a];
  }
  g.tm = P.mdot ? sgn * P.mdot[e * 12 + ip] : 0.0;
  mu_out = mu;
}

// The rows element e adds to its node r.  OWNED: the calling thread is the only one that ever touches these rows (row-owner
// kernel): plain adds, the residual is returned in res[].  Otherwise (element-colour kernel): value-less reductions.
template <bool OWNED>
__device__ __forceinline__ void row_contribution(const UrowParams &P, long long e, int r, double (&res)[3])
{
  int nodes[8];
  double x[8][3];
#pragma unroll
  for (int n = 0; n < 8; ++n) nodes[n] = P.elem_nodes[e * 8 + n];
  {
    // coordinates relative to the element's first node: the assembled system is translation invariant, the rounding of
    // differences of absolute coordinates is not
    const double o0 = P.coords[(size_t)nodes[0] * 3], o1 = P.coords[(size_t)nodes[0] * 3 + 1], o2 = P.coords[(size_t)nodes[0] * 3 + 2];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      x[n][0] = P.coords[(size_t)nodes[n] * 3] - o0;
      x[n][1] = P.coords[(size_t)nodes[n] * 3 + 1] - o1;
      x[n][2] = P.coords[(size_t)nodes[n] * 3 + 2] - o2;
    }
  }
  // everything the tail of the kernel gathers is requested now, so that its latency hides behind the geometry
  const int nd = nodes[r];
  const int rn = P.node_row[nd];
  const long long b0 = P.browptr[rn];
  const int nb = (int)(P.browptr[rn + 1] - b0);
  double m_rho = 0, m_rho_n = 0, m_rho_nm1 = 0, m_un[3] = {0, 0, 0}, m_unm1[3] = {0, 0, 0}, m_gp[3] = {0, 0, 0};
  if (P.mass) {
    m_rho = P.rho[nd];
    m_rho_n = P.rho_n[nd];
    m_rho_nm1 = P.rho_nm1[nd];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      m_un[i] = P.vel_n[(size_t)nd * 3 + i];
      m_unm1[i] = P.vel_nm1[(size_t)nd * 3 + i];
      m_gp[i] = P.dpdx[(size_t)nd * 3 + i];
    }
  }
  // lumped mass: the sub-control volume of corner r (HexSCV::determinant, Hex8CVFEM.C:149-178, in the closed form of the
  // tiled kernel: box5_tile.cuh) -- every quantity of the lumped kernel is the row node's own
  double scV = 0.0;
  if (P.mass) {
    box5::Coef C;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int m = 0; m < 8; ++m) C.X[i][m] = x[box::node_of_bits(m & 1, (m >> 1) & 1, m >> 2)][i];
      box::butterfly8(C.X[i]);
    }
    double vol[16];
    box5::scv_volumes_part<0>(C, vol);
    box5::scv_volumes_part<1>(C, vol);
    scV = 0.0;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n == r) scV = vol[n] + vol[8 + n];
  }
  IpGeom g[3];
  double mu_ip[3];
  ip_geometry(P, x, nodes, e, r, 0, g[0], mu_ip[0]);
  ip_geometry(P, x, nodes, e, r, 1, g[1], mu_ip[1]);
  ip_geometry(P, x, nodes, e, r, 2, g[2], mu_ip[2]);
  const double qa = (P.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25;
  double sa[3][3]; // advection shape-function location of the three points
#pragma unroll
  for (int k = 0; k < 3; ++k) hex8::scs_intg_loc(box::ip_of(r, k), qa, sa[k][0], sa[k][1], sa[k][2]);

  double *const row0 = P.values + 9 * b0;
  const unsigned char *slots = P.elem_slot + e * 64 + r * 8; // byte loads here: the packed word costs this function 2.8x (register pressure, measured)
  double divacc[3] = {0.0, 0.0, 0.0};

  double un[3] = {P.vel[(size_t)nodes[0] * 3], P.vel[(size_t)nodes[0] * 3 + 1], P.vel[(size_t)nodes[0] * 3 + 2]};
  // row-owner form: the current values of a 3 x 3 block are requested one column ahead, so that the read-modify-write of a
  // block does not wait for its own load (the thread is the only writer of these rows; a block read ahead is stale only if
  // the next column is the SAME column -- an element that reaches a node twice -- and is then re-read)
  double old[3][3];
  if (OWNED) {
    const double *b = row0 + slots[0] * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) old[i][j] = b[(long long)i * nb * 3 + j];
  }
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    double B[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, D = 0.0;
    const double u[3] = {un[0], un[1], un[2]};
    {
      // the next column's velocity is requested before this column's arithmetic
      const int nx = nodes[(c + 1) & 7];
      un[0] = P.vel[(size_t)nx * 3];
      un[1] = P.vel[(size_t)nx * 3 + 1];
      un[2] = P.vel[(size_t)nx * 3 + 2];
    }
    double cur[3][3];
    if (OWNED) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) cur[i][j] = old[i][j];
      if (c < 7) {
        const double *b = row0 + slots[c + 1] * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) old[i][j] = b[(long long)i * nb * 3 + j];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double d[3], h[3];
      hex8::shape_deriv(c, g[k].s[0], g[k].s[1], g[k].s[2], d);
#pragma unroll
      for (int i = 0; i < 3; ++i) h[i] = g[k].Gt[i][0] * d[0] + g[k].Gt[i][1] * d[1] + g[k].Gt[i][2] * d[2];
      // MomentumAdvDiffElemKernel.C:138-176: lhsfacAdv + lhs_riC_i on the diagonal, lhsfacDiff_j = -mu dN/dx_i A_j
      D += hex8::shape_fcn(c, sa[k][0], sa[k][1], sa[k][2]) * g[k].tm + h[0] * g[k].As[0] + h[1] * g[k].As[1] + h[2] * g[k].As[2];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B[i][j] += h[i] * g[k].As[j];
      divacc[k] += u[0] * h[0] + u[1] * h[1] + u[2] * h[2]; // = -mu_ip (u_c . grad N_c)
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) B[i][i] += D;
    if (P.mass && c == r) {
      // MomentumMassElemKernel.C:119-170 with the shifted (nodal) shape function
      const double rp = m_rho, lhsfac = P.g1_dt * rp * scV, densFac = rp * P.dens_fac + P.om_dens_fac;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        B[i][i] += lhsfac;
        // the residual below subtracts B . u: put back lhsfac u_i, then the kernel's own rhs
        res[i] += lhsfac * u[i] -
                  (P.g1_dt * rp * u[i] + P.g2_dt * m_rho_n * m_un[i] + P.g3_dt * m_rho_nm1 * m_unm1[i]) * scV - m_gp[i] * scV * densFac;
      }
    }
    double *blk = row0 + slots[c] * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      res[i] -= B[i][0] * u[0] + B[i][1] * u[1] + B[i][2] * u[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (OWNED) blk[(long long)i * nb * 3 + j] = cur[i][j] + B[i][j];
        else red_add(blk + (long long)i * nb * 3 + j, B[i][j]);
      }
    }
    if (OWNED && c < 7 && slots[c + 1] == slots[c]) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) old[i][j] = blk[(long long)i * nb * 3 + j];
    }
  }
  // divUstress = 2/3 mu divU A_i includeDivU, rhs -= sgn divUstress (MomentumAdvDiffElemKernel.C:118-136); divacc = -mu divU
#pragma unroll
  for (int i = 0; i < 3; ++i)
    res[i] += (2.0 / 3.0) * P.include_divu * (divacc[0] * g[0].As[i] + divacc[1] * g[1].As[i] + divacc[2] * g[2].As[i]);
  (void)mu_ip;
}

// The same for the lists with MomentumUpwAdvDiffElemKernel (UPW) and / or a buoyancy source (row-owner form only).  Kept apart
// from row_contribution: merged into one function the central list ran three times slower (the scheduling of its gathers).
template <bool OWNED, bool UPW>
__device__ __forceinline__ void row_contribution_ext(const UrowParams &P, long long e, int r, double (&res)[3])
{
  int nodes[8];
  double x[8][3];
#pragma unroll
  for (int n = 0; n < 8; ++n) nodes[n] = P.elem_nodes[e * 8 + n];
  {
    // coordinates relative to the element's first node: the assembled system is translation invariant, the rounding of
    // differences of absolute coordinates is not
    const double o0 = P.coords[(size_t)nodes[0] * 3], o1 = P.coords[(size_t)nodes[0] * 3 + 1], o2 = P.coords[(size_t)nodes[0] * 3 + 2];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      x[n][0] = P.coords[(size_t)nodes[n] * 3] - o0;
      x[n][1] = P.coords[(size_t)nodes[n] * 3 + 1] - o1;
      x[n][2] = P.coords[(size_t)nodes[n] * 3 + 2] - o2;
    }
  }
  // everything the tail of the kernel gathers is requested now, so that its latency hides behind the geometry
  const int nd = nodes[r];
  const int rn = P.node_row[nd];
  const long long b0 = P.browptr[rn];
  const int nb = (int)(P.browptr[rn + 1] - b0);
  double m_rho = 0, m_rho_n = 0, m_rho_nm1 = 0, m_un[3] = {0, 0, 0}, m_unm1[3] = {0, 0, 0}, m_gp[3] = {0, 0, 0};
  if (P.mass) {
    m_rho = P.rho[nd];
    m_rho_n = P.rho_n[nd];
    m_rho_nm1 = P.rho_nm1[nd];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      m_un[i] = P.vel_n[(size_t)nd * 3 + i];
      m_unm1[i] = P.vel_nm1[(size_t)nd * 3 + i];
      m_gp[i] = P.dpdx[(size_t)nd * 3 + i];
    }
  }
  // lumped mass: the sub-control volume of corner r (HexSCV::determinant, Hex8CVFEM.C:149-178, in the closed form of the
  // tiled kernel: box5_tile.cuh) -- every quantity of the lumped kernel is the row node's own
  double scV = 0.0;
  if (P.mass || P.buoy) {
    box5::Coef C;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int m = 0; m < 8; ++m) C.X[i][m] = x[box::node_of_bits(m & 1, (m >> 1) & 1, m >> 2)][i];
      box::butterfly8(C.X[i]);
    }
    double vol[16];
    box5::scv_volumes_part<0>(C, vol);
    box5::scv_volumes_part<1>(C, vol);
    scV = 0.0;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n == r) scV = vol[n] + vol[8 + n];
  }
  IpGeom g[3];
  double mu_ip[3];
  ip_geometry(P, x, nodes, e, r, 0, g[0], mu_ip[0]);
  ip_geometry(P, x, nodes, e, r, 1, g[1], mu_ip[1]);
  ip_geometry(P, x, nodes, e, r, 2, g[2], mu_ip[2]);
  const double qa = (P.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25;
  double sa[3][3]; // advection shape-function location of the three points
#pragma unroll
  for (int k = 0; k < 3; ++k) hex8::scs_intg_loc(box::ip_of(r, k), qa, sa[k][0], sa[k][1], sa[k][2]);

  // MomentumUpwAdvDiffElemKernel.C:150-243: per point the Peclet factor along the edge, the (limited) extrapolations from
  // both ends, the advective coefficient of the central part and the upwind factors of the two ends (signed)
  double advc[3], aLs[3], aRs[3], sg[3];
  int nbr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    advc[k] = g[k].tm;
    if (UPW) {
      const int ip = box::ip_of(r, k);
      const int il = hex8::scs_left(ip), ir = hex8::scs_right(ip);
      sg[k] = il == r ? 1.0 : -1.0;
      nbr[k] = il == r ? ir : il;
      // the two ends of the edge are read through the connectivity, not out of the register arrays: an index known only at
      // run time would push x[] and nodes[] into local memory
      const int nl = P.elem_nodes[e * 8 + il], nr = P.elem_nodes[e * 8 + ir];
      const int n0 = nodes[0];
      const double tmdot = sg[k] * g[k].tm;
      const double diffIp = 0.5 * (P.visc[nl] / P.rho[nl] + P.visc[nr] / P.rho[nr]);
      double cip[3], xl[3], xr[3], udotx = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        cip[j] = 0.0;
#pragma unroll
        for (int n = 0; n < 8; ++n) cip[j] += hex8::shape_fcn(n, sa[k][0], sa[k][1], sa[k][2]) * x[n][j];
        const double oj = P.coords[(size_t)n0 * 3 + j];
        xl[j] = P.coords[(size_t)nl * 3 + j] - oj;
        xr[j] = P.coords[(size_t)nr * 3 + j] - oj;
        udotx += 0.5 * (P.vrtm[(size_t)nl * 3 + j] + P.vrtm[(size_t)nr * 3 + j]) * (xr[j] - xl[j]);
      }
      const double pecfac = urow_peclet(P, fabs(udotx) / (diffIp + 1.0e-16)), om_pecfac = 1.0 - pecfac;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double duL = 0.0, duR = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          duL += (cip[j] - xl[j]) * P.gju[(size_t)nl * 9 + 3 * i + j];
          duR += (xr[j] - cip[j]) * P.gju[(size_t)nr * 9 + 3 * i + j];
        }
        const double ul = P.vel[(size_t)nl * 3 + i], ur = P.vel[(size_t)nr * 3 + i];
        double limL = 1.0, limR = 1.0;
        if (P.upw_options & NALU_OPT_USE_LIMITER) {
          const double dq = ur - ul;
          limL = urow_van_leer(2.0 * 2.0 * duL - dq, dq);
          limR = urow_van_leer(2.0 * 2.0 * duR - dq, dq);
        }
        // aflux_i = mdot (pecfac uUpwind_i + om_pecfac uCds_i) splits into a part that multiplies the interpolated u_ip --
        // it has the central kernel's form sum_c (N_c advc) u_ci and is taken in the column loop -- and the part of the
        // two extrapolated end values, which is known here
        const double uLi = ul + duL * P.ho_upwind * limL, uRi = ur - duR * P.ho_upwind * limR;
        res[i] -= g[k].tm * (pecfac * P.alpha_upw * (tmdot > 0 ? uLi : uRi) + om_pecfac * P.alpha * 0.5 * (uLi + uRi));
      }
      advc[k] = g[k].tm * (pecfac * (1.0 - P.alpha_upw) + om_pecfac * (1.0 - P.alpha));
      const double cds = 0.5 * P.alpha * om_pecfac * tmdot;
      aLs[k] = sg[k] * (0.5 * (tmdot + fabs(tmdot)) * pecfac * P.alpha_upw + cds);
      aRs[k] = sg[k] * (0.5 * (tmdot - fabs(tmdot)) * pecfac * P.alpha_upw + cds);
    }
  }
  if (P.buoy) {
    // MomentumBuoyancySrcElemKernel.C:61-90 / MomentumBuoyancyBoussinesqSrcElemKernel.C:66-94: the source of the sub-control
    // volume of corner r (its nearest node is r), SCV shape function never shifted
    double sIp = 0.0;
    const double v1 = hex8::sgn_x(r) * 0.25, v2 = hex8::sgn_y(r) * 0.25, v3 = hex8::sgn_z(r) * 0.25;
#pragma unroll
    for (int c = 0; c < 8; ++c) sIp += hex8::shape_fcn(c, v1, v2, v3) * P.bfield[nodes[c]];
    const double fac = P.buoy == 2 ? -P.rho_ref * P.beta * (sIp - P.t_ref) * scV : (sIp - P.rho_ref) * scV;
#pragma unroll
    for (int i = 0; i < 3; ++i) res[i] += fac * P.gravity[i];
  }

  double *const row0 = P.values + 9 * b0;
  const Slots slots = load_slots(P.elem_slot, e, r);
  double divacc[3] = {0.0, 0.0, 0.0};

  double un[3] = {P.vel[(size_t)nodes[0] * 3], P.vel[(size_t)nodes[0] * 3 + 1], P.vel[(size_t)nodes[0] * 3 + 2]};
  // row-owner form: the current values of a 3 x 3 block are requested one column ahead, so that the read-modify-write of a
  // block does not wait for its own load (the thread is the only writer of these rows; a block read ahead is stale only if
  // the next column is the SAME column -- an element that reaches a node twice -- and is then re-read)
  double old[3][3];
  if (OWNED) {
    const double *b = row0 + slots[0] * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) old[i][j] = b[(long long)i * nb * 3 + j];
  }
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    double B[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, D = 0.0, Dadv = 0.0, Dup = 0.0;
    const double u[3] = {un[0], un[1], un[2]};
    {
      // the next column's velocity is requested before this column's arithmetic
      const int nx = nodes[(c + 1) & 7];
      un[0] = P.vel[(size_t)nx * 3];
      un[1] = P.vel[(size_t)nx * 3 + 1];
      un[2] = P.vel[(size_t)nx * 3 + 2];
    }
    double cur[3][3];
    if (OWNED) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) cur[i][j] = old[i][j];
      if (c < 7) {
        const double *b = row0 + slots[c + 1] * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) old[i][j] = b[(long long)i * nb * 3 + j];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double d[3], h[3];
      hex8::shape_deriv(c, g[k].s[0], g[k].s[1], g[k].s[2], d);
#pragma unroll
      for (int i = 0; i < 3; ++i) h[i] = g[k].Gt[i][0] * d[0] + g[k].Gt[i][1] * d[1] + g[k].Gt[i][2] * d[2];
      // MomentumAdvDiffElemKernel.C:138-176: lhsfacAdv + lhs_riC_i on the diagonal, lhsfacDiff_j = -mu dN/dx_i A_j
      const double Nadv = hex8::shape_fcn(c, sa[k][0], sa[k][1], sa[k][2]);
      D += h[0] * g[k].As[0] + h[1] * g[k].As[1] + h[2] * g[k].As[2];
      Dadv += Nadv * advc[k];
      if (UPW) {
        // the upwind factors sit on the diagonal of the blocks of the two ends of the edge: r and its neighbour
        if (c == r) Dup += sg[k] > 0 ? aLs[k] : aRs[k];
        if (c == nbr[k]) Dup += sg[k] > 0 ? aRs[k] : aLs[k];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B[i][j] += h[i] * g[k].As[j];
      divacc[k] += u[0] * h[0] + u[1] * h[1] + u[2] * h[2]; // = -mu_ip (u_c . grad N_c)
    }
    // residual: -(diffusion row . u) - (N adv) u; the upwind end factors enter the matrix only (their flux was taken above)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      B[i][i] += D + Dadv;
      res[i] -= B[i][0] * u[0] + B[i][1] * u[1] + B[i][2] * u[2];
      B[i][i] += Dup;
    }
    if (P.mass && c == r) {
      // MomentumMassElemKernel.C:119-170 with the shifted (nodal) shape function
      const double rp = m_rho, lhsfac = P.g1_dt * rp * scV, densFac = rp * P.dens_fac + P.om_dens_fac;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        B[i][i] += lhsfac;
        res[i] -= (P.g1_dt * rp * u[i] + P.g2_dt * m_rho_n * m_un[i] + P.g3_dt * m_rho_nm1 * m_unm1[i]) * scV + m_gp[i] * scV * densFac;
      }
    }
    double *blk = row0 + slots[c] * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (OWNED) blk[(long long)i * nb * 3 + j] = cur[i][j] + B[i][j];
        else red_add(blk + (long long)i * nb * 3 + j, B[i][j]);
      }
    }
    if (OWNED && c < 7 && slots[c + 1] == slots[c]) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) old[i][j] = blk[(long long)i * nb * 3 + j];
    }
  }
  // divUstress = 2/3 mu divU A_i includeDivU, rhs -= sgn divUstress (MomentumAdvDiffElemKernel.C:118-136); divacc = -mu divU
#pragma unroll
  for (int i = 0; i < 3; ++i)
    res[i] += (2.0 / 3.0) * P.include_divu * (divacc[0] * g[0].As[i] + divacc[1] * g[1].As[i] + divacc[2] * g[2].As[i]);
  (void)mu_ip;
}

// one launch per element colour: thread = (element of the colour, row node)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) assemble_urow_kernel(const __grid_constant__ UrowParams P)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 3) >= P.n_elems) return;
  const long long e = P.elem_list[t >> 3];
  const int r = (int)(t & 7);
  double res[3] = {0.0, 0.0, 0.0};
  row_contribution<false>(P, e, r, res);
  const int rn = P.node_row[P.elem_nodes[e * 8 + r]];
#pragma unroll
  for (int i = 0; i < 3; ++i) red_add(P.rhs + (long long)rn * 3 + i, res[i]);
}

// ONE launch: thread = row node, walking the elements around it (rowelem: the transpose of the connectivity, built with
// the graph).  The thread owns its three rows, so every add is a plain read-modify-write by one thread in a fixed order:
// no atomics, no colouring, deterministic -- and an element that reaches a row through two of its nodes (a periodic box one
// cell wide) is simply visited twice.
template <int MINB, int KIND> // KIND 0: central list; 1: central + buoyancy source; 2: upwind list (+ buoyancy source)
__global__ void __launch_bounds__(128, MINB) assemble_urow_owner_kernel(const __grid_constant__ UrowParams P)
{
  const long long rn = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (rn >= P.n_rows) return;
  double res[3] = {0.0, 0.0, 0.0};
  const long long q1 = P.rowelem_ptr[rn + 1];
#pragma unroll 1
  for (long long q = P.rowelem_ptr[rn]; q < q1; ++q) {
    const int er = P.rowelem[q];
    if (KIND == 0) row_contribution<true>(P, er >> 3, er & 7, res);
    else row_contribution_ext<true, KIND == 2>(P, er >> 3, er & 7, res);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) P.rhs[rn * 3 + i] += res[i];
}


// ---- one dof per node: [ScalarMassElemKernel(lumped)]? + ScalarAdvDiffElemKernel / ScalarDiffElemKernel /
// ScalarUpwAdvDiffElemKernel (src/kernel/ScalarAdvDiffElemKernel.C:69-127, ScalarDiffElemKernel.C:64-110,
// ScalarUpwAdvDiffElemKernel.C:110-256, ScalarMassElemKernel.C:117-154), row-owner form.  Over the three points of node r:
// lhs(r, c) += sgn (N_c adv - Gamma_ip grad N_c . A) (+ the upwind factors on the two ends of the edge),
// rhs_r -= sgn (advective flux + sum_c (-Gamma grad N_c . A) q_c); per point only w_a = sum_i G~[i][a] A_i is needed.

template <bool UPW, bool SRC>
__device__ __forceinline__ void row_contribution_scalar(const UrowParams &P, long long e, int r, double &res)
{
  int nodes[8];
  double x[8][3];
#pragma unroll
  for (int n = 0; n < 8; ++n) nodes[n] = P.elem_nodes[e * 8 + n];
  {
    const double o0 = P.coords[(size_t)nodes[0] * 3], o1 = P.coords[(size_t)nodes[0] * 3 + 1], o2 = P.coords[(size_t)nodes[0] * 3 + 2];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      x[n][0] = P.coords[(size_t)nodes[n] * 3] - o0;
      x[n][1] = P.coords[(size_t)nodes[n] * 3 + 1] - o1;
      x[n][2] = P.coords[(size_t)nodes[n] * 3 + 2] - o2;
    }
  }
  const int rn = P.node_row[nodes[r]];
  double *const row0 = P.values + P.browptr[rn];
  const Slots slots = load_slots(P.elem_slot, e, r);
  // the row's current entries and the column values are requested before the geometry
  double old[8], qc[8];
  bool alias = false;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    old[c] = row0[slots[c]];
    qc[c] = P.q[nodes[c]];
#pragma unroll
    for (int b = 0; b < c; ++b) alias = alias || slots[b] == slots[c];
  }
  double w[3][3], tm[3], s[3][3], sa[3][3], mu;
  double advc[3], aL[3] = {0, 0, 0}, aR[3] = {0, 0, 0}, sg[3];
  int nbr[3]; // the other end of the edge leaving r in direction k
  const double qa = (P.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    IpGeom g;
    ip_geometry(P, x, nodes, e, r, k, g, mu);
#pragma unroll
    for (int a = 0; a < 3; ++a) w[k][a] = g.Gt[0][a] * g.As[0] + g.Gt[1][a] * g.As[1] + g.Gt[2][a] * g.As[2], s[k][a] = g.s[a];
    tm[k] = g.tm;
    const int ip = box::ip_of(r, k);
    hex8::scs_intg_loc(ip, qa, sa[k][0], sa[k][1], sa[k][2]);
    const int il = hex8::scs_left(ip), ir = hex8::scs_right(ip);
    sg[k] = il == r ? 1.0 : -1.0;
    nbr[k] = il == r ? ir : il;
    advc[k] = tm[k];
    if (UPW) {
      // ScalarUpwAdvDiffElemKernel.C:128-223: Peclet factor along the edge, limited extrapolation from both ends
      // the two ends of the edge are read through the connectivity, not out of the register arrays (see the momentum form)
      const int nl = P.elem_nodes[e * 8 + il], nr = P.elem_nodes[e * 8 + ir];
      const int n0 = nodes[0];
      const double qcl = P.q[nl], qcr = P.q[nr];
      const double tmdot = sg[k] * tm[k];
      const double diffIp = 0.5 * (P.visc[nl] / P.rho[nl] + P.visc[nr] / P.rho[nr]);
      double udotx = 0.0, dqL = 0.0, dqR = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double cip = 0.0; // adv-shape-function interpolation of the coordinates
#pragma unroll
        for (int n = 0; n < 8; ++n) cip += hex8::shape_fcn(n, sa[k][0], sa[k][1], sa[k][2]) * x[n][j];
        const double oj = P.coords[(size_t)n0 * 3 + j];
        const double xlj = P.coords[(size_t)nl * 3 + j] - oj, xrj = P.coords[(size_t)nr * 3 + j] - oj;
        udotx += 0.5 * (P.vel[(size_t)nl * 3 + j] + P.vel[(size_t)nr * 3 + j]) * (xrj - xlj);
        dqL += (cip - xlj) * P.gjq[(size_t)nl * 3 + j];
        dqR += (xrj - cip) * P.gjq[(size_t)nr * 3 + j];
      }
      const double pecfac = urow_peclet(P, fabs(udotx) / (diffIp + 1.0e-16)), om_pecfac = 1.0 - pecfac;
      double limitL = 1.0, limitR = 1.0;
      if (P.upw_options & NALU_OPT_USE_LIMITER) {
        const double dq = qcr - qcl;
        limitL = urow_van_leer(2.0 * 2.0 * dqL - dq, dq);
        limitR = urow_van_leer(2.0 * 2.0 * dqR - dq, dq);
      }
      // aflux = mdot (pecfac qUpwind + om_pecfac qCds): the part of the two extrapolated end values is known here, the part
      // that multiplies the interpolated q_ip has the central form sum_c (N_c advc) q_c and is taken in the column loop
      const double qLk = qcl + dqL * P.ho_upwind * limitL, qRk = qcr - dqR * P.ho_upwind * limitR;
      res -= tm[k] * (pecfac * P.alpha_upw * (tmdot > 0 ? qLk : qRk) + om_pecfac * P.alpha * 0.5 * (qLk + qRk));
      advc[k] = tm[k] * (pecfac * (1.0 - P.alpha_upw) + om_pecfac * (1.0 - P.alpha));
      const double cds = 0.5 * P.alpha * om_pecfac * tmdot;
      aL[k] = sg[k] * (0.5 * (tmdot + fabs(tmdot)) * pecfac * P.alpha_upw + cds);
      aR[k] = sg[k] * (0.5 * (tmdot - fabs(tmdot)) * pecfac * P.alpha_upw + cds);
    }
  }
  // lumped ScalarMassElemKernel: every quantity is the row node's own, times the sub-control volume of corner r
  double mdiag = 0.0;
  if (P.mass) {
    box5::Coef C;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int m = 0; m < 8; ++m) C.X[a][m] = x[box::node_of_bits(m & 1, (m >> 1) & 1, m >> 2)][a];
      box::butterfly8(C.X[a]);
    }
    double vol[16], scV = 0.0;
    box5::scv_volumes_part<0>(C, vol);
    box5::scv_volumes_part<1>(C, vol);
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n == r) scV = vol[n] + vol[8 + n];
    const int nd = nodes[r];
    const double rp = P.rho[nd];
    mdiag = P.g1_dt * rp * scV;
    res -= (P.g1_dt * rp * qc[r] + P.g2_dt * P.rho_n[nd] * P.q_n[nd] + P.g3_dt * P.rho_nm1[nd] * P.q_nm1[nd]) * scV;
  }
  // SST source of the sub-control volume of corner r (TurbKineticEnergySSTSrcElemKernel.C:95-174,
  // SpecificDissipationRateSSTSrcElemKernel.C:100-204): its nearest node is r, so it belongs to this row
  double sdiag[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (SRC) {
    box5::Coef C;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int m = 0; m < 8; ++m) C.X[a][m] = x[box::node_of_bits(m & 1, (m >> 1) & 1, m >> 2)][a];
      box::butterfly8(C.X[a]);
    }
    double vol[16], scV = 0.0;
    box5::scv_volumes_part<0>(C, vol);
    box5::scv_volumes_part<1>(C, vol);
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n == r) scV = vol[n] + vol[8 + n];
    // generic_grad_op at the SCV point of corner r
    const double qg = P.src_shifted ? 0.5 : 0.25, qs = P.src_lumped ? 0.5 : 0.25;
    const double g1 = hex8::sgn_x(r) * qg, g2 = hex8::sgn_y(r) * qg, g3 = hex8::sgn_z(r) * qg;
    const double v1 = hex8::sgn_x(r) * qs, v2 = hex8::sgn_y(r) * qs, v3 = hex8::sgn_z(r) * qs;
    double dndx[8][3];
    const double det = hex8::grad_op_ip(x, g1, g2, g3, dndx);
    if (!(det > 0.0)) *P.err = 1;
    double dudx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, dk[3] = {0, 0, 0}, dw[3] = {0, 0, 0};
    double rho = 0.0, tke = 0.0, sdr = 0.0, tvisc = 0.0, f1 = 0.0, N[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int nd = nodes[c];
      N[c] = hex8::shape_fcn(c, v1, v2, v3);
      const double kc = P.tke[nd], wc = P.sdr[nd];
      rho += N[c] * P.rho[nd];
      tke += N[c] * kc;
      sdr += N[c] * wc;
      tvisc += N[c] * P.tvisc[nd];
      if (P.src == 2) f1 += N[c] * P.f1[nd];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double ui = P.src_vel[(size_t)nd * 3 + i];
        dk[i] += dndx[c][i] * kc;
        dw[i] += dndx[c][i] * wc;
#pragma unroll
        for (int j = 0; j < 3; ++j) dudx[i][j] += dndx[c][j] * ui;
      }
    }
    double Pk = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Pk += dudx[i][j] * (dudx[i][j] + dudx[j][i]);
    Pk *= tvisc;
    const double betaStar = P.turb[NALU_TM_BETA_STAR], lim = P.turb[NALU_TM_TKE_PROD_LIMIT_RATIO];
    double f, rs;
    if (P.src == 1) {
      const double tkeFac = betaStar * rho * sdr, Dk = tkeFac * tke;
      Pk = fmin(Pk, lim * Dk);
      rs = (Pk - Dk) * scV;
      f = tkeFac * scV;
    } else {
      const double Dk = betaStar * rho * sdr * tke;
      Pk = fmin(Pk, lim * Dk);
      const double om = 1.0 - f1;
      const double beta = f1 * P.turb[NALU_TM_BETA_ONE] + om * P.turb[NALU_TM_BETA_TWO];
      const double gamma = f1 * P.turb[NALU_TM_GAMMA_ONE] + om * P.turb[NALU_TM_GAMMA_TWO];
      const double sigmaD = 2.0 * om * P.turb[NALU_TM_SIGMA_W_TWO];
      const double Pw = gamma * rho * Pk / fmax(tvisc, 1.0e-16);
      const double Dw = beta * rho * sdr * sdr;
      const double Sw = sigmaD * rho * (dk[0] * dw[0] + dk[1] * dw[1] + dk[2] * dw[2]) / sdr;
      rs = (Pw - Dw + Sw) * scV;
      f = (2.0 * beta * rho * sdr + fmax(Sw / sdr, 0.0)) * scV;
    }
    res += rs;
#pragma unroll
    for (int c = 0; c < 8; ++c) sdiag[c] = N[c] * f;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    double v = SRC ? sdiag[c] : 0.0, vd = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double d[3];
      hex8::shape_deriv(c, s[k][0], s[k][1], s[k][2], d);
      vd += hex8::shape_fcn(c, sa[k][0], sa[k][1], sa[k][2]) * advc[k] + w[k][0] * d[0] + w[k][1] * d[1] + w[k][2] * d[2];
      if (UPW) {
        // the upwind factors sit on the two ends of the edge: r itself and its neighbour in direction k (matrix only: their
        // flux was taken above)
        if (c == r) v += sg[k] > 0 ? aL[k] : aR[k];
        if (c == nbr[k]) v += sg[k] > 0 ? aR[k] : aL[k];
      }
    }
    res -= vd * qc[c];
    v += vd;
    if (c == r) v += mdiag;
    if (alias) row0[slots[c]] += v; // an element that reaches a node twice: entries read ahead may be stale
    else row0[slots[c]] = old[c] + v;
  }
}

// ---- ContinuityAdvElemKernel (src/kernel/ContinuityAdvElemKernel.C:86-148), row-owner form.  Per point of node r:
// lhs(r, c) += sgn (-grad N_c . A) with the (possibly shifted) lhs operator; rhs_r -= sgn mdot / tau,
// mdot = (rho u - tau (grad p - Gp))_ip . A.  grad N_c . A = w . dN_c/dxi with w_a = sum_i adj(J)[i][a] A_i / det J.
__device__ __forceinline__ void metric_dot_area(const double (&x)[8][3], int ip, double q, const double (&A)[3], double (&w)[3],
                                                double (&s)[3], int *err)
{
  hex8::scs_intg_loc(ip, q, s[0], s[1], s[2]);
  double jact[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    double d[3];
    hex8::shape_deriv(n, s[0], s[1], s[2], d);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) jact[i][j] += d[j] * x[n][i];
  }
  double adj[3][3];
  adj[0][0] = jact[1][1] * jact[2][2] - jact[2][1] * jact[1][2];
  adj[0][1] = jact[1][2] * jact[2][0] - jact[2][2] * jact[1][0];
  adj[0][2] = jact[1][0] * jact[2][1] - jact[2][0] * jact[1][1];
  adj[1][0] = jact[0][2] * jact[2][1] - jact[2][2] * jact[0][1];
  adj[1][1] = jact[0][0] * jact[2][2] - jact[2][0] * jact[0][2];
  adj[1][2] = jact[0][1] * jact[2][0] - jact[2][1] * jact[0][0];
  adj[2][0] = jact[0][1] * jact[1][2] - jact[1][1] * jact[0][2];
  adj[2][1] = jact[0][2] * jact[1][0] - jact[1][2] * jact[0][0];
  adj[2][2] = jact[0][0] * jact[1][1] - jact[1][0] * jact[0][1];
  const double det = jact[0][0] * adj[0][0] + jact[1][0] * adj[1][0] + jact[2][0] * adj[2][0];
  if (!(det > 0.0)) *err = 1;
  const double inv = 1.0 / det;
#pragma unroll
  for (int a = 0; a < 3; ++a) w[a] = (adj[0][a] * A[0] + adj[1][a] * A[1] + adj[2][a] * A[2]) * inv;
}

__device__ __forceinline__ void row_contribution_cont(const UrowParams &P, long long e, int r, double &res)
{
  int nodes[8];
  double x[8][3];
#pragma unroll
  for (int n = 0; n < 8; ++n) nodes[n] = P.elem_nodes[e * 8 + n];
  {
    const double o0 = P.coords[(size_t)nodes[0] * 3], o1 = P.coords[(size_t)nodes[0] * 3 + 1], o2 = P.coords[(size_t)nodes[0] * 3 + 2];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      x[n][0] = P.coords[(size_t)nodes[n] * 3] - o0;
      x[n][1] = P.coords[(size_t)nodes[n] * 3 + 1] - o1;
      x[n][2] = P.coords[(size_t)nodes[n] * 3 + 2] - o2;
    }
  }
  const int rn = P.node_row[nodes[r]];
  double *const row0 = P.values + P.browptr[rn];
  const Slots slots = load_slots(P.elem_slot, e, r);
  double old[8];
  bool alias = false;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    old[c] = row0[slots[c]];
#pragma unroll
    for (int b = 0; b < c; ++b) alias = alias || slots[b] == slots[c];
  }
  const bool shift_poisson = (P.options & NALU_OPT_SHIFT_POISSON) != 0;
  const bool lhs_shifted = shift_poisson || (P.options & NALU_OPT_REDUCED_SENS) != 0;
  const double qs = (P.options & NALU_OPT_SHIFT_MDOT) ? 0.5 : 0.25;
  double As[3][3], wr[3][3], wl[3][3], sr[3][3], sl[3][3], ss[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int ip = box::ip_of(r, k);
    const double sgn = hex8::scs_left(ip) == r ? 1.0 : -1.0;
    double A[3];
    hex8::scs_areav_ip(x, ip, A);
#pragma unroll
    for (int j = 0; j < 3; ++j) As[k][j] = sgn * A[j];
    metric_dot_area(x, ip, shift_poisson ? 0.5 : 0.25, As[k], wr[k], sr[k], P.err);
    if (lhs_shifted != shift_poisson) {
      metric_dot_area(x, ip, 0.5, As[k], wl[k], sl[k], P.err);
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) wl[k][a] = wr[k][a], sl[k][a] = sr[k][a];
    }
    hex8::scs_intg_loc(ip, qs, ss[k][0], ss[k][1], ss[k][2]);
  }
  double mdot = 0.0; // sum over the three points of sgn mdot
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int nd = nodes[c];
    const double p = P.pressure[nd], rho = P.rho[nd];
    const double u[3] = {P.vel[(size_t)nd * 3], P.vel[(size_t)nd * 3 + 1], P.vel[(size_t)nd * 3 + 2]};
    const double gp[3] = {P.dpdx[(size_t)nd * 3], P.dpdx[(size_t)nd * 3 + 1], P.dpdx[(size_t)nd * 3 + 2]};
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double d[3];
      hex8::shape_deriv(c, sl[k][0], sl[k][1], sl[k][2], d);
      v -= wl[k][0] * d[0] + wl[k][1] * d[1] + wl[k][2] * d[2];
      hex8::shape_deriv(c, sr[k][0], sr[k][1], sr[k][2], d);
      const double gpa = wr[k][0] * d[0] + wr[k][1] * d[1] + wr[k][2] * d[2]; // grad N_c . A
      const double N = hex8::shape_fcn(c, ss[k][0], ss[k][1], ss[k][2]);
      mdot += N * (rho * (u[0] * As[k][0] + u[1] * As[k][1] + u[2] * As[k][2]) + P.tau * (gp[0] * As[k][0] + gp[1] * As[k][1] + gp[2] * As[k][2])) -
              P.tau * gpa * p;
    }
    if (alias) row0[slots[c]] += v;
    else row0[slots[c]] = old[c] + v;
  }
  res -= mdot / P.tau;
}

template <int KIND>
__global__ void __launch_bounds__(128, 2) assemble_urow_owner_onedof_kernel(const __grid_constant__ UrowParams P)
{
  const long long rn = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (rn >= P.n_rows) return;
  double res = 0.0;
  const long long q1 = P.rowelem_ptr[rn + 1];
#pragma unroll 1
  for (long long q = P.rowelem_ptr[rn]; q < q1; ++q) {
    const int er = P.rowelem[q];
    if (KIND == 0) row_contribution_scalar<false, false>(P, er >> 3, er & 7, res);
    else if (KIND == 2) row_contribution_scalar<true, false>(P, er >> 3, er & 7, res);
    else if (KIND == 3) row_contribution_scalar<true, true>(P, er >> 3, er & 7, res);
    else if (KIND == 4) row_contribution_scalar<false, true>(P, er >> 3, er & 7, res);
    else row_contribution_cont(P, er >> 3, er & 7, res);
  }
  P.rhs[rn] += res;
}

} // namespace urow

namespace {
bool classify_urow(const nalu_b200_kernel_desc *k, int nk, int &ia, int &im, int &ib)
{
  ia = im = ib = -1;
  for (int i = 0; i < nk; ++i) {
    if (k[i].kind == NALU_K_MOMENTUM_ADV_DIFF || k[i].kind == NALU_K_MOMENTUM_UPW_ADV_DIFF) {
      if (ia >= 0) return false;
      ia = i;
    } else if (k[i].kind == NALU_K_MOMENTUM_MASS) {
      if (im >= 0 || !(k[i].options & NALU_OPT_LUMPED_MASS)) return false;
      im = i;
    } else if (k[i].kind == NALU_K_MOMENTUM_BUOYANCY_SRC || k[i].kind == NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC) {
      if (ib >= 0) return false;
      ib = i;
    } else {
      return false;
    }
  }
  return ia >= 0;
}
// the transposed connectivity can be (or has been) built for this graph
bool rowelem_available(const Graph &g)
{
  return g.d_rowelem || (g.n_elems > 0 && g.n_elems < (int64_t(1) << 28) && (int64_t)g.elem_nodes.size() == 8 * g.n_elems);
}
} // namespace

// row node -> (element, local node) pairs, elements in ascending order: a counting sort of the connectivity by row.  Built
// with the graph for non-lattice meshes (capi.cu) and on first use for lattice meshes whose kernel list the tiled kernels do
// not take; the handle is logically const (a cache).
static int ensure_rowelem(const Graph &cg)
{
  if (cg.d_rowelem) return NALU_B200_OK;
  Graph &g = const_cast<Graph &>(cg);
  const int nrow = g.n_row_nodes();
  std::vector<int64_t> ptr(size_t(nrow) + 1, 0);
  for (size_t q = 0; q < g.elem_nodes.size(); ++q) ++ptr[size_t(g.node_row[g.elem_nodes[q]]) + 1];
  for (int r = 0; r < nrow; ++r) ptr[r + 1] += ptr[r];
  std::vector<int32_t> re(g.elem_nodes.size());
  std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
  for (size_t q = 0; q < g.elem_nodes.size(); ++q) re[size_t(pos[g.node_row[g.elem_nodes[q]]]++)] = int32_t(q); // q = e * 8 + local
  NALU_CUDA(cudaMalloc((void **)&g.d_rowelem_ptr, ptr.size() * sizeof(int64_t)));
  NALU_CUDA(cudaMemcpy(g.d_rowelem_ptr, ptr.data(), ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
  NALU_CUDA(cudaMalloc((void **)&g.d_rowelem, re.size() * sizeof(int32_t)));
  NALU_CUDA(cudaMemcpy(g.d_rowelem, re.data(), re.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return NALU_B200_OK;
}

// one-dof lists of the row-owner kernel: ContinuityAdv alone, or [ScalarMass(lumped)]? + {ScalarAdvDiff, ScalarDiff, ScalarUpwAdvDiff}
static bool classify_urow_scalar(const nalu_b200_kernel_desc *k, int nk, int &ia, int &im, int &is)
{
  ia = im = is = -1;
  for (int i = 0; i < nk; ++i) {
    const int kind = k[i].kind;
    if (kind == NALU_K_TKE_SST_SRC || kind == NALU_K_SDR_SST_SRC) {
      if (is >= 0) return false;
      is = i;
      continue;
    }
    if (kind == NALU_K_SCALAR_ADV_DIFF || kind == NALU_K_SCALAR_DIFF || kind == NALU_K_SCALAR_UPW_ADV_DIFF) {
      if (ia >= 0) return false;
      ia = i;
    } else if (kind == NALU_K_SCALAR_MASS) {
      if (im >= 0 || !(k[i].options & NALU_OPT_LUMPED_MASS)) return false;
      im = i;
    } else {
      return false;
    }
  }
  return ia >= 0;
}
static bool urow_scalar_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (g.ndof != 1 || !rowelem_available(g) || !g.d_elem_slot || !tuning().urow_owner || nk < 1 || nk > 3) return false;
  if (nk == 1 && k[0].kind == NALU_K_CONTINUITY_ADV)
    return k[0].nodal[NALU_F_COORDINATES] && k[0].nodal[NALU_F_VELOCITY] && k[0].nodal[NALU_F_DENSITY] && k[0].nodal[NALU_F_PRESSURE] &&
           k[0].nodal[NALU_F_DPDX];
  int ia, im, is;
  if (!classify_urow_scalar(k, nk, ia, im, is)) return false;
  const nalu_b200_kernel_desc &a = k[ia];
  if (!a.nodal[NALU_F_COORDINATES] || !a.nodal[NALU_F_SCALARQ] || !a.nodal[NALU_F_VISCOSITY]) return false;
  if (is >= 0) {
    // the source's own scalar is the list's dof (one field, one pointer, as in Nalu)
    const nalu_b200_kernel_desc &sk = k[is];
    const bool tke = sk.kind == NALU_K_TKE_SST_SRC;
    if (sk.nodal[NALU_F_COORDINATES] != a.nodal[NALU_F_COORDINATES] || sk.nodal[tke ? NALU_F_TKE : NALU_F_SDR] != a.nodal[NALU_F_SCALARQ])
      return false;
    if (!sk.nodal[tke ? NALU_F_SDR : NALU_F_TKE] || !sk.nodal[NALU_F_VELOCITY] || !sk.nodal[NALU_F_DENSITY] || !sk.nodal[NALU_F_TVISC])
      return false;
    if (!tke && !sk.nodal[NALU_F_FONE_BLEND]) return false;
    if (a.kind == NALU_K_SCALAR_UPW_ADV_DIFF && sk.nodal[NALU_F_DENSITY] != a.nodal[NALU_F_DENSITY]) return false;
    if (im >= 0 && k[im].nodal[NALU_F_DENSITY] != sk.nodal[NALU_F_DENSITY]) return false;
  }
  if (a.kind != NALU_K_SCALAR_DIFF && !a.mass_flow_rate) return false;
  const double *rho = nullptr;
  if (a.kind == NALU_K_SCALAR_UPW_ADV_DIFF) {
    rho = a.nodal[NALU_F_DENSITY];
    if (!rho || !a.nodal[NALU_F_GJQ] || !(a.nodal[NALU_F_VELOCITY_RTM] || a.nodal[NALU_F_VELOCITY])) return false;
  }
  if (im >= 0) {
    const nalu_b200_kernel_desc &m = k[im];
    if (m.nodal[NALU_F_COORDINATES] != a.nodal[NALU_F_COORDINATES] || m.nodal[NALU_F_SCALARQ] != a.nodal[NALU_F_SCALARQ]) return false;
    for (int f : {NALU_F_SCALARQ_N, NALU_F_SCALARQ_NM1, NALU_F_DENSITY, NALU_F_DENSITY_N, NALU_F_DENSITY_NM1})
      if (!m.nodal[f]) return false;
    if (rho && m.nodal[NALU_F_DENSITY] != rho) return false;
  }
  return true;
}

bool urow_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (urow_scalar_supported(g, k, nk)) return true;
  const bool owner = tuning().urow_owner && rowelem_available(g);
  if (g.ndof != 3 || (!g.d_color_elems && !owner) || !g.d_elem_slot || nk < 1 || nk > 3) return false;
  int ia, im, ib;
  if (!classify_urow(k, nk, ia, im, ib)) return false;
  const nalu_b200_kernel_desc &a = k[ia];
  if (!a.nodal[NALU_F_COORDINATES] || !a.nodal[NALU_F_VELOCITY] || !a.nodal[NALU_F_VISCOSITY] || !a.mass_flow_rate) return false;
  // the upwind kernel and the buoyancy sources exist in the row-owner form only
  if ((a.kind == NALU_K_MOMENTUM_UPW_ADV_DIFF || ib >= 0) && !owner) return false;
  if (a.kind == NALU_K_MOMENTUM_UPW_ADV_DIFF && (!a.nodal[NALU_F_DENSITY] || !a.nodal[NALU_F_GJU])) return false;
  if (im >= 0) {
    const nalu_b200_kernel_desc &m = k[im];
    if (m.nodal[NALU_F_COORDINATES] != a.nodal[NALU_F_COORDINATES] || m.nodal[NALU_F_VELOCITY] != a.nodal[NALU_F_VELOCITY]) return false;
    for (int f : {NALU_F_VELOCITY_N, NALU_F_VELOCITY_NM1, NALU_F_DENSITY, NALU_F_DENSITY_N, NALU_F_DENSITY_NM1, NALU_F_DPDX})
      if (!m.nodal[f]) return false;
    if (a.kind == NALU_K_MOMENTUM_UPW_ADV_DIFF && m.nodal[NALU_F_DENSITY] != a.nodal[NALU_F_DENSITY]) return false;
  }
  if (ib >= 0) {
    const nalu_b200_kernel_desc &b = k[ib];
    if (b.nodal[NALU_F_COORDINATES] != a.nodal[NALU_F_COORDINATES]) return false;
    if (!b.nodal[b.kind == NALU_K_MOMENTUM_BUOYANCY_SRC ? NALU_F_DENSITY : NALU_F_TEMPERATURE]) return false;
  }
  return true;
}

int launch_assemble_urow(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs, cudaStream_t stream)
{
  if (!urow_supported(g, k, nk)) {
    set_error("assemble_urow: unsupported mesh / kernel list");
    return NALU_B200_ERR_ARG;
  }
  if (g.n_elems == 0) return NALU_B200_OK;
  if (tuning().urow_owner && rowelem_available(g)) {
    const int rc = ensure_rowelem(g);
    if (rc != NALU_B200_OK) return rc;
  }
  if (urow_scalar_supported(g, k, nk)) {
    urow::UrowParams P = {};
    P.elem_nodes = g.d_elem_nodes;
    P.node_row = g.d_node_row;
    P.browptr = (const long long *)g.d_browptr;
    P.elem_slot = g.d_elem_slot;
    P.values = d_values;
    P.rhs = d_rhs;
    P.err = g.d_err;
    P.coords = k[0].nodal[NALU_F_COORDINATES];
    P.n_rows = g.n_row_nodes();
    P.rowelem_ptr = (const long long *)g.d_rowelem_ptr;
    P.rowelem = g.d_rowelem;
    const unsigned grid = (unsigned)((P.n_rows + 127) / 128);
    if (nk == 1 && k[0].kind == NALU_K_CONTINUITY_ADV) {
      P.vel = k[0].nodal[NALU_F_VELOCITY];
      P.rho = k[0].nodal[NALU_F_DENSITY];
      P.pressure = k[0].nodal[NALU_F_PRESSURE];
      P.dpdx = k[0].nodal[NALU_F_DPDX];
      P.tau = k[0].dt / k[0].gamma1;
      P.options = k[0].options & (NALU_OPT_SHIFT_MDOT | NALU_OPT_SHIFT_POISSON | NALU_OPT_REDUCED_SENS);
      urow::assemble_urow_owner_onedof_kernel<1><<<grid, 128, 0, stream>>>(P);
    } else {
      int ia, im, is;
      classify_urow_scalar(k, nk, ia, im, is);
      const nalu_b200_kernel_desc &a = k[ia];
      if (is >= 0) {
        const nalu_b200_kernel_desc &sk = k[is];
        P.src = sk.kind == NALU_K_TKE_SST_SRC ? 1 : 2;
        P.src_lumped = (sk.options & NALU_OPT_LUMPED_MASS) != 0;
        P.src_shifted = (sk.options & NALU_OPT_SHIFTED_GRAD_OP) != 0;
        P.tke = sk.nodal[NALU_F_TKE];
        P.sdr = sk.nodal[NALU_F_SDR];
        P.tvisc = sk.nodal[NALU_F_TVISC];
        P.f1 = sk.nodal[NALU_F_FONE_BLEND];
        P.src_vel = sk.nodal[NALU_F_VELOCITY];
        P.rho = sk.nodal[NALU_F_DENSITY];
        for (int i = 0; i < 10; ++i) P.turb[i] = sk.turb[i];
      }
      P.q = a.nodal[NALU_F_SCALARQ];
      P.visc = a.nodal[NALU_F_VISCOSITY];
      P.mdot = a.kind == NALU_K_SCALAR_DIFF ? nullptr : a.mass_flow_rate;
      P.options = a.options & (NALU_OPT_SHIFTED_GRAD_OP | NALU_OPT_SKEW_SYMMETRIC);
      if (a.kind == NALU_K_SCALAR_UPW_ADV_DIFF) {
        P.upw = 1;
        P.upw_options = a.options & (NALU_OPT_USE_LIMITER | NALU_OPT_PECLET_TANH);
        P.vel = a.nodal[NALU_F_VELOCITY_RTM] ? a.nodal[NALU_F_VELOCITY_RTM] : a.nodal[NALU_F_VELOCITY];
        P.rho = a.nodal[NALU_F_DENSITY];
        P.gjq = a.nodal[NALU_F_GJQ];
        P.alpha = a.upw[NALU_UPW_ALPHA];
        P.alpha_upw = a.upw[NALU_UPW_ALPHA_UPW];
        P.ho_upwind = a.upw[NALU_UPW_HO_UPWIND];
        P.hybrid = a.upw[NALU_UPW_HYBRID_FACTOR];
        P.tanh_trans = a.upw[NALU_UPW_TANH_TRANS];
        P.tanh_width = a.upw[NALU_UPW_TANH_WIDTH];
      }
      if (im >= 0) {
        const nalu_b200_kernel_desc &m = k[im];
        P.mass = 1;
        P.rho = m.nodal[NALU_F_DENSITY];
        P.rho_n = m.nodal[NALU_F_DENSITY_N];
        P.rho_nm1 = m.nodal[NALU_F_DENSITY_NM1];
        P.q_n = m.nodal[NALU_F_SCALARQ_N];
        P.q_nm1 = m.nodal[NALU_F_SCALARQ_NM1];
        P.g1_dt = m.gamma1 / m.dt;
        P.g2_dt = m.gamma2 / m.dt;
        P.g3_dt = m.gamma3 / m.dt;
      }
      if (P.upw && P.src) urow::assemble_urow_owner_onedof_kernel<3><<<grid, 128, 0, stream>>>(P);
      else if (P.upw) urow::assemble_urow_owner_onedof_kernel<2><<<grid, 128, 0, stream>>>(P);
      else if (P.src) urow::assemble_urow_owner_onedof_kernel<4><<<grid, 128, 0, stream>>>(P);
      else urow::assemble_urow_owner_onedof_kernel<0><<<grid, 128, 0, stream>>>(P);
    }
    count_launch();
    NALU_CUDA(cudaGetLastError());
    return NALU_B200_OK;
  }
  int ia, im, ib;
  classify_urow(k, nk, ia, im, ib);
  const nalu_b200_kernel_desc &a = k[ia];
  urow::UrowParams P = {};
  P.elem_nodes = g.d_elem_nodes;
  P.node_row = g.d_node_row;
  P.browptr = (const long long *)g.d_browptr;
  P.elem_slot = g.d_elem_slot;
  P.values = d_values;
  P.rhs = d_rhs;
  P.err = g.d_err;
  P.coords = a.nodal[NALU_F_COORDINATES];
  P.vel = a.nodal[NALU_F_VELOCITY];
  P.visc = a.nodal[NALU_F_VISCOSITY];
  P.mdot = a.mass_flow_rate;
  P.options = a.options & (NALU_OPT_SHIFTED_GRAD_OP | NALU_OPT_SKEW_SYMMETRIC);
  P.include_divu = a.include_divu;
  if (a.kind == NALU_K_MOMENTUM_UPW_ADV_DIFF) {
    P.upw = 1;
    P.upw_options = a.options & (NALU_OPT_USE_LIMITER | NALU_OPT_PECLET_TANH);
    P.vrtm = a.nodal[NALU_F_VELOCITY_RTM] ? a.nodal[NALU_F_VELOCITY_RTM] : a.nodal[NALU_F_VELOCITY];
    P.rho = a.nodal[NALU_F_DENSITY];
    P.gju = a.nodal[NALU_F_GJU];
    P.alpha = a.upw[NALU_UPW_ALPHA];
    P.alpha_upw = a.upw[NALU_UPW_ALPHA_UPW];
    P.ho_upwind = a.upw[NALU_UPW_HO_UPWIND];
    P.hybrid = a.upw[NALU_UPW_HYBRID_FACTOR];
    P.tanh_trans = a.upw[NALU_UPW_TANH_TRANS];
    P.tanh_width = a.upw[NALU_UPW_TANH_WIDTH];
  }
  if (ib >= 0) {
    const nalu_b200_kernel_desc &b = k[ib];
    P.buoy = b.kind == NALU_K_MOMENTUM_BUOYANCY_SRC ? 1 : 2;
    P.bfield = b.nodal[P.buoy == 1 ? NALU_F_DENSITY : NALU_F_TEMPERATURE];
    for (int i = 0; i < 3; ++i) P.gravity[i] = b.gravity[i];
    P.rho_ref = b.rho_ref;
    P.t_ref = b.t_ref;
    P.beta = b.beta;
  }
  if (im >= 0) {
    const nalu_b200_kernel_desc &m = k[im];
    P.mass = 1;
    P.vel_n = m.nodal[NALU_F_VELOCITY_N];
    P.vel_nm1 = m.nodal[NALU_F_VELOCITY_NM1];
    P.rho = m.nodal[NALU_F_DENSITY];
    P.rho_n = m.nodal[NALU_F_DENSITY_N];
    P.rho_nm1 = m.nodal[NALU_F_DENSITY_NM1];
    P.dpdx = m.nodal[NALU_F_DPDX];
    P.g1_dt = m.gamma1 / m.dt;
    P.g2_dt = m.gamma2 / m.dt;
    P.g3_dt = m.gamma3 / m.dt;
    P.dens_fac = m.dens_fac;
    P.om_dens_fac = m.om_dens_fac;
  }
  const int minb = tuning().urow_minb;
  if (g.d_rowelem && tuning().urow_owner) {
    P.n_rows = g.n_row_nodes();
    P.rowelem_ptr = (const long long *)g.d_rowelem_ptr;
    P.rowelem = g.d_rowelem;
    const unsigned grid = (unsigned)((P.n_rows + 127) / 128);
    if (P.upw) urow::assemble_urow_owner_kernel<2, 2><<<grid, 128, 0, stream>>>(P);
    else if (P.buoy) urow::assemble_urow_owner_kernel<2, 1><<<grid, 128, 0, stream>>>(P);
    else if (minb >= 4) urow::assemble_urow_owner_kernel<4, 0><<<grid, 128, 0, stream>>>(P);
    else if (minb == 3) urow::assemble_urow_owner_kernel<3, 0><<<grid, 128, 0, stream>>>(P);
    else urow::assemble_urow_owner_kernel<2, 0><<<grid, 128, 0, stream>>>(P);
    count_launch();
    NALU_CUDA(cudaGetLastError());
    return NALU_B200_OK;
  }
  const long long ncolor = (long long)g.color_ptr.size() - 1;
  for (long long c = 0; c < ncolor; ++c) {
    P.elem_list = g.d_color_elems + g.color_ptr[c];
    P.n_elems = g.color_ptr[c + 1] - g.color_ptr[c];
    if (P.n_elems == 0) continue;
    const long long nt = 8 * P.n_elems;
    const unsigned grid = (unsigned)((nt + 127) / 128);
    if (minb >= 4) urow::assemble_urow_kernel<4><<<grid, 128, 0, stream>>>(P);
    else if (minb == 3) urow::assemble_urow_kernel<3><<<grid, 128, 0, stream>>>(P);
    else urow::assemble_urow_kernel<2><<<grid, 128, 0, stream>>>(P);
    count_launch();
  }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

} // namespace nalu_b200
