// face_elem.cu -- face/element boundary-condition assembly on the device.
//
// Replaces AssembleFaceElemSolverAlgorithm::execute (src/AssembleFaceElemSolverAlgorithm.C:78-108; the face loop is
// run_face_elem_algorithm, include/AssembleFaceElemSolverAlgorithm.h:45-133) for Quad4 faces of Hex8 elements with the kernels
//   MomentumSymmetryElemKernel (src/kernel/MomentumSymmetryElemKernel.C:68-163),
//   MomentumOpenAdvDiffElemKernel (src/kernel/MomentumOpenAdvDiffElemKernel.C:110-325),
//   ScalarOpenAdvElemKernel (src/kernel/ScalarOpenAdvElemKernel.C:90-195),
//   ContinuityOpenElemKernel (src/kernel/ContinuityOpenElemKernel.C:80-199),
// and ComputeGeometryBoundaryAlgorithm::execute (src/ComputeGeometryBoundaryAlgorithm.C:62-123).
//
// Decomposition: every face integration point adds to exactly ONE node's rows (its "nearest node", a node of the face),
// and the four points of a face have four different nearest nodes.  So a thread owns (face, integration point): it gathers
// the element, evaluates the face gradient operator at its own point only, builds the ndof x (8 ndof) row block of its
// nearest node in registers (the reference fills a (8 ndof)^2 scratch matrix of which 4 ndof rows are non-zero) and adds
// it to the CSR through the element's slot map.  Faces are coloured when the face set is created (faces of one colour
// share no node), one launch per colour, plain read-modify-writes: deterministic and atomic-free.  O(surface) work.
#include "hex8_dev.cuh"
#include "internal.h"

#include <algorithm>
#include <vector>

struct nalu_b200_face_set {
  int64_t n_faces = 0;
  int32_t n_colors = 0;
  std::vector<int64_t> color_ptr; // [n_colors + 1] into the colour-sorted lists
  int32_t *d_elem = nullptr;      // [n_faces] colour-sorted: local element
  int32_t *d_ord = nullptr;       //           face ordinal
  int32_t *d_face = nullptr;      //           index of the face in the caller's order (face fields)
};

namespace nalu_b200 {
namespace face {

// Hex8 side tables.  side_node_ordinals (Hex8CVFEM.C:414-422) doubles as the face ipNodeMap (:381-393);
// opposingNodes: Hex8CVFEM.C:288-300.
__constant__ int c_side_nodes[24] = {0, 1, 5, 4, 1, 2, 6, 5, 2, 3, 7, 6, 0, 4, 7, 3, 0, 3, 2, 1, 4, 5, 6, 7};
__constant__ int c_opp_node[24] = {3, 2, 6, 7, 0, 3, 7, 4, 1, 0, 4, 5, 1, 5, 6, 2, 4, 7, 6, 5, 0, 1, 2, 3};
static const int h_side_nodes[24] = {0, 1, 5, 4, 1, 2, 6, 5, 2, 3, 7, 6, 0, 4, 7, 3, 0, 3, 2, 1, 4, 5, 6, 7};

// Parametric location of face integration point ip of face `ord`: the face's node `ip` pulled half-way to the face centre
// in the two in-face directions (intgExpFace_, Hex8CVFEM.C:348-378), or the node itself when shifted (:424-441).
__device__ __forceinline__ void face_ip_loc(int ord, int ip, bool shifted, double &s1, double &s2, double &s3)
{
  const int n = c_side_nodes[ord * 4 + ip];
  const double q = shifted ? 0.5 : 0.25;
  const int axis = (ord == 1 || ord == 3) ? 0 : ((ord == 0 || ord == 2) ? 1 : 2); // the face's normal direction
  s1 = hex8::sgn_x(n) * (axis == 0 ? 0.5 : q);
  s2 = hex8::sgn_y(n) * (axis == 1 ? 0.5 : q);
  s3 = hex8::sgn_z(n) * (axis == 2 ? 0.5 : q);
}

// Quad3DSCS shape functions at face point ip (Quad43DCVFEM.C:31-43,130-148): bilinear on [-1/2,1/2]^2, the point sits at
// (+-q, +-q) towards node ip -> weight (1/2+q)^2 for node ip, (1/2+q)(1/2-q) for its two edge neighbours, (1/2-q)^2 opposite
__device__ __forceinline__ double quad_shape(int ip, int ic, bool shifted)
{
  const double q = shifted ? 0.5 : 0.25;
  const int d = (ic - ip) & 3;
  const double a = 0.5 + q, b = 0.5 - q;
  return d == 0 ? a * a : (d == 2 ? b * b : a * b);
}

// src/PecletFunction.C:67-71 (classic, A = 5), :98-101 (tanh)
__device__ __forceinline__ double peclet(const nalu_b200_face_kernel_desc &k, double x)
{
  if (k.options & NALU_OPT_PECLET_TANH) return 0.50 * (1.0 + tanh((x - k.upw[NALU_UPW_TANH_TRANS]) / k.upw[NALU_UPW_TANH_WIDTH]));
  const double m = k.upw[NALU_UPW_HYBRID_FACTOR] * x;
  return m * m / (5.0 + m * m);
}

struct FaceCtx {
  int nodes[8];      // element nodes
  int fn[4];         // element-local ordinals of the face's nodes
  int ip, ord;
  int nearest, opp;  // element-local ordinals
  long long face;    // index into the face fields
  double x[8][3];    // element coordinates
};

template <int NDOF>
struct RowBlock {
  double a[NDOF][8 * NDOF];
  double r[NDOF];
};

__device__ __forceinline__ double fnode(const double *f, int node) { return f[node]; }

// ---- MomentumSymmetryElemKernel / the diffusion part of MomentumOpenAdvDiffElemKernel --------------------------------------
// symmetry: F_i = P_il (stress . A)_l with P = n n^T; open: P = I - n n^T; (stress . A)_l = -mu (du_l/dx_j + du_j/dx_l) A_j.
// Per column node ic (c = dndx_ic, ca = c . A) the reference's four "lhsfac" statements per (ic, j, i, l) sum to
//   lhs(i, ic l) += -mu ca P_il          (MomentumSymmetryElemKernel.C:124-127,144-147; MomentumOpenAdvDiffElemKernel.C:286-289,304-307)
//   lhs(i, ic j) += -mu (P c)_i A_j      (:129-132,149-152; :291-294,309-312)
// and the explicit "divU" statements to rhs(i) -= [2/3 mu includeDivU sum_j c_j u_j A_j] W_i with W_i = n_i^2 + sum_{l != i}
// n_i n_l (symmetry) or (1 - n_i^2) + sum_{l != i} n_i n_l (open) -- taken literally from the reference (the weight of the
// open kernel is not the projector's row sum).
__device__ __forceinline__ void viscous_traction(const FaceCtx &F, const double dndx[8][3], const double *A, double visc, double divu,
                                                 const double *vel, bool open, RowBlock<3> &B)
{
  const double asq = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
  const double inv = 1.0 / sqrt(asq);
  const double nx[3] = {A[0] * inv, A[1] * inv, A[2] * inv};
  double P[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int l = 0; l < 3; ++l) P[i][l] = open ? ((i == l ? 1.0 : 0.0) - nx[i] * nx[l]) : nx[i] * nx[l];
  const double nsum = nx[0] + nx[1] + nx[2];
  double W[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) W[i] = (open ? 1.0 - nx[i] * nx[i] : nx[i] * nx[i]) + nx[i] * (nsum - nx[i]);
#pragma unroll 1
  for (int ic = 0; ic < 8; ++ic) {
    const double *c = dndx[ic];
    const double u[3] = {vel[(size_t)F.nodes[ic] * 3 + 0], vel[(size_t)F.nodes[ic] * 3 + 1], vel[(size_t)F.nodes[ic] * 3 + 2]};
    const double ca = c[0] * A[0] + c[1] * A[1] + c[2] * A[2];
    const double divUstress = (2.0 / 3.0) * visc * divu * (c[0] * u[0] * A[0] + c[1] * u[1] * A[1] + c[2] * u[2] * A[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double ri = 0.0;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        // -mu (du_l/dx_j A_j) P_il  and  -mu (du_j/dx_l A_j) P_il -> column (ic, j) gets -mu c_l A_j P_il summed over l
        const double e1 = -visc * ca * P[i][l];
        B.a[i][ic * 3 + l] += e1;
        ri += e1 * u[l];
      }
      double pc = P[i][0] * c[0] + P[i][1] * c[1] + P[i][2] * c[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double e2 = -visc * pc * A[j];
        B.a[i][ic * 3 + j] += e2;
        ri += e2 * u[j];
      }
      B.r[i] -= ri + divUstress * W[i];
    }
  }
}

__device__ void momentum_symmetry(const nalu_b200_face_kernel_desc &k, const FaceCtx &F, RowBlock<3> &B)
{
  double s1, s2, s3, dndx[8][3];
  face_ip_loc(F.ord, F.ip, (k.options & NALU_OPT_SHIFTED_GRAD_OP) != 0, s1, s2, s3);
  hex8::grad_op_ip(F.x, s1, s2, s3, dndx);
  const double *A = k.face[NALU_FF_EXPOSED_AREA_VEC] + 12 * F.face + 3 * F.ip;
  double visc = 0.0;
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) visc += quad_shape(F.ip, ic, false) * k.nodal[NALU_FN_VISCOSITY][F.nodes[F.fn[ic]]];
  viscous_traction(F, dndx, A, visc, k.include_divu, k.nodal[NALU_FN_VELOCITY], false, B);
}

// common to the two open advection kernels: Peclet factor along the edge opposing node -> nearest node
__device__ __forceinline__ double open_pecfac(const nalu_b200_face_kernel_desc &k, const FaceCtx &F, const double *vrtm)
{
  const int nn = F.nodes[F.nearest], on = F.nodes[F.opp];
  double udotx = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) udotx += 0.5 * (F.x[F.nearest][i] - F.x[F.opp][i]) * (vrtm[(size_t)on * 3 + i] + vrtm[(size_t)nn * 3 + i]);
  const double *g = k.nodal[NALU_FN_VISCOSITY], *rho = k.nodal[NALU_FN_DENSITY];
  const double diffIp = 0.5 * (g[on] / rho[on] + g[nn] / rho[nn]);
  return peclet(k, fabs(udotx) / (diffIp + 1.0e-16));
}

__device__ void scalar_open_adv(const nalu_b200_face_kernel_desc &k, const FaceCtx &F, RowBlock<1> &B)
{
  const bool skew = (k.options & NALU_OPT_SKEW_SYMMETRIC) != 0;
  const double alphaUpw = k.upw[NALU_UPW_ALPHA_UPW], hoUpwind = k.upw[NALU_UPW_HO_UPWIND];
  const double *q = k.nodal[NALU_FN_SCALARQ], *qbc = k.nodal[NALU_FN_BC_SCALARQ], *gjq = k.nodal[NALU_FN_GJQ];
  double qIp = 0.0, qEntrain = 0.0, xb[3] = {0, 0, 0}, w[4];
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) {
    w[ic] = quad_shape(F.ip, ic, skew);
    const int n = F.nodes[F.fn[ic]];
    qIp += w[ic] * q[n];
    qEntrain += w[ic] * qbc[n];
#pragma unroll
    for (int j = 0; j < 3; ++j) xb[j] += w[ic] * F.x[F.fn[ic]][j];
  }
  const int nn = F.nodes[F.nearest];
  double dqR = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) dqR += (xb[i] - F.x[F.nearest][i]) * gjq[(size_t)nn * 3 + i] * hoUpwind;
  const double qIpUpw = q[nn] + dqR;
  const double *vrtm = k.nodal[NALU_FN_VELOCITY_RTM] ? k.nodal[NALU_FN_VELOCITY_RTM] : k.nodal[NALU_FN_VELOCITY];
  const double pecfac = open_pecfac(k, F, vrtm), om_pecfac = 1.0 - pecfac;
  const double tmdot = k.face[NALU_FF_OPEN_MASS_FLOW_RATE][4 * F.face + F.ip];
  if (tmdot > 0) {
    // leaving: Peclet-blended upwind / central value of the interior
    const double qUpwind = alphaUpw * qIpUpw + (1.0 - alphaUpw) * qIp;
    B.r[0] -= tmdot * (pecfac * qUpwind + om_pecfac * qIp);
    B.a[0][F.nearest] += tmdot * pecfac * alphaUpw;
    const double fac = tmdot * (pecfac * (1.0 - alphaUpw) + om_pecfac);
#pragma unroll
    for (int ic = 0; ic < 4; ++ic) B.a[0][F.fn[ic]] += w[ic] * fac;
  } else {
    // entraining: advect the specified value in
    B.r[0] -= tmdot * qEntrain;
  }
}

__device__ void continuity_open(const nalu_b200_face_kernel_desc &k, const FaceCtx &F, RowBlock<1> &B)
{
  const bool shifted = (k.options & NALU_OPT_SHIFTED_GRAD_OP) != 0, reduced = (k.options & NALU_OPT_REDUCED_SENS) != 0;
  const double penaltyFac = 2.0; // include/kernel/ContinuityOpenElemKernel.h:67
  const double tau = k.dt / k.gamma1;
  double s1, s2, s3, dndx[8][3], dndx_lhs[8][3];
  face_ip_loc(F.ord, F.ip, shifted, s1, s2, s3);
  hex8::grad_op_ip(F.x, s1, s2, s3, dndx);
  const bool lhs_shifted = shifted || reduced;
  if (lhs_shifted != shifted) {
    face_ip_loc(F.ord, F.ip, lhs_shifted, s1, s2, s3);
    hex8::grad_op_ip(F.x, s1, s2, s3, dndx_lhs);
  }
  const double *A = k.face[NALU_FF_EXPOSED_AREA_VEC] + 12 * F.face + 3 * F.ip;
  const double aMag = sqrt(A[0] * A[0] + A[1] * A[1] + A[2] * A[2]);
  double invL = 0.0;
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) invL += dndx[F.fn[ic]][0] * A[0] + dndx[F.fn[ic]][1] * A[1] + dndx[F.fn[ic]][2] * A[2];
  invL /= aMag;
  const double *p = k.nodal[NALU_FN_PRESSURE], *pbc = k.nodal[NALU_FN_PRESSURE_BC], *rho = k.nodal[NALU_FN_DENSITY];
  const double *vrtm = k.nodal[NALU_FN_VELOCITY_RTM] ? k.nodal[NALU_FN_VELOCITY_RTM] : k.nodal[NALU_FN_VELOCITY];
  const double *gp = k.nodal[NALU_FN_DPDX];
  double pBip = 0.0, pbcBip = -k.face[NALU_FF_DYNAMIC_PRESSURE][4 * F.face + F.ip], rhoU[3] = {0, 0, 0}, G[3] = {0, 0, 0}, w[4];
  const bool shift_mdot = (k.options & NALU_OPT_SHIFT_MDOT) != 0;
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) {
    w[ic] = quad_shape(F.ip, ic, shift_mdot);
    const int n = F.nodes[F.fn[ic]];
    pBip += w[ic] * p[n];
    pbcBip += w[ic] * pbc[n];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      rhoU[j] += w[ic] * rho[n] * vrtm[(size_t)n * 3 + j];
      G[j] += w[ic] * gp[(size_t)n * 3 + j];
    }
  }
  double dp[3] = {0, 0, 0};
#pragma unroll
  for (int ic = 0; ic < 8; ++ic) {
    const double pIc = p[F.nodes[ic]];
#pragma unroll
    for (int j = 0; j < 3; ++j) dp[j] += dndx[ic][j] * pIc;
  }
  double mdot = penaltyFac * tau * invL * (pBip - pbcBip) * aMag;
#pragma unroll
  for (int j = 0; j < 3; ++j) mdot += (rhoU[j] - tau * (dp[j] - G[j])) * A[j];
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) B.a[0][F.fn[ic]] += w[ic] * penaltyFac * invL * aMag;
#pragma unroll
  for (int ic = 0; ic < 8; ++ic) {
    const double *c = lhs_shifted != shifted ? dndx_lhs[ic] : dndx[ic];
    B.a[0][ic] -= c[0] * A[0] + c[1] * A[1] + c[2] * A[2];
  }
  B.r[0] -= mdot / tau;
}

__device__ void momentum_open_adv_diff(const nalu_b200_face_kernel_desc &k, const FaceCtx &F, RowBlock<3> &B)
{
  const bool skew = (k.options & NALU_OPT_SKEW_SYMMETRIC) != 0, moves = (k.options & NALU_OPT_MESH_MOVES) != 0;
  const double alphaUpw = k.upw[NALU_UPW_ALPHA_UPW], hoUpwind = k.upw[NALU_UPW_HO_UPWIND];
  const double *vel = k.nodal[NALU_FN_VELOCITY];
  const double *vrtm = moves ? k.nodal[NALU_FN_VELOCITY_RTM] : vel, *vmesh = moves ? k.nodal[NALU_FN_MESH_VELOCITY] : vel;
  const double *rho = k.nodal[NALU_FN_DENSITY], *mu = k.nodal[NALU_FN_VISCOSITY], *ubc = k.nodal[NALU_FN_VELOCITY_BC];
  const double *A = k.face[NALU_FF_EXPOSED_AREA_VEC] + 12 * F.face + 3 * F.ip;
  const double amag = sqrt(A[0] * A[0] + A[1] * A[1] + A[2] * A[2]);
  double rhoBip = 0.0, viscBip = 0.0, uB[3] = {0, 0, 0}, rhoV[3] = {0, 0, 0}, uspec[3] = {0, 0, 0}, xb[3] = {0, 0, 0}, wa[4];
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) {
    const double r = quad_shape(F.ip, ic, false);
    wa[ic] = quad_shape(F.ip, ic, skew);
    const int n = F.nodes[F.fn[ic]];
    rhoBip += r * rho[n];
    viscBip += r * mu[n];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      uspec[j] += wa[ic] * ubc[(size_t)n * 3 + j];
      uB[j] += wa[ic] * vel[(size_t)n * 3 + j];
      xb[j] += wa[ic] * F.x[F.fn[ic]][j];
      rhoV[j] += r * rho[n] * vmesh[(size_t)n * 3 + j];
    }
  }
  const int nn = F.nodes[F.nearest];
  const double *gju = k.nodal[NALU_FN_GJUI] + (size_t)nn * 9;
  double nx[3], uExtrap[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    nx[i] = A[i] / amag;
    double duR = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) duR += (xb[j] - F.x[F.nearest][j]) * gju[3 * i + j] * hoUpwind;
    uExtrap[i] = vel[(size_t)nn * 3 + i] + duR;
  }
  const double pecfac = open_pecfac(k, F, vrtm), om_pecfac = 1.0 - pecfac;
  const double tmdot = k.face[NALU_FF_OPEN_MASS_FLOW_RATE][4 * F.face + F.ip];
  if (tmdot > 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double uiUpwind = alphaUpw * uExtrap[i] + (1.0 - alphaUpw) * uB[i];
      B.r[i] -= tmdot * (pecfac * uiUpwind + om_pecfac * uB[i]);
      B.a[i][F.nearest * 3 + i] += tmdot * pecfac * alphaUpw;
      const double fac = tmdot * (pecfac * (1.0 - alphaUpw) + om_pecfac);
#pragma unroll
      for (int ic = 0; ic < 4; ++ic) B.a[i][F.fn[ic] * 3 + i] += wa[ic] * fac;
    }
  } else {
    // entrainment: normal component from the mass flow (corrected for mesh motion), tangential from the specified velocity
    const double mvc = rhoV[0] * A[0] + rhoV[1] * A[1] + rhoV[2] * A[2];
    const double uEntrain = tmdot / (rhoBip * amag) + mvc / (rhoBip * amag) * (moves ? 1.0 : 0.0);
    const double uspecn = uspec[0] * nx[0] + uspec[1] * nx[1] + uspec[2] * nx[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) B.r[i] -= tmdot * uEntrain * nx[i] + tmdot * (uspec[i] - uspecn * nx[i]);
  }
  double s1, s2, s3, dndx[8][3];
  face_ip_loc(F.ord, F.ip, (k.options & NALU_OPT_SHIFTED_GRAD_OP) != 0, s1, s2, s3);
  hex8::grad_op_ip(F.x, s1, s2, s3, dndx);
  viscous_traction(F, dndx, A, viscBip, k.include_divu, vel, true, B);
}

constexpr int kMaxFaceKernels = 4;
struct FaceParams {
  long long n;          // faces of this colour
  const int *elem, *ord, *face;
  const int *elem_nodes, *node_row;
  const long long *browptr;
  const unsigned char *elem_slot;
  double *values, *rhs;
  int *err;
  int nk;
  nalu_b200_face_kernel_desc k[kMaxFaceKernels];
};

// ---- MomentumWallFunctionElemKernel (src/kernel/MomentumWallFunctionElemKernel.C:73-155) ------------------------------------
// lambda (I - n n^T) acting on the wall-parallel part of (u - u_wall) at the face point, lambda = rho kappa u_tau / ln(E y+) |A|
// in the log layer, mu / y_p |A| below y+_crit; columns = the face's four nodes only
__device__ void momentum_wall_function(const nalu_b200_face_kernel_desc &k, const FaceCtx &F, RowBlock<3> &B)
{
  const double *A = k.face[NALU_FF_EXPOSED_AREA_VEC] + 12 * F.face + 3 * F.ip;
  const double amag = sqrt(A[0] * A[0] + A[1] * A[1] + A[2] * A[2]);
  const double nx[3] = {A[0] / amag, A[1] / amag, A[2] / amag};
  double rho = 0.0, mu = 0.0, u[3] = {0.0, 0.0, 0.0}, ub[3] = {0.0, 0.0, 0.0}, sf[4];
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) {
    const int nd = F.nodes[F.fn[ic]];
    sf[ic] = quad_shape(F.ip, ic, false);
    rho += sf[ic] * k.nodal[NALU_FN_DENSITY][nd];
    mu += sf[ic] * k.nodal[NALU_FN_VISCOSITY][nd];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      u[j] += sf[ic] * k.nodal[NALU_FN_VELOCITY][(size_t)nd * 3 + j];
      ub[j] += sf[ic] * k.nodal[NALU_FN_WALL_VELOCITY_BC][(size_t)nd * 3 + j];
    }
  }
  const double yp = k.face[NALU_FF_WALL_NORMAL_DISTANCE_BIP][4 * F.face + F.ip];
  const double utau = k.face[NALU_FF_WALL_FRICTION_VELOCITY_BIP][4 * F.face + F.ip];
  const double yplus = rho * yp * utau / mu;
  const double lambda = yplus > k.wall[NALU_WALL_YPLUS_CRIT] ? rho * k.wall[NALU_WALL_KAPPA] * utau / log(k.wall[NALU_WALL_ELOG] * yplus) * amag
                                                             : mu / yp * amag;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double utan = 0.0, ubtan = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double pij = (i == j ? 1.0 : 0.0) - nx[i] * nx[j];
      utan += pij * u[j];
      ubtan += pij * ub[j];
#pragma unroll
      for (int ic = 0; ic < 4; ++ic) B.a[i][F.fn[ic] * 3 + j] += lambda * pij * sf[ic];
    }
    B.r[i] -= lambda * (utan - ubtan);
  }
}

template <int NDOF>
__global__ void __launch_bounds__(128) assemble_face_elem_kernel(const __grid_constant__ FaceParams p)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 2) >= p.n) return;
  const long long fi = t >> 2;
  FaceCtx F;
  F.ip = (int)(t & 3);
  F.ord = p.ord[fi];
  F.face = p.face[fi];
  const long long e = p.elem[fi];
  const double *X = p.k[0].nodal[NALU_FN_COORDINATES];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    F.nodes[n] = p.elem_nodes[e * 8 + n];
#pragma unroll
    for (int j = 0; j < 3; ++j) F.x[n][j] = X[(size_t)F.nodes[n] * 3 + j];
  }
#pragma unroll
  for (int ic = 0; ic < 4; ++ic) F.fn[ic] = c_side_nodes[F.ord * 4 + ic];
  F.nearest = F.fn[F.ip];
  F.opp = c_opp_node[F.ord * 4 + F.ip];
  RowBlock<NDOF> B;
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    B.r[i] = 0.0;
#pragma unroll
    for (int c = 0; c < 8 * NDOF; ++c) B.a[i][c] = 0.0;
  }
  for (int i = 0; i < p.nk; ++i) {
    const nalu_b200_face_kernel_desc &k = p.k[i];
    if constexpr (NDOF == 3) {
      if (k.kind == NALU_FK_MOMENTUM_SYMMETRY) momentum_symmetry(k, F, B);
      else if (k.kind == NALU_FK_MOMENTUM_WALL_FUNCTION) momentum_wall_function(k, F, B);
      else momentum_open_adv_diff(k, F, B);
    } else {
      if (k.kind == NALU_FK_SCALAR_OPEN_ADV) scalar_open_adv(k, F, B);
      else continuity_open(k, F, B);
    }
  }
  // apply_coeff -> sumInto (src/TpetraLinearSystem.C:1457-1522) for the one non-zero block row
  const int rn = p.node_row[F.nodes[F.nearest]];
  const long long b0 = p.browptr[rn];
  const int nb = (int)(p.browptr[rn + 1] - b0);
  const unsigned char *slots = p.elem_slot + e * 64 + F.nearest * 8;
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    double *row = p.values + (long long)NDOF * NDOF * b0 + (long long)i * nb * NDOF;
#pragma unroll
    for (int ic = 0; ic < 8; ++ic)
#pragma unroll
      for (int j = 0; j < NDOF; ++j) {
        const double v = B.a[i][ic * NDOF + j];
        if (v != 0.0) row[slots[ic] * NDOF + j] += v;
      }
    p.rhs[(long long)rn * NDOF + i] += B.r[i];
  }
}

__global__ void exposed_area_kernel(long long n, const int *elem, const int *ord, const int *face, const int *elem_nodes, const double *X,
                                    double *area)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((t >> 2) >= n) return;
  const long long fi = t >> 2;
  const int k = (int)(t & 3);
  const long long e = elem[fi];
  double p[4][3], c[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int node = elem_nodes[e * 8 + c_side_nodes[ord[fi] * 4 + i]];
#pragma unroll
    for (int j = 0; j < 3; ++j) p[i][j] = X[(size_t)node * 3 + j];
  }
  // quad3d_scs_det (MasterElementWork.F:1467-1586): A_k = 1/2 (centroid - p_k) x (e_{k-1} - e_k), e_k the mid-point of edge k
  const int kp = (k + 1) & 3, km = (k + 3) & 3;
  double d13[3], d24[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    c[j] = (p[0][j] + p[1][j] + p[2][j] + p[3][j]) * 0.25;
    d13[j] = c[j] - p[k][j];
    d24[j] = (p[km][j] + p[k][j]) * 0.5 - (p[k][j] + p[kp][j]) * 0.5;
  }
  double *a = area + 12 * (long long)face[fi] + 3 * k;
  a[0] = 0.5 * (d24[2] * d13[1] - d13[2] * d24[1]);
  a[1] = 0.5 * (d24[0] * d13[2] - d13[0] * d24[2]);
  a[2] = 0.5 * (d24[1] * d13[0] - d13[1] * d24[0]);
}

const char *missing_face_field(const nalu_b200_face_kernel_desc &k)
{
  auto need_n = [&](int slot, const char *name) -> const char * { return k.nodal[slot] ? nullptr : name; };
  auto need_f = [&](int slot, const char *name) -> const char * { return k.face[slot] ? nullptr : name; };
  const char *m = need_n(NALU_FN_COORDINATES, "coordinates");
  if (m) return m;
  const bool moves = (k.options & NALU_OPT_MESH_MOVES) != 0;
  switch (k.kind) {
  case NALU_FK_MOMENTUM_SYMMETRY:
    if ((m = need_n(NALU_FN_VELOCITY, "velocity")) || (m = need_n(NALU_FN_VISCOSITY, "viscosity")) ||
        (m = need_f(NALU_FF_EXPOSED_AREA_VEC, "exposed_area_vector")))
      return m;
    break;
  case NALU_FK_SCALAR_OPEN_ADV:
    if ((m = need_n(NALU_FN_SCALARQ, "scalarQ")) || (m = need_n(NALU_FN_BC_SCALARQ, "bcScalarQ")) || (m = need_n(NALU_FN_GJQ, "Gjq")) ||
        (m = need_n(NALU_FN_VISCOSITY, "diffFluxCoeff")) || (m = need_n(NALU_FN_DENSITY, "density")) ||
        (m = need_f(NALU_FF_OPEN_MASS_FLOW_RATE, "open_mass_flow_rate")))
      return m;
    if (!k.nodal[NALU_FN_VELOCITY_RTM] && !k.nodal[NALU_FN_VELOCITY]) return "velocity";
    break;
  case NALU_FK_CONTINUITY_OPEN:
    if ((m = need_n(NALU_FN_PRESSURE, "pressure")) || (m = need_n(NALU_FN_PRESSURE_BC, "pressure_bc")) ||
        (m = need_n(NALU_FN_DENSITY, "density")) || (m = need_n(NALU_FN_DPDX, "dpdx")) ||
        (m = need_f(NALU_FF_EXPOSED_AREA_VEC, "exposed_area_vector")) || (m = need_f(NALU_FF_DYNAMIC_PRESSURE, "dynamic_pressure")))
      return m;
    if (!k.nodal[NALU_FN_VELOCITY_RTM] && !k.nodal[NALU_FN_VELOCITY]) return "velocity";
    break;
  case NALU_FK_MOMENTUM_OPEN_ADV_DIFF:
    if ((m = need_n(NALU_FN_VELOCITY, "velocity")) || (m = need_n(NALU_FN_VISCOSITY, "viscosity")) ||
        (m = need_n(NALU_FN_DENSITY, "density")) || (m = need_n(NALU_FN_GJUI, "dudx")) ||
        (m = need_n(NALU_FN_VELOCITY_BC, "open_velocity_bc")) || (m = need_f(NALU_FF_EXPOSED_AREA_VEC, "exposed_area_vector")) ||
        (m = need_f(NALU_FF_OPEN_MASS_FLOW_RATE, "open_mass_flow_rate")))
      return m;
    if (moves && ((m = need_n(NALU_FN_VELOCITY_RTM, "velocity_rtm")) || (m = need_n(NALU_FN_MESH_VELOCITY, "mesh_velocity")))) return m;
    break;
  case NALU_FK_MOMENTUM_WALL_FUNCTION:
    if ((m = need_n(NALU_FN_VELOCITY, "velocity")) || (m = need_n(NALU_FN_WALL_VELOCITY_BC, "wall_velocity_bc")) ||
        (m = need_n(NALU_FN_DENSITY, "density")) || (m = need_n(NALU_FN_VISCOSITY, "viscosity")) ||
        (m = need_f(NALU_FF_EXPOSED_AREA_VEC, "exposed_area_vector")) ||
        (m = need_f(NALU_FF_WALL_FRICTION_VELOCITY_BIP, "wall_friction_velocity_bip")) ||
        (m = need_f(NALU_FF_WALL_NORMAL_DISTANCE_BIP, "wall_normal_distance_bip")))
      return m;
    break;
  default: return "kind";
  }
  return nullptr;
}

} // namespace face
} // namespace nalu_b200

using namespace nalu_b200;

extern "C" {

int nalu_b200_face_set_create(nalu_b200_face_set **out, const nalu_b200_graph *h, int64_t n_faces, const int32_t *face_elem,
                              const int32_t *face_ordinal)
{
  if (!out || !h || n_faces < 0 || (n_faces > 0 && (!face_elem || !face_ordinal))) return NALU_B200_ERR_ARG;
  const Graph &g = h->g;
  if (!g.finalized) return NALU_B200_ERR_STATE;
  if (g.npe != 8) {
    set_error("face_set_create: Quad4 faces of Hex8 elements only (the graph holds 27-node elements)");
    return NALU_B200_ERR_STATE;
  }
  // greedy colouring over the ROWS the faces add to: a face touches the rows of its four nodes (aliases of one row on
  // periodic meshes count as one), lowest colour not yet used at any of them
  std::vector<uint64_t> used((size_t)g.n_row_nodes(), 0);
  std::vector<int32_t> color((size_t)n_faces);
  int ncol = 0;
  for (int64_t f = 0; f < n_faces; ++f) {
    const int32_t e = face_elem[f], o = face_ordinal[f];
    if (e < 0 || e >= g.n_elems || o < 0 || o > 5) {
      set_error("face_set_create: face " + std::to_string(f) + " names element " + std::to_string(e) + ", ordinal " + std::to_string(o));
      return NALU_B200_ERR_ARG;
    }
    uint64_t m = 0;
    int rows[4];
    for (int i = 0; i < 4; ++i) {
      rows[i] = g.node_row[g.elem_nodes[(size_t)e * 8 + face::h_side_nodes[o * 4 + i]]];
      m |= used[rows[i]];
    }
    int c = 0;
    while (c < 64 && ((m >> c) & 1)) ++c;
    if (c == 64) {
      set_error("face_set_create: more than 64 face colours");
      return NALU_B200_ERR_ARG;
    }
    color[f] = c;
    for (int i = 0; i < 4; ++i) used[rows[i]] |= uint64_t(1) << c;
    ncol = std::max(ncol, c + 1);
  }
  auto *fs = new nalu_b200_face_set;
  fs->n_faces = n_faces;
  fs->n_colors = ncol;
  fs->color_ptr.assign(ncol + 1, 0);
  for (int64_t f = 0; f < n_faces; ++f) ++fs->color_ptr[color[f] + 1];
  for (int c = 0; c < ncol; ++c) fs->color_ptr[c + 1] += fs->color_ptr[c];
  std::vector<int32_t> se((size_t)n_faces), so((size_t)n_faces), sf((size_t)n_faces);
  std::vector<int64_t> pos(fs->color_ptr.begin(), fs->color_ptr.end() - (ncol ? 1 : 0));
  for (int64_t f = 0; f < n_faces; ++f) {
    const int64_t q = pos[color[f]]++;
    se[q] = face_elem[f];
    so[q] = face_ordinal[f];
    sf[q] = (int32_t)f;
  }
  if (n_faces > 0) {
    const size_t nb = sizeof(int32_t) * (size_t)n_faces;
    if (cudaMalloc(&fs->d_elem, nb) != cudaSuccess || cudaMalloc(&fs->d_ord, nb) != cudaSuccess ||
        cudaMalloc(&fs->d_face, nb) != cudaSuccess) {
      nalu_b200_face_set_destroy(fs);
      return cuda_fail(cudaGetLastError(), "face_set_create: cudaMalloc");
    }
    cudaMemcpy(fs->d_elem, se.data(), nb, cudaMemcpyHostToDevice);
    cudaMemcpy(fs->d_ord, so.data(), nb, cudaMemcpyHostToDevice);
    cudaMemcpy(fs->d_face, sf.data(), nb, cudaMemcpyHostToDevice);
  }
  *out = fs;
  return NALU_B200_OK;
}

void nalu_b200_face_set_destroy(nalu_b200_face_set *fs)
{
  if (!fs) return;
  cudaFree(fs->d_elem);
  cudaFree(fs->d_ord);
  cudaFree(fs->d_face);
  delete fs;
}

int nalu_b200_face_set_num_colors(const nalu_b200_face_set *fs, int32_t *n_colors)
{
  if (!fs || !n_colors) return NALU_B200_ERR_ARG;
  *n_colors = fs->n_colors;
  return NALU_B200_OK;
}

int nalu_b200_compute_exposed_area(const nalu_b200_graph *h, const nalu_b200_face_set *fs, const double *d_coords, double *d_area_vec,
                                   void *stream)
{
  if (!h || !fs || !h->g.d_elem_nodes) return NALU_B200_ERR_STATE;
  if (fs->n_faces == 0) return NALU_B200_OK;
  if (!d_coords || !d_area_vec) return NALU_B200_ERR_ARG;
  const long long nt = 4 * fs->n_faces;
  face::exposed_area_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, (cudaStream_t)stream>>>(fs->n_faces, fs->d_elem, fs->d_ord, fs->d_face,
                                                                                           h->g.d_elem_nodes, d_coords, d_area_vec);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

int nalu_b200_assemble_face_elem(const nalu_b200_graph *h, const nalu_b200_face_set *fs, const nalu_b200_face_kernel_desc *k, int32_t nk,
                                 double *d_values, double *d_rhs, void *stream)
{
  if (!h || !fs || !h->g.d_browptr || !h->g.d_elem_slot) return NALU_B200_ERR_STATE;
  if (!k || nk < 1 || nk > face::kMaxFaceKernels || !d_values || !d_rhs) return NALU_B200_ERR_ARG;
  const Graph &g = h->g;
  for (int i = 0; i < nk; ++i) {
    if (k[i].kind < 0 || k[i].kind >= NALU_FK_COUNT) {
      set_error("assemble_face_elem: unknown kernel kind");
      return NALU_B200_ERR_ARG;
    }
    const bool vec = k[i].kind == NALU_FK_MOMENTUM_SYMMETRY || k[i].kind == NALU_FK_MOMENTUM_OPEN_ADV_DIFF ||
                     k[i].kind == NALU_FK_MOMENTUM_WALL_FUNCTION;
    if ((vec && g.ndof != 3) || (!vec && g.ndof != 1)) {
      set_error("assemble_face_elem: kernel kind does not match the linear system's numDof");
      return NALU_B200_ERR_ARG;
    }
    if (const char *m = face::missing_face_field(k[i])) {
      set_error(std::string("assemble_face_elem: field missing for kernel ") + std::to_string(i) + ": " + m);
      return NALU_B200_ERR_ARG;
    }
  }
  face::FaceParams p = {};
  p.elem_nodes = g.d_elem_nodes;
  p.node_row = g.d_node_row;
  p.browptr = (const long long *)g.d_browptr;
  p.elem_slot = g.d_elem_slot;
  p.values = d_values;
  p.rhs = d_rhs;
  p.err = g.d_err;
  p.nk = nk;
  for (int i = 0; i < nk; ++i) p.k[i] = k[i];
  for (int c = 0; c < fs->n_colors; ++c) {
    const int64_t a = fs->color_ptr[c], b = fs->color_ptr[c + 1];
    if (b == a) continue;
    p.n = b - a;
    p.elem = fs->d_elem + a;
    p.ord = fs->d_ord + a;
    p.face = fs->d_face + a;
    const unsigned grid = (unsigned)((4 * p.n + 127) / 128);
    if (g.ndof == 3) face::assemble_face_elem_kernel<3><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    else face::assemble_face_elem_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    count_launch();
  }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

} // extern "C"
