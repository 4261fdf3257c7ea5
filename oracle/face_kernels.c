/*
 * face_kernels.c -- CPU restatement of the face/element boundary-condition path:
 *   AssembleFaceElemSolverAlgorithm::execute (src/AssembleFaceElemSolverAlgorithm.C:78-108,
 *   include/AssembleFaceElemSolverAlgorithm.h:45-133) with the Quad4 face / Hex8 element kernels
 *   MomentumSymmetryElemKernel, MomentumOpenAdvDiffElemKernel, ScalarOpenAdvElemKernel, ContinuityOpenElemKernel
 *   (src/kernel/ *.C), the face master element pieces they call (Quad3DSCS: src/master_element/Quad43DCVFEM.C,
 *   MasterElementWork.F quad3d_scs_det / quad3d_shape_fcn; HexSCS::face_grad_op, ipNodeMap, opposingNodes,
 *   side_node_ordinals: src/master_element/Hex8CVFEM.C) and ComputeGeometryBoundaryAlgorithm::execute
 *   (exposed_area_vector, src/ComputeGeometryBoundaryAlgorithm.C:45-126).
 *
 * TEST INFRASTRUCTURE ONLY (see nalu_oracle.h).  Scalar C in the reference's loop order.
 * PARITY-UNPINNED: the reference's own tests of these kernels (unit_tests/kernels/UnitTestFaceElemBasic.C) only
 * execute them, they hold no gold values; tests/test_oracle_golden.py checks this restatement against closed forms
 * (unit-cube area vectors, exactness of the face gradient operator on linear fields, flux balance of a uniform flow).
 */
#include "nalu_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- Hex8CVFEM.C:288-300 (oppNode_), :348-378 (intgExpFace_), :381-393 (ipNodeMap_), :414-422 (sideNodeOrdinals_) ---- */
static const int oppNode_[24] = {3, 2, 6, 7, 0, 3, 7, 4, 1, 0, 4, 5, 1, 5, 6, 2, 4, 7, 6, 5, 0, 1, 2, 3};
static const int faceIpNodeMap_[24] = {0, 1, 5, 4, 1, 2, 6, 5, 2, 3, 7, 6, 0, 4, 7, 3, 0, 3, 2, 1, 4, 5, 6, 7};
static const int sideNodeOrdinals_[24] = {0, 1, 5, 4, 1, 2, 6, 5, 2, 3, 7, 6, 0, 4, 7, 3, 0, 3, 2, 1, 4, 5, 6, 7};
static const double intgExpFace_[72] = {
    -0.25, -0.50, -0.25, 0.25,  -0.50, -0.25, 0.25,  -0.50, 0.25,  -0.25, -0.50, 0.25,  /* face 0 */
    0.50,  -0.25, -0.25, 0.50,  0.25,  -0.25, 0.50,  0.25,  0.25,  0.50,  -0.25, 0.25,  /* face 1 */
    0.25,  0.50,  -0.25, -0.25, 0.50,  -0.25, -0.25, 0.50,  0.25,  0.25,  0.50,  0.25,  /* face 2 */
    -0.50, -0.25, -0.25, -0.50, -0.25, 0.25,  -0.50, 0.25,  0.25,  -0.50, 0.25,  -0.25, /* face 3 */
    -0.25, -0.25, -0.50, -0.25, 0.25,  -0.50, 0.25,  0.25,  -0.50, 0.25,  -0.25, -0.50, /* face 4 */
    -0.25, -0.25, 0.50,  0.25,  -0.25, 0.50,  0.25,  0.25,  0.50,  -0.25, 0.25,  0.50}; /* face 5 */
/* Hex8CVFEM.C:424-441: the shifted face points are the face's nodes */
static const double nodeLocations[8][3] = {{-0.5, -0.5, -0.5}, {+0.5, -0.5, -0.5}, {+0.5, +0.5, -0.5}, {-0.5, +0.5, -0.5},
                                           {-0.5, -0.5, +0.5}, {+0.5, -0.5, +0.5}, {+0.5, +0.5, +0.5}, {-0.5, +0.5, +0.5}};

const int *orc_hex8_side_node_ordinals(int ordinal) { return &sideNodeOrdinals_[ordinal * 4]; }
const int *orc_hex8_face_ip_node_map(int ordinal) { return &faceIpNodeMap_[ordinal * 4]; }
int orc_hex8_opposing_node(int ordinal, int ip) { return oppNode_[ordinal * 4 + ip]; }

/* HexSCS::face_grad_op / shifted_face_grad_op (Hex8CVFEM.C:668-684,735-742): gradop [4][8][3] */
int orc_hex8_face_grad_op(int face_ordinal, int shifted, const double *coords, double *gradop)
{
  double exp_face[12], deriv[4 * 8 * 3];
  for (int k = 0; k < 4; ++k)
    for (int j = 0; j < 3; ++j)
      exp_face[3 * k + j] = shifted ? nodeLocations[sideNodeOrdinals_[4 * face_ordinal + k]][j] : intgExpFace_[12 * face_ordinal + 3 * k + j];
  orc_hex8_derivative(4, exp_face, deriv);
  return orc_generic_grad_op(4, deriv, coords, gradop, NULL);
}

/* quad3d_scs_det (MasterElementWork.F:1467-1586): area vectors of the 4 sub-faces of a Quad4 in 3-D; coords [4][3] */
void orc_quad3d_scs_areav(const double *cordel, double *area /*[4][3]*/)
{
  const double half = 0.5, one4th = 1.0 / 4.0;
  double p[4][3], e[4][3], c[3];
  for (int i = 0; i < 4; ++i)
    for (int idim = 0; idim < 3; ++idim) p[i][idim] = cordel[3 * i + idim];
  for (int idim = 0; idim < 3; ++idim) {
    e[0][idim] = (p[0][idim] + p[1][idim]) * half;
    e[1][idim] = (p[1][idim] + p[2][idim]) * half;
    e[2][idim] = (p[2][idim] + p[3][idim]) * half;
    e[3][idim] = (p[3][idim] + p[0][idim]) * half;
    c[idim] = (p[0][idim] + p[1][idim] + p[2][idim] + p[3][idim]) * one4th;
  }
  /* sub-face k: d13 = c - p(k), d24 = e(k-1) - e(k) */
  static const int prev[4] = {3, 0, 1, 2};
  for (int k = 0; k < 4; ++k) {
    const double dx13 = c[0] - p[k][0], dx24 = e[prev[k]][0] - e[k][0];
    const double dy13 = c[1] - p[k][1], dy24 = e[prev[k]][1] - e[k][1];
    const double dz13 = c[2] - p[k][2], dz24 = e[prev[k]][2] - e[k][2];
    area[3 * k + 0] = half * (dz24 * dy13 - dz13 * dy24);
    area[3 * k + 1] = half * (dx24 * dz13 - dx13 * dz24);
    area[3 * k + 2] = half * (dy24 * dx13 - dy13 * dx24);
  }
}

/* Quad3DSCS::shape_fcn / shifted_shape_fcn (Quad43DCVFEM.C:31-43,92-108; quad3d_shape_fcn MasterElementWork.F:3651-3695) */
void orc_quad3d_shape_fcn(int shifted, double shp[4][4])
{
  static const double intgLoc_[8] = {-0.25, -0.25, 0.25, -0.25, 0.25, 0.25, -0.25, 0.25};
  static const double intgLocShift_[8] = {-0.50, -0.50, 0.50, -0.50, 0.50, 0.50, -0.50, 0.50};
  const double *par = shifted ? intgLocShift_ : intgLoc_;
  const double one4th = 1.0 / 4.0, half = 1.0 / 2.0;
  for (int j = 0; j < 4; ++j) {
    const double s1 = par[2 * j], s2 = par[2 * j + 1];
    shp[j][0] = one4th + half * (-s1 - s2) + s1 * s2;
    shp[j][1] = one4th + half * (s1 - s2) - s1 * s2;
    shp[j][2] = one4th + half * (s1 + s2) + s1 * s2;
    shp[j][3] = one4th + half * (-s1 + s2) - s1 * s2;
  }
}

/* ComputeGeometryBoundaryAlgorithm::execute (src/ComputeGeometryBoundaryAlgorithm.C:62-123): exposed_area_vector of every
 * face = Quad3DSCS::determinant on the face's nodes (the element's side_node_ordinals of the face ordinal) */
void orc_compute_exposed_area(int64_t nFaces, const int32_t *faceElem, const int32_t *faceOrdinal, const int32_t *elemNodes,
                              const double *coords, double *areaVec /*[nFaces][4][3]*/)
{
  for (int64_t f = 0; f < nFaces; ++f) {
    double ws_coordinates[12];
    const int *ord = orc_hex8_side_node_ordinals(faceOrdinal[f]);
    for (int ni = 0; ni < 4; ++ni) {
      const int32_t node = elemNodes[8 * (size_t)faceElem[f] + ord[ni]];
      for (int j = 0; j < 3; ++j) ws_coordinates[3 * ni + j] = coords[3 * (size_t)node + j];
    }
    orc_quad3d_scs_areav(ws_coordinates, areaVec + 12 * f);
  }
}

/* ---- gathers: element nodes, face nodes (= element nodes in side_node_ordinals order; the reference's own test
 *      unit_tests/kernels/UnitTestFaceElemBasic.C:48-66 asserts that identity) ---- */
static void gather_elem(const double *field, int d, const int32_t *nodes, double *out)
{
  for (int n = 0; n < 8; ++n)
    for (int j = 0; j < d; ++j) out[n * d + j] = field[(size_t)nodes[n] * d + j];
}
static void gather_face(const double *field, int d, const int32_t *nodes, const int *ord, double *out)
{
  for (int n = 0; n < 4; ++n)
    for (int j = 0; j < d; ++j) out[n * d + j] = field[(size_t)nodes[ord[n]] * d + j];
}

/* src/PecletFunction.C:67-71,98-101 */
static double peclet_function(const orc_face_kernel_desc *k, double x)
{
  if (k->pecletTanh) return 0.50 * (1.0 + tanh((x - k->tanhTrans) / k->tanhWidth));
  const double modPeclet = k->hybridFactor * x;
  return modPeclet * modPeclet / (5.0 + modPeclet * modPeclet);
}

/* ---- src/kernel/MomentumSymmetryElemKernel.C:68-163 ---- */
static void momentum_symmetry(const orc_face_kernel_desc *k, const int32_t *nodes, int elemFaceOrdinal, int64_t face,
                              double *lhs, double *rhs)
{
  const int n = 24;
  const int *ord = orc_hex8_side_node_ordinals(elemFaceOrdinal);
  double vf_viscosity[4], v_uNp1[8][3], coords[8][3], v_dndx_fc_elem[4][8][3], vf_shape_function_[4][4], w_nx[3];
  const double *vf_exposedAreaVec = k->exposedAreaVec + 12 * face;
  gather_face(k->viscosity, 1, nodes, ord, vf_viscosity);
  gather_elem(k->velocity, 3, nodes, &v_uNp1[0][0]);
  gather_elem(k->coordinates, 3, nodes, &coords[0][0]);
  orc_hex8_face_grad_op(elemFaceOrdinal, k->shiftedGradOp, &coords[0][0], &v_dndx_fc_elem[0][0][0]);
  orc_quad3d_shape_fcn(0, vf_shape_function_);
  const double includeDivU_ = k->includeDivU;

  for (int ip = 0; ip < 4; ++ip) {
    const int nearestNode = orc_hex8_face_ip_node_map(elemFaceOrdinal)[ip];
    double asq = 0.0;
    for (int j = 0; j < 3; ++j) {
      const double axj = vf_exposedAreaVec[3 * ip + j];
      asq += axj * axj;
    }
    const double amag = sqrt(asq);
    for (int i = 0; i < 3; ++i) w_nx[i] = vf_exposedAreaVec[3 * ip + i] / amag;
    double viscBip = 0.0;
    for (int ic = 0; ic < 4; ++ic) viscBip += vf_shape_function_[ip][ic] * vf_viscosity[ic];

    for (int ic = 0; ic < 8; ++ic) {
      const int icNdim = ic * 3;
      for (int j = 0; j < 3; ++j) {
        const double axj = vf_exposedAreaVec[3 * ip + j];
        const double dndxj = v_dndx_fc_elem[ip][ic][j];
        const double uxj = v_uNp1[ic][j];
        const double divUstress = 2.0 / 3.0 * viscBip * dndxj * uxj * axj * includeDivU_;
        for (int i = 0; i < 3; ++i) {
          const int indexR = nearestNode * 3 + i;
          const double dndxi = v_dndx_fc_elem[ip][ic][i];
          const double uxi = v_uNp1[ic][i];
          const double nxi = w_nx[i];
          const double nxinxi = nxi * nxi;
          double lhsfac = -viscBip * dndxj * axj * nxinxi;
          lhs[indexR * n + icNdim + i] += lhsfac;
          rhs[indexR] -= lhsfac * uxi + divUstress * nxinxi;
          lhsfac = -viscBip * dndxi * axj * nxinxi;
          lhs[indexR * n + icNdim + j] += lhsfac;
          rhs[indexR] -= lhsfac * uxj;
          for (int l = 0; l < 3; ++l) {
            if (i != l) {
              const double nxinxl = nxi * w_nx[l];
              const double uxl = v_uNp1[ic][l];
              const double dndxl = v_dndx_fc_elem[ip][ic][l];
              lhsfac = -viscBip * dndxj * axj * nxinxl;
              lhs[indexR * n + icNdim + l] += lhsfac;
              rhs[indexR] -= lhsfac * uxl + divUstress * nxinxl;
              lhsfac = -viscBip * dndxl * axj * nxinxl;
              lhs[indexR * n + icNdim + j] += lhsfac;
              rhs[indexR] -= lhsfac * uxj;
            }
          }
        }
      }
    }
  }
}

/* ---- src/kernel/ScalarOpenAdvElemKernel.C:90-195 ---- */
static void scalar_open_adv(const orc_face_kernel_desc *k, const int32_t *nodes, int elemFaceOrdinal, int64_t face, double *lhs,
                            double *rhs)
{
  const int n = 8;
  const double small_ = 1.0e-16;
  const double alphaUpw_ = k->alphaUpw, om_alphaUpw_ = 1.0 - k->alphaUpw, hoUpwind_ = k->hoUpwind;
  const int *face_node_ordinals = orc_hex8_side_node_ordinals(elemFaceOrdinal);
  double vf_scalarQ[4], vf_bcScalarQ[4], vf_Gjq[4][3], vf_coordinates[4][3];
  double v_coordinates[8][3], v_vrtm[8][3], v_diffFluxCoeff[8], v_density[8], vf_adv_shape_function_[4][4], w_coordBip[3];
  const double *vf_openMassFlowRate = k->openMassFlowRate + 4 * face;
  gather_face(k->scalarQ, 1, nodes, face_node_ordinals, vf_scalarQ);
  gather_face(k->bcScalarQ, 1, nodes, face_node_ordinals, vf_bcScalarQ);
  gather_face(k->Gjq, 3, nodes, face_node_ordinals, &vf_Gjq[0][0]);
  gather_face(k->coordinates, 3, nodes, face_node_ordinals, &vf_coordinates[0][0]);
  gather_elem(k->coordinates, 3, nodes, &v_coordinates[0][0]);
  gather_elem(k->velocityRTM ? k->velocityRTM : k->velocity, 3, nodes, &v_vrtm[0][0]);
  gather_elem(k->viscosity, 1, nodes, v_diffFluxCoeff);
  gather_elem(k->density, 1, nodes, v_density);
  orc_quad3d_shape_fcn(k->skewSymmetric, vf_adv_shape_function_);
  static const int faceIpNodeMapQuad_[4] = {0, 1, 2, 3}; /* Quad43DCVFEM.C:29-34 */

  for (int ip = 0; ip < 4; ++ip) {
    const int opposingNode = orc_hex8_opposing_node(elemFaceOrdinal, ip);
    const int nearestNode = orc_hex8_face_ip_node_map(elemFaceOrdinal)[ip];
    const int localFaceNode = faceIpNodeMapQuad_[ip];
    for (int j = 0; j < 3; ++j) w_coordBip[j] = 0.0;
    double qIp = 0.0, qIpEntrain = 0.0;
    for (int ic = 0; ic < 4; ++ic) {
      const double rAdv = vf_adv_shape_function_[ip][ic];
      qIp += rAdv * vf_scalarQ[ic];
      qIpEntrain += rAdv * vf_bcScalarQ[ic];
      for (int j = 0; j < 3; ++j) w_coordBip[j] += rAdv * vf_coordinates[ic][j];
    }
    double udotx = 0.0, dqR = 0.0;
    for (int i = 0; i < 3; ++i) {
      const double dxi = v_coordinates[nearestNode][i] - v_coordinates[opposingNode][i];
      udotx += 0.5 * dxi * (v_vrtm[opposingNode][i] + v_vrtm[nearestNode][i]);
      const double dxBip = w_coordBip[i] - v_coordinates[nearestNode][i];
      dqR += dxBip * vf_Gjq[localFaceNode][i] * hoUpwind_;
    }
    const double qIpUpw = vf_scalarQ[localFaceNode] + dqR;
    const double diffIp = 0.5 * (v_diffFluxCoeff[opposingNode] / v_density[opposingNode] + v_diffFluxCoeff[nearestNode] / v_density[nearestNode]);
    const double pecFuncArg = fabs(udotx) / (diffIp + small_);
    const double pecfac = peclet_function(k, pecFuncArg);
    const double om_pecfac = 1.0 - pecfac;
    const double tmdot = vf_openMassFlowRate[ip];
    const double mdotLeaving = tmdot > 0 ? 1.0 : 0.0;
    const double om_mdotLeaving = 1.0 - mdotLeaving;
    const double qUpwind = alphaUpw_ * qIpUpw + om_alphaUpw_ * qIp;
    const double afluxLeaving = tmdot * (pecfac * qUpwind + om_pecfac * qIp);
    rhs[nearestNode] -= afluxLeaving * mdotLeaving;
    lhs[nearestNode * n + nearestNode] += tmdot * pecfac * alphaUpw_ * mdotLeaving;
    const double fac = tmdot * (pecfac * om_alphaUpw_ + om_pecfac) * mdotLeaving;
    for (int ic = 0; ic < 4; ++ic) {
      const int nn = face_node_ordinals[ic];
      lhs[nearestNode * n + nn] += vf_adv_shape_function_[ip][ic] * fac;
    }
    const double afluxEntraining = tmdot * qIpEntrain;
    rhs[nearestNode] -= afluxEntraining * om_mdotLeaving;
  }
}

/* ---- src/kernel/ContinuityOpenElemKernel.C:80-199 (setup :89-97: projTimeScale = dt / gamma1) ---- */
static void continuity_open(const orc_face_kernel_desc *k, const int32_t *nodes, int elemFaceOrdinal, int64_t face, double *lhs,
                            double *rhs)
{
  const int n = 8;
  const double penaltyFac_ = 2.0; /* include/kernel/ContinuityOpenElemKernel.h:67 */
  const double projTimeScale_ = k->dt / k->gamma1;
  const int *face_node_ordinals = orc_hex8_side_node_ordinals(elemFaceOrdinal);
  double vf_pressure[4], vf_pressureBc[4], vf_Gpdx[4][3], vf_density[4], vf_vrtm[4][3], v_pressure[8], coords[8][3];
  double dndx[4][8][3], dndx_shifted[4][8][3], vf_shape_function_[4][4];
  double w_rho_uBip[3], w_GpdxBip[3], w_dpdxBip[3];
  const double *vf_exposedAreaVec = k->exposedAreaVec + 12 * face;
  const double *vf_dynamicP = k->dynamicPressure + 4 * face;
  gather_face(k->pressure, 1, nodes, face_node_ordinals, vf_pressure);
  gather_face(k->pressureBc, 1, nodes, face_node_ordinals, vf_pressureBc);
  gather_face(k->density, 1, nodes, face_node_ordinals, vf_density);
  gather_face(k->velocityRTM ? k->velocityRTM : k->velocity, 3, nodes, face_node_ordinals, &vf_vrtm[0][0]);
  gather_face(k->Gpdx, 3, nodes, face_node_ordinals, &vf_Gpdx[0][0]);
  gather_elem(k->pressure, 1, nodes, v_pressure);
  gather_elem(k->coordinates, 3, nodes, &coords[0][0]);
  orc_hex8_face_grad_op(elemFaceOrdinal, 0, &coords[0][0], &dndx[0][0][0]);
  orc_hex8_face_grad_op(elemFaceOrdinal, 1, &coords[0][0], &dndx_shifted[0][0][0]);
  const double(*v_dndx_fc_elem)[8][3] = k->shiftedGradOp ? dndx_shifted : dndx;
  const double(*v_dndx_fc_elem_lhs)[8][3] = (k->shiftedGradOp || k->reducedSensitivities) ? dndx_shifted : dndx;
  orc_quad3d_shape_fcn(k->shiftMdot, vf_shape_function_);

  for (int ip = 0; ip < 4; ++ip) {
    const int nearestNode = orc_hex8_face_ip_node_map(elemFaceOrdinal)[ip];
    double aMag = 0.0;
    for (int j = 0; j < 3; ++j) {
      w_rho_uBip[j] = 0.0;
      w_GpdxBip[j] = 0.0;
      w_dpdxBip[j] = 0.0;
      const double axj = vf_exposedAreaVec[3 * ip + j];
      aMag += axj * axj;
    }
    aMag = sqrt(aMag);
    double inverseLengthScale = 0.0;
    for (int ic = 0; ic < 4; ++ic) {
      const int faceNodeNumber = face_node_ordinals[ic];
      for (int j = 0; j < 3; ++j) inverseLengthScale += v_dndx_fc_elem[ip][faceNodeNumber][j] * vf_exposedAreaVec[3 * ip + j];
    }
    inverseLengthScale /= aMag;
    double pBip = 0.0;
    double pbcBip = -vf_dynamicP[ip];
    for (int ic = 0; ic < 4; ++ic) {
      const double r = vf_shape_function_[ip][ic];
      const double rhoIC = vf_density[ic];
      pBip += r * vf_pressure[ic];
      pbcBip += r * vf_pressureBc[ic];
      for (int j = 0; j < 3; ++j) {
        w_rho_uBip[j] += r * rhoIC * vf_vrtm[ic][j];
        w_GpdxBip[j] += r * vf_Gpdx[ic][j];
      }
    }
    for (int ic = 0; ic < 8; ++ic) {
      const double pIc = v_pressure[ic];
      for (int j = 0; j < 3; ++j) w_dpdxBip[j] += v_dndx_fc_elem[ip][ic][j] * pIc;
    }
    double mdot = penaltyFac_ * projTimeScale_ * inverseLengthScale * (pBip - pbcBip) * aMag;
    for (int j = 0; j < 3; ++j) {
      const double axj = vf_exposedAreaVec[3 * ip + j];
      mdot += (w_rho_uBip[j] - projTimeScale_ * (w_dpdxBip[j] - w_GpdxBip[j])) * axj;
    }
    for (int ic = 0; ic < 4; ++ic) {
      const int faceNodeNumber = face_node_ordinals[ic];
      const double r = vf_shape_function_[ip][ic];
      lhs[nearestNode * n + faceNodeNumber] += r * penaltyFac_ * inverseLengthScale * aMag;
    }
    for (int ic = 0; ic < 8; ++ic) {
      double lhsFac = 0.0;
      for (int j = 0; j < 3; ++j) lhsFac += -v_dndx_fc_elem_lhs[ip][ic][j] * vf_exposedAreaVec[3 * ip + j];
      lhs[nearestNode * n + ic] += lhsFac;
    }
    rhs[nearestNode] -= mdot / projTimeScale_;
  }
}

/* ---- src/kernel/MomentumOpenAdvDiffElemKernel.C:110-325 ---- */
static void momentum_open_adv_diff(const orc_face_kernel_desc *k, const int32_t *nodes, int elemFaceOrdinal, int64_t face,
                                   double *lhs, double *rhs)
{
  const int n = 24;
  const double small_ = 1.0e-16;
  const double alphaUpw_ = k->alphaUpw, om_alphaUpw_ = 1.0 - k->alphaUpw, hoUpwind_ = k->hoUpwind;
  const double includeDivU_ = k->includeDivU;
  const double meshVelocityCorrection_ = k->meshMoves ? 1.0 : 0.0;
  const int *face_node_ordinals = orc_hex8_side_node_ordinals(elemFaceOrdinal);
  double vf_density[4], vf_viscosity[4], vf_velocityNp1[4][3], vf_meshVelocity[4][3], vf_bcVelocity[4][3], vf_Gjui[4][3][3],
      vf_coordinates[4][3];
  double v_coordinates[8][3], v_velocityNp1[8][3], v_vrtm[8][3], v_viscosity[8], v_density[8], v_dndx_fc_elem[4][8][3];
  double vf_shape_function_[4][4], vf_adv_shape_function_[4][4];
  double w_uBip[3], w_rho_vBip[3], w_uBipExtrap[3], w_uspecBip[3], w_coordBip[3], w_nx[3];
  const double *vf_exposedAreaVec = k->exposedAreaVec + 12 * face;
  const double *vf_openMassFlowRate = k->openMassFlowRate + 4 * face;
  /* without mesh motion velocityRTM_ = meshVelocity_ = velocityNp1_ (MomentumOpenAdvDiffElemKernel.C:53-60) */
  const double *rtm = k->meshMoves ? k->velocityRTM : k->velocity;
  const double *vmesh = k->meshMoves ? k->meshVelocity : k->velocity;
  gather_face(k->density, 1, nodes, face_node_ordinals, vf_density);
  gather_face(k->viscosity, 1, nodes, face_node_ordinals, vf_viscosity);
  gather_face(k->velocity, 3, nodes, face_node_ordinals, &vf_velocityNp1[0][0]);
  gather_face(vmesh, 3, nodes, face_node_ordinals, &vf_meshVelocity[0][0]);
  gather_face(k->velocityBc, 3, nodes, face_node_ordinals, &vf_bcVelocity[0][0]);
  gather_face(k->Gjui, 9, nodes, face_node_ordinals, &vf_Gjui[0][0][0]);
  gather_face(k->coordinates, 3, nodes, face_node_ordinals, &vf_coordinates[0][0]);
  gather_elem(k->coordinates, 3, nodes, &v_coordinates[0][0]);
  gather_elem(k->velocity, 3, nodes, &v_velocityNp1[0][0]);
  gather_elem(rtm, 3, nodes, &v_vrtm[0][0]);
  gather_elem(k->viscosity, 1, nodes, v_viscosity);
  gather_elem(k->density, 1, nodes, v_density);
  orc_hex8_face_grad_op(elemFaceOrdinal, k->shiftedGradOp, &v_coordinates[0][0], &v_dndx_fc_elem[0][0][0]);
  orc_quad3d_shape_fcn(0, vf_shape_function_);
  orc_quad3d_shape_fcn(k->skewSymmetric, vf_adv_shape_function_);
  static const int faceIpNodeMapQuad_[4] = {0, 1, 2, 3};

  for (int ip = 0; ip < 4; ++ip) {
    const int opposingNode = orc_hex8_opposing_node(elemFaceOrdinal, ip);
    const int nearestNode = orc_hex8_face_ip_node_map(elemFaceOrdinal)[ip];
    const int localFaceNode = faceIpNodeMapQuad_[ip];
    double asq = 0.0;
    for (int j = 0; j < 3; ++j) {
      w_uBip[j] = 0.0;
      w_rho_vBip[j] = 0.0;
      w_uspecBip[j] = 0.0;
      w_coordBip[j] = 0.0;
      const double axj = vf_exposedAreaVec[3 * ip + j];
      asq += axj * axj;
    }
    const double amag = sqrt(asq);
    double rhoBip = 0.0, viscBip = 0.0;
    for (int ic = 0; ic < 4; ++ic) {
      const double r = vf_shape_function_[ip][ic];
      const double rAdv = vf_adv_shape_function_[ip][ic];
      rhoBip += r * vf_density[ic];
      viscBip += r * vf_viscosity[ic];
      const double rhoIc = vf_density[ic];
      for (int j = 0; j < 3; ++j) {
        w_uspecBip[j] += rAdv * vf_bcVelocity[ic][j];
        w_uBip[j] += rAdv * vf_velocityNp1[ic][j];
        w_coordBip[j] += rAdv * vf_coordinates[ic][j];
        w_rho_vBip[j] += r * rhoIc * vf_meshVelocity[ic][j];
      }
    }
    double udotx = 0.0;
    for (int i = 0; i < 3; ++i) {
      const double dxi = v_coordinates[nearestNode][i] - v_coordinates[opposingNode][i];
      udotx += 0.5 * dxi * (v_vrtm[opposingNode][i] + v_vrtm[nearestNode][i]);
      w_nx[i] = vf_exposedAreaVec[3 * ip + i] / amag;
      double duR = 0.0;
      for (int j = 0; j < 3; ++j) {
        const double dxj = w_coordBip[j] - v_coordinates[nearestNode][j];
        duR += dxj * vf_Gjui[localFaceNode][i][j] * hoUpwind_;
      }
      w_uBipExtrap[i] = v_velocityNp1[nearestNode][i] + duR;
    }
    const double diffIp = 0.5 * (v_viscosity[opposingNode] / v_density[opposingNode] + v_viscosity[nearestNode] / v_density[nearestNode]);
    const double pecFuncArg = fabs(udotx) / (diffIp + small_);
    const double pecfac = peclet_function(k, pecFuncArg);
    const double om_pecfac = 1.0 - pecfac;
    const double tmdot = vf_openMassFlowRate[ip];
    const double mdotLeaving = tmdot > 0 ? 1.0 : 0.0;
    const double om_mdotLeaving = 1.0 - mdotLeaving;
    double mvc = 0.0;
    for (int j = 0; j < 3; ++j) mvc += w_rho_vBip[j] * vf_exposedAreaVec[3 * ip + j];
    const double uEntrain = tmdot / (rhoBip * amag) + mvc / (rhoBip * amag) * meshVelocityCorrection_;
    double uspecbipnx = 0.0;
    for (int j = 0; j < 3; ++j) uspecbipnx += w_uspecBip[j] * w_nx[j];

    for (int i = 0; i < 3; ++i) {
      const int indexR = nearestNode * 3 + i;
      const double uiIp = w_uBip[i];
      const double uiUpwind = alphaUpw_ * w_uBipExtrap[i] + om_alphaUpw_ * uiIp;
      const double afluxLeaving = tmdot * (pecfac * uiUpwind + om_pecfac * uiIp);
      rhs[indexR] -= afluxLeaving * mdotLeaving;
      lhs[indexR * n + nearestNode * 3 + i] += tmdot * pecfac * alphaUpw_ * mdotLeaving;
      const double fac = tmdot * (pecfac * om_alphaUpw_ + om_pecfac) * mdotLeaving;
      for (int ic = 0; ic < 4; ++ic) {
        const int nn = face_node_ordinals[ic];
        lhs[indexR * n + nn * 3 + i] += vf_adv_shape_function_[ip][ic] * fac;
      }
      const double nxi = w_nx[i];
      const double afluxEntraining = tmdot * uEntrain * nxi + tmdot * (w_uspecBip[i] - uspecbipnx * nxi);
      rhs[indexR] -= afluxEntraining * om_mdotLeaving;
    }

    for (int ic = 0; ic < 8; ++ic) {
      for (int j = 0; j < 3; ++j) {
        const double axj = vf_exposedAreaVec[3 * ip + j];
        const double dndxj = v_dndx_fc_elem[ip][ic][j];
        const double uxj = v_velocityNp1[ic][j];
        const double divUstress = 2.0 / 3.0 * viscBip * dndxj * uxj * axj * includeDivU_;
        for (int i = 0; i < 3; ++i) {
          const int indexR = nearestNode * 3 + i;
          const double dndxi = v_dndx_fc_elem[ip][ic][i];
          const double uxi = v_velocityNp1[ic][i];
          const double nxi = w_nx[i];
          const double om_nxinxi = 1.0 - nxi * nxi;
          double lhsfac = -viscBip * dndxj * axj * om_nxinxi;
          lhs[indexR * n + ic * 3 + i] += lhsfac;
          rhs[indexR] -= lhsfac * uxi + divUstress * om_nxinxi;
          lhsfac = -viscBip * dndxi * axj * om_nxinxi;
          lhs[indexR * n + ic * 3 + j] += lhsfac;
          rhs[indexR] -= lhsfac * uxj;
          for (int l = 0; l < 3; ++l) {
            if (i != l) {
              const double nxinxl = nxi * w_nx[l];
              const double uxl = v_velocityNp1[ic][l];
              const double dndxl = v_dndx_fc_elem[ip][ic][l];
              lhsfac = viscBip * dndxj * axj * nxinxl;
              lhs[indexR * n + ic * 3 + l] += lhsfac;
              rhs[indexR] -= lhsfac * uxl + divUstress * nxinxl;
              lhsfac = viscBip * dndxl * axj * nxinxl;
              lhs[indexR * n + ic * 3 + j] += lhsfac;
              rhs[indexR] -= lhsfac * uxj;
            }
          }
        }
      }
    }
  }
}

/* one face through the registered kernels: zero, run every kernel in order (AssembleFaceElemSolverAlgorithm.C:90-97);
 * lhs is [rhsSize][rhsSize] with rhsSize = 8 * numDof (the ELEMENT's nodes, include/AssembleFaceElemSolverAlgorithm.h:67) */
/* ---- src/kernel/MomentumWallFunctionElemKernel.C:73-155 ----
 * The reference registers this kernel on an AssembleElemSolverAlgorithm over the FACES of the wall part (lhs is 12 x 12 on the
 * face's 4 nodes).  Here it is written into the (8 nodes x 3)^2 matrix of the element the face belongs to, at the rows / columns
 * of the face's nodes (side_node_ordinals): the same CSR entries receive the same values. */
static void momentum_wall_function(const orc_face_kernel_desc *k, const int32_t *nodes, int elemFaceOrdinal, int64_t face,
                                   double *lhs, double *rhs)
{
  const int n = 24;
  const int *ord = orc_hex8_side_node_ordinals(elemFaceOrdinal);
  double v_uNp1[4][3], v_bcVelocity[4][3], v_density[4], v_viscosity[4], vf_shape_function_[4][4];
  double w_uNp1Bip[3], w_uBcBip[3], w_unitNormal[3];
  const double *vf_exposedAreaVec = k->exposedAreaVec + 12 * face;
  const double *vf_utau = k->wallFrictionVelocityBip + 4 * face;
  const double *vf_yp = k->wallNormalDistanceBip + 4 * face;
  gather_face(k->velocity, 3, nodes, ord, &v_uNp1[0][0]);
  gather_face(k->wallVelocityBc, 3, nodes, ord, &v_bcVelocity[0][0]);
  gather_face(k->density, 1, nodes, ord, v_density);
  gather_face(k->viscosity, 1, nodes, ord, v_viscosity);
  orc_quad3d_shape_fcn(0, vf_shape_function_);
  const double elog_ = k->elog, kappa_ = k->kappa, yplusCrit_ = k->yplusCrit;

  for (int ip = 0; ip < 4; ++ip) {
    /* Quad3DSCS::ipNodeMap is the identity (face-local); the face's node ip is element node ord[ip] */
    const int nearestNode = ord[ip];
    double aMag = 0.0;
    for (int j = 0; j < 3; ++j) {
      w_uNp1Bip[j] = 0.0;
      w_uBcBip[j] = 0.0;
      const double axj = vf_exposedAreaVec[3 * ip + j];
      aMag += axj * axj;
    }
    aMag = sqrt(aMag);
    double rhoBip = 0.0, muBip = 0.0;
    for (int ic = 0; ic < 4; ++ic) {
      const double r = vf_shape_function_[ip][ic];
      rhoBip += r * v_density[ic];
      muBip += r * v_viscosity[ic];
      for (int j = 0; j < 3; ++j) {
        w_uNp1Bip[j] += r * v_uNp1[ic][j];
        w_uBcBip[j] += r * v_bcVelocity[ic][j];
      }
    }
    for (int j = 0; j < 3; ++j) w_unitNormal[j] = vf_exposedAreaVec[3 * ip + j] / aMag;
    const double yp = vf_yp[ip];
    const double utau = vf_utau[ip];
    const double yplus = rhoBip * yp * utau / muBip;
    const double lambda = yplus > yplusCrit_ ? rhoBip * kappa_ * utau / log(elog_ * yplus) * aMag : muBip / yp * aMag;

    for (int i = 0; i < 3; ++i) {
      const int indexR = nearestNode * 3 + i;
      double uiTan = 0.0, uiBcTan = 0.0;
      for (int j = 0; j < 3; ++j) {
        const double ninj = w_unitNormal[i] * w_unitNormal[j];
        if (i == j) {
          const double om_nini = 1.0 - ninj;
          uiTan += om_nini * w_uNp1Bip[j];
          uiBcTan += om_nini * w_uBcBip[j];
          for (int ic = 0; ic < 4; ++ic) lhs[indexR * n + ord[ic] * 3 + i] += lambda * om_nini * vf_shape_function_[ip][ic];
        } else {
          uiTan -= ninj * w_uNp1Bip[j];
          uiBcTan -= ninj * w_uBcBip[j];
          for (int ic = 0; ic < 4; ++ic) lhs[indexR * n + ord[ic] * 3 + j] -= lambda * ninj * vf_shape_function_[ip][ic];
        }
      }
      rhs[indexR] -= lambda * (uiTan - uiBcTan);
    }
  }
}

void orc_face_elem_lhs_rhs(const orc_face_kernel_desc *kernels, int nk, int numDof, const int32_t *elemNodes, int elemFaceOrdinal,
                           int64_t faceIndex, double *lhs, double *rhs)
{
  const int n = 8 * numDof;
  memset(lhs, 0, sizeof(double) * (size_t)n * n);
  memset(rhs, 0, sizeof(double) * (size_t)n);
  for (int i = 0; i < nk; ++i) {
    const orc_face_kernel_desc *k = &kernels[i];
    switch (k->kind) {
    case ORC_FK_MOMENTUM_SYMMETRY: momentum_symmetry(k, elemNodes, elemFaceOrdinal, faceIndex, lhs, rhs); break;
    case ORC_FK_SCALAR_OPEN_ADV: scalar_open_adv(k, elemNodes, elemFaceOrdinal, faceIndex, lhs, rhs); break;
    case ORC_FK_CONTINUITY_OPEN: continuity_open(k, elemNodes, elemFaceOrdinal, faceIndex, lhs, rhs); break;
    case ORC_FK_MOMENTUM_OPEN_ADV_DIFF: momentum_open_adv_diff(k, elemNodes, elemFaceOrdinal, faceIndex, lhs, rhs); break;
    case ORC_FK_MOMENTUM_WALL_FUNCTION: momentum_wall_function(k, elemNodes, elemFaceOrdinal, faceIndex, lhs, rhs); break;
    default: abort();
    }
  }
}

/* AssembleFaceElemSolverAlgorithm::execute over the locally owned faces of `rank`: apply_coeff with the element's
 * connected nodes (AssembleFaceElemSolverAlgorithm.C:99-104) */
void orc_assemble_face_elem(orc_linsys *ls, int rank, int numDof, const int32_t *elemNodesAll, const orc_face_kernel_desc *kernels,
                            int nk, int64_t nFaces, const int32_t *faceElem, const int32_t *faceOrdinal)
{
  const int n = 8 * numDof;
  double *lhs = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *rhs = (double *)malloc(sizeof(double) * (size_t)n);
  for (int64_t f = 0; f < nFaces; ++f) {
    const int32_t *nodes = elemNodesAll + 8 * (size_t)faceElem[f];
    orc_face_elem_lhs_rhs(kernels, nk, numDof, nodes, faceOrdinal[f], f, lhs, rhs);
    orc_linsys_sum_into(ls, rank, 8, nodes, rhs, lhs);
  }
  free(lhs);
  free(rhs);
}
