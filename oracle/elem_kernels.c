/*
 * elem_kernels.c -- oracle restatement of the element Kernel::execute bodies and of
 * the per-element part of AssembleElemSolverAlgorithm::execute.
 * TEST INFRASTRUCTURE ONLY (see nalu_oracle.h).  Scalar, reference loop order.
 */
#include "nalu_oracle.h"
#include <math.h>
#include <string.h>

/* what ScratchViews + MasterElementViews hold for one element
 * (include/ScratchViews.h:85-108,630-683) -- only the views the path uses */
typedef struct {
  double coords[8][3];
  double scs_areav[12][3];
  double dndx[12][8][3];
  double dndx_shifted[12][8][3];
  double deriv[12][8][3];
  double scv_volume[8];
  double dndx_scv[8][8][3]; /* dndx_scv or dndx_scv_shifted, whichever the SST kernels requested */
  double gijUpper[12][3][3], gijLower[12][3][3];
  int have_areav, have_dndx, have_dndx_shifted, have_scv_volume;
} me_views;

/* gather_elem_node_field (src/ScratchViews.C:16-88): 8 nodes x d contiguous doubles */
static void gather(const double *field, int d, const int32_t *nodes, double *out)
{
  for (int n = 0; n < 8; ++n)
    for (int j = 0; j < d; ++j) out[n * d + j] = field[(size_t)nodes[n] * d + j];
}

/* Kernel ctor: get_scs_shape_fn_data / get_scv_shape_fn_data (include/kernel/Kernel.h:26-56).
 * The reference fills these tables once per kernel object; cache them the same way. */
static double g_scs_shape[2][12][8];
static double g_scv_shape[2][8][8];
static volatile int g_tables_ready = 0;
static void init_tables(void)
{
  if (g_tables_ready) return;
  for (int s = 0; s < 2; ++s) {
    double loc[36], locv[24];
    orc_hex8_scs_intg_loc(s, loc);
    orc_hex8_shape_fcn(12, loc, &g_scs_shape[s][0][0]);
    orc_hex8_scv_intg_loc(s, locv);
    orc_hex8_shape_fcn(8, locv, &g_scv_shape[s][0][0]);
  }
  g_tables_ready = 1;
}
static void scs_shape(int shifted, double shp[12][8])
{
  memcpy(shp, g_scs_shape[shifted ? 1 : 0], sizeof(double) * 96);
}
static void scv_shape(int shifted, double shp[8][8])
{
  memcpy(shp, g_scv_shape[shifted ? 1 : 0], sizeof(double) * 64);
}

/* ---- src/kernel/MomentumAdvDiffElemKernel.C:70-186 ---- */
static void momentum_adv_diff(const orc_kernel_desc *k, const int32_t *nodes, int64_t e,
                              const me_views *me, double *lhs, double *rhs)
{
  const int n = 24;
  double v_uNp1[8][3], v_viscosity[8];
  const double *v_mdot = k->massFlowRate + 12 * e;
  gather(k->velocity, 3, nodes, &v_uNp1[0][0]);
  gather(k->viscosity, 1, nodes, v_viscosity);
  double v_shape_function_[12][8], v_adv_shape_function_[12][8];
  scs_shape(0, v_shape_function_);
  scs_shape(k->skewSymmetric, v_adv_shape_function_);
  const double(*v_dndx)[8][3] = k->shiftedGradOp ? me->dndx_shifted : me->dndx;
  const int *lrscv = orc_hex8_lrscv();
  double w_uIp[3];

  for (int ip = 0; ip < 12; ++ip) {
    const int il = lrscv[2 * ip];
    const int ir = lrscv[2 * ip + 1];
    const int ilNdim = il * 3;
    const int irNdim = ir * 3;
    const double tmdot = v_mdot[ip];

    double muIp = 0.0;
    double divU = 0.0;
    for (int i = 0; i < 3; ++i) w_uIp[i] = 0.0;

    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      const double rAdv = v_adv_shape_function_[ip][ic];
      muIp += r * v_viscosity[ic];
      for (int j = 0; j < 3; ++j) {
        const double uj = v_uNp1[ic][j];
        w_uIp[j] += rAdv * uj;
        divU += uj * v_dndx[ip][ic][j];
      }
    }

    for (int i = 0; i < 3; ++i) {
      const double uiIp = w_uIp[i];
      const double aflux = tmdot * uiIp;
      const double divUstress = 2.0 / 3.0 * muIp * divU * me->scs_areav[ip][i] * k->includeDivU;
      const int indexL = ilNdim + i;
      const int indexR = irNdim + i;
      rhs[indexL] -= aflux + divUstress;
      rhs[indexR] += aflux + divUstress;
    }

    for (int ic = 0; ic < 8; ++ic) {
      const int icNdim = ic * 3;
      const double lhsfacAdv = v_adv_shape_function_[ip][ic] * tmdot;

      for (int i = 0; i < 3; ++i) {
        const int indexL = ilNdim + i;
        const int indexR = irNdim + i;

        lhs[indexL * n + icNdim + i] += lhsfacAdv;
        lhs[indexR * n + icNdim + i] -= lhsfacAdv;

        double lhs_riC_i = 0.0;
        for (int j = 0; j < 3; ++j) {
          const double axj = me->scs_areav[ip][j];
          const double uj = v_uNp1[ic][j];

          const double lhsfacDiff_i = -muIp * v_dndx[ip][ic][j] * axj;
          lhs_riC_i += lhsfacDiff_i;

          const double lhsfacDiff_j = -muIp * v_dndx[ip][ic][i] * axj;
          lhs[indexL * n + icNdim + j] += lhsfacDiff_j;
          lhs[indexR * n + icNdim + j] -= lhsfacDiff_j;
          rhs[indexL] -= lhsfacDiff_j * uj;
          rhs[indexR] += lhsfacDiff_j * uj;
        }

        lhs[indexL * n + icNdim + i] += lhs_riC_i;
        lhs[indexR * n + icNdim + i] -= lhs_riC_i;
        const double ui = v_uNp1[ic][i];
        rhs[indexL] -= lhs_riC_i * ui;
        rhs[indexR] += lhs_riC_i * ui;
      }
    }
  }
}

/* ---- src/kernel/ContinuityAdvElemKernel.C:86-148 (+ setup :78-84) ---- */
static void continuity_adv(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me,
                           double *lhs, double *rhs)
{
  const int n = 8;
  const double projTimeScale_ = k->dt / k->gamma1;
  double v_densityNp1[8], v_pressure[8], v_velocity[8][3], v_Gpdx[8][3];
  gather(k->density, 1, nodes, v_densityNp1);
  gather(k->pressure, 1, nodes, v_pressure);
  gather(k->velocity, 3, nodes, &v_velocity[0][0]);
  gather(k->Gpdx, 3, nodes, &v_Gpdx[0][0]);
  double v_shape_function_[12][8];
  scs_shape(k->shiftMdot, v_shape_function_);
  const double(*v_dndx)[8][3] = k->shiftPoisson ? me->dndx_shifted : me->dndx;
  const double(*v_dndx_lhs)[8][3] =
      (k->shiftPoisson || k->reducedSensitivities) ? me->dndx_shifted : me->dndx;
  const int *lrscv = orc_hex8_lrscv();
  double w_rho_uIp[3], w_Gpdx_Ip[3], w_dpdxIp[3];

  for (int ip = 0; ip < 12; ++ip) {
    const int il = lrscv[2 * ip];
    const int ir = lrscv[2 * ip + 1];

    for (int j = 0; j < 3; ++j) {
      w_rho_uIp[j] = 0.0;
      w_Gpdx_Ip[j] = 0.0;
      w_dpdxIp[j] = 0.0;
    }

    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      const double nodalPressure = v_pressure[ic];
      const double nodalRho = v_densityNp1[ic];

      double lhsfac = 0.0;
      for (int j = 0; j < 3; ++j) {
        w_Gpdx_Ip[j] += r * v_Gpdx[ic][j];
        w_rho_uIp[j] += r * nodalRho * v_velocity[ic][j];
        w_dpdxIp[j] += v_dndx[ip][ic][j] * nodalPressure;
        lhsfac += -v_dndx_lhs[ip][ic][j] * me->scs_areav[ip][j];
      }

      lhs[il * n + ic] += lhsfac;
      lhs[ir * n + ic] -= lhsfac;
    }

    double mdot = 0.0;
    for (int j = 0; j < 3; ++j) {
      mdot += (w_rho_uIp[j] - projTimeScale_ * (w_dpdxIp[j] - w_Gpdx_Ip[j])) * me->scs_areav[ip][j];
    }

    rhs[il] -= mdot / projTimeScale_;
    rhs[ir] += mdot / projTimeScale_;
  }
}

/* ---- src/kernel/ScalarAdvDiffElemKernel.C:69-127 ---- */
static void scalar_adv_diff(const orc_kernel_desc *k, const int32_t *nodes, int64_t e,
                            const me_views *me, double *lhs, double *rhs)
{
  const int n = 8;
  double v_scalarQ[8], v_diffFluxCoeff[8];
  const double *v_mdot = k->massFlowRate + 12 * e;
  gather(k->scalarQ, 1, nodes, v_scalarQ);
  gather(k->viscosity, 1, nodes, v_diffFluxCoeff);
  double v_shape_function_[12][8], v_adv_shape_function_[12][8];
  scs_shape(0, v_shape_function_);
  scs_shape(k->skewSymmetric, v_adv_shape_function_);
  const double(*v_dndx)[8][3] = k->shiftedGradOp ? me->dndx_shifted : me->dndx;
  const int *lrscv = orc_hex8_lrscv();

  for (int ip = 0; ip < 12; ++ip) {
    const int il = lrscv[2 * ip];
    const int ir = lrscv[2 * ip + 1];
    const double tmdot = v_mdot[ip];

    double diffFluxCoeffIp = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      diffFluxCoeffIp += r * v_diffFluxCoeff[ic];
    }

    double qAdv = 0.0;
    double qDiff = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double lhsfacAdv = v_adv_shape_function_[ip][ic] * tmdot;
      qAdv += lhsfacAdv * v_scalarQ[ic];

      double lhsfacDiff = 0.0;
      for (int j = 0; j < 3; ++j) {
        lhsfacDiff += -diffFluxCoeffIp * v_dndx[ip][ic][j] * me->scs_areav[ip][j];
      }
      qDiff += lhsfacDiff * v_scalarQ[ic];

      lhs[il * n + ic] += lhsfacAdv + lhsfacDiff;
      lhs[ir * n + ic] -= lhsfacAdv + lhsfacDiff;
    }

    rhs[il] -= qAdv + qDiff;
    rhs[ir] += qAdv + qDiff;
  }
}

/* ---- src/kernel/ScalarDiffElemKernel.C:64-110 ---- */
static void scalar_diff(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me,
                        double *lhs, double *rhs)
{
  const int n = 8;
  double v_scalarQ[8], v_diffFluxCoeff[8];
  gather(k->scalarQ, 1, nodes, v_scalarQ);
  gather(k->viscosity, 1, nodes, v_diffFluxCoeff);
  double v_shape_function_[12][8];
  scs_shape(0, v_shape_function_);
  const double(*v_dndx)[8][3] = k->shiftedGradOp ? me->dndx_shifted : me->dndx;
  const int *lrscv = orc_hex8_lrscv();

  for (int ip = 0; ip < 12; ++ip) {
    const int il = lrscv[2 * ip];
    const int ir = lrscv[2 * ip + 1];

    double diffFluxCoeffIp = 0.0;
    for (int ic = 0; ic < 8; ++ic) diffFluxCoeffIp += v_shape_function_[ip][ic] * v_diffFluxCoeff[ic];

    double qDiff = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      double lhsfacDiff = 0.0;
      for (int j = 0; j < 3; ++j) lhsfacDiff += -diffFluxCoeffIp * v_dndx[ip][ic][j] * me->scs_areav[ip][j];
      qDiff += lhsfacDiff * v_scalarQ[ic];
      lhs[il * n + ic] += lhsfacDiff;
      lhs[ir * n + ic] -= lhsfacDiff;
    }
    rhs[il] -= qDiff;
    rhs[ir] += qDiff;
  }
}

/* ---- src/kernel/MomentumMassElemKernel.C:98-171 (+ setup :88-96) ---- */
static void momentum_mass(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me,
                          double *lhs, double *rhs)
{
  const int n = 24;
  double v_densityNm1[8], v_densityN[8], v_densityNp1[8];
  double v_velocityNm1[8][3], v_velocityN[8][3], v_velocityNp1[8][3], v_Gpdx[8][3];
  gather(k->densityNm1, 1, nodes, v_densityNm1);
  gather(k->densityN, 1, nodes, v_densityN);
  gather(k->density, 1, nodes, v_densityNp1);
  gather(k->velocityNm1, 3, nodes, &v_velocityNm1[0][0]);
  gather(k->velocityN, 3, nodes, &v_velocityN[0][0]);
  gather(k->velocity, 3, nodes, &v_velocityNp1[0][0]);
  gather(k->Gpdx, 3, nodes, &v_Gpdx[0][0]);
  double v_shape_function_[8][8];
  scv_shape(k->lumpedMass, v_shape_function_);
  const int *ipNodeMap_ = orc_hex8_scv_ip_node_map();
  const double gamma1_ = k->gamma1, gamma2_ = k->gamma2, gamma3_ = k->gamma3, dt_ = k->dt;
  double w_uNm1[3], w_uN[3], w_uNp1[3], w_Gjp[3];

  for (int ip = 0; ip < 8; ++ip) {
    const int nearestNode = ipNodeMap_[ip];

    double rhoNm1 = 0.0, rhoN = 0.0, rhoNp1 = 0.0;
    for (int j = 0; j < 3; j++) {
      w_uNm1[j] = 0.0;
      w_uN[j] = 0.0;
      w_uNp1[j] = 0.0;
      w_Gjp[j] = 0.0;
    }

    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      rhoNm1 += r * v_densityNm1[ic];
      rhoN += r * v_densityN[ic];
      rhoNp1 += r * v_densityNp1[ic];
      for (int j = 0; j < 3; j++) {
        w_uNm1[j] += r * v_velocityNm1[ic][j];
        w_uN[j] += r * v_velocityN[ic][j];
        w_uNp1[j] += r * v_velocityNp1[ic][j];
        w_Gjp[j] += r * v_Gpdx[ic][j];
      }
    }

    const double scV = me->scv_volume[ip];
    const int nnNdim = nearestNode * 3;
    const double densFac = rhoNp1 * k->densFac + k->om_densFac;
    for (int j = 0; j < 3; ++j) {
      rhs[nnNdim + j] += -(gamma1_ * rhoNp1 * w_uNp1[j] + gamma2_ * rhoN * w_uN[j] +
                           gamma3_ * rhoNm1 * w_uNm1[j]) *
                             scV / dt_ -
                         w_Gjp[j] * scV * densFac;
    }

    for (int ic = 0; ic < 8; ++ic) {
      const int icNdim = ic * 3;
      const double r = v_shape_function_[ip][ic];
      const double lhsfac = r * gamma1_ * rhoNp1 * scV / dt_;
      for (int j = 0; j < 3; ++j) {
        const int indexNN = nnNdim + j;
        lhs[indexNN * n + icNdim + j] += lhsfac;
      }
    }
  }
}

/* ---- src/kernel/ContinuityMassElemKernel.C:82-119 ---- */
static void continuity_mass(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me,
                            double *lhs, double *rhs)
{
  (void)lhs;
  const double gamma1_ = k->gamma1, gamma2_ = k->gamma2, gamma3_ = k->gamma3, dt_ = k->dt;
  const double projTimeScale = dt_ / gamma1_;
  double v_densityNm1[8], v_densityN[8], v_densityNp1[8];
  gather(k->densityNm1, 1, nodes, v_densityNm1);
  gather(k->densityN, 1, nodes, v_densityN);
  gather(k->density, 1, nodes, v_densityNp1);
  double v_shape_function_[8][8];
  scv_shape(k->lumpedMass, v_shape_function_);
  const int *ipNodeMap_ = orc_hex8_scv_ip_node_map();

  for (int ip = 0; ip < 8; ++ip) {
    const int nearestNode = ipNodeMap_[ip];
    double rhoNm1 = 0.0, rhoN = 0.0, rhoNp1 = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      rhoNm1 += r * v_densityNm1[ic];
      rhoN += r * v_densityN[ic];
      rhoNp1 += r * v_densityNp1[ic];
    }
    const double scV = me->scv_volume[ip];
    rhs[nearestNode] += -(gamma1_ * rhoNp1 + gamma2_ * rhoN + gamma3_ * rhoNm1) * scV / dt_ / projTimeScale;
  }
}

/* ---- src/kernel/MomentumBuoyancySrcElemKernel.C:61-90 and
 *      src/kernel/MomentumBuoyancyBoussinesqSrcElemKernel.C:66-94 ---- */
static void momentum_buoyancy(const orc_kernel_desc *k, int boussinesq, const int32_t *nodes, const me_views *me,
                              double *lhs, double *rhs)
{
  (void)lhs; /* No LHS contributions */
  double v_s[8];
  gather(boussinesq ? k->temperature : k->density, 1, nodes, v_s);
  double v_shape_function_[8][8];
  scv_shape(0, v_shape_function_); /* meSCV->shape_fcn: never shifted (ctor :43 / :50) */
  const int *ipNodeMap_ = orc_hex8_scv_ip_node_map();
  for (int ip = 0; ip < 8; ++ip) {
    const int nearestNode = ipNodeMap_[ip];
    double sIp = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      sIp += r * v_s[ic];
    }
    const double scV = me->scv_volume[ip];
    const int nnNdim = nearestNode * 3;
    const double fac = boussinesq ? -k->rhoRef * k->beta * (sIp - k->tRef) * scV : (sIp - k->rhoRef) * scV;
    for (int j = 0; j < 3; ++j) rhs[nnNdim + j] += fac * k->gravity[j];
  }
}

/* ---- src/kernel/TurbKineticEnergySSTSrcElemKernel.C:95-174 (sdr_kernel = 0) and
 *      src/kernel/SpecificDissipationRateSSTSrcElemKernel.C:100-204 (sdr_kernel = 1) ---- */
static void sst_src(const orc_kernel_desc *k, int sdr_kernel, const int32_t *nodes, const me_views *me, double *lhs,
                    double *rhs)
{
  double v_tkeNp1[8], v_sdrNp1[8], v_densityNp1[8], v_velocityNp1[8][3], v_tvisc[8], v_fOneBlend[8];
  gather(k->tke, 1, nodes, v_tkeNp1);
  gather(k->sdr, 1, nodes, v_sdrNp1);
  gather(k->density, 1, nodes, v_densityNp1);
  gather(k->velocity, 3, nodes, &v_velocityNp1[0][0]);
  gather(k->tvisc, 1, nodes, v_tvisc);
  if (sdr_kernel) gather(k->fOneBlend, 1, nodes, v_fOneBlend);
  double v_shape_function_[8][8];
  scv_shape(k->lumpedMass, v_shape_function_);
  const int *ipNodeMap_ = orc_hex8_scv_ip_node_map();
  const double betaStar_ = k->betaStar, tkeProdLimitRatio_ = k->tkeProdLimitRatio;

  for (int ip = 0; ip < 8; ++ip) {
    const int nearestNode = ipNodeMap_[ip];
    const double scV = me->scv_volume[ip];
    double rho = 0.0, tke = 0.0, sdr = 0.0, tvisc = 0.0, fOneBlend = 0.0;
    double w_dudx[3][3], w_dkdx[3], w_dwdx[3];
    for (int i = 0; i < 3; ++i) {
      w_dkdx[i] = 0.0;
      w_dwdx[i] = 0.0;
      for (int j = 0; j < 3; ++j) w_dudx[i][j] = 0.0;
    }
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      rho += r * v_densityNp1[ic];
      tke += r * v_tkeNp1[ic];
      sdr += r * v_sdrNp1[ic];
      tvisc += r * v_tvisc[ic];
      if (sdr_kernel) fOneBlend += r * v_fOneBlend[ic];
      for (int i = 0; i < 3; ++i) {
        const double dni = me->dndx_scv[ip][ic][i];
        const double ui = v_velocityNp1[ic][i];
        w_dkdx[i] += dni * v_tkeNp1[ic];
        w_dwdx[i] += dni * v_sdrNp1[ic];
        for (int j = 0; j < 3; ++j) w_dudx[i][j] += me->dndx_scv[ip][ic][j] * ui;
      }
    }
    double Pk = 0.0, crossDiff = 0.0;
    for (int i = 0; i < 3; ++i) {
      crossDiff += w_dkdx[i] * w_dwdx[i];
      for (int j = 0; j < 3; ++j) Pk += w_dudx[i][j] * (w_dudx[i][j] + w_dudx[j][i]);
    }
    Pk *= tvisc;
    if (!sdr_kernel) {
      const double tkeFac = betaStar_ * rho * sdr;
      const double Dk = tkeFac * tke;
      Pk = fmin(Pk, tkeProdLimitRatio_ * Dk);
      rhs[nearestNode] += (Pk - Dk) * scV;
      for (int ic = 0; ic < 8; ++ic) lhs[nearestNode * 8 + ic] += v_shape_function_[ip][ic] * tkeFac * scV;
    } else {
      const double Dk = betaStar_ * rho * sdr * tke;
      Pk = fmin(Pk, tkeProdLimitRatio_ * Dk);
      const double om_fOneBlend = 1.0 - fOneBlend;
      const double beta = fOneBlend * k->betaOne + om_fOneBlend * k->betaTwo;
      const double gamma = fOneBlend * k->gammaOne + om_fOneBlend * k->gammaTwo;
      const double sigmaD = 2.0 * om_fOneBlend * k->sigmaWTwo;
      const double Pw = gamma * rho * Pk / fmax(tvisc, 1.0e-16);
      const double Dw = beta * rho * sdr * sdr;
      const double Sw = sigmaD * rho * crossDiff / sdr;
      rhs[nearestNode] += (Pw - Dw + Sw) * scV;
      for (int ic = 0; ic < 8; ++ic)
        lhs[nearestNode * 8 + ic] += v_shape_function_[ip][ic] * (2.0 * beta * rho * sdr + fmax(Sw / sdr, 0.0)) * scV;
    }
  }
}

/* ---- src/kernel/TurbKineticEnergySSTDESSrcElemKernel.C:106-198 ---- */
static void sst_des_src(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me, double *lhs, double *rhs)
{
  double v_tkeNp1[8], v_sdrNp1[8], v_densityNp1[8], v_velocityNp1[8][3], v_tvisc[8], v_maxLengthScale[8], v_fOneBlend[8];
  gather(k->tke, 1, nodes, v_tkeNp1);
  gather(k->sdr, 1, nodes, v_sdrNp1);
  gather(k->density, 1, nodes, v_densityNp1);
  gather(k->velocity, 3, nodes, &v_velocityNp1[0][0]);
  gather(k->tvisc, 1, nodes, v_tvisc);
  gather(k->maxLengthScale, 1, nodes, v_maxLengthScale);
  gather(k->fOneBlend, 1, nodes, v_fOneBlend);
  double v_shape_function_[8][8];
  scv_shape(k->lumpedMass, v_shape_function_);
  const int *ipNodeMap_ = orc_hex8_scv_ip_node_map();
  const double betaStar_ = k->betaStar, tkeProdLimitRatio_ = k->tkeProdLimitRatio, cDESkw_ = k->cDESkw, cDESke_ = k->cDESke;

  for (int ip = 0; ip < 8; ++ip) {
    const int nearestNode = ipNodeMap_[ip];
    const double scV = me->scv_volume[ip];
    double rho = 0.0, tke = 0.0, sdr = 0.0, tvisc = 0.0, maxLengthScale = 0.0, fOneBlend = 0.0;
    double w_dudx[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) w_dudx[i][j] = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      rho += r * v_densityNp1[ic];
      tke += r * v_tkeNp1[ic];
      sdr += r * v_sdrNp1[ic];
      tvisc += r * v_tvisc[ic];
      maxLengthScale += r * v_maxLengthScale[ic];
      fOneBlend += r * v_fOneBlend[ic];
      for (int i = 0; i < 3; ++i) {
        const double ui = v_velocityNp1[ic][i];
        for (int j = 0; j < 3; ++j) w_dudx[i][j] += me->dndx_scv[ip][ic][j] * ui;
      }
    }
    double Pk = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Pk += w_dudx[i][j] * (w_dudx[i][j] + w_dudx[j][i]);
    Pk *= tvisc;
    const double cDES = fOneBlend * cDESkw_ + (1.0 - fOneBlend) * cDESke_;
    const double sqrtTke = sqrt(tke);
    const double lSST = sqrtTke / betaStar_ / sdr;
    const double lDES = fmax(1.0e-16, fmin(lSST, cDES * maxLengthScale));
    const double tkeFac = rho * sqrtTke / lDES;
    const double Dk = tkeFac * tke;
    Pk = fmin(Pk, tkeProdLimitRatio_ * Dk);
    rhs[nearestNode] += (Pk - Dk) * scV;
    for (int ic = 0; ic < 8; ++ic) lhs[nearestNode * 8 + ic] += v_shape_function_[ip][ic] * tkeFac * scV;
  }
}

/* ---- src/kernel/TurbKineticEnergyKsgsSrcElemKernel.C:78-128: everything at the nearest node, lumped ---- */
static void ksgs_src(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me, double *lhs, double *rhs)
{
  double v_tkeNp1[8], v_densityNp1[8], v_tvisc[8], v_dualNodalVolume[8], v_cEps[8], v_Gju[8][3][3], v_visc[8], v_dsqrtkSq[8];
  gather(k->tke, 1, nodes, v_tkeNp1);
  gather(k->density, 1, nodes, v_densityNp1);
  gather(k->tvisc, 1, nodes, v_tvisc);
  gather(k->dualNodalVolume, 1, nodes, v_dualNodalVolume);
  gather(k->cEps, 1, nodes, v_cEps);
  gather(k->Gju, 9, nodes, &v_Gju[0][0][0]);
  gather(k->viscosity, 1, nodes, v_visc);
  gather(k->dsqrtkSq, 1, nodes, v_dsqrtkSq);
  const int *ipNodeMap_ = orc_hex8_scv_ip_node_map();
  const double lrksgsfac_ = k->lrksgsfac, tkeProdLimitRatio_ = k->tkeProdLimitRatio;
  for (int ip = 0; ip < 8; ++ip) {
    const int nearestNode = ipNodeMap_[ip];
    double Pk = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Pk += v_Gju[nearestNode][i][j] * (v_Gju[nearestNode][i][j] + v_Gju[nearestNode][j][i]);
    Pk *= v_tvisc[nearestNode];
    const double tke = v_tkeNp1[nearestNode];
    const double tkeFac = v_cEps[nearestNode] * v_densityNp1[nearestNode] * sqrt(tke) / cbrt(v_dualNodalVolume[nearestNode]);
    const double Dk = tkeFac * tke + lrksgsfac_ * 2.0 * v_visc[nearestNode] * v_dsqrtkSq[nearestNode];
    Pk = fmin(Pk, tkeProdLimitRatio_ * Dk);
    const double scvol = me->scv_volume[ip];
    rhs[nearestNode] += (Pk - Dk) * scvol;
    lhs[nearestNode * 8 + nearestNode] += 1.5 * tkeFac * scvol;
  }
}

/* ---- src/nso/MomentumNSOElemKernel.C:127-310 ---- */
static void momentum_nso(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me, double *lhs, double *rhs)
{
  double v_Gju[8][3][3], v_uNm1[8][3], v_uN[8][3], v_uNp1[8][3], v_velocityRTM[8][3], v_rhoNm1[8], v_rhoN[8], v_rhoNp1[8],
      v_viscosity[8], v_pressure[8];
  gather(k->Gju, 9, nodes, &v_Gju[0][0][0]);
  gather(k->velocityNm1, 3, nodes, &v_uNm1[0][0]);
  gather(k->velocityN, 3, nodes, &v_uN[0][0]);
  gather(k->velocity, 3, nodes, &v_uNp1[0][0]);
  gather(k->velocityRTM ? k->velocityRTM : k->velocity, 3, nodes, &v_velocityRTM[0][0]);
  gather(k->densityNm1, 1, nodes, v_rhoNm1);
  gather(k->densityN, 1, nodes, v_rhoN);
  gather(k->density, 1, nodes, v_rhoNp1);
  gather(k->viscosity, 1, nodes, v_viscosity);
  gather(k->pressure, 1, nodes, v_pressure);
  const double (*v_dndx)[8][3] = k->shiftedGradOp ? me->dndx_shifted : me->dndx;
  double v_shape_function_[12][8];
  scs_shape(0, v_shape_function_);
  const int *lrscv_ = orc_hex8_lrscv();
  const double gamma1_ = k->gamma1, gamma2_ = k->gamma2, gamma3_ = k->gamma3, dt_ = k->dt;
  const double fourthFac_ = k->fourthFac, altResFac_ = k->altResFac, om_altResFac_ = 1.0 - k->altResFac;
  const double includeDivU_ = k->includeDivU, Cupw_ = 0.1, nonConservedForm_ = 0.0, small_ = 1.0e-16;
  double w_dukdxScs[3], w_rhoVrtmScs[3], w_dpdxScs[3];

  for (int ip = 0; ip < 12; ++ip) {
    const int il = lrscv_[2 * ip];
    const int ir = lrscv_[2 * ip + 1];
    const int ilNdim = il * 3;
    const int irNdim = ir * 3;

    double rhoNm1Scs = 0.0, rhoNScs = 0.0, rhoNp1Scs = 0.0, dFdxCont = 0.0, divU = 0.0;
    for (int i = 0; i < 3; ++i) {
      w_rhoVrtmScs[i] = 0.0;
      w_dpdxScs[i] = 0.0;
    }
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      rhoNm1Scs += r * v_rhoNm1[ic];
      rhoNScs += r * v_rhoN[ic];
      rhoNp1Scs += r * v_rhoNp1[ic];
      const double pIC = v_pressure[ic];
      const double rhoIC = v_rhoNp1[ic];
      for (int j = 0; j < 3; ++j) {
        const double dnj = v_dndx[ip][ic][j];
        const double vrtmj = v_velocityRTM[ic][j];
        w_rhoVrtmScs[j] += r * rhoIC * vrtmj;
        divU += r * v_Gju[ic][j][j];
        dFdxCont += rhoIC * vrtmj * dnj;
        w_dpdxScs[j] += pIC * dnj;
      }
    }
    const double contRes = (gamma1_ * rhoNp1Scs + gamma2_ * rhoNScs + gamma3_ * rhoNm1Scs) / dt_ + dFdxCont;
    const double twoThirds = 2.0 / 3.0;

    for (int kk = 0; kk < 3; ++kk) {
      const int indexL = ilNdim + kk;
      const int indexR = irNdim + kk;
      double dFdxkAdv = 0.0, dFdxkDiff = 0.0, ukNm1Scs = 0.0, ukNScs = 0.0, ukNp1Scs = 0.0;
      for (int j = 0; j < 3; ++j) w_dukdxScs[j] = 0.0;

      for (int ic = 0; ic < 8; ++ic) {
        const double r = v_shape_function_[ip][ic];
        const double ukNp1 = v_uNp1[ic][kk];
        ukNm1Scs += r * v_uNm1[ic][kk];
        ukNScs += r * v_uN[ic][kk];
        ukNp1Scs += r * ukNp1;
        const double rhoIC = v_rhoNp1[ic];
        const double viscIC = v_viscosity[ic];
        for (int j = 0; j < 3; ++j) {
          const double dnj = v_dndx[ip][ic][j];
          const double kd = (kk == j) ? 1.0 : 0.0;
          w_dukdxScs[j] += ukNp1 * dnj;
          dFdxkAdv += rhoIC * v_velocityRTM[ic][j] * ukNp1 * dnj;
          dFdxkDiff += viscIC * (v_Gju[ic][kk][j] + v_Gju[ic][j][kk] - twoThirds * divU * kd * includeDivU_) * dnj;
        }
      }

      double residualAlt = dFdxkAdv - ukNp1Scs * dFdxCont;
      for (int j = 0; j < 3; ++j) residualAlt -= w_rhoVrtmScs[j] * w_dukdxScs[j];

      const double time = (gamma1_ * rhoNp1Scs * ukNp1Scs + gamma2_ * rhoNScs * ukNScs + gamma3_ * rhoNm1Scs * ukNm1Scs) / dt_;
      const double residualPde = time + dFdxkAdv - dFdxkDiff + w_dpdxScs[kk] - contRes * ukNp1Scs * nonConservedForm_;
      const double residual = residualAlt * altResFac_ + residualPde * om_altResFac_;

      double gUpperMagGradQ = 0.0, rhoVrtmiGLowerRhoVrtmj = 0.0;
      for (int i = 0; i < 3; ++i) {
        const double duidxScs = w_dukdxScs[i];
        const double rhoVrtmi = w_rhoVrtmScs[i];
        for (int j = 0; j < 3; ++j) {
          gUpperMagGradQ += duidxScs * me->gijUpper[ip][i][j] * w_dukdxScs[j];
          rhoVrtmiGLowerRhoVrtmj += rhoVrtmi * me->gijLower[ip][i][j] * w_rhoVrtmScs[j];
        }
      }
      const double nuResidual = sqrt((residual * residual) / (gUpperMagGradQ + small_));
      const double nuFirstOrder = sqrt(rhoVrtmiGLowerRhoVrtmj);
      const double nu = fmin(Cupw_ * nuFirstOrder, nuResidual);

      double gijFac = 0.0;
      for (int ic = 0; ic < 8; ++ic) {
        const double r = v_shape_function_[ip][ic];
        const int icNdim = ic * 3;
        const double ukNp1 = v_uNp1[ic][kk];
        double lhsfac = 0.0;
        for (int i = 0; i < 3; ++i) {
          const double axi = me->scs_areav[ip][i];
          for (int j = 0; j < 3; ++j) {
            const double fac = me->gijUpper[ip][i][j] * v_dndx[ip][ic][j] * axi;
            const double facGj = r * me->gijUpper[ip][i][j] * v_Gju[ic][kk][j] * axi;
            gijFac = fma(fac, ukNp1, gijFac) - facGj * fourthFac_;
            lhsfac -= fac;
          }
        }
        lhs[indexL * 24 + icNdim + kk] = fma(nu, lhsfac, lhs[indexL * 24 + icNdim + kk]);
        lhs[indexR * 24 + icNdim + kk] -= nu * lhsfac;
      }
      const double residualNSO = -nu * gijFac;
      rhs[indexL] -= residualNSO;
      rhs[indexR] += residualNSO;
    }
  }
}

/* ---- src/kernel/ScalarMassElemKernel.C:96-156 ---- */
static void scalar_mass(const orc_kernel_desc *k, const int32_t *nodes, const me_views *me,
                        double *lhs, double *rhs)
{
  const int n = 8;
  const double gamma1_ = k->gamma1, gamma2_ = k->gamma2, gamma3_ = k->gamma3, dt_ = k->dt;
  double v_qNm1[8], v_qN[8], v_qNp1[8], v_rhoNm1[8], v_rhoN[8], v_rhoNp1[8];
  gather(k->scalarQNm1, 1, nodes, v_qNm1);
  gather(k->scalarQN, 1, nodes, v_qN);
  gather(k->scalarQ, 1, nodes, v_qNp1);
  gather(k->densityNm1, 1, nodes, v_rhoNm1);
  gather(k->densityN, 1, nodes, v_rhoN);
  gather(k->density, 1, nodes, v_rhoNp1);
  double v_shape_function_[8][8];
  scv_shape(k->lumpedMass, v_shape_function_);
  const int *ipNodeMap_ = orc_hex8_scv_ip_node_map();

  for (int ip = 0; ip < 8; ++ip) {
    const int nearestNode = ipNodeMap_[ip];
    double qNm1Scv = 0.0, qNScv = 0.0, qNp1Scv = 0.0;
    double rhoNm1Scv = 0.0, rhoNScv = 0.0, rhoNp1Scv = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      qNm1Scv += r * v_qNm1[ic];
      qNScv += r * v_qN[ic];
      qNp1Scv += r * v_qNp1[ic];
      rhoNm1Scv += r * v_rhoNm1[ic];
      rhoNScv += r * v_rhoN[ic];
      rhoNp1Scv += r * v_rhoNp1[ic];
    }
    const double scV = me->scv_volume[ip];
    rhs[nearestNode] +=
        -(gamma1_ * rhoNp1Scv * qNp1Scv + gamma2_ * rhoNScv * qNScv + gamma3_ * rhoNm1Scv * qNm1Scv) * scV / dt_;
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      const double lhsfac = r * gamma1_ * rhoNp1Scv * scV / dt_;
      lhs[nearestNode * n + ic] += lhsfac;
    }
  }
}

/* ---- src/PecletFunction.C:67-71 (classic), :98-101 (tanh); built by EquationSystem::create_peclet_function
 *      (include/EquationSystem.h:306-325).  No reference unit test pins these two kernels: the restatement is
 *      anchored by its reduction to the golden tables of the central kernels (alpha = 0, hybridFactor = 0 makes
 *      pecfac = 0 and every upwind term vanish, tests/test_oracle_golden.py) -- the upwind / limiter branches
 *      themselves are PARITY-UNPINNED restatements. ---- */
static double peclet_function(const orc_kernel_desc *k, double x)
{
  if (k->pecletTanh) return 0.50 * (1.0 + tanh((x - k->tanhTrans) / k->tanhWidth));
  const double modPeclet = k->hybridFactor * x;
  return modPeclet * modPeclet / (5.0 + modPeclet * modPeclet);
}
/* ScalarUpwAdvDiffElemKernel.C:258-266 == MomentumUpwAdvDiffElemKernel.C:314-322 */
static double van_leer(double dqm, double dqp)
{
  const double small_ = 1.0e-16;
  return (2.0 * (dqm * dqp + fabs(dqm * dqp))) / ((dqm + dqp) * (dqm + dqp) + small_);
}

/* ---- src/kernel/ScalarUpwAdvDiffElemKernel.C:110-256 ---- */
static void scalar_upw_adv_diff(const orc_kernel_desc *k, const int32_t *nodes, int64_t e,
                                const me_views *me, double *lhs, double *rhs)
{
  const int n = 8;
  const double small_ = 1.0e-16;
  const double alpha_ = k->alpha, alphaUpw_ = k->alphaUpw, hoUpwind_ = k->hoUpwind;
  const double om_alpha_ = 1.0 - alpha_, om_alphaUpw_ = 1.0 - alphaUpw_;
  double v_velocityRTM[8][3], v_Gjq[8][3], v_scalarQ[8], v_density[8], v_diffFluxCoeff[8];
  const double(*v_coordinates)[3] = me->coords;
  const double *v_mdot = k->massFlowRate + 12 * e;
  gather(k->velocityRTM ? k->velocityRTM : k->velocity, 3, nodes, &v_velocityRTM[0][0]);
  gather(k->Gjq, 3, nodes, &v_Gjq[0][0]);
  gather(k->scalarQ, 1, nodes, v_scalarQ);
  gather(k->density, 1, nodes, v_density);
  gather(k->viscosity, 1, nodes, v_diffFluxCoeff);
  double v_shape_function_[12][8], v_adv_shape_function_[12][8];
  scs_shape(0, v_shape_function_);
  scs_shape(k->skewSymmetric, v_adv_shape_function_);
  const double(*v_dndx)[8][3] = k->shiftedGradOp ? me->dndx_shifted : me->dndx;
  const int *lrscv = orc_hex8_lrscv();
  double w_coordIp[3];

  for (int ip = 0; ip < 12; ++ip) {
    const int il = lrscv[2 * ip];
    const int ir = lrscv[2 * ip + 1];
    const double tmdot = v_mdot[ip];
    for (int j = 0; j < 3; ++j) w_coordIp[j] = 0.0;

    double qIp = 0.0;
    double diffFluxCoeffIp = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      const double rAdv = v_adv_shape_function_[ip][ic];
      qIp += rAdv * v_scalarQ[ic];
      diffFluxCoeffIp += r * v_diffFluxCoeff[ic];
      for (int i = 0; i < 3; ++i) w_coordIp[i] += rAdv * v_coordinates[ic][i];
    }

    /* Peclet factor; along the edge */
    const double diffIp = 0.5 * (v_diffFluxCoeff[il] / v_density[il] + v_diffFluxCoeff[ir] / v_density[ir]);
    double udotx = 0.0;
    for (int j = 0; j < 3; ++j) {
      const double dxj = v_coordinates[ir][j] - v_coordinates[il][j];
      const double uj = 0.5 * (v_velocityRTM[il][j] + v_velocityRTM[ir][j]);
      udotx += uj * dxj;
    }
    const double tmp = fabs(udotx) / (diffIp + small_);
    const double pecfac = peclet_function(k, tmp);
    const double om_pecfac = 1.0 - pecfac;

    /* left and right extrapolation */
    double dqL = 0.0;
    double dqR = 0.0;
    for (int j = 0; j < 3; ++j) {
      const double dxjL = w_coordIp[j] - v_coordinates[il][j];
      const double dxjR = v_coordinates[ir][j] - w_coordIp[j];
      dqL += dxjL * v_Gjq[il][j];
      dqR += dxjR * v_Gjq[ir][j];
    }

    double limitL = 1.0;
    double limitR = 1.0;
    if (k->useLimiter) {
      const double dq = v_scalarQ[ir] - v_scalarQ[il];
      const double dqMl = 2.0 * 2.0 * dqL - dq;
      const double dqMr = 2.0 * 2.0 * dqR - dq;
      limitL = van_leer(dqMl, dq);
      limitR = van_leer(dqMr, dq);
    }

    const double qIpL = v_scalarQ[il] + dqL * hoUpwind_ * limitL;
    const double qIpR = v_scalarQ[ir] - dqR * hoUpwind_ * limitR;

    const double qUpwind = tmdot > 0 ? alphaUpw_ * qIpL + om_alphaUpw_ * qIp : alphaUpw_ * qIpR + om_alphaUpw_ * qIp;

    const double qHatL = alpha_ * qIpL + om_alpha_ * qIp;
    const double qHatR = alpha_ * qIpR + om_alpha_ * qIp;
    const double qCds = 0.5 * (qHatL + qHatR);

    const double aflux = tmdot * (pecfac * qUpwind + om_pecfac * qCds);

    rhs[il] -= aflux;
    rhs[ir] += aflux;

    const double alhsfacL = 0.5 * (tmdot + fabs(tmdot)) * pecfac * alphaUpw_ + 0.5 * alpha_ * om_pecfac * tmdot;
    lhs[il * n + il] += alhsfacL;
    lhs[ir * n + il] -= alhsfacL;

    const double alhsfacR = 0.5 * (tmdot - fabs(tmdot)) * pecfac * alphaUpw_ + 0.5 * alpha_ * om_pecfac * tmdot;
    lhs[ir * n + ir] -= alhsfacR;
    lhs[il * n + ir] += alhsfacR;

    double qDiff = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double lhsfacAdv = v_adv_shape_function_[ip][ic] * tmdot * (pecfac * om_alphaUpw_ + om_pecfac * om_alpha_);
      lhs[il * n + ic] += lhsfacAdv;
      lhs[ir * n + ic] -= lhsfacAdv;

      double lhsfacDiff = 0.0;
      for (int j = 0; j < 3; ++j) lhsfacDiff += -diffFluxCoeffIp * v_dndx[ip][ic][j] * me->scs_areav[ip][j];
      qDiff += lhsfacDiff * v_scalarQ[ic];

      lhs[il * n + ic] += lhsfacDiff;
      lhs[ir * n + ic] -= lhsfacDiff;
    }

    rhs[il] -= qDiff;
    rhs[ir] += qDiff;
  }
}

/* ---- src/kernel/MomentumUpwAdvDiffElemKernel.C:108-311 ---- */
static void momentum_upw_adv_diff(const orc_kernel_desc *k, const int32_t *nodes, int64_t e,
                                  const me_views *me, double *lhs, double *rhs)
{
  const int n = 24;
  const double small_ = 1.0e-16;
  const double alpha_ = k->alpha, alphaUpw_ = k->alphaUpw, hoUpwind_ = k->hoUpwind;
  const double om_alpha_ = 1.0 - alpha_, om_alphaUpw_ = 1.0 - alphaUpw_;
  double w_uIp[3], w_uIpL[3], w_uIpR[3], w_coordIp[3], w_duL[3], w_duR[3], w_limitL[3], w_limitR[3];
  for (int i = 0; i < 3; ++i) {
    w_limitL[i] = 1.0;
    w_limitR[i] = 1.0;
  }
  double v_Gju[8][3][3], v_uNp1[8][3], v_density[8], v_viscosity[8], v_velocityRTM[8][3];
  const double(*v_coordinates)[3] = me->coords;
  const double *v_mdot = k->massFlowRate + 12 * e;
  gather(k->Gju, 9, nodes, &v_Gju[0][0][0]);
  gather(k->velocity, 3, nodes, &v_uNp1[0][0]);
  gather(k->density, 1, nodes, v_density);
  gather(k->viscosity, 1, nodes, v_viscosity);
  gather(k->velocityRTM ? k->velocityRTM : k->velocity, 3, nodes, &v_velocityRTM[0][0]);
  double v_shape_function_[12][8], v_adv_shape_function_[12][8];
  scs_shape(0, v_shape_function_);
  scs_shape(k->skewSymmetric, v_adv_shape_function_);
  const double(*v_dndx)[8][3] = k->shiftedGradOp ? me->dndx_shifted : me->dndx;
  const int *lrscv = orc_hex8_lrscv();

  for (int ip = 0; ip < 12; ++ip) {
    const int il = lrscv[2 * ip];
    const int ir = lrscv[2 * ip + 1];
    const int ilNdim = il * 3;
    const int irNdim = ir * 3;
    const double tmdot = v_mdot[ip];

    double muIp = 0.0;
    double divU = 0.0;
    for (int i = 0; i < 3; ++i) {
      w_uIp[i] = 0.0;
      w_coordIp[i] = 0.0;
    }
    for (int ic = 0; ic < 8; ++ic) {
      const double r = v_shape_function_[ip][ic];
      const double rAdv = v_adv_shape_function_[ip][ic];
      muIp += r * v_viscosity[ic];
      for (int j = 0; j < 3; ++j) {
        const double uj = v_uNp1[ic][j];
        w_uIp[j] += rAdv * uj;
        w_coordIp[j] += rAdv * v_coordinates[ic][j];
        divU += uj * v_dndx[ip][ic][j];
      }
    }

    double udotx = 0.0;
    for (int i = 0; i < 3; ++i) {
      const double dxi = v_coordinates[ir][i] - v_coordinates[il][i];
      const double ui = 0.5 * (v_velocityRTM[il][i] + v_velocityRTM[ir][i]);
      udotx += ui * dxi;
      w_duL[i] = 0.0;
      w_duR[i] = 0.0;
      for (int j = 0; j < 3; ++j) {
        const double dxjL = w_coordIp[j] - v_coordinates[il][j];
        const double dxjR = v_coordinates[ir][j] - w_coordIp[j];
        w_duL[i] += dxjL * v_Gju[il][i][j];
        w_duR[i] += dxjR * v_Gju[ir][i][j];
      }
    }

    const double diffIp = 0.5 * (v_viscosity[il] / v_density[il] + v_viscosity[ir] / v_density[ir]);
    const double pecFuncArg = fabs(udotx) / (diffIp + small_);
    const double pecfac = peclet_function(k, pecFuncArg);
    const double om_pecfac = 1.0 - pecfac;

    if (k->useLimiter) {
      for (int i = 0; i < 3; ++i) {
        const double dq = v_uNp1[ir][i] - v_uNp1[il][i];
        const double dqMl = 2.0 * 2.0 * w_duL[i] - dq;
        const double dqMr = 2.0 * 2.0 * w_duR[i] - dq;
        w_limitL[i] = van_leer(dqMl, dq);
        w_limitR[i] = van_leer(dqMr, dq);
      }
    }

    for (int i = 0; i < 3; ++i) {
      w_uIpL[i] = v_uNp1[il][i] + w_duL[i] * hoUpwind_ * w_limitL[i];
      w_uIpR[i] = v_uNp1[ir][i] - w_duR[i] * hoUpwind_ * w_limitR[i];
    }

    for (int i = 0; i < 3; ++i) {
      const double uiIp = w_uIp[i];
      const double uiUpwind =
          tmdot > 0 ? alphaUpw_ * w_uIpL[i] + om_alphaUpw_ * uiIp : alphaUpw_ * w_uIpR[i] + om_alphaUpw_ * uiIp;
      const double uiHatL = alpha_ * w_uIpL[i] + om_alpha_ * uiIp;
      const double uiHatR = alpha_ * w_uIpR[i] + om_alpha_ * uiIp;
      const double uiCds = 0.5 * (uiHatL + uiHatR);
      const double aflux = tmdot * (pecfac * uiUpwind + om_pecfac * uiCds);
      const double divUstress = 2.0 / 3.0 * muIp * divU * me->scs_areav[ip][i] * k->includeDivU;
      const int indexL = ilNdim + i;
      const int indexR = irNdim + i;
      rhs[indexL] -= aflux + divUstress;
      rhs[indexR] += aflux + divUstress;

      const double alhsfacL = 0.5 * (tmdot + fabs(tmdot)) * pecfac * alphaUpw_ + 0.5 * alpha_ * om_pecfac * tmdot;
      lhs[indexL * n + indexL] += alhsfacL;
      lhs[indexR * n + indexL] -= alhsfacL;

      const double alhsfacR = 0.5 * (tmdot - fabs(tmdot)) * pecfac * alphaUpw_ + 0.5 * alpha_ * om_pecfac * tmdot;
      lhs[indexR * n + indexR] -= alhsfacR;
      lhs[indexL * n + indexR] += alhsfacR;
    }

    for (int ic = 0; ic < 8; ++ic) {
      const int icNdim = ic * 3;
      const double lhsfacAdv = v_adv_shape_function_[ip][ic] * tmdot * (pecfac * om_alphaUpw_ + om_pecfac * om_alpha_);
      for (int i = 0; i < 3; ++i) {
        const int indexL = ilNdim + i;
        const int indexR = irNdim + i;
        lhs[indexL * n + icNdim + i] += lhsfacAdv;
        lhs[indexR * n + icNdim + i] -= lhsfacAdv;

        double lhs_riC_i = 0.0;
        for (int j = 0; j < 3; ++j) {
          const double axj = me->scs_areav[ip][j];
          const double uj = v_uNp1[ic][j];
          const double lhsfacDiff_i = -muIp * v_dndx[ip][ic][j] * axj;
          lhs_riC_i += lhsfacDiff_i;
          const double lhsfacDiff_j = -muIp * v_dndx[ip][ic][i] * axj;
          lhs[indexL * n + icNdim + j] += lhsfacDiff_j;
          lhs[indexR * n + icNdim + j] -= lhsfacDiff_j;
          rhs[indexL] -= lhsfacDiff_j * uj;
          rhs[indexR] += lhsfacDiff_j * uj;
        }
        lhs[indexL * n + icNdim + i] += lhs_riC_i;
        lhs[indexR * n + icNdim + i] -= lhs_riC_i;
        const double ui = v_uNp1[ic][i];
        rhs[indexL] -= lhs_riC_i * ui;
        rhs[indexR] += lhs_riC_i * ui;
      }
    }
  }
}

/* which ME quantities each kernel registers in its constructor */
static void me_needs(const orc_kernel_desc *k, int *areav, int *dndx, int *dndxs, int *scv)
{
  switch (k->kind) {
  case ORC_MOMENTUM_ADV_DIFF: /* MomentumAdvDiffElemKernel.C:58-63 */
  case ORC_SCALAR_ADV_DIFF:   /* ScalarAdvDiffElemKernel.C:57-62   */
  case ORC_SCALAR_DIFF:
  case ORC_SCALAR_UPW_ADV_DIFF:   /* ScalarUpwAdvDiffElemKernel.C:82-89   */
  case ORC_MOMENTUM_UPW_ADV_DIFF: /* MomentumUpwAdvDiffElemKernel.C:82-89 */
    *areav = 1;
    if (k->shiftedGradOp) *dndxs = 1; else *dndx = 1;
    break;
  case ORC_MOMENTUM_NSO: /* MomentumNSOElemKernel.C:98-104: SCS_AREAV, SCS_(SHIFTED_)GRAD_OP, SCS_GIJ */
    *areav = 1;
    if (k->shiftedGradOp) *dndxs = 1; else *dndx = 1;
    break;
  case ORC_CONTINUITY_ADV: /* ContinuityAdvElemKernel.C:59-65 */
    *areav = 1;
    if (!k->shiftPoisson || !k->reducedSensitivities) *dndx = 1;
    if (k->shiftPoisson || k->reducedSensitivities) *dndxs = 1;
    break;
  case ORC_MOMENTUM_MASS:
  case ORC_CONTINUITY_MASS:
  case ORC_SCALAR_MASS:
  case ORC_MOMENTUM_BUOYANCY_SRC:
  case ORC_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC:
  case ORC_TKE_KSGS_SRC: /* TurbKineticEnergyKsgsSrcElemKernel.C:64 */
    *scv = 1;
    break;
  default:
    break;
  }
}

/* src/AssembleElemSolverAlgorithm.C:74-100 + include/AssembleElemSolverAlgorithm.h:52-105
 * + fill_master_element_views_new_me (include/ScratchViews.h:476-582), one element */
void orc_element_lhs_rhs(const orc_kernel_desc *kernels, int nk, int numDof,
                         const int32_t *elemNodes, int64_t elemIndex, double *lhs, double *rhs)
{
  const int rhsSize = 8 * numDof;
  init_tables();
  me_views me;
  memset(&me, 0, sizeof(me));
  int areav = 0, dndx = 0, dndxs = 0, scv = 0;
  const double *coordField = NULL;
  for (int i = 0; i < nk; ++i) {
    me_needs(&kernels[i], &areav, &dndx, &dndxs, &scv);
    if (kernels[i].coordinates) coordField = kernels[i].coordinates;
  }
  if (coordField) {
    gather(coordField, 3, elemNodes, &me.coords[0][0]);
    if (areav) orc_hex8_scs_areav(&me.coords[0][0], &me.scs_areav[0][0]);
    if (dndx) orc_hex8_scs_grad_op(0, &me.coords[0][0], &me.dndx[0][0][0], &me.deriv[0][0][0]);
    if (dndxs) orc_hex8_scs_grad_op(1, &me.coords[0][0], &me.dndx_shifted[0][0][0], &me.deriv[0][0][0]);
    if (scv) orc_hex8_scv_volume(&me.coords[0][0], me.scv_volume);
    for (int i = 0; i < nk; ++i)
      if (kernels[i].kind == ORC_MOMENTUM_NSO) orc_hex8_scs_gij(&me.coords[0][0], &me.gijUpper[0][0][0], &me.gijLower[0][0][0]);
    for (int i = 0; i < nk; ++i)
      if (kernels[i].kind == ORC_TKE_SST_SRC || kernels[i].kind == ORC_SDR_SST_SRC || kernels[i].kind == ORC_TKE_SST_DES_SRC) {
        /* SCV_GRAD_OP or SCV_SHIFTED_GRAD_OP (TurbKineticEnergySSTSrcElemKernel.C:84-89) */
        orc_hex8_scv_volume(&me.coords[0][0], me.scv_volume);
        orc_hex8_scv_grad_op(kernels[i].shiftedGradOp, &me.coords[0][0], &me.dndx_scv[0][0][0], NULL);
      }
  }

  /* set_zero(simdrhs), set_zero(simdlhs) */
  for (int i = 0; i < rhsSize; ++i) rhs[i] = 0.0;
  for (int i = 0; i < rhsSize * rhsSize; ++i) lhs[i] = 0.0;

  for (int i = 0; i < nk; ++i) {
    const orc_kernel_desc *k = &kernels[i];
    switch (k->kind) {
    case ORC_MOMENTUM_ADV_DIFF: momentum_adv_diff(k, elemNodes, elemIndex, &me, lhs, rhs); break;
    case ORC_CONTINUITY_ADV: continuity_adv(k, elemNodes, &me, lhs, rhs); break;
    case ORC_SCALAR_ADV_DIFF: scalar_adv_diff(k, elemNodes, elemIndex, &me, lhs, rhs); break;
    case ORC_SCALAR_DIFF: scalar_diff(k, elemNodes, &me, lhs, rhs); break;
    case ORC_SCALAR_UPW_ADV_DIFF: scalar_upw_adv_diff(k, elemNodes, elemIndex, &me, lhs, rhs); break;
    case ORC_MOMENTUM_UPW_ADV_DIFF: momentum_upw_adv_diff(k, elemNodes, elemIndex, &me, lhs, rhs); break;
    case ORC_MOMENTUM_MASS: momentum_mass(k, elemNodes, &me, lhs, rhs); break;
    case ORC_CONTINUITY_MASS: continuity_mass(k, elemNodes, &me, lhs, rhs); break;
    case ORC_SCALAR_MASS: scalar_mass(k, elemNodes, &me, lhs, rhs); break;
    case ORC_MOMENTUM_BUOYANCY_SRC: momentum_buoyancy(k, 0, elemNodes, &me, lhs, rhs); break;
    case ORC_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC: momentum_buoyancy(k, 1, elemNodes, &me, lhs, rhs); break;
    case ORC_TKE_SST_SRC: sst_src(k, 0, elemNodes, &me, lhs, rhs); break;
    case ORC_SDR_SST_SRC: sst_src(k, 1, elemNodes, &me, lhs, rhs); break;
    case ORC_TKE_SST_DES_SRC: sst_des_src(k, elemNodes, &me, lhs, rhs); break;
    case ORC_TKE_KSGS_SRC: ksgs_src(k, elemNodes, &me, lhs, rhs); break;
    case ORC_MOMENTUM_NSO: momentum_nso(k, elemNodes, &me, lhs, rhs); break;
    case ORC_TEST_FIXED_BLOCK: /* unit_tests/UnitTestTpetra.C:125-139: lhs(i,j) = elemVals[i][j] */
      for (int q = 0; q < rhsSize * rhsSize; ++q) lhs[q] = k->fixedBlock[q];
      break;
    default: break;
    }
  }
}

/* unit_tests/kernels/UnitTestKernelUtils.C:547-628 */
void orc_calc_mass_flow_rate_scs(int64_t nElem, const int32_t *elemNodes, const double *coords,
                                 const double *density, const double *velocity, double *mdot)
{
  double v_shape_function[12][8];
  init_tables();
  scs_shape(0, v_shape_function);
  for (int64_t e = 0; e < nElem; ++e) {
    const int32_t *nodes = elemNodes + 8 * e;
    double c[8][3], v_densityNp1[8], v_velocityNp1[8][3], v_scs_areav[12][3];
    gather(coords, 3, nodes, &c[0][0]);
    gather(density, 1, nodes, v_densityNp1);
    gather(velocity, 3, nodes, &v_velocityNp1[0][0]);
    orc_hex8_scs_areav(&c[0][0], &v_scs_areav[0][0]);
    for (int ip = 0; ip < 12; ++ip) {
      double rhoU[3] = {0.0, 0.0, 0.0};
      for (int ic = 0; ic < 8; ++ic) {
        const double r = v_shape_function[ip][ic];
        for (int j = 0; j < 3; ++j) rhoU[j] += r * v_densityNp1[ic] * v_velocityNp1[ic][j];
      }
      double tmdot = 0.0;
      for (int j = 0; j < 3; j++) tmdot += rhoU[j] * v_scs_areav[ip][j];
      mdot[12 * e + ip] = tmdot;
    }
  }
}

/* src/ComputeMdotElemAlgorithm.C:85-254: mass_flow_rate_scs of every locally-owned element,
 * mdot_ip = (rho u_ip - projTimeScale (dpdx_ip - Gpdx_ip)) . areav_ip; shape function shifted when
 * shiftMdot (:153-156), grad-op shifted when shiftPoisson (:190-193).  No reference unit test holds a table
 * for it; the formula is the one ContinuityAdvElemKernel.C:135-147 evaluates, so it is pinned through the
 * continuity golds (tests/test_oracle_golden.py::test_compute_mdot_elem_matches_continuity_gold). */
void orc_compute_mdot_elem(int64_t nElem, const int32_t *elemNodes, const double *coords, const double *density,
                           const double *velocity, const double *pressure, const double *Gpdx, double projTimeScale,
                           int shiftMdot, int shiftPoisson, double *mdot)
{
  double p_shape_function[12][8];
  init_tables();
  scs_shape(shiftMdot, p_shape_function);
  for (int64_t e = 0; e < nElem; ++e) {
    const int32_t *nodes = elemNodes + 8 * e;
    double p_coordinates[8][3], p_density[8], p_pressure[8], p_vrtm[8][3], p_Gpdx[8][3];
    double p_scs_areav[12][3], p_dndx[12][8][3], deriv[12][8][3];
    gather(coords, 3, nodes, &p_coordinates[0][0]);
    gather(density, 1, nodes, p_density);
    gather(pressure, 1, nodes, p_pressure);
    gather(velocity, 3, nodes, &p_vrtm[0][0]);
    gather(Gpdx, 3, nodes, &p_Gpdx[0][0]);
    orc_hex8_scs_areav(&p_coordinates[0][0], &p_scs_areav[0][0]);
    orc_hex8_scs_grad_op(shiftPoisson, &p_coordinates[0][0], &p_dndx[0][0][0], &deriv[0][0][0]);
    for (int ip = 0; ip < 12; ++ip) {
      double p_rho_uIp[3] = {0.0, 0.0, 0.0}, p_GpdxIp[3] = {0.0, 0.0, 0.0}, p_dpdxIp[3] = {0.0, 0.0, 0.0};
      for (int ic = 0; ic < 8; ++ic) {
        const double r = p_shape_function[ip][ic];
        const double nodalPressure = p_pressure[ic];
        const double nodalRho = p_density[ic];
        for (int j = 0; j < 3; ++j) {
          p_GpdxIp[j] += r * p_Gpdx[ic][j];
          p_rho_uIp[j] += r * nodalRho * p_vrtm[ic][j];
          p_dpdxIp[j] += p_dndx[ip][ic][j] * nodalPressure;
        }
      }
      double tmdot = 0.0;
      for (int j = 0; j < 3; ++j) tmdot += (p_rho_uIp[j] - projTimeScale * (p_dpdxIp[j] - p_GpdxIp[j])) * p_scs_areav[ip][j];
      mdot[12 * e + ip] = tmdot;
    }
  }
}

/* src/AssembleNodalGradElemAlgorithm.C:77-133 (ncomp = 1: dqdx) and src/AssembleNodalGradUElemAlgorithm.C:120-195
 * (ncomp = 3: dudx, row_gradQ = i*nDim): interior Green-Gauss contributions of every locally-owned element,
 * grad(L, i*3+j) += q_i,ip areav_j / dualVolume_L,  grad(R, i*3+j) -= q_i,ip areav_j / dualVolume_R.
 * `grad` is accumulated into (the reference zeroes it before the interior and boundary algorithms run). */
void orc_assemble_nodal_grad_elem(int64_t nElem, const int32_t *elemNodes, const double *coords, const double *field,
                                  int ncomp, const double *dualNodalVolume, int useShifted, double *grad)
{
  double p_shape_function[12][8];
  init_tables();
  scs_shape(useShifted, p_shape_function);
  const int *lrscv = orc_hex8_lrscv();
  for (int64_t e = 0; e < nElem; ++e) {
    const int32_t *nodes = elemNodes + 8 * e;
    double p_coordinates[8][3], p_dualVolume[8], p_q[8][3], p_scs_areav[12][3];
    gather(coords, 3, nodes, &p_coordinates[0][0]);
    gather(dualNodalVolume, 1, nodes, p_dualVolume);
    for (int n = 0; n < 8; ++n)
      for (int i = 0; i < ncomp; ++i) p_q[n][i] = field[(size_t)nodes[n] * ncomp + i];
    orc_hex8_scs_areav(&p_coordinates[0][0], &p_scs_areav[0][0]);
    for (int ip = 0; ip < 12; ++ip) {
      const int il = lrscv[2 * ip];
      const int ir = lrscv[2 * ip + 1];
      double *gradQL = grad + (size_t)nodes[il] * ncomp * 3;
      double *gradQR = grad + (size_t)nodes[ir] * ncomp * 3;
      const double inv_volL = 1.0 / p_dualVolume[il];
      const double inv_volR = 1.0 / p_dualVolume[ir];
      for (int i = 0; i < ncomp; ++i) {
        double qIp = 0.0;
        for (int ic = 0; ic < 8; ++ic) qIp += p_shape_function[ip][ic] * p_q[ic][i];
        for (int j = 0; j < 3; ++j) {
          const double fac = qIp * p_scs_areav[ip][j];
          gradQL[i * 3 + j] += fac * inv_volL;
          gradQR[i * 3 + j] -= fac * inv_volR;
        }
      }
    }
  }
}
