// general_elem.cuh -- one Hex8 element through every registered kernel, warp-cooperatively.
//
// This is the general (any mesh, any kernel list) form of what the reference does per SIMD
// lane in AssembleElemSolverAlgorithm::execute (src/AssembleElemSolverAlgorithm.C:74-100):
// gather, master-element views, zero lhs/rhs, run each Kernel::execute in registration
// order, hand lhs/rhs to sumInto.  One warp owns one element; the element's lhs/rhs and
// master-element views live in that warp's shared-memory scratch.  Work is split into
// stages separated by warp barriers; inside a stage each lane writes only locations it owns
// (an integration point, a matrix column, or a matrix row), so no atomics are needed on
// shared memory.  The stage functions take the lane as an argument so that the very same
// code is executed lane-by-lane on the CPU by tests/cpu_harness (no GPU needed to check the
// arithmetic against the oracle).
#pragma once
#include <cmath>
#include "../../include/nalu_b200.h"
#include "hex8_dev.cuh"

namespace nalu_b200 {

struct MeNeeds {
  int areav, dndx, dndx_shifted, scv_volume, dndx_scv, dndx_scv_shifted, gij;
};

// which master-element views the kernel constructors request
// (e.g. src/kernel/MomentumAdvDiffElemKernel.C:58-63, ContinuityAdvElemKernel.C:59-65)
__host__ __device__ inline void me_needs_of(const nalu_b200_kernel_desc &k, MeNeeds &m)
{
  switch (k.kind) {
  case NALU_K_MOMENTUM_ADV_DIFF:
  case NALU_K_SCALAR_ADV_DIFF:
  case NALU_K_SCALAR_DIFF:
  case NALU_K_SCALAR_UPW_ADV_DIFF:   // ScalarUpwAdvDiffElemKernel.C:82-89
  case NALU_K_MOMENTUM_UPW_ADV_DIFF: // MomentumUpwAdvDiffElemKernel.C:82-89
    m.areav = 1;
    if (k.options & NALU_OPT_SHIFTED_GRAD_OP) m.dndx_shifted = 1; else m.dndx = 1;
    break;
  case NALU_K_MOMENTUM_NSO: // MomentumNSOElemKernel.C:98-104
    m.areav = 1;
    m.gij = 1;
    if (k.options & NALU_OPT_SHIFTED_GRAD_OP) m.dndx_shifted = 1; else m.dndx = 1;
    break;
  case NALU_K_CONTINUITY_ADV: {
    const bool sp = k.options & NALU_OPT_SHIFT_POISSON, rs = k.options & NALU_OPT_REDUCED_SENS;
    m.areav = 1;
    if (!sp || !rs) m.dndx = 1;
    if (sp || rs) m.dndx_shifted = 1;
    break;
  }
  case NALU_K_MOMENTUM_MASS:
  case NALU_K_CONTINUITY_MASS:
  case NALU_K_SCALAR_MASS:
  case NALU_K_MOMENTUM_BUOYANCY_SRC:            // MomentumBuoyancySrcElemKernel.C:50-53
  case NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC: // MomentumBuoyancyBoussinesqSrcElemKernel.C:52-58
  case NALU_K_TKE_KSGS_SRC:                     // TurbKineticEnergyKsgsSrcElemKernel.C:64
    m.scv_volume = 1;
    break;
  case NALU_K_TKE_SST_SRC: // TurbKineticEnergySSTSrcElemKernel.C:84-89
  case NALU_K_SDR_SST_SRC: // SpecificDissipationRateSSTSrcElemKernel.C:91-96
  case NALU_K_TKE_SST_DES_SRC: // TurbKineticEnergySSTDESSrcElemKernel.C:91-96
    m.scv_volume = 1;
    if (k.options & NALU_OPT_SHIFTED_GRAD_OP) m.dndx_scv_shifted = 1; else m.dndx_scv = 1;
    break;
  default:
    break;
  }
}

// per-warp scratch (doubles unless noted)
template <int NDOF>
struct ElemScratch {
  static constexpr int R = 8 * NDOF;
  double coords[8][3];
  double areav[12][3];
  double dndx[12][8][3];
  double dndxs[12][8][3];
  double scv[8];
  double dndx_scv[8][8][3]; // SCV (shifted) grad op, HexSCV::grad_op / shifted_grad_op (Hex8CVFEM.C:182-190,221-229)
  double fld[8][20];  // gathered nodal fields of the kernel currently running
  double ipd[12][8];  // per-integration-point results passed between stages
  double rhs[R];
  double lhs[R * R];
  long long rowbase[8]; // values offset of dof-row 0 of each element node's row
  int rownb[8];         // blocks in that row
  int rownode[8];
  int nodes[8];
  int bad_det;
};

// metric tensors of SCS point ip (NSO kernel): 12 doubles in the storage of the SCV grad-op
template <int NDOF>
__host__ __device__ __forceinline__ double *gij_of(ElemScratch<NDOF> &ws, int ip) { return &ws.dndx_scv[0][0][0] + 12 * ip; }
using hex8::sym3;

// ---- stage: geometry --------------------------------------------------------------------
template <int NDOF>
__host__ __device__ inline void stage_geometry(int lane, ElemScratch<NDOF> &ws, const MeNeeds &need)
{
  if (lane < 12) {
    const int ip = lane;
    if (need.areav) hex8::scs_areav_ip(ws.coords, ip, ws.areav[ip]);
    double s1, s2, s3;
    if (need.dndx) {
      hex8::scs_intg_loc(ip, 0.25, s1, s2, s3);
      const double det = hex8::grad_op_ip(ws.coords, s1, s2, s3, ws.dndx[ip]);
      if (!(det > 0.0)) ws.bad_det = 1;
    }
    if (need.dndx_shifted) {
      hex8::scs_intg_loc(ip, 0.5, s1, s2, s3);
      const double det = hex8::grad_op_ip(ws.coords, s1, s2, s3, ws.dndxs[ip]);
      if (!(det > 0.0)) ws.bad_det = 1;
    }
    if (need.gij) {
      // HexSCS::gij (Hex8CVFEM.C:800-808: always the unshifted points) + generic_gij_3d (MasterElementFunctions.h:171-236):
      // g^ij = J J^T, g_ij = its inverse; both symmetric: [0..5] upper 00 01 02 11 12 22, [6..11] lower.  A list with
      // the NSO kernel has three dofs per node, the SCV grad-op (one-dof SST kernels) is never in it: its storage is free
      hex8::scs_intg_loc(ip, 0.25, s1, s2, s3);
      hex8::gij_ip(ws.coords, s1, s2, s3, gij_of(ws, ip));
    }
  } else if (lane >= 16 && lane < 24 && need.scv_volume) {
    ws.scv[lane - 16] = hex8::scv_volume_ip(ws.coords, lane - 16);
  } else if (lane >= 24 && (need.dndx_scv || need.dndx_scv_shifted)) {
    const int ip = lane - 24;
    double s1, s2, s3;
    hex8::scv_intg_loc(ip, need.dndx_scv_shifted ? 0.5 : 0.25, s1, s2, s3);
    const double det = hex8::grad_op_ip(ws.coords, s1, s2, s3, ws.dndx_scv[ip]);
    if (!(det > 0.0)) ws.bad_det = 1;
  }
}

// ---- field gather: slot list per kernel kind ---------------------------------------------
// fld[n][..] layout per kind (component offsets)
enum { F_U = 0, F_MU = 3, F_RHO = 4, F_P = 5, F_GP = 6, F_Q = 9 };             // SCS kernels
enum { F_GJU = 5, F_RTM = 14 };                                                // momentum upwind: dudx[9], velocity_rtm[3]
enum { M_UNP1 = 0, M_UN = 3, M_UNM1 = 6, M_GP = 9, M_RNP1 = 12, M_RN = 13, M_RNM1 = 14 }; // momentum mass
enum { S_QNP1 = 0, S_QN = 1, S_QNM1 = 2, S_RNP1 = 3, S_RN = 4, S_RNM1 = 5 };  // scalar / continuity mass
enum { B_S = 0 };
enum { T_U = 0, T_RHO = 3, T_K = 4, T_W = 5, T_NUT = 6, T_F1 = 7, T_LMAX = 8 }; // SST / SST-DES sources
enum { N_U = 0, N_RTM = 3, N_RHO = 6, N_MU = 7, N_P = 8, N_GJU = 9 }; // momentum NSO (the old states are read from global memory)
enum { G_K = 0, G_RHO = 1, G_NUT = 2, G_VOL = 3, G_CEPS = 4, G_MU = 5, G_DSQ = 6, G_DUDX = 7 }; // Ksgs source

__host__ __device__ inline void gather_one(double *dst, const double *field, int ncomp, int node)
{
  for (int j = 0; j < ncomp; ++j) dst[j] = field[(size_t)node * ncomp + j];
}

template <int NDOF>
__host__ __device__ inline void stage_gather(int lane, ElemScratch<NDOF> &ws, const nalu_b200_kernel_desc &k)
{
  if (lane >= 8) return;
  const int n = lane, node = ws.nodes[n];
  double *f = ws.fld[n];
  switch (k.kind) {
  case NALU_K_MOMENTUM_ADV_DIFF:
    gather_one(f + F_U, k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + F_MU, k.nodal[NALU_F_VISCOSITY], 1, node);
    break;
  case NALU_K_CONTINUITY_ADV:
    gather_one(f + F_U, k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + F_RHO, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + F_P, k.nodal[NALU_F_PRESSURE], 1, node);
    gather_one(f + F_GP, k.nodal[NALU_F_DPDX], 3, node);
    break;
  case NALU_K_SCALAR_ADV_DIFF:
  case NALU_K_SCALAR_DIFF:
    gather_one(f + F_Q, k.nodal[NALU_F_SCALARQ], 1, node);
    gather_one(f + F_MU, k.nodal[NALU_F_VISCOSITY], 1, node);
    break;
  case NALU_K_SCALAR_UPW_ADV_DIFF: // ScalarUpwAdvDiffElemKernel.C:74-81; Gjq sits in the dpdx slot
    gather_one(f + F_U, k.nodal[NALU_F_VELOCITY_RTM] ? k.nodal[NALU_F_VELOCITY_RTM] : k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + F_GP, k.nodal[NALU_F_GJQ], 3, node);
    gather_one(f + F_Q, k.nodal[NALU_F_SCALARQ], 1, node);
    gather_one(f + F_RHO, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + F_MU, k.nodal[NALU_F_VISCOSITY], 1, node);
    break;
  case NALU_K_MOMENTUM_UPW_ADV_DIFF: // MomentumUpwAdvDiffElemKernel.C:74-81
    gather_one(f + F_U, k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + F_MU, k.nodal[NALU_F_VISCOSITY], 1, node);
    gather_one(f + F_RHO, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + F_GJU, k.nodal[NALU_F_GJU], 9, node);
    gather_one(f + F_RTM, k.nodal[NALU_F_VELOCITY_RTM] ? k.nodal[NALU_F_VELOCITY_RTM] : k.nodal[NALU_F_VELOCITY], 3, node);
    break;
  case NALU_K_MOMENTUM_MASS:
    gather_one(f + M_UNP1, k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + M_UN, k.nodal[NALU_F_VELOCITY_N], 3, node);
    gather_one(f + M_UNM1, k.nodal[NALU_F_VELOCITY_NM1], 3, node);
    gather_one(f + M_GP, k.nodal[NALU_F_DPDX], 3, node);
    gather_one(f + M_RNP1, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + M_RN, k.nodal[NALU_F_DENSITY_N], 1, node);
    gather_one(f + M_RNM1, k.nodal[NALU_F_DENSITY_NM1], 1, node);
    break;
  case NALU_K_SCALAR_MASS:
    gather_one(f + S_QNP1, k.nodal[NALU_F_SCALARQ], 1, node);
    gather_one(f + S_QN, k.nodal[NALU_F_SCALARQ_N], 1, node);
    gather_one(f + S_QNM1, k.nodal[NALU_F_SCALARQ_NM1], 1, node);
    // fallthrough: densities
  case NALU_K_CONTINUITY_MASS:
    gather_one(f + S_RNP1, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + S_RN, k.nodal[NALU_F_DENSITY_N], 1, node);
    gather_one(f + S_RNM1, k.nodal[NALU_F_DENSITY_NM1], 1, node);
    break;
  case NALU_K_MOMENTUM_BUOYANCY_SRC:
    gather_one(f + B_S, k.nodal[NALU_F_DENSITY], 1, node);
    break;
  case NALU_K_TKE_SST_DES_SRC:
    gather_one(f + T_LMAX, k.nodal[NALU_F_SST_MAX_LENGTH_SCALE], 1, node);
    // fallthrough
  case NALU_K_SDR_SST_SRC:
    gather_one(f + T_F1, k.nodal[NALU_F_FONE_BLEND], 1, node);
    // fallthrough
  case NALU_K_TKE_SST_SRC:
    gather_one(f + T_U, k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + T_RHO, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + T_K, k.nodal[NALU_F_TKE], 1, node);
    gather_one(f + T_W, k.nodal[NALU_F_SDR], 1, node);
    gather_one(f + T_NUT, k.nodal[NALU_F_TVISC], 1, node);
    break;
  case NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC:
    gather_one(f + B_S, k.nodal[NALU_F_TEMPERATURE], 1, node);
    break;
  case NALU_K_MOMENTUM_NSO: // MomentumNSOElemKernel.C:82-95
    gather_one(f + N_U, k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + N_RTM, k.nodal[NALU_F_VELOCITY_RTM] ? k.nodal[NALU_F_VELOCITY_RTM] : k.nodal[NALU_F_VELOCITY], 3, node);
    gather_one(f + N_RHO, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + N_MU, k.nodal[NALU_F_VISCOSITY], 1, node);
    gather_one(f + N_P, k.nodal[NALU_F_PRESSURE], 1, node);
    gather_one(f + N_GJU, k.nodal[NALU_F_GJU], 9, node);
    break;
  case NALU_K_TKE_KSGS_SRC: // TurbKineticEnergyKsgsSrcElemKernel.C:62-72
    gather_one(f + G_K, k.nodal[NALU_F_TKE], 1, node);
    gather_one(f + G_RHO, k.nodal[NALU_F_DENSITY], 1, node);
    gather_one(f + G_NUT, k.nodal[NALU_F_TVISC], 1, node);
    gather_one(f + G_VOL, k.nodal[NALU_F_DUAL_VOLUME], 1, node);
    gather_one(f + G_CEPS, k.nodal[NALU_F_C_EPSILON], 1, node);
    gather_one(f + G_MU, k.nodal[NALU_F_VISCOSITY], 1, node);
    gather_one(f + G_DSQ, k.nodal[NALU_F_DSQRTK_DX_SQ], 1, node);
    gather_one(f + G_DUDX, k.nodal[NALU_F_GJU], 9, node);
    break;
  default:
    break;
  }
}

// ---- upwind kernels: Peclet function and limiter --------------------------------------------
// ClassicPecletFunction (A = 5) / TanhFunction, src/PecletFunction.C:67-71,98-101
__host__ __device__ inline double peclet_function(const nalu_b200_kernel_desc &k, double x)
{
  if (k.options & NALU_OPT_PECLET_TANH) return 0.50 * (1.0 + tanh((x - k.upw[NALU_UPW_TANH_TRANS]) / k.upw[NALU_UPW_TANH_WIDTH]));
  const double m = k.upw[NALU_UPW_HYBRID_FACTOR] * x;
  return m * m / (5.0 + m * m);
}
// ScalarUpwAdvDiffElemKernel.C:258-266
__host__ __device__ inline double van_leer(double dqm, double dqp)
{
  return (2.0 * (dqm * dqp + fabs(dqm * dqp))) / ((dqm + dqp) * (dqm + dqp) + 1.0e-16);
}
// what both upwind kernels derive per integration point from the edge (il, ir): the Peclet blending factor and
// the operator-sensitivity factors (ScalarUpwAdvDiffElemKernel.C:151-163,208-223)
struct UpwIp {
  double pecfac, adv_eff, aL, aR;
};
template <int NDOF>
__host__ __device__ inline void upw_edge(const ElemScratch<NDOF> &ws, const nalu_b200_kernel_desc &k, int il, int ir,
                                         const double *uL, const double *uR, double tmdot, UpwIp &U)
{
  const double diffIp = 0.5 * (ws.fld[il][F_MU] / ws.fld[il][F_RHO] + ws.fld[ir][F_MU] / ws.fld[ir][F_RHO]);
  double udotx = 0.0;
  for (int j = 0; j < 3; ++j) udotx += 0.5 * (uL[j] + uR[j]) * (ws.coords[ir][j] - ws.coords[il][j]);
  U.pecfac = peclet_function(k, fabs(udotx) / (diffIp + 1.0e-16));
  const double om_pecfac = 1.0 - U.pecfac, alpha = k.upw[NALU_UPW_ALPHA], alphaUpw = k.upw[NALU_UPW_ALPHA_UPW];
  U.adv_eff = tmdot * (U.pecfac * (1.0 - alphaUpw) + om_pecfac * (1.0 - alpha));
  U.aL = 0.5 * (tmdot + fabs(tmdot)) * U.pecfac * alphaUpw + 0.5 * alpha * om_pecfac * tmdot;
  U.aR = 0.5 * (tmdot - fabs(tmdot)) * U.pecfac * alphaUpw + 0.5 * alpha * om_pecfac * tmdot;
}
// blended advected value: pecfac * upwind + (1 - pecfac) * generalized central (:186-198)
__host__ __device__ inline double upw_blend(const nalu_b200_kernel_desc &k, double pecfac, double tmdot, double qIp,
                                            double qIpL, double qIpR)
{
  const double alpha = k.upw[NALU_UPW_ALPHA], alphaUpw = k.upw[NALU_UPW_ALPHA_UPW];
  const double qUpwind = tmdot > 0 ? alphaUpw * qIpL + (1.0 - alphaUpw) * qIp : alphaUpw * qIpR + (1.0 - alphaUpw) * qIp;
  const double qHatL = alpha * qIpL + (1.0 - alpha) * qIp, qHatR = alpha * qIpR + (1.0 - alpha) * qIp;
  return pecfac * qUpwind + (1.0 - pecfac) * (0.5 * (qHatL + qHatR));
}

// ---- stage A: per integration point quantities (lane = ip) --------------------------------
template <int NDOF>
__host__ __device__ inline void stage_ip(int lane, ElemScratch<NDOF> &ws, const nalu_b200_kernel_desc &k,
                                         long long elem)
{
  const int kind = k.kind;
  if (kind == NALU_K_MOMENTUM_MASS || kind == NALU_K_CONTINUITY_MASS || kind == NALU_K_SCALAR_MASS ||
      kind == NALU_K_TEST_FIXED_BLOCK || kind == NALU_K_MOMENTUM_BUOYANCY_SRC ||
      kind == NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC || kind == NALU_K_TKE_SST_SRC || kind == NALU_K_SDR_SST_SRC ||
      kind == NALU_K_TKE_SST_DES_SRC || kind == NALU_K_TKE_KSGS_SRC)
    return;
  if (lane >= 12) return;
  const int ip = lane;
  double s1, s2, s3;
  hex8::scs_intg_loc(ip, 0.25, s1, s2, s3);
  const double (*A) = ws.areav[ip];

  if (kind == NALU_K_MOMENTUM_NSO) {
    // MomentumNSOElemKernel.C:141-308: per point the three artificial viscosities nu_k and residuals -nu_k gijFac_k
    const double (*g)[3] = (k.options & NALU_OPT_SHIFTED_GRAD_OP) ? ws.dndxs[ip] : ws.dndx[ip];
    const double *G = gij_of(ws, ip);
    const double *uN = k.nodal[NALU_F_VELOCITY_N], *uNm1 = k.nodal[NALU_F_VELOCITY_NM1];
    const double *rN = k.nodal[NALU_F_DENSITY_N], *rNm1 = k.nodal[NALU_F_DENSITY_NM1];
    double rhoNm1Scs = 0.0, rhoNScs = 0.0, rhoNp1Scs = 0.0, dFdxCont = 0.0, divU = 0.0, rhoVrtm[3] = {0, 0, 0}, dpdx[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, s1, s2, s3);
      const double *f = ws.fld[ic];
      rhoNm1Scs += r * rNm1[ws.nodes[ic]];
      rhoNScs += r * rN[ws.nodes[ic]];
      rhoNp1Scs += r * f[N_RHO];
      for (int j = 0; j < 3; ++j) {
        rhoVrtm[j] += r * f[N_RHO] * f[N_RTM + j];
        divU += r * f[N_GJU + 4 * j];
        dFdxCont += f[N_RHO] * f[N_RTM + j] * g[ic][j];
        dpdx[j] += f[N_P] * g[ic][j];
      }
    }
    const double contRes = (k.gamma1 * rhoNp1Scs + k.gamma2 * rhoNScs + k.gamma3 * rhoNm1Scs) / k.dt + dFdxCont;
    const double nonConservedForm = 0.0, Cupw = 0.1, small = 1.0e-16; // include/nso/MomentumNSOElemKernel.h:73-80
    // g_j = sum_i g^ij A_i: fac(ic) = g . dN_ic/dx
    double gA[3];
    for (int j = 0; j < 3; ++j) gA[j] = sym3(G, 0, j) * A[0] + sym3(G, 1, j) * A[1] + sym3(G, 2, j) * A[2];
    double rGr = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rGr += rhoVrtm[i] * sym3(G + 6, i, j) * rhoVrtm[j];
    const double nuFirstOrder = sqrt(rGr);
    for (int kk = 0; kk < 3; ++kk) {
      double dFdxkAdv = 0.0, dFdxkDiff = 0.0, ukNm1Scs = 0.0, ukNScs = 0.0, ukNp1Scs = 0.0, dukdx[3] = {0, 0, 0}, gijFac = 0.0;
      for (int ic = 0; ic < 8; ++ic) {
        const double r = hex8::shape_fcn(ic, s1, s2, s3);
        const double *f = ws.fld[ic];
        const double uk = f[N_U + kk];
        ukNm1Scs += r * uNm1[(size_t)ws.nodes[ic] * 3 + kk];
        ukNScs += r * uN[(size_t)ws.nodes[ic] * 3 + kk];
        ukNp1Scs += r * uk;
        double fac = 0.0, facG = 0.0;
        for (int j = 0; j < 3; ++j) {
          const double dnj = g[ic][j];
          dukdx[j] += uk * dnj;
          dFdxkAdv += f[N_RHO] * f[N_RTM + j] * uk * dnj;
          dFdxkDiff += f[N_MU] * (f[N_GJU + 3 * kk + j] + f[N_GJU + 3 * j + kk] - (kk == j ? 2.0 / 3.0 * divU * k.include_divu : 0.0)) * dnj;
          fac += gA[j] * dnj;
          facG += gA[j] * f[N_GJU + 3 * kk + j];
        }
        gijFac += fac * uk - r * facG * k.nso_fourth_fac;
      }
      double residualAlt = dFdxkAdv - ukNp1Scs * dFdxCont;
      for (int j = 0; j < 3; ++j) residualAlt -= rhoVrtm[j] * dukdx[j];
      const double time = (k.gamma1 * rhoNp1Scs * ukNp1Scs + k.gamma2 * rhoNScs * ukNScs + k.gamma3 * rhoNm1Scs * ukNm1Scs) / k.dt;
      const double residualPde = time + dFdxkAdv - dFdxkDiff + dpdx[kk] - contRes * ukNp1Scs * nonConservedForm;
      const double residual = residualAlt * k.nso_alt_res_fac + residualPde * (1.0 - k.nso_alt_res_fac);
      double gUpperMagGradQ = 0.0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) gUpperMagGradQ += dukdx[i] * sym3(G, i, j) * dukdx[j];
      const double nuResidual = sqrt((residual * residual) / (gUpperMagGradQ + small));
      const double nu = fmin(Cupw * nuFirstOrder, nuResidual);
      ws.ipd[ip][1 + kk] = -nu * gijFac; // residualNSO: rhs[L] -=, rhs[R] +=
      ws.ipd[ip][4 + kk] = nu;
    }
  } else if (kind == NALU_K_MOMENTUM_ADV_DIFF) {
    // MomentumAdvDiffElemKernel.C:102-136 and the rhs part of :138-184
    const double (*g)[3] = (k.options & NALU_OPT_SHIFTED_GRAD_OP) ? ws.dndxs[ip] : ws.dndx[ip];
    double a1 = s1, a2 = s2, a3 = s3; // advection shape function location
    if (k.options & NALU_OPT_SKEW_SYMMETRIC) hex8::scs_intg_loc(ip, 0.5, a1, a2, a3);
    const double tmdot = k.mass_flow_rate[elem * 12 + ip];
    double mu = 0.0, divU = 0.0, u[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, s1, s2, s3), ra = hex8::shape_fcn(ic, a1, a2, a3);
      mu += r * ws.fld[ic][F_MU];
      for (int j = 0; j < 3; ++j) {
        u[j] += ra * ws.fld[ic][F_U + j];
        divU += ws.fld[ic][F_U + j] * g[ic][j];
      }
    }
    double flux[3];
    for (int i = 0; i < 3; ++i) flux[i] = tmdot * u[i] + 2.0 / 3.0 * mu * divU * A[i] * k.include_divu;
    for (int ic = 0; ic < 8; ++ic) {
      const double *uc = &ws.fld[ic][F_U];
      const double gA = g[ic][0] * A[0] + g[ic][1] * A[1] + g[ic][2] * A[2];
      const double uA = uc[0] * A[0] + uc[1] * A[1] + uc[2] * A[2];
      for (int i = 0; i < 3; ++i) flux[i] += -mu * (g[ic][i] * uA + gA * uc[i]);
    }
    ws.ipd[ip][0] = mu;
    ws.ipd[ip][1] = flux[0];
    ws.ipd[ip][2] = flux[1];
    ws.ipd[ip][3] = flux[2];
    ws.ipd[ip][4] = tmdot;
  } else if (kind == NALU_K_MOMENTUM_UPW_ADV_DIFF) {
    // MomentumUpwAdvDiffElemKernel.C:140-311: the rhs flux of the point; the lhs factors for stage B
    const double (*g)[3] = (k.options & NALU_OPT_SHIFTED_GRAD_OP) ? ws.dndxs[ip] : ws.dndx[ip];
    double a1 = s1, a2 = s2, a3 = s3;
    if (k.options & NALU_OPT_SKEW_SYMMETRIC) hex8::scs_intg_loc(ip, 0.5, a1, a2, a3);
    const int il = hex8::scs_left(ip), ir = hex8::scs_right(ip);
    const double tmdot = k.mass_flow_rate[elem * 12 + ip];
    double mu = 0.0, divU = 0.0, u[3] = {0, 0, 0}, xip[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, s1, s2, s3), ra = hex8::shape_fcn(ic, a1, a2, a3);
      mu += r * ws.fld[ic][F_MU];
      for (int j = 0; j < 3; ++j) {
        u[j] += ra * ws.fld[ic][F_U + j];
        xip[j] += ra * ws.coords[ic][j];
        divU += ws.fld[ic][F_U + j] * g[ic][j];
      }
    }
    UpwIp U;
    upw_edge(ws, k, il, ir, &ws.fld[il][F_RTM], &ws.fld[ir][F_RTM], tmdot, U);
    double flux[3];
    for (int i = 0; i < 3; ++i) {
      double duL = 0.0, duR = 0.0;
      for (int j = 0; j < 3; ++j) {
        duL += (xip[j] - ws.coords[il][j]) * ws.fld[il][F_GJU + 3 * i + j];
        duR += (ws.coords[ir][j] - xip[j]) * ws.fld[ir][F_GJU + 3 * i + j];
      }
      double limL = 1.0, limR = 1.0;
      if (k.options & NALU_OPT_USE_LIMITER) {
        const double dq = ws.fld[ir][F_U + i] - ws.fld[il][F_U + i];
        limL = van_leer(2.0 * 2.0 * duL - dq, dq);
        limR = van_leer(2.0 * 2.0 * duR - dq, dq);
      }
      const double uL = ws.fld[il][F_U + i] + duL * k.upw[NALU_UPW_HO_UPWIND] * limL;
      const double uR = ws.fld[ir][F_U + i] - duR * k.upw[NALU_UPW_HO_UPWIND] * limR;
      flux[i] = tmdot * upw_blend(k, U.pecfac, tmdot, u[i], uL, uR) + 2.0 / 3.0 * mu * divU * A[i] * k.include_divu;
    }
    for (int ic = 0; ic < 8; ++ic) {
      const double *uc = &ws.fld[ic][F_U];
      const double gA = g[ic][0] * A[0] + g[ic][1] * A[1] + g[ic][2] * A[2];
      const double uA = uc[0] * A[0] + uc[1] * A[1] + uc[2] * A[2];
      for (int i = 0; i < 3; ++i) flux[i] += -mu * (g[ic][i] * uA + gA * uc[i]);
    }
    ws.ipd[ip][0] = mu;
    ws.ipd[ip][1] = flux[0];
    ws.ipd[ip][2] = flux[1];
    ws.ipd[ip][3] = flux[2];
    ws.ipd[ip][4] = U.adv_eff;
    ws.ipd[ip][5] = U.aL;
    ws.ipd[ip][6] = U.aR;
  } else if (kind == NALU_K_SCALAR_UPW_ADV_DIFF) {
    // ScalarUpwAdvDiffElemKernel.C:132-252
    const double (*g)[3] = (k.options & NALU_OPT_SHIFTED_GRAD_OP) ? ws.dndxs[ip] : ws.dndx[ip];
    double a1 = s1, a2 = s2, a3 = s3;
    if (k.options & NALU_OPT_SKEW_SYMMETRIC) hex8::scs_intg_loc(ip, 0.5, a1, a2, a3);
    const int il = hex8::scs_left(ip), ir = hex8::scs_right(ip);
    const double tmdot = k.mass_flow_rate[elem * 12 + ip];
    double qIp = 0.0, gam = 0.0, xip[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, s1, s2, s3), ra = hex8::shape_fcn(ic, a1, a2, a3);
      qIp += ra * ws.fld[ic][F_Q];
      gam += r * ws.fld[ic][F_MU];
      for (int j = 0; j < 3; ++j) xip[j] += ra * ws.coords[ic][j];
    }
    UpwIp U;
    upw_edge(ws, k, il, ir, &ws.fld[il][F_U], &ws.fld[ir][F_U], tmdot, U);
    double dqL = 0.0, dqR = 0.0;
    for (int j = 0; j < 3; ++j) {
      dqL += (xip[j] - ws.coords[il][j]) * ws.fld[il][F_GP + j];
      dqR += (ws.coords[ir][j] - xip[j]) * ws.fld[ir][F_GP + j];
    }
    double limL = 1.0, limR = 1.0;
    if (k.options & NALU_OPT_USE_LIMITER) {
      const double dq = ws.fld[ir][F_Q] - ws.fld[il][F_Q];
      limL = van_leer(2.0 * 2.0 * dqL - dq, dq);
      limR = van_leer(2.0 * 2.0 * dqR - dq, dq);
    }
    const double qL = ws.fld[il][F_Q] + dqL * k.upw[NALU_UPW_HO_UPWIND] * limL;
    const double qR = ws.fld[ir][F_Q] - dqR * k.upw[NALU_UPW_HO_UPWIND] * limR;
    double q = tmdot * upw_blend(k, U.pecfac, tmdot, qIp, qL, qR);
    for (int ic = 0; ic < 8; ++ic)
      q += -gam * (g[ic][0] * A[0] + g[ic][1] * A[1] + g[ic][2] * A[2]) * ws.fld[ic][F_Q];
    ws.ipd[ip][0] = gam;
    ws.ipd[ip][1] = q;
    ws.ipd[ip][4] = U.adv_eff;
    ws.ipd[ip][5] = U.aL;
    ws.ipd[ip][6] = U.aR;
  } else if (kind == NALU_K_CONTINUITY_ADV) {
    // ContinuityAdvElemKernel.C:110-147
    const bool sp = k.options & NALU_OPT_SHIFT_POISSON;
    const double (*g)[3] = sp ? ws.dndxs[ip] : ws.dndx[ip];
    double m1 = s1, m2 = s2, m3 = s3;
    if (k.options & NALU_OPT_SHIFT_MDOT) hex8::scs_intg_loc(ip, 0.5, m1, m2, m3);
    const double tau = k.dt / k.gamma1;
    double rhoU[3] = {0, 0, 0}, Gp[3] = {0, 0, 0}, dpdx[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, m1, m2, m3);
      const double p = ws.fld[ic][F_P], rho = ws.fld[ic][F_RHO];
      for (int j = 0; j < 3; ++j) {
        Gp[j] += r * ws.fld[ic][F_GP + j];
        rhoU[j] += r * rho * ws.fld[ic][F_U + j];
        dpdx[j] += g[ic][j] * p;
      }
    }
    double mdot = 0.0;
    for (int j = 0; j < 3; ++j) mdot += (rhoU[j] - tau * (dpdx[j] - Gp[j])) * A[j];
    ws.ipd[ip][0] = mdot / tau;
  } else { // NALU_K_SCALAR_ADV_DIFF, NALU_K_SCALAR_DIFF
    // ScalarAdvDiffElemKernel.C:83-125, ScalarDiffElemKernel.C:80-108
    const double (*g)[3] = (k.options & NALU_OPT_SHIFTED_GRAD_OP) ? ws.dndxs[ip] : ws.dndx[ip];
    const bool adv = kind == NALU_K_SCALAR_ADV_DIFF;
    double a1 = s1, a2 = s2, a3 = s3;
    if (k.options & NALU_OPT_SKEW_SYMMETRIC) hex8::scs_intg_loc(ip, 0.5, a1, a2, a3);
    const double tmdot = adv ? k.mass_flow_rate[elem * 12 + ip] : 0.0;
    double gam = 0.0;
    for (int ic = 0; ic < 8; ++ic) gam += hex8::shape_fcn(ic, s1, s2, s3) * ws.fld[ic][F_MU];
    double q = 0.0;
    for (int ic = 0; ic < 8; ++ic) {
      const double fadv = hex8::shape_fcn(ic, a1, a2, a3) * tmdot;
      const double fdiff = -gam * (g[ic][0] * A[0] + g[ic][1] * A[1] + g[ic][2] * A[2]);
      q += (fadv + fdiff) * ws.fld[ic][F_Q];
    }
    ws.ipd[ip][0] = gam;
    ws.ipd[ip][1] = q;
    ws.ipd[ip][4] = tmdot;
  }
}

// ---- stage B: lhs, one lane per matrix column (SCS kernels) or per SCV node (mass) -----------
template <int NDOF>
__host__ __device__ inline void stage_lhs(int lane, ElemScratch<NDOF> &ws, const nalu_b200_kernel_desc &k)
{
  constexpr int R = ElemScratch<NDOF>::R;
  const int kind = k.kind;
  if (kind == NALU_K_MOMENTUM_ADV_DIFF || kind == NALU_K_MOMENTUM_UPW_ADV_DIFF) {
    if (NDOF != 3 || lane >= 24) return;
    const int ic = lane / 3, j = lane % 3;
    const bool upw = kind == NALU_K_MOMENTUM_UPW_ADV_DIFF;
    const bool sh = k.options & NALU_OPT_SHIFTED_GRAD_OP;
    for (int ip = 0; ip < 12; ++ip) {
      const int L = hex8::scs_left(ip), Rn = hex8::scs_right(ip);
      double a1, a2, a3;
      hex8::scs_intg_loc(ip, (k.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25, a1, a2, a3);
      const double *g = sh ? ws.dndxs[ip][ic] : ws.dndx[ip][ic];
      const double *A = ws.areav[ip];
      const double mu = ws.ipd[ip][0];
      // ipd[4]: mdot, or mdot (pecfac (1 - alphaUpw) + (1 - pecfac)(1 - alpha)) for the upwind kernel, whose left /
      // right node columns also take alhsfacL / alhsfacR on the diagonal of the block (:246-255)
      double lhsfacAdv = hex8::shape_fcn(ic, a1, a2, a3) * ws.ipd[ip][4];
      if (upw && ic == L) lhsfacAdv += ws.ipd[ip][5];
      if (upw && ic == Rn) lhsfacAdv += ws.ipd[ip][6];
      const double diag = lhsfacAdv - mu * (g[0] * A[0] + g[1] * A[1] + g[2] * A[2]);
      for (int i = 0; i < 3; ++i) {
        double v = -mu * g[i] * A[j];
        if (i == j) v += diag;
        ws.lhs[(L * 3 + i) * R + lane] += v;
        ws.lhs[(Rn * 3 + i) * R + lane] -= v;
      }
    }
  } else if (kind == NALU_K_MOMENTUM_NSO) {
    // MomentumNSOElemKernel.C:282-303: lhs(L k, ic k) += nu_k lhsfac, lhs(R k, ic k) -= nu_k lhsfac,
    // lhsfac = -sum_ij g^ij dN_ic/dx_j A_i: no coupling between components
    if (NDOF != 3 || lane >= 24) return;
    const int ic = lane / 3, kk = lane % 3;
    const bool sh = k.options & NALU_OPT_SHIFTED_GRAD_OP;
    for (int ip = 0; ip < 12; ++ip) {
      const int L = hex8::scs_left(ip), Rn = hex8::scs_right(ip);
      const double *g = sh ? ws.dndxs[ip][ic] : ws.dndx[ip][ic];
      const double *A = ws.areav[ip];
      const double *G = gij_of(ws, ip);
      double lhsfac = 0.0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) lhsfac -= sym3(G, i, j) * g[j] * A[i];
      const double v = ws.ipd[ip][4 + kk] * lhsfac;
      ws.lhs[(L * 3 + kk) * R + lane] += v;
      ws.lhs[(Rn * 3 + kk) * R + lane] -= v;
    }
  } else if (kind == NALU_K_CONTINUITY_ADV || kind == NALU_K_SCALAR_ADV_DIFF || kind == NALU_K_SCALAR_DIFF ||
             kind == NALU_K_SCALAR_UPW_ADV_DIFF) {
    if (NDOF != 1 || lane >= 8) return;
    const int ic = lane;
    for (int ip = 0; ip < 12; ++ip) {
      const int L = hex8::scs_left(ip), Rn = hex8::scs_right(ip);
      const double *A = ws.areav[ip];
      double v;
      if (kind == NALU_K_CONTINUITY_ADV) {
        const bool shl = (k.options & NALU_OPT_SHIFT_POISSON) || (k.options & NALU_OPT_REDUCED_SENS);
        const double *g = shl ? ws.dndxs[ip][ic] : ws.dndx[ip][ic];
        v = -(g[0] * A[0] + g[1] * A[1] + g[2] * A[2]);
      } else {
        const double *g = (k.options & NALU_OPT_SHIFTED_GRAD_OP) ? ws.dndxs[ip][ic] : ws.dndx[ip][ic];
        double a1, a2, a3;
        hex8::scs_intg_loc(ip, (k.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25, a1, a2, a3);
        v = hex8::shape_fcn(ic, a1, a2, a3) * ws.ipd[ip][4] -
            ws.ipd[ip][0] * (g[0] * A[0] + g[1] * A[1] + g[2] * A[2]);
        if (kind == NALU_K_SCALAR_UPW_ADV_DIFF) { // ScalarUpwAdvDiffElemKernel.C:208-223
          if (ic == L) v += ws.ipd[ip][5];
          if (ic == Rn) v += ws.ipd[ip][6];
        }
      }
      ws.lhs[L * R + ic] += v;
      ws.lhs[Rn * R + ic] -= v;
    }
  } else if (kind == NALU_K_MOMENTUM_MASS) {
    // MomentumMassElemKernel.C:119-170; lane = SCV ip = nearest node
    if (NDOF != 3 || lane >= 8) return;
    const int ip = lane;
    double s1, s2, s3;
    hex8::scv_intg_loc(ip, (k.options & NALU_OPT_LUMPED_MASS) ? 0.5 : 0.25, s1, s2, s3);
    double rNm1 = 0, rN = 0, rNp1 = 0, uNm1[3] = {0, 0, 0}, uN[3] = {0, 0, 0}, uNp1[3] = {0, 0, 0}, Gp[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, s1, s2, s3);
      const double *f = ws.fld[ic];
      rNm1 += r * f[M_RNM1];
      rN += r * f[M_RN];
      rNp1 += r * f[M_RNP1];
      for (int j = 0; j < 3; ++j) {
        uNm1[j] += r * f[M_UNM1 + j];
        uN[j] += r * f[M_UN + j];
        uNp1[j] += r * f[M_UNP1 + j];
        Gp[j] += r * f[M_GP + j];
      }
    }
    const double scV = ws.scv[ip];
    const double densFac = rNp1 * k.dens_fac + k.om_dens_fac;
    for (int j = 0; j < 3; ++j)
      ws.rhs[ip * 3 + j] += -(k.gamma1 * rNp1 * uNp1[j] + k.gamma2 * rN * uN[j] + k.gamma3 * rNm1 * uNm1[j]) * scV / k.dt -
                            Gp[j] * scV * densFac;
    for (int ic = 0; ic < 8; ++ic) {
      const double lhsfac = hex8::shape_fcn(ic, s1, s2, s3) * k.gamma1 * rNp1 * scV / k.dt;
      for (int j = 0; j < 3; ++j) ws.lhs[(ip * 3 + j) * R + ic * 3 + j] += lhsfac;
    }
  } else if (kind == NALU_K_MOMENTUM_BUOYANCY_SRC || kind == NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC) {
    // MomentumBuoyancySrcElemKernel.C:69-86, MomentumBuoyancyBoussinesqSrcElemKernel.C:74-91; lane = SCV ip =
    // nearest node; the (never shifted) SCV shape function; no lhs
    if (NDOF != 3 || lane >= 8) return;
    const int ip = lane;
    double s1, s2, s3;
    hex8::scv_intg_loc(ip, 0.25, s1, s2, s3);
    double sIp = 0.0;
    for (int ic = 0; ic < 8; ++ic) sIp += hex8::shape_fcn(ic, s1, s2, s3) * ws.fld[ic][B_S];
    const double scV = ws.scv[ip];
    const double fac = kind == NALU_K_MOMENTUM_BUOYANCY_SRC ? (sIp - k.rho_ref) * scV
                                                             : -k.rho_ref * k.beta * (sIp - k.t_ref) * scV;
    for (int j = 0; j < 3; ++j) ws.rhs[ip * 3 + j] += fac * k.gravity[j];
  } else if (kind == NALU_K_TKE_SST_SRC || kind == NALU_K_SDR_SST_SRC || kind == NALU_K_TKE_SST_DES_SRC) {
    // TurbKineticEnergySSTSrcElemKernel.C:120-172, SpecificDissipationRateSSTSrcElemKernel.C:128-202,
    // TurbKineticEnergySSTDESSrcElemKernel.C:131-196; lane = SCV ip = nearest node
    if (NDOF != 1 || lane >= 8) return;
    const int ip = lane;
    double s1, s2, s3;
    hex8::scv_intg_loc(ip, (k.options & NALU_OPT_LUMPED_MASS) ? 0.5 : 0.25, s1, s2, s3);
    double rho = 0, tke = 0, sdr = 0, tvisc = 0, f1 = 0, lmax = 0, dudx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}},
           dkdx[3] = {0, 0, 0}, dwdx[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, s1, s2, s3);
      const double *f = ws.fld[ic];
      rho += r * f[T_RHO];
      tke += r * f[T_K];
      sdr += r * f[T_W];
      tvisc += r * f[T_NUT];
      f1 += r * f[T_F1];
      lmax += r * f[T_LMAX];
      const double *g = ws.dndx_scv[ip][ic];
      for (int i = 0; i < 3; ++i) {
        dkdx[i] += g[i] * f[T_K];
        dwdx[i] += g[i] * f[T_W];
        for (int j = 0; j < 3; ++j) dudx[i][j] += g[j] * f[T_U + i];
      }
    }
    double Pk = 0.0, crossDiff = 0.0;
    for (int i = 0; i < 3; ++i) {
      crossDiff += dkdx[i] * dwdx[i];
      for (int j = 0; j < 3; ++j) Pk += dudx[i][j] * (dudx[i][j] + dudx[j][i]);
    }
    Pk *= tvisc;
    const double betaStar = k.turb[NALU_TM_BETA_STAR], limit = k.turb[NALU_TM_TKE_PROD_LIMIT_RATIO];
    const double scV = ws.scv[ip];
    double Dk = betaStar * rho * sdr * tke;
    double lhsfac;
    if (kind == NALU_K_TKE_SST_DES_SRC) {
      // the DES length scale replaces k^(1/2) / (betaStar omega) where the grid resolves it (:165-178)
      const double cDES = f1 * k.turb[NALU_TM_CDES_KW] + (1.0 - f1) * k.turb[NALU_TM_CDES_KE];
      const double sqrtTke = sqrt(tke);
      const double lSST = sqrtTke / betaStar / sdr;
      const double lDES = fmax(1.0e-16, fmin(lSST, cDES * lmax));
      const double tkeFac = rho * sqrtTke / lDES;
      Dk = tkeFac * tke;
      Pk = fmin(Pk, limit * Dk);
      ws.rhs[ip] += (Pk - Dk) * scV;
      lhsfac = tkeFac * scV;
      for (int ic = 0; ic < 8; ++ic) ws.lhs[ip * R + ic] += hex8::shape_fcn(ic, s1, s2, s3) * lhsfac;
      return;
    }
    Pk = fmin(Pk, limit * Dk);
    if (kind == NALU_K_TKE_SST_SRC) {
      ws.rhs[ip] += (Pk - Dk) * scV;
      lhsfac = betaStar * rho * sdr * scV;
    } else {
      const double om = 1.0 - f1;
      const double beta = f1 * k.turb[NALU_TM_BETA_ONE] + om * k.turb[NALU_TM_BETA_TWO];
      const double gamma = f1 * k.turb[NALU_TM_GAMMA_ONE] + om * k.turb[NALU_TM_GAMMA_TWO];
      const double sigmaD = 2.0 * om * k.turb[NALU_TM_SIGMA_W_TWO];
      const double Pw = gamma * rho * Pk / fmax(tvisc, 1.0e-16);
      const double Dw = beta * rho * sdr * sdr;
      const double Sw = sigmaD * rho * crossDiff / sdr;
      ws.rhs[ip] += (Pw - Dw + Sw) * scV;
      lhsfac = (2.0 * beta * rho * sdr + fmax(Sw / sdr, 0.0)) * scV;
    }
    for (int ic = 0; ic < 8; ++ic) ws.lhs[ip * R + ic] += hex8::shape_fcn(ic, s1, s2, s3) * lhsfac;
  } else if (kind == NALU_K_TKE_KSGS_SRC) {
    // TurbKineticEnergyKsgsSrcElemKernel.C:90-126: everything evaluated at the nearest node, lumped diagonal
    if (NDOF != 1 || lane >= 8) return;
    const int ip = lane;
    const double *f = ws.fld[ip];
    const double *G = f + G_DUDX;
    double Pk = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Pk += G[3 * i + j] * (G[3 * i + j] + G[3 * j + i]);
    Pk *= f[G_NUT];
    const double tke = f[G_K];
    const double tkeFac = f[G_CEPS] * f[G_RHO] * sqrt(tke) / cbrt(f[G_VOL]);
    const double Dk = tkeFac * tke + k.turb[NALU_TM_LRKSGS_FAC] * 2.0 * f[G_MU] * f[G_DSQ];
    Pk = fmin(Pk, k.turb[NALU_TM_TKE_PROD_LIMIT_RATIO] * Dk);
    const double scV = ws.scv[ip];
    ws.rhs[ip] += (Pk - Dk) * scV;
    ws.lhs[ip * R + ip] += 1.5 * tkeFac * scV;
  } else if (kind == NALU_K_CONTINUITY_MASS || kind == NALU_K_SCALAR_MASS) {
    // ContinuityMassElemKernel.C:98-117, ScalarMassElemKernel.C:117-154
    if (NDOF != 1 || lane >= 8) return;
    const int ip = lane;
    double s1, s2, s3;
    hex8::scv_intg_loc(ip, (k.options & NALU_OPT_LUMPED_MASS) ? 0.5 : 0.25, s1, s2, s3);
    double q[3] = {0, 0, 0}, rho[3] = {0, 0, 0};
    for (int ic = 0; ic < 8; ++ic) {
      const double r = hex8::shape_fcn(ic, s1, s2, s3);
      const double *f = ws.fld[ic];
      for (int t = 0; t < 3; ++t) {
        rho[t] += r * f[S_RNP1 + t];
        if (kind == NALU_K_SCALAR_MASS) q[t] += r * f[S_QNP1 + t];
      }
    }
    const double scV = ws.scv[ip];
    if (kind == NALU_K_CONTINUITY_MASS) {
      const double tau = k.dt / k.gamma1;
      ws.rhs[ip] += -(k.gamma1 * rho[0] + k.gamma2 * rho[1] + k.gamma3 * rho[2]) * scV / k.dt / tau;
    } else {
      ws.rhs[ip] += -(k.gamma1 * rho[0] * q[0] + k.gamma2 * rho[1] * q[1] + k.gamma3 * rho[2] * q[2]) * scV / k.dt;
      for (int ic = 0; ic < 8; ++ic)
        ws.lhs[ip * R + ic] += hex8::shape_fcn(ic, s1, s2, s3) * k.gamma1 * rho[0] * scV / k.dt;
    }
  } else if (kind == NALU_K_TEST_FIXED_BLOCK) {
    // unit_tests/UnitTestTpetra.C:125-139: lhs(i,j) = elemVals[i][j] (assignment, not +=)
    for (int q = lane; q < R * R; q += 32) ws.lhs[q] = k.fixed_block[q];
  }
}

// ---- stage C: rhs of the SCS kernels, one lane per matrix row -----------------------------
template <int NDOF>
__host__ __device__ inline void stage_rhs(int lane, ElemScratch<NDOF> &ws, const nalu_b200_kernel_desc &k)
{
  constexpr int R = ElemScratch<NDOF>::R;
  const int kind = k.kind;
  if (lane >= R) return;
  int col;
  if (kind == NALU_K_MOMENTUM_ADV_DIFF || kind == NALU_K_MOMENTUM_UPW_ADV_DIFF || kind == NALU_K_MOMENTUM_NSO) col = 1 + lane % 3;
  else if (kind == NALU_K_CONTINUITY_ADV) col = 0;
  else if (kind == NALU_K_SCALAR_ADV_DIFF || kind == NALU_K_SCALAR_DIFF || kind == NALU_K_SCALAR_UPW_ADV_DIFF) col = 1;
  else return;
  const int node = lane / NDOF;
  double acc = 0.0;
  for (int ip = 0; ip < 12; ++ip) {
    if (hex8::scs_left(ip) == node) acc -= ws.ipd[ip][col];
    else if (hex8::scs_right(ip) == node) acc += ws.ipd[ip][col];
  }
  ws.rhs[lane] += acc;
}

} // namespace nalu_b200
