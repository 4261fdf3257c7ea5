// CPU emulation of the general element kernel's stage functions (nalu_b200/csrc/general_elem.cuh
// compiled as plain C++; each stage is run lane by lane, a warp barrier between stages) against
// the oracle, for every kernel kind and option combination.  Test scaffolding.
#define __host__
#define __device__
#define __forceinline__ inline
#include "../../nalu_b200/csrc/general_elem.cuh"
#include "../../oracle/nalu_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace nalu_b200;

static double frand() { return rand() / (double)RAND_MAX - 0.5; }

template <int NDOF>
static double run_case(const std::vector<nalu_b200_kernel_desc> &ks, const std::vector<orc_kernel_desc> &os,
                       const int *nodes, const double *coords)
{
  static ElemScratch<NDOF> ws;
  constexpr int R = 8 * NDOF;
  memset(&ws, 0, sizeof(ws));
  MeNeeds need = {0, 0, 0, 0};
  for (auto &k : ks) me_needs_of(k, need);
  for (int n = 0; n < 8; ++n) {
    ws.nodes[n] = nodes[n];
    for (int d = 0; d < 3; ++d) ws.coords[n][d] = coords[nodes[n] * 3 + d];
  }
  for (int lane = 0; lane < 32; ++lane) stage_geometry(lane, ws, need);
  for (auto &k : ks) {
    for (int lane = 0; lane < 32; ++lane) stage_gather(lane, ws, k);
    for (int lane = 0; lane < 32; ++lane) stage_ip(lane, ws, k, 0);
    for (int lane = 0; lane < 32; ++lane) stage_lhs(lane, ws, k);
    for (int lane = 0; lane < 32; ++lane) stage_rhs(lane, ws, k);
  }
  double lhs[R * R], rhs[R];
  orc_element_lhs_rhs(os.data(), (int)os.size(), NDOF, nodes, 0, lhs, rhs);
  double scale = 0, err = 0;
  for (int i = 0; i < R * R; ++i) scale = fmax(scale, fabs(lhs[i]));
  for (int i = 0; i < R; ++i) scale = fmax(scale, fabs(rhs[i]));
  for (int i = 0; i < R * R; ++i) err = fmax(err, fabs(lhs[i] - ws.lhs[i]));
  for (int i = 0; i < R; ++i) err = fmax(err, fabs(rhs[i] - ws.rhs[i]));
  if (ws.bad_det) return 1.0;
  return err / scale;
}

int main()
{
  srand(4242);
  const int NN = 12;
  double worst = 0;
  for (int trial = 0; trial < 40; ++trial) {
    // 12 nodes; the element uses a random-ish subset ordering of a perturbed cube + 4 decoys
    double coords[NN * 3], vel[NN * 3], velN[NN * 3], velNm1[NN * 3], visc[NN], rho[NN], rhoN[NN], rhoNm1[NN],
        p[NN], gp[NN * 3], q[NN], qN[NN], qNm1[NN], mdot[12];
    int nodes[8];
    for (int n = 0; n < 8; ++n) nodes[n] = (n * 5 + trial) % NN;  // distinct (5 coprime with 12)
    for (int n = 0; n < NN; ++n) {
      for (int d = 0; d < 3; ++d) {
        coords[n * 3 + d] = 10 * frand();
        vel[n * 3 + d] = frand(); velN[n * 3 + d] = frand(); velNm1[n * 3 + d] = frand(); gp[n * 3 + d] = frand();
      }
      visc[n] = 0.1 + 0.05 * frand(); rho[n] = 1 + 0.3 * frand(); rhoN[n] = 1 + 0.3 * frand(); rhoNm1[n] = 1 + 0.3 * frand();
      p[n] = frand(); q[n] = frand(); qN[n] = frand(); qNm1[n] = frand();
    }
    for (int n = 0; n < 8; ++n) {
      const double amp = 0.15;
      coords[nodes[n] * 3 + 0] = 2.0 + 1.3 * hex8::sgn_x(n) * 0.5 + amp * frand();
      coords[nodes[n] * 3 + 1] = -1.0 + 0.9 * hex8::sgn_y(n) * 0.5 + amp * frand();
      coords[nodes[n] * 3 + 2] = 0.5 + 1.1 * hex8::sgn_z(n) * 0.5 + amp * frand();
    }
    for (int i = 0; i < 12; ++i) mdot[i] = frand();

    // upwind adv-diff kernels: shifted grad-op / skew-symmetric / limiter x classic or tanh Peclet function
    for (int kind = NALU_K_SCALAR_UPW_ADV_DIFF; kind <= NALU_K_MOMENTUM_UPW_ADV_DIFF; ++kind) {
      for (int opt = 0; opt < 8; ++opt) {
        static double gjq[NN * 3], gju[NN * 9], rtm[NN * 3];
        for (int n = 0; n < NN * 3; ++n) { gjq[n] = frand(); rtm[n] = frand(); }
        for (int n = 0; n < NN * 9; ++n) gju[n] = frand();
        nalu_b200_kernel_desc k;
        orc_kernel_desc o;
        memset(&k, 0, sizeof(k));
        memset(&o, 0, sizeof(o));
        k.kind = kind; o.kind = kind;
        k.include_divu = o.includeDivU = (opt & 1) ? 1.0 : 0.0;
        int options = 0;
        if (opt & 1) { options |= NALU_OPT_SHIFTED_GRAD_OP; o.shiftedGradOp = 1; }
        if (opt & 2) { options |= NALU_OPT_SKEW_SYMMETRIC; o.skewSymmetric = 1; }
        if (opt & 4) { options |= NALU_OPT_USE_LIMITER; o.useLimiter = 1; }
        if (trial & 1) { options |= NALU_OPT_PECLET_TANH; o.pecletTanh = 1; }
        k.options = options;
        k.upw[NALU_UPW_ALPHA] = o.alpha = 0.1 * (trial % 3);
        k.upw[NALU_UPW_ALPHA_UPW] = o.alphaUpw = 1.0 - 0.25 * (trial % 2);
        k.upw[NALU_UPW_HO_UPWIND] = o.hoUpwind = (opt & 2) ? 1.0 : 0.5;
        k.upw[NALU_UPW_HYBRID_FACTOR] = o.hybridFactor = 0.7;
        k.upw[NALU_UPW_TANH_TRANS] = o.tanhTrans = 2.0;
        k.upw[NALU_UPW_TANH_WIDTH] = o.tanhWidth = 4.0;
        k.nodal[NALU_F_COORDINATES] = o.coordinates = coords;
        k.nodal[NALU_F_VELOCITY] = o.velocity = vel;
        k.nodal[NALU_F_VISCOSITY] = o.viscosity = visc;
        k.nodal[NALU_F_DENSITY] = o.density = rho;
        k.nodal[NALU_F_SCALARQ] = o.scalarQ = q;
        k.nodal[NALU_F_GJQ] = o.Gjq = gjq;
        k.nodal[NALU_F_GJU] = o.Gju = gju;
        if (opt & 4) k.nodal[NALU_F_VELOCITY_RTM] = o.velocityRTM = rtm; // mesh motion: separate velocity_rtm
        k.mass_flow_rate = o.massFlowRate = mdot;
        std::vector<nalu_b200_kernel_desc> ks{k};
        std::vector<orc_kernel_desc> os{o};
        const double e = kind == NALU_K_MOMENTUM_UPW_ADV_DIFF ? run_case<3>(ks, os, nodes, coords) : run_case<1>(ks, os, nodes, coords);
        if (e > 1e-13) { printf("FAIL upw kind %d opt %d trial %d relerr %.3e\n", kind, opt, trial, e); return 1; }
        worst = fmax(worst, e);
      }
    }

    // turbulence source kernels: SST tke / sdr, SST-DES tke, Ksgs tke; lumped x shifted grad-op
    for (int kind : {NALU_K_TKE_SST_SRC, NALU_K_SDR_SST_SRC, NALU_K_TKE_SST_DES_SRC, NALU_K_TKE_KSGS_SRC}) {
      for (int opt = 0; opt < 4; ++opt) {
        static double tke[NN], sdr[NN], tvisc[NN], f1[NN], lmax[NN], ceps[NN], dsq[NN], dvol[NN], gju[NN * 9];
        for (int n = 0; n < NN; ++n) {
          tke[n] = 2.0 + frand(); sdr[n] = 2.0 + frand(); tvisc[n] = 0.3 + 0.2 * frand(); f1[n] = 0.5 + frand();
          lmax[n] = 8.0 + 14.0 * frand(); ceps[n] = 0.845 + 0.1 * frand(); dsq[n] = 0.5 + frand(); dvol[n] = 1.0 + 0.5 * frand();
          for (int j = 0; j < 9; ++j) gju[n * 9 + j] = (n & 1 ? 6.0 : 1.0) * frand();
        }
        nalu_b200_kernel_desc k;
        orc_kernel_desc o;
        memset(&k, 0, sizeof(k));
        memset(&o, 0, sizeof(o));
        k.kind = kind; o.kind = kind;
        int options = 0;
        if (opt & 1) { options |= NALU_OPT_SHIFTED_GRAD_OP; o.shiftedGradOp = 1; }
        if (opt & 2) { options |= NALU_OPT_LUMPED_MASS; o.lumpedMass = 1; }
        k.options = options;
        k.turb[NALU_TM_BETA_STAR] = o.betaStar = 0.09;
        k.turb[NALU_TM_TKE_PROD_LIMIT_RATIO] = o.tkeProdLimitRatio = (trial & 1) ? 0.5 : 10.0;
        k.turb[NALU_TM_SIGMA_W_TWO] = o.sigmaWTwo = 0.856;
        k.turb[NALU_TM_BETA_ONE] = o.betaOne = 0.075;
        k.turb[NALU_TM_BETA_TWO] = o.betaTwo = 0.0828;
        k.turb[NALU_TM_GAMMA_ONE] = o.gammaOne = 5.0 / 9.0;
        k.turb[NALU_TM_GAMMA_TWO] = o.gammaTwo = 0.44;
        k.turb[NALU_TM_CDES_KW] = o.cDESkw = 0.78;
        k.turb[NALU_TM_CDES_KE] = o.cDESke = 0.61;
        k.turb[NALU_TM_LRKSGS_FAC] = o.lrksgsfac = (opt & 1) ? 1.0 : 0.0;
        k.nodal[NALU_F_COORDINATES] = o.coordinates = coords;
        k.nodal[NALU_F_VELOCITY] = o.velocity = vel;
        k.nodal[NALU_F_DENSITY] = o.density = rho;
        k.nodal[NALU_F_VISCOSITY] = o.viscosity = visc;
        k.nodal[NALU_F_TKE] = o.tke = tke;
        k.nodal[NALU_F_SDR] = o.sdr = sdr;
        k.nodal[NALU_F_TVISC] = o.tvisc = tvisc;
        k.nodal[NALU_F_FONE_BLEND] = o.fOneBlend = f1;
        k.nodal[NALU_F_SST_MAX_LENGTH_SCALE] = o.maxLengthScale = lmax;
        k.nodal[NALU_F_C_EPSILON] = o.cEps = ceps;
        k.nodal[NALU_F_DSQRTK_DX_SQ] = o.dsqrtkSq = dsq;
        k.nodal[NALU_F_DUAL_VOLUME] = o.dualNodalVolume = dvol;
        k.nodal[NALU_F_GJU] = o.Gju = gju;
        std::vector<nalu_b200_kernel_desc> ks{k};
        std::vector<orc_kernel_desc> os{o};
        const double e = run_case<1>(ks, os, nodes, coords);
        if (e > 1e-13) { printf("FAIL turbulence source kind %d opt %d trial %d relerr %.3e\n", kind, opt, trial, e); return 1; }
        worst = fmax(worst, e);
      }
    }

    // momentum NSO: shifted grad-op x divU, pde / linearized / blended residual, fourth-order term, BDF2 time term
    for (int opt = 0; opt < 8; ++opt) {
      static double gju[NN * 9], rtm[NN * 3];
      for (int n = 0; n < NN * 9; ++n) gju[n] = frand();
      for (int n = 0; n < NN * 3; ++n) rtm[n] = frand();
      nalu_b200_kernel_desc k;
      orc_kernel_desc o;
      memset(&k, 0, sizeof(k));
      memset(&o, 0, sizeof(o));
      k.kind = NALU_K_MOMENTUM_NSO; o.kind = ORC_MOMENTUM_NSO;
      if (opt & 1) { k.options |= NALU_OPT_SHIFTED_GRAD_OP; o.shiftedGradOp = 1; }
      k.include_divu = o.includeDivU = (opt & 2) ? 1.0 : 0.0;
      k.nso_alt_res_fac = o.altResFac = (opt & 4) ? 1.0 : 0.25 * (trial % 3);
      k.nso_fourth_fac = o.fourthFac = (trial & 1) ? 1.0 : 0.0;
      k.dt = o.dt = 0.05; k.gamma1 = o.gamma1 = 1.5; k.gamma2 = o.gamma2 = -2.0; k.gamma3 = o.gamma3 = 0.5;
      k.nodal[NALU_F_COORDINATES] = o.coordinates = coords;
      k.nodal[NALU_F_VELOCITY] = o.velocity = vel;
      k.nodal[NALU_F_VELOCITY_N] = o.velocityN = velN;
      k.nodal[NALU_F_VELOCITY_NM1] = o.velocityNm1 = velNm1;
      if (opt & 2) k.nodal[NALU_F_VELOCITY_RTM] = o.velocityRTM = rtm;
      k.nodal[NALU_F_DENSITY] = o.density = rho;
      k.nodal[NALU_F_DENSITY_N] = o.densityN = rhoN;
      k.nodal[NALU_F_DENSITY_NM1] = o.densityNm1 = rhoNm1;
      k.nodal[NALU_F_VISCOSITY] = o.viscosity = visc;
      k.nodal[NALU_F_PRESSURE] = o.pressure = p;
      k.nodal[NALU_F_GJU] = o.Gju = gju;
      std::vector<nalu_b200_kernel_desc> ks{k};
      std::vector<orc_kernel_desc> os{o};
      const double e = run_case<3>(ks, os, nodes, coords);
      if (e > 1e-13) { printf("FAIL momentum NSO opt %d trial %d relerr %.3e\n", opt, trial, e); return 1; }
      worst = fmax(worst, e);
    }

    for (int kind = 0; kind < 7; ++kind) {
      for (int opt = 0; opt < 4; ++opt) {
        nalu_b200_kernel_desc k;
        orc_kernel_desc o;
        memset(&k, 0, sizeof(k));
        memset(&o, 0, sizeof(o));
        k.kind = kind; o.kind = kind;
        k.include_divu = o.includeDivU = (opt & 1) ? 1.0 : 0.0;
        k.dens_fac = o.densFac = (opt & 2) ? 1.0 : 0.0;
        k.om_dens_fac = o.om_densFac = (opt & 2) ? 0.0 : 1.0;
        k.dt = o.dt = 0.1 + 0.01 * trial; k.gamma1 = o.gamma1 = 1.5; k.gamma2 = o.gamma2 = -2.0; k.gamma3 = o.gamma3 = 0.5;
        int options = 0;
        if (kind == NALU_K_MOMENTUM_ADV_DIFF || kind == NALU_K_SCALAR_ADV_DIFF || kind == NALU_K_SCALAR_DIFF) {
          if (opt & 1) { options |= NALU_OPT_SHIFTED_GRAD_OP; o.shiftedGradOp = 1; }
          if (opt & 2) { options |= NALU_OPT_SKEW_SYMMETRIC; o.skewSymmetric = 1; }
        } else if (kind == NALU_K_CONTINUITY_ADV) {
          if (opt & 1) { options |= NALU_OPT_SHIFT_POISSON; o.shiftPoisson = 1; }
          if (opt & 2) { options |= NALU_OPT_REDUCED_SENS; o.reducedSensitivities = 1; }
          if (trial & 1) { options |= NALU_OPT_SHIFT_MDOT; o.shiftMdot = 1; }
        } else {
          if (opt & 1) { options |= NALU_OPT_LUMPED_MASS; o.lumpedMass = 1; }
        }
        k.options = options;
        k.nodal[NALU_F_COORDINATES] = o.coordinates = coords;
        k.nodal[NALU_F_VELOCITY] = o.velocity = vel;
        k.nodal[NALU_F_VELOCITY_N] = o.velocityN = velN;
        k.nodal[NALU_F_VELOCITY_NM1] = o.velocityNm1 = velNm1;
        k.nodal[NALU_F_VISCOSITY] = o.viscosity = visc;
        k.nodal[NALU_F_DENSITY] = o.density = rho;
        k.nodal[NALU_F_DENSITY_N] = o.densityN = rhoN;
        k.nodal[NALU_F_DENSITY_NM1] = o.densityNm1 = rhoNm1;
        k.nodal[NALU_F_PRESSURE] = o.pressure = p;
        k.nodal[NALU_F_DPDX] = o.Gpdx = gp;
        k.nodal[NALU_F_SCALARQ] = o.scalarQ = q;
        k.nodal[NALU_F_SCALARQ_N] = o.scalarQN = qN;
        k.nodal[NALU_F_SCALARQ_NM1] = o.scalarQNm1 = qNm1;
        k.mass_flow_rate = o.massFlowRate = mdot;
        const bool vec = kind == NALU_K_MOMENTUM_ADV_DIFF || kind == NALU_K_MOMENTUM_MASS;
        std::vector<nalu_b200_kernel_desc> ks{k};
        std::vector<orc_kernel_desc> os{o};
        // also exercise accumulation of two kernels on one scratch (mass + adv-diff)
        if (kind == NALU_K_MOMENTUM_ADV_DIFF || kind == NALU_K_SCALAR_ADV_DIFF) {
          nalu_b200_kernel_desc k2 = k; orc_kernel_desc o2 = o;
          k2.kind = o2.kind = vec ? NALU_K_MOMENTUM_MASS : NALU_K_SCALAR_MASS;
          k2.options = NALU_OPT_LUMPED_MASS; o2.lumpedMass = 1;
          ks.insert(ks.begin(), k2); os.insert(os.begin(), o2);
        }
        const double e = vec ? run_case<3>(ks, os, nodes, coords) : run_case<1>(ks, os, nodes, coords);
        if (e > 1e-13) { printf("FAIL kind %d opt %d trial %d relerr %.3e\n", kind, opt, trial, e); return 1; }
        worst = fmax(worst, e);
      }
    }
  }
  printf("worst relerr %.3e\n", worst);
  return 0;
}
