// assemble_box.cu -- dispatch of the tiled write-once MomentumAdvDiff assembly on structured Hex8 blocks: the
// row-thread kernel (assemble_box5.cu / box5_tile.cuh) by default, the column-pair kernel (assemble_box2.cu /
// box2_tile.cuh) for the option combinations the row-thread kernel does not take (shifted gradient operator without
// skew-symmetric advection or the reverse).  Both implement AssembleElemSolverAlgorithm::execute +
// MomentumAdvDiffElemKernel::execute + TpetraLinearSystem::sumInto (src/AssembleElemSolverAlgorithm.C:74-100,
// src/kernel/MomentumAdvDiffElemKernel.C:70-186, src/TpetraLinearSystem.C:1457-1522).
#include "box_tile.cuh"
#include "internal.h"

namespace nalu_b200 {

using box::BoxParams;

// which descriptor of the list is the adv-diff kernel / the lumped mass kernel the row-thread kernel fuses with it
// (the momentum list of the low-Mach equation system, src/LowMachEquationSystem.C:1304-1398); -1: none
static bool split_list(const nalu_b200_kernel_desc *k, int nk, int &adv, int &mass)
{
  adv = mass = -1;
  if (nk < 1 || nk > 2) return false;
  for (int i = 0; i < nk; ++i) {
    if (k[i].kind == NALU_K_MOMENTUM_ADV_DIFF && adv < 0) adv = i;
    else if (k[i].kind == NALU_K_MOMENTUM_MASS && (k[i].options & NALU_OPT_LUMPED_MASS) && mass < 0) mass = i;
    else return false;
  }
  return adv >= 0;
}

bool box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (g.box[0] <= 0 || !g.d_node_slot27 || g.ndof != 3) return false;
  int adv, mass;
  if (!split_list(k, nk, adv, mass)) return false;
  if (missing_field(k[adv])) return false;
  if (mass >= 0) {
    // the fused list needs the row-thread kernel; both kernels must see the same coordinates
    BoxParams P;
    P.options = k[adv].options;
    if (tuning().box_kernel != 5 || !box5_supported(P) || missing_field(k[mass]) ||
        k[mass].nodal[NALU_F_COORDINATES] != k[adv].nodal[NALU_F_COORDINATES])
      return false;
  }
  return true;
}

int launch_assemble_box(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs,
                        int overwrite, cudaStream_t stream, int z_begin, int z_end)
{
  if (!box_supported(g, k, nk)) {
    set_error("assemble_box: unsupported mesh / kernel list");
    return NALU_B200_ERR_ARG;
  }
  if (g.n_elems == 0) return NALU_B200_OK;
  int adv, mass;
  split_list(k, nk, adv, mass);
  BoxParams P;
  P.nx = g.box[0]; P.ny = g.box[1]; P.nz = g.box[2];
  P.NX = P.nx + 1; P.NY = P.ny + 1; P.NZ = P.nz + 1;
  P.coords = k[adv].nodal[NALU_F_COORDINATES];
  P.vel = k[adv].nodal[NALU_F_VELOCITY];
  P.visc = k[adv].nodal[NALU_F_VISCOSITY];
  P.mdot = k[adv].mass_flow_rate;
  P.fuse_mass = mass >= 0;
  if (mass >= 0) {
    const nalu_b200_kernel_desc &m = k[mass];
    if (m.nodal[NALU_F_VELOCITY] != P.vel) {
      set_error("assemble_box: the mass and adv-diff kernels of a fused list must gather the same velocity (NP1) field");
      return NALU_B200_ERR_ARG;
    }
    P.rho_np1 = m.nodal[NALU_F_DENSITY]; P.rho_n = m.nodal[NALU_F_DENSITY_N]; P.rho_nm1 = m.nodal[NALU_F_DENSITY_NM1];
    P.vel_n = m.nodal[NALU_F_VELOCITY_N]; P.vel_nm1 = m.nodal[NALU_F_VELOCITY_NM1]; P.dpdx = m.nodal[NALU_F_DPDX];
    P.g1_dt = m.gamma1 / m.dt; P.g2_dt = m.gamma2 / m.dt; P.g3_dt = m.gamma3 / m.dt;
    P.dens_fac = m.dens_fac; P.om_dens_fac = m.om_dens_fac;
  }
  P.node_row = g.d_node_row;
  P.browptr = (const long long *)g.d_browptr;
  P.slot27 = g.d_node_slot27;
  P.node_desc = (const unsigned long long *)g.d_node_desc;
  P.values = d_values;
  P.rhs = d_rhs;
  P.err = g.d_err;
  P.include_divu = k[adv].include_divu;
  P.options = k[adv].options;
  P.overwrite = overwrite;
  P.z_first = 0;
  P.z_last = P.NZ;
  if (z_begin >= 0) {
    if (z_begin >= z_end || z_end > P.NZ) {
      set_error("assemble_box: bad node-plane range");
      return NALU_B200_ERR_ARG;
    }
    P.z_first = z_begin;
    P.z_last = z_end;
  }
  if (tuning().box_kernel == 5 && box5_supported(P)) return launch_assemble_box5(P, stream);
  return launch_assemble_box2(P, stream);
}

} // namespace nalu_b200
