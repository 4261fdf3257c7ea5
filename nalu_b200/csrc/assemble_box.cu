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

bool box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (g.box[0] <= 0 || !g.d_node_slot27 || g.ndof != 3 || nk != 1) return false;
  if (k[0].kind != NALU_K_MOMENTUM_ADV_DIFF) return false;
  if (!k[0].nodal[NALU_F_COORDINATES] || !k[0].nodal[NALU_F_VELOCITY] || !k[0].nodal[NALU_F_VISCOSITY] ||
      !k[0].mass_flow_rate)
    return false;
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
  BoxParams P;
  P.nx = g.box[0]; P.ny = g.box[1]; P.nz = g.box[2];
  P.NX = P.nx + 1; P.NY = P.ny + 1; P.NZ = P.nz + 1;
  P.coords = k[0].nodal[NALU_F_COORDINATES];
  P.vel = k[0].nodal[NALU_F_VELOCITY];
  P.visc = k[0].nodal[NALU_F_VISCOSITY];
  P.mdot = k[0].mass_flow_rate;
  P.node_row = g.d_node_row;
  P.browptr = (const long long *)g.d_browptr;
  P.slot27 = g.d_node_slot27;
  P.node_desc = (const unsigned long long *)g.d_node_desc;
  P.values = d_values;
  P.rhs = d_rhs;
  P.err = g.d_err;
  P.include_divu = k[0].include_divu;
  P.options = k[0].options;
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
