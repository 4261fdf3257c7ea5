// assemble_slist_box.cu -- launch side of the tiled write-once assembly of the scalar transport kernel LISTS (slist_tile.cuh):
//   [ScalarMassElemKernel(lumped)]? + one of {ScalarUpwAdvDiff, ScalarAdvDiff, ScalarDiff}ElemKernel
//   + [TurbKineticEnergySSTSrcElemKernel | SpecificDissipationRateSSTSrcElemKernel]?
// in ONE launch -- the tke and sdr lists of BASELINE config 4 (src/TurbKineticEnergyEquationSystem.C:447-494,
// src/SpecificDissipationRateEquationSystem.C:334-361) through AssembleElemSolverAlgorithm::execute
// (src/AssembleElemSolverAlgorithm.C:74-100) + TpetraLinearSystem::sumInto (src/TpetraLinearSystem.C:1457-1522).
//
// Grid: one CTA (128 threads) per (15 x 7 node tile, z chunk), two CTAs per SM.  Per element layer ez a CTA runs
//   1b  thread = element: the 8 SCV records (volume, source lhs factor, source rhs)        [only with a mass / source kernel]
//   2b  thread = node: SCV records -> rows of planes ez and ez+1, dual volume
//   1a  thread = element: the 12 SCS records (diffusion vector, mdot, Peclet factor, flux)
//   2a  thread = node: SCS records -> rows of planes ez and ez+1  ||  cp.async of node plane ez+2 into the slot of plane ez
//   the row of plane ez is complete: lumped mass from the dual volume, one write to HBM
// with one CTA barrier after each phase (the two record passes share one shared-memory buffer).
#include "slist_tile.cuh"
#include "internal.h"

namespace nalu_b200 {

using slist::SlistParams;
using SC = slist::Cfg;

template <bool SH, bool SK>
__global__ void __launch_bounds__(SC::NT, 2) assemble_slist_box_kernel(const __grid_constant__ SlistParams P)
{
  extern __shared__ __align__(16) double sm[];
  double *const ring = sm + SC::OFF_RING;
  const int tid = threadIdx.x;
  const int ntile = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % ntile, zc = blockIdx.x / ntile;
  const int x0 = (tile % P.tiles_x) * SC::TBX, y0 = (tile / P.tiles_x) * SC::TBY;
  const int zs = P.z_first + zc * P.z_chunk;
  const int ze = (zs + P.z_chunk < P.z_last) ? zs + P.z_chunk : P.z_last;
  const int lx = tid % SC::ELX, ly = tid / SC::ELX;
  const bool in_tile = lx < SC::TBX && ly < SC::TBY;
  const bool active = in_tile && x0 + lx < P.NX && y0 + ly < P.NY;
  const bool scv_pass = P.mass || P.src;
  const bool src_lumped = P.src_lumped != 0;
  const slist::UpwK K = {P.alpha, P.alpha_upw};

  double accA[27], accB[27], rA = 0.0, rB = 0.0, volA = 0.0, volB = 0.0;
#pragma unroll
  for (int k = 0; k < 27; ++k) accA[k] = accB[k] = 0.0;

  // node planes zs-1 and zs (ring slot of plane z: (z + 2) % 2)
  slist::stage_plane(P, ring + ((zs + 1) & 1) * SC::PLN, tid, x0, y0, zs - 1);
  slist::stage_plane(P, ring + (zs & 1) * SC::PLN, tid, x0, y0, zs);
  cont::stage_wait();
  __syncthreads();

  for (int ez = zs - 1; ez < ze; ++ez) {
    const double *pl0 = ring + ((ez + 2) & 1) * SC::PLN, *pl1 = ring + ((ez + 1) & 1) * SC::PLN;
    if (scv_pass) {
      slist::p1b_element(P, sm, tid, x0, y0, ez, pl0, pl1);
      __syncthreads();
      if (in_tile) slist::p2b_layer(sm, lx, ly, src_lumped, accA, rA, volA, accB, rB, volB);
      __syncthreads();
    }
    slist::p1a_element(P, sm, tid, x0, y0, ez, pl0, pl1);
    __syncthreads();
    // plane ez is dead: plane ez+2 lands in its slot while phase 2a runs
    slist::stage_plane(P, ring + ((ez + 2) & 1) * SC::PLN, tid, x0, y0, ez + 2);
    if (in_tile) slist::p2a_layer<SH, SK>(K, sm, lx, ly, accA, rA, accB, rB);
    if (active && ez >= zs) {
      if (P.mass) slist::mass_apply(P, x0 + lx, y0 + ly, ez, volA, accA[13], rA);
      cont::write_row(P, x0 + lx, y0 + ly, ez, accA, rA);
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      accA[k] = accB[k];
      accB[k] = 0.0;
    }
    rA = rB;
    rB = 0.0;
    volA = volB;
    volB = 0.0;
    cont::stage_wait();
    __syncthreads();
  }
}

namespace {

bool is_advdiff(int kind) { return kind == NALU_K_SCALAR_UPW_ADV_DIFF || kind == NALU_K_SCALAR_ADV_DIFF || kind == NALU_K_SCALAR_DIFF; }
bool is_src(int kind) { return kind == NALU_K_TKE_SST_SRC || kind == NALU_K_SDR_SST_SRC; }

// positions of the list's members; false when the list is not of the supported form
bool classify(const nalu_b200_kernel_desc *k, int nk, int &ia, int &im, int &is)
{
  ia = im = is = -1;
  for (int i = 0; i < nk; ++i) {
    if (is_advdiff(k[i].kind)) {
      if (ia >= 0) return false;
      ia = i;
    } else if (k[i].kind == NALU_K_SCALAR_MASS) {
      if (im >= 0 || !(k[i].options & NALU_OPT_LUMPED_MASS)) return false;
      im = i;
    } else if (is_src(k[i].kind)) {
      if (is >= 0) return false;
      is = i;
    } else {
      return false;
    }
  }
  return ia >= 0;
}

} // namespace

bool slist_box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (g.box[0] <= 0 || !g.d_node_slot27 || g.ndof != 1 || nk < 1 || nk > 3) return false;
  int ia, im, is;
  if (!classify(k, nk, ia, im, is)) return false;
  const nalu_b200_kernel_desc &a = k[ia];
  if (nk == 1 && a.kind != NALU_K_SCALAR_UPW_ADV_DIFF) return false; // the central kernels alone: assemble_scalar_box
  const double *coords = a.nodal[NALU_F_COORDINATES], *q = a.nodal[NALU_F_SCALARQ];
  if (!coords || !q || !a.nodal[NALU_F_VISCOSITY]) return false;
  if (a.kind != NALU_K_SCALAR_DIFF && !a.mass_flow_rate) return false;
  const double *vel = nullptr, *rho = nullptr;
  if (a.kind == NALU_K_SCALAR_UPW_ADV_DIFF) {
    vel = a.nodal[NALU_F_VELOCITY_RTM] ? a.nodal[NALU_F_VELOCITY_RTM] : a.nodal[NALU_F_VELOCITY];
    rho = a.nodal[NALU_F_DENSITY];
    if (!vel || !rho || !a.nodal[NALU_F_GJQ]) return false;
  }
  if (im >= 0) {
    const nalu_b200_kernel_desc &m = k[im];
    if (m.nodal[NALU_F_COORDINATES] != coords || m.nodal[NALU_F_SCALARQ] != q) return false;
    if (!m.nodal[NALU_F_SCALARQ_N] || !m.nodal[NALU_F_SCALARQ_NM1] || !m.nodal[NALU_F_DENSITY] || !m.nodal[NALU_F_DENSITY_N] ||
        !m.nodal[NALU_F_DENSITY_NM1])
      return false;
    if (rho && m.nodal[NALU_F_DENSITY] != rho) return false;
    rho = m.nodal[NALU_F_DENSITY];
  }
  if (is >= 0) {
    const nalu_b200_kernel_desc &s = k[is];
    const bool tke = s.kind == NALU_K_TKE_SST_SRC;
    if (s.nodal[NALU_F_COORDINATES] != coords) return false;
    // the source's own scalar is the list's dof
    if (s.nodal[tke ? NALU_F_TKE : NALU_F_SDR] != q || !s.nodal[tke ? NALU_F_SDR : NALU_F_TKE]) return false;
    if (!s.nodal[NALU_F_VELOCITY] || !s.nodal[NALU_F_DENSITY] || !s.nodal[NALU_F_TVISC]) return false;
    if (!tke && !s.nodal[NALU_F_FONE_BLEND]) return false;
    if (vel && s.nodal[NALU_F_VELOCITY] != vel) return false;
    if (rho && s.nodal[NALU_F_DENSITY] != rho) return false;
  }
  return true;
}

template <bool SH, bool SK>
static int launch_slist(SlistParams &P, cudaStream_t stream)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  P.tiles_x = (P.NX + SC::TBX - 1) / SC::TBX;
  P.tiles_y = (P.NY + SC::TBY - 1) / SC::TBY;
  const long long ntile = (long long)P.tiles_x * P.tiles_y;
  // z chunks: enough CTAs for ~12 waves of 2 CTAs per SM, at least 16 planes per chunk (a chunk recomputes one element layer)
  long long want = ((long long)sms * 2 * 12 + ntile - 1) / ntile;
  const int nplanes = P.z_last - P.z_first;
  int zchunks = (int)(want < 1 ? 1 : want);
  int zc = (nplanes + zchunks - 1) / zchunks;
  if (zc < 16) zc = 16;
  if (tuning().box_zchunk > 0) zc = tuning().box_zchunk;
  if (zc > nplanes) zc = nplanes;
  P.z_chunk = zc;
  zchunks = (nplanes + zc - 1) / zc;
  const size_t smem = sizeof(double) * SC::SMEM_DOUBLES;
  NALU_CUDA(cudaFuncSetAttribute(assemble_slist_box_kernel<SH, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NALU_CUDA(cudaFuncSetAttribute(assemble_slist_box_kernel<SH, SK>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_slist_box: grid too large");
    return NALU_B200_ERR_ARG;
  }
  assemble_slist_box_kernel<SH, SK><<<(unsigned)ctas, SC::NT, smem, stream>>>(P);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

int launch_assemble_slist_box(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs,
                              int overwrite, cudaStream_t stream)
{
  if (!slist_box_supported(g, k, nk)) {
    set_error("assemble_slist_box: unsupported mesh / kernel list");
    return NALU_B200_ERR_ARG;
  }
  if (g.n_elems == 0) return NALU_B200_OK;
  int ia, im, is;
  classify(k, nk, ia, im, is);
  const nalu_b200_kernel_desc &a = k[ia];
  SlistParams P = {};
  P.nx = g.box[0]; P.ny = g.box[1]; P.nz = g.box[2];
  P.NX = P.nx + 1; P.NY = P.ny + 1; P.NZ = P.nz + 1;
  P.coords = a.nodal[NALU_F_COORDINATES];
  P.q = a.nodal[NALU_F_SCALARQ];
  P.gam = a.nodal[NALU_F_VISCOSITY];
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
    P.q_n = m.nodal[NALU_F_SCALARQ_N];
    P.q_nm1 = m.nodal[NALU_F_SCALARQ_NM1];
    P.rho_n = m.nodal[NALU_F_DENSITY_N];
    P.rho_nm1 = m.nodal[NALU_F_DENSITY_NM1];
    P.g1_dt = m.gamma1 / m.dt;
    P.g2_dt = m.gamma2 / m.dt;
    P.g3_dt = m.gamma3 / m.dt;
  }
  if (is >= 0) {
    const nalu_b200_kernel_desc &s = k[is];
    const bool tke = s.kind == NALU_K_TKE_SST_SRC;
    P.src = tke ? 1 : 2;
    P.src_lumped = (s.options & NALU_OPT_LUMPED_MASS) != 0;
    P.src_shifted = (s.options & NALU_OPT_SHIFTED_GRAD_OP) != 0;
    P.other = s.nodal[tke ? NALU_F_SDR : NALU_F_TKE];
    P.vel = s.nodal[NALU_F_VELOCITY];
    P.rho = s.nodal[NALU_F_DENSITY];
    P.tvisc = s.nodal[NALU_F_TVISC];
    P.f1 = tke ? nullptr : s.nodal[NALU_F_FONE_BLEND];
    for (int i = 0; i < 10; ++i) P.turb[i] = s.turb[i];
  }
  P.node_row = g.d_node_row;
  P.browptr = (const long long *)g.d_browptr;
  P.slot27 = g.d_node_slot27;
  P.values = d_values;
  P.rhs = d_rhs;
  P.err = g.d_err;
  P.overwrite = overwrite;
  P.z_first = 0;
  P.z_last = P.NZ;
  const bool sh = (P.options & NALU_OPT_SHIFTED_GRAD_OP) != 0, sk = (P.options & NALU_OPT_SKEW_SYMMETRIC) != 0;
  if (sh) return sk ? launch_slist<true, true>(P, stream) : launch_slist<true, false>(P, stream);
  return sk ? launch_slist<false, true>(P, stream) : launch_slist<false, false>(P, stream);
}

} // namespace nalu_b200
