// assemble_cont_box.cu -- launch side of the tiled write-once continuity assembly (see cont_tile.cuh).
//
// Grid: one CTA (128 threads) per (15 x 7 node tile, z chunk), two CTAs per SM.  Per element layer ez a CTA runs
//   1  phase 1: thread = element of the layer (tile + halo): 12 records  ||  cp.async of node plane ez+2
//   2  phase 2: thread = node: the layer's contribution to the rows of planes ez and ez+1; the row of plane ez
//      is complete and goes to HBM
// with one CTA barrier after each.
#include "cont_tile.cuh"
#include "internal.h"

#include <cstdlib>

namespace nalu_b200 {

using cont::ContParams;
using CC = cont::Cfg;

template <bool SH>
__global__ void __launch_bounds__(CC::NT, 2) assemble_cont_box_kernel(const __grid_constant__ ContParams P)
{
  extern __shared__ __align__(16) double sm[];
  double *const ring = sm + CC::OFF_RING;
  const int tid = threadIdx.x;
  const int ntile = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % ntile, zc = blockIdx.x / ntile;
  const int x0 = (tile % P.tiles_x) * CC::TBX, y0 = (tile / P.tiles_x) * CC::TBY;
  const int zs = P.z_first + zc * P.z_chunk;
  const int ze = (zs + P.z_chunk < P.z_last) ? zs + P.z_chunk : P.z_last;
  const int lx = tid % CC::ELX, ly = tid / CC::ELX;
  const bool in_tile = lx < CC::TBX && ly < CC::TBY;
  const bool active = in_tile && x0 + lx < P.NX && y0 + ly < P.NY;

  double accA[27], accB[27], rA = 0.0, rB = 0.0;
#pragma unroll
  for (int k = 0; k < 27; ++k) accA[k] = accB[k] = 0.0;

  // node planes zs-1 and zs (ring slot of plane z: (z + 3) % 3)
  cont::stage_plane(P, ring + ((zs + 2) % 3) * CC::PLN, tid, x0, y0, zs - 1);
  cont::stage_plane(P, ring + (zs % 3) * CC::PLN, tid, x0, y0, zs);
  cont::stage_wait();
  __syncthreads();

  for (int ez = zs - 1; ez < ze; ++ez) {
    // plane ez+2 lands while phase 1 and phase 2 of this layer run (its slot held plane ez-1, dead)
    cont::stage_plane(P, ring + ((ez + 2 + 3) % 3) * CC::PLN, tid, x0, y0, ez + 2);
    cont::p1_element_c(P, sm, tid, x0, y0, ez, ring + ((ez + 3) % 3) * CC::PLN, ring + ((ez + 1 + 3) % 3) * CC::PLN);
    __syncthreads();
    if (in_tile) cont::p2_layer<SH>(sm, lx, ly, accA, rA, accB, rB);
    if (active && ez >= zs) cont::write_row(P, x0 + lx, y0 + ly, ez, accA, rA);
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      accA[k] = accB[k];
      accB[k] = 0.0;
    }
    rA = rB;
    rB = 0.0;
    cont::stage_wait();
    __syncthreads();
  }
}

bool cont_box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (g.box[0] <= 0 || !g.d_node_slot27 || g.ndof != 1 || nk != 1) return false;
  if (k[0].kind != NALU_K_CONTINUITY_ADV) return false;
  if (!k[0].nodal[NALU_F_COORDINATES] || !k[0].nodal[NALU_F_VELOCITY] || !k[0].nodal[NALU_F_DENSITY] ||
      !k[0].nodal[NALU_F_PRESSURE] || !k[0].nodal[NALU_F_DPDX])
    return false;
  return true;
}

template <bool SH>
static int launch_cont(ContParams &P, cudaStream_t stream)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  P.tiles_x = (P.NX + CC::TBX - 1) / CC::TBX;
  P.tiles_y = (P.NY + CC::TBY - 1) / CC::TBY;
  const long long ntile = (long long)P.tiles_x * P.tiles_y;
  // z chunks: enough CTAs for ~12 waves of 2 CTAs per SM, at least 16 planes per chunk (a chunk recomputes
  // one element layer)
  long long want = ((long long)sms * 2 * 12 + ntile - 1) / ntile;
  const int nplanes = P.z_last - P.z_first;
  int zchunks = (int)(want < 1 ? 1 : want);
  int zc = (nplanes + zchunks - 1) / zchunks;
  if (zc < 16) zc = 16;
  if (tuning().box_zchunk > 0) zc = tuning().box_zchunk;
  if (zc > nplanes) zc = nplanes;
  P.z_chunk = zc;
  zchunks = (nplanes + zc - 1) / zc;
  const size_t smem = sizeof(double) * CC::SMEM_DOUBLES;
  NALU_CUDA(cudaFuncSetAttribute(assemble_cont_box_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NALU_CUDA(cudaFuncSetAttribute(assemble_cont_box_kernel<SH>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_cont_box: grid too large");
    return NALU_B200_ERR_ARG;
  }
  assemble_cont_box_kernel<SH><<<(unsigned)ctas, CC::NT, smem, stream>>>(P);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

int launch_assemble_cont_box(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs,
                             int overwrite, cudaStream_t stream)
{
  if (!cont_box_supported(g, k, nk)) {
    set_error("assemble_cont_box: unsupported mesh / kernel list");
    return NALU_B200_ERR_ARG;
  }
  if (g.n_elems == 0) return NALU_B200_OK;
  ContParams P;
  P.nx = g.box[0]; P.ny = g.box[1]; P.nz = g.box[2];
  P.NX = P.nx + 1; P.NY = P.ny + 1; P.NZ = P.nz + 1;
  P.coords = k[0].nodal[NALU_F_COORDINATES];
  P.vel = k[0].nodal[NALU_F_VELOCITY];
  P.rho = k[0].nodal[NALU_F_DENSITY];
  P.p = k[0].nodal[NALU_F_PRESSURE];
  P.gp = k[0].nodal[NALU_F_DPDX];
  P.node_row = g.d_node_row;
  P.browptr = (const long long *)g.d_browptr;
  P.slot27 = g.d_node_slot27;
  P.values = d_values;
  P.rhs = d_rhs;
  P.err = g.d_err;
  P.tau = k[0].dt / k[0].gamma1;
  P.options = k[0].options;
  P.overwrite = overwrite;
  P.z_first = 0;
  P.z_last = P.NZ;
  const bool sh = (P.options & (NALU_OPT_SHIFT_POISSON | NALU_OPT_REDUCED_SENS)) != 0;
  return sh ? launch_cont<true>(P, stream) : launch_cont<false>(P, stream);
}

} // namespace nalu_b200
