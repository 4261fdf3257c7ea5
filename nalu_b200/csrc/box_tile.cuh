// box_tile.cuh -- what the tiled write-once kernels for structured Hex8 blocks share: launch parameters, Hex8 numbering
// helpers (Hex8CVFEM.C:264-276, 396-412) and the Walsh-Hadamard butterfly of the 8 nodal values of an element.
#pragma once
#include "../../include/nalu_b200.h"
#include "hex8_dev.cuh"

namespace nalu_b200 {
namespace box {

struct BoxParams {
  int nx, ny, nz;    // elements of the local block
  int NX, NY, NZ;    // nodes
  int z_chunk;       // node planes per CTA
  int tiles_x, tiles_y;
  const double *coords; // [nNodes][3]
  const double *vel;    // [nNodes][3]
  const double *visc;   // [nNodes]
  const double *mdot;   // [nElem][12]
  const int *node_row;
  const long long *browptr;
  const unsigned char *slot27;
  const unsigned long long *node_desc; // [nNodes][3] row descriptors (graph.h)
  double *values;
  double *rhs;
  int *err;
  double include_divu;
  int options;
  int overwrite;
  int z_first, z_last; // node planes [z_first, z_last) this launch assembles (default all)
  // row-thread kernel: magnitudes of the shape-function derivatives (box5_tile.cuh, p2k_init), read by the FMAs
  // straight from the constant bank the kernel parameters live in
  double kt[3], kh[2], f23;
  // fused lumped MomentumMassElemKernel (row-thread kernel only; fuse_mass = 0: adv-diff alone)
  int fuse_mass;
  const double *rho_np1, *rho_n, *rho_nm1; // [nNodes]
  const double *vel_n, *vel_nm1, *dpdx;    // [nNodes][3]
  double g1_dt, g2_dt, g3_dt;              // gamma / dt
  double dens_fac, om_dens_fac;
};

// ---- small integer helpers on the Hex8 numbering (Hex8CVFEM.C:264-276, 396-412) -------------
__host__ __device__ __forceinline__ int node_of_bits(int bx, int by, int bz) { return bz * 4 + by * 2 + (bx ^ by); }
__host__ __device__ __forceinline__ int bit_x(int n) { return (n ^ (n >> 1)) & 1; }
__host__ __device__ __forceinline__ int bit_y(int n) { return (n >> 1) & 1; }
__host__ __device__ __forceinline__ int bit_z(int n) { return (n >> 2) & 1; }
// the SCS ip on the edge that leaves node k in direction dir (0 x, 1 y, 2 z)
__host__ __device__ __forceinline__ int ip_of(int k, int dir)
{
  const int bz = bit_z(k);
  if (dir == 0) return bit_y(k) ? (bz ? 6 : 2) : (bz ? 4 : 0);
  if (dir == 1) return bit_x(k) ? (bz ? 5 : 1) : (bz ? 7 : 3);
  return 8 + (k & 3);
}

// C[mask] = sum_m prod_{axis in mask} sign_axis(m) f_m   (mask bits = axes; f is overwritten)
__host__ __device__ __forceinline__ void butterfly8(double *f)
{
#pragma unroll
  for (int h = 1; h < 8; h <<= 1)
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (!(m & h)) {
        const double lo = f[m], hi = f[m | h];
        f[m] = hi + lo;
        f[m | h] = hi - lo;
      }
}

} // namespace box
} // namespace nalu_b200
