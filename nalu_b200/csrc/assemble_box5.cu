// assemble_box5.cu -- launch side of the row-thread tiled assembly (see box5_tile.cuh).
//
// Grid: one CTA (128 threads = 4 warps = 4 x lines of 8 nodes, a quad of lanes per node) per (8x4 node tile, z chunk),
// two CTAs per SM.  Per node plane z a CTA runs
//   A  lower half: the 4 elements below every node (top-face + z-edge points of layer z-1): dz=-1 group finished,
//      written out; dz=0 group half done (stays in registers)
//   B  phase 1a: bottom-face + z-edge points of layer z (face units on warps 0-1, z units on warps 2-3)
//   C  upper half: dz=0 and dz=+1 groups finished, written out; rhs
//   D  phase 1b: top-face points of layer z (x-edge pairs on warps 0-1, y-edge pairs on warps 2-3), metadata of plane
//      z+1, field prefetch of plane z+2
// with one CTA barrier between consecutive phases; inside A and C a warp synchronises only with itself.
#include "box5_tile.cuh"
#include "internal.h"

#include <cstdint>

namespace nalu_b200 {

using box::BoxParams;
using C5 = box5::Cfg;

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// diagonal term + rhs share of one finished dz group (box5_tile.cuh, fin_apply), quad exchange by shuffles
__device__ __forceinline__ void finalize_group(const box5::Thr &N, const double *sm, int uoff, double (&acc)[9][3], double &rr,
                                               double f23, int lane)
{
  double D[9], T[3];
#pragma unroll
  for (int c = 0; c < 9; ++c) D[c] = shfl_d(acc[c][0], lane | 3);
  const double r = box5::fin_apply(N, sm, uoff, acc, D, T);
  double t = 0.0;
  if (f23 != 0.0) {
    const int q = lane & ~3;
    t = T[0] + shfl_d(T[1], q | ((N.i + 2) % 3)) + shfl_d(T[2], q | ((N.i + 1) % 3));
  }
  rr += f23 * t - r;
}

// the 24 runs (8 nodes x 3 dof rows) of dz group g of the warp's x line: registers -> warp scratch -> CSR
template <bool OW>
__device__ __forceinline__ void output_group(const box5::Thr &N, const box5::Meta &M, const unsigned char *sl, int g, bool natural,
                                             double *scr, const double (&acc)[9][3], double *values, int lane)
{
  const int n0 = C5::TBX * N.ly;
  if (natural) {
    box5::out_stage_natural(N, scr, acc);
    __syncwarp();
    double *base = values + M.rowbase(n0) + 3 * M.s0(n0, g) + lane;
    if (lane < 27) {
#pragma unroll
      for (int r = 0; r < 24; ++r) {
        const double v = scr[r * 27 + lane];
        double *dst = base + (r / 3) * 243 + (r % 3) * 81;
        if (OW) *dst = v;
        else *dst += v;
      }
    }
    __syncwarp();
  } else {
    long long G = 0;
    int L = 0;
    if (N.i < 3 && N.active) {
      box5::out_stage_slots(N, M, sl, g, scr, acc);
      G = box5::run_first(M, N.node, N.i, g);
      L = 3 * M.cnt(N.node, g);
    }
    __syncwarp();
#pragma unroll 4
    for (int r = 0; r < 24; ++r) {
      const int src = (r / 3) * 4 + r % 3;
      const long long Gr = __shfl_sync(FULL, G, src);
      const int Lr = __shfl_sync(FULL, L, src);
      if (lane < Lr) {
        const double v = scr[r * 27 + lane];
        double *dst = values + Gr + lane;
        if (OW) *dst = v;
        else *dst += v;
      }
    }
    __syncwarp();
  }
}

template <bool OW>
__global__ void __launch_bounds__(C5::NT, 2) assemble_box5_kernel(const __grid_constant__ BoxParams P)
{
  extern __shared__ __align__(16) double sm[];
  double *const rec = sm + C5::OFF_REC, *const ring = sm + C5::OFF_RING, *const mdb = sm + C5::OFF_MD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double *const scr = sm + C5::OFF_SCR + warp * C5::SCRW;
  const int ntile = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % ntile, zc = blockIdx.x / ntile;
  const int x0 = (tile % P.tiles_x) * C5::TBX, y0 = (tile / P.tiles_x) * C5::TBY;
  const int zs = P.z_first + zc * P.z_chunk;
  const int ze = (zs + P.z_chunk < P.z_last) ? zs + P.z_chunk : P.z_last;

  // what lives across phases: the dz=0 group between the two halves of a plane (across phase 1a), the rhs of dof row i.
  // The accumulators are given their zeros right before the sweep, never after the output: a value that stays live
  // across phase 1, even as zero, costs phase 1 its registers (measured: it ends up in local memory).
  double acc1[9][3], rr = 0.0, vol = 0.0;
  double *const scv = sm + C5::OFF_SCV;

  // records of elements outside the domain are never written: they must read zero (phase 2 does not test)
  for (int i = tid; i < C5::RECSZ; i += C5::NT) rec[i] = 0.0;
  for (int i = tid; i < C5::NEL * 16; i += C5::NT) scv[i] = 0.0;
  // ---- prologue: node planes zs-1 .. zs+1, the mass flow rates of the layer below the first plane
  if (warp < 3) {
    box5::stage_plane(P, ring + ((zs + 2 + warp) % 3) * C5::PLN, lane, x0, y0, zs - 1 + warp);
    box5::stage_plane(P, ring + ((zs + 2 + warp) % 3) * C5::PLN, lane + 32, x0, y0, zs - 1 + warp);
  }
  box5::stage_mdot(P, mdb, tid, x0, y0, zs - 1);
  if (warp == 3) box5::meta_load(P, sm, zs & 1, lane, x0, y0, zs);
  box5::stage_wait();
  __syncthreads();

  // One iteration = one z half of a plane (H = 0: the elements below, H = 1: above) followed by the phase 1 that
  // prepares the next half; every piece of code appears once (instruction-cache footprint).  A chunk that does not
  // start at the bottom plane runs one extra leading iteration that only computes the top-face and z-edge points of
  // the element layer below its first plane.
  for (int hs = zs >= 1 ? 2 * zs - 1 : 2 * zs; hs < 2 * ze; ++hs) {
    const int z = hs >> 1, H = hs & 1;
    const bool lead = hs < 2 * zs;
    const bool lower = z >= 1, upper = z < P.nz;
    if (!lead) {
      const box5::Thr N = box5::thr_make(P, tid, x0, y0);
      const box5::Meta M = box5::meta_at(sm, z & 1);
      const bool act = H ? upper : lower;
      // ---- A (H = 0): dz=-1 group finished, dz=0 half done / C (H = 1): dz=0 and dz=+1 groups finished, rhs
      if (!H) {
        box5::stage_mdot(P, mdb, tid, x0, y0, z); // lands while the lower half is computed
#pragma unroll
        for (int c = 0; c < 9; ++c) acc1[c][0] = acc1[c][1] = acc1[c][2] = 0.0;
      }
      double acc2[9][3];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc2[c][0] = acc2[c][1] = acc2[c][2] = 0.0;
      const bool mass_lane = P.fuse_mass && H && N.i < 3 && N.active;
      const long long gnode = (x0 + N.lx) + (long long)P.NX * ((y0 + N.ly) + (long long)P.NY * z);
      if (mass_lane) box5::mass_prefetch(P, gnode, N.i);
      if (act) box5::sweep_half(N, sm, P, H, acc1, acc2);
      if (P.fuse_mass && act) vol += box5::dual_volume_half(N, sm, H);
      box5::MassNode mn;
      if (mass_lane) box5::mass_fetch(P, gnode, N.i, mn);
      // is the warp's x line stored back to back in lattice order?  (then the runs of a group sit at fixed offsets)
      const bool natural = __all_sync(FULL, box5::node_in_natural_line(M, N.node));
      const unsigned char *sl = P.slot27 + 27 * ((x0 + N.lx) + (long long)P.NX * ((y0 + N.ly) + (long long)P.NY * z));
      if (H) {
        finalize_group(N, sm, box5::vel_off(N, z % 3), acc1, rr, P.f23, lane);
        if (P.fuse_mass) {
          // lumped mass: diagonal entries of the node's own block (column 4 of the dz=0 group; the diagonal is rotated
          // index 0) and the residual, from the dual volume gathered over the two halves
          if (mass_lane) box5::mass_apply(P, mn, vol, acc1[4][0], rr);
          vol = 0.0;
        }
        output_group<OW>(N, M, sl, 1, natural, scr, acc1, P.values, lane);
      }
      if (act) {
        finalize_group(N, sm, box5::vel_off(N, (z + (H ? 1 : 2)) % 3), acc2, rr, P.f23, lane);
        output_group<OW>(N, M, sl, H ? 2 : 0, natural, scr, acc2, P.values, lane);
      }
      if (H) {
        if (N.i < 3 && N.active) {
          double *r = P.rhs + (long long)M.row(N.node) * 3 + N.i;
          *r = OW ? rr : *r + rr;
          if (OW && !M.natural(N.node)) box5::zero_uncovered(M, N.node, N.i, P.values);
        }
        rr = 0.0;
      } else {
        box5::stage_wait();
      }
      __syncthreads();
    }
    // ---- B (H = 0) / leading iteration: phase 1a: the face units of layer z (bottom face; top face in the leading
    //      iteration) on warps 0-1, its z units on warps 2-3: no warp runs both
    //      D (H = 1): phase 1b: top-face points of layer z, x-edge pairs on warps 0-1, y-edge pairs on warps 2-3;
    //      metadata of the next plane, field prefetch
    if (upper) {
      const double *pl0 = ring + (z % 3) * C5::PLN, *pl1 = ring + ((z + 1) % 3) * C5::PLN;
      const int t = tid & 63;
      if (!H || lead) {
        if (tid < 64) {
          if (t < C5::NEL) box5::p1_face_unit(P, t, lead ? 1 : 0, x0, y0, pl0, pl1, mdb, rec, scv);
        } else if (t < C5::NEL) box5::p1_z_unit(P, t, x0, y0, pl0, pl1, mdb, rec, scv);
      } else if (z + 1 < ze) {
        if (tid < 64) {
          if (t < C5::NEL) box5::p1_face_half<0>(P, t, 1, x0, y0, pl0, pl1, mdb, rec);
        } else if (t < C5::NEL) box5::p1_face_half<1>(P, t, 1, x0, y0, pl0, pl1, mdb, rec);
      }
    }
    if (H && !lead && z + 1 < ze) {
      if (warp == 3) box5::meta_load(P, sm, (z + 1) & 1, lane, x0, y0, z + 1);
      // node plane z+2 -> the ring slot of plane z-1 (dead since the dz=-1 group was finalized)
      if (warp == 1 || warp == 2) box5::stage_plane(P, ring + ((z + 2) % 3) * C5::PLN, lane + 32 * (warp - 1), x0, y0, z + 2);
    }
    __syncthreads();
  }
}

} // namespace

bool box5_supported(const BoxParams &P) { return box5::options_supported(P.options); }

int launch_assemble_box5(BoxParams &P, cudaStream_t stream)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  box5::p2k_init(P);
  P.tiles_x = (P.NX + C5::TBX - 1) / C5::TBX;
  P.tiles_y = (P.NY + C5::TBY - 1) / C5::TBY;
  const long long ntile = (long long)P.tiles_x * P.tiles_y;
  // z chunks: enough CTAs for ~24 waves of 2 CTAs per SM (tail effect; measured at 256^3: 2 chunks 840, 4-6 chunks 851
  // Melem/s), but at least 16 planes per chunk (each chunk recomputes the top-face / z-edge points of one element
  // layer in its leading iteration)
  long long want = ((long long)sms * 2 * 24 + ntile - 1) / ntile;
  int zchunks = (int)(want < 1 ? 1 : want);
  const int nplanes = P.z_last - P.z_first;
  int zc = (nplanes + zchunks - 1) / zchunks;
  if (zc < 16) zc = 16;
  if (tuning().box_zchunk > 0) zc = tuning().box_zchunk;
  if (zc > nplanes) zc = nplanes;
  P.z_chunk = zc;
  zchunks = (nplanes + zc - 1) / zc;
  const size_t smem = sizeof(double) * C5::SMEM_DOUBLES;
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_box5: grid too large");
    return NALU_B200_ERR_ARG;
  }
  if (P.overwrite) {
    NALU_CUDA(cudaFuncSetAttribute(assemble_box5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NALU_CUDA(cudaFuncSetAttribute(assemble_box5_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    assemble_box5_kernel<true><<<(unsigned)ctas, C5::NT, smem, stream>>>(P);
  } else {
    NALU_CUDA(cudaFuncSetAttribute(assemble_box5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NALU_CUDA(cudaFuncSetAttribute(assemble_box5_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    assemble_box5_kernel<false><<<(unsigned)ctas, C5::NT, smem, stream>>>(P);
  }
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

} // namespace nalu_b200
