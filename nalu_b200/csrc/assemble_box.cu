// assemble_box.cu -- launch side of the tiled write-once assembly (see box_tile.cuh).
//
// Grid: one CTA per (node tile, z chunk).  Per node plane a CTA runs
//   phase 1 (ip data of the element layer above the plane)  -> sync
//   lower octants (layer below)                              -> sync (that layer's buffer becomes the
//                                                               row staging area)
//   upper octants, block flush, rhs                           -> sync
//   coalesced copy of the finished rows to HBM               -> sync
// Two CTAs are resident per SM (6x4 node tile, 192 threads, ~101 KB shared memory) so that one
// CTA's copy/phase-1 overlaps the other's FP64-bound phase 2.
#include "box_tile.cuh"
#include "internal.h"

#include <cstdlib>

namespace nalu_b200 {

using box::BoxParams;

template <class C>
__global__ void __launch_bounds__(C::NT, (C::NT <= 256 ? 2 : 1)) assemble_box_kernel(const __grid_constant__ BoxParams P)
{
  extern __shared__ __align__(16) double sm[];
  double *const tab = sm + C::OFF_TAB;
  double *const ring = sm + C::OFF_RING;
  double *const mdb = sm + C::OFF_MD;
  long long *const s_rb = reinterpret_cast<long long *>(sm + C::OFF_RB);
  int *const s_len = reinterpret_cast<int *>(sm + C::OFF_LEN);
  constexpr int PLN = C::RN * C::NF;
  const int tid = threadIdx.x;
  const int ntile = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % ntile, zc = blockIdx.x / ntile;
  const int x0 = (tile % P.tiles_x) * C::TBX, y0 = (tile / P.tiles_x) * C::TBY;
  const int zs = zc * P.z_chunk;
  const int ze = (zs + P.z_chunk < P.NZ) ? zs + P.z_chunk : P.NZ;
  const bool divu = P.include_divu != 0.0;
  const long long plane_nodes = (long long)P.NX * P.NY;

  box::NodeCtx N;
  const int nl = tid >> 3;
  box::node_ctx_init<C>(P, N, tid, x0, y0);
  const int dz_fin = (N.cls & 4) ? 1 : 0;

  double regP[C::NPL], regM[C::NML];
  // ---- prologue: table, node planes zs-1..zs+1, the layer below the first plane, first node ctx
  for (int i = tid; i < 192; i += C::NT) box::build_table_entry<C>(i, P.options, tab);
  for (int pl = zs - 1; pl <= zs + 1; ++pl) {
    box::plane_load<C>(P, tid, x0, y0, pl, regP);
    box::plane_store<C>(tid, regP, ring + ((pl + 3) % 3) * PLN);
  }
  box::mdot_load<C>(P, tid, x0, y0, zs - 1, regM);
  box::mdot_store<C>(tid, regM, mdb);
  __syncthreads();
  if (zs >= 1)
    for (int it = tid; it < C::NITEM; it += C::NT)
      box::phase1_item<C>(P, it, x0, y0, ring + ((zs - 1) % 3) * PLN, ring + (zs % 3) * PLN, mdb, sm + ((zs - 1) & 1) * C::LBUF);
  __syncthreads();
  box::mdot_load<C>(P, tid, x0, y0, zs, regM);
  box::mdot_store<C>(tid, regM, mdb);
  N.n = N.x + (long long)P.NX * (N.y + (long long)P.NY * zs);
  N.row = 0; N.nb = 0; N.rowbase = 0; N.slots_mid = N.slots_fin = 0xFFFFFFFFu;
  if (N.active) {
    N.row = P.node_row[N.n];
    const long long b0 = P.browptr[N.row];
    N.nb = (int)(P.browptr[N.row + 1] - b0);
    N.rowbase = 9 * b0;
    N.slots_mid = box::load_slots(P.slot27 + N.n * 27, N.cls, -1);
    N.slots_fin = box::load_slots(P.slot27 + N.n * 27, N.cls, dz_fin);
  }
  __syncthreads();

  for (int z = zs; z < ze; ++z) {
    const bool more = z + 1 < ze;
    // loads for the next plane / layer, parked in registers
    box::plane_load<C>(P, tid, x0, y0, z + 2, regP);
    box::mdot_load<C>(P, tid, x0, y0, z + 1, regM);
    int row_n = 0;
    if (N.active && more) row_n = P.node_row[N.n + plane_nodes];
    const double *pl_m = ring + ((z + 2) % 3) * PLN; // plane z-1
    const double *pl_0 = ring + (z % 3) * PLN;
    const double *pl_p = ring + ((z + 1) % 3) * PLN;
    if (z < P.nz)
      for (int it = tid; it < C::NITEM; it += C::NT) box::phase1_item<C>(P, it, x0, y0, pl_0, pl_p, mdb, sm + (z & 1) * C::LBUF);
    __syncthreads();
    box::mdot_store<C>(tid, regM, mdb);
    const double *lower = sm + ((z + 1) & 1) * C::LBUF;
    const double *upper = sm + (z & 1) * C::LBUF;
    double *stg = sm + ((z + 1) & 1) * C::LBUF;
    if (N.cls == 0) {
      s_rb[nl] = N.rowbase;
      s_len[nl] = N.active ? N.nb * 9 : 0;
    }
    long long b0_n = 0, b1_n = 0;
    unsigned sm_n = 0xFFFFFFFFu, sf_n = 0xFFFFFFFFu;
    if (N.active && more) {
      b0_n = P.browptr[row_n];
      b1_n = P.browptr[row_n + 1];
      const unsigned char *sl = P.slot27 + (N.n + plane_nodes) * 27;
      sm_n = box::load_slots(sl, N.cls, -1);
      sf_n = box::load_slots(sl, N.cls, dz_fin);
    }
    box::P2State S;
    box::p2_begin(S);
    N.z = z;
    int ncov = 0;
    if (N.active) {
      ncov = ((N.x > 0) + 1 + (N.x < P.NX - 1)) * ((N.y > 0) + 1 + (N.y < P.NY - 1)) * ((z > 0) + 1 + (z < P.NZ - 1));
      box::p2_pass<C, 0>(P, N, S, lower, tab + 2 * N.cls, divu);
    }
    __syncthreads(); // every thread is done with the lower layer: its buffer is now the staging area
    if (N.active && N.nb != ncov) {
      // row also holds columns no local element touches (remote columns announced by the ranks
      // sharing this node): they must read zero
      double *row = stg + nl * C::SROW;
      for (int q = N.cls; q < N.nb * 9; q += 8) row[q] = 0.0;
    }
    __syncwarp();
    if (N.active) {
      if (N.cls & 4) {
        box::p2_flush<C>(N, S, N.slots_mid, pl_m, stg);
        box::p2_zero(S);
      }
      box::p2_pass<C, 1>(P, N, S, upper, tab + 2 * N.cls, divu);
      box::p2_flush<C>(N, S, N.slots_fin, dz_fin ? pl_p : pl_0, stg);
    }
    // rhs: sum the 8 class partials of the node (classes 0..2 also carry the divU stress components)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double v = S.rr[i] + (N.cls == i ? S.dv : 0.0);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      S.rr[i] = v;
    }
    if (N.active && N.cls < 3) {
      double *r = P.rhs + (long long)N.row * 3 + N.cls;
      const double v = N.cls == 0 ? S.rr[0] : (N.cls == 1 ? S.rr[1] : S.rr[2]);
      *r = P.overwrite ? v : *r + v;
    }
    __syncthreads();
    box::plane_store<C>(tid, regP, ring + ((z + 2) % 3) * PLN); // plane z-1 is no longer needed
    // finished rows -> HBM, one warp per node, 256 B per store instruction
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int m = warp; m < C::NNODE; m += C::NT / 32) {
        const int len = s_len[m];
        double *dst = P.values + s_rb[m];
        const double *src = stg + m * C::SROW;
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (lane + 32 * k < len) v[k] = src[lane + 32 * k];
        if (P.overwrite) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (lane + 32 * k < len) dst[lane + 32 * k] = v[k];
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (lane + 32 * k < len) dst[lane + 32 * k] += v[k];
        }
      }
    }
    __syncthreads();
    // rotate the node context
    N.n += plane_nodes;
    N.row = row_n;
    N.nb = (int)(b1_n - b0_n);
    N.rowbase = 9 * b0_n;
    N.slots_mid = sm_n;
    N.slots_fin = sf_n;
  }
}

bool box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (g.box[0] <= 0 || !g.d_node_slot27 || g.ndof != 3 || nk != 1) return false;
  if (k[0].kind != NALU_K_MOMENTUM_ADV_DIFF) return false;
  if (!k[0].nodal[NALU_F_COORDINATES] || !k[0].nodal[NALU_F_VELOCITY] || !k[0].nodal[NALU_F_VISCOSITY] ||
      !k[0].mass_flow_rate)
    return false;
  return true;
}

template <class C>
static int launch_cfg(const Graph &g, BoxParams &P, cudaStream_t stream)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  P.tiles_x = (P.NX + C::TBX - 1) / C::TBX;
  P.tiles_y = (P.NY + C::TBY - 1) / C::TBY;
  const long long ntile = (long long)P.tiles_x * P.tiles_y;
  // z chunks: enough CTAs for ~16 waves, but at least 16 planes per chunk (each chunk recomputes
  // one element layer)
  const int occ = C::NT <= 256 ? 2 : 1;
  long long want = ((long long)sms * occ * 16 + ntile - 1) / ntile;
  int zchunks = (int)(want < 1 ? 1 : want);
  int zc = (P.NZ + zchunks - 1) / zchunks;
  if (zc < 16) zc = 16;
  if (const char *e = getenv("NALU_BOX_ZCHUNK")) {
    const int v = atoi(e);
    if (v > 0) zc = v;
  }
  if (zc > P.NZ) zc = P.NZ;
  P.z_chunk = zc;
  zchunks = (P.NZ + zc - 1) / zc;
  const size_t smem = sizeof(double) * C::SMEM_DOUBLES;
  NALU_CUDA(cudaFuncSetAttribute(assemble_box_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_box: grid too large");
    return NALU_B200_ERR_ARG;
  }
  assemble_box_kernel<C><<<(unsigned)ctas, C::NT, smem, stream>>>(P);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
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
  P.values = d_values;
  P.rhs = d_rhs;
  P.err = g.d_err;
  P.include_divu = k[0].include_divu;
  P.options = k[0].options;
  P.overwrite = overwrite;
  P.trace = nullptr;
  P.z_first = 0;
  P.z_last = P.NZ;
  if (z_begin >= 0) {
    // a range of node planes: only the column-pair kernel takes one
    if (z_begin >= z_end || z_end > P.NZ) {
      set_error("assemble_box: bad node-plane range");
      return NALU_B200_ERR_ARG;
    }
    P.z_first = z_begin;
    P.z_last = z_end;
    if (const char *e = getenv("NALU_BOX_KERNEL"))
      if (atoi(e) == 4) return launch_assemble_box4(P, stream);
    return launch_assemble_box2(P, stream);
  }
  int tile = 0;
  if (const char *e = getenv("NALU_BOX_TILE")) tile = atoi(e);
  // default: the column-pair kernel (assemble_box2.cu); NALU_BOX_TILE selects the first-generation
  // class kernel of this file (kept for cross-checks and A/B measurements)
  if (tile == 0) {
    if (const char *e = getenv("NALU_BOX_KERNEL"))
      if (atoi(e) == 3) return launch_assemble_box3(P, stream);
    if (const char *e = getenv("NALU_BOX_KERNEL"))
      if (atoi(e) == 4) return launch_assemble_box4(P, stream);
    return launch_assemble_box2(P, stream);
  }
  if (tile == 84) return launch_cfg<box::Cfg<8, 4>>(g, P, stream);
  if (tile == 86) return launch_cfg<box::Cfg<8, 6>>(g, P, stream);
  return launch_cfg<box::Cfg<6, 4>>(g, P, stream);
}

} // namespace nalu_b200
