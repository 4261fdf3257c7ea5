// assemble_box2.cu -- launch side of the column-pair tiled assembly (see box2_tile.cuh).
//
// Grid: one CTA (128 threads = 4 class warps x 32 nodes, two column nodes per thread) per (8x4 node tile, z chunk), two CTAs per SM.
// Per node plane z a CTA runs
//   A  lower octants (face ips TF(z-1), z-edge ips ZE(z-1)): dz=-1 group -> S1, dz=0 partial -> S0
//   B  bulk copy of S1 to HBM  ||  phase 1a: BF(z), ZE(z)
//   C  upper octants: dz=0 completed in S0, dz=+1 group -> S1, rhs partials
//   D  bulk copy of S0, S1, rhs  ||  phase 1b: TF(z)  ||  metadata / field prefetch of the next plane
// with one CTA barrier between consecutive phases.
#include "box2_tile.cuh"
#include "internal.h"

#include <cstdint>

namespace nalu_b200 {

using box::BoxParams;
using C = box2::Cfg;

// Copy-out: one warp per run, lane k moves double k; the run descriptors of the plane are stored by build_run_desc
// at the head of phase A, so a run costs a broadcast descriptor load, one staging load and one predicated store,
// eight independent runs in flight per warp.  (Round-1 measurements, profiles/r2u_phase_trace.txt,
// r5b_phase_trace_*.txt: per-lane cp.async.bulk issue was slower -- a bulk copy occupies its warp for 58 to 250
// cycles whatever its size.)
template <int NB>
__device__ __forceinline__ void copy_runs_desc(const double *sm, int first, int n, int lane, double *values, int overwrite)
{
  const int4 *rd = reinterpret_cast<const int4 *>(sm + C::OFF_RD);
  const double *S = sm + C::OFF_S;
#pragma unroll 1
  for (int k0 = 0; k0 < n; k0 += NB) {
    int4 d[NB];
    double v[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) d[u] = rd[first + k0 + u];
#pragma unroll
    for (int u = 0; u < NB; ++u) v[u] = S[d[u].z + lane];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      double *dst = reinterpret_cast<double *>(((unsigned long long)(unsigned)d[u].y << 32) | (unsigned)d[u].x) + lane;
      if (lane < d[u].w) *dst = overwrite ? v[u] : *dst + v[u];
    }
  }
}

// phase 1 over one item list.  kind 0: prologue (top face + z edges of the layer below the first
// plane), 1: phase 1a (bottom face + z edges), 2: phase 1b (top face).
__device__ __forceinline__ void p1_phase(const BoxParams &P, int kind, int x0, int y0, const double *pl0, const double *pl1,
                                         const double *mdb, double *F, double *ZE)
{
  const int n = kind == 2 ? C::NP1B : C::NP1A;
  for (int it = threadIdx.x; it < n; it += C::NT) {
    const int k = it / C::NEL, el = it - k * C::NEL;
    const int g = k < 2 ? k : 2;
    const int side = k < 2 ? (kind == 1 ? 0 : 1) : k - 2;
    box2::p1_item(P, el, g, side, x0, y0, pl0, pl1, mdb, F, ZE);
  }
}
// out-of-line copy for the prologue (runs once per CTA; keeps the second copy of phase 1 out of the loop's
// register allocation)
__device__ __noinline__ void p1_prologue(const BoxParams &P, int x0, int y0, const double *pl0, const double *pl1,
                                         const double *mdb, double *F, double *ZE)
{
  p1_phase(P, 0, x0, y0, pl0, pl1, mdb, F, ZE);
}

__global__ void __launch_bounds__(C::NT, 2) assemble_box2_kernel(const __grid_constant__ BoxParams P)
{
  extern __shared__ __align__(16) double sm[];
  double *const F = sm + C::OFF_F, *const ZE = sm + C::OFF_ZE, *const S0 = sm + C::OFF_S, *const S1 = sm + C::OFF_S;
  double *const ring = sm + C::OFF_RING, *const mdb = sm + C::OFF_MD, *const rp = sm + C::OFF_RP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntile = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % ntile, zc = blockIdx.x / ntile;
  const int x0 = (tile % P.tiles_x) * C::TBX, y0 = (tile / P.tiles_x) * C::TBY;
  const int zs = P.z_first + zc * P.z_chunk;
  const int ze = (zs + P.z_chunk < P.z_last) ? zs + P.z_chunk : P.z_last;

  box2::T2 N;
  box2::t2_init(P, N, lane, x0, y0);
  // class constants live in shared memory (12 doubles per class warp) and are fetched at the head of
  // every phase 2 pass, so that they do not occupy registers during phase 1
  double *const qk = sm + C::OFF_QK + warp * 24;
  if (lane == 0) {
    box2::P2K K;
    box2::ClsK Q;
    box2::p2k_init(K, P.options, P.include_divu);
    box2::cls_init(Q, K, warp);
    box2::cls_store(Q, qk);
  }
  const double f23 = 2.0 / 3.0 * P.include_divu;

  // records of elements outside the domain are never written: they must read zero (phase 2 does not test)
  for (int i = tid; i < 2 * C::NEL * C::ES; i += C::NT) sm[C::OFF_F + i] = 0.0;
  // ---- prologue: node planes zs-1, zs; the face/z-edge ips of the layer below the first plane
  box2::stage_plane(P, ring + ((zs + 2) % 3) * C::PLN, tid, x0, y0, zs - 1);
  box2::stage_plane(P, ring + (zs % 3) * C::PLN, tid, x0, y0, zs);
  box2::stage_mdot(P, mdb, tid, x0, y0, zs - 1);
  if (warp == C::NCLS - 1) box2::meta_load(P, box2::meta_at(sm, zs & 1), lane, x0, y0, zs);
  box2::stage_wait();
  __syncthreads();
  if (zs >= 1) p1_prologue(P, x0, y0, ring + ((zs - 1) % 3) * C::PLN, ring + (zs % 3) * C::PLN, mdb, F, ZE);
  // the last warp (least phase 1 work) owns the node metadata; the row index is fetched one phase ahead
  const bool meta_warp = warp == C::NCLS - 1;
  int row_next = meta_warp ? box2::meta_row(P, lane, x0, y0, zs + 1) : -1;
  // first pass of the loop expects node plane zs+1 already in flight (see phase D below)
  box2::stage_plane(P, ring + ((zs + 1) % 3) * C::PLN, tid, x0, y0, zs + 1);
  __syncthreads();

  // One iteration = one z half of a plane (H = 0: elements below, H = 1: above) followed by the
  // work that overlaps it; every piece of code appears once (instruction-cache footprint).
  for (int hs = 2 * zs; hs < 2 * ze; ++hs) {
    const int z = hs >> 1, H = hs & 1;
    const box2::Meta M = box2::meta_at(sm, z & 1);
    const bool lower = z >= 1, upper = z < P.nz;
    // ---- A (H = 0) / C (H = 1)
    if (!H) {
      // the mass flow rates of layer z land while the lower octants are processed (node plane z+1 has
      // been in flight since phase D of the previous plane)
      box2::stage_mdot(P, mdb, tid, x0, y0, z);
      N.rr[0] = N.rr[1] = N.rr[2] = 0.0;
      box2::build_run_desc(M, sm, tid, P.values);
    }
    if (H ? upper : lower) {
      box2::ClsK Q;
      box2::cls_fetch(Q, qk, warp);
      box2::p2_half(N, sm, Q, f23, H, z, lower);
    }
    if (H) {
      rp[(warp * C::NNODE + lane) * 3 + 0] = N.rr[0];
      rp[(warp * C::NNODE + lane) * 3 + 1] = N.rr[1];
      rp[(warp * C::NNODE + lane) * 3 + 2] = N.rr[2];
    }
    if (!H) box2::stage_wait();
    __syncthreads();
    // ---- B (H = 0): dz=-1 group out, phase 1a / D (H = 1): dz=0,+1 groups and rhs out, phase 1b,
    //      metadata and field prefetch of the next plane
    box2::MetaRaw MR;
    const bool do_meta = !H && z + 1 < ze && meta_warp;
    if (do_meta) box2::meta_fetch(P, MR, lane, x0, y0, z + 1, row_next);
    if (H) {
      // node plane z+2 -> the ring slot of plane z-1 (dead since phase A); row indices of plane z+2
      box2::stage_plane(P, ring + ((z + 2) % 3) * C::PLN, tid, x0, y0, z + 2);
      if (meta_warp) row_next = box2::meta_row(P, lane, x0, y0, z + 2);
    }
    // H = 0: the 96 runs of dz=-1; H = 1: the 96 of dz=0 and the 96 of dz=+1; 24 / 48 per warp
    if (H) copy_runs_desc<8>(sm, 3 * C::NNODE + warp * (upper ? 48 : 24), upper ? 48 : 24, lane, P.values, P.overwrite);
    else if (lower) copy_runs_desc<8>(sm, warp * 24, 24, lane, P.values, P.overwrite);
    if (H && tid < 3 * C::NNODE && P.overwrite) box2::zero_uncovered(M, lane, warp, P.values);
    if (H && tid >= C::NT - 3 * C::NNODE) {
      // rhs: sum the class partials; thread -> (node, dof) so that an x line of nodes is one run
      const int t = tid - (C::NT - 3 * C::NNODE);
      const int node = t / 3, i = t - node * 3;
      if (M.cnt[node * 4 + 1]) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < C::NCLS; ++c) v += rp[(c * C::NNODE + node) * 3 + i];
        double *r = P.rhs + (long long)M.row[node] * 3 + i;
        *r = P.overwrite ? v : *r + v;
      }
    }
    if (upper) p1_phase(P, H ? 2 : 1, x0, y0, ring + (z % 3) * C::PLN, ring + ((z + 1) % 3) * C::PLN, mdb, F, ZE);
    if (do_meta) box2::meta_store(P, box2::meta_at(sm, (z + 1) & 1), MR, lane);
    __syncthreads();
  }
}

int launch_assemble_box2(BoxParams &P, cudaStream_t stream)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  P.tiles_x = (P.NX + C::TBX - 1) / C::TBX;
  P.tiles_y = (P.NY + C::TBY - 1) / C::TBY;
  const long long ntile = (long long)P.tiles_x * P.tiles_y;
  // z chunks: enough CTAs for ~12 waves of 2 CTAs per SM, but at least 16 planes per chunk (each chunk
  // recomputes one element layer in its prologue)
  long long want = ((long long)sms * 2 * 12 + ntile - 1) / ntile;
  int zchunks = (int)(want < 1 ? 1 : want);
  const int nplanes = P.z_last - P.z_first;
  int zc = (nplanes + zchunks - 1) / zchunks;
  if (zc < 16) zc = 16;
  if (tuning().box_zchunk > 0) zc = tuning().box_zchunk;
  if (zc > nplanes) zc = nplanes;
  P.z_chunk = zc;
  zchunks = (nplanes + zc - 1) / zc;
  const size_t smem = sizeof(double) * C::SMEM_DOUBLES;
  NALU_CUDA(cudaFuncSetAttribute(assemble_box2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NALU_CUDA(cudaFuncSetAttribute(assemble_box2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_box2: grid too large");
    return NALU_B200_ERR_ARG;
  }
  assemble_box2_kernel<<<(unsigned)ctas, C::NT, smem, stream>>>(P);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

} // namespace nalu_b200
