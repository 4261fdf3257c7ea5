// assemble_box4.cu -- launch side of the row-thread tiled momentum assembly (see box4_tile.cuh).
//
// Grid and phase structure of assemble_box2.cu (one CTA of 128 threads per (8x4 node tile, z chunk), two per SM);
// per node plane z
//   A1  lower octants into registers (rows: warps 0-2, advection operator: warp 3); exchange put
//   A2  exchange get: diagonal term, rhs share; dz=-1 group -> CSR, dz=0 partial parked  ||  phase 1a: BF(z), ZE(z)
//   C1  upper octants; exchange put
//   C2  exchange get; + parked partial; dz=0 and dz=+1 groups and the rhs -> CSR  ||  phase 1b: TF(z)  ||  metadata / field prefetch
// with one CTA barrier between consecutive phases.  Selected with NALU_BOX_KERNEL=4.
#include "box4_tile.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace nalu_b200 {

using box::BoxParams;
using C = box2::Cfg;

__device__ __forceinline__ void p1_phase4(const BoxParams &P, int kind, int x0, int y0, const double *pl0, const double *pl1,
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
__device__ __noinline__ void p1_prologue4(const BoxParams &P, int x0, int y0, const double *pl0, const double *pl1, const double *mdb,
                                          double *F, double *ZE)
{
  p1_phase4(P, 0, x0, y0, pl0, pl1, mdb, F, ZE);
}

// the 32 staged runs of the warp's slice -> CSR: one run (<= 27 contiguous doubles) per warp instruction
__device__ __forceinline__ void copy_staged_warp(const BoxParams &P, double *sm, const box4::Row &N, const box4::RunW &run, int lane)
{
  const double *S = box4::warp_slice(sm, N);
  __syncwarp();
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const long long G = __shfl_sync(0xffffffffu, run.G, r);
    const int L = __shfl_sync(0xffffffffu, run.L, r);
    const double v = S[r * 27 + lane];
    if (lane < L) {
      double *d = P.values + G + lane;
      *d = P.overwrite ? v : *d + v;
    }
  }
  __syncwarp();
}

// profiling aid (NALU_BOX_TRACE): warp w of CTA 1500 stamps the clock at phase boundaries, format of assemble_box2.cu
#define TRACE4(hs, slot)                                                                      \
  do {                                                                                        \
    if (P.trace && blockIdx.x == 1500 && lane == 0 && (hs) < 2 * zs + 16)                     \
      P.trace[(((hs) - 2 * zs) * 8 + warp) * 8 + (slot)] = clock64();                         \
  } while (0)

__global__ void __launch_bounds__(C::NT, 2) assemble_box4_kernel(const __grid_constant__ BoxParams P)
{
  extern __shared__ __align__(16) double sm[];
  double *const F = sm + C::OFF_F, *const ZE = sm + C::OFF_ZE;
  double *const ring = sm + C::OFF_RING, *const mdb = sm + C::OFF_MD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntile = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % ntile, zc = blockIdx.x / ntile;
  const int x0 = (tile % P.tiles_x) * C::TBX, y0 = (tile / P.tiles_x) * C::TBY;
  const int zs = P.z_first + zc * P.z_chunk;
  const int ze = (zs + P.z_chunk < P.z_last) ? zs + P.z_chunk : P.z_last;

  box4::Row N;
  N.node = lane;
  N.lx = lane % C::TBX;
  N.ly = lane / C::TBX;
  N.i = warp;
  box2::P2K Kc;
  box2::p2k_init(Kc, P.options, P.include_divu);
  const bool rowthread = warp < 3;
  const bool direct = (P.options & (1 << 20)) == 0; // NALU_BOX_ABLATE=16: warp-staged coalesced output (measured slower)

  for (int i = tid; i < 2 * C::NEL * C::ES; i += C::NT) sm[C::OFF_F + i] = 0.0;
  box2::stage_plane(P, ring + ((zs + 2) % 3) * C::PLN, tid, x0, y0, zs - 1);
  box2::stage_plane(P, ring + (zs % 3) * C::PLN, tid, x0, y0, zs);
  box2::stage_mdot(P, mdb, tid, x0, y0, zs - 1);
  if (warp == 3) box2::meta_load(P, box2::meta_at(sm, zs & 1), lane, x0, y0, zs);
  box2::stage_wait();
  __syncthreads();
  if (zs >= 1) p1_prologue4(P, x0, y0, ring + ((zs - 1) % 3) * C::PLN, ring + (zs % 3) * C::PLN, mdb, F, ZE);
  int row_next = warp == 3 ? box2::meta_row(P, lane, x0, y0, zs + 1) : -1;
  box2::stage_plane(P, ring + ((zs + 1) % 3) * C::PLN, tid, x0, y0, zs + 1);
  __syncthreads();

  for (int z = zs; z < ze; ++z) {
    const box2::Meta M = box2::meta_at(sm, z & 1);
    const bool lower = z >= 1, upper = z < P.nz;
    // ---- A1: lower octants
    TRACE4(2 * z, 0);
    box2::stage_mdot(P, mdb, tid, x0, y0, z);
    double acc[18][3], rr = 0.0;
    if (lower) {
      if (rowthread) {
#pragma unroll
        for (int c = 0; c < 18; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
        box4::half_rows<0>(sm, N, Kc, acc);
        box4::exchange_put_row<0>(sm, N, acc, z);
      } else {
        double dm[18];
#pragma unroll
        for (int c = 0; c < 18; ++c) dm[c] = 0.0;
        box4::half_adv<0>(sm, N, Kc, dm);
        box4::exchange_put_adv(sm, N, dm);
      }
    }
    box2::stage_wait();
    TRACE4(2 * z, 1);
    __syncthreads();
    TRACE4(2 * z, 2);
    // ---- A2 + B: finalize the lower half, dz=-1 group out, dz=0 partial parked; phase 1a; metadata of plane z+1
    box2::MetaRaw MR;
    const bool do_meta = z + 1 < ze && warp == 3;
    if (do_meta) box2::meta_fetch(P, MR, lane, x0, y0, z + 1, row_next);
    if (rowthread) {
      if (lower) {
        box4::finalize_row<0>(sm, N, acc, rr, Kc.f23, z);
        if (direct) box4::store_group(P, M, N, 0, 9, acc);
        else copy_staged_warp(P, sm, N, box4::stage_group(sm, M, N, 0, 9, acc), lane);
      } else {
#pragma unroll
        for (int c = 0; c < 9; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
      }
      box4::park(sm, N, acc, rr);
    }
    TRACE4(2 * z, 3);
    if (upper) p1_phase4(P, 1, x0, y0, ring + (z % 3) * C::PLN, ring + ((z + 1) % 3) * C::PLN, mdb, F, ZE);
    if (do_meta) box2::meta_store(P, box2::meta_at(sm, (z + 1) & 1), MR, lane);
    TRACE4(2 * z, 4);
    TRACE4(2 * z, 5);
    __syncthreads();
    TRACE4(2 * z, 6);
    TRACE4(2 * z + 1, 0);
    // ---- C1: upper octants
    rr = 0.0;
    if (rowthread) {
#pragma unroll
      for (int c = 0; c < 18; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
      if (upper) {
        box4::half_rows<1>(sm, N, Kc, acc);
        box4::exchange_put_row<1>(sm, N, acc, z);
      }
    } else if (upper) {
      double dm[18];
#pragma unroll
      for (int c = 0; c < 18; ++c) dm[c] = 0.0;
      box4::half_adv<1>(sm, N, Kc, dm);
      box4::exchange_put_adv(sm, N, dm);
    }
    TRACE4(2 * z + 1, 1);
    __syncthreads();
    TRACE4(2 * z + 1, 2);
    // ---- C2 + D: finalize, dz=0 / dz=+1 groups and rhs out; phase 1b; prefetch of plane z+2
    box2::stage_plane(P, ring + ((z + 2) % 3) * C::PLN, tid, x0, y0, z + 2);
    if (warp == 3) row_next = box2::meta_row(P, lane, x0, y0, z + 2);
    if (rowthread) {
      // the exchange carries only the upper half's share; the lower half's (diagonal term included) waits in the park
      if (upper) box4::finalize_row<1>(sm, N, acc, rr, Kc.f23, z);
      box4::unpark_add(sm, N, acc, rr);
      if (direct) {
        box4::store_group(P, M, N, 1, 0, acc);
        if (upper) box4::store_group(P, M, N, 2, 9, acc);
      } else {
        __syncwarp(); // every lane has read its parked share before the slice becomes the staging area
        copy_staged_warp(P, sm, N, box4::stage_group(sm, M, N, 1, 0, acc), lane);
        if (upper) copy_staged_warp(P, sm, N, box4::stage_group(sm, M, N, 2, 9, acc), lane);
      }
      if (P.overwrite) box2::zero_uncovered(M, lane, warp, P.values);
      if (M.cnt[lane * 4 + 1]) {
        double *r = P.rhs + (long long)M.row[lane] * 3 + warp;
        *r = P.overwrite ? rr : *r + rr;
      }
    }
    TRACE4(2 * z + 1, 3);
    if (upper) p1_phase4(P, 2, x0, y0, ring + (z % 3) * C::PLN, ring + ((z + 1) % 3) * C::PLN, mdb, F, ZE);
    TRACE4(2 * z + 1, 4);
    TRACE4(2 * z + 1, 5);
    __syncthreads();
    TRACE4(2 * z + 1, 6);
  }
}

int launch_assemble_box4(BoxParams &P, cudaStream_t stream)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  P.tiles_x = (P.NX + C::TBX - 1) / C::TBX;
  P.tiles_y = (P.NY + C::TBY - 1) / C::TBY;
  const long long ntile = (long long)P.tiles_x * P.tiles_y;
  long long want = ((long long)sms * 2 * 12 + ntile - 1) / ntile;
  int zchunks = (int)(want < 1 ? 1 : want);
  const int nplanes = P.z_last - P.z_first;
  int zc = (nplanes + zchunks - 1) / zchunks;
  if (zc < 16) zc = 16;
  if (const char *e = getenv("NALU_BOX_ZCHUNK")) {
    const int v = atoi(e);
    if (v > 0) zc = v;
  }
  if (zc > nplanes) zc = nplanes;
  P.z_chunk = zc;
  zchunks = (nplanes + zc - 1) / zc;
  const size_t smem = sizeof(double) * C::SMEM_DOUBLES;
  NALU_CUDA(cudaFuncSetAttribute(assemble_box4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NALU_CUDA(cudaFuncSetAttribute(assemble_box4_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_box4: grid too large");
    return NALU_B200_ERR_ARG;
  }
  if (const char *e = getenv("NALU_BOX_ABLATE")) P.options |= atoi(e) << 16;
  static long long *d_trace = nullptr;
  const bool trace = getenv("NALU_BOX_TRACE") != nullptr;
  constexpr int TN = 16 * 8 * 8;
  if (trace && !d_trace) NALU_CUDA(cudaMalloc(&d_trace, TN * sizeof(long long)));
  P.trace = trace ? d_trace : nullptr;
  if (trace) NALU_CUDA(cudaMemsetAsync(d_trace, 0, TN * sizeof(long long), stream));
  assemble_box4_kernel<<<(unsigned)ctas, C::NT, smem, stream>>>(P);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  if (trace) {
    static long long h[TN];
    NALU_CUDA(cudaStreamSynchronize(stream));
    NALU_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
    bool any = false;
    for (int k = 0; k < TN && !any; ++k) any = h[k] != 0;
    if (FILE *f = any ? fopen(getenv("NALU_BOX_TRACE"), "w") : nullptr) {
      for (int hs = 0; hs < 16; ++hs)
        for (int w = 0; w < 8; ++w) {
          fprintf(f, "%d %d", hs, w);
          for (int k = 0; k < 7; ++k) fprintf(f, " %lld", h[(hs * 8 + w) * 8 + k] ? h[(hs * 8 + w) * 8 + k] - h[0] : -1LL);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
  return NALU_B200_OK;
}

} // namespace nalu_b200
