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
#include <cstdio>
#include <cstdlib>

namespace nalu_b200 {

using box::BoxParams;
using C = box2::Cfg;

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(double *dst, const double *src, int bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_add(double *dst, const double *src, int bytes)
{
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst),
               "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// dof row i of dz group g of one node: staging -> values.  The staging offset carries the parity of
// the destination, so the 16-byte aligned middle goes out as one bulk copy and at most one double
// at either end as a plain store.
template <bool BULK>
__device__ __forceinline__ void copy_run(const box2::Meta &M, int node, int i, int g, const double *S, double *values,
                                         int overwrite)
{
  const box2::Run r = box2::run_of(M, node, i, g);
  if (r.L == 0) return;
  const double *src = S + r.soff;
  double *dst = values + r.G;
  if (BULK) {
    int L = r.L;
    if (r.G & 1) {
      *dst = overwrite ? *src : *dst + *src;
      ++dst; ++src; --L;
    }
    if (L & 1) {
      --L;
      dst[L] = overwrite ? src[L] : dst[L] + src[L];
    }
    if (L) {
      if (overwrite) bulk_store(dst, src, L * 8);
      else bulk_add(dst, src, L * 8);
    }
  } else {
    for (int k = 0; k < r.L; ++k) dst[k] = overwrite ? src[k] : dst[k] + src[k];
  }
}

// Cooperative variant: one warp per run, lane k moves double k (27 lanes busy, one 216-byte run per
// instruction pair).  Runs [r0, r0 + n) of the phase's run list; run r -> group g0 + r / 96, dof row
// (r % 96) / 32, node r % 32.
__device__ __forceinline__ void copy_runs_warp(const box2::Meta &M, int warp, int lane, int nk, int g0, const double *S0,
                                               const double *S1, double *values, int overwrite)
{
  // warp w takes runs w, w + 8, ...; batches of four: all descriptors, then all staging loads, then
  // all stores, so that the descriptor -> data -> store chains of the four runs overlap
  for (int k0 = 0; k0 < nk; k0 += 4) {
    box2::Run run[4];
    const double *src[4];
    double v[4], w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = warp + C::NCLS * (k0 + u);
      const int g = g0 + r / 96, rr = r % 96;
      run[u] = box2::run_of(M, rr & 31, rr >> 5, g);
      src[u] = (g == 1 ? S0 : S1) + run[u].soff;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = lane < run[u].L ? src[u][lane] : 0.0;
      w[u] = lane + 32 < run[u].L ? src[u][lane + 32] : 0.0; // merged dz=0 / dz=+1 runs are 54 long
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double *dst = values + run[u].G + lane;
      if (lane < run[u].L) *dst = overwrite ? v[u] : *dst + v[u];
      if (lane + 32 < run[u].L) dst[32] = overwrite ? w[u] : dst[32] + w[u];
    }
  }
}

// Descriptor variant (the default): one warp per run, lane k moves double k; the run descriptors of the plane
// were stored by build_run_desc at the head of phase A, so a run costs a broadcast descriptor load, one staging
// load and one predicated store, eight independent runs in flight per warp.  No bulk-copy issue: a
// cp.async.bulk occupies its warp for 58 (idle SM) to 250 (inside this kernel) cycles whatever its size, and a
// plane has 192 of them (profiles/r2u_phase_trace.txt: 9.8 k of a plane's 31 k cycles).
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

// profiling aid: warp w of CTA `TRACE_CTA` stamps the clock at phase boundaries
#define TRACE_CTA 1500
#define TRACE(slot)                                                                                   \
  do {                                                                                                \
    if (P.trace && blockIdx.x == TRACE_CTA && lane == 0 && hs < 2 * zs + 16)                           \
      P.trace[((hs - 2 * zs) * 8 + warp) * 8 + (slot)] = clock64();                                    \
  } while (0)

// MODE 0: cooperative copy-out with in-loop run_of, 1: cp.async.bulk, 2: descriptor copy-out
template <int MODE>
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
  constexpr bool BULK = MODE == 1;

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
    TRACE(0);
    if (!H) {
      // the mass flow rates of layer z land while the lower octants are processed (node plane z+1 has
      // been in flight since phase D of the previous plane)
      box2::stage_mdot(P, mdb, tid, x0, y0, z);
      N.rr[0] = N.rr[1] = N.rr[2] = 0.0;
      if (MODE == 2) box2::build_run_desc(M, sm, tid, P.values);
    }
    if ((H ? upper : lower) && !(P.options & (1 << 18))) {
      // profiling aid: 9 stamps inside the pass for half-steps 4 and 5 of the traced CTA
      N.trace = (P.trace && blockIdx.x == TRACE_CTA && lane == 0 && (hs == 2 * zs + 4 || hs == 2 * zs + 5))
                    ? P.trace + 16 * 8 * 8 + ((hs - 2 * zs - 4) * 4 + warp) * 16
                    : nullptr;
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
    if (BULK) fence_async_smem();
    TRACE(1);
    __syncthreads();
    TRACE(2);
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
    if (P.options & (1 << 16)) {
    } else if (MODE == 2) {
      // H = 0: the 96 runs of dz=-1; H = 1: the 96 of dz=0 and the 96 of dz=+1; 24 / 48 per warp
      if (H) copy_runs_desc<8>(sm, 3 * C::NNODE + warp * (upper ? 48 : 24), upper ? 48 : 24, lane, P.values, P.overwrite);
      // (an uneven 0 / 8 / 44 / 44 split against the phase 1a load measured 1 % slower: the stores drain at a
      // rate set by the memory system, not by the issuing warp)
      else if (lower) copy_runs_desc<8>(sm, warp * 24, 24, lane, P.values, P.overwrite);
    } else if (BULK) {
      // the run list of this half (H = 0: 96 runs of the dz=-1 group; H = 1: 96 of dz=0, then 96 of
      // dz=+1) spread over all 8 warps: the per-lane bulk issue serialises inside a warp
      if (H) {
        // every warp takes 24 runs of dz=0 (which carry dz=+1 along when merged) and 24 of dz=+1
#pragma unroll 1
        for (int l = lane; l < 48; l += 32) {
          const int g = 1 + l / 24, rr = warp * 24 + l % 24;
          if (g == 2 ? upper : true) copy_run<BULK>(M, rr & 31, rr >> 5, g, S0, P.values, P.overwrite);
        }
      } else if (lower) {
        // the 96 runs of dz=-1, shared out against the phase 1a load (180 items on 128 threads: warp 0
        // has two items per lane, warp 1 two on 20 lanes): 0 / 8 / 44 / 44 runs
        const int first = warp < 2 ? 0 : 8 + (warp - 2) * 44, cnt = warp == 0 ? 0 : (warp == 1 ? 8 : 44);
#pragma unroll 1
        for (int l = lane; l < cnt; l += 32) {
          const int rr = first + l;
          copy_run<BULK>(M, rr & 31, rr >> 5, 0, S0, P.values, P.overwrite);
        }
      }
      bulk_commit();
    } else {
      // run list of this half: H = 0: the 96 runs of the dz=-1 group; H = 1: 96 runs of dz=0, then 96 of dz=+1
      const int nrun = H ? (upper ? 192 : 96) : (lower ? 96 : 0);
      copy_runs_warp(M, warp, lane, nrun / C::NCLS, H ? 1 : 0, S0, S1, P.values, P.overwrite);
    }
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
    TRACE(3);
    if (upper && !(P.options & (1 << 17))) p1_phase(P, H ? 2 : 1, x0, y0, ring + (z % 3) * C::PLN, ring + ((z + 1) % 3) * C::PLN, mdb, F, ZE);
    if (do_meta) box2::meta_store(P, box2::meta_at(sm, (z + 1) & 1), MR, lane);
    TRACE(4);
    if (BULK) bulk_wait_read();
    TRACE(5);
    __syncthreads();
    TRACE(6);
  }
}

template <int MODE>
static int launch2(BoxParams &P, cudaStream_t stream)
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
  if (const char *e = getenv("NALU_BOX_ZCHUNK")) {
    const int v = atoi(e);
    if (v > 0) zc = v;
  }
  if (zc > nplanes) zc = nplanes;
  P.z_chunk = zc;
  zchunks = (nplanes + zc - 1) / zc;
  size_t smem = sizeof(double) * C::SMEM_DOUBLES;
  if (getenv("NALU_BOX_ONE_CTA")) smem = 120 * 1024; // profiling aid: one CTA per SM
  NALU_CUDA(cudaFuncSetAttribute(assemble_box2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NALU_CUDA(cudaFuncSetAttribute(assemble_box2_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_box2: grid too large");
    return NALU_B200_ERR_ARG;
  }
  static long long *d_trace = nullptr;
  const bool trace = getenv("NALU_BOX_TRACE") != nullptr;
  if (trace && !d_trace) NALU_CUDA(cudaMalloc(&d_trace, (16 * 8 * 8 + 128) * sizeof(long long)));
  P.trace = trace ? d_trace : nullptr;
  if (trace) NALU_CUDA(cudaMemsetAsync(d_trace, 0, (16 * 8 * 8 + 128) * sizeof(long long), stream));
  assemble_box2_kernel<MODE><<<(unsigned)ctas, C::NT, smem, stream>>>(P);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  if (trace) {
    static long long h[16 * 8 * 8 + 128];
    NALU_CUDA(cudaStreamSynchronize(stream));
    NALU_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
    bool any = false; // launches without the traced CTA (small z ranges) leave the file alone
    for (int k = 0; k < 16 * 8 * 8 && !any; ++k) any = h[k] != 0;
    if (FILE *f = any ? fopen(getenv("NALU_BOX_TRACE"), "w") : nullptr) {
      for (int hs = 0; hs < 16; ++hs)
        for (int w = 0; w < 8; ++w) {
          fprintf(f, "%d %d", hs, w);
          for (int k = 0; k < 7; ++k) fprintf(f, " %lld", h[(hs * 8 + w) * 8 + k] ? h[(hs * 8 + w) * 8 + k] - h[0] : -1LL);
          fprintf(f, "\n");
        }
      for (int k = 0; k < 8; ++k) {
        fprintf(f, "p2 half %d warp %d:", k / 4, k % 4);
        const long long *t = h + 16 * 8 * 8 + k * 16;
        for (int j = 1; j < 9; ++j) fprintf(f, " %lld", t[j] && t[j - 1] ? t[j] - t[j - 1] : -1LL);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return NALU_B200_OK;
}

int launch_assemble_box2(BoxParams &P, cudaStream_t stream)
{
  // the bulk copies need 16-byte aligned destinations: staging parity is matched to the offset in
  // `values`, so the base pointer itself must be aligned (cudaMalloc gives 256 bytes)
  if (const char *e = getenv("NALU_BOX_ABLATE")) P.options |= atoi(e) << 16; // profiling aid: skip copy (1) / phase 1 (2) / phase 2 (4)
  // copy-out: 2 = run descriptors + plain stores (default), 1 = cp.async.bulk (needs a 16-byte aligned
  // `values`), 0 = first cooperative variant; NALU_BOX_BULK selects for A/B runs
  int mode = 2;
  if (const char *e = getenv("NALU_BOX_BULK")) mode = atoi(e);
  if (mode == 1 && ((uintptr_t)P.values & 15) != 0) mode = 2;
  return mode == 1 ? launch2<1>(P, stream) : (mode == 0 ? launch2<0>(P, stream) : launch2<2>(P, stream));
}

} // namespace nalu_b200
