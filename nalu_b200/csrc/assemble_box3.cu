// assemble_box3.cu -- warp-specialised column-pair assembly (see box3_tile.cuh): one 384-thread CTA per
// SM; warps 0-3 consume (phase 2, one class each), warps 4-7 produce (phase 1), warps 8-10 stream the
// staging rows to HBM (cp.async.bulk) and write the rhs, warp 11 stages node planes / mass flow rates /
// node metadata one plane ahead.  The roles meet at mbarriers that complete once per node plane:
//   loaded(z)   plane z+1, mass flow rates of layer z, metadata of plane z are in shared memory
//   fullA(z)    bottom-face + z-edge records of layer z written      (producers -> upper half of plane z)
//   fullB(z)    top-face records of layer z written                  (producers -> lower half of plane z+1)
//   doneL(z)    lower half of plane z finished: F_TF / ZE[(z-1)&1] free, SB holds the dz=-1 group
//   doneU(z)    upper half finished: F_BF free, SA holds dz=0 | dz=+1, rhs partials written
//   emptyB(z)   SB read by the bulk copies;   emptyA(z)   SA and the rhs partials read
// Selected with NALU_BOX_KERNEL=3.  EXPERIMENTAL: parity-green (tests/test_gpu_parity.py::test_box3_*), but at
// 408 Melem/s (256^3) slower than assemble_box2 (604): with four different instruction streams resident the
// straight-line phase-1 / phase-2 bodies no longer fit the instruction cache (56-60 % of the stall samples
// inside the FMA code are no-instruction stalls, profiles/r3c_box3_128_summary.txt), and a single
// consumer group caps the SM at 32 nodes in flight.  Kept as the scaffold for a compact-loop version.
#include "box3_tile.cuh"
#include "internal.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace nalu_b200 {

using box::BoxParams;
using C3 = box3::Cfg3;
using B2 = box2::Cfg;

namespace {
__device__ __forceinline__ unsigned sa(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sa(b)) : "memory");
}
__device__ long long g_wait_acc[12 * 8];
__device__ __forceinline__ void mbar_wait_raw(unsigned long long *b, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\n"
               "bra WAIT_%=;\n"
               "DONE_%=:\n"
               "}" ::"r"(sa(b)), "r"(parity)
               : "memory");
}
#ifndef NALU_B3_PROFILE
#define NALU_B3_PROFILE 0 // build with -DNALU_B3_PROFILE=1 for the per-role wait accounting below
#endif
#if NALU_B3_PROFILE
// profiling aid: per-warp cycles spent waiting on each barrier (CTA 100), dumped with NALU_BOX_TRACE
#define mbar_wait(b, par)                                                              \
  do {                                                                                 \
    const long long _t0 = clock64();                                                   \
    mbar_wait_raw(b, par);                                                             \
    if (blockIdx.x == 100 && lane == 0) g_wait_acc[warp * 8 + (int)((b)-bar)] += clock64() - _t0; \
  } while (0)
#else
#define mbar_wait(b, par) mbar_wait_raw(b, par)
#endif
__device__ __forceinline__ void fence_async_smem3() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store3(double *dst, const double *src, int bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(sa(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_add3(double *dst, const double *src, int bytes)
{
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(sa(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_wait3()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void copy_run3(const box2::Run &r, const double *sm, double *values, int overwrite, bool bulk)
{
  if (r.L == 0) return;
  const double *src = sm + r.soff;
  double *dst = values + r.G;
  if (bulk) {
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
      if (overwrite) bulk_store3(dst, src, L * 8);
      else bulk_add3(dst, src, L * 8);
    }
  } else {
    for (int k = 0; k < r.L; ++k) dst[k] = overwrite ? src[k] : dst[k] + src[k];
  }
}

// loader item of plane z: node plane z+1, mass flow rates of layer z, metadata of plane z (32 lanes)
__device__ __forceinline__ void load_item(const BoxParams &P, double *sm, int lane, int x0, int y0, int z)
{
  // stage_plane / stage_mdot of box2 with this kernel's buffers and a 32-thread stride
  const long long pn = (long long)P.NX * P.NY;
  {
    double *ring_plane = sm + C3::OFF_RING + ((z + 1) & 3) * C3::PLN;
    const int zp = z + 1;
    const bool zin = zp >= 0 && zp < P.NZ;
#pragma unroll 1
    for (int idx = lane; idx < B2::PLN; idx += 32) {
      const int rn = idx / B2::NF, f = idx - rn * B2::NF;
      const int x = x0 - 1 + rn % B2::RNX, y = y0 - 1 + rn / B2::RNX;
      const bool valid = zin && x >= 0 && y >= 0 && x < P.NX && y < P.NY;
      const long long n = valid ? x + (long long)P.NX * y + pn * zp : 0;
      const double *src = f < 3 ? P.coords + n * 3 + f : (f == 3 ? P.visc + n : P.vel + n * 3 + (f - 4));
      box2::copy8_or_zero(ring_plane + idx, src, valid);
    }
  }
  {
    double *mdb = sm + C3::OFF_MD + (z & 1) * C3::MDB;
    const long long ln = (long long)P.nx * P.ny;
    const bool zin = z >= 0 && z < P.nz;
#pragma unroll 1
    for (int idx = lane; idx < B2::MDB; idx += 32) {
      const int el = idx / 12, ip = idx - el * 12;
      const int ex = x0 - 1 + el % B2::ELX, ey = y0 - 1 + el / B2::ELX;
      const bool valid = zin && ex >= 0 && ey >= 0 && ex < P.nx && ey < P.ny;
      const long long e = valid ? ex + (long long)P.nx * ey + ln * z : 0;
      box2::copy8_or_zero(mdb + idx, P.mdot + e * 12 + ip, valid);
    }
  }
  box2::MetaRaw R;
  box2::meta_fetch(P, R, lane, x0, y0, z, box2::meta_row(P, lane, x0, y0, z));
  box3::meta_store3(P, box3::meta3_at(sm, z), R, lane);
  box2::stage_wait();
}
} // namespace

template <bool BULK>
__global__ void __launch_bounds__(C3::NT, 1) assemble_box3_kernel(const __grid_constant__ BoxParams P)
{
  extern __shared__ __align__(16) double sm[];
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + C3::OFF_BAR);
  unsigned long long *const b_loaded = bar + 0, *const b_fullA = bar + 1, *const b_fullB = bar + 2, *const b_doneL = bar + 3,
                           *const b_doneU = bar + 4, *const b_emptyA = bar + 5, *const b_emptyB = bar + 6;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntile = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % ntile, zc = blockIdx.x / ntile;
  const int x0 = (tile % P.tiles_x) * B2::TBX, y0 = (tile / P.tiles_x) * B2::TBY;
  const int zs = zc * P.z_chunk;
  const int ze = (zs + P.z_chunk < P.NZ) ? zs + P.z_chunk : P.NZ;
  const int nk = ze - zs;

  // ---- prologue (all warps): barriers, zero records, class constants, planes zs-1..zs+1, layers zs-1, zs
  if (tid == 0) {
    mbar_init(b_loaded, 1);
    mbar_init(b_fullA, C3::NPROD);
    mbar_init(b_fullB, C3::NPROD);
    mbar_init(b_doneL, C3::NCONS);
    mbar_init(b_doneU, C3::NCONS);
    mbar_init(b_emptyA, 96);
    mbar_init(b_emptyB, 96);
  }
  for (int i = tid; i < 4 * C3::NEL * C3::ES; i += C3::NT) sm[C3::OFF_BF + i] = 0.0;
  if (warp < 4 && lane == 0) {
    box2::P2K K;
    box2::ClsK Q;
    box2::p2k_init(K, P.options, P.include_divu);
    box2::cls_init(Q, K, warp);
    box2::cls_store(Q, sm + C3::OFF_QK + warp * 24);
  }
  if (warp == 11) {
    // plane zs-1 and zs (items zs-2, zs-1: only their node plane matters), then the full item of plane zs
    load_item(P, sm, lane, x0, y0, zs - 2);
    load_item(P, sm, lane, x0, y0, zs - 1);
    load_item(P, sm, lane, x0, y0, zs);
  }
  __syncthreads();
  // records the lower half of the first plane needs: top face + z edges of layer zs-1
  if (zs >= 1) box3::p1_step3(P, sm, 0, x0, y0, zs - 1, tid, C3::NT);
  __syncthreads();

  const double f23 = 2.0 / 3.0 * P.include_divu;
  const long long t_role0 = clock64();
  if (warp < 4) {
    // ================= consumers: phase 2, class = warp =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    box2::T2 N;
    box2::t2_init(P, N, lane, x0, y0);
    const double *qk = sm + C3::OFF_QK + warp * 24;
    double *rp = sm + C3::OFF_RP;
    for (int hs = 0; hs < 2 * nk; ++hs) {
      const int k = hs >> 1, H = hs & 1, z = zs + k;
      const bool lower = z >= 1, upper = z < P.nz;
      if (!H) {
        if (k > 0) {
          mbar_wait(b_fullB, (k - 1) & 1);  // top-face records of layer z-1
          mbar_wait(b_emptyB, (k - 1) & 1); // SB read
        }
        mbar_wait(b_loaded, k & 1);
        if (!upper && k > 0) mbar_wait(b_emptyA, (k - 1) & 1); // top plane: the lower half writes dz=0 straight to SA
        N.rr[0] = N.rr[1] = N.rr[2] = 0.0;
      } else {
        mbar_wait(b_fullA, k & 1);                     // bottom-face + z-edge records of layer z
        if (k > 0) mbar_wait(b_emptyA, (k - 1) & 1);   // SA and the rhs partials read
      }
      if (H ? upper : lower) {
        box2::ClsK Q;
        box2::cls_fetch(Q, qk, warp);
        box3::p2_half3(N, sm, Q, f23, H, z, lower, upper);
      }
      if (H) {
        rp[(warp * B2::NNODE + lane) * 3 + 0] = N.rr[0];
        rp[(warp * B2::NNODE + lane) * 3 + 1] = N.rr[1];
        rp[(warp * B2::NNODE + lane) * 3 + 2] = N.rr[2];
      }
      if (BULK) fence_async_smem3();
      mbar_arrive(H ? b_doneU : b_doneL);
    }
  } else if (warp < 8) {
    // ================= producers: phase 1 =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 152;");
    const int t = tid - C3::NCONS;
    for (int hs = 0; hs < 2 * nk; ++hs) {
      const int k = hs >> 1, H = hs & 1, z = zs + k;
      const bool upper = z < P.nz;
      if (!H) {
        // F_BF free; ZE[z&1] was last read by the lower half of plane z-1, whose doneL this warp has
        // already passed in its previous step (a second wait with a stale parity could be lapped)
        if (k > 0) mbar_wait(b_doneU, (k - 1) & 1);
        mbar_wait(b_loaded, k & 1);
      } else {
        mbar_wait(b_doneL, k & 1); // F_TF free
      }
      if (upper) box3::p1_step3(P, sm, H ? 2 : 1, x0, y0, z, t, C3::NPROD);
      mbar_arrive(H ? b_fullB : b_fullA);
    }
  } else if (warp < 11) {
    // ================= copy-out: dof row i = warp - 8 of node = lane =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int i = warp - 8, t = tid - 256;
    const double *rp = sm + C3::OFF_RP;
    for (int hs = 0; hs < 2 * nk; ++hs) {
      const int k = hs >> 1, H = hs & 1, z = zs + k;
      const bool lower = z >= 1, upper = z < P.nz;
      const box2::Meta M = box3::meta3_at(sm, z);
      if (!H) {
        mbar_wait(b_doneL, k & 1);
        if (lower) copy_run3(box3::run_of3(M, lane, i, 0), sm, P.values, P.overwrite, BULK);
        if (BULK) bulk_commit_wait3();
        mbar_arrive(b_emptyB);
      } else {
        mbar_wait(b_doneU, k & 1);
        copy_run3(box3::run_of3(M, lane, i, 1), sm, P.values, P.overwrite, BULK);
        if (upper) copy_run3(box3::run_of3(M, lane, i, 2), sm, P.values, P.overwrite, BULK);
        if (P.overwrite) box2::zero_uncovered(M, lane, i, P.values);
        {
          const int node = t / 3, d = t - node * 3;
          if (M.cnt[node * 4 + 1]) {
            const double v = rp[node * 3 + d] + rp[(B2::NNODE + node) * 3 + d] + rp[(2 * B2::NNODE + node) * 3 + d] +
                             rp[(3 * B2::NNODE + node) * 3 + d];
            double *r = P.rhs + (long long)M.row[node] * 3 + d;
            *r = P.overwrite ? v : *r + v;
          }
        }
        if (BULK) bulk_commit_wait3();
        mbar_arrive(b_emptyA);
      }
    }
  } else {
    // ================= loader: one plane ahead =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) mbar_arrive(b_loaded); // plane zs was staged in the prologue
    for (int k = 0; k + 1 < nk; ++k) {
      mbar_wait(b_doneL, k & 1); // plane z-2's ring slot, layer z-1's mass flow buffer and metadata z-2 are dead
      mbar_wait(b_fullA, k & 1); // the producers have consumed loaded(k): nobody can be lapped by loaded(k+1)
      load_item(P, sm, lane, x0, y0, zs + k + 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(b_loaded);
    }
  }
#if NALU_B3_PROFILE
  if (blockIdx.x == 100 && lane == 0) g_wait_acc[warp * 8 + 7] += clock64() - t_role0;
#endif
}

template <bool BULK>
static int launch3(BoxParams &P, cudaStream_t stream)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  P.tiles_x = (P.NX + B2::TBX - 1) / B2::TBX;
  P.tiles_y = (P.NY + B2::TBY - 1) / B2::TBY;
  const long long ntile = (long long)P.tiles_x * P.tiles_y;
  long long want = ((long long)sms * 12 + ntile - 1) / ntile;
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
  const size_t smem = sizeof(double) * C3::SMEM_DOUBLES;
  NALU_CUDA(cudaFuncSetAttribute(assemble_box3_kernel<BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ctas = ntile * zchunks;
  if (ctas > 0x7fffffffLL) {
    set_error("assemble_box3: grid too large");
    return NALU_B200_ERR_ARG;
  }
  P.trace = nullptr;
  const char *tr = getenv("NALU_BOX_TRACE");
  if (tr) {
    long long z[96] = {0};
    NALU_CUDA(cudaMemcpyToSymbol(g_wait_acc, z, sizeof(z)));
  }
  assemble_box3_kernel<BULK><<<(unsigned)ctas, C3::NT, smem, stream>>>(P);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  if (tr) {
    long long h[96];
    NALU_CUDA(cudaStreamSynchronize(stream));
    NALU_CUDA(cudaMemcpyFromSymbol(h, g_wait_acc, sizeof(h)));
    if (FILE *f = fopen(tr, "w")) {
      fprintf(f, "# warp: cycles waiting on loaded fullA fullB doneL doneU emptyA emptyB | total cycles in role loop\n");
      for (int w = 0; w < 12; ++w) {
        fprintf(f, "%2d:", w);
        for (int k = 0; k < 8; ++k) fprintf(f, " %9lld", h[w * 8 + k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return NALU_B200_OK;
}

int launch_assemble_box3(BoxParams &P, cudaStream_t stream)
{
  bool bulk = ((uintptr_t)P.values & 15) == 0;
  if (const char *e = getenv("NALU_BOX_BULK")) bulk = bulk && atoi(e) != 0;
  return bulk ? launch3<true>(P, stream) : launch3<false>(P, stream);
}

} // namespace nalu_b200
