// ubench.cu -- micro-benchmarks that ground the design of the tiled assembly kernel:
//   lds   : shared-memory return bandwidth of LDS.64 / LDS.128 for distinct, 8-lane-broadcast and
//           uniform addresses (is a broadcast 64-bit load cheaper than a distinct one?)
//   dfma  : DFMA issue rate for w warps per SM with ILP independent chains per thread
//   bulk  : cp.async.bulk shared->global rate for small runs (208 B ... 15.5 KB)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ double lds64(unsigned a)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds128(unsigned a, double &x, double &y)
{
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}

// mode 0: lane-distinct, 1: 8-lane broadcast (4 addresses per warp), 2: uniform
template <int WIDE>
__global__ void k_lds(double *out, int mode, int iters, long long *cyc)
{
  extern __shared__ __align__(16) double s[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int idx = mode == 0 ? lane : (mode == 1 ? (lane >> 3) : 0);
  unsigned base = (unsigned)__cvta_generic_to_shared(s) + (warp * 64 + idx * (WIDE ? 2 : 1)) * 8;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (WIDE) {
        double x, y;
        lds128(base + u * 1024, x, y);
        a0 += x; a1 += y;
      } else {
        a0 += lds64(base + u * 1024);
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void k_dfma(double *out, double a, double b, int iters, long long *cyc)
{
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// store drain: every warp writes `nrun` runs of 27 doubles (216 bytes, lane k -> double k) back to back to its own
// contiguous stretch of a buffer much larger than L2 -- what the copy-out of the tiled momentum kernel does.
// `vec`: lanes 0..13 store a double2 instead (the 27th double is dropped; alignment of the run start permitting)
__global__ void k_store_runs(double *dst, long long nrun, int vec, long long *cyc)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double *base = dst + ((long long)blockIdx.x * nw + warp) * nrun * 28; // 28: keeps the double2 variant 16-byte aligned
  const double v = (double)lane;
  long long t0 = clock64();
  if (vec == 2) { // 27-double runs back to back, no holes: sectors shared by two runs are completed by the next run
    for (long long r = 0; r < nrun; ++r)
      if (lane < 27) base[r * 27 + lane] = v;
  } else if (!vec) {
    for (long long r = 0; r < nrun; ++r)
      if (lane < 27) base[r * 28 + lane] = v;
  } else {
    for (long long r = 0; r < nrun; ++r)
      if (lane < 14) reinterpret_cast<double2 *>(base + r * 28)[lane] = make_double2(v, v);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// every thread < nrun issues one bulk copy of `bytes` from its own smem run to its own global run
__global__ void k_bulk(double *dst, int bytes, int nrun, int iters, long long *cyc)
{
  extern __shared__ __align__(128) double s[];
  for (int i = threadIdx.x; i < nrun * (bytes / 8); i += blockDim.x) s[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  if ((int)threadIdx.x < nrun) {
    const unsigned src = (unsigned)__cvta_generic_to_shared(s) + threadIdx.x * bytes;
    char *g = (char *)dst + ((size_t)blockIdx.x * nrun + threadIdx.x) * (size_t)bytes;
    for (int i = 0; i < iters; ++i) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(src), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// issue cost: `nw` warps, every lane issues `nops` bulk copies back to back (waterfall per warp instruction),
// clock per warp-level issue loop; no wait inside the timed region
__global__ void k_bulk_issue(double *dst, int bytes, int nops, long long *cyc)
{
  extern __shared__ __align__(128) double s[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const unsigned src = (unsigned)__cvta_generic_to_shared(s) + (threadIdx.x & 63) * 512;
  char *g = (char *)dst + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 512;
  long long t0 = clock64();
  for (int i = 0; i < nops; ++i)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(src), "r"(bytes) : "memory");
  long long t1 = clock64();
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double *out;
  long long *cyc, h;
  CK(cudaMalloc(&out, 1ull << 30));
  CK(cudaMalloc(&cyc, 8));
  printf("{\"sms\": %d", sms);
  // ---- LDS
  const char *mn[3] = {"distinct", "bcast8", "uniform"};
  for (int wide = 0; wide < 2; ++wide)
    for (int mode = 0; mode < 3; ++mode) {
      const int threads = 256, iters = 2000;
      if (wide) {
        CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        k_lds<1><<<sms, threads, 65536>>>(out, mode, iters, cyc);
      } else {
        CK(cudaFuncSetAttribute(k_lds<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        k_lds<0><<<sms, threads, 65536>>>(out, mode, iters, cyc);
      }
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
      // 8 warps each issue iters*8 loads: cycles per warp-load per SM
      printf(",\n \"lds%d_%s_cyc_per_warp_instr\": %.3f", wide ? 128 : 64, mn[mode], (double)h / (iters * 8.0 * (threads / 32)));
    }
  // ---- DFMA
#define RUN_DFMA(ILP, W)                                                                             \
  {                                                                                                  \
    const int iters = 2000;                                                                          \
    k_dfma<ILP><<<sms, 32 * W>>>(out, 0.999, 0.001, iters, cyc);                                     \
    CK(cudaDeviceSynchronize());                                                                     \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                              \
    printf(",\n \"dfma_ilp%d_w%d_warp_instr_per_clk_per_sm\": %.3f", ILP, W, iters * 8.0 * ILP * W / (double)h); \
  }
  RUN_DFMA(1, 1) RUN_DFMA(2, 1) RUN_DFMA(4, 1) RUN_DFMA(8, 1)
  RUN_DFMA(1, 4) RUN_DFMA(2, 4) RUN_DFMA(4, 4) RUN_DFMA(8, 4)
  RUN_DFMA(1, 8) RUN_DFMA(2, 8) RUN_DFMA(4, 8) RUN_DFMA(2, 16)
  // ---- store drain to HBM (8 GB footprint >> L2), 2 CTAs per SM, 4 / 8 / 16 warps per CTA
  {
    double *big = nullptr;
    const size_t bytes_total = (size_t)8 << 30;
    CK(cudaMalloc(&big, bytes_total));
    for (int vec = 0; vec < 3; ++vec)
      for (int nw = 4; nw <= 16; nw *= 2) {
        const long long nrun = (long long)(bytes_total / 8 / 28) / ((long long)sms * 2 * nw);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_store_runs<<<sms * 2, 32 * nw>>>(big, nrun / 8, vec, cyc);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_store_runs<<<sms * 2, 32 * nw>>>(big, nrun, vec, cyc);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double b = (double)sms * 2 * nw * nrun * (vec == 1 ? 224.0 : 216.0);
        printf(",\n \"store_runs_%s_w%d\": {\"GBps\": %.1f, \"bytes_per_clk_per_sm\": %.2f, \"cyc_per_run_per_warp\": %.1f}",
               vec == 1 ? "double2" : (vec == 2 ? "double_contig" : "double_holes"), nw, b / (ms * 1e-3) / 1e9, b / sms / ((double)h), (double)h / nrun);
      }
    CK(cudaFree(big));
  }
  // ---- bulk stores
  const int sizes[5] = {208, 432, 1936, 15552, 208};
  const int nruns[5] = {96, 96, 32, 4, 32};
  for (int k = 0; k < 5; ++k) {
    const int iters = 400, bytes = sizes[k], nrun = nruns[k];
    const int smem = bytes * nrun;
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_bulk<<<sms * 2, 128, smem>>>(out, bytes, nrun, 10, cyc);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_bulk<<<sms * 2, 128, smem>>>(out, bytes, nrun, iters, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double ops = (double)sms * 2 * nrun * iters;
    printf(",\n \"bulk_%dB_x%d\": {\"GBps\": %.1f, \"Mops_per_s_per_sm\": %.2f, \"cyc_per_round\": %.1f}", bytes, nrun,
           ops * bytes / (ms * 1e-3) / 1e9, ops / (ms * 1e-3) / 1e6 / sms, (double)h / iters);
  }
  for (int nw = 1; nw <= 4; nw *= 4)
    for (int bytes = 208; bytes <= 432; bytes += 224) {
      CK(cudaFuncSetAttribute(k_bulk_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      k_bulk_issue<<<sms, 32 * nw, 65536>>>(out, bytes, 8, cyc);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
      printf(",\n \"bulk_issue_%dB_w%d_cyc_per_lane_op\": %.1f", bytes, nw, (double)h / (8.0 * 32));
    }
  printf("}\n");
  return 0;
}
