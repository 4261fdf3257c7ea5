// fp64_peak.cu -- micro-benchmarks for the FP64 roof of this part (not in MEASURED_PEAKS.json):
//   dfma : dependent-chain-free DFMA throughput (8 independent chains / thread)
//   dmma : mma.sync.m8n8k4.f64 throughput
//   mixed: both interleaved (are the DFMA and DMMA pipes additive?)
//   shfl / lds : rate of the "free" companion instructions next to DFMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double *out, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double *out, double a, double b)
{
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    dmma(c0, c1, a, b); dmma(c2, c3, a, b); dmma(c4, c5, a, b); dmma(c6, c7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

__global__ void k_mixed(double *out, double a, double b)
{
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    dmma(c0, c1, a, b);
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    dmma(c2, c3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// DFMA with one LDS.64 (shared-memory operand) per DFMA
__global__ void k_dfma_lds(double *out, double a)
{
  __shared__ double s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = i * 1e-3;
  __syncthreads();
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  int j = threadIdx.x;
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    x0 = fma(x0, a, s[(j) & 1023]); x1 = fma(x1, a, s[(j + 32) & 1023]);
    x2 = fma(x2, a, s[(j + 64) & 1023]); x3 = fma(x3, a, s[(j + 96) & 1023]);
    j += 128;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

// DFMA with one 64-bit shuffle (2 SHFL) per DFMA
__global__ void k_dfma_shfl(double *out, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    x0 = fma(__shfl_xor_sync(0xffffffffu, x1, 1), a, b); x1 = fma(__shfl_xor_sync(0xffffffffu, x2, 2), a, b);
    x2 = fma(__shfl_xor_sync(0xffffffffu, x3, 4), a, b); x3 = fma(__shfl_xor_sync(0xffffffffu, x0, 8), a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

template <typename F>
static float timeit(F f, int reps)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main()
{
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int threads = 512, ctas = sms * 4;
  double *out;
  cudaMalloc(&out, sizeof(double) * threads * ctas);
  const double nthreads = (double)threads * ctas;
  float ms;
  ms = timeit([&] { k_dfma<<<ctas, threads>>>(out, 0.999, 0.001); }, 10);
  double dfma = nthreads * ITERS * 8 * 2 / (ms * 1e-3) / 1e12;
  printf("{\"sms\": %d, \"clock_mhz_attr\": %d,\n \"dfma_tflops\": %.2f, \"dfma_per_clk_per_sm_at_attr_clock\": %.1f,\n", sms, clk / 1000,
         dfma, dfma * 1e12 / 2 / sms / (clk * 1e3));
  ms = timeit([&] { k_dmma<<<ctas, threads>>>(out, 0.999, 0.001); }, 10);
  // each warp-level mma.m8n8k4 = 8*8*4 = 256 FMA = 512 flop; 4 per iter per warp
  double dmma_t = (nthreads / 32) * ITERS * 4 * 512 / (ms * 1e-3) / 1e12;
  printf(" \"dmma_tflops\": %.2f,\n", dmma_t);
  ms = timeit([&] { k_mixed<<<ctas, threads>>>(out, 0.999, 0.001); }, 10);
  double mixed = ((nthreads / 32) * ITERS * 2 * 512 + nthreads * ITERS * 8 * 2) / (ms * 1e-3) / 1e12;
  printf(" \"mixed_dfma_dmma_tflops\": %.2f,\n", mixed);
  ms = timeit([&] { k_dfma_lds<<<ctas, threads>>>(out, 0.999); }, 10);
  printf(" \"dfma_with_lds64_tflops\": %.2f,\n", nthreads * ITERS * 4 * 2 / (ms * 1e-3) / 1e12);
  ms = timeit([&] { k_dfma_shfl<<<ctas, threads>>>(out, 0.999, 0.001); }, 10);
  printf(" \"dfma_with_shfl64_tflops\": %.2f}\n", nthreads * ITERS * 4 * 2 / (ms * 1e-3) / 1e12);
  return 0;
}
