// fp64_peak.cu — measured fp64 roofs of the box (SURVEY.md 8d asks for a DFMA loop):
//   * vector DFMA: 8 independent FMA chains per thread, grid = SMs x 2 blocks of 1024 threads
//   * tensor DMMA: mma.sync.aligned.m8n8k4.f64, 4 independent accumulator chains per warp
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void k_dfma(double *out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

__global__ void k_dmma(double *out, int iters, double a, double b) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

int main() {
  int dev = 0, sms = 0, mhz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
  double *out;
  cudaMalloc(&out, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 16, threads = 1024, blocks = sms * 2;
  float best_fma = 1e30f, best_mma = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best_fma) best_fma = ms;
    cudaEventRecord(e0);
    k_dmma<<<blocks, threads>>>(out, iters / 4, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best_mma) best_mma = ms;
  }
  const double fma_flops = 2.0 * 8 * iters * (double)threads * blocks;
  const double mma_flops = 2.0 * 256 * 4 * (iters / 4) * (double)(threads / 32) * blocks;  // 8x8x4 FMAs per warp instruction
  cudaError_t e = cudaGetLastError();
  printf("{\"device_sms\": %d, \"clock_mhz\": %d, \"dfma_tflops\": %.3f, \"dfma_ms\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, "
         "\"dmma_ms\": %.3f, \"dfma_per_clk_per_sm\": %.2f, \"dmma_warp_inst_per_clk_per_sm\": %.4f, \"cuda_error\": \"%s\"}\n",
         sms, mhz / 1000, fma_flops / (best_fma * 1e-3) / 1e12, best_fma, mma_flops / (best_mma * 1e-3) / 1e12, best_mma,
         fma_flops / 2 / (best_fma * 1e-3) / sms / (mhz * 1e3), mma_flops / 512 / (best_mma * 1e-3) / sms / (mhz * 1e3),
         cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}
