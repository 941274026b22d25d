// Microbenchmark: FP64 FMA (DFMA) vs FP64 tensor (mma.sync m8n8k4 f64) throughput on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) dfma_k(double* out, int iters, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}
__global__ void __launch_bounds__(256) dmma_k(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048, blocks = sms * 8, threads = 256;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); dfma_k<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 64 * iters * (double)blocks * threads;
    printf("DFMA  %.2f TFLOP/s (%.3f ms)\n", fl / ms / 1e9, ms);
    cudaEventRecord(e0); dmma_k<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // per warp per mma: 8*8*4 FMA = 512 flop; 32 mma per iter per warp
    fl = 512.0 * 32 * iters * (double)blocks * (threads / 32);
    printf("DMMA  %.2f TFLOP/s (%.3f ms)\n", fl / ms / 1e9, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
