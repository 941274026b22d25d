// Microbenchmark: do DFMA (FP64 FMA pipe) and DMMA.8x8x4 (FP64 tensor path) overlap on this GPU?
// Runs NF independent DFMA chains and NM independent DMMA chains per loop iteration in the same warp and
// compares the time with the two alone.  If time(NF, NM) ~ max(time(NF, 0), time(0, NM)) the pipes are
// independent and a pair kernel can move its accumulation onto the tensor path for free.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NM>
__global__ void __launch_bounds__(256) mix_k(double* out, int iters, double a, double b) {
  double x[NF > 0 ? NF : 1];
  double c[NM > 0 ? NM : 1][2];
#pragma unroll
  for (int i = 0; i < NF; ++i) x[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < NF; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
      for (int i = 0; i < NM; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
template <int NF, int NM>
void run(double* out, int sms) {
  const int iters = 1024, blocks = sms * 8, threads = 256;
  float best = 1e30f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); mix_k<NF, NM><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double nf = 4.0 * NF * iters * (double)blocks * (threads / 32);   // warp-level DFMA instructions
  double nm = 4.0 * NM * iters * (double)blocks * (threads / 32);   // warp-level DMMA instructions
  printf("NF=%2d NM=%2d  %.3f ms   DFMA %.1f Ginst/s (%.2f TF)  DMMA %.1f Ginst/s (%.2f TF)\n", NF, NM, best,
         nf / best / 1e6, nf * 64 / best / 1e9, nm / best / 1e6, nm * 512 / best / 1e9);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8);
  run<16, 0>(out, sms); run<0, 2>(out, sms); run<0, 4>(out, sms); run<16, 2>(out, sms); run<16, 4>(out, sms);
  run<15, 1>(out, sms); run<15, 2>(out, sms); run<8, 4>(out, sms); run<30, 2>(out, sms); run<30, 4>(out, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
