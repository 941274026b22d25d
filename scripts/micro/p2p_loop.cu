// Microbenchmark: the Laplace pair loop in isolation (sources from a shared tile, no global traffic), to find
// what bounds it: FP64 issue rate, operand fetch, MUFU, or dependent-chain latency at a given occupancy.
// Prints clocks per warp-level pair iteration per SM sub-partition (ideal = 18 FP64 instr x 2.2 clk = 40).
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__device__ __forceinline__ void pair(const double4 t, const double4 sq, double& pot, double& fx, double& fy, double& fz) {
  double dx = sq.x - t.x, dy = sq.y - t.y, dz = sq.z - t.z;
  double r2 = dx * dx + dy * dy + dz * dz;
  if (V == 3) {   // same count of FP64 instructions, no MUFU, no select: issue-rate reference
    double a = fma(r2, dx, dy), b = fma(r2, dy, dz), c = fma(r2, dz, dx), d = fma(a, b, c), e = fma(b, c, a);
    double f = fma(c, a, b), g = fma(d, e, f), h = fma(e, f, d);
    pot += g; fx = fma(dx, h, fx); fy = fma(dy, h, fy); fz = fma(dz, h, fz);
    return;
  }
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
  if (V == 4) {   // mask on the seed's high word (MUFU.RSQ64H leaves the low word 0): y0 = 0 -> inv = 0 exactly
    int h = __double2hiint(y0);
    if (__double_as_longlong(r2) < __double_as_longlong(1e-8)) h = 0;
    y0 = __hiloint2double(h, 0);
  }
  if (V == 5) {   // high-word-only compare (inexact within 2^-20 of the threshold): measurement only
    int h = __double2hiint(y0);
    if (__double2hiint(r2) < 0x3e45798e) h = 0;
    y0 = __hiloint2double(h, 0);
  }
  if (V == 6) {   // mask folded into the charge: one 64-bit select on q (off the dependent chain)
    double e6 = fma(-(r2 * y0), y0, 1.0);
    double inv6 = fma(y0 * e6, fma(0.375, e6, 0.5), y0);
    double q6 = (__double_as_longlong(r2) < __double_as_longlong(1e-8)) ? 0.0 : sq.w;
    double qi6 = q6 * inv6, w6 = qi6 * (inv6 * inv6);
    pot += qi6; fx = fma(dx, w6, fx); fy = fma(dy, w6, fy); fz = fma(dz, w6, fz);
    return;
  }
  double inv;
  if (V == 2) {   // Newton step (error ~1e-13)
    double e = fma(-(r2 * y0), y0, 1.0);
    inv = fma(0.5 * y0, e, y0);
  } else {
    double e = fma(-(r2 * y0), y0, 1.0);
    inv = fma(y0 * e, fma(0.375, e, 0.5), y0);
  }
  if (V != 1 && V != 4 && V != 5) { if (__double_as_longlong(r2) < __double_as_longlong(1e-8)) inv = 0.0; }
  double qi = sq.w * inv;
  double qi3 = qi * (inv * inv);
  pot += qi;
  fx = fma(dx, qi3, fx); fy = fma(dy, qi3, fy); fz = fma(dz, qi3, fz);
}

template <int V, int UNROLL, int TPL>   // TPL = targets per lane
__global__ void __launch_bounds__(128) loop_k(double4* out, int reps, double seed) {
  __shared__ double4 tile[4][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  tile[wl][lane] = make_double4(seed * lane, seed * (lane + 1), seed * (lane + 2), 1.0 + lane);
  __syncwarp();
  double4 t[TPL];
  double pot[TPL], fx[TPL], fy[TPL], fz[TPL];
#pragma unroll
  for (int i = 0; i < TPL; ++i) {
    t[i] = make_double4(0.3 + 0.01 * lane + i, 0.7 + i, 0.2 - 0.01 * lane, 0);
    pot[i] = fx[i] = fy[i] = fz[i] = 0;
  }
  for (int r = 0; r < reps; ++r) {
#pragma unroll UNROLL
    for (int k = 0; k < 32; ++k) {
      const double4 s = tile[wl][k];
#pragma unroll
      for (int i = 0; i < TPL; ++i) pair<V>(t[i], s, pot[i], fx[i], fy[i], fz[i]);
    }
  }
  double4 acc = make_double4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < TPL; ++i) { acc.x += pot[i]; acc.y += fx[i]; acc.z += fy[i]; acc.w += fz[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V, int UNROLL, int TPL>
void run(double4* out, int sms, int blocks_per_sm, double ghz) {
  const int reps = 256 / TPL;
  const int blocks = sms * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); loop_k<V, UNROLL, TPL><<<blocks, 128>>>(out, reps, 0.01); cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double warp_iters = (double)blocks * 4 * reps * 32 * TPL;     // warp-level pair iterations
  double clk = best * 1e-3 * ghz * 1e9;
  printf("V=%d unroll=%2d targets/lane=%d warps/SM=%2d : %.3f ms  %.1f clk per warp-pair per SMSP  (%.2f Tpairs/s)\n", V,
         UNROLL, TPL, blocks_per_sm * 4, best, clk / (warp_iters / (sms * 4)), warp_iters * 32 / best / 1e9);
}
int main() {
  int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz / 1e6;
  printf("SMs %d, clock %.3f GHz\n", sms, ghz);
  double4* out; cudaMalloc(&out, sizeof(double4) * sms * 16 * 128);
  run<0, 4, 1>(out, sms, 8, ghz); run<0, 8, 1>(out, sms, 8, ghz); run<0, 4, 1>(out, sms, 4, ghz); run<0, 4, 1>(out, sms, 2, ghz);
  run<0, 4, 1>(out, sms, 12, ghz); run<0, 4, 1>(out, sms, 16, ghz);
  run<1, 4, 1>(out, sms, 8, ghz); run<2, 4, 1>(out, sms, 8, ghz); run<3, 4, 1>(out, sms, 8, ghz);
  run<4, 4, 1>(out, sms, 8, ghz); run<5, 4, 1>(out, sms, 8, ghz); run<6, 4, 1>(out, sms, 8, ghz);
  run<4, 4, 2>(out, sms, 8, ghz); run<4, 2, 2>(out, sms, 8, ghz); run<4, 4, 2>(out, sms, 6, ghz); run<5, 4, 2>(out, sms, 8, ghz);
  run<1, 4, 2>(out, sms, 8, ghz);
  run<0, 4, 2>(out, sms, 8, ghz); run<0, 2, 2>(out, sms, 8, ghz); run<0, 2, 4>(out, sms, 4, ghz); run<2, 4, 2>(out, sms, 8, ghz);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
