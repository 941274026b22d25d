// Round-2 probes (run once on a B200; results quoted in DESIGN.md):
//  1. L2 read bandwidth at natural clocks for the access shapes of the blocked translation kernel
//     (LDG.64 in 64-byte segments like the T_c fragments, and LDG.128 streaming), 44 MB working set.
//  2. Error of the MUFU.RSQ64H seed: max |e| and mean e^2 of e = 1 - r2*y0^2 (bounds the Newton-only pair kernel).
//  3. Latency of a dependent DMMA.8x8x4 chain (how many independent accumulators a warp needs).
//  4. 1-D TMA bulk copy (cp.async.bulk.shared::cluster.global + mbarrier complete_tx): functional check + rate.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(512) l2_read64(const double* __restrict__ T, size_t n_mats, int iters, double* out) {
  // every warp reads "A fragments": for a matrix c, k = 4kb + lk, row = 8rt + lr  ->  T[c][k][row], ld 64
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3, rt = w & 7;
  double acc = 0;
  size_t c = (blockIdx.x * 977u + w * 131u) % n_mats;
  for (int it = 0; it < iters; ++it) {
    const double* Tc = T + c * 4096;
    double a[16];
#pragma unroll
    for (int kb = 0; kb < 16; ++kb) a[kb] = __ldg(Tc + (size_t)(4 * kb + lk) * 64 + 8 * rt + lr);
#pragma unroll
    for (int kb = 0; kb < 16; ++kb) acc += a[kb];
    c = (c + 7919u) % n_mats;
  }
  if (acc == 1.2345) out[0] = acc;
}
__global__ void __launch_bounds__(512) l2_read128(const double2* __restrict__ T, size_t n16, int iters, double* out) {
  double acc = 0;
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) % n16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double2 v = __ldg(T + i);
      acc += v.x + v.y;
      i += stride; if (i >= n16) i -= n16;
    }
  }
  if (acc == 1.2345) out[0] = acc;
}

__global__ void seed_error(unsigned long long n, double* out) {  // out[0] = max |e|, out[1] = sum e^2, out[2] = sum e
  double mx = 0, s2 = 0, s1 = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    // r2 log-uniform in [1e-8, 4), mantissa bits scrambled
    unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    double r2 = exp(log(1e-8) + u * (log(4.0) - log(1e-8)));
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
    double e = fma(-(r2 * y0), y0, 1.0);
    mx = fmax(mx, fabs(e)); s2 += e * e; s1 += e;
  }
  for (int o = 16; o; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    s2 += __shfl_xor_sync(0xffffffffu, s2, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(mx));
    atomicAdd(out + 1, s2); atomicAdd(out + 2, s1);
  }
}

template <int CH>
__global__ void __launch_bounds__(32) dmma_chain(double* out, int iters, double a, double b, long long* clk) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

// ---- TMA 1-D bulk copy ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
      ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// each 1-warp block streams `tiles` tiles of 1 KB (32 double4) through a 2-stage ring and sums them
__global__ void __launch_bounds__(32) tma_stream(const double4* __restrict__ src, size_t n, int tiles, double* out, int* bad) {
  __shared__ __align__(128) double4 buf[2][32];
  __shared__ __align__(8) uint64_t bar[2];
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  size_t base = ((size_t)blockIdx.x * 4099u * 32u) % (n - (size_t)tiles * 32 - 32);
  base &= ~(size_t)0;   // 32-byte elements: always 16-byte aligned
  if (lane == 0) { mbar_expect_tx(&bar[0], 1024); bulk_g2s(buf[0], src + base, 1024, &bar[0]); }
  double acc = 0;
  int nbad = 0;
  for (int t = 0; t < tiles; ++t) {
    const int s = t & 1;
    if (lane == 0 && t + 1 < tiles) { mbar_expect_tx(&bar[s ^ 1], 1024); bulk_g2s(buf[s ^ 1], src + base + (size_t)(t + 1) * 32, 1024, &bar[s ^ 1]); }
    mbar_wait(&bar[s], (t >> 1) & 1);
    const double4 v = buf[s][lane];
    const double4 w = src[base + (size_t)t * 32 + lane];
    if (v.x != w.x || v.w != w.w) ++nbad;
    acc += v.x + v.y + v.z + v.w;
    __syncwarp();
  }
  if (nbad) atomicAdd(bad, nbad);
  if (acc == 1.2345) out[0] = acc;
}

int main() {
  int sms = 0, khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  printf("SMs %d clock %.3f GHz\n", sms, khz / 1e6);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double* out; CK(cudaMalloc(&out, 64)); CK(cudaMemset(out, 0, 64));
  // 1. L2 bandwidth
  for (size_t mats : {1360u, 512u, 8192u}) {
    double* T; CK(cudaMalloc(&T, mats * 32768)); CK(cudaMemset(T, 0, mats * 32768));
    for (int cta : {1, 2}) {
      const int iters = 400, blocks = sms * cta;
      float best = 1e30f;
      for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); l2_read64<<<blocks, 512>>>(T, mats, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
      }
      double bytes = (double)blocks * 16 * iters * 16 * 32 * 8;
      printf("L2 read, T-fragment shape (LDG.64, 64-B segments), %zu matrices (%.0f MB), %d CTA/SM x 512 thr: %.3f ms  %.2f TB/s\n", mats,
             mats * 32768 / 1e6, cta, best, bytes / best / 1e9);
      float best2 = 1e30f;
      const int it2 = 200;
      for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); l2_read128<<<blocks, 512>>>((const double2*)T, mats * 2048, it2, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best2) best2 = ms;
      }
      printf("L2 read, LDG.128 streaming, same set: %.3f ms  %.2f TB/s\n", best2, (double)blocks * 512 * it2 * 8 * 16 / best2 / 1e9);
    }
    cudaFree(T);
  }
  // 2. seed error
  {
    CK(cudaMemset(out, 0, 64));
    const unsigned long long n = 1ull << 28;
    seed_error<<<sms * 8, 256>>>(n, out);
    double h[3]; CK(cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost));
    printf("rsqrt.approx.ftz.f64 seed: max|e| = %.4e (2^%.2f), rms e = %.4e, mean e = %.4e  (e = 1 - r2 y0^2 over %llu log-uniform r2)\n",
           h[0], log2(h[0]), sqrt(h[1] / n), h[2] / n, n);
    printf("  Newton-only inverse root: relative error <= 3/8 max e^2 = %.3e, mean bias 3/8 E[e^2] = %.3e\n", 0.375 * h[0] * h[0], 0.375 * h[1] / n);
  }
  // 3. DMMA chain latency
  {
    long long* clk; CK(cudaMalloc(&clk, 8));
    long long h;
    const int iters = 4096;
    dmma_chain<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9, clk); CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
    printf("DMMA.8x8x4 dependent chain, 1 warp: %.1f clk per DMMA\n", (double)h / iters);
    dmma_chain<2><<<1, 32>>>(out, iters, 1.0000001, 1e-9, clk); CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
    printf("  2 independent chains: %.1f clk per DMMA\n", (double)h / iters / 2);
    dmma_chain<4><<<1, 32>>>(out, iters, 1.0000001, 1e-9, clk); CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
    printf("  4 independent chains: %.1f clk per DMMA\n", (double)h / iters / 4);
    dmma_chain<8><<<1, 32>>>(out, iters, 1.0000001, 1e-9, clk); CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
    printf("  8 independent chains: %.1f clk per DMMA\n", (double)h / iters / 8);
  }
  // 4. TMA bulk
  {
    const size_t n = 8u << 20;   // 8M double4 = 256 MB
    double4* src; CK(cudaMalloc(&src, n * sizeof(double4)));
    double* hp = (double*)malloc(1 << 20);
    for (int i = 0; i < (1 << 17); ++i) hp[i] = i * 0.5 + 1;
    for (size_t off = 0; off < n * 32; off += 1 << 20) CK(cudaMemcpy((char*)src + off, hp, 1 << 20, cudaMemcpyHostToDevice));
    int* bad; CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
    const int tiles = 64, blocks = sms * 20 * 4;
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0); tma_stream<<<blocks, 32>>>(src, n, tiles, out, bad); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int hb; CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
    printf("TMA 1-D bulk copy (cp.async.bulk + mbarrier), 1 KB tiles, 1-warp blocks: mismatches %d, %.3f ms, %.2f TB/s staged\n", hb, best,
           (double)blocks * tiles * 1024 / best / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
