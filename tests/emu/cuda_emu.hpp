// tests/emu/cuda_emu.hpp -- TEST INFRASTRUCTURE ONLY.
// Lock-step CPU emulation of a CUDA launch, enough to execute the warp-synchronous kernels of
// fmm_bem_relaxed_b200/csrc/stokes.cu and stokes_bem.cu UNCHANGED on a machine without a GPU: one std::thread per CUDA
// thread of a block, blocks one after the other, __syncwarp / __syncthreads as std::barrier, static __shared__ arrays
// as function-local statics (one block runs at a time), the dynamic shared segment as one global buffer, __constant__
// arrays as ordinary globals (cudaMemcpyToSymbol = memcpy).  No warp shuffles, no PTX: kernels that need them are
// compiled but must not be launched here.
// This is a checker for kernel LOGIC (indexing, staging, the arithmetic of every lane); it says nothing about
// performance and does not replace a run on the device.
#pragma once
#define __global__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __launch_bounds__(...)
#include <cuda_runtime.h>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include <algorithm>

namespace emu { template <class T> void* symbol(T& x) { return (void*)&x; } }     // arrays and objects alike
#define cudaMemcpyToSymbol(dst, src, n) (std::memcpy(emu::symbol(dst), (src), (n)), cudaSuccess)
#define cudaGetDevice(p) (*(p) = 0, cudaSuccess)

namespace emu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
inline thread_local Idx threadIdx_, blockIdx_, blockDim_, gridDim_;
inline thread_local std::barrier<>* warp_bar = nullptr;
inline thread_local std::barrier<>* block_bar = nullptr;
alignas(16) inline unsigned char dyn_shared[232448];

inline size_t dyn_shared_limit = sizeof dyn_shared;     // bytes the current launch asked for (guard bytes follow)
inline long launches = 0, guard_failures = 0;

template <class F>
void launch(dim3 grid, dim3 block, F&& body) {
  if (block.x % 32 != 0 || block.y != 1 || block.z != 1) { fprintf(stderr, "emu: block must be a multiple of 32 x 1 x 1\n"); exit(3); }
  const unsigned nw = block.x / 32;
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      std::vector<std::unique_ptr<std::barrier<>>> wb;
      for (unsigned w = 0; w < nw; ++w) wb.emplace_back(new std::barrier<>(32));
      std::barrier<> bb(block.x);
      std::vector<std::thread> th;
      for (unsigned t = 0; t < block.x; ++t)
        th.emplace_back([&, t] {
          threadIdx_ = {t, 0, 0}; blockIdx_ = {bx, by, 0}; blockDim_ = {block.x, 1, 1}; gridDim_ = {grid.x, grid.y, 1};
          warp_bar = wb[t / 32].get(); block_bar = &bb;
          body();
          // a CUDA thread that has exited no longer takes part in barriers
          warp_bar->arrive_and_drop(); block_bar->arrive_and_drop();
        });
      for (auto& t : th) t.join();
    }
}

// A launch as the host code writes it, kernel<<<grid, block, shmem, stream>>>(args): the dynamic shared segment is
// exactly `shmem` bytes, followed by guard bytes that a kernel indexing past its request would overwrite.
template <class F>
void launch_cfg_impl(dim3 grid, dim3 block, size_t shmem, F&& body) {
  ++launches;
  if (grid.x == 0 || grid.y == 0 || block.x == 0 || block.x > 1024 || shmem + 64 > sizeof dyn_shared || shmem > 232448 - 1024) {
    fprintf(stderr, "emu: invalid launch configuration grid (%u, %u) block %u shared %zu\n", grid.x, grid.y, block.x, shmem);
    exit(5);
  }
  std::memset(dyn_shared + shmem, 0xA5, 64);
  launch(grid, block, body);
  for (int i = 0; i < 64; ++i)
    if (dyn_shared[shmem + i] != 0xA5) { ++guard_failures; break; }
}
inline dim3 to_dim3(dim3 d) { return d; }
template <class T> dim3 to_dim3(T v) { return dim3((unsigned)v); }
template <class G, class B, class F>
void launch_cfg(G grid, B block, size_t shmem, F&& body) { launch_cfg_impl(to_dim3(grid), to_dim3(block), shmem, body); }
}  // namespace emu

// ---- the few runtime calls the host side of a csrc/*.cu file makes, on host memory (after the CUDA headers: these
// ---- function-like macros replace the calls, not the declarations)
#define cudaMalloc(pp, bytes) (*(pp) = std::malloc(bytes), cudaSuccess)
#define cudaFree(p) (std::free(p), cudaSuccess)
#define cudaMemcpyAsync(dst, src, bytes, kind, stream) (std::memcpy((dst), (src), (bytes)), cudaSuccess)
#define cudaMemsetAsync(dst, v, bytes, stream) (std::memset((dst), (v), (bytes)), cudaSuccess)
#define cudaStreamSynchronize(s) (cudaSuccess)
#define cudaEventRecord(e, s) (cudaSuccess)
#define cudaStreamWaitEvent(s, e, f) (cudaSuccess)
#define cudaGetLastError() (cudaSuccess)
#define cudaFuncSetAttribute(f, a, v) (cudaSuccess)
#define cudaFuncGetAttributes(a, f) (cudaSuccess)
#define cudaLaunchCooperativeKernel(f, g, b, a, sh, s) (cudaErrorNotSupported)   // grid barriers are not emulated

#define threadIdx emu::threadIdx_
#define blockIdx emu::blockIdx_
#define blockDim emu::blockDim_
#define gridDim emu::gridDim_
inline void __syncwarp() { emu::warp_bar->arrive_and_wait(); }
inline void __syncthreads() { emu::block_bar->arrive_and_wait(); }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __shfl_sync(unsigned, double, int) { fprintf(stderr, "emu: warp shuffles are not emulated\n"); abort(); }
inline double __shfl_xor_sync(unsigned, double, int) { fprintf(stderr, "emu: warp shuffles are not emulated\n"); abort(); }
inline unsigned __shfl_sync(unsigned, unsigned, int) { fprintf(stderr, "emu: warp shuffles are not emulated\n"); abort(); }
inline unsigned __shfl_up_sync(unsigned, unsigned, int) { fprintf(stderr, "emu: warp shuffles are not emulated\n"); abort(); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline long long clock64() { return 0; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, sizeof d); return d; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }   // blocks run one at a time and
inline void __threadfence() {}                                                               // one thread per block calls it
using std::min;
using std::max;
