// csrc/m2l_classes.cu -- batched M2L: translation classes + dense FP64 contraction.
//
// The reference applies kernel/LaplaceSpherical.hpp:296-329 pair by pair, re-evaluating the
// (2P)^2 singular harmonics of c_tgt - c_src every time.  In an octree the translation vector
// takes few distinct values: box centres sit on a lattice of half finest cells, so two pairs
// with the same integer offset share one translation operator.  M2L is linear in the source
// multipole, hence for a class c
//      L_real[:, pair] = T_c (P^2 x P^2, real)  *  M_real[:, pair]
// where a multipole is stored as P^2 reals (Re M_n^m at n^2+n+m, Im M_n^m at n^2+n-m; the
// imaginary part of m = 0 is identically zero).  T_c is built once per plan and order from the
// same Cnm and evalLocal terms the reference uses.
//
// Phase 1 (m2l_gemm_kernel): one CTA per (class, 128 pairs): gathers the source multipoles,
//   multiplies by T_c out of shared memory with an 8xRPT register tile per thread (FP64 FMA) and
//   writes one column per pair into a scratch array ordered target-major.
// Phase 2 (m2l_reduce_kernel): one CTA per target box sums its (contiguous) columns -- no
//   atomics, fixed order, bit-reproducible -- and writes the packed complex local expansion.
// Pairs whose class is too small to batch (tails of adaptive trees, top levels) stay on the
// per-pair kernel in laplace.cu, which accumulates on top.
#include "common.cuh"
#include "laplace_ops.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace fmmb {

namespace {

using namespace ops;

constexpr int kNB = 64;           // pairs (columns) per pipeline stage
constexpr int kMinPopM2L = 1;     // smallest M2L class that is worth a GEMM tile

__host__ __device__ __forceinline__ unsigned compact10(unsigned x) {
  x &= 0x09249249u;
  x = (x | (x >> 2)) & 0x030C30C3u;
  x = (x | (x >> 4)) & 0x0300F00Fu;
  x = (x | (x >> 8)) & 0x030000FFu;
  x = (x | (x >> 16)) & 0x000003FFu;
  return x;
}

struct Temp {
  DevBuf<char> buf;
  void* get(size_t bytes) { if (bytes > buf.cap) buf.resize(bytes); return buf.p; }
};
inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

// centre of a box in units of half a finest cell (exact integers)
__device__ __forceinline__ int3 centre_half_cells(unsigned key, unsigned L) {
  unsigned m = (key & 0x7fffffffu) << (3u * (10 - L));
  m &= ~(1u << 30);
  int half = L >= 10 ? 1 : (1 << (10 - L));       // box spans 2^(11-L) half cells
  return make_int3(2 * (int)compact10(m) + half, 2 * (int)compact10(m >> 1) + half,
                   2 * (int)compact10(m >> 2) + half);
}

__global__ void slot_targets(const int* __restrict__ off, int nb, int* __restrict__ tgt) {
  int b = blockIdx.x;
  if (b >= nb) return;
  for (int e = off[b] + threadIdx.x; e < off[b + 1]; e += blockDim.x) tgt[e] = b;
}
// M2M: slot = child box c, pair (src = c, tgt = parent[c]); L2L: (src = parent[c], tgt = c)
__global__ void parent_child_pairs(const unsigned* __restrict__ parent, int nb, int child_is_target,
                                   int* __restrict__ tgt, int* __restrict__ src) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb) return;
  int p = c == 0 ? 0 : (int)parent[c];
  tgt[c] = child_is_target ? c : p;
  src[c] = child_is_target ? p : c;
}

// class key of slot first+e: integer centre offset (36 bits) [+ target level above it]
__global__ void class_keys(const int* __restrict__ tgt, const int* __restrict__ src, int first,
                           const int* __restrict__ slot_list, int64_t n,
                           const unsigned* __restrict__ key, const unsigned* __restrict__ lvl, int with_level,
                           unsigned long long* __restrict__ ckey, int* __restrict__ slot) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  int sl = slot_list ? slot_list[e] : first + (int)e;
  int t = tgt[sl], s = src[sl];
  int3 a = centre_half_cells(key[t], lvl[t]), b = centre_half_cells(key[s], lvl[s]);
  unsigned long long dx = (unsigned)(a.x - b.x + 2048), dy = (unsigned)(a.y - b.y + 2048),
                     dz = (unsigned)(a.z - b.z + 2048);
  unsigned long long k = (dx << 24) | (dy << 12) | dz;
  if (with_level) k |= (unsigned long long)lvl[t] << 36;
  ckey[e] = k;
  slot[e] = sl;
}

// per class: number of GEMM items (0 if the class is too small)
__global__ void class_items(const int* __restrict__ count, int nclasses, int minpop, int* __restrict__ nitems) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nclasses) return;
  nitems[c] = (c < nclasses && count[c] >= minpop) ? (count[c] + kNB - 1) / kNB : 0;
}
__global__ void fill_items(const int* __restrict__ count, const int* __restrict__ start,
                           const int* __restrict__ item_off, int nclasses, int* __restrict__ item_class,
                           int* __restrict__ item_start, int* __restrict__ item_count, int4* __restrict__ item_desc) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nclasses) return;
  int ni = item_off[c + 1] - item_off[c];
  for (int i = 0; i < ni; ++i) {
    int it = item_off[c] + i;
    item_class[it] = c;
    item_start[it] = start[c] + i * kNB;
    item_count[it] = min(kNB, count[c] - i * kNB);
    item_desc[it] = make_int4(c, start[c] + i * kNB, min(kNB, count[c] - i * kNB), 0);
  }
}
__global__ void gather_sorted_src(const int* __restrict__ sorted_slot, const int* __restrict__ src, int64_t n,
                                  int* __restrict__ sorted_src) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sorted_src[i] = src[sorted_slot[i]];
}
// mark slots that the batched path covers; the rest go to the per-pair kernel
__global__ void mark_batched(const int* __restrict__ count, const int* __restrict__ start, int nclasses, int minpop,
                             const int* __restrict__ sorted_slot, unsigned char* __restrict__ batched) {
  int c = blockIdx.x;
  if (c >= nclasses) return;
  unsigned char v = count[c] >= minpop;
  for (int i = threadIdx.x; i < count[c]; i += blockDim.x) batched[sorted_slot[start[c] + i]] = v;
}
__global__ void residual_flags(const unsigned char* __restrict__ batched, int64_t n, int* __restrict__ flag) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n) flag[e] = batched[e] ? 0 : 1;
  if (e == n) flag[e] = 0;
}
__global__ void residual_compact(const int* __restrict__ flag, const int* __restrict__ pos, int64_t n,
                                 const int* __restrict__ src, int* __restrict__ res_src) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n && flag[e]) res_src[pos[e]] = src[e];
}
__global__ void residual_offsets(const int* __restrict__ off, const int* __restrict__ pos, int nb,
                                 int* __restrict__ res_off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= nb) res_off[b] = pos[off[b]];
}
// Representative translation vector of a class: the integer centre offset times half a finest cell.
// (The floating-point centre differences of the pairs of one class agree with it to ~1e-16 relative;
// using the lattice value makes T_c independent of which pairs a rank happens to hold.)
__global__ void class_vectors(const int* __restrict__ start, int nclasses, const int* __restrict__ sorted_slot,
                              const int* __restrict__ tgt, const int* __restrict__ src,
                              const unsigned* __restrict__ key, const unsigned* __restrict__ lvl, double3 half_cell,
                              double4* __restrict__ vec) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nclasses) return;
  int e = sorted_slot[start[c]];
  int t = tgt[e], s = src[e];
  int3 a = centre_half_cells(key[t], lvl[t]), b = centre_half_cells(key[s], lvl[s]);
  vec[c] = make_double4((a.x - b.x) * half_cell.x, (a.y - b.y) * half_cell.y, (a.z - b.z) * half_cell.z, 0.0);
}

// ---- translation matrices: Tt[c][col][row], rows contiguous, ld = P^2 -------------------------------
// real layout of an expansion: Re X_n^m at n^2+n+m (m >= 0), Im X_n^m at n^2+n-m (m > 0)
__global__ void __launch_bounds__(256)
build_T_m2l(int P, const double4* __restrict__ vec, double* __restrict__ Tt) {
  extern __shared__ double2 Y[];       // (2P)^2
  const int pp = P * P, c = blockIdx.x;
  double4 v = vec[c];
  Sph s = to_sph(v.x, v.y, v.z);
  for (int m = threadIdx.x; m < 2 * P; m += blockDim.x) harmonics_column<true>(m, 2 * P, s, 1.0, Y);
  __syncthreads();
  double* T = Tt + (size_t)c * pp * pp;
  for (int idx = threadIdx.x; idx < pp * pp; idx += blockDim.x) {
    int col = idx / pp, row = idx % pp;
    int j = 0; while ((j + 1) * (j + 1) <= row) ++j;
    int kk = row - j * j - j;            // >= 0: Re L_j^k, < 0: Im L_j^{-kk}
    int n = 0; while ((n + 1) * (n + 1) <= col) ++n;
    int mm = col - n * n - n;            // >= 0: Re M_n^m, < 0: Im M_n^{-mm}
    int k = abs(kk), m = abs(mm);
    int base = (j + n) * (j + n) + j + n - k;
    // W(+-m) = Cnm(+-m) * Y_{j+n}^{+-m-k};  M = a + ib contributes a (W+ + W-) + i b (W+ - W-)
    double cp_ = cnm_real(j, k, n, m);
    double2 yp = Y[base + m];
    double wpr = cp_ * yp.x, wpi = cp_ * yp.y;
    double val;
    if (m == 0) {
      val = kk >= 0 ? wpr : wpi;
    } else {
      double cm_ = cnm_real(j, k, n, -m);
      double2 ym = Y[base - m];
      double wmr = cm_ * ym.x, wmi = cm_ * ym.y;
      if (kk >= 0) val = (mm >= 0) ? (wpr + wmr) : -(wpi - wmi);
      else val = (mm >= 0) ? (wpi + wmi) : (wpr - wmr);
    }
    T[(size_t)col * pp + row] = val;
  }
}

// M2M / L2L matrices by probing the operator with unit vectors (the operators contain a complex
// conjugation, so they are linear over the reals only).
template <int KIND>   // 1 = M2M, 2 = L2L
__global__ void __launch_bounds__(128)
build_T_probe(int P, const double4* __restrict__ vec, double* __restrict__ Tt) {
  extern __shared__ double2 shp[];
  const int pp = P * P, nc = P * (P + 1) / 2, c = blockIdx.x;
  double2* Y = shp;          // pp
  double2* E = shp + pp;     // nc
  double4 v = vec[c];
  Sph s = to_sph(v.x, v.y, v.z);
  for (int m = threadIdx.x; m < P; m += blockDim.x) harmonics_column<false>(m, P, s, KIND == 1 ? -1.0 : 1.0, Y);
  double* T = Tt + (size_t)c * pp * pp;
  for (int col = 0; col < pp; ++col) {
    int n = 0; while ((n + 1) * (n + 1) <= col) ++n;
    int mm = col - n * n - n;
    int hot = n * (n + 1) / 2 + abs(mm);
    __syncthreads();
    for (int i = threadIdx.x; i < nc; i += blockDim.x)
      E[i] = i == hot ? (mm >= 0 ? make_double2(1, 0) : make_double2(0, 1)) : make_double2(0, 0);
    __syncthreads();
    for (int jks = threadIdx.x; jks < nc; jks += blockDim.x) {
      int j, k;
      unpack_nm(jks, j, k);
      double2 o = KIND == 1 ? m2m_entry(E, Y, j, k) : l2l_entry(E, Y, j, k, P);
      T[(size_t)col * pp + j * j + j + k] = o.x;
      if (k > 0) T[(size_t)col * pp + j * j + j - k] = o.y;
    }
  }
}

// ---- phase 1: C[rows x 64] = T_c * B, persistent and software pipelined --------------------------------
// A CTA owns a contiguous range of items (an item = one class x <=64 pairs).  T_c stays in shared
// memory while consecutive items share the class; the source expansions of the NEXT item are
// fetched with cp.async (16 B = one complex coefficient per copy, raw packed layout) while the
// current one is multiplied.  256 threads = 8 warps as 4 (rows) x 2 (columns); inside a warp lanes
// are 4 (rows) x 8 (columns); a thread owns RT x 4 outputs: rows wr*4RT + rg*RT + r, columns
// wc*32 + cg + 8j.  Shared-memory reads per k: RT doubles of T (4 distinct addresses per warp) and
// 4 doubles of B (8 distinct, conflict-free because the column stride is an odd multiple of 16 B).
// ACC = false: column of slot s goes to tmp[s][0..pp);  ACC = true: it is added to the packed
// complex expansion Out[s] (L2L: every slot is written exactly once).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Tile geometry per expansion order.  8 warps = WR (rows) x WC (columns); a warp owns RB row blocks
// and CB column blocks of 8x8 (the DMMA shape), CB = WR so that WC * CB * 8 = 64 columns.
template <int P>
struct GemmCfg {
  static constexpr int PP = P * P;
  static constexpr int XS = (PP + 1) & ~1;              // doubles per expansion in global memory
  static constexpr int KB = (PP + 3) / 4;               // k blocks of 4
  static constexpr int WR = PP > 32 ? 4 : (PP > 16 ? 2 : 1);
  static constexpr int RB = PP > 48 ? 2 : (PP > 32 ? 2 : (PP > 24 ? 3 : (PP > 8 ? 2 : 1)));
  static constexpr int WC = 8 / WR;
  static constexpr int CB = WR;
  static constexpr int KMAX = KB * 4 > XS ? KB * 4 : XS;
  // column stride: multiple of 2 (16-byte cp.async rows), congruent 4 mod 16 so that the 8 columns x 4 k
  // of one B fragment cover all 16 eight-byte banks exactly twice
  static constexpr int LDB = KMAX + ((4 - KMAX % 16) + 16) % 16;
  static_assert(WR * RB * 8 >= PP, "row tiles must cover the expansion");
};

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[rows x 64] = T_c * B with FP64 tensor-core MMAs (DMMA.8x8x4).  The T_c fragments of a warp live in
// REGISTERS for as long as consecutive items share the class; only the source expansions pass through
// shared memory (cp.async double buffer).
template <int P, bool ACC>
__global__ void __launch_bounds__(256, 2)
trans_gemm_kernel(int ldT, const double* __restrict__ Tt, int n_items, const int4* __restrict__ item_desc,
                  const int* __restrict__ sorted_slot, const int* __restrict__ sorted_src,
                  const double* __restrict__ X, double* __restrict__ tmp, double* __restrict__ Out) {
  using G = GemmCfg<P>;
  constexpr int PP = G::PP, XS = G::XS, KB = G::KB, RB = G::RB, CB = G::CB, LDB = G::LDB;
  // Software pipeline with ONE block barrier per item: source expansions are copied two items ahead into a ring
  // of three buffers, their slot / source indices are staged three items ahead into a ring of four.
  constexpr int NBUF = 3, NIDX = 4;
  extern __shared__ __align__(16) double smem[];        // NBUF x [kNB][LDB]
  __shared__ int s_slot[NIDX][kNB], s_src[NIDX][kNB];
  const int per = (n_items + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n_items, i0 + per);
  if (i0 >= i1) return;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = w % G::WR, wc = w / G::WR;              // wc < G::WC
  static_assert(G::WR * G::WC == 8, "8 warps");
  const int lr = lane >> 2, lk = lane & 3;
  const int row_base = wr * RB * 8 + lr;                 // + 8 rb
  const int col_base = wc * CB * 8;                      // + 8 cb
  constexpr int chunks = XS / 2;

  // k positions beyond the copied expansion (k-block padding) stay zero for the whole kernel
  if constexpr (G::KMAX > XS) {
    constexpr int extra = G::KMAX - XS;
    for (int idx = threadIdx.x; idx < NBUF * kNB * extra; idx += 256) {
      int colb = idx / extra, k = XS + idx % extra;
      smem[(size_t)colb * LDB + k] = 0.0;
    }
  }

  // Index pipeline without a dependent global load on the path of an item: the descriptor of item it + 4 is
  // fetched into a register while the slots / sources of item it + 3 (descriptor fetched one turn earlier) load
  // beside the MMAs of item it; they are stored to the ring behind the MMA loop and read after the next barrier.
  // (Before: class, count, slot and source of an item were loaded back to back in front of the MMA loop --
  // up to three dependent L2 round trips per item on the warps that stage the indices, one on every warp.)
  auto desc = [&](int it) { return it < i1 ? item_desc[it] : make_int4(-1, 0, 0, 0); };
  auto stage_indices = [&](int it, int buf) {            // prologue only (blocking)
    if (threadIdx.x < kNB) {
      const int4 d = item_desc[it];
      const bool in = threadIdx.x < d.z;
      s_slot[buf][threadIdx.x] = in ? sorted_slot[d.y + threadIdx.x] : -1;
      s_src[buf][threadIdx.x] = in ? sorted_src[d.y + threadIdx.x] : -1;
    }
  };
  auto stage_copy = [&](int ib, int buf) {
    double* Bs = smem + (size_t)buf * kNB * LDB;
    for (int idx = threadIdx.x; idx < kNB * chunks; idx += 256) {
      int col = idx / chunks, ch = idx - col * chunks;
      int src = s_src[ib][col];
      double* dst = Bs + col * LDB + 2 * ch;
      if (src >= 0) cp_async16(dst, X + (size_t)src * XS + 2 * ch);
      else { dst[0] = 0.0; dst[1] = 0.0; }
    }
    cp_async_commit();
  };

  double A[RB][KB];
  int cur = -1;
  stage_indices(i0, 0);
  if (i0 + 1 < i1) stage_indices(i0 + 1, 1);
  if (i0 + 2 < i1) stage_indices(i0 + 2, 2);
  __syncthreads();
  stage_copy(0, 0);
  if (i0 + 1 < i1) stage_copy(1, 1); else cp_async_commit();
  int cls0 = desc(i0).x, cls1 = desc(i0 + 1).x, cls2 = desc(i0 + 2).x;
  int4 d3 = desc(i0 + 3);
  for (int it = i0; it < i1; ++it) {
    const int j = it - i0, buf = j % NBUF, ib = j % NIDX;
    const int c = cls0;
    const int4 d4 = desc(it + 4);
    int sl_r = -1, src_r = -1;
    if (threadIdx.x < kNB && (int)threadIdx.x < d3.z) {   // d3.z = 0 behind the last item
      sl_r = sorted_slot[d3.y + threadIdx.x];
      src_r = sorted_src[d3.y + threadIdx.x];
    }
    if (c != cur) {
      const double* Tc = Tt + (size_t)c * ldT * ldT;
#pragma unroll
      for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          int row = row_base + 8 * rb, k = kb * 4 + lk;
          A[rb][kb] = (row < PP && k < PP) ? Tc[(size_t)k * ldT + row] : 0.0;
        }
      cur = c;
    }
    cp_async_wait<1>();                 // this thread's copies of item `it` have landed (item it+1 may be in flight)
    __syncthreads();                    // ... everyone's; all warps are done with item it-1
    if (it + 2 < i1) stage_copy((j + 2) % NIDX, (j + 2) % NBUF); else cp_async_commit();

    const double* Bs = smem + (size_t)buf * kNB * LDB + (size_t)(col_base + lr) * LDB + lk;
    double C[RB][CB][2];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) C[rb][cb][0] = C[rb][cb][1] = 0.0;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      double bf[CB];
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) bf[cb] = Bs[cb * 8 * LDB + kb * 4];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) dmma8x8x4(C[rb][cb][0], C[rb][cb][1], A[rb][kb], bf[cb]);
    }
    if (threadIdx.x < kNB && it + 3 < i1) {               // ring slot of item it - 1: everyone is past its epilogue
      s_slot[(j + 3) % NIDX][threadIdx.x] = sl_r;
      s_src[(j + 3) % NIDX][threadIdx.x] = src_r;
    }
    cls0 = cls1; cls1 = cls2; cls2 = d3.x; d3 = d4;
#pragma unroll
    for (int cb = 0; cb < CB; ++cb)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int sl = s_slot[ib][col_base + cb * 8 + 2 * lk + i];
        if (sl < 0) continue;
        double* o = (ACC ? Out : tmp) + (size_t)sl * XS;
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          int row = row_base + 8 * rb;
          if (row < PP) {
            if (ACC) o[row] += C[rb][cb][i];
            else o[row] = C[rb][cb][i];
          }
        }
      }
  }
}

// ---- phase 2 (M2L): L[t] = sum of its columns -------------------------------------------------------
__global__ void __launch_bounds__(256)
m2l_reduce_kernel(int nboxes, const int* __restrict__ off, const unsigned char* __restrict__ batched, int P,
                  const double* __restrict__ tmp, double* __restrict__ L) {
  int b = blockIdx.x;
  if (b >= nboxes) return;
  const int pp = P * P, xs = xstride(P);
  int e0 = off[b], e1 = off[b + 1];
  for (int row = threadIdx.x; row < pp; row += blockDim.x) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int e = e0;
    for (; e + 4 <= e1; e += 4) {
      double v0 = batched[e] ? tmp[(size_t)e * xs + row] : 0.0;
      double v1 = batched[e + 1] ? tmp[(size_t)(e + 1) * xs + row] : 0.0;
      double v2 = batched[e + 2] ? tmp[(size_t)(e + 2) * xs + row] : 0.0;
      double v3 = batched[e + 3] ? tmp[(size_t)(e + 3) * xs + row] : 0.0;
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; e < e1; ++e)
      if (batched[e]) s0 += tmp[(size_t)e * xs + row];
    L[(size_t)b * xs + row] = (s0 + s1) + (s2 + s3);
  }
}

// The same sums (same order of additions, same bits) with a footprint that leaves most of an SM to the kernel running
// beside it.  m2l_reduce_kernel keeps its bytes in flight in registers and takes all 32 block slots of an SM with its
// 64-thread blocks: while it runs (0.32 ms at N = 1M, HBM-bound), the near field on the other stream -- one-warp
// blocks that need the FP64 pipe the reduction leaves idle -- is shut out.  Here the bytes in flight live in SHARED
// memory: every warp owns a ring of kRedStages 4 KB stages, one lane brings the next columns of the warp's box in by
// 1-D TMA bulk copies (cp.async.bulk, completion in bytes on the stage's mbarrier; the columns of a box are
// contiguous), the warp adds them up out of shared memory (lane = two rows).  Two blocks of four warps per SM keep
// 96 KB per SM in flight with 15 K registers (58 per thread); warps never meet at a block barrier.
constexpr int kRedStages = 4, kRedStageBytes = 4096, kRedWarps = 4;
constexpr size_t kRedShared = (size_t)kRedWarps * kRedStages * (kRedStageBytes + sizeof(uint64_t) + sizeof(unsigned));

__device__ __forceinline__ void red_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void red_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void red_mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
      " bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void red_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
               "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__global__ void __launch_bounds__(32 * kRedWarps)
m2l_reduce_tma_kernel(int nboxes, const int* __restrict__ off, const unsigned char* __restrict__ batched, int P,
                      const double* __restrict__ tmp, double* __restrict__ L) {
  extern __shared__ __align__(128) unsigned char red_sh[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* data = red_sh + (size_t)w * kRedStages * kRedStageBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(red_sh + (size_t)kRedWarps * kRedStages * kRedStageBytes) + w * kRedStages;
  unsigned* cmask = reinterpret_cast<unsigned*>(red_sh + (size_t)kRedWarps * kRedStages * (kRedStageBytes + sizeof(uint64_t))) +
                    w * kRedStages;
  if (lane == 0)
    for (int s = 0; s < kRedStages; ++s) red_mbar_init(&bar[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int xs = xstride(P), nv = xs >> 1;            // xs is even: a column is nv double2
  const unsigned colb = (unsigned)xs * sizeof(double);
  // columns per chunk: a multiple of 4 (the four partial sums of m2l_reduce_kernel then line up with the chunks),
  // at most 32 (one flag per lane)
  const int CH = min(32, (int)(kRedStageBytes / colb) & ~3);
  const int nw = gridDim.x * kRedWarps;
  // a cursor walks the chunks of the warp's boxes (b = first, first + nw, ...); a box without columns is one empty chunk
  struct Cur { int b, e, e1; };
  auto load_box = [&](Cur& c) { if (c.b < nboxes) { c.e = off[c.b]; c.e1 = off[c.b + 1]; } };
  auto step = [&](Cur& c, int n) { c.e += n; if (c.e >= c.e1) { c.b += nw; load_box(c); } };
  auto issue = [&](Cur& c, int st) {
    const int n = min(CH, c.e1 - c.e);
    if (n > 0) {
      const unsigned m = __ballot_sync(0xffffffffu, lane < n && batched[c.e + lane] != 0);
      if (lane == 0) {
        cmask[st] = m;
        red_mbar_expect_tx(&bar[st], (unsigned)n * colb);
        red_bulk_g2s(data + (size_t)st * kRedStageBytes, tmp + (size_t)c.e * xs, (unsigned)n * colb, &bar[st]);
      }
    } else if (lane == 0) {
      red_mbar_arrive(&bar[st]);                       // keeps the phase of the stage in step with the chunk count
    }
    step(c, n);
  };
  Cur pc{(int)(blockIdx.x * kRedWarps) + w, 0, 0};
  load_box(pc);
  Cur cc = pc;
  int pi = 0, ci = 0;
  while (pi < kRedStages - 1 && pc.b < nboxes) { issue(pc, pi % kRedStages); ++pi; }
  __syncwarp();
  const double2 zero = make_double2(0.0, 0.0);
  double2 s0 = zero, s1 = zero, s2 = zero, s3 = zero;
  while (cc.b < nboxes) {
    // the stage of chunk ci - 1 was released by the warp barrier that ended the previous turn
    if (pc.b < nboxes) { issue(pc, pi % kRedStages); ++pi; }
    const int n = min(CH, cc.e1 - cc.e);
    const int st = ci % kRedStages;
    red_mbar_wait(&bar[st], (unsigned)(ci / kRedStages) & 1u);
    if (n > 0 && lane < nv) {
      const unsigned m = cmask[st];
      const double2* col = reinterpret_cast<const double2*>(data + (size_t)st * kRedStageBytes) + lane;
      int j = 0;
#pragma unroll 2
      for (; j + 4 <= n; j += 4) {
        const double2 v0 = (m >> j) & 1u ? col[(size_t)j * nv] : zero;
        const double2 v1 = (m >> (j + 1)) & 1u ? col[(size_t)(j + 1) * nv] : zero;
        const double2 v2 = (m >> (j + 2)) & 1u ? col[(size_t)(j + 2) * nv] : zero;
        const double2 v3 = (m >> (j + 3)) & 1u ? col[(size_t)(j + 3) * nv] : zero;
        s0.x += v0.x; s0.y += v0.y; s1.x += v1.x; s1.y += v1.y;
        s2.x += v2.x; s2.y += v2.y; s3.x += v3.x; s3.y += v3.y;
      }
      for (; j < n; ++j)                               // only the last chunk of a box has a remainder
        if ((m >> j) & 1u) { const double2 v = col[(size_t)j * nv]; s0.x += v.x; s0.y += v.y; }
    }
    if (cc.e + n >= cc.e1) {                           // last chunk of the box
      if (lane < nv) {
        // the padding double of an odd-sized expansion stays zero (the scratch columns do not define theirs)
        const bool pad = 2 * lane + 1 >= P * P;
        reinterpret_cast<double2*>(L)[(size_t)cc.b * nv + lane] =
            make_double2((s0.x + s1.x) + (s2.x + s3.x), pad ? 0.0 : (s0.y + s1.y) + (s2.y + s3.y));
      }
      s0 = zero; s1 = zero; s2 = zero; s3 = zero;
    }
    __syncwarp();                                      // every lane is done with stage st before it is refilled
    step(cc, n);
    ++ci;
  }
}

// ---- phase 2 (M2M): M[parent] = sum over its children's columns, in child order ---------------------
__global__ void __launch_bounds__(64)
m2m_reduce_kernel(int lo, int hi, const unsigned* __restrict__ key, const unsigned* __restrict__ cbegin,
                  const unsigned* __restrict__ cend, const unsigned char* __restrict__ mask, int P,
                  const double* __restrict__ tmp, double* __restrict__ M) {
  int b = lo + blockIdx.x;
  if (b >= hi || (key[b] >> 31) || (mask && !mask[b])) return;
  const int pp = P * P, xs = xstride(P);
  unsigned c0 = cbegin[b], c1 = cend[b];
  for (int row = threadIdx.x; row < pp; row += blockDim.x) {
    double s = 0;
    for (unsigned c = c0; c < c1; ++c) s += tmp[(size_t)c * xs + row];
    M[(size_t)b * xs + row] = s;
  }
}

template <int P, bool ACC>
void launch_gemm_t(const TransBatch& B, int first, int count, const double* X, double* tmp, double* out,
                   cudaStream_t s) {
  size_t sh = (size_t)3 * kNB * GemmCfg<P>::LDB * sizeof(double);
  // per call, not once per process: function attributes and the SM count belong to the CURRENT device
  // (a process may hold plans on several devices); both calls are host-side and cheap
  int sms = 0, dev = 0;
  FMMB_CUDA(cudaFuncSetAttribute(trans_gemm_kernel<P, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  FMMB_CUDA(cudaGetDevice(&dev));
  FMMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = std::min(count, 2 * sms);
  trans_gemm_kernel<P, ACC><<<grid, 256, sh, s>>>(B.built_p * B.built_p, B.T.p, count, B.item_desc.p + first,
                                                 B.sorted_slot.p, B.sorted_src.p, X, tmp, out);
  FMMB_CUDA(cudaGetLastError());
}
template <bool ACC>
void launch_gemm(const TransBatch& B, int P, int first, int count, const double* X, double* tmp, double* out,
                 cudaStream_t s) {
  if (count <= 0) return;
  switch (P) {
    case 1: launch_gemm_t<1, ACC>(B, first, count, X, tmp, out, s); break;
    case 2: launch_gemm_t<2, ACC>(B, first, count, X, tmp, out, s); break;
    case 3: launch_gemm_t<3, ACC>(B, first, count, X, tmp, out, s); break;
    case 4: launch_gemm_t<4, ACC>(B, first, count, X, tmp, out, s); break;
    case 5: launch_gemm_t<5, ACC>(B, first, count, X, tmp, out, s); break;
    case 6: launch_gemm_t<6, ACC>(B, first, count, X, tmp, out, s); break;
    case 7: launch_gemm_t<7, ACC>(B, first, count, X, tmp, out, s); break;
    default: launch_gemm_t<8, ACC>(B, first, count, X, tmp, out, s); break;
  }
}

// Sorts the pairs of a batch by class, builds classes / items / (for M2L) the residual lists.
// tgt/src are indexed by slot; slots [first, first+n) take part.
void classify(fmmb_plan* plan, TransBatch& B, const int* tgt, const int* src, int first, const int* slot_list,
              int64_t n, int minpop, bool by_level) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  Temp tmp;
  DevBuf<unsigned long long> k0, k1, uniq;
  DevBuf<int> v0, count, nruns;
  k0.resize(n); k1.resize(n); v0.resize(n); B.sorted_slot.resize(n);
  class_keys<<<nblk(n, 256), 256, 0, s>>>(tgt, src, first, slot_list, n, T.key.p, T.level.p, by_level ? 1 : 0, k0.p,
                                          v0.p);
  FMMB_CUDA(cudaGetLastError());
  {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, v0.p, B.sorted_slot.p, n, 0, 40, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, k0.p, k1.p, v0.p, B.sorted_slot.p, n, 0, 40, s));
  }
  uniq.resize(n); count.resize(n + 1); nruns.resize(1);
  {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, bytes, k1.p, uniq.p, count.p, nruns.p, n, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceRunLengthEncode::Encode(t, bytes, k1.p, uniq.p, count.p, nruns.p, n, s));
  }
  int ncls = nruns.to_host(s)[0];
  DevBuf<int> start, nitems, item_off;
  start.resize(ncls + 1); nitems.resize(ncls + 1); item_off.resize(ncls + 1);
  {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, count.p, start.p, ncls + 1, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, count.p, start.p, ncls + 1, s));
    class_items<<<nblk(ncls + 1, 256), 256, 0, s>>>(count.p, ncls, minpop, nitems.p);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, nitems.p, item_off.p, ncls + 1, s));
    t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, nitems.p, item_off.p, ncls + 1, s));
  }
  int n_items = 0;
  FMMB_CUDA(cudaMemcpyAsync(&n_items, item_off.p + ncls, sizeof(int), cudaMemcpyDeviceToHost, s));
  FMMB_CUDA(cudaStreamSynchronize(s));
  B.n_classes = ncls;
  B.n_items = n_items;
  B.item_class.resize(n_items); B.item_start.resize(n_items); B.item_count.resize(n_items);
  B.item_desc.resize(n_items);
  if (n_items)
    fill_items<<<nblk(ncls, 128), 128, 0, s>>>(count.p, start.p, item_off.p, ncls, B.item_class.p, B.item_start.p,
                                              B.item_count.p, B.item_desc.p);
  B.sorted_src.resize(n);
  if (n) gather_sorted_src<<<nblk(n, 256), 256, 0, s>>>(B.sorted_slot.p, src, n, B.sorted_src.p);
  B.class_vec.resize(ncls);
  class_vectors<<<nblk(ncls, 128), 128, 0, s>>>(start.p, ncls, B.sorted_slot.p, tgt, src, T.key.p, T.level.p,
                                               make_double3(0.5 * T.cell[0], 0.5 * T.cell[1], 0.5 * T.cell[2]),
                                               B.class_vec.p);
  FMMB_CUDA(cudaGetLastError());
  if (by_level) {
    // items are sorted by target level (top bits of the class key): record the ranges on the host
    B.level_item_off.assign(T.nlevels + 1, 0);
    std::vector<unsigned long long> hk(ncls);
    FMMB_CUDA(cudaMemcpyAsync(hk.data(), uniq.p, ncls * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    std::vector<int> hoff = item_off.to_host(s);
    for (int c = 0; c < ncls; ++c) {
      int l = (int)(hk[c] >> 36);
      if (l + 1 <= T.nlevels) B.level_item_off[l + 1] = hoff[c + 1];
    }
    for (int l = 1; l <= T.nlevels; ++l) B.level_item_off[l] = std::max(B.level_item_off[l], B.level_item_off[l - 1]);
  } else {
    // M2L: slots of small classes stay on the per-pair kernel
    B.batched.resize(n);
    mark_batched<<<ncls, 128, 0, s>>>(count.p, start.p, ncls, minpop, B.sorted_slot.p, B.batched.p);
    DevBuf<int> flag, pos;
    flag.resize(n + 1); pos.resize(n + 1);
    residual_flags<<<nblk(n + 1, 256), 256, 0, s>>>(B.batched.p, n, flag.p);
    {
      size_t bytes = 0;
      FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, pos.p, n + 1, s));
      void* t = tmp.get(bytes);
      FMMB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, flag.p, pos.p, n + 1, s));
    }
    int n_res = 0;
    FMMB_CUDA(cudaMemcpyAsync(&n_res, pos.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    B.n_res = n_res;
    B.n_pairs = n - n_res;
    B.res_src.resize(n_res); B.res_off.resize(T.nboxes + 1);
    if (n_res) residual_compact<<<nblk(n, 256), 256, 0, s>>>(flag.p, pos.p, n, src, B.res_src.p);
    residual_offsets<<<nblk(T.nboxes + 1, 256), 256, 0, s>>>(T.m2l_off.p, pos.p, T.nboxes, B.res_off.p);
    {
      // boxes that own residual pairs (few): the per-pair kernel is launched over these only
      std::vector<int> ho = B.res_off.to_host(s), list;
      for (int b = 0; b < T.nboxes; ++b) if (ho[b + 1] > ho[b]) list.push_back(b);
      B.n_res_boxes = (int)list.size();
      B.res_boxes.from_host(list.data(), list.size(), s);
    }
  }
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));
}

void ensure_T(fmmb_plan* plan, TransBatch& B, cudaStream_t s) {
  if (B.built_p >= plan->p || B.n_classes == 0) return;
  const int bp = 8;   // largest order of the GEMM family; smaller orders use the leading sub-block
  B.T.resize((size_t)B.n_classes * bp * bp * bp * bp);
  if (B.kind == 0)
    build_T_m2l<<<(int)B.n_classes, 256, (size_t)4 * bp * bp * sizeof(double2), s>>>(bp, B.class_vec.p, B.T.p);
  else if (B.kind == 1)
    build_T_probe<1><<<(int)B.n_classes, 128, (size_t)(bp * bp + bp * (bp + 1) / 2) * sizeof(double2), s>>>(
        bp, B.class_vec.p, B.T.p);
  else
    build_T_probe<2><<<(int)B.n_classes, 128, (size_t)(bp * bp + bp * (bp + 1) / 2) * sizeof(double2), s>>>(
        bp, B.class_vec.p, B.T.p);
  FMMB_CUDA(cudaGetLastError());
  B.built_p = bp;
  ++plan->launches;
}

}  // namespace

void m2l_init_tables() { upload_laplace_tables(); }

// Plan-time: classify M2L, M2M and L2L translations and build the GEMM work items.
void build_m2l_classes(fmmb_plan* plan) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  const int nb = T.nboxes;
  TransBatch& C = plan->cls;
  C.kind = 0; C.n_classes = 0; C.n_pairs = 0; C.n_items = 0; C.n_res = T.n_lr_local; C.built_p = 0;
  plan->m2m.kind = 1; plan->m2m.n_items = 0; plan->m2m.built_p = 0;
  plan->l2l.kind = 2; plan->l2l.n_items = 0; plan->l2l.built_p = 0;
  if (plan->opts.m2l_mode == 1) return;
  if (T.n_lr >= (1ll << 31)) throw StatusError{FMMB_ERR_INVALID, "more than 2^31 M2L pairs"};
  if (T.n_lr_local > 0) {
    C.slot_tgt.resize(T.n_lr_local);
    slot_targets<<<nb, 64, 0, s>>>(T.m2l_off.p, nb, C.slot_tgt.p);
    C.slot_src_p = T.m2l_src.p;
    classify(plan, C, C.slot_tgt.p, T.m2l_src.p, 0, nullptr, T.n_lr_local, kMinPopM2L, false);
  }
  const std::vector<unsigned char>& need = T.need_M_host;     // parents whose multipole a matvec reads
  std::vector<unsigned> par;
  if (nb > 1) par = T.parent.to_host(s);
  if (nb > 1 && T.nranks > 1) {
    // M2M restricted to parents whose bodies all belong to this rank (owned upward pass)
    TransBatch& B = plan->m2m_own;
    B.kind = 1; B.n_items = 0; B.built_p = 0;
    B.slot_tgt.resize(nb); B.slot_src.resize(nb);
    parent_child_pairs<<<nblk(nb, 256), 256, 0, s>>>(T.parent.p, nb, 0, B.slot_tgt.p, B.slot_src.p);
    B.slot_src_p = B.slot_src.p;
    std::vector<unsigned char> ins = T.up_inside.to_host(s), mask(nb, 0);
    std::vector<int> list;
    for (int c = 1; c < nb; ++c) if (ins[par[c]] && need[par[c]]) { list.push_back(c); mask[par[c]] = 1; }
    T.m2m_mask_own.from_host(mask.data(), mask.size(), s);
    DevBuf<int> dl;
    dl.from_host(list.data(), list.size(), s);
    if (!list.empty()) classify(plan, B, B.slot_tgt.p, B.slot_src.p, 0, dl.p, (int64_t)list.size(), 1, true);
    else B.level_item_off.assign(T.nlevels + 1, 0);
    B.n_pairs = (int64_t)list.size();
  }
  if (nb > 1) {
    for (int kind = 1; kind <= 2; ++kind) {
      TransBatch& B = kind == 1 ? plan->m2m : plan->l2l;
      B.slot_tgt.resize(nb); B.slot_src.resize(nb);
      parent_child_pairs<<<nblk(nb, 256), 256, 0, s>>>(T.parent.p, nb, kind == 2, B.slot_tgt.p, B.slot_src.p);
      B.slot_src_p = B.slot_src.p;
      std::vector<int> list;
      if (kind == 1) {
        // M2M only into parents whose multipole is read (M2L sources and what lies below them)
        std::vector<unsigned char> mask(nb, 0);
        for (int c = 1; c < nb; ++c) if (need[par[c]]) { list.push_back(c); mask[par[c]] = 1; }
        T.m2m_mask_all.from_host(mask.data(), mask.size(), s);
      } else {
        // L2L only into boxes that are targets on this rank, from parents that carry a local expansion
        std::vector<unsigned char> act = T.active.to_host(s), hl = T.has_local.to_host(s);
        for (int c = 1; c < nb; ++c) if (act[c] && hl[par[c]]) list.push_back(c);
      }
      DevBuf<int> dl;
      dl.from_host(list.data(), list.size(), s);
      if (!list.empty()) classify(plan, B, B.slot_tgt.p, B.slot_src.p, 0, dl.p, (int64_t)list.size(), 1, true);
      else { B.level_item_off.assign(T.nlevels + 1, 0); B.n_items = 0; }
      B.n_pairs = (int64_t)list.size();
      FMMB_CUDA(cudaStreamSynchronize(s));
    }
  }
}

// Batched far field for the current order.  Returns false if there is nothing batched
// (the caller then runs the per-pair kernel over the full lists).
bool m2l_batched(fmmb_plan* plan, cudaStream_t s) {
  TransBatch& C = plan->cls;
  Tree& T = plan->tree;
  const int P = plan->p, pp = P * P, xs = xstride(P);
  if (C.n_items == 0 || P > 8) return false;
  ensure_T(plan, C, s);
  C.tmp.resize((size_t)T.n_lr_local * xs);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(plan->ev[13], s));
  launch_gemm<false>(C, P, 0, C.n_items, plan->M.p, C.tmp.p, nullptr, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(plan->ev[14], s));
  plan->m2l_gemm_timed = true;
  ++plan->launches;
  if (plan->hook_after_m2l_gemm) plan->hook_after_m2l_gemm();   // laplace_execute: the near field starts here
  if (plan->m2l_reduce == 1) {
    int sms = 148;
    FMMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device));
    FMMB_CUDA(cudaFuncSetAttribute(m2l_reduce_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRedShared));
    const int grid = std::min(nblk(T.nboxes, kRedWarps), plan->m2l_reduce_bps * sms);
    m2l_reduce_tma_kernel<<<grid, 32 * kRedWarps, kRedShared, s>>>(T.nboxes, T.m2l_off.p, C.batched.p, P, C.tmp.p, plan->L.p);
  } else {
    int threads = pp < 64 ? 64 : (pp > 256 ? 256 : pp);
    m2l_reduce_kernel<<<T.nboxes, threads, 0, s>>>(T.nboxes, T.m2l_off.p, C.batched.p, P, C.tmp.p, plan->L.p);
  }
  FMMB_CUDA(cudaGetLastError());
  ++plan->launches;
  return true;
}

// Batched M2M level sweep (finest parents first).  False -> caller uses the per-box kernels.
bool m2m_batched(fmmb_plan* plan, cudaStream_t s, bool owned_only) {
  TransBatch& B = owned_only ? plan->m2m_own : plan->m2m;
  Tree& T = plan->tree;
  const int P = plan->p, xs = xstride(P);
  if (B.n_items == 0 || P > 8 || plan->opts.m2l_mode == 1) return false;
  ensure_T(plan, B, s);
  B.tmp.resize((size_t)T.nboxes * xs);
  for (int l = T.nlevels - 2; l >= 0; --l) {
    int i0 = B.level_item_off[l], i1 = B.level_item_off[l + 1];
    if (i1 <= i0) continue;
    launch_gemm<false>(B, P, i0, i1 - i0, plan->M.p, B.tmp.p, nullptr, s);
    int lo = T.level_off[l], hi = T.level_off[l + 1];
    m2m_reduce_kernel<<<hi - lo, 64, 0, s>>>(
        lo, hi, T.key.p, T.cbegin.p, T.cend.p, owned_only ? T.m2m_mask_own.p : T.m2m_mask_all.p, P, B.tmp.p, plan->M.p);
    plan->launches += 2;
  }
  FMMB_CUDA(cudaGetLastError());
  return true;
}

// Batched L2L level sweep (coarsest children first); adds into the locals M2L left behind.
bool l2l_batched(fmmb_plan* plan, cudaStream_t s) {
  TransBatch& B = plan->l2l;
  Tree& T = plan->tree;
  const int P = plan->p;
  if (B.n_items == 0 || P > 8 || plan->opts.m2l_mode == 1) return false;
  ensure_T(plan, B, s);
  for (int l = 1; l < T.nlevels; ++l) {
    int i0 = B.level_item_off[l], i1 = B.level_item_off[l + 1];
    if (i1 <= i0) continue;
    launch_gemm<true>(B, P, i0, i1 - i0, plan->L.p, nullptr, plan->L.p, s);
    ++plan->launches;
  }
  return true;
}

}  // namespace fmmb
