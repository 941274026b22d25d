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
#include "laplace_tables.cuh"
#include <cub/cub.cuh>

namespace fmmb {

namespace {

constexpr int kNB = 128;          // pairs (columns) per CTA
constexpr int kMinPop = 24;       // smallest class that is worth a GEMM tile

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

__global__ void class_keys(const int* __restrict__ tgt, const int* __restrict__ src, int64_t n,
                           const unsigned* __restrict__ key, const unsigned* __restrict__ lvl,
                           unsigned long long* __restrict__ ckey, int* __restrict__ slot) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  int t = tgt[e], s = src[e];
  int3 a = centre_half_cells(key[t], lvl[t]), b = centre_half_cells(key[s], lvl[s]);
  unsigned long long dx = (unsigned)(a.x - b.x + 2048), dy = (unsigned)(a.y - b.y + 2048),
                     dz = (unsigned)(a.z - b.z + 2048);
  ckey[e] = (dx << 24) | (dy << 12) | dz;
  slot[e] = (int)e;
}

// per class: number of GEMM items (0 if the class is too small)
__global__ void class_items(const int* __restrict__ count, int nclasses, int* __restrict__ nitems) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nclasses) return;
  nitems[c] = (c < nclasses && count[c] >= kMinPop) ? (count[c] + kNB - 1) / kNB : 0;
}
__global__ void fill_items(const int* __restrict__ count, const int* __restrict__ start,
                           const int* __restrict__ item_off, int nclasses, int* __restrict__ item_class,
                           int* __restrict__ item_start, int* __restrict__ item_count) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nclasses) return;
  int ni = item_off[c + 1] - item_off[c];
  for (int i = 0; i < ni; ++i) {
    int it = item_off[c] + i;
    item_class[it] = c;
    item_start[it] = start[c] + i * kNB;
    item_count[it] = min(kNB, count[c] - i * kNB);
  }
}
// mark slots that the batched path covers; the rest go to the per-pair kernel
__global__ void mark_batched(const int* __restrict__ count, const int* __restrict__ start, int nclasses,
                             const int* __restrict__ sorted_slot, unsigned char* __restrict__ batched) {
  int c = blockIdx.x;
  if (c >= nclasses) return;
  unsigned char v = count[c] >= kMinPop;
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
__global__ void class_vectors(const int* __restrict__ start, int nclasses, const int* __restrict__ sorted_slot,
                              const int* __restrict__ tgt, const int* __restrict__ src,
                              const double4* __restrict__ center, double4* __restrict__ vec) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nclasses) return;
  int e = sorted_slot[start[c]];
  double4 a = center[tgt[e]], b = center[src[e]];
  vec[c] = make_double4(a.x - b.x, a.y - b.y, a.z - b.z, 0.0);
}

// ---- translation matrices ------------------------------------------------------------------------
struct Sph { double r, x, y, cp, sp; };
__device__ __forceinline__ Sph to_sph(double dx, double dy, double dz) {
  Sph s;
  s.r = sqrt(dx * dx + dy * dy + dz * dz) + 1e-12;
  s.x = __ddiv_rn(dz, s.r);
  s.y = sqrt((1.0 - s.x) * (1.0 + s.x));
  double ax = fabs(dx), ay = fabs(dy);
  if (ax + ay < 1e-12) { s.cp = 1.0; s.sp = 0.0; }
  else if (ax < 1e-12) { s.cp = 0.0; s.sp = dy > 0 ? 1.0 : -1.0; }
  else { double h = sqrt(dx * dx + dy * dy); s.cp = dx / h; s.sp = dy / h; }
  return s;
}
// singular harmonics rho^{-n-1} Y_n^m, n < top, column m (evalLocal, LaplaceSpherical.hpp:491-524)
__device__ void local_column(int m, int top, const Sph& s, double2* Y) {
  double pn = 1, fact = 1, er = 1, ei = 0, rhom = 1.0 / s.r;
  for (int k = 0; k < m; ++k) {
    pn = -pn * fact * s.y; fact += 2;
    double t = er * s.cp - ei * s.sp; ei = er * s.sp + ei * s.cp; er = t;
    rhom /= s.r;
  }
  double p = pn;
  int npn = m * m + 2 * m, nmn = m * m;
  double a = rhom * p * c_pref[npn];
  Y[npn] = make_double2(a * er, a * ei);
  Y[nmn] = make_double2(a * er, -a * ei);
  double p1 = p;
  p = s.x * (2 * m + 1) * p1;
  rhom /= s.r;
  double rhon = rhom;
  for (int n = m + 1; n < top; ++n) {
    int npm = n * n + n + m, nmm = n * n + n - m;
    a = rhon * p * c_pref[npm];
    Y[npm] = make_double2(a * er, a * ei);
    Y[nmm] = make_double2(a * er, -a * ei);
    double p2 = p1; p1 = p;
    p = (s.x * (2 * n + 1) * p1 - (n + m) * p2) / (n - m + 1);
    rhon /= s.r;
  }
}
__device__ __forceinline__ double cnm_real(int j, int k, int n, int m) {
  int e = abs(k - m) - abs(k) - abs(m);
  double sgn = ((e / 2) & 1) ? -1.0 : 1.0;
  double oj = (j & 1) ? -1.0 : 1.0;
  return sgn * oj * c_anm[n * n + n + m] * c_anm[j * j + j + k] / c_anm[(j + n) * (j + n) + j + n + m - k];
}
// Tt[c][col][row] (column-major in the GEMM sense: k-major, rows contiguous), ld = P^2
__global__ void __launch_bounds__(256)
build_T(int P, const double4* __restrict__ vec, double* __restrict__ Tt) {
  extern __shared__ double2 Y[];       // (2P)^2
  const int pp = P * P, c = blockIdx.x;
  double4 v = vec[c];
  Sph s = to_sph(v.x, v.y, v.z);
  for (int m = threadIdx.x; m < 2 * P; m += blockDim.x) local_column(m, 2 * P, s, Y);
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
    // W(+m) and W(-m) = Cnm * Y_{j+n}^{+-m-k}
    double cp_ = cnm_real(j, k, n, m);
    double2 yp = Y[base + m];
    double wpr = cp_ * yp.x, wpi = cp_ * yp.y;
    double val;
    if (m == 0) {
      val = (mm < 0) ? 0.0 : (kk >= 0 ? wpr : wpi);
    } else {
      double cm_ = cnm_real(j, k, n, -m);
      double2 ym = Y[base - m];
      double wmr = cm_ * ym.x, wmi = cm_ * ym.y;
      if (kk >= 0) val = (mm >= 0) ? (wpr + wmr) : -(wpi - wmi);
      else val = (mm >= 0) ? (wpi + wmi) : (wpr - wmr);
    }
    if (kk < 0 && k == 0) val = 0.0;
    T[(size_t)col * pp + row] = val;
  }
}

// ---- phase 1: C[rows x 128] = T_c * B -------------------------------------------------------------
// 256 threads: warp w owns rows [w*RPT, (w+1)*RPT), lane l owns columns l, l+32, l+64, l+96.
template <int RPT>
__global__ void __launch_bounds__(256, 2)
m2l_gemm_kernel(int P, int ldT, const double* __restrict__ Tt, const int* __restrict__ item_class,
                const int* __restrict__ item_start, const int* __restrict__ item_count,
                const int* __restrict__ sorted_slot, const int* __restrict__ slot_src,
                const double2* __restrict__ M, double* __restrict__ tmp) {
  constexpr int ROWS = 8 * RPT;
  const int pp = P * P, nc = P * (P + 1) / 2, ldb = pp + 1;
  extern __shared__ double smem[];
  double* Ts = smem;                       // [pp][ROWS]
  double* Bs = smem + (size_t)pp * ROWS;   // [kNB][ldb]
  __shared__ int s_slot[kNB];
  const int item = blockIdx.x;
  const int c = item_class[item], start = item_start[item], cnt = item_count[item];
  const double* Tc = Tt + (size_t)c * ldT * ldT;
  for (int idx = threadIdx.x; idx < pp * ROWS; idx += 256) {
    int k = idx / ROWS, row = idx % ROWS;
    Ts[idx] = row < pp ? Tc[(size_t)k * ldT + row] : 0.0;
  }
  if (threadIdx.x < kNB) s_slot[threadIdx.x] = threadIdx.x < cnt ? sorted_slot[start + threadIdx.x] : -1;
  __syncthreads();
  // gather: thread per (column, packed coefficient)
  for (int idx = threadIdx.x; idx < kNB * nc; idx += 256) {
    int col = idx / nc, nms = idx % nc;
    int n = 0; while ((n + 1) * (n + 2) / 2 <= nms) ++n;
    int m = nms - n * (n + 1) / 2;
    double2 v = make_double2(0, 0);
    int sl = s_slot[col];
    if (sl >= 0) v = M[(size_t)slot_src[sl] * nc + nms];
    Bs[col * ldb + n * n + n + m] = v.x;
    if (m > 0) Bs[col * ldb + n * n + n - m] = v.y;
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[RPT][4];
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[r][j] = 0.0;
  const double* a_ptr = Ts + w * RPT;
  const double* b_ptr = Bs + lane * ldb;
#pragma unroll 4
  for (int k = 0; k < pp; ++k) {
    double a[RPT], b[4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) a[r] = a_ptr[k * ROWS + r];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = b_ptr[j * 32 * ldb + k];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = fma(a[r], b[j], acc[r][j]);
  }
  __syncthreads();
  // stage through shared memory (reuse Bs as [col][ldb]) for coalesced column stores
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      int row = w * RPT + r;
      if (row < pp) Bs[(lane + 32 * j) * ldb + row] = acc[r][j];
    }
  __syncthreads();
  for (int idx = threadIdx.x; idx < cnt * pp; idx += 256) {
    int col = idx / pp, row = idx % pp;
    tmp[(size_t)s_slot[col] * pp + row] = Bs[col * ldb + row];
  }
}

// ---- phase 2: L[t] = sum of its columns -------------------------------------------------------------
__global__ void __launch_bounds__(256)
m2l_reduce_kernel(int nboxes, const int* __restrict__ off, const unsigned char* __restrict__ batched, int P,
                  const double* __restrict__ tmp, double2* __restrict__ L) {
  extern __shared__ double sum[];
  int b = blockIdx.x;
  if (b >= nboxes) return;
  const int pp = P * P, nc = P * (P + 1) / 2;
  int e0 = off[b], e1 = off[b + 1];
  for (int row = threadIdx.x; row < pp; row += blockDim.x) {
    double s = 0;
    for (int e = e0; e < e1; ++e)
      if (batched[e]) s += tmp[(size_t)e * pp + row];
    sum[row] = s;
  }
  __syncthreads();
  for (int nms = threadIdx.x; nms < nc; nms += blockDim.x) {
    int n = 0; while ((n + 1) * (n + 2) / 2 <= nms) ++n;
    int m = nms - n * (n + 1) / 2;
    L[(size_t)b * nc + nms] = make_double2(sum[n * n + n + m], m > 0 ? sum[n * n + n - m] : 0.0);
  }
}

template <int RPT>
void launch_gemm(fmmb_plan* plan, int P, double* tmp, cudaStream_t s) {
  M2LClasses& C = plan->cls;
  const int pp = P * P;
  size_t sh = ((size_t)pp * 8 * RPT + (size_t)kNB * (pp + 1)) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    FMMB_CUDA(cudaFuncSetAttribute(m2l_gemm_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    attr_set = true;
  }
  m2l_gemm_kernel<RPT><<<C.n_items, 256, sh, s>>>(P, C.built_p * C.built_p, C.T.p, C.item_class.p, C.item_start.p,
                                                 C.item_count.p, C.sorted_slot.p, plan->tree.m2l_src.p,
                                                 plan->M.p, tmp);
}

}  // namespace

void m2l_init_tables() { upload_laplace_tables(); }

// Plan-time: classify the M2L pairs and build the work items.
void build_m2l_classes(fmmb_plan* plan) {
  Tree& T = plan->tree;
  M2LClasses& C = plan->cls;
  cudaStream_t s = plan->stream;
  const int64_t n = T.n_lr;
  const int nb = T.nboxes;
  C.n_classes = 0; C.n_pairs = 0; C.n_items = 0; C.n_res = n; C.built_p = 0;
  if (n == 0 || plan->opts.m2l_mode == 1) return;
  if (n >= (1ll << 31)) throw StatusError{FMMB_ERR_INVALID, "more than 2^31 M2L pairs"};
  Temp tmp;
  C.slot_tgt.resize(n);
  slot_targets<<<nb, 64, 0, s>>>(T.m2l_off.p, nb, C.slot_tgt.p);
  DevBuf<unsigned long long> k0, k1;
  DevBuf<int> v0;
  k0.resize(n); k1.resize(n); v0.resize(n); C.sorted_slot.resize(n);
  class_keys<<<nblk(n, 256), 256, 0, s>>>(C.slot_tgt.p, T.m2l_src.p, n, T.key.p, T.level.p, k0.p, v0.p);
  FMMB_CUDA(cudaGetLastError());
  {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, v0.p, C.sorted_slot.p, n, 0, 36, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, k0.p, k1.p, v0.p, C.sorted_slot.p, n, 0, 36, s));
  }
  // run-length encode -> classes
  DevBuf<unsigned long long> uniq;
  DevBuf<int> count, nruns;
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
    class_items<<<nblk(ncls + 1, 256), 256, 0, s>>>(count.p, ncls, nitems.p);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, nitems.p, item_off.p, ncls + 1, s));
    t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, nitems.p, item_off.p, ncls + 1, s));
  }
  int n_items = 0;
  FMMB_CUDA(cudaMemcpyAsync(&n_items, item_off.p + ncls, sizeof(int), cudaMemcpyDeviceToHost, s));
  FMMB_CUDA(cudaStreamSynchronize(s));
  C.n_classes = ncls;
  C.n_items = n_items;
  C.item_class.resize(n_items); C.item_start.resize(n_items); C.item_count.resize(n_items);
  if (n_items)
    fill_items<<<nblk(ncls, 128), 128, 0, s>>>(count.p, start.p, item_off.p, ncls, C.item_class.p, C.item_start.p,
                                              C.item_count.p);
  C.batched.resize(n);
  mark_batched<<<ncls, 128, 0, s>>>(count.p, start.p, ncls, C.sorted_slot.p, C.batched.p);
  // residual CSR (target-major, list order preserved)
  DevBuf<int> flag, pos;
  flag.resize(n + 1); pos.resize(n + 1);
  residual_flags<<<nblk(n + 1, 256), 256, 0, s>>>(C.batched.p, n, flag.p);
  {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, pos.p, n + 1, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, flag.p, pos.p, n + 1, s));
  }
  int n_res = 0;
  FMMB_CUDA(cudaMemcpyAsync(&n_res, pos.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  FMMB_CUDA(cudaStreamSynchronize(s));
  C.n_res = n_res;
  C.n_pairs = n - n_res;
  C.res_src.resize(n_res); C.res_off.resize(nb + 1);
  if (n_res) residual_compact<<<nblk(n, 256), 256, 0, s>>>(flag.p, pos.p, n, T.m2l_src.p, C.res_src.p);
  residual_offsets<<<nblk(nb + 1, 256), 256, 0, s>>>(T.m2l_off.p, pos.p, nb, C.res_off.p);
  // representative translation vector per class
  C.class_vec.resize(ncls);
  class_vectors<<<nblk(ncls, 128), 128, 0, s>>>(start.p, ncls, C.sorted_slot.p, C.slot_tgt.p, T.m2l_src.p,
                                               T.center.p, C.class_vec.p);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));
}

// Batched far field for the current order.  Returns false if there is nothing batched
// (the caller then runs the per-pair kernel over the full lists).
bool m2l_batched(fmmb_plan* plan, cudaStream_t s) {
  M2LClasses& C = plan->cls;
  Tree& T = plan->tree;
  const int P = plan->p, pp = P * P;
  if (C.n_items == 0 || P > 8) return false;
  if (C.built_p < P) {
    // build for the largest order this kernel family handles, once
    int bp = 8;
    C.T.resize((size_t)C.n_classes * bp * bp * bp * bp);
    build_T<<<(int)C.n_classes, 256, (size_t)4 * bp * bp * sizeof(double2), s>>>(bp, C.class_vec.p, C.T.p);
    FMMB_CUDA(cudaGetLastError());
    C.built_p = bp;
    ++plan->launches;
  }
  C.tmp.resize((size_t)T.n_lr * pp);
  switch (P) {
    case 1: case 2: launch_gemm<1>(plan, P, C.tmp.p, s); break;
    case 3: case 4: launch_gemm<2>(plan, P, C.tmp.p, s); break;
    case 5: launch_gemm<4>(plan, P, C.tmp.p, s); break;
    case 6: launch_gemm<5>(plan, P, C.tmp.p, s); break;
    case 7: launch_gemm<7>(plan, P, C.tmp.p, s); break;
    default: launch_gemm<8>(plan, P, C.tmp.p, s); break;
  }
  FMMB_CUDA(cudaGetLastError());
  ++plan->launches;
  int threads = pp < 64 ? 64 : (pp > 256 ? 256 : pp);
  m2l_reduce_kernel<<<T.nboxes, threads, pp * sizeof(double), s>>>(T.nboxes, T.m2l_off.p, C.batched.p, P,
                                                                  C.tmp.p, plan->L.p);
  FMMB_CUDA(cudaGetLastError());
  ++plan->launches;
  return true;
}

}  // namespace fmmb
