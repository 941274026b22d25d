// csrc/yukawa.cu -- YukawaCartesian on the GPU: K(t,s) = exp(-kappa |t-s|) / |t-s| with Cartesian Taylor
// expansions, (P+1)(P+2)(P+3)/6 real coefficients per box over multi-indices n = (i,j,k), i+j+k <= P,
// enumerated i-major like the reference (kernel/YukawaCartesian.hpp:111-121).
//
// Replaces (reference kernel/YukawaCartesian.hpp):
//   operator() :148-159 applied pair by pair (Direct.hpp:99-125)   -> yk_p2p_kernel (warp per <= 32 targets)
//   P2M :169-180   M_n += q (c - x)^n / n!                           -> yk_p2m_kernel (warp per leaf)
//   M2M :190-210   M_n += sum_{m <= n} M'_m d^(n-m) / (n-m)!         -> yk_m2m_kernel (block per parent, level sweep)
//   M2L :250-291 + getCoeff :356-672                                 -> yk_table_kernel + yk_m2l_kernel
//        L_k += sum_{|n+k| <= P} a'_{n+k} M_n,  a'_m = m! a_m, a_m = Taylor coefficients of exp(-kappa R)/R at the
//        translation vector.  a depends on the translation only: it is tabulated ONCE per translation class
//        (box centres sit on a lattice, csrc/m2l_classes.cu) instead of per pair, by the one recurrence that all
//        hand-unrolled cases of getCoeff are instances of (s = |n|):
//          b_n = -kappa/s ( sum_d x_d a_{n-e_d} + sum_d a_{n-2e_d} )
//          a_n = 1/(s R^2) ( -kappa (sum_d x_d b_{n-e_d} + sum_d b_{n-2e_d}) - (2s-1) sum_d x_d a_{n-e_d}
//                            - (s-1) sum_d a_{n-2e_d} )
//   L2L :301-321   L_n += sum_{k >= n} L'_k t^(k-n) / (k-n)!         -> yk_l2l_kernel (block per child, level sweep)
//   L2P :332-352   phi = L_n d^n / n!; gradient as phi n_d / d_d with the |d_d| < 1e-12 guard -> yk_l2p_kernel
// The reference's operators carry a trailing `unsigned p` that its own executor cannot supply; the
// engine defines set_p(p) as "all tables at order p" (SURVEY Q16).  Orders 1..kYkMaxP are built.
#include "common.cuh"
#include "../hostcxx/bem_math.hpp"
#include <algorithm>
#include <cmath>

namespace fmmb {

// panel data of a BEM plan (csrc/bem.cu)
const bem::Panel* bem_panels(const BemData* b);
const int* bem_bc(const BemData* b);

constexpr int kYkMaxP = 16;                                                 // = FMMB_MAX_P: index tables for every order
constexpr int kYkMaxT = (kYkMaxP + 1) * (kYkMaxP + 2) * (kYkMaxP + 3) / 6;   // 969
// Kernels that keep per-term state in registers, local or static shared memory are compiled twice: for orders up to 10
// (286 terms -- every BASELINE configuration) and up to 16 (969 terms), so that the common orders do not pay for the
// arrays of the large ones (YkCap<MP>); the host picks by the plan's order (YK_MP).
template <int MP>
struct YkCap {
  static constexpr int T = (MP + 1) * (MP + 2) * (MP + 3) / 6;
  static constexpr int ACC = (T + 31) / 32, ACC128 = (T + 127) / 128;
};

struct YukawaData {
  double kappa = 0.125;
  DevBuf<double> M, L;               // box-major, nt(P) doubles per box
  DevBuf<double> res_near, res_far;  // double4 per body, tree order (allocated as doubles)
  DevBuf<int> slot_class;            // M2L slot (target-major CSR position) -> translation class
  bool have_classes = false;
  std::map<int, DevBuf<double>*> tables;   // per order: [class][nt], a' of the class's translation vector
  ~YukawaData() { for (auto& kv : tables) delete kv.second; }
};

void yukawa_free(YukawaData* d) { delete d; }
double yukawa_kappa(const YukawaData* d) { return d->kappa; }

namespace {

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }
__host__ __device__ inline int yk_terms(int P) { return (P + 1) * (P + 2) * (P + 3) / 6; }

// per order P: multi-index of each term and the start of each i-slab
__constant__ unsigned char c_yI[kYkMaxP + 1][kYkMaxT], c_yJ[kYkMaxP + 1][kYkMaxT], c_yK[kYkMaxP + 1][kYkMaxT];
__constant__ short c_ybase[kYkMaxP + 1][kYkMaxP + 2];
__constant__ double c_yfact[2 * kYkMaxP + 2], c_yrfact[2 * kYkMaxP + 2];

__device__ __forceinline__ int yk_idx(int P, int i, int j, int k) {
  return c_ybase[P][i] + j * (P - i + 1) - (j * (j - 1)) / 2 + k;
}

void upload_tables() {
  static bool done[64] = {false};
  int dev = 0;
  FMMB_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return;
  std::vector<unsigned char> I((kYkMaxP + 1) * kYkMaxT, 0), J(I), K(I);
  std::vector<short> base((kYkMaxP + 1) * (kYkMaxP + 2), 0);
  for (int P = 1; P <= kYkMaxP; ++P) {
    int t = 0;
    for (int i = 0; i <= P; ++i) {
      base[P * (kYkMaxP + 2) + i] = (short)t;
      for (int j = 0; j <= P - i; ++j)
        for (int k = 0; k <= P - i - j; ++k) {
          I[P * kYkMaxT + t] = (unsigned char)i; J[P * kYkMaxT + t] = (unsigned char)j; K[P * kYkMaxT + t] = (unsigned char)k;
          ++t;
        }
    }
    base[P * (kYkMaxP + 2) + P + 1] = (short)t;
  }
  double f[2 * kYkMaxP + 2], rf[2 * kYkMaxP + 2];
  f[0] = 1.0;
  for (int i = 1; i < 2 * kYkMaxP + 2; ++i) f[i] = i * f[i - 1];
  for (int i = 0; i < 2 * kYkMaxP + 2; ++i) rf[i] = 1.0 / f[i];
  FMMB_CUDA(cudaMemcpyToSymbol(c_yI, I.data(), I.size()));
  FMMB_CUDA(cudaMemcpyToSymbol(c_yJ, J.data(), J.size()));
  FMMB_CUDA(cudaMemcpyToSymbol(c_yK, K.data(), K.size()));
  FMMB_CUDA(cudaMemcpyToSymbol(c_ybase, base.data(), base.size() * sizeof(short)));
  FMMB_CUDA(cudaMemcpyToSymbol(c_yfact, f, sizeof f));
  FMMB_CUDA(cudaMemcpyToSymbol(c_yrfact, rf, sizeof rf));
  if (dev < 64) done[dev] = true;
}

__global__ void yk_gather(const double* __restrict__ q, const unsigned* __restrict__ perm, int64_t n,
                          double4* __restrict__ body) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) body[i].w = q[perm[i]];
}

// d^e / e!, e = 0..P, for the three coordinates of one vector: out[c * (P+1) + e]
__device__ __forceinline__ void scaled_powers(int P, double dx, double dy, double dz, double* out) {
  const double d[3] = {dx, dy, dz};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double v = 1.0;
    out[c * (P + 1)] = 1.0;
    for (int e = 1; e <= P; ++e) { v *= d[c]; out[c * (P + 1) + e] = v * c_yrfact[e]; }
  }
}

// ---- P2M: warp per leaf; lanes stage (c - x)^e / e! of 32 bodies, then lane = coefficient --------------------
template <int MP>
__global__ void __launch_bounds__(128)
yk_p2m_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
              const unsigned* __restrict__ be, const double4* __restrict__ center,
              const double4* __restrict__ body, int P, double* __restrict__ M) {
  extern __shared__ double yk_sh[];
  const int nt = yk_terms(P), pw = 3 * (P + 1) + 1;       // + charge
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double* tile = yk_sh + (size_t)wl * 32 * pw;
  const int b = leaves[w];
  const double4 c = center[b];
  const unsigned b0 = bb[b], b1 = be[b];
  double acc[YkCap<MP>::ACC];
#pragma unroll
  for (int i = 0; i < YkCap<MP>::ACC; ++i) acc[i] = 0.0;
  for (unsigned base = b0; base < b1; base += 32) {
    const int cnt = (int)min(32u, b1 - base);
    __syncwarp();
    if (lane < cnt) {
      const double4 p = body[base + lane];
      scaled_powers(P, c.x - p.x, c.y - p.y, c.z - p.z, tile + lane * pw);
      tile[lane * pw + 3 * (P + 1)] = p.w;
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < YkCap<MP>::ACC; ++a) {
      const int t = lane + 32 * a;
      if (t < nt) {
        const int i = c_yI[P][t], j = P + 1 + c_yJ[P][t], k = 2 * (P + 1) + c_yK[P][t];
        double s = 0;
        for (int m = 0; m < cnt; ++m) {
          const double* r = tile + m * pw;
          s += r[3 * (P + 1)] * r[i] * r[j] * r[k];
        }
        acc[a] += s;
      }
    }
  }
  double* Mb = M + (size_t)b * nt;
#pragma unroll
  for (int a = 0; a < YkCap<MP>::ACC; ++a) {
    const int t = lane + 32 * a;
    if (t < nt) Mb[t] = acc[a];
  }
}

// ---- M2M: block per parent of one level, children in index order -------------------------------------------
template <int MP>
__global__ void __launch_bounds__(128)
yk_m2m_kernel(int lo, int hi, const unsigned* __restrict__ key, const unsigned* __restrict__ cbegin,
              const unsigned* __restrict__ cend, const double4* __restrict__ center, int P, double* __restrict__ M) {
  const int b = lo + blockIdx.x;
  if (b >= hi || (key[b] >> 31)) return;          // leaves got their multipole from P2M
  __shared__ double Ms[YkCap<MP>::T];
  __shared__ double pw[3 * (MP + 1)];
  const int nt = yk_terms(P);
  const double4 cp = center[b];
  double acc[YkCap<MP>::ACC128];
#pragma unroll
  for (int a = 0; a < YkCap<MP>::ACC128; ++a) acc[a] = 0.0;
  for (unsigned c = cbegin[b]; c < cend[b]; ++c) {
    const double4 cc = center[c];
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += blockDim.x) Ms[t] = M[(size_t)c * nt + t];
    if (threadIdx.x == 0) scaled_powers(P, cp.x - cc.x, cp.y - cc.y, cp.z - cc.z, pw);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < YkCap<MP>::ACC128; ++a) {
      const int t = threadIdx.x + 128 * a;
      if (t < nt) {
        const int I = c_yI[P][t], J = c_yJ[P][t], K = c_yK[P][t];
        double s = 0;
        for (int ii = 0; ii <= I; ++ii)
          for (int jj = 0; jj <= J; ++jj) {
            const double f = pw[I - ii] * pw[P + 1 + J - jj];
            const int row = yk_idx(P, ii, jj, 0);
            for (int kk = 0; kk <= K; ++kk) s += Ms[row + kk] * f * pw[2 * (P + 1) + K - kk];
          }
        acc[a] += s;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < YkCap<MP>::ACC128; ++a) {
    const int t = threadIdx.x + 128 * a;
    if (t < nt) M[(size_t)b * nt + t] = acc[a];
  }
}

// ---- derivative table of one translation vector (block-cooperative, a and b in shared memory) ---------------
// On return a[] holds a'_n = n! a_n for all |n| <= P.
__device__ void yk_coeff_table(int P, double kappa, double x, double y, double z, double* a, double* b) {
  const int nt = yk_terms(P);
  const double R2 = x * x + y * y + z * z, R = sqrt(R2), R2_1 = 1.0 / R2;
  const double xv[3] = {x, y, z};
  if (threadIdx.x == 0) { b[0] = exp(-kappa * R); a[0] = b[0] / R; }
  __syncthreads();
  for (int s = 1; s <= P; ++s) {
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
      const int n[3] = {c_yI[P][t], c_yJ[P][t], c_yK[P][t]};
      if (n[0] + n[1] + n[2] != s) continue;
      double xa = 0, xb = 0, a2 = 0, b2 = 0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (n[d] >= 1) {
          const int q = yk_idx(P, n[0] - (d == 0), n[1] - (d == 1), n[2] - (d == 2));
          xa += xv[d] * a[q]; xb += xv[d] * b[q];
        }
        if (n[d] >= 2) {
          const int q = yk_idx(P, n[0] - 2 * (d == 0), n[1] - 2 * (d == 1), n[2] - 2 * (d == 2));
          a2 += a[q]; b2 += b[q];
        }
      }
      b[t] = -kappa / s * (xa + a2);
      a[t] = R2_1 / s * (-kappa * (xb + b2) - (2 * s - 1) * xa - (s - 1) * a2);
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < nt; t += blockDim.x) a[t] *= c_yfact[c_yI[P][t]] * c_yfact[c_yJ[P][t]] * c_yfact[c_yK[P][t]];
  __syncthreads();
}

template <int MP>
__global__ void __launch_bounds__(128)
yk_table_kernel(int P, double kappa, const double4* __restrict__ vec, double* __restrict__ table) {
  __shared__ double a[YkCap<MP>::T], b[YkCap<MP>::T];
  const double4 v = vec[blockIdx.x];
  yk_coeff_table(P, kappa, v.x, v.y, v.z, a, b);
  const int nt = yk_terms(P);
  for (int t = threadIdx.x; t < nt; t += blockDim.x) table[(size_t)blockIdx.x * nt + t] = a[t];
}

// ---- M2P (treecode evaluator of YukawaCartesianBEM, reference kernel/YukawaCartesianBEM.hpp:298-327): warp per
// leaf, lane per target panel.  For every source box accepted for the leaf or one of its ancestors the Taylor
// table a'_n of exp(-kappa R)/R at (panel centre - box centre) is built in the lane's private arrays by the
// recurrence of yk_coeff_table above (order by order: entries of order s need orders s-1 and s-2), then
// phi += sum_n a'_n M_n.  Only panels whose BC selects this set are touched; set 0 adds, set 1 subtracts.  The
// reference's FMM evaluator is broken for this kernel class while its treecode agrees with Direct, so this
// is the far-field path of BASELINE config 3 that is pinned to the reference.
template <int SET, int MP>
__global__ void __launch_bounds__(128)
yk_bem_m2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const unsigned* __restrict__ parent, const int* __restrict__ off,
                  const int* __restrict__ src, const double4* __restrict__ center, const bem::Panel* __restrict__ pan,
                  const int* __restrict__ bc, int P, double kappa, const double* __restrict__ M,
                  double* __restrict__ res) {
  const int nt = yk_terms(P);
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int leaf = leaves[w];
  double a[YkCap<MP>::T], b[YkCap<MP>::T];
  for (unsigned i = bb[leaf] + lane; i < be[leaf]; i += 32) {
    if (bc[i] != SET) continue;
    const double px = pan[i].c[0], py = pan[i].c[1], pz = pan[i].c[2];
    double acc = 0;
    for (int anc = leaf;; anc = (int)parent[anc]) {
      for (int e = off[anc]; e < off[anc + 1]; ++e) {
        const int sb = src[e];
        const double4 c = center[sb];
        const double xv[3] = {px - c.x, py - c.y, pz - c.z};
        const double R2 = xv[0] * xv[0] + xv[1] * xv[1] + xv[2] * xv[2], R = sqrt(R2), R2_1 = 1.0 / R2;
        b[0] = exp(-kappa * R);
        a[0] = b[0] / R;
        for (int s = 1; s <= P; ++s)
          for (int i0 = 0; i0 <= s; ++i0)
            for (int j0 = 0; j0 <= s - i0; ++j0) {
              const int n[3] = {i0, j0, s - i0 - j0};
              const int t = yk_idx(P, n[0], n[1], n[2]);
              double xa = 0, xb = 0, a2 = 0, b2 = 0;
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                if (n[d] >= 1) {
                  const int q = yk_idx(P, n[0] - (d == 0), n[1] - (d == 1), n[2] - (d == 2));
                  xa += xv[d] * a[q]; xb += xv[d] * b[q];
                }
                if (n[d] >= 2) {
                  const int q = yk_idx(P, n[0] - 2 * (d == 0), n[1] - 2 * (d == 1), n[2] - 2 * (d == 2));
                  a2 += a[q]; b2 += b[q];
                }
              }
              b[t] = -kappa / s * (xa + a2);
              a[t] = R2_1 / s * (-kappa * (xb + b2) - (2 * s - 1) * xa - (s - 1) * a2);
            }
        const double* Ms = M + (size_t)sb * nt;
        double v = 0;
        for (int t = 0; t < nt; ++t) v += a[t] * Ms[t] * (c_yfact[c_yI[P][t]] * c_yfact[c_yJ[P][t]] * c_yfact[c_yK[P][t]]);
        acc += v;
      }
      if (anc == 0) break;
    }
    res[i] += SET == 0 ? acc : -acc;
  }
}

// ---- M2P of the point kernel (treecode, reference kernel/YukawaCartesian.hpp:221-240): warp per leaf, lane per body.
// Same lane-private table as above; the gradient uses ax_m = (m_x + 1) a_{m + e_x} for |m| < P and nothing for
// |m| = P, which is what getCoeff's `ax[Im1x] = a[I] * i` lines leave in ax / ay / az (:388-672).
template <int MP>
__global__ void __launch_bounds__(128)
yk_m2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
              const unsigned* __restrict__ be, const unsigned* __restrict__ parent, const int* __restrict__ off,
              const int* __restrict__ src, const double4* __restrict__ center, const double4* __restrict__ body, int P,
              double kappa, const double* __restrict__ M, double4* __restrict__ res) {
  const int nt = yk_terms(P);
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int leaf = leaves[w];
  double a[YkCap<MP>::T], b[YkCap<MP>::T];
  for (unsigned i = bb[leaf] + lane; i < be[leaf]; i += 32) {
    const double4 p = body[i];
    double pot = 0, gx = 0, gy = 0, gz = 0;
    for (int anc = leaf;; anc = (int)parent[anc]) {
      for (int e = off[anc]; e < off[anc + 1]; ++e) {
        const int sb = src[e];
        const double4 c = center[sb];
        const double xv[3] = {p.x - c.x, p.y - c.y, p.z - c.z};
        const double R2 = xv[0] * xv[0] + xv[1] * xv[1] + xv[2] * xv[2], R = sqrt(R2), R2_1 = 1.0 / R2;
        b[0] = exp(-kappa * R);
        a[0] = b[0] / R;
        for (int s = 1; s <= P; ++s)
          for (int i0 = 0; i0 <= s; ++i0)
            for (int j0 = 0; j0 <= s - i0; ++j0) {
              const int n[3] = {i0, j0, s - i0 - j0};
              const int t = yk_idx(P, n[0], n[1], n[2]);
              double xa = 0, xb = 0, a2 = 0, b2 = 0;
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                if (n[d] >= 1) {
                  const int q = yk_idx(P, n[0] - (d == 0), n[1] - (d == 1), n[2] - (d == 2));
                  xa += xv[d] * a[q]; xb += xv[d] * b[q];
                }
                if (n[d] >= 2) {
                  const int q = yk_idx(P, n[0] - 2 * (d == 0), n[1] - 2 * (d == 1), n[2] - 2 * (d == 2));
                  a2 += a[q]; b2 += b[q];
                }
              }
              b[t] = -kappa / s * (xa + a2);
              a[t] = R2_1 / s * (-kappa * (xb + b2) - (2 * s - 1) * xa - (s - 1) * a2);
            }
        const double* Ms = M + (size_t)sb * nt;
        for (int t = 0; t < nt; ++t) {
          const int ii = c_yI[P][t], jj = c_yJ[P][t], kk = c_yK[P][t];
          const double mf = Ms[t] * (c_yfact[ii] * c_yfact[jj] * c_yfact[kk]);
          pot += a[t] * mf;
          if (ii + jj + kk < P) {
            gx += a[yk_idx(P, ii + 1, jj, kk)] * (ii + 1) * mf;
            gy += a[yk_idx(P, ii, jj + 1, kk)] * (jj + 1) * mf;
            gz += a[yk_idx(P, ii, jj, kk + 1)] * (kk + 1) * mf;
          }
        }
      }
      if (anc == 0) break;
    }
    res[i] = make_double4(pot, gx, gy, gz);
  }
}

__global__ void yk_slot_class_kernel(int n_items, const int* __restrict__ item_class, const int* __restrict__ item_start,
                                     const int* __restrict__ item_count, const int* __restrict__ sorted_slot,
                                     int* __restrict__ slot_class) {
  const int it = blockIdx.x;
  if (it >= n_items) return;
  for (int j = threadIdx.x; j < item_count[it]; j += blockDim.x) slot_class[sorted_slot[item_start[it] + j]] = item_class[it];
}

// ---- M2L: block per target box, sources in LR_list order; thread = output coefficient k --------------------
template <int MP>
__global__ void __launch_bounds__(128)
yk_m2l_kernel(int nboxes, const int* __restrict__ off, const int* __restrict__ src, const int* __restrict__ slot_class,
              const double* __restrict__ table, const double4* __restrict__ center, int P, double kappa,
              const double* __restrict__ M, double* __restrict__ L) {
  const int b = blockIdx.x;
  if (b >= nboxes) return;
  __shared__ double a[YkCap<MP>::T], bt[YkCap<MP>::T], Ms[YkCap<MP>::T];
  const int nt = yk_terms(P);
  double acc[YkCap<MP>::ACC128];
#pragma unroll
  for (int i = 0; i < YkCap<MP>::ACC128; ++i) acc[i] = 0.0;
  const double4 ct = center[b];
  for (int e = off[b]; e < off[b + 1]; ++e) {
    const int sb = src[e];
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += blockDim.x) Ms[t] = M[(size_t)sb * nt + t];
    if (slot_class) {
      const double* tab = table + (size_t)slot_class[e] * nt;
      for (int t = threadIdx.x; t < nt; t += blockDim.x) a[t] = tab[t];
      __syncthreads();
    } else {
      const double4 cs = center[sb];
      yk_coeff_table(P, kappa, ct.x - cs.x, ct.y - cs.y, ct.z - cs.z, a, bt);
    }
#pragma unroll
    for (int i = 0; i < YkCap<MP>::ACC128; ++i) {
      const int t = threadIdx.x + 128 * i;
      if (t < nt) {
        const int ik = c_yI[P][t], jk = c_yJ[P][t], kk = c_yK[P][t];
        const int rem = P - ik - jk - kk;              // |n| <= rem
        double s = 0;
        for (int in = 0; in <= rem; ++in)
          for (int jn = 0; jn <= rem - in; ++jn) {
            const int mrow = yk_idx(P, in, jn, 0), arow = yk_idx(P, ik + in, jk + jn, kk);
            for (int kn = 0; kn <= rem - in - jn; ++kn) s += a[arow + kn] * Ms[mrow + kn];
          }
        acc[i] += s;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < YkCap<MP>::ACC128; ++i) {
    const int t = threadIdx.x + 128 * i;
    if (t < nt) L[(size_t)b * nt + t] = acc[i];
  }
}

// ---- L2L: block per child of one level ------------------------------------------------------------------------
template <int MP>
__global__ void __launch_bounds__(128)
yk_l2l_kernel(int lo, int hi, const unsigned* __restrict__ parent, const unsigned char* __restrict__ has_local,
              const double4* __restrict__ center, int P, double* __restrict__ L) {
  const int b = lo + blockIdx.x;
  if (b >= hi) return;
  const int par = (int)parent[b];
  if (!has_local[par]) return;
  __shared__ double Ls[YkCap<MP>::T];
  __shared__ double pw[3 * (MP + 1)];
  const int nt = yk_terms(P);
  const double4 cc = center[b], cp = center[par];
  for (int t = threadIdx.x; t < nt; t += blockDim.x) Ls[t] = L[(size_t)par * nt + t];
  if (threadIdx.x == 0) scaled_powers(P, cc.x - cp.x, cc.y - cp.y, cc.z - cp.z, pw);
  __syncthreads();
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const int I = c_yI[P][t], J = c_yJ[P][t], K = c_yK[P][t];
    double s = 0;
    for (int ii = I; ii <= P; ++ii)
      for (int jj = J; jj <= P - ii; ++jj) {
        const double f = pw[ii - I] * pw[P + 1 + jj - J];
        const int row = yk_idx(P, ii, jj, 0);
        for (int kk = K; kk <= P - ii - jj; ++kk) s += Ls[row + kk] * f * pw[2 * (P + 1) + kk - K];
      }
    L[(size_t)b * nt + t] += s;
  }
}

// ---- L2P: warp per leaf, lane per body ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
yk_l2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
              const unsigned* __restrict__ be, const double4* __restrict__ center,
              const unsigned char* __restrict__ has_local, const double4* __restrict__ body, int P,
              const double* __restrict__ L, double4* __restrict__ res) {
  extern __shared__ double yk_sh[];
  const int nt = yk_terms(P), pw = 3 * (P + 1);
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int b = leaves[w];
  const unsigned b0 = bb[b], b1 = be[b];
  if (!has_local[b]) {
    for (unsigned i = b0 + lane; i < b1; i += 32) res[i] = make_double4(0, 0, 0, 0);
    return;
  }
  double* Ls = yk_sh + (size_t)wl * (nt + 32 * pw);
  double* powers = Ls + nt + lane * pw;
  for (int t = lane; t < nt; t += 32) Ls[t] = L[(size_t)b * nt + t];
  __syncwarp();
  const double4 c = center[b];
  for (unsigned i = b0 + lane; i < b1; i += 32) {
    const double4 p = body[i];
    const double dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
    scaled_powers(P, dx, dy, dz, powers);
    const double ix = fabs(dx) < 1e-12 ? 0.0 : 1.0 / dx, iy = fabs(dy) < 1e-12 ? 0.0 : 1.0 / dy,
                 iz = fabs(dz) < 1e-12 ? 0.0 : 1.0 / dz;
    double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    for (int t = 0; t < nt; ++t) {
      const int I = c_yI[P][t], J = c_yJ[P][t], K = c_yK[P][t];
      const double phi = Ls[t] * powers[I] * powers[P + 1 + J] * powers[2 * (P + 1) + K];
      r0 += phi; r1 += phi * I * ix; r2 += phi * J * iy; r3 += phi * K * iz;
    }
    res[i] = make_double4(r0, r1, r2, r3);
  }
}

// ---- near field --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void yk_pair(double kappa, const double4 t, const double4 s, double& pot, double& fx,
                                        double& fy, double& fz) {
  const double dx = t.x - s.x, dy = t.y - s.y, dz = t.z - s.z;      // t - s (:150)
  const double r2 = dx * dx + dy * dy + dz * dz;
  const double r = sqrt(r2);
  double invR = 1.0 / r, invR2 = 1.0 / r2;
  if (r < 1e-8) { invR = 0; invR2 = 0; }
  const double p = exp(-kappa * r) * invR;
  const double f = p * (kappa * r + 1) * invR2 * s.w;
  pot += p * s.w;
  fx -= dx * f; fy -= dy * f; fz -= dz * f;
}

constexpr int kYkWarps = 4;
__global__ void __launch_bounds__(32 * kYkWarps)
yk_p2p_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
              const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
              const double4* __restrict__ body, double kappa, double4* __restrict__ res) {
  __shared__ double4 tiles[kYkWarps][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kYkWarps + wl;
  if (item >= nitems) return;
  double4* tile = tiles[wl];
  const int4 it = items[item];
  const int r = it.z, S = 32 / r;
  const int ti = lane % r, sp = lane / r;
  const bool act = sp < S;
  const double4 t = body[it.y + ti];
  double pot = 0, fx = 0, fy = 0, fz = 0;
  for (int e = off[it.x]; e < off[it.x + 1]; ++e) {
    const int sb = src[e];
    const unsigned c0 = bb[sb], c1 = be[sb];
    for (unsigned base = c0; base < c1; base += 32) {
      const int cnt = (int)min(32u, c1 - base);
      __syncwarp();
      if (lane < cnt) tile[lane] = body[base + lane];
      __syncwarp();
      if (act)
        for (int k = sp; k < cnt; k += S) yk_pair(kappa, t, tile[k], pot, fx, fy, fz);
    }
  }
  for (int q = 1; q < S; ++q) {
    const int from = (lane + q * r) & 31;
    const double a = __shfl_sync(0xffffffffu, pot, from), b1 = __shfl_sync(0xffffffffu, fx, from),
                 b2 = __shfl_sync(0xffffffffu, fy, from), b3 = __shfl_sync(0xffffffffu, fz, from);
    if (lane < r) { pot += a; fx += b1; fy += b2; fz += b3; }
  }
  if (lane < r) res[it.y + lane] = make_double4(pot, fx, fy, fz);
}


__global__ void __launch_bounds__(128)
yk_direct_kernel(const double* __restrict__ spts, const double* __restrict__ q, int64_t ns,
                 const double* __restrict__ tpts, int64_t nt, double kappa, double4* __restrict__ out) {
  __shared__ double4 tile[128];
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool act = i < nt;
  const double4 t = act ? make_double4(tpts[3 * i], tpts[3 * i + 1], tpts[3 * i + 2], 0) : make_double4(0, 0, 0, 0);
  double pot = 0, fx = 0, fy = 0, fz = 0;
  for (int64_t base = 0; base < ns; base += 128) {
    const int cnt = (int)min((int64_t)128, ns - base);
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      const int64_t j = base + threadIdx.x;
      tile[threadIdx.x] = make_double4(spts[3 * j], spts[3 * j + 1], spts[3 * j + 2], q[j]);
    }
    __syncthreads();
    if (act)
      for (int k = 0; k < cnt; ++k) yk_pair(kappa, t, tile[k], pot, fx, fy, fz);
  }
  if (act) out[i] = make_double4(pot, fx, fy, fz);
}

// ---- YukawaCartesianBEM far field (reference kernel/YukawaCartesianBEM.hpp:240-273 P2M, :340-361 L2P) ------------
// Two expansion sets, processed one after the other through the same M / L buffers: set 0 (single layer) is fed by
// POTENTIAL panels and read by POTENTIAL targets, set 1 (double layer) by NORMAL_DERIV panels / targets.
// P2M: warp per leaf; an entry is a (panel, quadrature point) pair; lanes stage (c - q)^e / e!, the weight
// q w_j Area and the panel normal of 32 entries, then lane = coefficient sums over the entries:
//   set 0:  M_n += mult C_n,               C_n = dX^n / n!
//   set 1:  M_n -= mult (n . grad_dX) C_n   (the reference writes the derivative as C_n n_d / dX_d)
template <int SET, int MP>
__global__ void __launch_bounds__(128)
yk_bem_p2m_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const double4* __restrict__ center,
                  const double4* __restrict__ body, const bem::Panel* __restrict__ pan, const int* __restrict__ bc,
                  bem::Rule rule, int P, double* __restrict__ M) {
  extern __shared__ double yk_sh[];
  const int nt = yk_terms(P), pw = 3 * (P + 1) + 4;       // powers, mult, normal
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double* tile = yk_sh + (size_t)wl * 32 * pw;
  const int b = leaves[w];
  const double4 c = center[b];
  const unsigned b0 = bb[b], b1 = be[b];
  const int K = rule.n;
  const int nent = (int)(b1 - b0) * K;
  double acc[YkCap<MP>::ACC];
#pragma unroll
  for (int i = 0; i < YkCap<MP>::ACC; ++i) acc[i] = 0.0;
  for (int base = 0; base < nent; base += 32) {
    const int ent = base + lane;
    const int cnt = min(32, nent - base);
    __syncwarp();
    if (ent < nent) {
      const unsigned i = b0 + ent / K;
      const int qi = ent % K;
      double* r = tile + lane * pw;
      const bem::Panel& s = pan[i];
      double q[3];
      bem::quad_point(s, rule.pt[qi], q);
      scaled_powers(P, c.x - q[0], c.y - q[1], c.z - q[2], r);
      r[3 * (P + 1)] = bc[i] == SET ? body[i].w * rule.w[qi] * s.area : 0.0;
      r[3 * (P + 1) + 1] = s.nrm[0]; r[3 * (P + 1) + 2] = s.nrm[1]; r[3 * (P + 1) + 3] = s.nrm[2];
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < YkCap<MP>::ACC; ++a) {
      const int t = lane + 32 * a;
      if (t < nt) {
        const int I = c_yI[P][t], J = c_yJ[P][t], Kk = c_yK[P][t];
        const int ix = I, iy = P + 1 + J, iz = 2 * (P + 1) + Kk;
        double sum = 0;
        for (int m = 0; m < cnt; ++m) {
          const double* r = tile + m * pw;
          const double mult = r[3 * (P + 1)];
          if (SET == 0) {
            sum += mult * r[ix] * r[iy] * r[iz];
          } else {
            double g = 0;
            if (I > 0) g += r[3 * (P + 1) + 1] * r[ix - 1] * r[iy] * r[iz];
            if (J > 0) g += r[3 * (P + 1) + 2] * r[ix] * r[iy - 1] * r[iz];
            if (Kk > 0) g += r[3 * (P + 1) + 3] * r[ix] * r[iy] * r[iz - 1];
            sum -= mult * g;
          }
        }
        acc[a] += sum;
      }
    }
  }
  double* Mb = M + (size_t)b * nt;
#pragma unroll
  for (int a = 0; a < YkCap<MP>::ACC; ++a) {
    const int t = lane + 32 * a;
    if (t < nt) Mb[t] = acc[a];
  }
}

// L2P: warp per leaf, lane per target panel whose boundary condition selects this set
template <int SET>
__global__ void __launch_bounds__(128)
yk_bem_l2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const double4* __restrict__ center,
                  const unsigned char* __restrict__ has_local, const bem::Panel* __restrict__ pan,
                  const int* __restrict__ bc, int P, const double* __restrict__ L, double* __restrict__ res) {
  extern __shared__ double yk_sh[];
  const int nt = yk_terms(P), pw = 3 * (P + 1);
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int b = leaves[w];
  if (!has_local[b]) return;
  double* Ls = yk_sh + (size_t)wl * (nt + 32 * pw);
  double* powers = Ls + nt + lane * pw;
  for (int t = lane; t < nt; t += 32) Ls[t] = L[(size_t)b * nt + t];
  __syncwarp();
  const double4 c = center[b];
  for (unsigned i = bb[b] + lane; i < be[b]; i += 32) {
    if (bc[i] != SET) continue;
    const bem::Panel& tp = pan[i];
    scaled_powers(P, tp.c[0] - c.x, tp.c[1] - c.y, tp.c[2] - c.z, powers);
    double acc = 0;
    for (int t = 0; t < nt; ++t) acc += Ls[t] * powers[c_yI[P][t]] * powers[P + 1 + c_yJ[P][t]] * powers[2 * (P + 1) + c_yK[P][t]];
    res[i] = SET == 0 ? acc : -acc;
  }
}

// launch of a kernel compiled for orders <= 10 and <= 16 (YkCap)
#define YK_MP(P, KERNEL, ...)                      \
  do {                                             \
    if ((P) <= 10) KERNEL<10> __VA_ARGS__;         \
    else KERNEL<16> __VA_ARGS__;                   \
  } while (0)
// dynamic shared memory above the 48 KB default (orders 14..16 of the tile kernels)
template <typename K>
void yk_shared_limit(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) FMMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// derivative tables of the translation classes at order P (built once per order, kept)
const double* yk_class_tables(fmmb_plan* plan, YukawaData* d, int P, cudaStream_t s) {
  const bool use_classes = d->have_classes && plan->opts.m2l_mode != 1;
  if (!use_classes) return nullptr;
  const int nt = yk_terms(P);
  auto it = d->tables.find(P);
  if (it != d->tables.end()) return it->second->p;
  DevBuf<double>* buf = new DevBuf<double>();
  d->tables[P] = buf;
  buf->resize((size_t)plan->cls.n_classes * nt);
  YK_MP(P, yk_table_kernel, <<<(int)plan->cls.n_classes, 128, 0, s>>>(P, d->kappa, plan->cls.class_vec.p, buf->p));
  FMMB_CUDA(cudaGetLastError());
  ++plan->launches;
  return buf->p;
}

// M2M sweep -> M2L -> L2L sweep on d->M / d->L (leaf multipoles in, complete locals out)
void yk_translations(fmmb_plan* plan, YukawaData* d, int P, const double* table, cudaStream_t s) {
  Tree& T = plan->tree;
  const int nb = T.nboxes;
  cudaEvent_t* ev = plan->ev;
  for (int l = T.nlevels - 2; l >= 0; --l) {
    const int lo = T.level_off[l], hi = T.level_off[l + 1];
    YK_MP(P, yk_m2m_kernel, <<<hi - lo, 128, 0, s>>>(lo, hi, T.key.p, T.cbegin.p, T.cend.p, T.center.p, P, d->M.p));
    ++plan->launches;
  }
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[2], s));
  if (plan->opts.evaluator == FMMB_EVAL_TREECODE) {       // treecode: multipoles only, M2P does the rest
    if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[3], s));
    return;
  }
  YK_MP(P, yk_m2l_kernel, <<<nb, 128, 0, s>>>(nb, T.m2l_off.p, T.m2l_src.p, table ? d->slot_class.p : nullptr, table,
                                               T.center.p, P, d->kappa, d->M.p, d->L.p));
  ++plan->launches;
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[3], s));
  for (int l = 1; l < T.nlevels; ++l) {
    const int lo = T.level_off[l], hi = T.level_off[l + 1];
    YK_MP(P, yk_l2l_kernel, <<<hi - lo, 128, 0, s>>>(lo, hi, T.parent.p, T.has_local.p, T.center.p, P, d->L.p));
    ++plan->launches;
  }
}

}  // namespace

void yukawa_setup(fmmb_plan* plan, double kappa) {
  Tree& T = plan->tree;
  if (plan->p > kYkMaxP) throw StatusError{FMMB_ERR_UNSUPPORTED, "YukawaCartesian is built for orders 1..16"};
  YukawaData* d = new YukawaData();
  plan->yukawa = d;
  d->kappa = kappa;
  upload_tables();
  // translation classes of the M2L pairs (built by build_m2l_classes unless m2l_mode = 1): slot -> class
  TransBatch& C = plan->cls;
  if (C.n_items > 0 && C.n_res == 0 && T.n_lr_local > 0) {
    d->slot_class.resize(T.n_lr_local);
    yk_slot_class_kernel<<<C.n_items, 64, 0, plan->stream>>>(C.n_items, C.item_class.p, C.item_start.p, C.item_count.p,
                                                            C.sorted_slot.p, d->slot_class.p);
    FMMB_CUDA(cudaGetLastError());
    FMMB_CUDA(cudaStreamSynchronize(plan->stream));
    d->have_classes = true;
  }
}

void yukawa_execute(fmmb_plan* plan, const double* d_charges, double* d_results) {
  Tree& T = plan->tree;
  YukawaData* d = plan->yukawa;
  const int P = plan->p;
  if (P > kYkMaxP) throw StatusError{FMMB_ERR_UNSUPPORTED, "YukawaCartesian is built for orders 1..16"};
  const int nt = yk_terms(P);
  const int64_t n = T.n;
  const int nb = T.nboxes;
  cudaStream_t s = plan->stream, s2 = plan->overlap_p2p ? plan->stream2 : plan->stream;
  cudaEvent_t* ev = plan->ev;
  d->M.resize((size_t)nb * nt);
  d->L.resize((size_t)nb * nt);
  d->res_near.resize(4 * (size_t)n);
  d->res_far.resize(4 * (size_t)n);
  double4* near = reinterpret_cast<double4*>(d->res_near.p);
  double4* far = reinterpret_cast<double4*>(d->res_far.p);
  plan->launches = 0;
  const double* table = plan->opts.evaluator == FMMB_EVAL_TREECODE ? nullptr : yk_class_tables(plan, d, P, s);

  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[0], s));
  yk_gather<<<nblk(n, 256), 256, 0, s>>>(exec_charges(plan, d_charges), exec_perm(plan), n, T.body.p);
  ++plan->launches;
  FMMB_CUDA(cudaEventRecord(ev[1], s));

  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s2, ev[1], 0));
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[6], s2));
  if (T.n_p2p_items) {
    yk_p2p_kernel<<<nblk(T.n_p2p_items, kYkWarps), 32 * kYkWarps, 0, s2>>>(T.p2p_items.p, T.n_p2p_items, T.bbegin.p,
                                                                          T.bend.p, T.p2p_off.p, T.p2p_src.p, T.body.p,
                                                                          d->kappa, near);
    ++plan->launches;
  }
  FMMB_CUDA(cudaEventRecord(ev[7], s2));

  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[12], s));
  {
    const size_t sh = (size_t)4 * 32 * (3 * (P + 1) + 1) * sizeof(double);
    if (P <= 10) yk_shared_limit(yk_p2m_kernel<10>, sh); else yk_shared_limit(yk_p2m_kernel<16>, sh);
    YK_MP(P, yk_p2m_kernel, <<<nblk(T.nleaves, 4), 128, sh, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p, T.center.p,
                                                                 T.body.p, P, d->M.p));
    ++plan->launches;
  }
  yk_translations(plan, d, P, table, s);      // treecode: stops after the upward pass
  if (plan->opts.evaluator == FMMB_EVAL_TREECODE) {
    if (T.n_own_leaves)
      YK_MP(P, yk_m2p_kernel, <<<nblk(T.n_own_leaves, 4), 128, 0, s>>>(T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p,
                                                                       T.parent.p, T.m2l_off.p, T.m2l_src.p, T.center.p,
                                                                       T.body.p, P, d->kappa, d->M.p, far));
    ++plan->launches;
  } else {
    const size_t sh = (size_t)4 * (nt + 32 * 3 * (P + 1)) * sizeof(double);
    yk_shared_limit(yk_l2p_kernel, sh);
    if (T.n_own_leaves)
    yk_l2p_kernel<<<nblk(T.n_own_leaves, 4), 128, sh, s>>>(T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p,
                                                          T.center.p, T.has_local.p, T.body.p, P, d->L.p, far);
    ++plan->launches;
  }
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[4], s));
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s, ev[7], 0));
  finish_results(plan, d->res_near.p, d->res_far.p, 4, d_results, s);   // multi-GPU: upward pass replicated
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[5], s));
  FMMB_CUDA(cudaGetLastError());
  plan->timed = true;
}

void yukawa_direct_raw(double kappa, const double* d_spts, const double* d_q, int64_t ns, const double* d_tpts,
                       int64_t nt, double* d_out, cudaStream_t s) {
  yk_direct_kernel<<<nblk(nt, 128), 128, 0, s>>>(d_spts, d_q, ns, d_tpts, nt, kappa, reinterpret_cast<double4*>(d_out));
  FMMB_CUDA(cudaGetLastError());
}


// YukawaCartesianBEM matvec: cached near field (csrc/bem.cu, assembled with the Yukawa panel integrals), then per
// active expansion set P2M -> translations -> L2P.  Multi-GPU: upward pass replicated, results all-gathered.
void yukawa_bem_execute(fmmb_plan* plan, const double* d_charges, double* d_results) {
  Tree& T = plan->tree;
  YukawaData* d = plan->yukawa;
  BemData* B = plan->bem;
  const int P = plan->p;
  if (P > kYkMaxP) throw StatusError{FMMB_ERR_UNSUPPORTED, "YukawaCartesianBEM is built for orders 1..16"};
  const int nt = yk_terms(P);
  cudaStream_t s = plan->stream;
  cudaEvent_t* ev = plan->ev;
  d->M.resize((size_t)T.nboxes * nt);
  d->L.resize((size_t)T.nboxes * nt);
  plan->launches = 0;
  const double* table = (plan->near_only || plan->opts.evaluator == FMMB_EVAL_TREECODE) ? nullptr
                                                                                        : yk_class_tables(plan, d, P, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[0], s));
  FMMB_CUDA(cudaEventRecord(ev[1], s));
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[6], s));
  bem_begin(plan, d_charges, s);                       // charges to tree order, cached near field, far = 0
  FMMB_CUDA(cudaEventRecord(ev[7], s));
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[12], s));
  const bem::Rule rule = bem::make_rule(bem_rule_points(B));
  const size_t sh_p2m = (size_t)4 * 32 * (3 * (P + 1) + 4) * sizeof(double);
  const size_t sh_l2p = (size_t)4 * (nt + 32 * 3 * (P + 1)) * sizeof(double);
  for (int set = 0; set < 2; ++set) {
    if (!bem_set_active(B, set) || plan->near_only) continue;
#define YK_BEM_P2M(SET, MP)                                                                                             \
  do {                                                                                                                 \
    yk_shared_limit(yk_bem_p2m_kernel<SET, MP>, sh_p2m);                                                               \
    yk_bem_p2m_kernel<SET, MP><<<nblk(T.nleaves, 4), 128, sh_p2m, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p,     \
                                                                        T.center.p, T.body.p, bem_panels(B), bem_bc(B), \
                                                                        rule, P, d->M.p);                               \
  } while (0)
    if (set == 0) { if (P <= 10) YK_BEM_P2M(0, 10); else YK_BEM_P2M(0, 16); }
    else { if (P <= 10) YK_BEM_P2M(1, 10); else YK_BEM_P2M(1, 16); }
#undef YK_BEM_P2M
    ++plan->launches;
    yk_translations(plan, d, P, table, s);
    if (T.n_own_leaves && plan->opts.evaluator == FMMB_EVAL_TREECODE) {
#define YK_BEM_M2P(SET, MP)                                                                                         \
  yk_bem_m2p_kernel<SET, MP><<<nblk(T.n_own_leaves, 4), 128, 0, s>>>(                                               \
      T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.parent.p, T.m2l_off.p, T.m2l_src.p, T.center.p,       \
      bem_panels(B), bem_bc(B), P, d->kappa, d->M.p, bem_res_far(B))
      if (set == 0) { if (P <= 10) YK_BEM_M2P(0, 10); else YK_BEM_M2P(0, 16); }
      else { if (P <= 10) YK_BEM_M2P(1, 10); else YK_BEM_M2P(1, 16); }
#undef YK_BEM_M2P
      ++plan->launches;
    } else if (T.n_own_leaves) {
      yk_shared_limit(yk_bem_l2p_kernel<0>, sh_l2p);
      yk_shared_limit(yk_bem_l2p_kernel<1>, sh_l2p);
      if (set == 0)
        yk_bem_l2p_kernel<0><<<nblk(T.n_own_leaves, 4), 128, sh_l2p, s>>>(T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p,
                                                                         T.center.p, T.has_local.p, bem_panels(B), bem_bc(B),
                                                                         P, d->L.p, bem_res_far(B));
      else
        yk_bem_l2p_kernel<1><<<nblk(T.n_own_leaves, 4), 128, sh_l2p, s>>>(T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p,
                                                                         T.center.p, T.has_local.p, bem_panels(B), bem_bc(B),
                                                                         P, d->L.p, bem_res_far(B));
      ++plan->launches;
    }
  }
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[4], s));
  finish_results(plan, bem_res_near(B), bem_res_far(B), 1, d_results, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[5], s));
  FMMB_CUDA(cudaGetLastError());
  plan->timed = true;
}

}  // namespace fmmb
