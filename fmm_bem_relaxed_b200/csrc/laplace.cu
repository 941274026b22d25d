// csrc/laplace.cu -- LaplaceSpherical operators as sm_100a kernels (FP64 throughout).
//
// Replaces the operator bodies of reference kernel/LaplaceSpherical.hpp and the executor
// loops that call them (reference include/executor/EvalInteractionLazy.hpp:122-153,239-300):
//   P2M  :213-235 (+ evalMultipole :455-488, cart2sph :528-541)   -> p2m_kernel   (warp per leaf)
//   M2M  :245-285                                                 -> m2m_kernel   (block per parent, level sweep)
//   M2L  :296-329 (+ evalLocal :491-524, Cnm :106-116)            -> m2l_pair_kernel (block per target box)
//   L2L  :378-411                                                 -> l2l_kernel   (block per child, level sweep)
//   L2P  :422-450 (+ sph2cart :546-561)                           -> l2p_kernel   (warp per leaf)
//   P2P  Direct.hpp:116-124 with operator() :153-162              -> p2p_kernel   (block per target leaf)
//
// Conventions kept from the reference: packed index n(n+1)/2+m for m>=0,
// negative m by conjugation; r = |d| + 1e-12 in cart2sph; P2P pairs with R2 < 1e-8 contribute 0.
// Differences that stay far below the 1e-10 parity tolerance: the EPS scale factors that cancel
// algebraically are dropped; i^k factors are exact (+-1) instead of std::pow(complex) values;
// cos/sin of the polar angle come from z/r and sqrt((1-x)(1+x)) instead of cos(acos(.)), and
// exp(i m phi) from (x+iy)/|xy| by recurrence; sums run in a different order.
#include "common.cuh"
#include "laplace_ops.cuh"
#include <cub/cub.cuh>
#include <cmath>
#include <algorithm>

namespace fmmb {

namespace {

using namespace ops;

// ---- charges into tree order ------------------------------------------------------------------
__global__ void gather_charges(const double* __restrict__ q, const unsigned* __restrict__ perm, int64_t n,
                               double4* __restrict__ body) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) body[i].w = q[perm[i]];
}

// ---- P2M: one warp per leaf -----------------------------------------------------------------------
// Lane = body: evaluates q * rho^n Y_n^m(alpha,-beta) into a warp-private shared tile [body][P^2]
// (real layout).  Then lane = coefficient: sums its column over the bodies.  One pass per 32 bodies.
__global__ void __launch_bounds__(128)
p2m_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
           const unsigned* __restrict__ be, const double4* __restrict__ center,
           const double4* __restrict__ body, int P, double* __restrict__ M) {
  extern __shared__ double p2m_sh[];
  const int pp = P * P, ld = pp | 1;               // odd stride: conflict-free row writes and column reads
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double* tile = p2m_sh + (size_t)wl * 32 * ld;
  const int b = leaves[w];
  const double4 c = center[b];
  const unsigned b0 = bb[b], b1 = be[b];
  double acc[(FMMB_MAX_P * FMMB_MAX_P + 31) / 32];
#pragma unroll
  for (int i = 0; i < (FMMB_MAX_P * FMMB_MAX_P + 31) / 32; ++i) acc[i] = 0.0;
  for (unsigned base = b0; base < b1; base += 32) {
    const unsigned i = base + lane;
    const int cnt = (int)min(32u, b1 - base);
    __syncwarp();
    if (i < b1) {
      const double4 p = body[i];
      const double q = p.w;
      const Sph s = to_sph(p.x - c.x, p.y - c.y, p.z - c.z);
      double* row = tile + lane * ld;
      regular_harmonics<false>(P, s, -1.0, [&](int n, int m, double yr, double yi, double, double) {
        row[n * n + n + m] = q * yr;
        if (m > 0) row[n * n + n - m] = q * yi;
      });
    }
    __syncwarp();
#pragma unroll
    for (int i2 = 0; i2 < (FMMB_MAX_P * FMMB_MAX_P + 31) / 32; ++i2) {
      const int col = lane + 32 * i2;
      if (col < pp) {
        double sum = 0;
        for (int k = 0; k < cnt; ++k) sum += tile[k * ld + col];
        acc[i2] += sum;
      }
    }
  }
  double* Mb = M + (size_t)b * xstride(P);
#pragma unroll
  for (int i2 = 0; i2 < (FMMB_MAX_P * FMMB_MAX_P + 31) / 32; ++i2) {
    const int col = lane + 32 * i2;
    if (col < pp) Mb[col] = acc[i2];
  }
}

// ---- P2M with a narrow transposition tile ---------------------------------------------------------------------------
// p2m_kernel holds all P^2 values of 32 bodies in shared memory (16.6 KB per warp at P = 8: 12 warps per SM, FP64
// pipe 26 % busy, issue-latency bound).  Here the columns of the harmonics table are flushed in chunks of <= kP2MCols
// values as the m-major recurrence produces them: 6.4 KB per warp, more than twice the resident warps.  Same sums in
// the same order (bodies ascending per coefficient), so the multipoles are bit-identical to p2m_kernel's.
constexpr int kP2MCols = 24;
__global__ void __launch_bounds__(128)
p2m_cols_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                const unsigned* __restrict__ be, const double4* __restrict__ center,
                const double4* __restrict__ body, int P, double* __restrict__ M) {
  constexpr int ld = kP2MCols | 1;
  __shared__ double tiles[4][32 * ld];
  __shared__ int colidx[4][kP2MCols];
  const int pp = P * P;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double* tile = tiles[wl];
  int* cidx = colidx[wl];
  const int b = leaves[w];
  const double4 c = center[b];
  const unsigned b0 = bb[b], b1 = be[b];
  double* Mb = M + (size_t)b * xstride(P);
  bool first = true;
  for (unsigned base = b0; base < b1; base += 32) {
    const unsigned i = base + lane;
    const int cnt = (int)min(32u, b1 - base);
    const bool live = i < b1;
    const double4 p = live ? body[i] : make_double4(c.x + 0.125, c.y + 0.25, c.z + 0.5, 0.0);   // idle lanes: any regular point
    const double q = p.w;
    const Sph s = to_sph(p.x - c.x, p.y - c.y, p.z - c.z);
    double* row = tile + lane * ld;
    int pos = 0;                                    // values of the current chunk written so far (warp-uniform)
    auto flush = [&]() {
      __syncwarp();
      if (lane < pos) {
        double sum = 0;
        for (int k = 0; k < cnt; ++k) sum += tile[k * ld + lane];
        double* o = Mb + cidx[lane];
        *o = first ? sum : *o + sum;
      }
      __syncwarp();
      pos = 0;
    };
    regular_harmonics<false>(
        P, s, -1.0,
        [&](int n, int m, double yr, double yi, double, double) {
          row[pos] = q * yr;
          if (lane == 0) cidx[pos] = n * n + n + m;
          ++pos;
          if (m > 0) {
            row[pos] = q * yi;
            if (lane == 0) cidx[pos] = n * n + n - m;
            ++pos;
          }
        },
        [&](int m) {
          // room for the next column (2 (P - m - 1) values)?  otherwise flush what the tile holds
          const int next = m + 1 < P ? 2 * (P - m - 1) : kP2MCols + 1;
          if (pos + next > kP2MCols) flush();
        });
    first = false;
  }
  if (lane == 0 && xstride(P) > pp) Mb[pp] = 0.0;   // padding double of odd-sized expansions
}

// ---- M2M: block per parent box of one level; children accumulate in index order ----------------
__global__ void __launch_bounds__(64)
m2m_kernel(int lo, int hi, const int* __restrict__ box_list, const unsigned* __restrict__ key,
           const unsigned* __restrict__ cbegin, const unsigned* __restrict__ cend,
           const double4* __restrict__ center, int P, double* __restrict__ M) {
  int b = box_list ? box_list[blockIdx.x] : lo + blockIdx.x;
  if ((!box_list && b >= hi) || (key[b] >> 31)) return;   // leaves got their multipole from P2M
  extern __shared__ double2 sh[];
  const int nc = P * (P + 1) / 2, pp = P * P;
  double2* Y = sh;          // pp
  double2* Ms = sh + pp;    // nc
  double4 cpar = center[b];
  const int xs = xstride(P);
  for (int r = threadIdx.x; r < pp; r += blockDim.x) M[(size_t)b * xs + r] = 0.0;
  for (unsigned c = cbegin[b]; c < cend[b]; ++c) {
    double4 cc = center[c];
    Sph s = to_sph(cpar.x - cc.x, cpar.y - cc.y, cpar.z - cc.z);
    __syncthreads();
    for (int m = threadIdx.x; m < P; m += blockDim.x) harmonics_column<false>(m, P, s, -1.0, Y);
    for (int i = threadIdx.x; i < nc; i += blockDim.x) {
      int n, m;
      unpack_nm(i, n, m);
      Ms[i] = load_coef(M + (size_t)c * xs, n, m);
    }
    __syncthreads();
    for (int jks = threadIdx.x; jks < nc; jks += blockDim.x) {
      int j, k;
      unpack_nm(jks, j, k);
      double2 v = m2m_entry(Ms, Y, j, k);
      add_coef(M + (size_t)b * xs, j, k, v);
    }
  }
}

// ---- multi-GPU: multipoles of the boxes that straddle a partition cut ------------------------------
// One block per (straddling box, descendant) pair: direct M2M over any number of levels (M2M composes
// exactly), written to a scratch column; a second kernel sums the columns of a box in list order.
__global__ void __launch_bounds__(64)
m2m_direct_kernel(const int* __restrict__ pair_box, const int* __restrict__ strad_box,
                  const int* __restrict__ desc, const double4* __restrict__ center, int P,
                  const double* __restrict__ M, double* __restrict__ tmp) {
  extern __shared__ double2 sh[];
  const int nc = P * (P + 1) / 2, pp = P * P, xs = xstride(P);
  double2* Y = sh;
  double2* Ms = sh + pp;
  const int e = blockIdx.x;
  const int b = strad_box[pair_box[e]], d = desc[e];
  const double4 cb = center[b], cd = center[d];
  const Sph s = to_sph(cb.x - cd.x, cb.y - cd.y, cb.z - cd.z);
  for (int m = threadIdx.x; m < P; m += blockDim.x) harmonics_column<false>(m, P, s, -1.0, Y);
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    int n, m;
    unpack_nm(i, n, m);
    Ms[i] = load_coef(M + (size_t)d * xs, n, m);
  }
  __syncthreads();
  double* out = tmp + (size_t)e * xs;
  for (int jks = threadIdx.x; jks < nc; jks += blockDim.x) {
    int j, k;
    unpack_nm(jks, j, k);
    store_coef(out, j, k, m2m_entry(Ms, Y, j, k));
  }
}
__global__ void __launch_bounds__(64)
strad_reduce_kernel(const int* __restrict__ strad_box, const int* __restrict__ off, int P,
                    const double* __restrict__ tmp, double* __restrict__ M) {
  const int i = blockIdx.x, pp = P * P, xs = xstride(P);
  const int b = strad_box[i];
  for (int r = threadIdx.x; r < pp; r += blockDim.x) {
    double sum = 0;
    for (int e = off[i]; e < off[i + 1]; ++e) sum += tmp[(size_t)e * xs + r];
    M[(size_t)b * xs + r] = sum;
  }
}

// ---- M2L, generic per-pair path: block per target box, sources in LR_list order -----------------
// Creal[jks * P^2 + nm] = i^(|k-m|-|k|-|m|) (-1)^j Anm[nm] Anm[jk] / Anm[(j+n)^2+(j+n)+m-k]   (real)
__global__ void m2l_coeff_kernel(int P, double* __restrict__ C) {
  int pp = P * P, nc = P * (P + 1) / 2;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * pp) return;
  int jks = t / pp, nm = t % pp;
  int j = 0; while ((j + 1) * (j + 2) / 2 <= jks) ++j;
  int k = jks - j * (j + 1) / 2;
  int n = 0; while ((n + 1) * (n + 1) <= nm) ++n;
  int m = nm - n * n - n;
  int e = abs(k - m) - abs(k) - abs(m);          // always even
  double sgn = ((e / 2) & 1) ? -1.0 : 1.0;
  int jnkm = (j + n) * (j + n) + j + n + m - k;
  C[t] = sgn * oddeven(j) * c_anm[nm] * c_anm[j * j + j + k] / c_anm[jnkm];
}

__global__ void __launch_bounds__(256)
m2l_pair_kernel(int nboxes, const int* __restrict__ box_list, const int* __restrict__ off, const int* __restrict__ src,
                const double4* __restrict__ center, int P, const double* __restrict__ C,
                const double* __restrict__ M, double* __restrict__ L, int accumulate) {
  if (blockIdx.x >= nboxes) return;
  const int b = box_list ? box_list[blockIdx.x] : blockIdx.x;
  extern __shared__ double2 sh[];
  const int nc = P * (P + 1) / 2, pp = P * P;
  double2* Y = sh;                 // 4 pp
  double2* Mf = sh + 4 * pp;       // pp : full (-n..n) source multipole
  const int groups = blockDim.x / nc > 0 ? blockDim.x / nc : 1;
  const int g = threadIdx.x / nc, jks = threadIdx.x % nc;
  const bool worker = g < groups && blockDim.x >= nc;
  int j = 0, k = 0;
  if (worker) { while ((j + 1) * (j + 2) / 2 <= jks) ++j; k = jks - j * (j + 1) / 2; }
  const double* Crow = C + (size_t)jks * pp;
  double4 ct = center[b];
  double ar = 0, ai = 0;
  int s0 = off[b], s1 = off[b + 1];
  for (int it = s0; it < s1; ++it) {
    int sb = src[it];
    double4 cs = center[sb];
    Sph s = to_sph(ct.x - cs.x, ct.y - cs.y, ct.z - cs.z);
    __syncthreads();
    for (int m = threadIdx.x; m < 2 * P; m += blockDim.x) harmonics_column<true>(m, 2 * P, s, 1.0, Y);
    for (int nm = threadIdx.x; nm < pp; nm += blockDim.x) {
      int n = 0; while ((n + 1) * (n + 1) <= nm) ++n;
      int m = nm - n * n - n;
      double2 v = load_coef(M + (size_t)sb * xstride(P), n, abs(m));
      if (m < 0) v.y = -v.y;
      Mf[nm] = v;
    }
    __syncthreads();
    if (worker) {
      // group g takes every groups-th (n) row
      for (int n = g; n < P; n += groups) {
        int base = (j + n) * (j + n) + j + n - k;
        for (int m = -n; m <= n; ++m) {
          int nm = n * n + n + m;
          double c = Crow[nm];
          double2 a = Mf[nm], y = Y[base + m];
          ar += c * (a.x * y.x - a.y * y.y);
          ai += c * (a.x * y.y + a.y * y.x);
        }
      }
    }
  }
  // reduce the groups
  __syncthreads();
  double2* red = sh;
  if (worker) red[g * nc + jks] = make_double2(ar, ai);
  __syncthreads();
  if (threadIdx.x < nc) {
    double rr = 0, ri = 0;
    for (int q = 0; q < groups; ++q) { rr += red[q * nc + threadIdx.x].x; ri += red[q * nc + threadIdx.x].y; }
    double* Lb = L + (size_t)b * xstride(P);
    if (accumulate) add_coef(Lb, j, k, make_double2(rr, ri));
    else store_coef(Lb, j, k, make_double2(rr, ri));
  }
}

// ---- L2L: block per child box of one level -------------------------------------------------------
__global__ void __launch_bounds__(64)
l2l_kernel(int lo, int hi, const unsigned* __restrict__ parent, const unsigned char* __restrict__ has_local,
           const double4* __restrict__ center, int P, double* __restrict__ L) {
  int b = lo + blockIdx.x;
  if (b >= hi) return;
  int par = parent[b];
  if (!has_local[par]) return;
  extern __shared__ double2 sh[];
  const int nc = P * (P + 1) / 2, pp = P * P;
  double2* Y = sh;
  double2* Ls = sh + pp;
  double4 cc = center[b], cp = center[par];
  Sph s = to_sph(cc.x - cp.x, cc.y - cp.y, cc.z - cp.z);
  for (int m = threadIdx.x; m < P; m += blockDim.x) harmonics_column<false>(m, P, s, 1.0, Y);
  const int xs = xstride(P);
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    int n, m;
    unpack_nm(i, n, m);
    Ls[i] = load_coef(L + (size_t)par * xs, n, m);
  }
  __syncthreads();
  for (int jks = threadIdx.x; jks < nc; jks += blockDim.x) {
    int j, k;
    unpack_nm(jks, j, k);
    double2 v = l2l_entry(Ls, Y, j, k, P);
    add_coef(L + (size_t)b * xs, j, k, v);
  }
}

// ---- L2P: warp per leaf, lane per body ------------------------------------------------------------
__global__ void __launch_bounds__(128)
l2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
           const unsigned* __restrict__ be, const double4* __restrict__ center,
           const unsigned char* __restrict__ has_local, const double4* __restrict__ body, int P,
           const double* __restrict__ L, double4* __restrict__ res) {
  extern __shared__ double2 sh[];
  const int nc = P * (P + 1) / 2;
  int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  int b = leaves[w];
  unsigned b0 = bb[b], b1 = be[b];
  if (!has_local[b]) {
    for (unsigned i = b0 + lane; i < b1; i += 32) res[i] = make_double4(0, 0, 0, 0);
    return;
  }
  double2* Ls = sh + wl * nc;
  for (int i = lane; i < nc; i += 32) {
    int n, m;
    unpack_nm(i, n, m);
    Ls[i] = load_coef(L + (size_t)b * xstride(P), n, m);
  }
  __syncwarp();
  double4 c = center[b];
  for (unsigned i = b0 + lane; i < b1; i += 32) {
    double4 p = body[i];
    Sph s = to_sph(p.x - c.x, p.y - c.y, p.z - c.z);
    double pot = 0, s0 = 0, s1 = 0, s2 = 0;
    const double inv_r = 1.0 / s.r;
    regular_harmonics<true>(P, s, 1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
      double2 l = Ls[n * (n + 1) / 2 + m];
      double w2 = m == 0 ? 1.0 : 2.0;
      double re = l.x * yr - l.y * yi;             // Re(L Y)
      pot += w2 * re;
      s0 += w2 * re * inv_r * n;
      s1 += w2 * (l.x * tr - l.y * ti);            // Re(L Ytheta)
      s2 -= w2 * (l.x * yi + l.y * yr) * m;        // Re(L Y i) m = -Im(L Y) m
    });
    // sph2cart (:546-561): theta -> (s.x = cos, s.y = sin), phi -> (cp, sp)
    const double inv_ry = inv_r / s.y;
    double fx = s.y * s.cp * s0 + s.x * s.cp * inv_r * s1 - s.sp * inv_ry * s2;
    double fy = s.y * s.sp * s0 + s.x * s.sp * inv_r * s1 + s.cp * inv_ry * s2;
    double fz = s.x * s0 - s.y * inv_r * s1;
    res[i] = make_double4(pot, fx, fy, fz);
  }
}

// ---- L2P, leaves packed four to a warp -------------------------------------------------------------------------------
// l2p_kernel gives a warp one leaf: with ~30 bodies per leaf and leaves of up to 64 bodies the second pass over a leaf
// is mostly empty (22 of 32 lanes on average, ncu).  Here a warp takes kL2PLeaves consecutive leaves IN BODY ORDER --
// their bodies are one contiguous stretch of the tree-ordered array -- stages all their local expansions and walks the
// stretch 32 bodies at a time; a lane picks the expansion of the leaf its body lies in.  Same arithmetic per body.
constexpr int kL2PLeaves = 4;
__global__ void __launch_bounds__(128)
l2p_packed_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const double4* __restrict__ center,
                  const unsigned char* __restrict__ has_local, const double4* __restrict__ body, int P,
                  const double* __restrict__ L, double4* __restrict__ res) {
  extern __shared__ double2 sh[];
  __shared__ double4 s_cen[4][kL2PLeaves];
  __shared__ unsigned s_end[4][kL2PLeaves];
  __shared__ unsigned char s_has[4][kL2PLeaves];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  const int l0 = w * kL2PLeaves;
  if (l0 >= nleaves) return;
  const int nl = min(kL2PLeaves, nleaves - l0);
  double2* Ls = sh + (size_t)wl * kL2PLeaves * nc;
  if (lane < kL2PLeaves) {
    const int b = leaves[l0 + min(lane, nl - 1)];
    s_cen[wl][lane] = center[b];
    s_end[wl][lane] = lane < nl ? be[b] : 0xffffffffu;
    s_has[wl][lane] = has_local[b];
  }
  for (int j = 0; j < nl; ++j) {
    const int b = leaves[l0 + j];
    if (!has_local[b]) continue;
    for (int i = lane; i < nc; i += 32) {
      int n, m;
      unpack_nm(i, n, m);
      Ls[j * nc + i] = load_coef(L + (size_t)b * xstride(P), n, m);
    }
  }
  __syncwarp();
  const unsigned begin = bb[leaves[l0]], end = be[leaves[l0 + nl - 1]];
  const unsigned e0 = s_end[wl][0], e1 = s_end[wl][1], e2 = s_end[wl][2];
  for (unsigned i = begin + lane; i < end; i += 32) {
    const int j = (i >= e0) + (i >= e1) + (i >= e2);
    if (!s_has[wl][j]) { res[i] = make_double4(0, 0, 0, 0); continue; }
    const double4 c = s_cen[wl][j];
    const double2* Lj = Ls + j * nc;
    const double4 p = body[i];
    const Sph s = to_sph(p.x - c.x, p.y - c.y, p.z - c.z);
    double pot = 0, s0 = 0, s1 = 0, s2 = 0;
    const double inv_r = 1.0 / s.r;
    regular_harmonics<true>(P, s, 1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
      const double2 l = Lj[n * (n + 1) / 2 + m];
      const double w2 = m == 0 ? 1.0 : 2.0;
      const double re = l.x * yr - l.y * yi;             // Re(L Y)
      pot += w2 * re;
      s0 += w2 * re * inv_r * n;
      s1 += w2 * (l.x * tr - l.y * ti);            // Re(L Ytheta)
      s2 -= w2 * (l.x * yi + l.y * yr) * m;        // Re(L Y i) m = -Im(L Y) m
    });
    const double inv_ry = inv_r / s.y;
    const double fx = s.y * s.cp * s0 + s.x * s.cp * inv_r * s1 - s.sp * inv_ry * s2;
    const double fy = s.y * s.sp * s0 + s.x * s.sp * inv_r * s1 + s.cp * inv_ry * s2;
    const double fz = s.x * s0 - s.y * inv_r * s1;
    res[i] = make_double4(pot, fx, fy, fz);
  }
}

// ---- M2P (treecode, FMMOptions::TREECODE): warp per leaf, lane per body ---------------------------------------
// Every accepted pair (source box, target box) of the traversal evaluates the source multipole at the bodies of
// the target box (LaplaceSpherical.hpp:340-368 through EvalInteractionLazy.hpp:271-282).  A body receives from
// the lists of its leaf AND of every ancestor, so the warp walks up the parent chain; the source multipole is
// staged in a warp-private shared tile and each lane sums its own body in a fixed order (no atomics).
__global__ void __launch_bounds__(128)
m2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
           const unsigned* __restrict__ be, const unsigned* __restrict__ parent, const int* __restrict__ off,
           const int* __restrict__ src, const double4* __restrict__ center, const double4* __restrict__ body, int P,
           const double* __restrict__ M, double4* __restrict__ res) {
  extern __shared__ double2 sh[];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double2* Ms = sh + wl * nc;
  const int leaf = leaves[w];
  const unsigned b0 = bb[leaf], b1 = be[leaf];
  for (unsigned base = b0; base < b1; base += 32) {
    const unsigned i = base + lane;
    const bool act = i < b1;
    const double4 p = act ? body[i] : make_double4(0, 0, 0, 0);
    double pot = 0, fx = 0, fy = 0, fz = 0;
    for (int a = leaf;; a = (int)parent[a]) {
      for (int e = off[a]; e < off[a + 1]; ++e) {
        const int sb = src[e];
        __syncwarp();
        for (int k = lane; k < nc; k += 32) {
          int n, m;
          unpack_nm(k, n, m);
          Ms[k] = load_coef(M + (size_t)sb * xstride(P), n, m);
        }
        __syncwarp();
        if (act) {
          const double4 c = center[sb];
          const Sph s = to_sph(p.x - c.x, p.y - c.y, p.z - c.z);
          const double inv_r = 1.0 / s.r;
          double v = 0, s0 = 0, s1 = 0, s2 = 0;
          regular_harmonics<true, true>(P, s, 1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
            const double2 c2 = Ms[n * (n + 1) / 2 + m];
            const double w2 = m == 0 ? 1.0 : 2.0;
            const double re = w2 * (c2.x * yr - c2.y * yi);   // Re(M Y)
            v += re;
            s0 -= re * inv_r * (n + 1);
            s1 += w2 * (c2.x * tr - c2.y * ti);               // Re(M Ytheta)
            s2 -= w2 * (c2.x * yi + c2.y * yr) * m;           // Re(M Y i) m
          });
          const double inv_ry = inv_r / s.y;
          pot += v;
          fx += s.y * s.cp * s0 + s.x * s.cp * inv_r * s1 - s.sp * inv_ry * s2;
          fy += s.y * s.sp * s0 + s.x * s.sp * inv_r * s1 + s.cp * inv_ry * s2;
          fz += s.x * s0 - s.y * inv_r * s1;
        }
      }
      if (a == 0) break;
    }
    if (act) res[i] = make_double4(pot, fx, fy, fz);
  }
}

// ---- P2P: one warp per (target leaf, chunk of <= 32 targets) ---------------------------------------
// Sources stream through a warp-private shared tile (32 bodies = 1 KB), so there is no block
// barrier.  A chunk with r < 32 targets is replicated S = 32/r times across the lanes and every
// replica takes every S-th source; the replicas are summed with shuffles at the end.  That keeps
// the FP64 pipe fed for leaves whose body count is not a multiple of 32.
constexpr int kP2PWarps = 4;

__device__ __forceinline__ void p2p_accumulate(const double4 t, const double4 sq, double& pot, double& fx,
                                               double& fy, double& fz) {
  double dx = sq.x - t.x, dy = sq.y - t.y, dz = sq.z - t.z;
  double r2 = dx * dx + dy * dy + dz * dz;
  // 1/sqrt(r2): MUFU.RSQ64H seed (relative error < 2^-22) + one cubic step, branch free so that the
  // unrolled sources interleave.  r2 = 0 gives inf/NaN here and is discarded by the select below.
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
  double e = fma(-(r2 * y0), y0, 1.0);
  double inv = fma(y0 * e, fma(0.375, e, 0.5), y0);
  // LaplaceSpherical.hpp:158, R2 < 1e-8 -> 0.  r2 >= 0, so the ordering of the doubles is the ordering of their
  // bit patterns: the compare runs on the integer pipe and leaves the FP64 pipe to the arithmetic.
  if (__double_as_longlong(r2) < __double_as_longlong(1e-8)) inv = 0.0;
  double qi = sq.w * inv;
  double qi3 = qi * (inv * inv);
  pot += qi;
  fx = fma(dx, qi3, fx); fy = fma(dy, qi3, fy); fz = fma(dz, qi3, fz);
}

// work items: x = target box, y = first target body, z = number of targets (<= 32)
// mode 0: chunks of 32 and one remainder chunk (BEM: one cached block per chunk).
// mode 1: chunks of 32, then the remainder in power-of-two pieces: a piece of r = 2^k targets is replicated
//         32/r times across the lanes (source splitting), so every lane of every warp does useful pairs.
__global__ void p2p_count_items(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                                const unsigned* __restrict__ be, int mode, int chunk, int* __restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nleaves) return;
  int c = i < nleaves ? (int)(be[leaves[i]] - bb[leaves[i]]) : 0;
  cnt[i] = mode ? c / 32 + __popc(c % 32) : (c + chunk - 1) / chunk;
}
__global__ void p2p_fill_items(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                               const unsigned* __restrict__ be, const int* __restrict__ off,
                               const int* __restrict__ p2p_off, const int* __restrict__ p2p_src, int mode,
                               int chunk, int4* __restrict__ items, unsigned* __restrict__ work) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nleaves) return;
  int b = leaves[i];
  unsigned t0 = bb[b], t1 = be[b];
  int o = off[i];
  unsigned ns = 0;                       // source bodies of this leaf's list = sequential pair steps of a full chunk
  for (int e = p2p_off[b]; e < p2p_off[b + 1]; ++e) ns += be[p2p_src[e]] - bb[p2p_src[e]];
  unsigned t = t0;
  if (!mode) chunk = min(chunk, 32); else chunk = 32;
  for (; t + chunk <= t1 || (!mode && t < t1); t += chunk) {
    int r = (int)min((unsigned)chunk, t1 - t);
    work[o] = ns;
    items[o++] = make_int4(b, (int)t, r, 0);
  }
  if (mode)
    for (int r = 16; r >= 1; r >>= 1)
      if ((t1 - t) & r) {
        work[o] = (ns * r + 31) / 32;
        items[o++] = make_int4(b, (int)t, r, 0);
        t += r;
      }
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
p2p_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
           const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
           const double4* __restrict__ body, double4* __restrict__ res) {
  __shared__ double4 tiles[WARPS][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * WARPS + wl;
  if (item >= nitems) return;
  double4* tile = tiles[wl];
  const int4 it = items[item];
  const int r = it.z;                      // targets in this chunk
  const int S = 32 / r;                    // source splits (1 when r > 16)
  const int ti = lane % r, sp = lane / r;
  const bool act = sp < S;
  const double4 t = body[it.y + ti];
  double pot = 0, fx = 0, fy = 0, fz = 0;
  const int s0 = off[it.x], s1 = off[it.x + 1];
  for (int e = s0; e < s1; ++e) {
    const int sb = src[e];
    const unsigned c0 = bb[sb], c1 = be[sb];
    for (unsigned base = c0; base < c1; base += 32) {
      const int cnt = (int)min(32u, c1 - base);
      __syncwarp();
      if (lane < cnt) tile[lane] = body[base + lane];
      __syncwarp();
      if (act) {
        if (S == 1) {
#pragma unroll 4
          for (int k = 0; k < cnt; ++k) p2p_accumulate(t, tile[k], pot, fx, fy, fz);
        } else {
          for (int k = sp; k < cnt; k += S) p2p_accumulate(t, tile[k], pot, fx, fy, fz);
        }
      }
    }
  }
  if (S > 1) {
    // lanes ti, ti + r, ti + 2r, ... hold partial sums of the same target
    for (int q = 1; q < S; ++q) {
      int from = lane + q * r;
      double a = __shfl_sync(0xffffffffu, pot, from & 31), bx = __shfl_sync(0xffffffffu, fx, from & 31),
             by = __shfl_sync(0xffffffffu, fy, from & 31), bz = __shfl_sync(0xffffffffu, fz, from & 31);
      if (lane < r) { pot += a; fx += bx; fy += by; fz += bz; }
    }
  }
  if (lane < r) res[it.y + lane] = make_double4(pot, fx, fy, fz);
}

// ---- P2P over merged source runs --------------------------------------------------------------------
// Plan time: the source leaves of a target leaf are sorted by body index and adjacent body ranges are merged
// into runs (Morton neighbours are contiguous in the tree-ordered body array: ~27 leaves become ~10 runs).
// The kernel walks the runs as ONE virtual source stream: every tile holds exactly 32 sources (only the last
// one is padded, with zero-charge dummies far outside the domain, which contribute exactly 0), the pair loop
// has a fixed trip count (fully unrolled in groups, no remainder path), and the next tile is fetched into
// registers while the current one is being used.
__global__ void run_keys_kernel(const int* __restrict__ off, const int* __restrict__ src,
                                const unsigned* __restrict__ bb, int nb, unsigned long long* __restrict__ key,
                                int* __restrict__ val) {
  int b = blockIdx.x;
  if (b >= nb) return;
  for (int e = off[b] + threadIdx.x; e < off[b + 1]; e += blockDim.x) {
    key[e] = ((unsigned long long)b << 32) | bb[src[e]];
    val[e] = src[e];
  }
}
__global__ void run_flags_kernel(const unsigned long long* __restrict__ key, const int* __restrict__ val,
                                 const unsigned* __restrict__ bb, const unsigned* __restrict__ be, int64_t n,
                                 int* __restrict__ flag) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > n) return;
  if (e == n) { flag[e] = 0; return; }
  flag[e] = (e == 0 || (key[e] >> 32) != (key[e - 1] >> 32) || bb[val[e]] != be[val[e - 1]]) ? 1 : 0;
}
__global__ void run_fill_kernel(const int* __restrict__ val, const unsigned* __restrict__ bb,
                                const unsigned* __restrict__ be, const int* __restrict__ flag,
                                const int* __restrict__ pos, int64_t n, int2* __restrict__ runs) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (flag[e]) runs[pos[e]].x = (int)bb[val[e]];
  if (flag[e + 1] || e + 1 == n) runs[pos[e + 1] - 1].y = (int)be[val[e]];
}
__global__ void run_offsets_kernel(const int* __restrict__ off, const int* __restrict__ pos, int nb,
                                   int* __restrict__ run_off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= nb) run_off[b] = pos[off[b]];
}

template <int WARPS, int UNROLL>
__global__ void __launch_bounds__(32 * WARPS)
p2p_run_kernel(const int4* __restrict__ items, int nitems, const int* __restrict__ run_off,
               const int2* __restrict__ runs, const double4* __restrict__ body, double4 dummy,
               double4* __restrict__ res) {
  __shared__ double4 tiles[WARPS][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * WARPS + wl;
  if (item >= nitems) return;
  double4* tile = tiles[wl];
  const int4 it = items[item];
  const int r = it.z;                      // targets in this chunk
  const int S = 32 / r;                    // source splits (1 when r > 16)
  const int ti = lane % r, sp = lane / r;
  const bool act = sp < S;
  const double4 t = body[it.y + ti];
  double pot = 0, fx = 0, fy = 0, fz = 0;
  int j = run_off[it.x];
  const int j1 = run_off[it.x + 1];
  // cursor over the virtual source stream: current run [p, pend), descriptor of the following run prefetched
  int p = 0, pend = 0;
  int2 rnext = make_int2(0, 0);
  if (j < j1) { const int2 rr = runs[j]; p = rr.x; pend = rr.y; }
  if (j + 1 < j1) rnext = runs[j + 1];
  auto fetch = [&]() -> double4 {          // next 32 sources (warp-uniform control flow)
    double4 v = dummy;
    int filled = 0;
    while (filled < 32 && j < j1) {
      const int take = min(32 - filled, pend - p);
      const int l = lane - filled;
      if (l >= 0 && l < take) v = body[p + l];
      p += take; filled += take;
      if (p == pend) {
        ++j;
        p = rnext.x; pend = rnext.y;
        if (j + 1 < j1) rnext = runs[j + 1];
      }
    }
    return v;
  };
  bool have = j < j1;
  double4 nxt = dummy;
  if (have) nxt = fetch();
  while (have) {
    __syncwarp();
    tile[lane] = nxt;
    __syncwarp();
    have = j < j1;
    if (have) nxt = fetch();               // in flight while the current tile is consumed
    if (act) {
      if (S == 1) {
#pragma unroll UNROLL
        for (int k = 0; k < 32; ++k) p2p_accumulate(t, tile[k], pot, fx, fy, fz);
      } else {
#pragma unroll 2
        for (int k = sp; k < 32; k += S) p2p_accumulate(t, tile[k], pot, fx, fy, fz);
      }
    }
  }
  if (S > 1) {
    for (int q = 1; q < S; ++q) {
      int from = lane + q * r;
      double a = __shfl_sync(0xffffffffu, pot, from & 31), bx = __shfl_sync(0xffffffffu, fx, from & 31),
             by = __shfl_sync(0xffffffffu, fy, from & 31), bz = __shfl_sync(0xffffffffu, fz, from & 31);
      if (lane < r) { pot += a; fx += bx; fy += by; fz += bz; }
    }
  }
  if (lane < r) res[it.y + lane] = make_double4(pot, fx, fy, fz);
}

// ---- P2P, two targets per lane ------------------------------------------------------------------------
// Measured on B200 (scripts/micro/p2p_loop.cu): the pair loop is bound by FP64 issue (18 FP64 instructions per
// pair); everything else in the loop body costs issue slots on top.  Two measures cut that overhead:
//  * every lane owns TWO targets, so one shared-memory source load feeds two pairs.  A chunk of r targets uses
//    G = ceil(r/2) lanes per replica and S = 32/G replicas that split the sources (S = 2 for a full chunk);
//  * the R2 < 1e-8 select (LaplaceSpherical.hpp:158) only runs on tiles that can contain such a pair: tiles
//    that overlap the target leaf's own bodies (self pairs) and, for leaves that the plan found to have a
//    source of ANOTHER leaf closer than 1e-4 (p2p_close_kernel), every tile.  All other tiles take the
//    select-free loop.  r2 = 0 cannot occur there: equal positions have equal Morton codes, hence one leaf.
template <bool MASKED>
__device__ __forceinline__ void p2p_pair(const double tx, const double ty, const double tz, const double4 sq,
                                         double& pot, double& fx, double& fy, double& fz) {
  const double dx = sq.x - tx, dy = sq.y - ty, dz = sq.z - tz;
  const double r2 = dx * dx + dy * dy + dz * dz;
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
  const double e = fma(-(r2 * y0), y0, 1.0);
  double inv = fma(y0 * e, fma(0.375, e, 0.5), y0);
  if (MASKED) { if (__double_as_longlong(r2) < __double_as_longlong(1e-8)) inv = 0.0; }
  const double qi = sq.w * inv;
  const double qi3 = qi * (inv * inv);
  pot += qi;
  fx = fma(dx, qi3, fx); fy = fma(dy, qi3, fy); fz = fma(dz, qi3, fz);
}

// plan time: per target leaf, the number of (target, source) pairs with R2 < 1e-8 over its whole P2P list
__global__ void __launch_bounds__(128)
p2p_close_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                 const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
                 const double4* __restrict__ body, unsigned char* __restrict__ close_flag) {
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int b = leaves[w];
  const unsigned t0 = bb[b], t1 = be[b];
  int cnt = 0;
  for (unsigned ti = t0 + lane; ti < t1; ti += 32) {
    const double4 t = body[ti];
    for (int e = off[b]; e < off[b + 1]; ++e) {
      const int sb = src[e];
      if (sb == b) continue;
      for (unsigned k = bb[sb]; k < be[sb]; ++k) {
        const double4 s = body[k];
        const double dx = s.x - t.x, dy = s.y - t.y, dz = s.z - t.z;
        if (dx * dx + dy * dy + dz * dz < 1e-8) ++cnt;
      }
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) close_flag[b] = cnt > 0;
}

__global__ void p2p_items_ext_kernel(int4* __restrict__ items, int n, const unsigned* __restrict__ bb,
                                     const unsigned* __restrict__ be, const int* __restrict__ run_off,
                                     const int2* __restrict__ runs, const unsigned char* __restrict__ close_flag,
                                     int4* __restrict__ ext) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 it = items[i];
  const int b = it.x, j0 = run_off[b], j1 = run_off[b + 1];
  it.w = j0 < j1 ? runs[j0].x : 0;
  items[i] = it;
  ext[i] = make_int4(j0, j1, (int)bb[b], (int)((be[b] - bb[b]) << 1) | (close_flag[b] ? 1 : 0));
}

// PERSIST (option "p2p_wps"): the grid is a fixed number of one-warp blocks per SM that pull items from a counter.
// The near field then occupies a fixed share of every SM for its whole duration and the far-field kernels of the
// other stream find room beside it (a grid of 44 000 one-warp blocks refills every slot a finished block frees, so
// the large CTAs of the DMMA GEMM never accumulate the registers they need until the near field is done).  Measured:
// the far field does hide under the near field, but at 8 warps per SM the pair loop is latency-bound per warp and
// the matvec is not faster than with the plain grid -- kept as an option, not the default.
template <int UNROLL, bool PERSIST, int MINB = 1>
__global__ void __launch_bounds__(32, MINB)
p2p_pair2_kernel(const int4* __restrict__ items, const int4* __restrict__ items_ext, int nitems,
                 const int2* __restrict__ runs, const double4* __restrict__ body, double4 dummy,
                 double4* __restrict__ res, unsigned* __restrict__ counter) {
  __shared__ double4 tile[32];
  const int lane = threadIdx.x;
  for (;;) {
  int item = blockIdx.x;
  if (PERSIST) {
    if (lane == 0) item = (int)atomicAdd(counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
  }
  if (item >= nitems) return;
  const int4 it = items[item];
  const int r = it.z;                      // targets in this chunk (<= 32)
  const int G = (r + 1) >> 1;              // lanes per replica, two targets each
  const int S = 32 / G;                    // replicas: each takes every S-th source
  const int g = lane % G, sp = lane / G;
  const bool act = sp < S;
  const bool even_split = (32 % S) == 0;
  const int steps = 32 / S;
  const bool hasb = g + G < r;
  const double4 ta = body[it.y + g];
  const double4 tb = hasb ? body[it.y + g + G] : ta;
  double pa = 0, ax = 0, ay = 0, az = 0, pb = 0, bx = 0, by = 0, bz = 0;
  // everything needed to start streaming sits in the two item words: no dependent index loads
  const int4 ix = items_ext[item];         // x, y: run range; z: own bodies begin; w: (own count << 1) | close flag
  const int self0 = ix.z, self1 = ix.z + (ix.w >> 1);
  const bool all_masked = (ix.w & 1) != 0;
  int j = ix.x;
  const int j1 = ix.y;
  int p = 0, pend = 0;
  int2 rnext = make_int2(0, 0);
  if (j < j1) { p = it.w; pend = runs[j].y; }    // it.w = begin of the first run
  if (j + 1 < j1) rnext = runs[j + 1];
  bool mask_nxt = false;
  auto fetch = [&]() -> double4 {          // next 32 sources of the virtual stream (warp-uniform control flow)
    double4 v = dummy;
    int filled = 0;
    mask_nxt = all_masked;
    while (filled < 32 && j < j1) {
      const int take = min(32 - filled, pend - p);
      const int l = lane - filled;
      if (l >= 0 && l < take) v = body[p + l];
      mask_nxt = mask_nxt || (p < self1 && p + take > self0);
      p += take; filled += take;
      if (p == pend) {
        ++j;
        p = rnext.x; pend = rnext.y;
        if (j + 1 < j1) rnext = runs[j + 1];
      }
    }
    return v;
  };
  bool have = j < j1;
  double4 nxt = dummy;
  if (have) nxt = fetch();
  while (have) {
    __syncwarp();
    tile[lane] = nxt;
    const bool masked = mask_nxt;
    __syncwarp();
    have = j < j1;
    if (have) nxt = fetch();               // in flight while the current tile is consumed
    if (act) {
      if (S == 2) {                        // full chunks (17..32 targets): compile-time trip count
        const double4* ts = tile + sp;
        if (masked) {
#pragma unroll UNROLL
          for (int kk = 0; kk < 16; ++kk) {
            const double4 s = ts[2 * kk];
            p2p_pair<true>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair<true>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        } else {
#pragma unroll UNROLL
          for (int kk = 0; kk < 16; ++kk) {
            const double4 s = ts[2 * kk];
            p2p_pair<false>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair<false>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        }
      } else if (even_split) {             // S divides 32: every replica takes exactly 32 / S sources of the tile
        const double4* ts = tile + sp;
        if (masked) {
#pragma unroll UNROLL
          for (int kk = 0; kk < steps; ++kk) {
            const double4 s = ts[kk * S];
            p2p_pair<true>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair<true>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        } else {
#pragma unroll UNROLL
          for (int kk = 0; kk < steps; ++kk) {
            const double4 s = ts[kk * S];
            p2p_pair<false>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair<false>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        }
      } else if (masked) {
#pragma unroll 2
        for (int k = sp; k < 32; k += S) {
          const double4 s = tile[k];
          p2p_pair<true>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
          p2p_pair<true>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
        }
      } else {
#pragma unroll 2
        for (int k = sp; k < 32; k += S) {
          const double4 s = tile[k];
          p2p_pair<false>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
          p2p_pair<false>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
        }
      }
    }
  }
  // replicas sp = 1..S-1 hold partial sums of the same targets as replica 0
  for (int q = 1; q < S; ++q) {
    const int from = (lane + q * G) & 31;
    const double a0 = __shfl_sync(0xffffffffu, pa, from), a1 = __shfl_sync(0xffffffffu, ax, from),
                 a2 = __shfl_sync(0xffffffffu, ay, from), a3 = __shfl_sync(0xffffffffu, az, from),
                 b0 = __shfl_sync(0xffffffffu, pb, from), b1 = __shfl_sync(0xffffffffu, bx, from),
                 b2 = __shfl_sync(0xffffffffu, by, from), b3 = __shfl_sync(0xffffffffu, bz, from);
    if (lane < G) { pa += a0; ax += a1; ay += a2; az += a3; pb += b0; bx += b1; by += b2; bz += b3; }
  }
  if (lane < G) {
    res[it.y + g] = make_double4(pa, ax, ay, az);
    if (hasb) res[it.y + g + G] = make_double4(pb, bx, by, bz);
  }
  if (!PERSIST) return;
  __syncwarp();                                              // the tile is reused by the next item
  }
}

// ---- P2P, source tiles staged by TMA -----------------------------------------------------------------------------
// The same work decomposition and pair loop as p2p_pair2_kernel; what changes is how a 32-source tile reaches
// shared memory.  The merged source runs are contiguous stretches of the tree-ordered body array, so a tile is one
// to three 1-D bulk copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP) issued by one lane, completion
// counted in bytes on an mbarrier: no LDG -> register -> STS round trip (the 10 M shared-store bank conflicts and
// the 8 staging registers per lane of p2p_pair2_kernel), and the next tile is in flight while this one is consumed.
// NEWTON: one Newton step on the MUFU.RSQ64H seed instead of the cubic step -- 16 instead of 18 FP64 instructions
// per pair; the inverse root is then accurate to 3/8 e^2 <= 1.3e-12 relative (measured max |e| = 1.86e-6 over
// 2.7e8 arguments, scripts/micro/probe_r2.cu), inside the 1e-10 parity tolerance but not at round-off.
__device__ __forceinline__ void p2p_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void p2p_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void p2p_mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
      " bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void p2p_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
               "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <bool MASKED, bool NEWTON>
__device__ __forceinline__ void p2p_pair_n(const double tx, const double ty, const double tz, const double4 sq,
                                           double& pot, double& fx, double& fy, double& fz) {
  const double dx = sq.x - tx, dy = sq.y - ty, dz = sq.z - tz;
  const double r2 = dx * dx + dy * dy + dz * dz;
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
  const double e = fma(-(r2 * y0), y0, 1.0);
  double inv;
  if (NEWTON) {
    // y0 / 2 by an exponent decrement on the integer pipe (y0 is a normal number for every r2 in range)
    const double h = __hiloint2double(__double2hiint(y0) - 0x00100000, __double2loint(y0));
    inv = fma(h, e, y0);
  } else {
    inv = fma(y0 * e, fma(0.375, e, 0.5), y0);
  }
  if (MASKED) { if (__double_as_longlong(r2) < __double_as_longlong(1e-8)) inv = 0.0; }
  const double qi = sq.w * inv;
  const double qi3 = qi * (inv * inv);
  pot += qi;
  fx = fma(dx, qi3, fx); fy = fma(dy, qi3, fy); fz = fma(dz, qi3, fz);
}

template <int UNROLL, bool NEWTON>
__global__ void __launch_bounds__(32)
p2p_tma_kernel(const int4* __restrict__ items, const int4* __restrict__ items_ext, int nitems,
               const int2* __restrict__ runs, const double4* __restrict__ body, double4 dummy,
               double4* __restrict__ res) {
  __shared__ __align__(128) double4 tile[2][32];
  __shared__ __align__(8) uint64_t bar[2];
  const int lane = threadIdx.x;
  const int item = blockIdx.x;
  if (item >= nitems) return;
  if (lane == 0) { p2p_mbar_init(&bar[0], 1); p2p_mbar_init(&bar[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int4 it = items[item];
  const int r = it.z;                      // targets in this chunk (<= 32)
  const int G = (r + 1) >> 1;              // lanes per replica, two targets each
  const int S = 32 / G;                    // replicas: each takes every S-th source
  const int g = lane % G, sp = lane / G;
  const bool act = sp < S;
  const bool even_split = (32 % S) == 0;
  const int steps = 32 / S;
  const bool hasb = g + G < r;
  const double4 ta = body[it.y + g];
  const double4 tb = hasb ? body[it.y + g + G] : ta;
  double pa = 0, ax = 0, ay = 0, az = 0, pb = 0, bx = 0, by = 0, bz = 0;
  const int4 ix = items_ext[item];         // x, y: run range; z: own bodies begin; w: (own count << 1) | close flag
  const int self0 = ix.z, self1 = ix.z + (ix.w >> 1);
  const bool all_masked = (ix.w & 1) != 0;
  int j = ix.x;
  const int j1 = ix.y;
  // length of the virtual source stream (sum of the run lengths): lanes add up the runs
  int total = 0;
  for (int k = j + lane; k < j1; k += 32) { const int2 rr = runs[k]; total += rr.y - rr.x; }
  total = __reduce_add_sync(0xffffffffu, total);
  const int ntiles = (total + 31) >> 5;
  int p = 0, pend = 0;
  int2 rnext = make_int2(0, 0);
  if (j < j1) { p = it.w; pend = runs[j].y; }    // it.w = begin of the first run
  if (j + 1 < j1) rnext = runs[j + 1];
  // tile t -> stage t & 1: one to three bulk copies (pieces of consecutive runs), issued by lane 0; every lane
  // advances the cursor and learns whether the tile can hold a pair with R2 < 1e-8
  auto issue = [&](int t) -> bool {
    const int s = t & 1;
    const int cnt = min(32, total - 32 * t);
    bool masked = all_masked;
    if (lane == 0) p2p_mbar_expect_tx(&bar[s], (unsigned)cnt * 32u);
    int filled = 0;
    while (filled < cnt) {
      const int take = min(cnt - filled, pend - p);
      if (lane == 0) p2p_bulk_g2s(&tile[s][filled], body + p, (unsigned)take * 32u, &bar[s]);
      masked = masked || (p < self1 && p + take > self0);
      p += take; filled += take;
      if (p == pend) {
        ++j;
        p = rnext.x; pend = rnext.y;
        if (j + 1 < j1) rnext = runs[j + 1];
      }
    }
    if (lane >= cnt) tile[s][lane] = dummy;      // last tile: zero-charge sources far outside the domain
    return masked;
  };
  bool mask_cur = false, mask_nxt = false;
  if (ntiles > 0) mask_cur = issue(0);
  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) mask_nxt = issue(t + 1);           // in flight while tile t is consumed
    p2p_mbar_wait(&bar[t & 1], (unsigned)(t >> 1) & 1u);
    __syncwarp();                                          // the dummy stores of the last tile
    const double4* tl = tile[t & 1];
    const bool masked = mask_cur;
    if (act) {
      if (S == 2) {                        // full chunks (17..32 targets): compile-time trip count
        const double4* ts = tl + sp;
        if (masked) {
#pragma unroll UNROLL
          for (int kk = 0; kk < 16; ++kk) {
            const double4 s = ts[2 * kk];
            p2p_pair_n<true, NEWTON>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair_n<true, NEWTON>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        } else {
#pragma unroll UNROLL
          for (int kk = 0; kk < 16; ++kk) {
            const double4 s = ts[2 * kk];
            p2p_pair_n<false, NEWTON>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair_n<false, NEWTON>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        }
      } else if (even_split) {             // S divides 32: every replica takes exactly 32 / S sources of the tile
        const double4* ts = tl + sp;
        if (masked) {
#pragma unroll UNROLL
          for (int kk = 0; kk < steps; ++kk) {
            const double4 s = ts[kk * S];
            p2p_pair_n<true, NEWTON>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair_n<true, NEWTON>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        } else {
#pragma unroll UNROLL
          for (int kk = 0; kk < steps; ++kk) {
            const double4 s = ts[kk * S];
            p2p_pair_n<false, NEWTON>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
            p2p_pair_n<false, NEWTON>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
          }
        }
      } else if (masked) {
#pragma unroll 2
        for (int k = sp; k < 32; k += S) {
          const double4 s = tl[k];
          p2p_pair_n<true, NEWTON>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
          p2p_pair_n<true, NEWTON>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
        }
      } else {
#pragma unroll 2
        for (int k = sp; k < 32; k += S) {
          const double4 s = tl[k];
          p2p_pair_n<false, NEWTON>(ta.x, ta.y, ta.z, s, pa, ax, ay, az);
          p2p_pair_n<false, NEWTON>(tb.x, tb.y, tb.z, s, pb, bx, by, bz);
        }
      }
    }
    mask_cur = mask_nxt;
    __syncwarp();                                          // everyone is done with the stage before it is refilled
  }
  // replicas sp = 1..S-1 hold partial sums of the same targets as replica 0
  for (int q = 1; q < S; ++q) {
    const int from = (lane + q * G) & 31;
    const double a0 = __shfl_sync(0xffffffffu, pa, from), a1 = __shfl_sync(0xffffffffu, ax, from),
                 a2 = __shfl_sync(0xffffffffu, ay, from), a3 = __shfl_sync(0xffffffffu, az, from),
                 b0 = __shfl_sync(0xffffffffu, pb, from), b1 = __shfl_sync(0xffffffffu, bx, from),
                 b2 = __shfl_sync(0xffffffffu, by, from), b3 = __shfl_sync(0xffffffffu, bz, from);
    if (lane < G) { pa += a0; ax += a1; ay += a2; az += a3; pb += b0; bx += b1; by += b2; bz += b3; }
  }
  if (lane < G) {
    res[it.y + g] = make_double4(pa, ax, ay, az);
    if (hasb) res[it.y + g + G] = make_double4(pb, bx, by, bz);
  }
}

// ---- results back to the caller's order ----------------------------------------------------------
__global__ void scatter_results(const double4* __restrict__ near, const double4* __restrict__ far,
                                const unsigned* __restrict__ perm, int64_t i0, int64_t i1,
                                double4* __restrict__ out) {
  int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= i1) return;
  double4 a = near[i], b = far[i];
  out[perm[i]] = make_double4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
// sharded call: charges arrive as per-rank slices in tree order (padded all-gather staging) -> body[].w
__global__ void place_charges(const double* __restrict__ stage, const long long* __restrict__ cuts, int nranks,
                              long long chunk, double4* __restrict__ body) {
  const int q = blockIdx.y;
  const long long b0 = cuts[q], len = cuts[q + 1] - b0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
    body[b0 + i].w = stage[(size_t)q * chunk + i];
}
__global__ void place_charges_local(const double* __restrict__ q, int64_t n, double4* __restrict__ body) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) body[i].w = q[i];
}
// sharded call: near + far of the owned range, written as the rank's result slice (tree order)
__global__ void combine_slice(const double4* __restrict__ near, const double4* __restrict__ far, int64_t i0,
                              int64_t i1, double4* __restrict__ out) {
  int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= i1) return;
  double4 a = near[i], b = far[i];
  out[i - i0] = make_double4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
// multi-GPU: near + far of the owned range in tree order, ready for the all-gather
__global__ void combine_results(const double4* __restrict__ near, const double4* __restrict__ far, int64_t i0,
                                int64_t i1, double4* __restrict__ tree) {
  int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= i1) return;
  double4 a = near[i], b = far[i];
  tree[i] = make_double4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__global__ void scatter_tree(const double4* __restrict__ tree, const unsigned* __restrict__ perm, int64_t n,
                             double4* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[perm[i]] = tree[i];
}

// ---- brute force (Direct.hpp:99-125) ---------------------------------------------------------------
__global__ void __launch_bounds__(128)
direct_kernel(const double* __restrict__ spts, const double* __restrict__ q, int64_t ns,
              const double* __restrict__ tpts, int64_t nt, double4* __restrict__ out) {
  __shared__ double4 tile[128];
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool act = i < nt;
  double4 t = act ? make_double4(tpts[3 * i], tpts[3 * i + 1], tpts[3 * i + 2], 0) : make_double4(0, 0, 0, 0);
  double pot = 0, fx = 0, fy = 0, fz = 0;
  for (int64_t base = 0; base < ns; base += 128) {
    int64_t j = base + threadIdx.x;
    __syncthreads();
    tile[threadIdx.x] = j < ns ? make_double4(spts[3 * j], spts[3 * j + 1], spts[3 * j + 2], q[j])
                               : make_double4(1e30, 1e30, 1e30, 0);
    __syncthreads();
    int cnt = (int)min((int64_t)128, ns - base);
    if (act)
      for (int k = 0; k < cnt; ++k) p2p_accumulate(t, tile[k], pot, fx, fy, fz);
  }
  if (act) out[i] = make_double4(pot, fx, fy, fz);
}

// ---- FP64 peaks: independent chains of DFMA, and of DMMA.8x8x4 (FP64 tensor core) ------------------
__global__ void __launch_bounds__(256)
dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256)
dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

}  // namespace

void laplace_init_tables(fmmb_plan* plan) {
  upload_laplace_tables();
  m2l_init_tables();
  plan->tab.pmax = FMMB_MAX_P;
}

// per-order real M2L coefficient table, cached on the plan
static const double* m2l_coeffs(fmmb_plan* plan, int P) {
  auto it = plan->m2l_coeff.find(P);
  if (it != plan->m2l_coeff.end()) return it->second->p;
  DevBuf<double>* buf = new DevBuf<double>();
  plan->m2l_coeff[P] = buf;
  int cnt = P * (P + 1) / 2 * P * P;
  buf->resize(cnt);
  m2l_coeff_kernel<<<nblk(cnt, 256), 256, 0, plan->stream>>>(P, buf->p);
  FMMB_CUDA(cudaGetLastError());
  return buf->p;
}

// Which far-field engine runs.  fmmb_options.m2l_mode: 1 = the per-pair kernels above; 2 = the class-major GEMM +
// column reduction of m2l_classes.cu (orders <= 8; per-pair kernels above that); 3 = the output-stationary fused sweep
// of trans_blocked.cu (all orders); 0 (auto) = whichever is faster on a B200: measured at N = 1M, P = 8 the
// class-major engine runs the far field in 1.72 ms (4.5 GB of DRAM traffic through its 2 GB column scratch), the
// fused sweep in 1.87 ms (0.14 GB, no scratch, one launch), so auto takes 2 for orders <= 8 and 3 for orders 9..16,
// where the class-major GEMM does not exist.
static int far_engine(const fmmb_plan* plan) {
  const int m = plan->opts.m2l_mode;
  if (m == 1) return 1;
  if (m == 3) return 3;
  if (m == 2) return plan->p <= 8 ? 2 : 1;
  return plan->p <= 8 ? 2 : 3;
}

void laplace_build_far(fmmb_plan* plan) {
  if (plan->near_only) return;
  const int e = far_engine(plan);
  if (e == 2 && !plan->far_built_classes) { build_m2l_classes(plan); plan->far_built_classes = true; }
  if (e == 3 && !plan->far_built_blocked) { build_blocked_batches(plan); plan->far_built_blocked = true; }
}

// the few boxes that straddle a partition cut, from their maximal single-rank descendants (per-pair engines)
static void straddlers_direct(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P;
  if (!T.n_strad_pairs) return;
  const size_t sh_mm = (size_t)(pp + nc) * sizeof(double2);
  T.strad_tmp.resize((size_t)T.n_strad_pairs * xstride(P));
  m2m_direct_kernel<<<T.n_strad_pairs, 64, sh_mm, s>>>(T.strad_pair_box.p, T.strad_box.p, T.strad_desc.p, T.center.p,
                                                      P, plan->M.p, T.strad_tmp.p);
  strad_reduce_kernel<<<T.n_strad, 64, 0, s>>>(T.strad_box.p, T.strad_off.p, P, T.strad_tmp.p, plan->M.p);
  plan->launches += 2;
}

// M2M sweep -> M2L -> L2L sweep on plan->M / plan->L at the current order (expects the leaf multipoles
// in plan->M; leaves the complete local expansions in plan->L).  Records ev[2] / ev[3] around M2L.
void laplace_translations(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P;
  const int nb = T.nboxes;
  cudaEvent_t* ev = plan->ev;
  const size_t sh_mm = (size_t)(pp + nc) * sizeof(double2);
  const int engine = far_engine(plan);
  if (!plan->capturing) laplace_build_far(plan);
  // multi-GPU: upward pass inside the owned subtrees, exchange, then the boxes that straddle a cut.  The gate is
  // the same on every rank (Tree::owned_upward): a collective follows.
  const bool owned_up = laplace_owned_upward(plan);
  auto exchange = [&] {
    if (plan->hook_after_owned_m2m) plan->hook_after_owned_m2m();   // laplace_execute: start the near field now
    if (plan->peer_ready && plan->kind == FMMB_LAPLACE_SPHERICAL) exchange_multipoles_peer(plan, s);
    else exchange_multipoles(plan, s);
  };

  // ---- engine 3: the whole far field is one launch (two around the exchange on a multi-GPU plan); the phases
  // inside it are ordered by device-side counters (trans_blocked.cu)
  if (engine == 3) {
    const bool tree_only = plan->opts.evaluator == FMMB_EVAL_TREECODE;
    // every target box is written exactly once by the block that owns it; boxes without M2L pairs stay zero
    if (!tree_only) FMMB_CUDA(cudaMemsetAsync(plan->L.p, 0, (size_t)nb * xstride(P) * sizeof(double), s));
    if (!plan->capturing) FMMB_CUDA(cudaEventRecord(plan->ev[13], s));
    if (owned_up) {
      run_sweep(plan, plan_sweep(plan, 2), s);
      exchange();
      run_sweep(plan, plan_sweep(plan, tree_only ? 4 : 3), s);
    } else {
      run_sweep(plan, plan_sweep(plan, tree_only ? 1 : 0), s);
    }
    if (!plan->capturing) {
      FMMB_CUDA(cudaEventRecord(plan->ev[14], s));
      FMMB_CUDA(cudaEventRecord(ev[2], s));
      FMMB_CUDA(cudaEventRecord(ev[3], s));
    }
    plan->m2l_gemm_timed = true;
    return;
  }
  // ---- upward sweep
  {
    bool up_done = false;
    if (owned_up) {
      m2m_batched(plan, s, /*owned_only=*/true);
      exchange();
      straddlers_direct(plan, s);
      up_done = true;
    } else if (engine == 2) {
      up_done = m2m_batched(plan, s);
    }
    for (int l = T.nlevels - 2; l >= 0 && !up_done; --l) {
      int lo = T.level_off[l], hi = T.level_off[l + 1];
      m2m_kernel<<<hi - lo, 64, sh_mm, s>>>(lo, hi, nullptr, T.key.p, T.cbegin.p, T.cend.p, T.center.p, P, plan->M.p);
      ++plan->launches;
    }
  }
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[2], s));
  if (plan->opts.evaluator == FMMB_EVAL_TREECODE) {        // treecode: multipoles only, M2P does the rest
    if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[3], s));
    return;
  }

  // ---- far field translations
  {
    const double* C = m2l_coeffs(plan, P);
    int threads = 128;
    while (threads < nc) threads += 32;
    size_t sh = (size_t)(5 * pp) * sizeof(double2);
    size_t red = (size_t)(threads / nc) * nc * sizeof(double2);
    if (red > sh) sh = red;
    bool batched = engine == 2 && m2l_batched(plan, s);
    if (!batched) {
      m2l_pair_kernel<<<nb, threads, sh, s>>>(nb, nullptr, T.m2l_off.p, T.m2l_src.p, T.center.p, P, C, plan->M.p,
                                             plan->L.p, 0);
      ++plan->launches;
    } else if (plan->cls.n_res > 0) {
      m2l_pair_kernel<<<plan->cls.n_res_boxes, threads, sh, s>>>(plan->cls.n_res_boxes, plan->cls.res_boxes.p,
                                                                plan->cls.res_off.p, plan->cls.res_src.p,
                                                                T.center.p, P, C, plan->M.p, plan->L.p, 1);
      ++plan->launches;
    }
  }
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[3], s));

  // ---- downward sweep
  {
    const bool down_batched = engine == 2 && l2l_batched(plan, s);
    for (int l = 1; l < T.nlevels && !down_batched; ++l) {
      int lo = T.level_off[l], hi = T.level_off[l + 1];
      l2l_kernel<<<hi - lo, 64, sh_mm, s>>>(lo, hi, T.parent.p, T.has_local.p, T.center.p, P, plan->L.p);
      ++plan->launches;
    }
  }
}

// The owned upward pass (P2M and M2M inside the rank's subtrees, multipole exchange, straddling boxes) runs when the
// plan is one of several ranks with an exchange path and some rank owns a parent box -- facts every rank agrees on.
bool laplace_owned_upward(const fmmb_plan* plan) {
  const Tree& T = plan->tree;
  const int engine = far_engine(plan);
  return T.nranks > 1 && (plan->comm || plan->peer_ready) && T.owned_upward && engine != 1 && (engine == 3 || plan->p <= 8);
}

// Sizes the expansion buffers for the current order (real layout, see laplace_ops.cuh).
void laplace_prepare_expansions(fmmb_plan* plan) {
  const int P = plan->p, pp = P * P;
  const int xs = (pp + 1) & ~1;
  // one more expansion than boxes: row nboxes stays all zero (the source of absent pairs in trans_blocked.cu)
  plan->M.resize((size_t)(plan->tree.nboxes + 1) * xs);
  plan->L.resize((size_t)(plan->tree.nboxes + 1) * xs);
  if (plan->p_alloc != P) {
    // the padding double of odd-sized expansions is read (times zero) by the GEMM: keep it finite.
    // An exported (peer) multipole array was zeroed once and is never cleared again: a faster peer may already be
    // writing the next matvec's rows into it.
    if (!plan->peer_alloc) plan->M.zero(plan->stream);
    else FMMB_CUDA(cudaMemsetAsync(plan->M.p + (size_t)plan->tree.nboxes * xs, 0, (size_t)xs * sizeof(double), plan->stream));
    plan->L.zero(plan->stream);
    plan->p_alloc = P;
  }
}

// Near-field launch on stream s2 (after the charges are in place, event ev[1]); records ev[6] / ev[7] around it.
static void launch_near_field(fmmb_plan* plan, cudaStream_t s, cudaStream_t s2) {
  Tree& T = plan->tree;
  cudaEvent_t* ev = plan->ev;
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s2, ev[1], 0));
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[6], s2));
  // zero-charge padding sources for the last tile of a source stream, far outside the domain
  const double ext = 1024.0 * std::max(T.cell[0], std::max(T.cell[1], T.cell[2]));
  const double4 dummy = make_double4(T.pmin[0] - 1000.0 * ext, T.pmin[1] - 1000.0 * ext, T.pmin[2] - 1000.0 * ext, 0.0);
  const int ni = T.n_p2p_items;
  const int w = plan->p2p_warps, u = plan->p2p_unroll;
#define FMMB_P2P_RUN(W, U)                                                                                        \
  p2p_run_kernel<W, U><<<nblk(ni, W), 32 * W, 0, s2>>>(T.p2p_items.p, ni, T.p2p_run_off.p, T.p2p_runs.p, T.body.p, \
                                                       dummy, plan->res_near.p)
#define FMMB_P2P_LEAF(W)                                                                                           \
  p2p_kernel<W><<<nblk(ni, W), 32 * W, 0, s2>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p, T.p2p_off.p, T.p2p_src.p, \
                                                T.body.p, plan->res_near.p)
  if (ni > 0) {
    if (plan->p2p_kernel == 3) {                    // two targets per lane, merged runs, tiles staged by TMA
      if (plan->p2p_newton)
        p2p_tma_kernel<4, true><<<ni, 32, 0, s2>>>(T.p2p_items.p, T.p2p_items_ext.p, ni, T.p2p_runs.p, T.body.p, dummy,
                                                   plan->res_near.p);
      else
        p2p_tma_kernel<4, false><<<ni, 32, 0, s2>>>(T.p2p_items.p, T.p2p_items_ext.p, ni, T.p2p_runs.p, T.body.p, dummy,
                                                    plan->res_near.p);
    } else if (plan->p2p_kernel == 2 && plan->p2p_wps > 0 && s2 != s) {
      // default when the near field runs beside the far field: p2p_wps persistent one-warp blocks per SM
      int sms = 148;
      FMMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device));
      const int wps = plan->p2p_wps;
      const int grid = std::min(ni, sms * wps);
      // no residency cap beyond the grid size: the launch finds the GPU (nearly) empty, and the block scheduler
      // deals a grid of sms x wps small blocks evenly over the SMs.  (Capping the residency with a dynamic shared
      // memory request was tried: it takes the shared memory the far-field kernels need.)
      const size_t cap = 0;
      plan->p2p_counter.resize(1);
      FMMB_CUDA(cudaMemsetAsync(plan->p2p_counter.p, 0, sizeof(unsigned), s2));
      // the 72-register build: 14 resident blocks leave half of the register file to the far-field kernels
      p2p_pair2_kernel<4, true, 28><<<grid, 32, cap, s2>>>(T.p2p_items.p, T.p2p_items_ext.p, ni, T.p2p_runs.p, T.body.p,
                                                           dummy, plan->res_near.p, plan->p2p_counter.p);
    } else if (plan->p2p_kernel == 2 && plan->p2p_occ > 20) {   // the same kernel compiled for more resident warps
#define FMMB_P2P_OCC(B)                                                                                              \
      p2p_pair2_kernel<4, false, B><<<ni, 32, 0, s2>>>(T.p2p_items.p, T.p2p_items_ext.p, ni, T.p2p_runs.p, T.body.p, \
                                                       dummy, plan->res_near.p, nullptr)
      if (plan->p2p_occ >= 32) FMMB_P2P_OCC(32);
      else if (plan->p2p_occ >= 28) FMMB_P2P_OCC(28);
      else FMMB_P2P_OCC(24);
#undef FMMB_P2P_OCC
    } else if (plan->p2p_kernel == 2) {             // two targets per lane, merged runs, register prefetch
      if (u == 8)
        p2p_pair2_kernel<8, false><<<ni, 32, 0, s2>>>(T.p2p_items.p, T.p2p_items_ext.p, ni, T.p2p_runs.p, T.body.p, dummy,
                                                      plan->res_near.p, nullptr);
      else
        p2p_pair2_kernel<4, false><<<ni, 32, 0, s2>>>(T.p2p_items.p, T.p2p_items_ext.p, ni, T.p2p_runs.p, T.body.p, dummy,
                                                      plan->res_near.p, nullptr);
    } else if (plan->p2p_kernel == 1) {             // one target per lane over merged runs
      if (w == 1 && u == 4) FMMB_P2P_RUN(1, 4);
      else if (w == 1) FMMB_P2P_RUN(1, 8);
      else if (w == 2 && u == 4) FMMB_P2P_RUN(2, 4);
      else if (w == 2) FMMB_P2P_RUN(2, 8);
      else if (u == 4) FMMB_P2P_RUN(4, 4);
      else FMMB_P2P_RUN(4, 8);
    } else {                                        // one tile per source leaf
      if (w == 1) FMMB_P2P_LEAF(1);
      else if (w == 2) FMMB_P2P_LEAF(2);
      else FMMB_P2P_LEAF(kP2PWarps);
    }
  }
#undef FMMB_P2P_RUN
#undef FMMB_P2P_LEAF
  ++plan->launches;
  FMMB_CUDA(cudaEventRecord(ev[7], s2));
}

// Results of this rank (near + far, tree order) to the caller: sharded slice, all-gathered full vector, or the
// owned part of the full vector when there is no communicator.
static void deliver_results(fmmb_plan* plan, double* d_results, cudaStream_t s) {
  Tree& T = plan->tree;
  const int64_t n = T.n, own = T.own_b1 - T.own_b0;
  double4* out = reinterpret_cast<double4*>(d_results);
  if (plan->call_sharded) {
    // results stay sharded by target: the rank's slice in tree order, no collective
    if (own > 0)
      combine_slice<<<nblk(own, 256), 256, 0, s>>>(plan->res_near.p, plan->res_far.p, T.own_b0, T.own_b1, out);
  } else if (T.nranks > 1 && plan->comm) {
    // all-gather the per-rank result slices (tree order), then un-permute;
    // + pad: the padded all-gather reads a full chunk starting at the owned slice
    long long chunk = 0;
    for (int q = 0; q < T.nranks; ++q) chunk = std::max<long long>(chunk, T.body_cuts[q + 1] - T.body_cuts[q]);
    plan->res_tree.resize((size_t)n + (size_t)chunk);
    if (own > 0)
      combine_results<<<nblk(own, 256), 256, 0, s>>>(plan->res_near.p, plan->res_far.p, T.own_b0, T.own_b1,
                                                    plan->res_tree.p);
    allgather_results(plan, s);
    scatter_tree<<<nblk(n, 256), 256, 0, s>>>(plan->res_tree.p, T.perm.p, n, out);
    ++plan->launches;
  } else if (own > 0) {
    scatter_results<<<nblk(own, 256), 256, 0, s>>>(plan->res_near.p, plan->res_far.p, T.perm.p, T.own_b0, T.own_b1, out);
  }
  ++plan->launches;
}

void laplace_execute(fmmb_plan* plan, const double* d_charges, double* d_results) {
  Tree& T = plan->tree;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P;
  const int64_t n = T.n;
  cudaStream_t s = plan->stream, s2 = plan->overlap_p2p ? plan->stream2 : plan->stream;
  laplace_prepare_expansions(plan);
  plan->res_near.resize(n);
  plan->res_far.resize(n);
  cudaEvent_t* ev = plan->ev;
  plan->launches = 0;

  // ---- charges into the tree-ordered bodies
  NvtxRange r_exec("fmmb: LaplaceSpherical launches (gather, P2P, P2M, translations, L2P, scatter)");
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[0], s));
  if (plan->call_sharded) {
    // charges = this rank's slice in tree order, no permutation
    if (T.nranks > 1 && plan->peer_ready) {
      peer_exchange_charges(plan, d_charges, s);       // slices written straight into the peers (NVLink), no NCCL
    } else if (T.nranks > 1 && plan->comm) {
      allgather_charges(plan, d_charges, s);
      dim3 grid(64, T.nranks);
      place_charges<<<grid, 256, 0, s>>>(plan->chg_stage.p, plan->cuts_dev.p, T.nranks, plan->chg_chunk, T.body.p);
    } else {
      place_charges_local<<<nblk(n, 256), 256, 0, s>>>(d_charges, n, T.body.p);
    }
  } else {
    gather_charges<<<nblk(n, 256), 256, 0, s>>>(d_charges, T.perm.p, n, T.body.p);
  }
  ++plan->launches;
  FMMB_CUDA(cudaEventRecord(ev[1], s));

  // ---- near field on the second stream: needs only the charges.
  // One GPU: it starts right away and fills the gaps of the latency-bound upward chain.  Sharded with an owned
  // upward pass: it starts once the owned M2M sweep is enqueued (hook below), so that the short dependent kernels
  // before the multipole exchange are not queued behind its blocks and it overlaps the exchange instead.
  // One GPU, class-major engine (p2p_order 1): it starts when the M2L GEMM has finished.  The GEMM and the pair
  // kernel both live on the FP64 pipe, so side by side they only share it; one after the other, the rest of the
  // far-field chain (column reduction: HBM-bound; L2L, L2P: latency-bound) runs beside the near field, on the
  // stream with the higher priority, instead of after it.
  const bool p2m_owned = laplace_owned_upward(plan);
  const bool defer_p2p = p2m_owned && s2 != s && !plan->near_only && plan->p2p_defer;
  const bool behind_gemm = !defer_p2p && s2 != s && !plan->near_only && plan->p2p_order == 1 && T.nranks == 1 &&
                           plan->opts.evaluator != FMMB_EVAL_TREECODE && far_engine(plan) == 2 && P <= 8;
  bool near_launched = false;
  if (!defer_p2p && !behind_gemm) { launch_near_field(plan, s, s2); near_launched = true; }

  // ---- far field
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[12], s));
  if (plan->near_only) {
    // plans for preconditioners (FMMOptions::local_evaluation / block_diagonal): no far field at all
    FMMB_CUDA(cudaMemsetAsync(plan->res_far.p, 0, (size_t)n * sizeof(double4), s));
    if (!plan->capturing) { FMMB_CUDA(cudaEventRecord(ev[2], s)); FMMB_CUDA(cudaEventRecord(ev[3], s)); }
  } else {
    const int p2m_warps = pp <= 64 ? 4 : 1;          // shared tile: warps x 32 bodies x P^2 doubles
    const size_t p2m_sh = (size_t)p2m_warps * 32 * (pp | 1) * sizeof(double);
    // per call: function attributes belong to the current device
    FMMB_CUDA(cudaFuncSetAttribute(p2m_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    const int* p2m_list = p2m_owned ? T.own_leaves.p : T.leaves.p;
    const int p2m_n = p2m_owned ? T.n_own_leaves : T.nleaves;
    // narrow-tile kernel for the orders where the full tile limits the occupancy (P >= 7: measured 0.222 -> 0.195 ms
    // of upward phase at N = 1M, P = 8; at P = 5 the full tile is already small and the extra flushes cost 15 us)
    if (p2m_n && plan->p2m_kernel == 1 && P >= 7 && 2 * (P - 1) <= kP2MCols)
      p2m_cols_kernel<<<nblk(p2m_n, 4), 128, 0, s>>>(p2m_list, p2m_n, T.bbegin.p, T.bend.p, T.center.p, T.body.p, P,
                                                     plan->M.p);
    else if (p2m_n)
      p2m_kernel<<<nblk(p2m_n, p2m_warps), 32 * p2m_warps, p2m_sh, s>>>(p2m_list, p2m_n, T.bbegin.p, T.bend.p, T.center.p,
                                                                       T.body.p, P, plan->M.p);
    ++plan->launches;
    if (defer_p2p) {
      plan->hook_after_owned_m2m = [&]() {
        FMMB_CUDA(cudaEventRecord(ev[15], s));
        FMMB_CUDA(cudaStreamWaitEvent(s2, ev[15], 0));
        launch_near_field(plan, s, s2);
      };
    }
    if (behind_gemm) {
      plan->hook_after_m2l_gemm = [&]() {
        FMMB_CUDA(cudaEventRecord(ev[15], s));
        FMMB_CUDA(cudaStreamWaitEvent(s2, ev[15], 0));
        launch_near_field(plan, s, s2);
        near_launched = true;
      };
    }
    try { laplace_translations(plan, s); } catch (...) {
      plan->hook_after_owned_m2m = nullptr; plan->hook_after_m2l_gemm = nullptr; throw;
    }
    plan->hook_after_owned_m2m = nullptr;
    plan->hook_after_m2l_gemm = nullptr;
    if (behind_gemm && !near_launched) { launch_near_field(plan, s, s2); near_launched = true; }   // nothing was batched
    if (plan->opts.evaluator == FMMB_EVAL_TREECODE) {
      if (T.n_own_leaves)
        m2p_kernel<<<nblk(T.n_own_leaves, 4), 128, 4 * nc * sizeof(double2), s>>>(
            T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.parent.p, T.m2l_off.p, T.m2l_src.p, T.center.p,
            T.body.p, P, plan->M.p, plan->res_far.p);
    } else if (T.n_own_leaves && plan->l2p_kernel == 1 && T.own_leaves_body.n == (size_t)T.n_own_leaves) {
      const int nw = (T.n_own_leaves + kL2PLeaves - 1) / kL2PLeaves;
      l2p_packed_kernel<<<nblk(nw, 4), 128, (size_t)4 * kL2PLeaves * nc * sizeof(double2), s>>>(
          T.own_leaves_body.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.center.p, T.has_local.p, T.body.p, P, plan->L.p,
          plan->res_far.p);
    } else if (T.n_own_leaves) {
      l2p_kernel<<<nblk(T.n_own_leaves, 4), 128, 4 * nc * sizeof(double2), s>>>(
          T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.center.p, T.has_local.p, T.body.p, P, plan->L.p,
          plan->res_far.p);
    }
    ++plan->launches;
  }
  // peer exchange: this rank no longer reads its multipole array -> peers may push the next matvec's rows
  if (plan->peer_ready) peer_read_done(plan, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[4], s));

  // ---- results
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s, ev[7], 0));
  deliver_results(plan, d_results, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[5], s));
  FMMB_CUDA(cudaGetLastError());
  plan->timed = true;
}

// Plan-time: P2P work items (target leaf chunks of <= 32 bodies), in leaf order.
void build_p2p_items(fmmb_plan* plan) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  if (T.p2p_run_off.n == 0 && T.n_p2p > 0) {
    // merged source runs per target box (see p2p_run_kernel)
    const int64_t ne = T.n_p2p;
    const int nb = T.nboxes;
    DevBuf<unsigned long long> k0, k1;
    DevBuf<int> v0, v1, flag, pos;
    DevBuf<char> tmp;
    k0.resize(ne); k1.resize(ne); v0.resize(ne); v1.resize(ne); flag.resize(ne + 1); pos.resize(ne + 1);
    run_keys_kernel<<<nb, 64, 0, s>>>(T.p2p_off.p, T.p2p_src.p, T.bbegin.p, nb, k0.p, v0.p);
    int bits = 32;
    while ((1ll << (bits - 32)) < nb) ++bits;
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, v0.p, v1.p, ne, 0, bits, s));
    tmp.resize(bytes);
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, k0.p, k1.p, v0.p, v1.p, ne, 0, bits, s));
    run_flags_kernel<<<nblk(ne + 1, 256), 256, 0, s>>>(k1.p, v1.p, T.bbegin.p, T.bend.p, ne, flag.p);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, pos.p, ne + 1, s));
    tmp.resize(bytes);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, flag.p, pos.p, ne + 1, s));
    int nruns = 0;
    FMMB_CUDA(cudaMemcpyAsync(&nruns, pos.p + ne, sizeof(int), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    T.n_p2p_runs = nruns;
    T.p2p_runs.resize(nruns);
    T.p2p_run_off.resize(nb + 1);
    run_fill_kernel<<<nblk(ne, 256), 256, 0, s>>>(v1.p, T.bbegin.p, T.bend.p, flag.p, pos.p, ne, T.p2p_runs.p);
    run_offsets_kernel<<<nblk(nb + 1, 256), 256, 0, s>>>(T.p2p_off.p, pos.p, nb, T.p2p_run_off.p);
    // leaves with a foreign source closer than 1e-4 to one of their bodies (see p2p_pair2_kernel)
    T.p2p_close.resize(nb);
    T.p2p_close.zero(s);
    p2p_close_kernel<<<nblk(T.nleaves, 4), 128, 0, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p, T.p2p_off.p,
                                                       T.p2p_src.p, T.body.p, T.p2p_close.p);
    FMMB_CUDA(cudaGetLastError());
    FMMB_CUDA(cudaStreamSynchronize(s));
  }
  if (T.own_leaves_body.n != (size_t)T.n_own_leaves && T.n_own_leaves > 0) {
    // the rank's leaves in BODY order (consecutive entries own consecutive body ranges): l2p_packed_kernel
    std::vector<int> lv = T.own_leaves.to_host(s);
    std::vector<unsigned> hb = T.bbegin.to_host(s);
    std::sort(lv.begin(), lv.end(), [&](int a, int b) { return hb[a] < hb[b]; });
    T.own_leaves_body.from_host(lv.data(), lv.size(), s);
    FMMB_CUDA(cudaStreamSynchronize(s));
  }
  DevBuf<int> cnt;
  const int nl = T.n_own_leaves;
  cnt.resize(nl + 1);
  const int mode = plan->p2p_item_mode;
  // Target chunk per warp: 32 bodies, or less when the rank has too few leaves to fill the GPU several times over
  // (multi-GPU shards, small problems): a chunk of 16 or 8 targets splits its sources 4 or 8 ways across the
  // lanes at the same lane efficiency, and the shorter items even out the tail of the grid.
  int chunk = 32, sms = 148;
  if (const char* e = std::getenv("FMMB_P2P_CHUNK")) {   // development knob: first chunk size tried (32, 16 or 8)
    const int v = std::atoi(e);
    if (v == 32 || v == 16 || v == 8) chunk = v;
  }
  FMMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device));
  std::vector<int> h, off(nl + 1, 0);
  for (;;) {
    p2p_count_items<<<nblk(nl + 1, 256), 256, 0, s>>>(T.own_leaves.p, nl, T.bbegin.p, T.bend.p, mode, chunk, cnt.p);
    FMMB_CUDA(cudaGetLastError());
    h = cnt.to_host(s);
    for (int i = 0; i < nl; ++i) off[i + 1] = off[i] + h[i];
    if (plan->kind == FMMB_LAPLACE_SPHERICAL_BEM || plan->kind == FMMB_YUKAWA_CARTESIAN_BEM ||
        plan->kind == FMMB_STOKES_SPHERICAL_BEM || mode != 0 || chunk <= plan->p2p_min_chunk || off[nl] >= 6 * 20 * sms) break;
    chunk >>= 1;
  }
  plan->p2p_chunk = chunk;
  T.n_p2p_items = off[nl];
  DevBuf<int> doff;
  doff.from_host(off.data(), off.size(), s);
  const int ni = T.n_p2p_items;
  DevBuf<int4> unsorted;
  DevBuf<unsigned> work, work_sorted;
  unsorted.resize(ni); work.resize(ni); work_sorted.resize(ni);
  T.p2p_items.resize(ni);
  int4* fill_to = mode ? unsorted.p : T.p2p_items.p;
  if (nl) p2p_fill_items<<<nblk(nl, 256), 256, 0, s>>>(T.own_leaves.p, nl, T.bbegin.p, T.bend.p, doff.p, T.p2p_off.p,
                                                       T.p2p_src.p, mode, chunk, fill_to, work.p);
  FMMB_CUDA(cudaGetLastError());
  if (mode && ni) {
    // longest items first: the tail of the grid is made of the short pieces (stable: ties keep leaf order)
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, work.p, work_sorted.p, unsorted.p,
                                                        T.p2p_items.p, ni, 0, 32, s));
    DevBuf<char> tmp;
    tmp.resize(bytes);
    FMMB_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, bytes, work.p, work_sorted.p, unsorted.p,
                                                        T.p2p_items.p, ni, 0, 32, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
  }
  T.p2p_items_ext.resize(ni);
  if (ni && T.n_p2p > 0)
    p2p_items_ext_kernel<<<nblk(ni, 256), 256, 0, s>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p, T.p2p_run_off.p,
                                                      T.p2p_runs.p, T.p2p_close.p, T.p2p_items_ext.p);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));
}

void laplace_direct_raw(const double* d_spts, const double* d_q, int64_t ns, const double* d_tpts, int64_t nt,
                        double* d_out, cudaStream_t s) {
  direct_kernel<<<nblk(nt, 128), 128, 0, s>>>(d_spts, d_q, ns, d_tpts, nt, reinterpret_cast<double4*>(d_out));
  FMMB_CUDA(cudaGetLastError());
}

void measure_fp64_peak(double* dfma, double* dmma) {
  int dev = 0, sms = 0;
  FMMB_CUDA(cudaGetDevice(&dev));
  FMMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* out = nullptr;
  FMMB_CUDA(cudaMalloc(&out, 8));
  cudaEvent_t a, b;
  FMMB_CUDA(cudaEventCreate(&a));
  FMMB_CUDA(cudaEventCreate(&b));
  const int iters = 2048, blocks = sms * 8, threads = 256;
  double best[2] = {0, 0};
  for (int rep = 0; rep < 4; ++rep)
    for (int which = 0; which < 2; ++which) {
      FMMB_CUDA(cudaEventRecord(a));
      if (which == 0) dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      else dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      FMMB_CUDA(cudaEventRecord(b));
      FMMB_CUDA(cudaEventSynchronize(b));
      float ms = 0;
      FMMB_CUDA(cudaEventElapsedTime(&ms, a, b));
      // DFMA: 64 fma per thread per iteration; DMMA: 32 mma x 256 fma per warp per iteration
      double flops = which == 0 ? 2.0 * 64.0 * iters * (double)blocks * threads
                                : 2.0 * 256.0 * 32.0 * iters * (double)blocks * (threads / 32);
      double tf = flops / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best[which]) best[which] = tf;
    }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  *dfma = best[0];
  *dmma = best[1];
}

}  // namespace fmmb
