// csrc/laplace_ops.cuh -- device building blocks of the LaplaceSpherical operators shared by the
// per-box kernels (laplace.cu) and the translation-matrix builder (m2l_classes.cu).
// Reference: kernel/LaplaceSpherical.hpp (cart2sph :528-541, evalMultipole :455-488,
// evalLocal :491-524, M2M :245-285, M2L :296-329, L2L :378-411).
#pragma once
#include "common.cuh"
#include "laplace_tables.cuh"

namespace fmmb {
namespace ops {

constexpr double kEps = 1e-12;

struct Sph { double r, x, y, cp, sp; };

// cart2sph (LaplaceSpherical.hpp:528-541) without the inverse trig round trip
__device__ __forceinline__ Sph to_sph(double dx, double dy, double dz) {
  Sph s;
  s.r = sqrt(dx * dx + dy * dy + dz * dz) + kEps;
  s.x = __ddiv_rn(dz, s.r);
  s.y = sqrt((1.0 - s.x) * (1.0 + s.x));
  double ax = fabs(dx), ay = fabs(dy);
  if (ax + ay < kEps) { s.cp = 1.0; s.sp = 0.0; }
  else if (ax < kEps) { s.cp = 0.0; s.sp = dy > 0 ? 1.0 : -1.0; }
  else { double h = sqrt(dx * dx + dy * dy); s.cp = dx / h; s.sp = dy / h; }
  return s;
}

// All rho^n Y_n^m, 0 <= m <= n < P, visited m-major exactly like evalMultipole (:455-488).
// f(n, m, Yre, Yim, Ytre, Ytim); sign = +1 for e^{+i m phi}, -1 for e^{-i m phi}.
// SINGULAR: rho^{-n-1} Y_n^m instead (evalLocal :491-524, used by M2P for n < P).
struct NoColumnHook { __device__ __forceinline__ void operator()(int) const {} };

// g(m) is called after the last term of column m (all n for that m): P2M flushes its transposition tile there.
template <bool THETA, bool SINGULAR = false, typename F, typename G = NoColumnHook>
__device__ __forceinline__ void regular_harmonics(int P, const Sph& s, double sign, F&& f, G&& g = G()) {
  const double step = SINGULAR ? 1.0 / s.r : s.r;
  double fact = 1, pn = 1, rhom = SINGULAR ? step : 1.0;
  double er = 1, ei = 0;
  const double cp = s.cp, sp = sign * s.sp;
  // divisions of the recurrences become multiplications by tabulated 1/k and one 1/sin(alpha)
  const double inv_y = THETA ? 1.0 / s.y : 0.0;
  for (int m = 0; m < P; ++m) {
    double p = pn;
    int npn = m * m + 2 * m;
    double a = rhom * p * c_pref[npn];
    double p1 = p;
    p = s.x * (2 * m + 1) * p1;
    double at = 0;
    if (THETA) at = rhom * (p - (m + 1) * s.x * p1) * inv_y * c_pref[npn];
    f(m, m, a * er, a * ei, at * er, at * ei);
    rhom *= step;
    double rhon = rhom;
    for (int n = m + 1; n < P; ++n) {
      int npm = n * n + n + m;
      a = rhon * p * c_pref[npm];
      double p2 = p1;
      p1 = p;
      p = (s.x * (2 * n + 1) * p1 - (n + m) * p2) * c_rcp[n - m + 1];
      if (THETA) at = rhon * ((n - m + 1) * p - (n + 1) * s.x * p1) * inv_y * c_pref[npm];
      f(n, m, a * er, a * ei, at * er, at * ei);
      rhon *= step;
    }
    g(m);
    pn = -pn * fact * s.y;
    fact += 2;
    double t = er * cp - ei * sp;
    ei = er * sp + ei * cp;
    er = t;
  }
}

// One column (fixed m >= 0) of the harmonics table, written for +m and -m (conjugate).
// SINGULAR: rho^{-n-1} Y_n^m for n < top (evalLocal :491-524); else rho^n Y_n^m.
template <bool SINGULAR>
__device__ __forceinline__ void harmonics_column(int m, int top, const Sph& s, double sign, double2* Y) {
  double pn = 1, fact = 1, er = 1, ei = 0;
  double rhom = SINGULAR ? 1.0 / s.r : 1.0;
  const double cp = s.cp, sp = sign * s.sp;
  for (int k = 0; k < m; ++k) {
    pn = -pn * fact * s.y;
    fact += 2;
    double t = er * cp - ei * sp;
    ei = er * sp + ei * cp;
    er = t;
    if (SINGULAR) rhom /= s.r; else rhom *= s.r;
  }
  double p = pn;
  int npn = m * m + 2 * m, nmn = m * m;
  double a = rhom * p * c_pref[npn];
  Y[npn] = make_double2(a * er, a * ei);
  Y[nmn] = make_double2(a * er, -a * ei);
  double p1 = p;
  p = s.x * (2 * m + 1) * p1;
  if (SINGULAR) rhom /= s.r; else rhom *= s.r;
  double rhon = rhom;
  for (int n = m + 1; n < top; ++n) {
    int npm = n * n + n + m, nmm = n * n + n - m;
    a = rhon * p * c_pref[npm];
    Y[npm] = make_double2(a * er, a * ei);
    Y[nmm] = make_double2(a * er, -a * ei);
    double p2 = p1;
    p1 = p;
    p = (s.x * (2 * n + 1) * p1 - (n + m) * p2) / (n - m + 1);
    if (SINGULAR) rhon /= s.r; else rhon *= s.r;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ double oddeven(int n) { return (n & 1) ? -1.0 : 1.0; }


// i^(|k-m|-|k|-|m|) (-1)^j Anm[nm] Anm[jk] / Anm[(j+n)^2+(j+n)+m-k]: the exponent is always even,
// so the M2L coefficient Cnm (LaplaceSpherical.hpp:106-116) is real.
__device__ __forceinline__ double cnm_real(int j, int k, int n, int m) {
  int e = abs(k - m) - abs(k) - abs(m);
  double sgn = ((e / 2) & 1) ? -1.0 : 1.0;
  return sgn * oddeven(j) * c_anm[n * n + n + m] * c_anm[j * j + j + k] /
         c_anm[(j + n) * (j + n) + j + n + m - k];
}

// One output coefficient (j,k), k >= 0, of M2M (LaplaceSpherical.hpp:255-281).
// Ms: packed source multipole; Y: rho^n Y_n^m(alpha,-beta) of c_parent - c_child, full (n^2+n+m).
__device__ __forceinline__ double2 m2m_entry(const double2* Ms, const double2* Y, int j, int k) {
  int jk = j * j + j + k;
  double inv_ajk = 1.0 / c_anm[jk];
  double ar = 0, ai = 0;
  for (int n = 0; n <= j; ++n) {
    int mtop = min(k - 1, n);
    for (int m = -n; m <= mtop; ++m) {
      if (j - n >= k - m) {
        int jnkm = (j - n) * (j - n) + j - n + k - m;
        int jnkms = (j - n) * (j - n + 1) / 2 + k - m;
        int nm = n * n + n + m;
        // i^(m-|m|) = (-1)^m for m < 0, 1 otherwise
        double f = ((m < 0 && (m & 1)) ? -1.0 : 1.0) * oddeven(n) * c_anm[nm] * c_anm[jnkm] * inv_ajk;
        double2 a = Ms[jnkms], y = Y[nm];
        ar += f * (a.x * y.x - a.y * y.y);
        ai += f * (a.x * y.y + a.y * y.x);
      }
    }
    for (int m = k; m <= n; ++m) {
      if (j - n >= m - k) {
        int jnkm = (j - n) * (j - n) + j - n + k - m;
        int jnkms = (j - n) * (j - n + 1) / 2 - k + m;
        int nm = n * n + n + m;
        double f = oddeven(k + n + m) * c_anm[nm] * c_anm[jnkm] * inv_ajk;
        double2 a = Ms[jnkms], y = Y[nm];   // conj(a) * y
        ar += f * (a.x * y.x + a.y * y.y);
        ai += f * (a.x * y.y - a.y * y.x);
      }
    }
  }
  return make_double2(ar, ai);
}

// One output coefficient (j,k), k >= 0, of L2L (LaplaceSpherical.hpp:385-409).
// Ls: packed parent local; Y: rho^n Y_n^m(alpha,+beta) of c_child - c_parent.
__device__ __forceinline__ double2 l2l_entry(const double2* Ls, const double2* Y, int j, int k, int P) {
  int jk = j * j + j + k;
  double ajk = c_anm[jk];
  double ar = 0, ai = 0;
  for (int n = j; n < P; ++n) {
    for (int m = j + k - n; m < 0; ++m) {
      int jnkm = (n - j) * (n - j) + n - j + m - k;
      int nm = n * n + n - m, nms = n * (n + 1) / 2 - m;
      double f = oddeven(k) * c_anm[jnkm] * ajk / c_anm[nm];
      double2 a = Ls[nms], y = Y[jnkm];       // conj(a) * y
      ar += f * (a.x * y.x + a.y * y.y);
      ai += f * (a.x * y.y - a.y * y.x);
    }
    for (int m = 0; m <= n; ++m) {
      if (n - j >= abs(m - k)) {
        int jnkm = (n - j) * (n - j) + n - j + m - k;
        int nm = n * n + n + m, nms = n * (n + 1) / 2 + m;
        // i^(m-k-|m-k|) = (-1)^(m-k) for m < k, 1 otherwise
        double f = ((m < k && ((k - m) & 1)) ? -1.0 : 1.0) * c_anm[jnkm] * ajk / c_anm[nm];
        double2 a = Ls[nms], y = Y[jnkm];
        ar += f * (a.x * y.x - a.y * y.y);
        ai += f * (a.x * y.y + a.y * y.x);
      }
    }
  }
  return make_double2(ar, ai);
}

// Expansions live in global memory in a REAL layout: P^2 doubles per box, Re X_n^m at n^2+n+m
// (m >= 0) and Im X_n^m at n^2+n-m (m > 0); Im X_n^0 is identically zero and not stored.  The
// per-box stride is padded to an even count so every box starts on a 16-byte boundary.
__host__ __device__ __forceinline__ int xstride(int P) { return (P * P + 1) & ~1; }
__device__ __forceinline__ double2 load_coef(const double* X, int n, int m) {
  return make_double2(X[n * n + n + m], m > 0 ? X[n * n + n - m] : 0.0);
}
__device__ __forceinline__ void store_coef(double* X, int n, int m, double2 v) {
  X[n * n + n + m] = v.x;
  if (m > 0) X[n * n + n - m] = v.y;
}
__device__ __forceinline__ void add_coef(double* X, int n, int m, double2 v) {
  X[n * n + n + m] += v.x;
  if (m > 0) X[n * n + n - m] += v.y;
}

// packed index n(n+1)/2+m  ->  (n, m)
__device__ __forceinline__ void unpack_nm(int nms, int& n, int& m) {
  n = 0;
  while ((n + 1) * (n + 2) / 2 <= nms) ++n;
  m = nms - n * (n + 1) / 2;
}

}  // namespace ops
}  // namespace fmmb
