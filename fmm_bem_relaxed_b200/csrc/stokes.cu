// csrc/stokes.cu -- StokesSpherical on the GPU: Stokeslet and stresslet far field through four Laplace
// expansion sets per box, near field by a warp-tiled pair kernel.
//
// Replaces (reference kernel/StokesSpherical.hpp):
//   :47-59    init_multipole / init_local: 4 Laplace expansions per box     -> StokesData::M4 / L4
//   :62-78    Stokeslet operator() (Mat3 per pair) applied by Direct.hpp    -> stokes_p2p_kernel<false>
//   :85-116   stresslet P2P (vector form)                                   -> stokes_p2p_kernel<true>
//   :122-147  Stokeslet P2M: sets f0 Y, f1 Y, f2 Y, (f.x) Y                 -> stokes_p2m_kernel<false>
//   :151-188  stresslet P2M: sets (w_s . grad)(rho^n Y), w_s = g_s n + n_s g (s < 3),
//             w_3 = (x.g) n + (n.x) g                                       -> stokes_p2m_kernel<true>
//   :190-196, :293-307  M2M / M2L / L2L on each of the 4 sets               -> laplace_translations() per set
//   :318-401  L2P: u_s += c (phi_s - x_s d_k phi_k ... ) i.e. per set s the potential (s < 3) and the Cartesian
//             gradient times -x_s (s < 3) or 1 (s = 3); c = 1 (Stokeslet) or 1/6 (stresslet) -> stokes_l2p_kernel
//   :207-291  M2P (treecode evaluator): the same on the singular harmonics                -> stokes_m2p_kernel
// x is the ABSOLUTE body position (not relative to the box centre), exactly as in the reference.
//
// The stresslet path corresponds to the reference compiled with -DSTRESSLET and the two compile fixes listed in
// SURVEY.md section 8(c) (rdotn / rdotg complex); the Stokeslet path to the unmodified reference.
#include "common.cuh"
#include "laplace_ops.cuh"
#include <algorithm>

namespace fmmb {

struct StokesData {
  bool stresslet = false;
  int cd = 3;                    // doubles per charge: 3 (f) or 6 (g, n)
  int rec = 6;                   // doubles per source record: position + charge
  DevBuf<double> src;            // tree order, [n][rec]: x, y, z, charge...
  DevBuf<double> M4[4], L4[4];   // the four expansion sets, real layout, box-major
  DevBuf<double> res_near, res_far;  // tree order, 3 per body
  int p_alloc = 0;
};

void stokes_free(StokesData* d) { delete d; }

namespace {

using namespace ops;

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

__global__ void stokes_positions(const double4* __restrict__ body, int64_t n, int rec, double* __restrict__ src) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = body[i];
  double* o = src + (size_t)i * rec;
  o[0] = p.x; o[1] = p.y; o[2] = p.z;
}

// charges (original order, cd per body) into the tree-ordered source records
__global__ void stokes_gather(const double* __restrict__ q, const unsigned* __restrict__ perm, int64_t n, int cd,
                              int rec, double* __restrict__ src) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n * cd) return;
  const int64_t i = t / cd;
  const int c = (int)(t - i * cd);
  src[(size_t)i * rec + 3 + c] = q[(size_t)perm[i] * cd + c];
}

// ---- P2M: warp per (leaf, set); lane = body writes its row, then lane = coefficient sums the column ----
template <bool STRESSLET>
__global__ void __launch_bounds__(128)
stokes_p2m_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const double4* __restrict__ center,
                  const double* __restrict__ src, int P, double* __restrict__ M0, double* __restrict__ M1,
                  double* __restrict__ M2, double* __restrict__ M3) {
  extern __shared__ double stk_sh[];
  constexpr int REC = STRESSLET ? 9 : 6;
  const int pp = P * P, ld = pp | 1;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  const int set = blockIdx.y;
  if (w >= nleaves) return;
  double* tile = stk_sh + (size_t)wl * 32 * ld;
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
      const double* r = src + (size_t)i * REC;
      const double x = r[0], y = r[1], z = r[2];
      const Sph s = to_sph(x - c.x, y - c.y, z - c.z);
      double* row = tile + lane * ld;
      if (!STRESSLET) {
        const double f0 = r[3], f1 = r[4], f2 = r[5];
        const double mult = set == 0 ? f0 : (set == 1 ? f1 : (set == 2 ? f2 : f0 * x + f1 * y + f2 * z));
        regular_harmonics<false>(P, s, -1.0, [&](int n, int m, double yr, double yi, double, double) {
          row[n * n + n + m] = mult * yr;
          if (m > 0) row[n * n + n - m] = mult * yi;
        });
      } else {
        const double g0 = r[3], g1 = r[4], g2 = r[5], n0 = r[6], n1 = r[7], n2 = r[8];
        // direction of the derivative for this set (see the header comment)
        double a, bq;
        if (set == 0) { a = g0; bq = n0; }
        else if (set == 1) { a = g1; bq = n1; }
        else if (set == 2) { a = g2; bq = n2; }
        else { a = x * g0 + y * g1 + z * g2; bq = n0 * x + n1 * y + n2 * z; }
        const double w0 = a * n0 + bq * g0, w1 = a * n1 + bq * g1, w2 = a * n2 + bq * g2;
        // spherical basis vectors over the metric: grad = e_r d/drho + e_a/rho d/dalpha + e_b/(rho sin) d/dbeta
        const double ir = 1.0 / s.r, iry = ir / s.y;
        const double wa = w0 * (s.y * s.cp) + w1 * (s.y * s.sp) + w2 * s.x;
        const double wb = (w0 * (s.x * s.cp) + w1 * (s.x * s.sp) - w2 * s.y) * ir;
        const double wc = (-w0 * s.sp + w1 * s.cp) * iry;
        regular_harmonics<true>(P, s, -1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
          // brh = n/rho Y, bal = Ytheta, bbe = -i m Y = (m yi, -m yr)
          const double fr = n * ir;
          row[n * n + n + m] = wa * fr * yr + wb * tr + wc * (m * yi);
          if (m > 0) row[n * n + n - m] = wa * fr * yi + wb * ti - wc * (m * yr);
        });
      }
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
  double* Mb = (set == 0 ? M0 : (set == 1 ? M1 : (set == 2 ? M2 : M3))) + (size_t)b * xstride(P);
#pragma unroll
  for (int i2 = 0; i2 < (FMMB_MAX_P * FMMB_MAX_P + 31) / 32; ++i2) {
    const int col = lane + 32 * i2;
    if (col < pp) Mb[col] = acc[i2];
  }
}

// ---- L2P: warp per leaf, lane per body, all four sets in one pass over the harmonics -----------------
__global__ void __launch_bounds__(128)
stokes_l2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const double4* __restrict__ center,
                  const unsigned char* __restrict__ has_local, const double4* __restrict__ body, int P,
                  const double* __restrict__ L0, const double* __restrict__ L1, const double* __restrict__ L2,
                  const double* __restrict__ L3, double scale, double* __restrict__ res) {
  extern __shared__ double2 stk_ls[];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int b = leaves[w];
  const unsigned b0 = bb[b], b1 = be[b];
  if (!has_local[b]) {
    for (unsigned i = 3 * b0 + lane; i < 3 * b1; i += 32) res[i] = 0.0;
    return;
  }
  double2* Ls = stk_ls + (size_t)wl * 4 * nc;
  for (int i = lane; i < 4 * nc; i += 32) {
    const int set = i / nc, e = i - set * nc;
    int n, m;
    unpack_nm(e, n, m);
    const double* L = set == 0 ? L0 : (set == 1 ? L1 : (set == 2 ? L2 : L3));
    Ls[i] = load_coef(L + (size_t)b * xstride(P), n, m);
  }
  __syncwarp();
  const double4 c = center[b];
  for (unsigned i = b0 + lane; i < b1; i += 32) {
    const double4 p = body[i];
    const Sph s = to_sph(p.x - c.x, p.y - c.y, p.z - c.z);
    const double inv_r = 1.0 / s.r;
    double pot[4] = {0, 0, 0, 0}, ga[4] = {0, 0, 0, 0}, gb[4] = {0, 0, 0, 0}, gc[4] = {0, 0, 0, 0};
    regular_harmonics<true>(P, s, 1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
      const double w2 = m == 0 ? 1.0 : 2.0;
      const int e = n * (n + 1) / 2 + m;
#pragma unroll
      for (int set = 0; set < 4; ++set) {
        const double2 l = Ls[set * nc + e];
        const double re = w2 * (l.x * yr - l.y * yi);     // Re(L Y)
        pot[set] += re;
        ga[set] += re * inv_r * n;
        gb[set] += w2 * (l.x * tr - l.y * ti);            // Re(L Ytheta)
        gc[set] -= w2 * (l.x * yi + l.y * yr) * m;        // Re(L Y i) m
      }
    });
    const double inv_ry = inv_r / s.y;
    const double xs_[3] = {p.x, p.y, p.z};
    double u[3] = {0, 0, 0};
#pragma unroll
    for (int set = 0; set < 4; ++set) {
      const double cx = s.y * s.cp * ga[set] + s.x * s.cp * inv_r * gb[set] - s.sp * inv_ry * gc[set];
      const double cy = s.y * s.sp * ga[set] + s.x * s.sp * inv_r * gb[set] + s.cp * inv_ry * gc[set];
      const double cz = s.x * ga[set] - s.y * inv_r * gb[set];
      const double f = set < 3 ? -xs_[set < 3 ? set : 0] : 1.0;
      u[0] += f * cx; u[1] += f * cy; u[2] += f * cz;
    }
    res[3 * (size_t)i + 0] = scale * (pot[0] + u[0]);
    res[3 * (size_t)i + 1] = scale * (pot[1] + u[1]);
    res[3 * (size_t)i + 2] = scale * (pot[2] + u[2]);
  }
}

// ---- M2P (treecode, FMMOptions::TREECODE; StokesSpherical.hpp:207-291): warp per leaf, lane per body; every source
// box accepted for the leaf or one of its ancestors is evaluated at the bodies -- the L2P arithmetic on the singular
// harmonics (radial factor -(n+1)/r), converted to Cartesian per source box.  Fixed order, no atomics.
__global__ void __launch_bounds__(128)
stokes_m2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const unsigned* __restrict__ parent, const int* __restrict__ off,
                  const int* __restrict__ srcbox, const double4* __restrict__ center, const double4* __restrict__ body,
                  int P, const double* __restrict__ M0, const double* __restrict__ M1, const double* __restrict__ M2,
                  const double* __restrict__ M3, double scale, double* __restrict__ res) {
  extern __shared__ double2 stk_ms[];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double2* Ms = stk_ms + (size_t)wl * 4 * nc;
  const int leaf = leaves[w];
  const unsigned b0 = bb[leaf], b1 = be[leaf];
  for (unsigned base = b0; base < b1; base += 32) {
    const unsigned i = base + lane;
    const bool act = i < b1;
    const double4 p = act ? body[i] : make_double4(0, 0, 0, 0);
    const double xs_[3] = {p.x, p.y, p.z};
    double u[3] = {0, 0, 0};
    for (int a = leaf;; a = (int)parent[a]) {
      for (int e = off[a]; e < off[a + 1]; ++e) {
        const int sb = srcbox[e];
        __syncwarp();
        for (int k = lane; k < 4 * nc; k += 32) {
          const int set = k / nc, c = k - set * nc;
          int n, m;
          unpack_nm(c, n, m);
          const double* M = set == 0 ? M0 : (set == 1 ? M1 : (set == 2 ? M2 : M3));
          Ms[k] = load_coef(M + (size_t)sb * xstride(P), n, m);
        }
        __syncwarp();
        if (act) {
          const double4 c = center[sb];
          const Sph s = to_sph(p.x - c.x, p.y - c.y, p.z - c.z);
          const double inv_r = 1.0 / s.r;
          double pot[4] = {0, 0, 0, 0}, ga[4] = {0, 0, 0, 0}, gb[4] = {0, 0, 0, 0}, gc[4] = {0, 0, 0, 0};
          regular_harmonics<true, true>(P, s, 1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
            const double w2 = m == 0 ? 1.0 : 2.0;
            const int q = n * (n + 1) / 2 + m;
#pragma unroll
            for (int set = 0; set < 4; ++set) {
              const double2 l = Ms[set * nc + q];
              const double re = w2 * (l.x * yr - l.y * yi);     // Re(M Y)
              pot[set] += re;
              ga[set] -= re * inv_r * (n + 1);
              gb[set] += w2 * (l.x * tr - l.y * ti);            // Re(M Ytheta)
              gc[set] -= w2 * (l.x * yi + l.y * yr) * m;        // Re(M Y i) m
            }
          });
          const double inv_ry = inv_r / s.y;
#pragma unroll
          for (int set = 0; set < 4; ++set) {
            const double cx = s.y * s.cp * ga[set] + s.x * s.cp * inv_r * gb[set] - s.sp * inv_ry * gc[set];
            const double cy = s.y * s.sp * ga[set] + s.x * s.sp * inv_r * gb[set] + s.cp * inv_ry * gc[set];
            const double cz = s.x * ga[set] - s.y * inv_r * gb[set];
            const double f = set < 3 ? -xs_[set < 3 ? set : 0] : 1.0;
            u[0] += f * cx; u[1] += f * cy; u[2] += f * cz;
          }
          u[0] += pot[0]; u[1] += pot[1]; u[2] += pot[2];
        }
      }
      if (a == 0) break;
    }
    if (act) {
      res[3 * (size_t)i + 0] = scale * u[0];
      res[3 * (size_t)i + 1] = scale * u[1];
      res[3 * (size_t)i + 2] = scale * u[2];
    }
  }
}

// ---- near field: one warp per (target leaf, <= 32 targets); sources through a warp-private tile ------
constexpr int kStkWarps = 4;

__device__ __forceinline__ double rsqrt_masked(double r2) {
  // 1/sqrt(r2) by MUFU.RSQ64H + one cubic step (same as the Laplace pair kernel); 0 for r2 < 1e-8
  // (StokesSpherical.hpp:69,101)
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
  const double e = fma(-(r2 * y0), y0, 1.0);
  double inv = fma(y0 * e, fma(0.375, e, 0.5), y0);
  if (r2 < 1e-8) inv = 0.0;
  return inv;
}

template <bool STRESSLET>
__device__ __forceinline__ void stokes_pair(double tx, double ty, double tz, const double* __restrict__ s,
                                            double& u0, double& u1, double& u2) {
  if (STRESSLET) {
    const double dx = tx - s[0], dy = ty - s[1], dz = tz - s[2];          // target - source (:96)
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double inv = rsqrt_masked(r2);
    const double inv2 = inv * inv;
    const double dn = dx * s[6] + dy * s[7] + dz * s[8];
    const double dg = dx * s[3] + dy * s[4] + dz * s[5];
    const double H = (inv2 * inv2) * (inv * dn) * dg;                     // (dx.n)(dx.g) / r^5
    u0 = fma(H, dx, u0); u1 = fma(H, dy, u1); u2 = fma(H, dz, u2);
  } else {
    const double dx = s[0] - tx, dy = s[1] - ty, dz = s[2] - tz;          // source - target (:66)
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double inv = rsqrt_masked(r2);
    const double inv3 = inv * inv * inv;
    const double df = (dx * s[3] + dy * s[4] + dz * s[5]) * inv3;
    u0 += fma(inv, s[3], df * dx); u1 += fma(inv, s[4], df * dy); u2 += fma(inv, s[5], df * dz);
  }
}

template <bool STRESSLET>
__global__ void __launch_bounds__(32 * kStkWarps)
stokes_p2p_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                  const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ srcbox,
                  const double* __restrict__ src, double* __restrict__ res) {
  constexpr int REC = STRESSLET ? 9 : 6;
  __shared__ double tiles[kStkWarps][32 * REC];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kStkWarps + wl;
  if (item >= nitems) return;
  double* tile = tiles[wl];
  const int4 it = items[item];
  const int r = it.z;
  const int S = 32 / r;                    // source splits (1 when r > 16)
  const int ti = lane % r, sp = lane / r;
  const bool act = sp < S;
  const double* tp = src + (size_t)(it.y + ti) * REC;
  const double tx = tp[0], ty = tp[1], tz = tp[2];
  double u0 = 0, u1 = 0, u2 = 0;
  const int s0 = off[it.x], s1 = off[it.x + 1];
  for (int e = s0; e < s1; ++e) {
    const int sb = srcbox[e];
    const unsigned c0 = bb[sb], c1 = be[sb];
    for (unsigned base = c0; base < c1; base += 32) {
      const int cnt = (int)min(32u, c1 - base);
      __syncwarp();
      const double* g = src + (size_t)base * REC;
      for (int k = lane; k < cnt * REC; k += 32) tile[k] = g[k];       // contiguous records: coalesced
      __syncwarp();
      if (act) {
        if (S == 1) {
#pragma unroll 4
          for (int k = 0; k < cnt; ++k) stokes_pair<STRESSLET>(tx, ty, tz, tile + k * REC, u0, u1, u2);
        } else {
          for (int k = sp; k < cnt; k += S) stokes_pair<STRESSLET>(tx, ty, tz, tile + k * REC, u0, u1, u2);
        }
      }
    }
  }
  if (S > 1) {
    for (int q = 1; q < S; ++q) {
      const int from = (lane + q * r) & 31;
      const double a = __shfl_sync(0xffffffffu, u0, from), b2 = __shfl_sync(0xffffffffu, u1, from),
                   c2 = __shfl_sync(0xffffffffu, u2, from);
      if (lane < r) { u0 += a; u1 += b2; u2 += c2; }
    }
  }
  if (lane < r) {
    double* o = res + 3 * (size_t)(it.y + lane);
    o[0] = u0; o[1] = u1; o[2] = u2;
  }
}


// brute force over all sources (Direct::matvec with the kernel's own pair rule)
template <bool STRESSLET>
__global__ void __launch_bounds__(128)
stokes_direct_kernel(const double* __restrict__ spts, const double* __restrict__ q, int64_t ns,
                     const double* __restrict__ tpts, int64_t nt, double* __restrict__ out) {
  constexpr int REC = STRESSLET ? 9 : 6, CD = REC - 3;
  __shared__ double tile[128 * REC];
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool act = i < nt;
  const double tx = act ? tpts[3 * i] : 0, ty = act ? tpts[3 * i + 1] : 0, tz = act ? tpts[3 * i + 2] : 0;
  double u0 = 0, u1 = 0, u2 = 0;
  for (int64_t base = 0; base < ns; base += 128) {
    const int cnt = (int)min((int64_t)128, ns - base);
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      const int64_t j = base + threadIdx.x;
      double* o = tile + threadIdx.x * REC;
      o[0] = spts[3 * j]; o[1] = spts[3 * j + 1]; o[2] = spts[3 * j + 2];
#pragma unroll
      for (int c = 0; c < CD; ++c) o[3 + c] = q[(size_t)j * CD + c];
    }
    __syncthreads();
    if (act)
      for (int k = 0; k < cnt; ++k) stokes_pair<STRESSLET>(tx, ty, tz, tile + k * REC, u0, u1, u2);
  }
  if (act) { out[3 * i] = u0; out[3 * i + 1] = u1; out[3 * i + 2] = u2; }
}

// swaps the storage of two device buffers (used to run the Laplace translations on one set at a time)
void swap_buf(DevBuf<double>& a, DevBuf<double>& b) {
  std::swap(a.p, b.p); std::swap(a.cap, b.cap); std::swap(a.n, b.n);
}
struct SetGuard {            // plan->M / plan->L temporarily ARE set k of the Stokes data
  fmmb_plan* plan; StokesData* d; int k;
  SetGuard(fmmb_plan* pl, StokesData* dd, int kk) : plan(pl), d(dd), k(kk) { swap_buf(plan->M, d->M4[k]); swap_buf(plan->L, d->L4[k]); }
  ~SetGuard() { swap_buf(plan->M, d->M4[k]); swap_buf(plan->L, d->L4[k]); }
};

}  // namespace

void stokes_setup(fmmb_plan* plan, bool stresslet) {
  Tree& T = plan->tree;
  StokesData* d = new StokesData();
  plan->stokes = d;
  d->stresslet = stresslet;
  d->cd = stresslet ? 6 : 3;
  d->rec = 3 + d->cd;
  upload_laplace_tables();   // this translation unit's copy of the factorial tables
  d->src.resize((size_t)T.n * d->rec);
  d->src.zero(plan->stream);
  stokes_positions<<<nblk(T.n, 256), 256, 0, plan->stream>>>(T.body.p, T.n, d->rec, d->src.p);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(plan->stream));
}

// The four expansion sets at the plan's current order.  A change of order moves every row (and the all-zero row
// nboxes that trans_blocked.cu reads for absent pairs), so the arrays are cleared then; run_matvec calls this before
// it replays a cached graph as well -- a graph holds the launches of a matvec, not this clearing.
void stokes_prepare_expansions(fmmb_plan* plan) {
  StokesData* d = plan->stokes;
  const int P = plan->p, xs = xstride(P);
  for (int k = 0; k < 4; ++k) {
    d->M4[k].resize((size_t)(plan->tree.nboxes + 1) * xs);
    d->L4[k].resize((size_t)(plan->tree.nboxes + 1) * xs);
    if (d->p_alloc != P) { d->M4[k].zero(plan->stream); d->L4[k].zero(plan->stream); }
  }
  d->p_alloc = P;
}

void stokes_execute(fmmb_plan* plan, const double* d_charges, double* d_results) {
  Tree& T = plan->tree;
  StokesData* d = plan->stokes;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P;
  const int xs = xstride(P);
  const int64_t n = T.n;
  cudaStream_t s = plan->stream, s2 = plan->overlap_p2p ? plan->stream2 : plan->stream;
  cudaEvent_t* ev = plan->ev;
  stokes_prepare_expansions(plan);
  d->res_near.resize(3 * (size_t)n);
  d->res_far.resize(3 * (size_t)n);
  plan->launches = 0;

  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[0], s));
  stokes_gather<<<nblk(n * d->cd, 256), 256, 0, s>>>(exec_charges(plan, d_charges), exec_perm(plan), n, d->cd, d->rec, d->src.p);
  ++plan->launches;
  FMMB_CUDA(cudaEventRecord(ev[1], s));

  // near field on the second stream
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s2, ev[1], 0));
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[6], s2));
  if (T.n_p2p_items) {
    if (d->stresslet)
      stokes_p2p_kernel<true><<<nblk(T.n_p2p_items, kStkWarps), 32 * kStkWarps, 0, s2>>>(
          T.p2p_items.p, T.n_p2p_items, T.bbegin.p, T.bend.p, T.p2p_off.p, T.p2p_src.p, d->src.p, d->res_near.p);
    else
      stokes_p2p_kernel<false><<<nblk(T.n_p2p_items, kStkWarps), 32 * kStkWarps, 0, s2>>>(
          T.p2p_items.p, T.n_p2p_items, T.bbegin.p, T.bend.p, T.p2p_off.p, T.p2p_src.p, d->src.p, d->res_near.p);
    ++plan->launches;
  }
  FMMB_CUDA(cudaEventRecord(ev[7], s2));

  // upward: all four sets in one launch (grid.y = set)
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[12], s));
  const int warps = pp <= 64 ? 4 : 1;
  const size_t sh = (size_t)warps * 32 * (pp | 1) * sizeof(double);
  // per call: function attributes belong to the current device
  FMMB_CUDA(cudaFuncSetAttribute(stokes_p2m_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  FMMB_CUDA(cudaFuncSetAttribute(stokes_p2m_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  // multi-GPU with a communicator: the upward pass is owned (own leaves, multipoles exchanged per set by
  // laplace_translations); otherwise it is replicated
  const bool p2m_owned = laplace_owned_upward(plan);
  const int* p2m_list = p2m_owned ? T.own_leaves.p : T.leaves.p;
  const int p2m_n = p2m_owned ? T.n_own_leaves : T.nleaves;
  const dim3 pg(nblk(p2m_n, warps), 4);
  if (p2m_n && d->stresslet)
    stokes_p2m_kernel<true><<<pg, 32 * warps, sh, s>>>(p2m_list, p2m_n, T.bbegin.p, T.bend.p, T.center.p,
                                                      d->src.p, P, d->M4[0].p, d->M4[1].p, d->M4[2].p, d->M4[3].p);
  else if (p2m_n)
    stokes_p2m_kernel<false><<<pg, 32 * warps, sh, s>>>(p2m_list, p2m_n, T.bbegin.p, T.bend.p, T.center.p,
                                                       d->src.p, P, d->M4[0].p, d->M4[1].p, d->M4[2].p, d->M4[3].p);
  ++plan->launches;
  // translations: the Laplace operators applied to each set (StokesSpherical.hpp:190-196,293-307)
  for (int k = 0; k < 4; ++k) {
    SetGuard g(plan, d, k);
    laplace_translations(plan, s);
  }
  if (T.n_own_leaves && plan->opts.evaluator == FMMB_EVAL_TREECODE)      // the translations stopped after the upward pass
    stokes_m2p_kernel<<<nblk(T.n_own_leaves, 4), 128, (size_t)4 * 4 * nc * sizeof(double2), s>>>(
        T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.parent.p, T.m2l_off.p, T.m2l_src.p, T.center.p, T.body.p, P,
        d->M4[0].p, d->M4[1].p, d->M4[2].p, d->M4[3].p, d->stresslet ? 1.0 / 6 : 1.0, d->res_far.p);
  else if (T.n_own_leaves)
  stokes_l2p_kernel<<<nblk(T.n_own_leaves, 4), 128, (size_t)4 * 4 * nc * sizeof(double2), s>>>(
      T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.center.p, T.has_local.p, T.body.p, P, d->L4[0].p,
      d->L4[1].p, d->L4[2].p, d->L4[3].p, d->stresslet ? 1.0 / 6 : 1.0, d->res_far.p);
  ++plan->launches;
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[4], s));

  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s, ev[7], 0));
  finish_results(plan, d->res_near.p, d->res_far.p, 3, d_results, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[5], s));
  FMMB_CUDA(cudaGetLastError());
  plan->timed = true;
}

void stokes_direct_raw(bool stresslet, const double* d_spts, const double* d_q, int64_t ns, const double* d_tpts,
                       int64_t nt, double* d_out, cudaStream_t s) {
  if (stresslet) stokes_direct_kernel<true><<<nblk(nt, 128), 128, 0, s>>>(d_spts, d_q, ns, d_tpts, nt, d_out);
  else stokes_direct_kernel<false><<<nblk(nt, 128), 128, 0, s>>>(d_spts, d_q, ns, d_tpts, nt, d_out);
  FMMB_CUDA(cudaGetLastError());
}

bool stokes_is_stresslet(const StokesData* d) { return d->stresslet; }

}  // namespace fmmb
