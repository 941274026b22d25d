// csrc/stokes_bem.cu -- StokesSphericalBEM on the GPU: triangular panels, 3 x 3 block near field cached once per
// plan, far field through two groups of four Laplace expansion sets.
//
// Replaces (reference kernel/StokesSphericalBEM.hpp):
//   :60-96    Panel(p0, p1, p2): centre, normal, area                       -> sbem_setup_kernel (bem::make_panel)
//   :377-390  operator() -> :257-375 eval_velocity_integral / :160-255 eval_traction_integral, evaluated for every
//             near pair into a CSR matrix of Mat3 by include/executor/EvalP2P.hpp:47-97
//                                                                          -> sbem_assemble_kernel (once per plan)
//   include/Matvec.hpp:14-33 (Mat3 x Vec<3> CSR matvec) in EvalInteractionLazySparse.hpp:134-151
//                                                                          -> sbem_near_kernel (per matvec)
//   :392-466  P2M with K quadrature points per panel: VELOCITY panels feed group 0 with the Stokeslet sets
//             f Y, (f.x) Y (f = Area w_i charge, x = quadrature point); TRACTION panels feed group 1 with the
//             stresslet sets (w_s . grad)(rho^n Y), g = Area w_i charge, n = panel normal
//                                                                          -> sbem_p2m_kernel<GROUP>
//   :468-472, :494-506  M2M / M2L / L2L of each set                         -> laplace_translations() per set
//   :508-527  L2P: r = StokesSpherical::L2P (kernel/StokesSpherical.hpp:318-401, scale 1) of the group the TARGET's
//             boundary condition picks; result += r / (2 mu) (VELOCITY) or 0.5 r (TRACTION) -> sbem_l2p_kernel
//   :474-492  M2P (treecode evaluator, `StokesBEM -eval TREE`), same scaling                -> sbem_m2p_kernel
//
// Near-field layout: block dense like csrc/bem.cu -- a work item is <= 32 targets of one leaf against the leaf's
// whole source list.  Both layers give SYMMETRIC 3 x 3 blocks (r^2 I + d d' and d d' scaled, the self terms too), so
// only the upper triangle is kept: entry e of (xx, xy, xz, yy, yz, zz) of (target lane, source j) sits at
// val[6 base + (j * 6 + e) * cnt + lane]; the per-matvec kernel streams 48 bytes per pair with coalesced loads
// (HBM bound: 48 B for 18 flop) instead of the 72 B of the reference's Mat3 CSR plus its indices.
//
// Reference quirks kept for parity: near-field entries as the reference computes them when
// compiled (K-point rule for every pair) unless FMMB_FLAG_STOKES_BEM_AS_WRITTEN asks for the branches of its source
// text, whose self term is the reference's reading of Fata's closed form (stokes_bem_math.hpp); the TRACTION far
// field enters with +0.5 although the near field carries -3 x the same integral (the reference's own FMM and Direct
// disagree for TRACTION targets; its driver only uses that plan for a right-hand side it then overwrites,
// examples/StokesBEM.cpp:256-271).
#include "common.cuh"
#include "laplace_ops.cuh"
#include "../hostcxx/stokes_bem_math.hpp"
#include <algorithm>

namespace fmmb {

struct StokesBemData {
  int K = 4, kfine = 19;
  double mu = 1e-3;
  bool as_written = false;                 // FMMB_FLAG_STOKES_BEM_AS_WRITTEN
  bool group_active[2] = {false, false};   // some panel carries VELOCITY (0) / TRACTION (1)
  DevBuf<bem::Panel> pan;          // tree order
  DevBuf<int> bc;                  // tree order: 0 VELOCITY, 1 TRACTION
  DevBuf<double> chg;              // tree order, 3 per panel
  DevBuf<double> nf_val;           // cached near field, 6 doubles per pair (upper triangle), block layout (see above)
  DevBuf<long long> nf_base;       // per work item: offset of its block in PAIRS
  int64_t nnz = 0;                 // pairs
  DevBuf<double> M4[4], L4[4];     // the four expansion sets of the group being evaluated
  DevBuf<double> res_near, res_far;  // tree order, 3 per panel
  int p_alloc = 0;
};

void stokes_bem_free(StokesBemData* d) { delete d; }
int64_t stokes_bem_nnz(const StokesBemData* d) { return d->nnz; }

namespace {

using namespace ops;

__constant__ bem::Rule c_srule;   // K-point panel rule
__constant__ bem::Rule c_sfine;   // fine rule for panels closer than 2 sqrt(2 Area)

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

__global__ void sbem_setup_kernel(const double* __restrict__ verts, const int* __restrict__ bc,
                                  const unsigned* __restrict__ perm, int64_t n, bem::Panel* __restrict__ pan,
                                  int* __restrict__ bc_tree) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* v = verts + 9 * (size_t)perm[i];
  bem::Panel p;
  bem::make_panel(v, v + 3, v + 6, p);
  pan[i] = p;
  bc_tree[i] = bc ? bc[perm[i]] : 0;
}

// per work item: pairs = targets x total source panels of the leaf's list
__global__ void sbem_count_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                                  const unsigned* __restrict__ be, const int* __restrict__ off,
                                  const int* __restrict__ src, long long* __restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nitems) return;
  long long c = 0;
  if (i < nitems) {
    const int4 it = items[i];
    long long ns = 0;
    for (int e = off[it.x]; e < off[it.x + 1]; ++e) ns += be[src[e]] - bb[src[e]];
    c = ns * it.z;
  }
  cnt[i] = c;
}

constexpr int kSbemWarps = 4;
constexpr int kSbemEntries = 6;   // stored entries per 3 x 3 block: xx, xy, xz, yy, yz, zz

// one warp per work item: lane = target, source panels staged through a warp-private tile
__global__ void __launch_bounds__(32 * kSbemWarps)
sbem_assemble_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                     const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
                     const bem::Panel* __restrict__ pan, const int* __restrict__ bc,
                     const long long* __restrict__ base, double mu, bool as_written, double* __restrict__ val) {
  __shared__ bem::Panel tiles[kSbemWarps][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kSbemWarps + wl;
  if (item >= nitems) return;
  bem::Panel* tile = tiles[wl];
  const int4 it = items[item];
  const int cnt = it.z;
  const bool act = lane < cnt;
  double tc[3] = {0, 0, 0};
  int tbc = 0;
  if (act) {
    const bem::Panel& t = pan[it.y + lane];
    tc[0] = t.c[0]; tc[1] = t.c[1]; tc[2] = t.c[2];
    tbc = bc[it.y + lane];
  }
  double* out = val + kSbemEntries * base[item];
  long long j = 0;
  for (int e = off[it.x]; e < off[it.x + 1]; ++e) {
    const int sb = src[e];
    const unsigned c0 = bb[sb], c1 = be[sb];
    for (unsigned b0 = c0; b0 < c1; b0 += 32) {
      const int ns = (int)min(32u, c1 - b0);
      __syncwarp();
      if (lane < ns) tile[lane] = pan[b0 + lane];
      __syncwarp();
      if (act)
        for (int k = 0; k < ns; ++k) {
          double m[9];
          bem::stokes_kernel(tbc, tc, tile[k], c_srule, c_sfine, mu, as_written, m);
          const double up[kSbemEntries] = {m[0], m[1], m[2], m[4], m[5], m[8]};     // m[3] = m[1], m[6] = m[2], m[7] = m[5]
#pragma unroll
          for (int q = 0; q < kSbemEntries; ++q) out[((j + k) * kSbemEntries + q) * cnt + lane] = up[q];
        }
      j += ns;
    }
  }
}

// results(targets of the item) = block * charges(sources); same traversal order as the assembly
__global__ void __launch_bounds__(32 * kSbemWarps)
sbem_near_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                 const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
                 const double* __restrict__ chg, const long long* __restrict__ base,
                 const double* __restrict__ val, double* __restrict__ res) {
  __shared__ double tiles[kSbemWarps][96];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kSbemWarps + wl;
  if (item >= nitems) return;
  double* tile = tiles[wl];
  const int4 it = items[item];
  const int cnt = it.z;
  const bool act = lane < cnt;
  const double* a = val + kSbemEntries * base[item] + lane;      // walks the block: one pair = 6 rows of cnt doubles
  const size_t cs = (size_t)cnt;
  double u0 = 0, u1 = 0, u2 = 0;
  for (int e = off[it.x]; e < off[it.x + 1]; ++e) {
    const int sb = src[e];
    const unsigned c0 = bb[sb], c1 = be[sb];
    for (unsigned b0 = c0; b0 < c1; b0 += 32) {
      const int ns = (int)min(32u, c1 - b0);
      __syncwarp();
      for (int k = lane; k < 3 * ns; k += 32) tile[k] = chg[3 * (size_t)b0 + k];   // contiguous: coalesced
      __syncwarp();
      if (act) {
#pragma unroll 2
        for (int k = 0; k < ns; ++k) {
          const double f0 = tile[3 * k], f1 = tile[3 * k + 1], f2 = tile[3 * k + 2];
          const double xx = a[0], xy = a[cs], xz = a[2 * cs], yy = a[3 * cs], yz = a[4 * cs], zz = a[5 * cs];
          a += kSbemEntries * cs;
          u0 = fma(xx, f0, fma(xy, f1, fma(xz, f2, u0)));
          u1 = fma(xy, f0, fma(yy, f1, fma(yz, f2, u1)));
          u2 = fma(xz, f0, fma(yz, f1, fma(zz, f2, u2)));
        }
      }
    }
  }
  if (act) {
    double* o = res + 3 * (size_t)(it.y + lane);
    o[0] = u0; o[1] = u1; o[2] = u2;
  }
}

// The same product with kSbemSplit warps per work item (round 2; see bem_near_split_kernel in csrc/bem.cu): the
// list is walked 32 source leaves at a time, row offsets from a warp scan of the leaf sizes; the charges of a chunk
// are staged in shared memory and its pairs split evenly over the warps, four pairs (24 loads) in flight per lane;
// partial sums are added in warp order (same bits every run).  A chunk with more pairs than the staging buffer deals
// its leaves to the warps instead.
constexpr int kSbemSplit = 8;
constexpr int kSbemRows = 1024;

__device__ __forceinline__ void sbem_pair(const double* __restrict__ a, size_t cs, double g0, double g1, double g2,
                                          double& u0, double& u1, double& u2) {
  double m[kSbemEntries];
#pragma unroll
  for (int q = 0; q < kSbemEntries; ++q) m[q] = __ldg(a + q * cs);
  u0 = fma(m[0], g0, fma(m[1], g1, fma(m[2], g2, u0)));
  u1 = fma(m[1], g0, fma(m[3], g1, fma(m[4], g2, u1)));
  u2 = fma(m[2], g0, fma(m[4], g1, fma(m[5], g2, u2)));
}

__global__ void __launch_bounds__(32 * kSbemSplit, 2)
sbem_near_split_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                       const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
                       const double* __restrict__ chg, const long long* __restrict__ base,
                       const double* __restrict__ val, double* __restrict__ res) {
  __shared__ double qs[3 * kSbemRows];
  __shared__ double part[kSbemSplit][3][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int4 it = items[item];
  const int cnt = it.z;
  const bool act = lane < cnt;
  const double* in = val + kSbemEntries * base[item] + (act ? lane : 0);
  const size_t cs = (size_t)cnt;
  double u0 = 0, u1 = 0, u2 = 0;
  const int e0 = off[it.x], e1 = off[it.x + 1];
  long long jbase = 0;
  for (int ec = e0; ec < e1; ec += 32) {
    unsigned c0 = 0, ns_l = 0;
    if (ec + lane < e1) {
      const int sb = src[ec + lane];
      c0 = bb[sb];
      ns_l = be[sb] - c0;
    }
    unsigned incl = ns_l;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const int nent = min(32, e1 - ec);
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    if (total <= (unsigned)kSbemRows) {
      for (int l = wl; l < nent; l += kSbemSplit) {
        const unsigned b0 = __shfl_sync(0xffffffffu, c0, l);
        const unsigned ns = __shfl_sync(0xffffffffu, ns_l, l);
        const unsigned r0 = __shfl_sync(0xffffffffu, incl, l) - ns;
        for (unsigned t = lane; t < 3 * ns; t += 32) qs[3 * r0 + t] = chg[3 * (size_t)b0 + t];   // contiguous
      }
      __syncthreads();
      const int ra = (int)((unsigned long long)total * wl / kSbemSplit);
      const int rb = (int)((unsigned long long)total * (wl + 1) / kSbemSplit);
      if (act) {
        const double* a = in + (size_t)(jbase + ra) * kSbemEntries * cs;
        int k = ra;
        for (; k + 4 <= rb; k += 4, a += 4 * kSbemEntries * cs) {
          double m[4][kSbemEntries];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < kSbemEntries; ++q) m[u][q] = __ldg(a + ((size_t)u * kSbemEntries + q) * cs);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double g0 = qs[3 * (k + u)], g1 = qs[3 * (k + u) + 1], g2 = qs[3 * (k + u) + 2];
            u0 = fma(m[u][0], g0, fma(m[u][1], g1, fma(m[u][2], g2, u0)));
            u1 = fma(m[u][1], g0, fma(m[u][3], g1, fma(m[u][4], g2, u1)));
            u2 = fma(m[u][2], g0, fma(m[u][4], g1, fma(m[u][5], g2, u2)));
          }
        }
        for (; k < rb; ++k, a += kSbemEntries * cs) sbem_pair(a, cs, qs[3 * k], qs[3 * k + 1], qs[3 * k + 2], u0, u1, u2);
      }
      __syncthreads();                       // qs is overwritten by the next chunk
    } else {
      for (int l = wl; l < nent; l += kSbemSplit) {
        const unsigned b0 = __shfl_sync(0xffffffffu, c0, l);
        const int ns = (int)__shfl_sync(0xffffffffu, ns_l, l);
        const long long j0 = jbase + (long long)(__shfl_sync(0xffffffffu, incl, l) - (unsigned)ns);
        for (int t0 = 0; t0 < ns; t0 += 32) {
          const int nt = min(32, ns - t0);
          double f0 = 0, f1 = 0, f2 = 0;       // lane k holds the three charge components of source t0 + k
          if (lane < nt) {
            const double* c = chg + 3 * (size_t)(b0 + t0 + lane);
            f0 = c[0]; f1 = c[1]; f2 = c[2];
          }
          const double* a = in + (size_t)(j0 + t0) * kSbemEntries * cs;
          for (int k = 0; k < nt; ++k, a += kSbemEntries * cs) {
            const double g0 = __shfl_sync(0xffffffffu, f0, k), g1 = __shfl_sync(0xffffffffu, f1, k),
                         g2 = __shfl_sync(0xffffffffu, f2, k);
            if (act) sbem_pair(a, cs, g0, g1, g2, u0, u1, u2);
          }
        }
      }
    }
    jbase += total;
  }
  part[wl][0][lane] = u0; part[wl][1][lane] = u1; part[wl][2][lane] = u2;
  __syncthreads();
  if (wl < 3 && act) {                       // warp c sums component c
    double t = part[0][wl][lane];
#pragma unroll
    for (int w = 1; w < kSbemSplit; ++w) t += part[w][wl][lane];
    res[3 * (size_t)(it.y + lane) + wl] = t;
  }
}

// charges (original order, 3 per panel) into tree order
__global__ void sbem_gather(const double* __restrict__ q, const unsigned* __restrict__ perm, int64_t n,
                            double* __restrict__ chg) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  const int64_t i = t / 3;
  const int c = (int)(t - 3 * i);
  chg[t] = q[3 * (size_t)perm[i] + c];
}

// ---- P2M: warp per (leaf, set); lane = (panel, quadrature point) writes its row, then lane = coefficient sums
// the column.  GROUP 0: VELOCITY panels, Stokeslet sets; GROUP 1: TRACTION panels, stresslet sets.
template <int GROUP>
__global__ void __launch_bounds__(128)
sbem_p2m_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                const unsigned* __restrict__ be, const double4* __restrict__ center,
                const bem::Panel* __restrict__ pan, const int* __restrict__ bc, const double* __restrict__ chg, int P,
                double* __restrict__ M0, double* __restrict__ M1, double* __restrict__ M2, double* __restrict__ M3) {
  extern __shared__ double sbem_sh[];
  const int pp = P * P, ld = pp | 1;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  const int set = blockIdx.y;
  if (w >= nleaves) return;
  double* tile = sbem_sh + (size_t)wl * 32 * ld;
  const int b = leaves[w];
  const double4 c = center[b];
  const unsigned b0 = bb[b], b1 = be[b];
  const int K = c_srule.n;
  const int nent = (int)(b1 - b0) * K;
  double acc[(FMMB_MAX_P * FMMB_MAX_P + 31) / 32];
#pragma unroll
  for (int i = 0; i < (FMMB_MAX_P * FMMB_MAX_P + 31) / 32; ++i) acc[i] = 0.0;
  for (int base = 0; base < nent; base += 32) {
    const int ent = base + lane;
    const int cnt = min(32, nent - base);
    __syncwarp();
    if (ent < nent) {
      const unsigned i = b0 + ent / K;
      const int qi = ent % K;
      double* row = tile + lane * ld;
      if (bc[i] != GROUP) {
        for (int r = 0; r < pp; ++r) row[r] = 0.0;
      } else {
        const bem::Panel& s = pan[i];
        double q[3];
        bem::quad_point(s, c_srule.pt[qi], q);
        const double wa = s.area * c_srule.w[qi];
        const double g0 = wa * chg[3 * (size_t)i], g1 = wa * chg[3 * (size_t)i + 1], g2 = wa * chg[3 * (size_t)i + 2];
        const Sph sp = to_sph(q[0] - c.x, q[1] - c.y, q[2] - c.z);
        if (GROUP == 0) {
          const double mult = set == 0 ? g0 : (set == 1 ? g1 : (set == 2 ? g2 : g0 * q[0] + g1 * q[1] + g2 * q[2]));
          regular_harmonics<false>(P, sp, -1.0, [&](int n, int m, double yr, double yi, double, double) {
            row[n * n + n + m] = mult * yr;
            if (m > 0) row[n * n + n - m] = mult * yi;
          });
        } else {
          const double n0 = s.nrm[0], n1 = s.nrm[1], n2 = s.nrm[2];
          // direction of the derivative for this set: w_s = g_s n + n_s g (s < 3), w_3 = (x.g) n + (n.x) g
          double a, bq;
          if (set == 0) { a = g0; bq = n0; }
          else if (set == 1) { a = g1; bq = n1; }
          else if (set == 2) { a = g2; bq = n2; }
          else { a = q[0] * g0 + q[1] * g1 + q[2] * g2; bq = n0 * q[0] + n1 * q[1] + n2 * q[2]; }
          const double w0 = a * n0 + bq * g0, w1 = a * n1 + bq * g1, w2 = a * n2 + bq * g2;
          // spherical basis vectors over the metric: grad = e_r d/drho + e_a/rho d/dalpha + e_b/(rho sin) d/dbeta
          const double ir = 1.0 / sp.r, iry = ir / sp.y;
          const double wa_ = w0 * (sp.y * sp.cp) + w1 * (sp.y * sp.sp) + w2 * sp.x;
          const double wb = (w0 * (sp.x * sp.cp) + w1 * (sp.x * sp.sp) - w2 * sp.y) * ir;
          const double wc = (-w0 * sp.sp + w1 * sp.cp) * iry;
          regular_harmonics<true>(P, sp, -1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
            // brh = n/rho Y, bal = Ytheta, bbe = -i m Y = (m yi, -m yr)
            const double fr = n * ir;
            row[n * n + n + m] = wa_ * fr * yr + wb * tr + wc * (m * yi);
            if (m > 0) row[n * n + n - m] = wa_ * fr * yi + wb * ti - wc * (m * yr);
          });
        }
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

// ---- L2P: warp per leaf, lane per target panel (its centre), all four sets in one pass over the harmonics; only
// targets whose boundary condition selects GROUP are written (res is zeroed before the first group)
__global__ void __launch_bounds__(128)
sbem_l2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                const unsigned* __restrict__ be, const double4* __restrict__ center,
                const unsigned char* __restrict__ has_local, const bem::Panel* __restrict__ pan,
                const int* __restrict__ bc, int group, int P, const double* __restrict__ L0,
                const double* __restrict__ L1, const double* __restrict__ L2, const double* __restrict__ L3,
                double scale, double* __restrict__ res) {
  extern __shared__ double2 sbem_ls[];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int b = leaves[w];
  if (!has_local[b]) return;
  const unsigned b0 = bb[b], b1 = be[b];
  double2* Ls = sbem_ls + (size_t)wl * 4 * nc;
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
    if (bc[i] != group) continue;
    const double px = pan[i].c[0], py = pan[i].c[1], pz = pan[i].c[2];
    const Sph s = to_sph(px - c.x, py - c.y, pz - c.z);
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
    const double xs_[3] = {px, py, pz};
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

// ---- M2P (treecode, `StokesBEM -eval TREE`; :474-492 over StokesSpherical.hpp:207-291): warp per leaf, lane per
// target panel; every source box accepted for the leaf or one of its ancestors is evaluated at the centres of the
// panels whose boundary condition selects GROUP -- the L2P arithmetic on the singular harmonics (radial factor
// -(n+1)/r), converted to Cartesian per source box.  res is zeroed before the first group.
__global__ void __launch_bounds__(128)
sbem_m2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
                const unsigned* __restrict__ be, const unsigned* __restrict__ parent, const int* __restrict__ off,
                const int* __restrict__ srcbox, const double4* __restrict__ center, const bem::Panel* __restrict__ pan,
                const int* __restrict__ bc, int group, int P, const double* __restrict__ M0,
                const double* __restrict__ M1, const double* __restrict__ M2, const double* __restrict__ M3,
                double scale, double* __restrict__ res) {
  extern __shared__ double2 sbem_ms[];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double2* Ms = sbem_ms + (size_t)wl * 4 * nc;
  const int leaf = leaves[w];
  const unsigned b0 = bb[leaf], b1 = be[leaf];
  for (unsigned base = b0; base < b1; base += 32) {
    const unsigned i = base + lane;
    const bool act = i < b1 && bc[i] == group;
    double xs_[3] = {0, 0, 0};
    if (act) { xs_[0] = pan[i].c[0]; xs_[1] = pan[i].c[1]; xs_[2] = pan[i].c[2]; }
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
          const Sph s = to_sph(xs_[0] - c.x, xs_[1] - c.y, xs_[2] - c.z);
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

// Direct::matvec with the panel kernel (include/Direct.hpp:99-125 over operator(), :377-390): block per target panel,
// threads stride over ALL source panels of the plan, block-wide sums in a fixed tree order.
__global__ void __launch_bounds__(128)
sbem_direct_kernel(const bem::Panel* __restrict__ pan, const double* __restrict__ chg, int64_t ns,
                   const double* __restrict__ tverts, const int* __restrict__ tbc, double mu, bool as_written,
                   double* __restrict__ out) {
  __shared__ double part[3][128];
  const int64_t t = blockIdx.x;
  const double* v = tverts + 9 * (size_t)t;
  const double tc[3] = {((v[0] + v[3]) + v[6]) / 3, ((v[1] + v[4]) + v[7]) / 3, ((v[2] + v[5]) + v[8]) / 3};
  const int bc = tbc ? tbc[t] : 0;
  double u0 = 0, u1 = 0, u2 = 0;
  for (int64_t j = threadIdx.x; j < ns; j += blockDim.x) {
    double m[9];
    bem::stokes_kernel(bc, tc, pan[j], c_srule, c_sfine, mu, as_written, m);
    const double f0 = chg[3 * j], f1 = chg[3 * j + 1], f2 = chg[3 * j + 2];
    u0 += m[0] * f0 + m[1] * f1 + m[2] * f2;
    u1 += m[3] * f0 + m[4] * f1 + m[5] * f2;
    u2 += m[6] * f0 + m[7] * f1 + m[8] * f2;
  }
  part[0][threadIdx.x] = u0; part[1][threadIdx.x] = u1; part[2][threadIdx.x] = u2;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w)
      for (int c = 0; c < 3; ++c) part[c][threadIdx.x] += part[c][threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x < 3) out[3 * (size_t)t + threadIdx.x] = part[threadIdx.x][0];
}

void swap_buf(DevBuf<double>& a, DevBuf<double>& b) {
  std::swap(a.p, b.p); std::swap(a.cap, b.cap); std::swap(a.n, b.n);
}
struct SetGuard {            // plan->M / plan->L temporarily ARE set k of the group being evaluated
  fmmb_plan* plan; StokesBemData* d; int k;
  SetGuard(fmmb_plan* pl, StokesBemData* dd, int kk) : plan(pl), d(dd), k(kk) { swap_buf(plan->M, d->M4[k]); swap_buf(plan->L, d->L4[k]); }
  ~SetGuard() { swap_buf(plan->M, d->M4[k]); swap_buf(plan->L, d->L4[k]); }
};

}  // namespace

// Plan-time: panel geometry in tree order, then the cached 3 x 3 block near field.
void stokes_bem_setup(fmmb_plan* plan, const double* verts_host, const int32_t* bc_host, int quad_k, int quad_kfine,
                      double mu) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  if (quad_kfine <= 0) quad_kfine = 25;      // StokesSphericalBEM(p, k, mu) leaves K_fine = 25 (:136)
  if (!bem::rule_supported(quad_k) || !bem::rule_supported(quad_kfine))
    throw StatusError{FMMB_ERR_UNSUPPORTED, "quad_k / quad_kfine must be keys of the reference's Gauss table: 1, 3, 4, 7, 13, 17, 19, 25 or 79"};
  if (!(mu > 0)) throw StatusError{FMMB_ERR_INVALID, "StokesSphericalBEM needs a positive viscosity (fmmb_kernel_desc.kappa)"};
  StokesBemData* B = new StokesBemData();
  plan->sbem = B;
  B->K = quad_k; B->kfine = quad_kfine; B->mu = mu;
  B->as_written = (plan->opts.kernel_flags & FMMB_FLAG_STOKES_BEM_AS_WRITTEN) != 0;
  upload_laplace_tables();   // this translation unit's copy of the factorial tables
  const bem::Rule rule = bem::make_rule(quad_k), fine = bem::make_rule(quad_kfine);
  FMMB_CUDA(cudaMemcpyToSymbol(c_srule, &rule, sizeof rule));
  FMMB_CUDA(cudaMemcpyToSymbol(c_sfine, &fine, sizeof fine));
  const int64_t n = T.n;
  DevBuf<double> verts;
  DevBuf<int> bc;
  verts.from_host(verts_host, 9 * (size_t)n, s);
  for (int64_t i = 0; i < n; ++i) {
    const int v = bc_host ? bc_host[i] : 0;
    if (v != 0 && v != 1) throw StatusError{FMMB_ERR_INVALID, "bc entries must be 0 (VELOCITY) or 1 (TRACTION)"};
    B->group_active[v] = true;
  }
  if (bc_host) bc.from_host(bc_host, n, s);
  B->pan.resize(n); B->bc.resize(n); B->chg.resize(3 * (size_t)n);
  sbem_setup_kernel<<<nblk(n, 128), 128, 0, s>>>(verts.p, bc_host ? bc.p : nullptr, T.perm.p, n, B->pan.p, B->bc.p);
  FMMB_CUDA(cudaGetLastError());
  const int ni = T.n_p2p_items;
  DevBuf<long long> cnt;
  cnt.resize(ni + 1);
  sbem_count_kernel<<<nblk(ni + 1, 128), 128, 0, s>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p, T.p2p_off.p,
                                                     T.p2p_src.p, cnt.p);
  FMMB_CUDA(cudaGetLastError());
  std::vector<long long> h = cnt.to_host(s), off(ni + 1, 0);
  for (int i = 0; i < ni; ++i) off[i + 1] = off[i] + h[i];
  B->nnz = off[ni];
  B->nf_base.from_host(off.data(), off.size(), s);
  B->nf_val.resize(kSbemEntries * (size_t)B->nnz);
  if (ni)
    sbem_assemble_kernel<<<nblk(ni, kSbemWarps), 32 * kSbemWarps, 0, s>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p,
                                                                         T.p2p_off.p, T.p2p_src.p, B->pan.p, B->bc.p,
                                                                         B->nf_base.p, B->mu, B->as_written, B->nf_val.p);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));
}

// fmmb_plan_direct_panels for StokesSphericalBEM plans (device pointers; charges original order, 3 per panel)
void stokes_bem_direct(fmmb_plan* plan, const double* d_charges, int64_t nt, const double* d_tverts, const int* d_tbc,
                       double* d_out, cudaStream_t s) {
  Tree& T = plan->tree;
  StokesBemData* B = plan->sbem;
  DevBuf<double> chg;
  chg.resize(3 * (size_t)T.n);
  sbem_gather<<<nblk(3 * T.n, 256), 256, 0, s>>>(d_charges, T.perm.p, T.n, chg.p);
  if (nt) sbem_direct_kernel<<<(unsigned)nt, 128, 0, s>>>(B->pan.p, chg.p, T.n, d_tverts, d_tbc, B->mu, B->as_written, d_out);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));      // chg is released on return
}

// see stokes_prepare_expansions (stokes.cu)
void stokes_bem_prepare_expansions(fmmb_plan* plan) {
  StokesBemData* B = plan->sbem;
  const int P = plan->p, xs = xstride(P);
  for (int k = 0; k < 4; ++k) {
    B->M4[k].resize((size_t)(plan->tree.nboxes + 1) * xs);
    B->L4[k].resize((size_t)(plan->tree.nboxes + 1) * xs);
    if (B->p_alloc != P) { B->M4[k].zero(plan->stream); B->L4[k].zero(plan->stream); }
  }
  B->p_alloc = P;
}

void stokes_bem_execute(fmmb_plan* plan, const double* d_charges, double* d_results) {
  Tree& T = plan->tree;
  StokesBemData* B = plan->sbem;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P;
  const int xs = xstride(P);
  const int64_t n = T.n;
  cudaStream_t s = plan->stream;
  cudaEvent_t* ev = plan->ev;
  stokes_bem_prepare_expansions(plan);
  B->res_near.resize(3 * (size_t)n);
  B->res_far.resize(3 * (size_t)n);
  plan->launches = 0;

  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[0], s));
  sbem_gather<<<nblk(3 * n, 256), 256, 0, s>>>(exec_charges(plan, d_charges), exec_perm(plan), n, B->chg.p);
  ++plan->launches;
  FMMB_CUDA(cudaEventRecord(ev[1], s));

  // cached near field (one pass over 48 bytes per pair) on the second stream, beside the far-field chain of short
  // dependent kernels (round 2; joined before the results are combined)
  cudaStream_t s2 = plan->overlap_p2p ? plan->stream2 : s;
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s2, ev[1], 0));
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[6], s2));
  const int ni = T.n_p2p_items;
  if (ni) {
    if (plan->bem_near_kernel)
      sbem_near_split_kernel<<<ni, 32 * kSbemSplit, 0, s2>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p, T.p2p_off.p,
                                                            T.p2p_src.p, B->chg.p, B->nf_base.p, B->nf_val.p, B->res_near.p);
    else
      sbem_near_kernel<<<nblk(ni, kSbemWarps), 32 * kSbemWarps, 0, s2>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p,
                                                                        T.p2p_off.p, T.p2p_src.p, B->chg.p, B->nf_base.p,
                                                                        B->nf_val.p, B->res_near.p);
    ++plan->launches;
  }
  FMMB_CUDA(cudaEventRecord(ev[7], s2));
  B->res_far.zero(s);
  ++plan->launches;

  // far field, one group of four sets at a time
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[12], s));
  const int warps = pp <= 64 ? 4 : 1;
  const size_t sh = (size_t)warps * 32 * (pp | 1) * sizeof(double);
  // per call: function attributes belong to the current device
  FMMB_CUDA(cudaFuncSetAttribute(sbem_p2m_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  FMMB_CUDA(cudaFuncSetAttribute(sbem_p2m_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  for (int group = 0; group < 2; ++group) {
    if (!B->group_active[group] || plan->near_only) continue;   // near_only: plans for preconditioners
    const dim3 pg(nblk(T.nleaves, warps), 4);
    if (group == 0)
      sbem_p2m_kernel<0><<<pg, 32 * warps, sh, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p, T.center.p, B->pan.p,
                                                    B->bc.p, B->chg.p, P, B->M4[0].p, B->M4[1].p, B->M4[2].p,
                                                    B->M4[3].p);
    else
      sbem_p2m_kernel<1><<<pg, 32 * warps, sh, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p, T.center.p, B->pan.p,
                                                    B->bc.p, B->chg.p, P, B->M4[0].p, B->M4[1].p, B->M4[2].p,
                                                    B->M4[3].p);
    ++plan->launches;
    for (int k = 0; k < 4; ++k) {
      SetGuard g(plan, B, k);
      laplace_translations(plan, s);
    }
    // VELOCITY: result += r / (2 mu);  TRACTION: result += 0.5 r  (:508-527)
    const double scale = group == 0 ? 1. / 2 / B->mu : 0.5;
    if (T.n_own_leaves && plan->opts.evaluator == FMMB_EVAL_TREECODE)    // the translations stopped after the upward pass
      sbem_m2p_kernel<<<nblk(T.n_own_leaves, 4), 128, (size_t)4 * 4 * nc * sizeof(double2), s>>>(
          T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.parent.p, T.m2l_off.p, T.m2l_src.p, T.center.p, B->pan.p,
          B->bc.p, group, P, B->M4[0].p, B->M4[1].p, B->M4[2].p, B->M4[3].p, scale, B->res_far.p);
    else if (T.n_own_leaves)
      sbem_l2p_kernel<<<nblk(T.n_own_leaves, 4), 128, (size_t)4 * 4 * nc * sizeof(double2), s>>>(
          T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.center.p, T.has_local.p, B->pan.p, B->bc.p, group, P,
          B->L4[0].p, B->L4[1].p, B->L4[2].p, B->L4[3].p, scale, B->res_far.p);
    ++plan->launches;
  }
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[4], s));
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s, ev[7], 0));
  finish_results(plan, B->res_near.p, B->res_far.p, 3, d_results, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[5], s));
  FMMB_CUDA(cudaGetLastError());
  plan->timed = true;
}

}  // namespace fmmb
