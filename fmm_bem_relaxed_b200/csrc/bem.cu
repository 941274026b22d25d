// csrc/bem.cu -- LaplaceSphericalBEM on the GPU: panel sources, Gauss-quadrature P2M, cached near field.
//
// Replaces (reference paths):
//   kernel/LaplaceSphericalBEM.hpp:61-97     Panel geometry                 -> bem_setup_kernel
//   :273-297 operator() + :159-264 eval_G / eval_dGdn + examples/BEM/SemiAnalytical.hpp
//        evaluated for every near pair by executor/EvalP2P.hpp:47-97        -> bem_assemble_kernel (once per plan)
//   include/Matvec.hpp:14-33 CSR matvec in EvalInteractionLazySparse.hpp:134-151 -> bem_near_kernel (per matvec)
//   :307-352 P2M with K quadrature points per panel, two expansion sets     -> bem_p2m_kernel<SET>
//   :448-476 L2P (scalar result, set and sign picked by the target's BC)     -> bem_l2p_kernel<SET>
//   :394-422 vector M2P (treecode evaluator, `LaplaceBEM -eval TREE`)        -> bem_m2p_kernel<SET>
//   M2M / M2L / L2L of each set (:362-383,432-437) are the Laplace translations -> laplace_translations()
//
// Near field layout.  All targets of a leaf share one source list (the P2P list of the leaf), so the
// cached near-field matrix is block dense: for a work item (<= 32 targets of one leaf) the block is
// stored source-major, val[base + j * cnt + lane], without column indices; the per-matvec kernel
// streams it once (8 bytes per entry) with coalesced loads -- HBM bound.
//
// Expansion sets: set 0 = single layer (G), fed by POTENTIAL panels and read by
// POTENTIAL targets; set 1 = double layer (dG/dn), fed and read by NORMAL_DERIV panels.  Only the sets
// that have panels are run (the reference's examples use one boundary condition at a time).
#include "common.cuh"
#include "laplace_ops.cuh"
#include "../hostcxx/bem_math.hpp"

namespace fmmb {

struct BemData {
  int K = 4;
  double kappa = -1.0;           // >= 0: YukawaCartesianBEM near field; < 0: LaplaceSphericalBEM
  bool set_active[2] = {false, false};
  DevBuf<bem::Panel> pan;        // tree order
  DevBuf<int> bc;                // tree order: 0 POTENTIAL, 1 NORMAL_DERIV
  DevBuf<double> nf_val;         // cached near field, block layout (see above)
  DevBuf<long long> nf_base;     // per P2P work item: offset of its block
  DevBuf<double> res_near, res_far;
  int64_t nnz = 0;
};

void bem_free(BemData* b) { delete b; }
int64_t bem_nnz(const BemData* b) { return b->nnz; }

namespace {

using namespace ops;

__constant__ bem::Rule c_rule;   // K-point panel rule
__constant__ bem::Rule c_fine;   // the rule filed under 17 (16 points) for the near-singular double layer

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

__global__ void bem_setup_kernel(const double* __restrict__ verts, const int* __restrict__ bc,
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

// per work item: number of cached entries = targets x total source bodies of the leaf's list
__global__ void bem_count_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                                 const unsigned* __restrict__ be, const int* __restrict__ off,
                                 const int* __restrict__ src, long long* __restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nitems) return;
  long long c = 0;
  if (i < nitems) {
    int4 it = items[i];
    long long ns = 0;
    for (int e = off[it.x]; e < off[it.x + 1]; ++e) ns += be[src[e]] - bb[src[e]];
    c = ns * it.z;
  }
  cnt[i] = c;
}

constexpr int kBemWarps = 4;

// one warp per work item: lane = target, sources staged through a warp-private tile of panels
__global__ void __launch_bounds__(32 * kBemWarps)
bem_assemble_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                    const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
                    const bem::Panel* __restrict__ pan, const int* __restrict__ bc,
                    const long long* __restrict__ base, double kappa, double* __restrict__ val) {
  __shared__ bem::Panel tiles[kBemWarps][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kBemWarps + wl;
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
  double* out = val + base[item];
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
        for (int k = 0; k < ns; ++k)
          out[(j + k) * cnt + lane] = kappa < 0 ? bem::kernel(tbc, tc, tile[k], c_rule, c_fine)
                                                : bem::kernel_yk(tbc, tc, tile[k], c_rule, kappa);
      j += ns;
    }
  }
}

// results(targets of the item) = block * charges(sources); same traversal order as the assembly
__global__ void __launch_bounds__(32 * kBemWarps)
bem_near_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
                const double4* __restrict__ body, const long long* __restrict__ base,
                const double* __restrict__ val, double* __restrict__ res) {
  __shared__ double tiles[kBemWarps][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kBemWarps + wl;
  if (item >= nitems) return;
  double* tile = tiles[wl];
  const int4 it = items[item];
  const int cnt = it.z;
  const bool act = lane < cnt;
  const double* in = val + base[item] + lane;
  double a0 = 0, a1 = 0;
  long long j = 0;
  for (int e = off[it.x]; e < off[it.x + 1]; ++e) {
    const int sb = src[e];
    const unsigned c0 = bb[sb], c1 = be[sb];
    for (unsigned b0 = c0; b0 < c1; b0 += 32) {
      const int ns = (int)min(32u, c1 - b0);
      __syncwarp();
      if (lane < ns) tile[lane] = body[b0 + lane].w;
      __syncwarp();
      if (act) {
        int k = 0;
        for (; k + 2 <= ns; k += 2) {
          a0 = fma(in[(j + k) * cnt], tile[k], a0);
          a1 = fma(in[(j + k + 1) * cnt], tile[k + 1], a1);
        }
        if (k < ns) a0 = fma(in[(j + k) * cnt], tile[k], a0);
      }
      j += ns;
    }
  }
  if (act) res[it.y + lane] = a0 + a1;
}

// The same product with kNearSplit warps per work item (round 2).  One warp per item leaves ~11 warps per SM, each
// with two loads in flight: 139 MB of config C2's cached entries took 221 us (0.6 TB/s).  Here a block of eight
// warps takes the item.  The list is walked 32 source leaves at a time (row offsets from a warp scan of the leaf
// sizes); the charges of such a chunk are staged in shared memory and its rows are split EVENLY over the warps
// (dealing whole leaves left the warps 2 or 3 leaves each: ncu showed 6.4 barrier stalls per issue), every lane
// keeps eight independent loads in flight, and the warps' partial sums are added in warp order (fixed order: same
// bits on every run).  A chunk with more rows than the staging buffer (huge ncrit) deals its leaves to the warps.
constexpr int kNearSplit = 8;
constexpr int kNearRows = 2048;

__global__ void __launch_bounds__(32 * kNearSplit, 4)
bem_near_split_kernel(const int4* __restrict__ items, int nitems, const unsigned* __restrict__ bb,
                      const unsigned* __restrict__ be, const int* __restrict__ off, const int* __restrict__ src,
                      const double4* __restrict__ body, const long long* __restrict__ base,
                      const double* __restrict__ val, double* __restrict__ res) {
  __shared__ double qs[kNearRows];
  __shared__ double part[kNearSplit][32];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int4 it = items[item];
  const int cnt = it.z;
  const bool act = lane < cnt;
  const double* in = val + base[item] + (act ? lane : 0);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int e0 = off[it.x], e1 = off[it.x + 1];
  long long jbase = 0;
  for (int ec = e0; ec < e1; ec += 32) {
    // lane l: source leaf ec + l of the list -- first body, size, offset of its rows in the block
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
    if (total <= (unsigned)kNearRows) {
      // stage the chunk's charges (leaves dealt to the warps), then an even share of its rows per warp
      for (int l = wl; l < nent; l += kNearSplit) {
        const unsigned b0 = __shfl_sync(0xffffffffu, c0, l);
        const unsigned ns = __shfl_sync(0xffffffffu, ns_l, l);
        const unsigned r0 = __shfl_sync(0xffffffffu, incl, l) - ns;
        for (unsigned t = lane; t < ns; t += 32) qs[r0 + t] = body[b0 + t].w;
      }
      __syncthreads();
      const int ra = (int)((unsigned long long)total * wl / kNearSplit);
      const int rb = (int)((unsigned long long)total * (wl + 1) / kNearSplit);
      const double* row = in + (jbase + ra) * cnt;
      int k = ra;
      if (act) {
        for (; k + 8 <= rb; k += 8, row += (size_t)8 * cnt) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldg(row + (size_t)u * cnt);
#pragma unroll
          for (int u = 0; u < 8; u += 4) {
            a0 = fma(v[u], qs[k + u], a0);
            a1 = fma(v[u + 1], qs[k + u + 1], a1);
            a2 = fma(v[u + 2], qs[k + u + 2], a2);
            a3 = fma(v[u + 3], qs[k + u + 3], a3);
          }
        }
        for (; k < rb; ++k, row += cnt) a0 = fma(__ldg(row), qs[k], a0);
      }
      __syncthreads();                       // qs is overwritten by the next chunk
    } else {
      for (int l = wl; l < nent; l += kNearSplit) {
        const unsigned b0 = __shfl_sync(0xffffffffu, c0, l);
        const int ns = (int)__shfl_sync(0xffffffffu, ns_l, l);
        const long long j0 = jbase + (long long)(__shfl_sync(0xffffffffu, incl, l) - (unsigned)ns);
        for (int t0 = 0; t0 < ns; t0 += 32) {
          const int nt = min(32, ns - t0);
          const double q = lane < nt ? body[b0 + t0 + lane].w : 0.0;
          const double* row = in + (j0 + t0) * cnt;
          for (int k = 0; k < nt; ++k) {
            const double v = act ? __ldg(row + (size_t)k * cnt) : 0.0;
            a0 = fma(v, __shfl_sync(0xffffffffu, q, k), a0);
          }
        }
      }
    }
    jbase += total;
  }
  part[wl][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (wl == 0 && act) {
    double t = part[0][lane];
#pragma unroll
    for (int w = 1; w < kNearSplit; ++w) t += part[w][lane];
    res[it.y + lane] = t;
  }
}

// P2M: warp per leaf; lane = (panel, quadrature point); rows go through a shared tile, then lane = coefficient
template <int SET>
__global__ void __launch_bounds__(128)
bem_p2m_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
               const unsigned* __restrict__ be, const double4* __restrict__ center,
               const double4* __restrict__ body, const bem::Panel* __restrict__ pan, const int* __restrict__ bc,
               int P, double* __restrict__ M) {
  extern __shared__ double bem_sh[];
  const int pp = P * P, ld = pp | 1;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double* tile = bem_sh + (size_t)wl * 32 * ld;
  const int b = leaves[w];
  const double4 c = center[b];
  const unsigned b0 = bb[b], b1 = be[b];
  const int K = c_rule.n;
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
      if (bc[i] != SET) {
        for (int r = 0; r < pp; ++r) row[r] = 0.0;
      } else {
        const bem::Panel& s = pan[i];
        double q[3];
        bem::quad_point(s, c_rule.pt[qi], q);
        const double mult = body[i].w * c_rule.w[qi] * s.area;
        const Sph sp = to_sph(q[0] - c.x, q[1] - c.y, q[2] - c.z);
        if (SET == 0) {
          regular_harmonics<false>(P, sp, -1.0, [&](int n, int m, double yr, double yi, double, double) {
            row[n * n + n + m] = mult * yr;
            if (m > 0) row[n * n + n - m] = mult * yi;
          });
        } else {
          // (n . grad)(rho^n Y_n^m): spherical components -> Cartesian, LaplaceSphericalBEM.hpp:331-344
          const double ir = 1.0 / sp.r, iry = ir / sp.y;
          const double ax = sp.y * sp.cp, ay = sp.y * sp.sp, az = sp.x;                       // d/d rho
          const double bx = sp.x * sp.cp * ir, by = sp.x * sp.sp * ir, bz = -sp.y * ir;        // d/d alpha
          const double cx = -sp.sp * iry, cy = sp.cp * iry;                                    // d/d beta
          const double na = s.nrm[0] * ax + s.nrm[1] * ay + s.nrm[2] * az;
          const double nb_ = s.nrm[0] * bx + s.nrm[1] * by + s.nrm[2] * bz;
          const double ncc = s.nrm[0] * cx + s.nrm[1] * cy;
          regular_harmonics<true>(P, sp, -1.0, [&](int n, int m, double yr, double yi, double tr, double ti) {
            // brh = n/rho Y, bal = Ytheta, bbe = -i m Y = (m yi, -m yr)
            const double fr = n * ir;
            const double vr = na * fr * yr + nb_ * tr + ncc * (m * yi);
            const double vi = na * fr * yi + nb_ * ti - ncc * (m * yr);
            row[n * n + n + m] = mult * vr;
            if (m > 0) row[n * n + n - m] = mult * vi;
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
  double* Mb = M + (size_t)b * xstride(P);
#pragma unroll
  for (int i2 = 0; i2 < (FMMB_MAX_P * FMMB_MAX_P + 31) / 32; ++i2) {
    const int col = lane + 32 * i2;
    if (col < pp) Mb[col] = acc[i2];
  }
}

// L2P: warp per leaf, lane per target panel; only panels whose BC selects this set are touched
template <int SET>
__global__ void __launch_bounds__(128)
bem_l2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
               const unsigned* __restrict__ be, const double4* __restrict__ center,
               const unsigned char* __restrict__ has_local, const bem::Panel* __restrict__ pan,
               const int* __restrict__ bc, int P, const double* __restrict__ L, double* __restrict__ res) {
  extern __shared__ double2 bem_ls[];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  const int b = leaves[w];
  if (!has_local[b]) return;
  double2* Ls = bem_ls + wl * nc;
  for (int i = lane; i < nc; i += 32) {
    int n, m;
    unpack_nm(i, n, m);
    Ls[i] = load_coef(L + (size_t)b * xstride(P), n, m);
  }
  __syncwarp();
  const double4 c = center[b];
  for (unsigned i = bb[b] + lane; i < be[b]; i += 32) {
    if (bc[i] != SET) continue;
    const bem::Panel& t = pan[i];
    const Sph s = to_sph(t.c[0] - c.x, t.c[1] - c.y, t.c[2] - c.z);
    double acc = 0;
    regular_harmonics<false>(P, s, 1.0, [&](int n, int m, double yr, double yi, double, double) {
      const double2 l = Ls[n * (n + 1) / 2 + m];
      acc += (m == 0 ? 1.0 : 2.0) * (l.x * yr - l.y * yi);
    });
    res[i] = SET == 0 ? acc : -acc;
  }
}

// M2P (treecode, `LaplaceBEM -eval TREE`): warp per leaf, lane per target panel; every source box accepted for the
// leaf or one of its ancestors is evaluated at the panel centres (vector M2P, LaplaceSphericalBEM.hpp:394-422):
// only panels whose BC selects this set are touched, set 0 adds, set 1 subtracts.  Accumulates into res (zeroed by
// the caller), sources in list order per box, ancestors bottom-up -- a fixed order, no atomics.
template <int SET>
__global__ void __launch_bounds__(128)
bem_m2p_kernel(const int* __restrict__ leaves, int nleaves, const unsigned* __restrict__ bb,
               const unsigned* __restrict__ be, const unsigned* __restrict__ parent, const int* __restrict__ off,
               const int* __restrict__ src, const double4* __restrict__ center, const bem::Panel* __restrict__ pan,
               const int* __restrict__ bc, int P, const double* __restrict__ M, double* __restrict__ res) {
  extern __shared__ double2 bem_ms[];
  const int nc = P * (P + 1) / 2;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + wl;
  if (w >= nleaves) return;
  double2* Ms = bem_ms + wl * nc;
  const int leaf = leaves[w];
  const unsigned b0 = bb[leaf], b1 = be[leaf];
  for (unsigned base = b0; base < b1; base += 32) {
    const unsigned i = base + lane;
    const bool act = i < b1 && bc[i] == SET;
    double px = 0, py = 0, pz = 0;
    if (act) { px = pan[i].c[0]; py = pan[i].c[1]; pz = pan[i].c[2]; }
    double acc = 0;
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
          const Sph s = to_sph(px - c.x, py - c.y, pz - c.z);
          double v = 0;
          regular_harmonics<false, true>(P, s, 1.0, [&](int n, int m, double yr, double yi, double, double) {
            const double2 c2 = Ms[n * (n + 1) / 2 + m];
            v += (m == 0 ? 1.0 : 2.0) * (c2.x * yr - c2.y * yi);   // Re(M Y)
          });
          acc += v;
        }
      }
      if (a == 0) break;
    }
    if (act) res[i] += SET == 0 ? acc : -acc;
  }
}

// Direct::matvec with the panel kernel (include/Direct.hpp:99-125 over operator(), :273-297): block per target panel,
// threads stride over ALL source panels of the plan, block-wide sum in a fixed tree order.
__global__ void __launch_bounds__(128)
bem_direct_kernel(const bem::Panel* __restrict__ pan, const double* __restrict__ chg, int64_t ns,
                  const double* __restrict__ tverts, const int* __restrict__ tbc, double kappa,
                  double* __restrict__ out) {
  __shared__ double part[128];
  const int64_t t = blockIdx.x;
  const double* v = tverts + 9 * (size_t)t;
  const double tc[3] = {((v[0] + v[3]) + v[6]) / 3, ((v[1] + v[4]) + v[7]) / 3, ((v[2] + v[5]) + v[8]) / 3};
  const int bc = tbc ? tbc[t] : 0;
  double acc = 0;
  for (int64_t j = threadIdx.x; j < ns; j += blockDim.x)
    acc += (kappa < 0 ? bem::kernel(bc, tc, pan[j], c_rule, c_fine) : bem::kernel_yk(bc, tc, pan[j], c_rule, kappa)) * chg[j];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[t] = part[0];
}

__global__ void bem_gather_plain(const double* __restrict__ q, const unsigned* __restrict__ perm, int64_t n,
                                 double* __restrict__ chg) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) chg[i] = q[perm[i]];
}

__global__ void bem_gather_charges(const double* __restrict__ q, const unsigned* __restrict__ perm, int64_t n,
                                   double4* __restrict__ body) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) body[i].w = q[perm[i]];
}

}  // namespace

// Plan-time: panel geometry in tree order, then the cached near field.
void bem_setup(fmmb_plan* plan, const double* verts_host, const int32_t* bc_host, int quad_k, double kappa) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  if (!bem::rule_supported(quad_k))
    throw StatusError{FMMB_ERR_UNSUPPORTED, "quad_k must be a key of the reference's Gauss table: 1, 3, 4, 7, 13, 17, 19, 25 or 79"};
  BemData* B = new BemData();
  plan->bem = B;
  B->K = quad_k == 7 ? 4 : quad_k;
  B->kappa = kappa;
  upload_laplace_tables();   // this translation unit's copy of the factorial tables
  bem::Rule rule = bem::make_rule(B->K), fine = bem::make_rule(17);
  FMMB_CUDA(cudaMemcpyToSymbol(c_rule, &rule, sizeof rule));
  FMMB_CUDA(cudaMemcpyToSymbol(c_fine, &fine, sizeof fine));
  const int64_t n = T.n;
  DevBuf<double> verts;
  DevBuf<int> bc;
  verts.from_host(verts_host, 9 * (size_t)n, s);
  for (int64_t i = 0; i < n; ++i) {
    int v = bc_host ? bc_host[i] : 0;
    if (v != 0 && v != 1) throw StatusError{FMMB_ERR_INVALID, "bc entries must be 0 (POTENTIAL) or 1 (NORMAL_DERIV)"};
    B->set_active[v] = true;
  }
  if (bc_host) bc.from_host(bc_host, n, s);
  B->pan.resize(n); B->bc.resize(n);
  bem_setup_kernel<<<nblk(n, 128), 128, 0, s>>>(verts.p, bc_host ? bc.p : nullptr, T.perm.p, n, B->pan.p, B->bc.p);
  FMMB_CUDA(cudaGetLastError());
  // block offsets of the cached near field
  const int ni = T.n_p2p_items;
  DevBuf<long long> cnt;
  cnt.resize(ni + 1);
  bem_count_kernel<<<nblk(ni + 1, 128), 128, 0, s>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p, T.p2p_off.p,
                                                    T.p2p_src.p, cnt.p);
  FMMB_CUDA(cudaGetLastError());
  std::vector<long long> h = cnt.to_host(s), off(ni + 1, 0);
  for (int i = 0; i < ni; ++i) off[i + 1] = off[i] + h[i];
  B->nnz = off[ni];
  B->nf_base.from_host(off.data(), off.size(), s);
  B->nf_val.resize((size_t)B->nnz);
  if (ni)
    bem_assemble_kernel<<<nblk(ni, kBemWarps), 32 * kBemWarps, 0, s>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p,
                                                                      T.p2p_off.p, T.p2p_src.p, B->pan.p, B->bc.p,
                                                                      B->nf_base.p, B->kappa, B->nf_val.p);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));
}

static void launch_bem_near(fmmb_plan* plan, cudaStream_t s) {
  Tree& T = plan->tree;
  BemData* B = plan->bem;
  const int ni = T.n_p2p_items;
  if (plan->bem_near_kernel)
    bem_near_split_kernel<<<ni, 32 * kNearSplit, 0, s>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p, T.p2p_off.p,
                                                        T.p2p_src.p, T.body.p, B->nf_base.p, B->nf_val.p, B->res_near.p);
  else
    bem_near_kernel<<<nblk(ni, kBemWarps), 32 * kBemWarps, 0, s>>>(T.p2p_items.p, ni, T.bbegin.p, T.bend.p, T.p2p_off.p,
                                                                  T.p2p_src.p, T.body.p, B->nf_base.p, B->nf_val.p,
                                                                  B->res_near.p);
}

// pieces of the BEM matvec shared with YukawaCartesianBEM (csrc/yukawa.cu): charges into tree order + cached near
// field, and access to the panel data
void bem_begin(fmmb_plan* plan, const double* d_charges, cudaStream_t s) {
  Tree& T = plan->tree;
  BemData* B = plan->bem;
  const int64_t n = T.n;
  B->res_near.resize(n); B->res_far.resize(n);
  bem_gather_charges<<<nblk(n, 256), 256, 0, s>>>(exec_charges(plan, d_charges), exec_perm(plan), n, T.body.p);
  const int ni = T.n_p2p_items;
  if (ni)
    launch_bem_near(plan, s);
  B->res_far.zero(s);
  plan->launches += 3;
  FMMB_CUDA(cudaGetLastError());
}
const bem::Panel* bem_panels(const BemData* b) { return b->pan.p; }
const int* bem_bc(const BemData* b) { return b->bc.p; }
bool bem_set_active(const BemData* b, int set) { return b->set_active[set]; }
double* bem_res_near(BemData* b) { return b->res_near.p; }
double* bem_res_far(BemData* b) { return b->res_far.p; }
int bem_rule_points(const BemData* b) { return b->K; }

// fmmb_plan_direct_panels for LaplaceSphericalBEM / YukawaCartesianBEM plans (device pointers; charges original order)
void bem_direct(fmmb_plan* plan, const double* d_charges, int64_t nt, const double* d_tverts, const int* d_tbc,
                double* d_out, cudaStream_t s) {
  Tree& T = plan->tree;
  BemData* B = plan->bem;
  DevBuf<double> chg;
  chg.resize(T.n);
  bem_gather_plain<<<nblk(T.n, 256), 256, 0, s>>>(d_charges, T.perm.p, T.n, chg.p);
  if (nt) bem_direct_kernel<<<(unsigned)nt, 128, 0, s>>>(B->pan.p, chg.p, T.n, d_tverts, d_tbc, B->kappa, d_out);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));      // chg is released on return
}

void bem_execute(fmmb_plan* plan, const double* d_charges, double* d_results) {
  Tree& T = plan->tree;
  BemData* B = plan->bem;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P;
  const int64_t n = T.n;
  cudaStream_t s = plan->stream;
  cudaEvent_t* ev = plan->ev;
  laplace_prepare_expansions(plan);
  B->res_near.resize(n); B->res_far.resize(n);
  plan->launches = 0;
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[0], s));
  bem_gather_charges<<<nblk(n, 256), 256, 0, s>>>(exec_charges(plan, d_charges), exec_perm(plan), n, T.body.p);
  FMMB_CUDA(cudaEventRecord(ev[1], s));
  // cached near field on the second stream, beside the far-field chain of short dependent kernels (round 2)
  cudaStream_t s2 = plan->overlap_p2p ? plan->stream2 : s;
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s2, ev[1], 0));
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[6], s2));
  const int ni = T.n_p2p_items;
  if (ni)
    launch_bem_near(plan, s2);
  FMMB_CUDA(cudaEventRecord(ev[7], s2));
  B->res_far.zero(s);
  plan->launches += 3;
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[12], s));
  const int warps = pp <= 64 ? 4 : 1;
  const size_t sh = (size_t)warps * 32 * (pp | 1) * sizeof(double);
  // per call: function attributes belong to the current device
  FMMB_CUDA(cudaFuncSetAttribute(bem_p2m_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  FMMB_CUDA(cudaFuncSetAttribute(bem_p2m_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  for (int set = 0; set < 2; ++set) {
    if (!B->set_active[set] || plan->near_only) continue;   // near_only: plans for preconditioners, no far field
    if (set == 0)
      bem_p2m_kernel<0><<<nblk(T.nleaves, warps), 32 * warps, sh, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p,
                                                                      T.center.p, T.body.p, B->pan.p, B->bc.p, P,
                                                                      plan->M.p);
    else
      bem_p2m_kernel<1><<<nblk(T.nleaves, warps), 32 * warps, sh, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p,
                                                                      T.center.p, T.body.p, B->pan.p, B->bc.p, P,
                                                                      plan->M.p);
    ++plan->launches;
    laplace_translations(plan, s);          // treecode: stops after the upward pass
    if (plan->opts.evaluator == FMMB_EVAL_TREECODE) {
      if (!T.n_own_leaves) continue;
      if (set == 0)
        bem_m2p_kernel<0><<<nblk(T.n_own_leaves, 4), 128, 4 * nc * sizeof(double2), s>>>(
            T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.parent.p, T.m2l_off.p, T.m2l_src.p, T.center.p,
            B->pan.p, B->bc.p, P, plan->M.p, B->res_far.p);
      else
        bem_m2p_kernel<1><<<nblk(T.n_own_leaves, 4), 128, 4 * nc * sizeof(double2), s>>>(
            T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.parent.p, T.m2l_off.p, T.m2l_src.p, T.center.p,
            B->pan.p, B->bc.p, P, plan->M.p, B->res_far.p);
      ++plan->launches;
      continue;
    }
    if (!T.n_own_leaves) continue;         // a rank may own no leaf at all (more ranks than leaves)
    if (set == 0)
      bem_l2p_kernel<0><<<nblk(T.n_own_leaves, 4), 128, 4 * nc * sizeof(double2), s>>>(
          T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.center.p, T.has_local.p, B->pan.p, B->bc.p, P,
          plan->L.p, B->res_far.p);
    else
      bem_l2p_kernel<1><<<nblk(T.n_own_leaves, 4), 128, 4 * nc * sizeof(double2), s>>>(
          T.own_leaves.p, T.n_own_leaves, T.bbegin.p, T.bend.p, T.center.p, T.has_local.p, B->pan.p, B->bc.p, P,
          plan->L.p, B->res_far.p);
    ++plan->launches;
  }
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[4], s));
  if (s2 != s) FMMB_CUDA(cudaStreamWaitEvent(s, ev[7], 0));
  finish_results(plan, B->res_near.p, B->res_far.p, 1, d_results, s);
  if (!plan->capturing) FMMB_CUDA(cudaEventRecord(ev[5], s));
  FMMB_CUDA(cudaGetLastError());
  plan->timed = true;
}

}  // namespace fmmb
