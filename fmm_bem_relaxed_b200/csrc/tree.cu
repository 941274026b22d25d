// csrc/tree.cu -- device octree and interaction lists, bit-exact to the reference.
//
// Replaces (reference paths):
//   include/tree/Octree.hpp:67-79      get_boundingbox        -> bbox_partial + host finish
//   include/tree/Octree.hpp:118-129    MortonCoder::code       -> morton_codes
//   include/tree/Octree.hpp:617-692    construct_tree          -> radix sort + per-level split kernels
//   include/tree/Octree.hpp:226-248,334-355  Box geometry     -> box_geometry
//   include/FMMOptions.hpp:21-31       DefaultMAC              -> mac_accept (no FMA contraction)
//   include/executor/EvalInteractionLazy.hpp:59-105,224-237  dual traversal -> frontier kernels
//
// The reference sorts bodies with a recursive STABLE 8-way bucket sort, so bodies inside a leaf
// keep their input order.  Here: (1) stable radix sort by the full 30-bit Morton code gives the
// box structure; (2) a second stable radix sort by the code masked to each body's leaf level
// gives exactly the reference permutation.
//
// The reference's FIFO traversal processes the pair queue generation by generation; expanding
// a whole generation in parallel with prefix-sum output offsets reproduces the sequential
// append order of LR_list and P2P_lists exactly.
#include "common.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace fmmb {
namespace {

__host__ __device__ __forceinline__ unsigned spread10(unsigned x) {
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
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

// ---- bounding box -------------------------------------------------------------------------
__global__ void bbox_partial(const double* __restrict__ pts, int64_t n, double* __restrict__ out) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double v = pts[3 * i + k];
      mn[k] = fmin(mn[k], v);
      mx[k] = fmax(mx[k], v);
    }
  }
  __shared__ double sh[6][256];
  for (int k = 0; k < 3; ++k) { sh[k][threadIdx.x] = mn[k]; sh[3 + k][threadIdx.x] = mx[k]; }
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) {
        sh[k][threadIdx.x] = fmin(sh[k][threadIdx.x], sh[k][threadIdx.x + s]);
        sh[3 + k][threadIdx.x] = fmax(sh[3 + k][threadIdx.x], sh[3 + k][threadIdx.x + s]);
      }
    __syncthreads();
  }
  if (threadIdx.x < 6) out[blockIdx.x * 6 + threadIdx.x] = sh[threadIdx.x][0];
}

// ---- Morton codes: subtraction, division, truncation, in that order (Octree.hpp:118-129) ----
__global__ void morton_codes(const double* __restrict__ pts, int64_t n, double3 pmin, double3 cell,
                             unsigned* __restrict__ code, unsigned* __restrict__ idx, int* err) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double vx = __ddiv_rn(__dsub_rn(pts[3 * i], pmin.x), cell.x);
  double vy = __ddiv_rn(__dsub_rn(pts[3 * i + 1], pmin.y), cell.y);
  double vz = __ddiv_rn(__dsub_rn(pts[3 * i + 2], pmin.z), cell.z);
  // all bodies coincide (zero-size cube): the reference would assert; put them in cell 0 instead
  if (cell.x == 0.0 && cell.y == 0.0 && cell.z == 0.0) vx = vy = vz = 0.0;
  unsigned qx = (unsigned)vx, qy = (unsigned)vy, qz = (unsigned)vz;
  if (!(vx >= 0.0 && vy >= 0.0 && vz >= 0.0) || qx >= 1024u || qy >= 1024u || qz >= 1024u) {
    atomicOr(err, 1);
    qx &= 1023u; qy &= 1023u; qz &= 1023u;
  }
  code[i] = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
  idx[i] = (unsigned)i;
}

__device__ __forceinline__ unsigned lower_bound_u32(const unsigned* a, unsigned lo, unsigned hi, unsigned v) {
  while (lo < hi) {
    unsigned mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- one BFS level of construct_tree (Octree.hpp:636-682): 8 lanes per parent box ------------
__global__ void split_count(const unsigned* __restrict__ sc, const unsigned* __restrict__ bb,
                            const unsigned* __restrict__ be, int lo, int nparents, int level,
                            unsigned ncrit, unsigned* __restrict__ tmp_lo, unsigned* __restrict__ tmp_hi,
                            unsigned* __restrict__ child_mask, int* __restrict__ nchild, int* err) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int j = t >> 3, c = t & 7;
  bool active = j < nparents;
  unsigned l = 0, h = 0;
  if (active) {
    int b = lo + j;
    unsigned b0 = bb[b], b1 = be[b];
    if (b1 - b0 > ncrit) {
      if (level >= 10) {
        if (c == 0) atomicOr(err, 2);
      } else {
        unsigned shift = 3u * (10 - level - 1);
        unsigned base = sc[b0] & ~((1u << (shift + 3)) - 1u);
        l = lower_bound_u32(sc, b0, b1, base | ((unsigned)c << shift));
        h = (c == 7) ? b1 : lower_bound_u32(sc, b0, b1, base | ((unsigned)(c + 1) << shift));
      }
    }
  }
  unsigned full = __ballot_sync(0xffffffffu, h > l);
  if (active) {
    unsigned m = (full >> ((threadIdx.x & 31) & ~7)) & 0xffu;
    tmp_lo[t] = l; tmp_hi[t] = h;
    if (c == 0) { child_mask[j] = m; nchild[j] = __popc(m); }
  }
}

__global__ void split_write(int lo, int nparents, int level, int next_off,
                            const unsigned* __restrict__ tmp_lo, const unsigned* __restrict__ tmp_hi,
                            const unsigned* __restrict__ child_mask, const int* __restrict__ child_off,
                            unsigned* __restrict__ key, unsigned* __restrict__ parent,
                            unsigned* __restrict__ cbegin, unsigned* __restrict__ cend,
                            unsigned* __restrict__ bb, unsigned* __restrict__ be,
                            unsigned* __restrict__ lvl) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int j = t >> 3, c = t & 7;
  if (j >= nparents) return;
  int b = lo + j;
  unsigned m = child_mask[j];
  int first = next_off + child_off[j];
  if (c == 0) {
    if (m == 0) {          // leaf: keep body offsets, set the leaf bit (Octree.hpp:641-644)
      key[b] |= 0x80000000u;
      cbegin[b] = bb[b]; cend[b] = be[b];
    } else {
      cbegin[b] = first; cend[b] = first + __popc(m);
    }
  }
  if (m & (1u << c)) {
    int r = first + __popc(m & ((1u << c) - 1u));
    key[r] = (key[b] & 0x7fffffffu) << 3 | (unsigned)c;
    parent[r] = b;
    bb[r] = tmp_lo[t]; be[r] = tmp_hi[t];
    cbegin[r] = 0; cend[r] = 0;
    lvl[r] = level + 1;
  }
}

__global__ void flag_leaves(const unsigned* __restrict__ key, int nb, int* __restrict__ flag) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb) flag[b] = (key[b] >> 31) & 1;
  if (b == nb) flag[b] = 0;
}
__global__ void compact_leaves(const int* __restrict__ flag, const int* __restrict__ pos, int nb,
                               int* __restrict__ leaves) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb && flag[b]) leaves[pos[b]] = b;
}

// masked code per ORIGINAL body: code with the digits below the body's leaf level cleared
__global__ void leaf_masked_codes(const int* __restrict__ leaves, int nleaves,
                                  const unsigned* __restrict__ bb, const unsigned* __restrict__ be,
                                  const unsigned* __restrict__ lvl, const unsigned* __restrict__ sc,
                                  const unsigned* __restrict__ sidx, unsigned* __restrict__ masked_orig,
                                  unsigned* __restrict__ iota) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nleaves) return;
  int b = leaves[w];
  unsigned shift = 3u * (10 - lvl[b]);
  unsigned mask = shift >= 32 ? 0u : ~((1u << shift) - 1u);
  for (unsigned i = bb[b] + lane; i < be[b]; i += 32) {
    unsigned o = sidx[i];
    masked_orig[o] = sc[i] & mask;
    iota[o] = o;
  }
}

__global__ void gather_bodies(const double* __restrict__ pts, const unsigned* __restrict__ perm,
                              const unsigned* __restrict__ code_orig, int64_t n,
                              double4* __restrict__ body, unsigned* __restrict__ code_tree) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned o = perm[i];
  body[i] = make_double4(pts[3 * (size_t)o], pts[3 * (size_t)o + 1], pts[3 * (size_t)o + 2], 0.0);
  code_tree[i] = code_orig[o];
}

// ---- box geometry in the reference's operation order (Octree.hpp:243-248,334-355) -------------
__global__ void box_geometry(const unsigned* __restrict__ key, const unsigned* __restrict__ lvl, int nb,
                             double3 pmin, double3 cell, double4* __restrict__ center) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  unsigned L = lvl[b];
  unsigned m = (key[b] & 0x7fffffffu) << (3u * (10 - L));   // marker now at bit 30
  m &= ~(1u << 30);
  double ix = (double)compact10(m), iy = (double)compact10(m >> 1), iz = (double)compact10(m >> 2);
  // level-10 leaves: the reference shifts by -1 (undefined); use the true half cell there
  double mult = L >= 10 ? 0.5 : (double)(1 << (9 - (int)L));
  double cminx = __dadd_rn(pmin.x, __dmul_rn(cell.x, ix));
  double cminy = __dadd_rn(pmin.y, __dmul_rn(cell.y, iy));
  double cminz = __dadd_rn(pmin.z, __dmul_rn(cell.z, iz));
  double cx = __dadd_rn(cminx, __dmul_rn(__dsub_rn(__dadd_rn(cminx, cell.x), cminx), mult));
  double cy = __dadd_rn(cminy, __dmul_rn(__dsub_rn(__dadd_rn(cminy, cell.y), cminy), mult));
  double cz = __dadd_rn(cminz, __dmul_rn(__dsub_rn(__dadd_rn(cminz, cell.z), cminz), mult));
  double ext = __dsub_rn(__dadd_rn(pmin.x, __dmul_rn(1024.0, cell.x)), pmin.x);
  double side = __ddiv_rn(ext, (double)(1 << L));
  center[b] = make_double4(cx, cy, cz, side);
}

// accept iff |c1-c2|^2 > ((r1+r2)/theta)^2, r = side/2  (FMMOptions.hpp:26-30), no contraction
__device__ __forceinline__ bool mac_accept(const double4 a, const double4 b, double theta) {
  double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
  double r0 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  double rhs = __ddiv_rn(__dadd_rn(__ddiv_rn(a.w, 2.0), __ddiv_rn(b.w, 2.0)), theta);
  return r0 > __dmul_rn(rhs, rhs);
}

// ---- dual tree traversal, one generation of the reference's FIFO queue ------------------------
struct TreeView {
  const unsigned* key; const unsigned* cbegin; const unsigned* cend; const double4* center;
  double theta;
};
// which box is split: 0 = leaf x leaf (P2P), 1 = children of b1 against b2, 2 = children of b2 against b1
__device__ __forceinline__ int split_side(const TreeView& T, int b1, int b2) {
  bool l1 = T.key[b1] >> 31, l2 = T.key[b2] >> 31;
  if (l1 && l2) return 0;
  if (l1) return 2;
  if (l2) return 1;
  return T.center[b1].w > T.center[b2].w ? 1 : 2;
}

__global__ void traverse_count(TreeView T, const int2* __restrict__ front, int nfront,
                               unsigned long long* __restrict__ cnt, int* __restrict__ cnt_p2p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nfront) return;
  if (i == nfront) { cnt[i] = 0; cnt_p2p[i] = 0; return; }
  int2 pr = front[i];
  int side = split_side(T, pr.x, pr.y);
  if (side == 0) { cnt[i] = 0; cnt_p2p[i] = 1; return; }
  unsigned nlr = 0, nq = 0;
  if (side == 1) {
    double4 cb = T.center[pr.y];
    for (unsigned c = T.cbegin[pr.x]; c < T.cend[pr.x]; ++c)
      if (mac_accept(T.center[c], cb, T.theta)) ++nlr; else ++nq;
  } else {
    double4 ca = T.center[pr.x];
    for (unsigned c = T.cbegin[pr.y]; c < T.cend[pr.y]; ++c)
      if (mac_accept(ca, T.center[c], T.theta)) ++nlr; else ++nq;
  }
  cnt[i] = (unsigned long long)nlr | ((unsigned long long)nq << 32);
  cnt_p2p[i] = 0;
}

__global__ void traverse_write(TreeView T, const int2* __restrict__ front, int nfront,
                               const unsigned long long* __restrict__ off, const int* __restrict__ off_p2p,
                               int2* __restrict__ lr, int2* __restrict__ next, int2* __restrict__ p2p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfront) return;
  int2 pr = front[i];
  int side = split_side(T, pr.x, pr.y);
  if (side == 0) { p2p[off_p2p[i]] = pr; return; }
  unsigned long long o = off[i];
  unsigned olr = (unsigned)(o & 0xffffffffu), oq = (unsigned)(o >> 32);
  if (side == 1) {
    double4 cb = T.center[pr.y];
    for (unsigned c = T.cbegin[pr.x]; c < T.cend[pr.x]; ++c) {
      int2 np = make_int2((int)c, pr.y);
      if (mac_accept(T.center[c], cb, T.theta)) lr[olr++] = np; else next[oq++] = np;
    }
  } else {
    double4 ca = T.center[pr.x];
    for (unsigned c = T.cbegin[pr.y]; c < T.cend[pr.y]; ++c) {
      int2 np = make_int2(pr.x, (int)c);
      if (mac_accept(ca, T.center[c], T.theta)) lr[olr++] = np; else next[oq++] = np;
    }
  }
}

// split (a,b) pair list into key / value arrays for a stable sort by `key`
__global__ void unzip_pairs(const int2* __restrict__ pr, int64_t n, int key_is_y,
                            unsigned* __restrict__ k, int* __restrict__ v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int2 p = pr[i];
  k[i] = key_is_y ? (unsigned)p.y : (unsigned)p.x;
  v[i] = key_is_y ? p.x : p.y;
}
__global__ void csr_offsets(const unsigned* __restrict__ sorted_keys, int64_t n, int nb, int* __restrict__ off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < (unsigned)b) lo = mid + 1; else hi = mid;
  }
  off[b] = (int)lo;
}
__global__ void mark_targets(const int2* __restrict__ lr, int64_t n, unsigned char* __restrict__ has) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) has[lr[i].y] = 1;
}
__global__ void mark_sources(const int2* __restrict__ lr, int64_t n, unsigned char* __restrict__ need) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) need[lr[i].x] = 1;
}
__global__ void inherit_local(const unsigned* __restrict__ parent, int lo, int hi, unsigned char* __restrict__ has) {
  int b = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (b < hi && has[parent[b]]) has[b] = 1;
}
__global__ void count_body_pairs(const int2* __restrict__ p2p, int64_t n, const unsigned* __restrict__ bb,
                                 const unsigned* __restrict__ be, unsigned long long* total) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (i < n) {
    int2 p = p2p[i];
    v = (unsigned long long)(be[p.x] - bb[p.x]) * (unsigned long long)(be[p.y] - bb[p.y]);
  }
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, v);
}

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

void sort_pairs_u32(Temp& tmp, const unsigned* kin, unsigned* kout, const unsigned* vin, unsigned* vout,
                    int64_t n, int end_bit, cudaStream_t s) {
  size_t bytes = 0;
  FMMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
  void* t = tmp.get(bytes);
  FMMB_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
}
void sort_pairs_i32(Temp& tmp, const unsigned* kin, unsigned* kout, const int* vin, int* vout,
                    int64_t n, int end_bit, cudaStream_t s) {
  size_t bytes = 0;
  FMMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
  void* t = tmp.get(bytes);
  FMMB_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
}
template <typename T>
void exclusive_sum(Temp& tmp, const T* in, T* out, int64_t n, cudaStream_t s) {
  size_t bytes = 0;
  FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  void* t = tmp.get(bytes);
  FMMB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, n, s));
}
int bits_for(int64_t v) { int b = 1; while ((1ll << b) <= v) ++b; return b; }

// pair list -> CSR by one side, preserving list order inside each row
void pairs_to_csr(Temp& tmp, const DevBuf<int2>& pairs, int64_t n, int key_is_y, int nb,
                  DevBuf<int>& off, DevBuf<int>& val, cudaStream_t s) {
  DevBuf<unsigned> k0, k1;
  DevBuf<int> v0;
  k0.resize(n); k1.resize(n); v0.resize(n);
  val.resize(n);
  off.resize(nb + 1);
  if (n) {
    unzip_pairs<<<nblk(n, 256), 256, 0, s>>>(pairs.p, n, key_is_y, k0.p, v0.p);
    sort_pairs_i32(tmp, k0.p, k1.p, v0.p, val.p, n, bits_for(nb), s);
  }
  csr_offsets<<<nblk(nb + 1, 256), 256, 0, s>>>(k1.p, n, nb, off.p);
  FMMB_CUDA(cudaGetLastError());
  FMMB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace

void partition_ranges(const double* w, int64_t n, int nranks, int64_t* cuts) {
  double total = 0;
  for (int64_t i = 0; i < n; ++i) total += w[i];
  cuts[0] = 0;
  double run = 0;
  int64_t i = 0;
  for (int r = 1; r < nranks; ++r) {
    double goal = total * r / nranks;
    while (i < n && run + 0.5 * w[i] < goal) run += w[i++];
    cuts[r] = i;
  }
  cuts[nranks] = n;
  for (int r = 1; r <= nranks; ++r) cuts[r] = std::max(cuts[r], cuts[r - 1]);
}

// Multi-GPU: choose this rank's leaf range and restrict the target-major M2L lists to it.
static void partition_tree(fmmb_plan* plan) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  const int nb = T.nboxes;
  T.rank = plan->opts.nranks > 1 ? plan->opts.rank : 0;
  T.nranks = plan->opts.nranks > 1 ? plan->opts.nranks : 1;
  T.n_lr_local = T.n_lr;
  T.body_cuts.assign(T.nranks + 1, 0);
  T.body_cuts[T.nranks] = T.n;
  T.own_b0 = 0; T.own_b1 = T.n;
  T.active.resize(nb);
  if (T.nranks == 1) {
    T.n_own_leaves = T.nleaves;
    T.own_leaves.resize(T.nleaves);
    FMMB_CUDA(cudaMemcpyAsync(T.own_leaves.p, T.leaves.p, T.nleaves * sizeof(int), cudaMemcpyDeviceToDevice, s));
    FMMB_CUDA(cudaMemsetAsync(T.active.p, 1, nb, s));
    return;
  }
  if (T.rank < 0 || T.rank >= T.nranks) throw StatusError{FMMB_ERR_INVALID, "rank must be in [0, nranks)"};
  std::vector<int> leaves = T.leaves.to_host(s), m2l_off = T.m2l_off.to_host(s), m2l_src = T.m2l_src.to_host(s),
                   p2p_off = T.p2p_off.to_host(s), p2p_src = T.p2p_src.to_host(s);
  std::vector<unsigned> bb = T.bbegin.to_host(s), be = T.bend.to_host(s), par = T.parent.to_host(s),
                        hcb = T.cbegin.to_host(s), hce = T.cend.to_host(s);
  // work estimate per box: M2L pairs into it, inherited from the ancestors in proportion to the bodies
  // (one M2L pair costs about as much as 230 P2P body pairs at P = 8 on a B200)
  std::vector<double> far(nb, 0.0);
  for (int b = 0; b < nb; ++b) {
    double own = m2l_off[b + 1] - m2l_off[b];
    double inh = b == 0 ? 0.0 : far[par[b]] * double(be[b] - bb[b]) / double(be[par[b]] - bb[par[b]]);
    far[b] = own + inh;   // parents precede children in BFS order
  }
  std::vector<int> order(leaves);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return bb[a] < bb[b]; });
  std::vector<double> w(order.size());
  for (size_t i = 0; i < order.size(); ++i) {
    int b = order[i];
    double near = 0;
    for (int e = p2p_off[b]; e < p2p_off[b + 1]; ++e) near += double(be[p2p_src[e]] - bb[p2p_src[e]]);
    w[i] = near * double(be[b] - bb[b]) + 230.0 * far[b];
  }
  std::vector<int64_t> cuts(T.nranks + 1);
  partition_ranges(w.data(), (int64_t)w.size(), T.nranks, cuts.data());
  for (int r = 0; r <= T.nranks; ++r)
    T.body_cuts[r] = cuts[r] >= (int64_t)order.size() ? T.n : (int64_t)bb[order[cuts[r]]];
  T.own_b0 = T.body_cuts[T.rank];
  T.own_b1 = T.body_cuts[T.rank + 1];
  std::vector<int> own;
  for (int b : leaves) if (bb[b] >= T.own_b0 && bb[b] < T.own_b1) own.push_back(b);
  T.n_own_leaves = (int)own.size();
  T.own_leaves.from_host(own.data(), own.size(), s);
  std::vector<unsigned char> act(nb);
  std::vector<int> noff(nb + 1, 0), nsrc;
  for (int b = 0; b < nb; ++b) {
    act[b] = (int64_t)bb[b] < T.own_b1 && (int64_t)be[b] > T.own_b0;
    if (act[b]) nsrc.insert(nsrc.end(), m2l_src.begin() + m2l_off[b], m2l_src.begin() + m2l_off[b + 1]);
    noff[b + 1] = (int)nsrc.size();
  }
  // upward-pass ownership
  {
    std::vector<unsigned char> inside_me(nb, 0);
    std::vector<int> owner(nb, -1);        // rank whose range contains the whole box, -1 = straddles a cut
    for (int b = 0; b < nb; ++b) {
      int q = (int)(std::upper_bound(T.body_cuts.begin(), T.body_cuts.end(), (int64_t)bb[b]) - T.body_cuts.begin()) - 1;
      if (q >= 0 && q < T.nranks && (int64_t)be[b] <= T.body_cuts[q + 1]) owner[b] = q;
      inside_me[b] = owner[b] == T.rank;
    }
    T.up_inside.from_host(inside_me.data(), inside_me.size(), s);
    std::vector<unsigned> key = T.key.to_host(s);
    // the gate of the owned upward pass must come out the same on every rank (a collective follows it): it is
    // computed from the owners of ALL boxes, not from what this rank happens to hold
    T.box_owner = owner;
    T.owned_upward = false;
    for (int b = 0; b < nb; ++b) if (owner[b] >= 0 && !(key[b] >> 31)) { T.owned_upward = true; break; }
    // straddlers and their maximal single-rank descendants (children of straddlers that are not straddlers)
    std::vector<int> sb, soff(1, 0), sdesc, spair;
    for (int b = 0; b < nb; ++b) {
      if (owner[b] >= 0 || (key[b] >> 31) || !T.need_M_host[b]) continue;
      // descendants of b form contiguous index ranges per level; walk down through straddling children only
      std::vector<int> stack(1, b);
      while (!stack.empty()) {
        int x = stack.back();
        stack.pop_back();
        std::vector<unsigned> cb1 = {}, ce1 = {};
        (void)cb1; (void)ce1;
        for (unsigned c = hcb[x]; c < hce[x]; ++c) {
          if (owner[c] >= 0) { sdesc.push_back((int)c); spair.push_back((int)sb.size()); }
          else stack.push_back((int)c);
        }
      }
      sb.push_back(b);
      soff.push_back((int)sdesc.size());
    }
    T.n_strad = (int)sb.size();
    T.n_strad_pairs = (int)sdesc.size();
    T.strad_box.from_host(sb.data(), sb.size(), s);
    T.strad_off.from_host(soff.data(), soff.size(), s);
    T.strad_desc.from_host(sdesc.data(), sdesc.size(), s);
    T.strad_pair_box.from_host(spair.data(), spair.size(), s);
    std::vector<int> xl;
    T.xchg_off.assign(T.nranks + 1, 0);
    T.xchg_max = 0;
    for (int q = 0; q < T.nranks; ++q) {
      for (int b = 0; b < nb; ++b) if (owner[b] == q) xl.push_back(b);
      T.xchg_off[q + 1] = (int)xl.size();
      T.xchg_max = std::max(T.xchg_max, T.xchg_off[q + 1] - T.xchg_off[q]);
    }
    T.xchg_list.from_host(xl.data(), xl.size(), s);
  }
  T.active.from_host(act.data(), act.size(), s);
  T.m2l_off.from_host(noff.data(), noff.size(), s);
  T.m2l_src.from_host(nsrc.data(), nsrc.size(), s);
  T.n_lr_local = (int64_t)nsrc.size();
  FMMB_CUDA(cudaStreamSynchronize(s));
}

void build_tree(fmmb_plan* plan, const double* points_host, int64_t n) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  Temp tmp;
  T.n = n;
  T.ncrit = plan->opts.ncrit;
  T.theta = plan->opts.theta;
  if (n >= (1ll << 31)) throw StatusError{FMMB_ERR_INVALID, "more than 2^31 bodies"};

  DevBuf<double>& pts = T.pts_orig;
  pts.from_host(points_host, 3 * (size_t)n, s);
  DevBuf<int> err;
  err.resize(1); err.zero(s);

  // bounding cube (Octree.hpp:67-79, BoundingBox.hpp:122-133)
  {
    const int nb = 256;
    DevBuf<double> part;
    part.resize(nb * 6);
    bbox_partial<<<nb, 256, 0, s>>>(pts.p, n, part.p);
    FMMB_CUDA(cudaGetLastError());
    std::vector<double> h = part.to_host(s);
    double mn[3], mx[3];
    for (int k = 0; k < 3; ++k) { mn[k] = h[k]; mx[k] = h[3 + k]; }
    for (int b = 1; b < nb; ++b)
      for (int k = 0; k < 3; ++k) {
        mn[k] = std::min(mn[k], h[6 * b + k]);
        mx[k] = std::max(mx[k], h[6 * b + 3 + k]);
      }
    volatile double ext = 0;
    for (int k = 0; k < 3; ++k) { volatile double d = std::fabs(mx[k] - mn[k]); if (d > ext) ext = d; }
    volatile double sc = ext * (1 + 1e-6);
    for (int k = 0; k < 3; ++k) {
      volatile double a = mn[k] + sc;
      mx[k] = std::max(mx[k], (double)a);
      T.pmin[k] = mn[k];
      volatile double d = mx[k] - mn[k];
      T.cell[k] = d / 1024.0;
    }
  }
  double3 pmin = make_double3(T.pmin[0], T.pmin[1], T.pmin[2]);
  double3 cell = make_double3(T.cell[0], T.cell[1], T.cell[2]);

  // Morton codes + stable sort by full code
  DevBuf<unsigned> code_orig, idx0, sc, sidx;
  code_orig.resize(n); idx0.resize(n); sc.resize(n); sidx.resize(n);
  morton_codes<<<nblk(n, 256), 256, 0, s>>>(pts.p, n, pmin, cell, code_orig.p, idx0.p, err.p);
  FMMB_CUDA(cudaGetLastError());
  sort_pairs_u32(tmp, code_orig.p, sc.p, idx0.p, sidx.p, n, 30, s);

  // BFS levels
  size_t cap = 1024;
  auto grow_boxes = [&](size_t need) {
    T.key.grow(need, s); T.parent.grow(need, s); T.cbegin.grow(need, s); T.cend.grow(need, s);
    T.bbegin.grow(need, s); T.bend.grow(need, s); T.level.grow(need, s);
  };
  grow_boxes(cap);
  {
    unsigned root[7] = {1u, 0u, 0u, 0u, 0u, (unsigned)n, 0u};
    FMMB_CUDA(cudaMemcpyAsync(T.key.p, &root[0], 4, cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaMemcpyAsync(T.parent.p, &root[1], 4, cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaMemcpyAsync(T.cbegin.p, &root[2], 4, cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaMemcpyAsync(T.cend.p, &root[3], 4, cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaMemcpyAsync(T.bbegin.p, &root[4], 4, cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaMemcpyAsync(T.bend.p, &root[5], 4, cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaMemcpyAsync(T.level.p, &root[6], 4, cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
  }
  T.level_off.clear();
  T.level_off.push_back(0);
  int lo = 0, hi = 1;
  DevBuf<unsigned> tmp_lo, tmp_hi, cmask;
  DevBuf<int> nchild, choff;
  for (int level = 0; hi > lo; ++level) {
    int np = hi - lo;
    tmp_lo.resize((size_t)np * 8); tmp_hi.resize((size_t)np * 8); cmask.resize(np);
    nchild.resize(np + 1); choff.resize(np + 1);
    nchild.zero(s);
    split_count<<<nblk((int64_t)np * 8, 256), 256, 0, s>>>(sc.p, T.bbegin.p, T.bend.p, lo, np, level, T.ncrit,
                                                          tmp_lo.p, tmp_hi.p, cmask.p, nchild.p, err.p);
    FMMB_CUDA(cudaGetLastError());
    exclusive_sum(tmp, nchild.p, choff.p, np + 1, s);
    int total = 0;
    FMMB_CUDA(cudaMemcpyAsync(&total, choff.p + np, sizeof(int), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    grow_boxes((size_t)hi + total);
    split_write<<<nblk((int64_t)np * 8, 256), 256, 0, s>>>(lo, np, level, hi, tmp_lo.p, tmp_hi.p, cmask.p, choff.p,
                                                          T.key.p, T.parent.p, T.cbegin.p, T.cend.p,
                                                          T.bbegin.p, T.bend.p, T.level.p);
    FMMB_CUDA(cudaGetLastError());
    lo = hi; hi += total;
    if (total > 0) T.level_off.push_back(lo);
  }
  T.nboxes = hi;
  T.level_off.push_back(T.nboxes);
  T.nlevels = (int)T.level_off.size() - 1;
  {
    int herr = 0;
    FMMB_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    if (herr & 1) throw StatusError{FMMB_ERR_INVALID, "a source point is not finite or falls outside the bounding cube"};
    if (herr & 2)
      throw StatusError{FMMB_ERR_TREE_DEPTH,
                        "a level-10 box still holds more than ncrit bodies (32-bit Morton codes resolve 10 "
                        "levels; the reference does not terminate on this input)"};
  }
  const int nb = T.nboxes;
  T.key.n = T.parent.n = T.cbegin.n = T.cend.n = T.bbegin.n = T.bend.n = T.level.n = nb;

  // leaves
  {
    DevBuf<int> flag, pos;
    flag.resize(nb + 1); pos.resize(nb + 1);
    flag_leaves<<<nblk(nb + 1, 256), 256, 0, s>>>(T.key.p, nb, flag.p);
    exclusive_sum(tmp, flag.p, pos.p, nb + 1, s);
    int nl = 0;
    FMMB_CUDA(cudaMemcpyAsync(&nl, pos.p + nb, sizeof(int), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    T.nleaves = nl;
    T.leaves.resize(nl);
    compact_leaves<<<nblk(nb, 256), 256, 0, s>>>(flag.p, pos.p, nb, T.leaves.p);
    FMMB_CUDA(cudaGetLastError());
  }

  // reference permutation: stable sort by the code masked to the leaf level
  {
    DevBuf<unsigned> masked, iota, mk_sorted;
    masked.resize(n); iota.resize(n); mk_sorted.resize(n);
    T.perm.resize(n); T.code.resize(n); T.body.resize(n);
    leaf_masked_codes<<<nblk((int64_t)T.nleaves * 32, 256), 256, 0, s>>>(T.leaves.p, T.nleaves, T.bbegin.p, T.bend.p,
                                                                        T.level.p, sc.p, sidx.p, masked.p, iota.p);
    FMMB_CUDA(cudaGetLastError());
    sort_pairs_u32(tmp, masked.p, mk_sorted.p, iota.p, T.perm.p, n, 30, s);
    gather_bodies<<<nblk(n, 256), 256, 0, s>>>(pts.p, T.perm.p, code_orig.p, n, T.body.p, T.code.p);
    FMMB_CUDA(cudaGetLastError());
  }

  // geometry
  T.center.resize(nb);
  box_geometry<<<nblk(nb, 256), 256, 0, s>>>(T.key.p, T.level.p, nb, pmin, cell, T.center.p);
  FMMB_CUDA(cudaGetLastError());

  // dual traversal (EvalInteractionLazy.hpp:65-105)
  TreeView tv{T.key.p, T.cbegin.p, T.cend.p, T.center.p, T.theta};
  DevBuf<int2> front, next, p2p_pairs;
  DevBuf<unsigned long long> cnt, off;
  DevBuf<int> cntp, offp;
  int2 rootpair = make_int2(0, 0);
  front.resize(1);
  FMMB_CUDA(cudaMemcpyAsync(front.p, &rootpair, sizeof(int2), cudaMemcpyHostToDevice, s));
  int64_t nfront = 1, n_lr = 0, n_p2p = 0;
  T.lr.resize(1024); T.lr.n = 0;
  p2p_pairs.resize(1024); p2p_pairs.n = 0;
  while (nfront > 0) {
    if (nfront >= (1ll << 31) - 1) throw StatusError{FMMB_ERR_INVALID, "traversal frontier exceeds 2^31 pairs"};
    cnt.resize(nfront + 1); off.resize(nfront + 1); cntp.resize(nfront + 1); offp.resize(nfront + 1);
    traverse_count<<<nblk(nfront + 1, 128), 128, 0, s>>>(tv, front.p, (int)nfront, cnt.p, cntp.p);
    FMMB_CUDA(cudaGetLastError());
    exclusive_sum(tmp, cnt.p, off.p, nfront + 1, s);
    exclusive_sum(tmp, cntp.p, offp.p, nfront + 1, s);
    unsigned long long tot = 0; int totp = 0;
    FMMB_CUDA(cudaMemcpyAsync(&tot, off.p + nfront, sizeof(tot), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaMemcpyAsync(&totp, offp.p + nfront, sizeof(int), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    int64_t add_lr = (int64_t)(tot & 0xffffffffull), add_q = (int64_t)(tot >> 32);
    T.lr.n = n_lr; T.lr.grow(n_lr + add_lr, s);
    p2p_pairs.n = n_p2p; p2p_pairs.grow(n_p2p + totp, s);
    next.resize(add_q);
    traverse_write<<<nblk(nfront, 128), 128, 0, s>>>(tv, front.p, (int)nfront, off.p, offp.p,
                                                    T.lr.p + n_lr, next.p, p2p_pairs.p + n_p2p);
    FMMB_CUDA(cudaGetLastError());
    FMMB_CUDA(cudaStreamSynchronize(s));
    n_lr += add_lr; n_p2p += totp;
    std::swap(front.p, next.p); std::swap(front.cap, next.cap); std::swap(front.n, next.n);
    nfront = add_q;
  }
  T.n_lr = n_lr; T.n_p2p = n_p2p;
  T.lr.n = n_lr; p2p_pairs.n = n_p2p;

  // target-major CSRs
  pairs_to_csr(tmp, T.lr, n_lr, /*key_is_y=*/1, nb, T.m2l_off, T.m2l_src, s);
  pairs_to_csr(tmp, p2p_pairs, n_p2p, /*key_is_y=*/0, nb, T.p2p_off, T.p2p_src, s);

  // which boxes carry a local expansion (propagate_local, EvalInteractionLazy.hpp:184-205)
  T.has_local.resize(nb); T.has_local.zero(s);
  if (n_lr) mark_targets<<<nblk(n_lr, 256), 256, 0, s>>>(T.lr.p, n_lr, T.has_local.p);
  for (int l = 1; l < T.nlevels; ++l) {
    int a = T.level_off[l], b = T.level_off[l + 1];
    if (b > a) inherit_local<<<nblk(b - a, 256), 256, 0, s>>>(T.parent.p, a, b, T.has_local.p);
  }
  FMMB_CUDA(cudaGetLastError());
  // which multipoles a matvec needs: the sources of the accepted pairs and everything below them (a parent's
  // multipole is built from its children's).  The root and the first levels under it are never a source at
  // theta < 1, so the upward sweep stops early -- like the reference's lazy lists (EvalInteractionLazy.hpp:158-183).
  T.need_M.resize(nb); T.need_M.zero(s);
  if (n_lr) mark_sources<<<nblk(n_lr, 256), 256, 0, s>>>(T.lr.p, n_lr, T.need_M.p);
  for (int l = 1; l < T.nlevels; ++l) {
    int a = T.level_off[l], b = T.level_off[l + 1];
    if (b > a) inherit_local<<<nblk(b - a, 256), 256, 0, s>>>(T.parent.p, a, b, T.need_M.p);
  }
  FMMB_CUDA(cudaGetLastError());
  T.need_M_host = T.need_M.to_host(s);

  // work count
  {
    DevBuf<unsigned long long> total;
    total.resize(1); total.zero(s);
    if (n_p2p) count_body_pairs<<<nblk(n_p2p, 256), 256, 0, s>>>(p2p_pairs.p, n_p2p, T.bbegin.p, T.bend.p, total.p);
    FMMB_CUDA(cudaGetLastError());
    std::vector<unsigned long long> h = total.to_host(s);
    T.n_p2p_body_pairs = (int64_t)h[0];
  }
  FMMB_CUDA(cudaStreamSynchronize(s));
  partition_tree(plan);
  FMMB_CUDA(cudaStreamSynchronize(s));
}


// FMMOptions::block_diagonal (reference include/executor/EvalDiagonalSparse.hpp:32-47): the near field keeps only
// the interaction of every leaf with itself.  Rewrites the target-major P2P lists in place; everything built from
// them afterwards (work items, source runs, cached BEM blocks) follows.
namespace {
__global__ void self_list_offsets(const unsigned* __restrict__ key, int nb, int* __restrict__ flag) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= nb) flag[b] = b < nb ? (int)(key[b] >> 31) : 0;
}
__global__ void self_list_fill(const unsigned* __restrict__ key, const int* __restrict__ off, int nb, int* __restrict__ src) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb && (key[b] >> 31)) src[off[b]] = b;
}
}  // namespace

void restrict_p2p_to_self(fmmb_plan* plan) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  const int nb = T.nboxes;
  DevBuf<int> flag;
  flag.resize(nb + 1);
  self_list_offsets<<<(nb + 256) / 256, 256, 0, s>>>(T.key.p, nb, flag.p);
  T.p2p_off.resize(nb + 1);
  size_t bytes = 0;
  FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, T.p2p_off.p, nb + 1, s));
  DevBuf<char> tmp;
  tmp.resize(bytes);
  FMMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, flag.p, T.p2p_off.p, nb + 1, s));
  T.n_p2p = T.nleaves;
  T.p2p_src.resize(T.n_p2p);
  self_list_fill<<<(nb + 255) / 256, 256, 0, s>>>(T.key.p, T.p2p_off.p, nb, T.p2p_src.p);
  FMMB_CUDA(cudaGetLastError());
  std::vector<unsigned> key = T.key.to_host(s), b0 = T.bbegin.to_host(s), b1 = T.bend.to_host(s);
  int64_t pairs = 0;
  for (int b = 0; b < nb; ++b) if (key[b] >> 31) pairs += (int64_t)(b1[b] - b0[b]) * (b1[b] - b0[b]);
  T.n_p2p_body_pairs = pairs;
}

}  // namespace fmmb
