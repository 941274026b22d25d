// csrc/trans_blocked.cu -- output-stationary box-to-box translations (M2L, M2M, L2L) on the FP64 tensor path.
//
// The reference applies kernel/LaplaceSpherical.hpp:296-329 (M2L), :245-285 (M2M), :378-411 (L2L) pair by pair
// (include/executor/EvalInteractionLazy.hpp:122-153).  A translation is linear in the source expansion and its
// vector takes few distinct values in an octree (box centres sit on a lattice of half finest cells), so for a
// translation class c
//        X_out[:, target] += T_c (P^2 x P^2, real) * X_in[:, source]
// with expansions stored as P^2 reals (laplace_ops.cuh).  Round 1 ran this class-major (one GEMM per class, columns
// written to a 2 GB scratch array, a second kernel summed the columns of a target).  Here the OUTPUT is stationary:
//
//  * targets of one level are grouped into blocks of 16 sibling families (<= 128 boxes).  Column of a box inside its
//    block = octant * 16 + family slot, i.e. the columns are ordered by the parity of the box coordinates.  Which
//    targets own a pair of a given class depends (away from the domain boundary) only on that parity, so the active
//    columns of a class come in aligned groups of 8 = one DMMA.8x8x4 column tile; only active tiles are computed.
//  * a CTA (8 warps, one per 8-row tile of T_c) owns one block, keeps its 64 x 128 accumulator in registers and walks
//    the (class, active tile mask) items of the block.  The operands of an item arrive by 1-D TMA bulk copies
//    (cp.async.bulk.shared::cluster.global, completion in bytes on an mbarrier) into a two-stage ring: the T_c
//    slice (fragment-major, 32 KB, one copy) and the source expansion of every active column (one copy each).
//    The accumulator is written once: no scratch columns, no reduction kernel, fixed summation order.
//  * M2M level by level, M2L and L2L level by level are ONE launch: work units are listed phase by phase, CTAs are
//    dispatched in index order, and a unit waits on a device-side counter until the previous phase has written
//    its expansions (acquire / release at GPU scope).  No launch gaps between the short top-of-tree steps.
//  * small phases (multi-GPU shards, top levels, BEM meshes) split the item list of a block over 2..8 CTAs; the
//    partial accumulators meet in a small scratch array and the last CTA to arrive adds them in split order.
//  * orders 9..16: the contraction is cut into k-chunks of 64 and the rows into blocks of 64.
//
// Where the time goes (B200, N = 1M, P = 8, the M2L phase of the sweep, 1.77 ms): with the DMMAs removed the phase
// takes 0.96 ms, with the operand copies removed 1.44 ms, with both removed 0.33 ms (the per-item loop: barrier,
// descriptors).  The math alone (1.1 ms) is at the DMMA pipe's rate for the 16.8 % padded tile count; the copies move
// 5.5 GB out of L2 (3.2 GB of T slices, 2.3 GB of source expansions) at the 8.9 TB/s this access shape measures
// (scripts/micro/probe_r2.cu) and overlap the math poorly at a prefetch distance of one item, because half of the
// items are the two-tile items of the classes with three odd offset components.  The class-major engine of
// m2l_classes.cu keeps T_c in registers across ~190 items and pays with a column scratch instead; at orders <= 8 it
// is the faster of the two on a uniform tree (1.58 ms) and the default there.
#include "common.cuh"
#include "laplace_ops.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstring>

namespace fmmb {

namespace {

using namespace ops;

constexpr int kCols = 128;       // columns (target boxes) per block
constexpr int kThreads = 512;    // 16 warps
constexpr int kWarps = kThreads / 32;
constexpr int kPartial = 16 * 64 * 2 * 4;   // doubles of one partial accumulator: 64 rows x 128 columns

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

struct Temp {
  DevBuf<char> buf;
  void* get(size_t bytes) { if (bytes > buf.cap) buf.resize(bytes); return buf.p; }
};

template <int P>
struct Cfg {
  static constexpr int PP = P * P;
  static constexpr int XS = (PP + 1) & ~1;                 // doubles per expansion in global memory
  static constexpr int KC = (PP + 63) / 64;                // k chunks of 64
  static constexpr int RBLK = KC;                          // row blocks of 64 (grid.y)
  static constexpr int KCH = PP < 64 ? PP : 64;            // k (and rows) of a full chunk
  static constexpr int KB = (KCH + 3) / 4;                 // DMMA k-steps per chunk
  static constexpr int KB2 = (KB + 1) / 2;                 // ... loaded two at a time
  static constexpr int NRT = (KCH + 7) / 8;                // row tiles per row block
  // a warp owns ONE row tile (its T fragments stay in registers for the item) and TPW column tiles.  Measured
  // alternatives at N = 1M, P = 8: 8 warps x (1 row tile, 16 column tiles) 2.27 ms, 8 warps x (2 row tiles, 8 column
  // tiles; half the shared-memory operand traffic) 1.86 ms, 16 warps x (1 row tile, 8 column tiles) 1.70 ms: with
  // one CTA per SM the warp count matters more than the operand reuse, and 16 x 32 x 128 registers is the file.
  static constexpr int NRTW = NRT > 4 ? 8 : (NRT > 2 ? 4 : (NRT > 1 ? 2 : 1));   // warps along the rows
  static constexpr int NCG = kWarps / NRTW;                // warps along the columns
  static constexpr int TPW = 16 / NCG;                     // column tiles per warp
  static constexpr int KMAX = 8 * KB2 > XS ? 8 * KB2 : (XS < 64 ? XS : 64);
  // column stride of a stage: even (16-byte bulk-copy destinations), congruent 4 mod 16 so that the 8 columns x 4 k
  // of one B fragment cover all 16 eight-byte banks exactly twice
  static constexpr int LDB = KMAX + ((4 - KMAX % 16) + 16) % 16;
  static constexpr int TSLICE = NRTW * KB2 * 64;           // doubles of one (class, row block, k chunk) slice of T
  static constexpr int CLS_STRIDE = RBLK * KC * TSLICE;    // doubles per class, fragment-major
  static constexpr int STAGE = kCols * LDB + TSLICE;       // doubles per pipeline stage: source columns | T slice
  static constexpr size_t SMEM = (size_t)2 * STAGE * sizeof(double);
  // a bulk copy overwrites the first XS (or chunk) doubles of a column; what the k loop reads beyond must be finite
  static constexpr bool NEED_ZERO = 4 * KB > XS || KC > 1;
};

struct FragShape { int KC, NRTW, KB2, cls_stride; };
template <int P> FragShape frag_shape() { return FragShape{Cfg<P>::KC, Cfg<P>::NRTW, Cfg<P>::KB2, Cfg<P>::CLS_STRIDE}; }

// fragment-major position of T_c[row][k]: [row block][k chunk][row tile][k-step pair][lane][2]
__device__ __forceinline__ int frag_off(const FragShape f, int row, int k) {
  const int rb = row >> 6, rr = row & 63, rt = rr >> 3, lr = rr & 7;
  const int kc = k >> 6, kk = k & 63, kb = kk >> 2, lk = kk & 3;
  return ((((rb * f.KC + kc) * f.NRTW + rt) * f.KB2 + (kb >> 1)) * 32 + lr * 4 + lk) * 2 + (kb & 1);
}

// ---- translation matrices (real layout: Re X_n^m at n^2+n+m, Im X_n^m at n^2+n-m) ---------------------------------
__global__ void __launch_bounds__(256)
build_T_m2l_frag(int P, FragShape f, const double4* __restrict__ vec, double* __restrict__ Tt) {
  extern __shared__ double2 Y[];       // (2P)^2
  const int pp = P * P, c = blockIdx.x;
  const double4 v = vec[c];
  const Sph s = to_sph(v.x, v.y, v.z);
  for (int m = threadIdx.x; m < 2 * P; m += blockDim.x) harmonics_column<true>(m, 2 * P, s, 1.0, Y);
  __syncthreads();
  double* T = Tt + (size_t)c * f.cls_stride;
  for (int idx = threadIdx.x; idx < pp * pp; idx += blockDim.x) {
    const int col = idx / pp, row = idx % pp;
    int j = 0; while ((j + 1) * (j + 1) <= row) ++j;
    const int kk = row - j * j - j;            // >= 0: Re L_j^k, < 0: Im L_j^{-kk}
    int n = 0; while ((n + 1) * (n + 1) <= col) ++n;
    const int mm = col - n * n - n;            // >= 0: Re M_n^m, < 0: Im M_n^{-mm}
    const int k = abs(kk), m = abs(mm);
    const int base = (j + n) * (j + n) + j + n - k;
    // W(+-m) = Cnm(+-m) * Y_{j+n}^{+-m-k};  M = a + ib contributes a (W+ + W-) + i b (W+ - W-)
    const double cp_ = cnm_real(j, k, n, m);
    const double2 yp = Y[base + m];
    const double wpr = cp_ * yp.x, wpi = cp_ * yp.y;
    double val;
    if (m == 0) {
      val = kk >= 0 ? wpr : wpi;
    } else {
      const double cm_ = cnm_real(j, k, n, -m);
      const double2 ym = Y[base - m];
      const double wmr = cm_ * ym.x, wmi = cm_ * ym.y;
      if (kk >= 0) val = (mm >= 0) ? (wpr + wmr) : -(wpi - wmi);
      else val = (mm >= 0) ? (wpi + wmi) : (wpr - wmr);
    }
    T[frag_off(f, row, col)] = val;
  }
}

// M2M / L2L matrices by probing the operator with unit vectors (the operators contain a complex conjugation, so
// they are linear over the reals only).
template <int KIND>   // 1 = M2M, 2 = L2L
__global__ void __launch_bounds__(128)
build_T_probe_frag(int P, FragShape f, const double4* __restrict__ vec, double* __restrict__ Tt) {
  extern __shared__ double2 shp[];
  const int pp = P * P, nc = P * (P + 1) / 2, c = blockIdx.x;
  double2* Y = shp;          // pp
  double2* E = shp + pp;     // nc
  const double4 v = vec[c];
  const Sph s = to_sph(v.x, v.y, v.z);
  for (int m = threadIdx.x; m < P; m += blockDim.x) harmonics_column<false>(m, P, s, KIND == 1 ? -1.0 : 1.0, Y);
  double* T = Tt + (size_t)c * f.cls_stride;
  for (int col = 0; col < pp; ++col) {
    int n = 0; while ((n + 1) * (n + 1) <= col) ++n;
    const int mm = col - n * n - n;
    const int hot = n * (n + 1) / 2 + abs(mm);
    __syncthreads();
    for (int i = threadIdx.x; i < nc; i += blockDim.x)
      E[i] = i == hot ? (mm >= 0 ? make_double2(1, 0) : make_double2(0, 1)) : make_double2(0, 0);
    __syncthreads();
    for (int jks = threadIdx.x; jks < nc; jks += blockDim.x) {
      int j, k;
      unpack_nm(jks, j, k);
      const double2 o = KIND == 1 ? m2m_entry(E, Y, j, k) : l2l_entry(E, Y, j, k, P);
      T[frag_off(f, j * j + j + k, col)] = o.x;
      if (k > 0) T[frag_off(f, j * j + j - k, col)] = o.y;
    }
  }
}

// ---- plan time: classes, blocks, items -----------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned compact10(unsigned x) {
  x &= 0x09249249u;
  x = (x | (x >> 2)) & 0x030C30C3u;
  x = (x | (x >> 4)) & 0x0300F00Fu;
  x = (x | (x >> 8)) & 0x030000FFu;
  x = (x | (x >> 16)) & 0x000003FFu;
  return x;
}
// centre of a box in units of half a finest cell (exact integers)
__device__ __forceinline__ int3 centre_half_cells(unsigned key, unsigned L) {
  unsigned m = (key & 0x7fffffffu) << (3u * (10 - L));
  m &= ~(1u << 30);
  const int half = L >= 10 ? 1 : (1 << (10 - L));       // box spans 2^(11-L) half cells
  return make_int3(2 * (int)compact10(m) + half, 2 * (int)compact10(m >> 1) + half, 2 * (int)compact10(m >> 2) + half);
}
__global__ void pair_class_keys(const int* __restrict__ tgt, const int* __restrict__ src, int64_t n,
                                const unsigned* __restrict__ key, const unsigned* __restrict__ lvl,
                                unsigned long long* __restrict__ ckey, int* __restrict__ idx) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int t = tgt[e], s = src[e];
  const int3 a = centre_half_cells(key[t], lvl[t]), b = centre_half_cells(key[s], lvl[s]);
  const unsigned long long dx = (unsigned)(a.x - b.x + 2048), dy = (unsigned)(a.y - b.y + 2048),
                           dz = (unsigned)(a.z - b.z + 2048);
  ckey[e] = (dx << 24) | (dy << 12) | dz;
  idx[e] = (int)e;
}
__global__ void head_flags(const unsigned long long* __restrict__ k, int64_t n, int shift, int* __restrict__ flag) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n) flag[e] = (e == 0 || (k[e] >> shift) != (k[e - 1] >> shift)) ? 1 : 0;
}
// class id of every pair + the representative translation vector of every class (the integer centre offset times
// half a finest cell: independent of which pairs a rank happens to hold)
__global__ void scatter_class_ids(const int* __restrict__ sorted_idx, const int* __restrict__ head,
                                  const int* __restrict__ scan, int64_t n, const int* __restrict__ tgt,
                                  const int* __restrict__ src, const unsigned* __restrict__ key,
                                  const unsigned* __restrict__ lvl, double3 half_cell, int* __restrict__ cid,
                                  double4* __restrict__ vec) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int c = scan[e] - 1, i = sorted_idx[e];
  cid[i] = c;
  if (head[e]) {
    const int t = tgt[i], s = src[i];
    const int3 a = centre_half_cells(key[t], lvl[t]), b = centre_half_cells(key[s], lvl[s]);
    vec[c] = make_double4((a.x - b.x) * half_cell.x, (a.y - b.y) * half_cell.y, (a.z - b.z) * half_cell.z, 0.0);
  }
}
__global__ void mark_targets_and_parents(const int* __restrict__ tgt, int64_t n, const unsigned* __restrict__ parent,
                                         int* __restrict__ is_tgt, int* __restrict__ pflag) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int t = tgt[e];
  is_tgt[t] = 1;
  if (t > 0) pflag[parent[t]] = 1;
}
struct LevelTable { int blk_base[16]; int pr_base[16]; };   // per TARGET level: first block, rank offset of its parents
__global__ void place_targets(const int* __restrict__ is_tgt, int nb, const unsigned* __restrict__ key,
                              const unsigned* __restrict__ lvl, const unsigned* __restrict__ parent,
                              const int* __restrict__ pr, LevelTable lt, int* __restrict__ box_blk,
                              int* __restrict__ box_col, int* __restrict__ blk_cols) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb || !is_tgt[b]) return;
  const int l = (int)lvl[b];
  int blk, col;
  if (l == 0) { blk = lt.blk_base[0]; col = 0; }
  else {
    const int r = pr[parent[b]] - lt.pr_base[l];
    blk = lt.blk_base[l] + (r >> 4);
    col = (int)(key[b] & 7u) * 16 + (r & 15);
  }
  box_blk[b] = blk; box_col[b] = col;
  blk_cols[(size_t)blk * kCols + col] = b;
}
__global__ void pair_sort_keys(const int* __restrict__ tgt, const int* __restrict__ cid, const int* __restrict__ box_blk,
                               const int* __restrict__ box_col, int64_t n, unsigned long long* __restrict__ k) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int t = tgt[e];
  k[e] = ((unsigned long long)box_blk[t] << 27) | ((unsigned long long)cid[e] << 7) | (unsigned)box_col[t];
}
__global__ void fill_items_tiles(const unsigned long long* __restrict__ k, const int* __restrict__ src_sorted, int64_t n,
                                 const int* __restrict__ item_head, const int* __restrict__ item_scan,
                                 const int* __restrict__ tile_head, const int* __restrict__ tile_scan,
                                 int2* __restrict__ items, int* __restrict__ tile_src, int* __restrict__ blk_first) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const unsigned long long key = k[e];
  const int col = (int)(key & 127u), cls = (int)((key >> 7) & 0xfffffu), blk = (int)(key >> 27);
  const int item = item_scan[e] - 1, tile = tile_scan[e] - 1;
  tile_src[(size_t)tile * 8 + (col & 7)] = src_sorted[e];
  if (tile_head[e]) atomicOr(&items[item].x, 1 << (16 + (col >> 3)));
  if (item_head[e]) {
    atomicOr(&items[item].x, cls);
    items[item].y = tile;
    if (e == 0 || (k[e - 1] >> 27) != (unsigned long long)blk) blk_first[blk] = item;
  }
}
__global__ void fill_int(int* __restrict__ a, int64_t n, int v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void block_weights(const int* __restrict__ blk_item_off, const int2* __restrict__ items, int n_blocks,
                              int n_items, int n_tiles, unsigned* __restrict__ w, int* __restrict__ id) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const int i0 = blk_item_off[b], i1 = blk_item_off[b + 1];
  const int t0 = items[i0].y, t1 = i1 < n_items ? items[i1].y : n_tiles;
  w[b] = (unsigned)(t1 - t0);
  id[b] = b;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
      " bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP.S.G)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
               "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

struct BlkArgs {
  const double* T;               // fragment-major matrices of this order
  const int2* items;             // x = class | tile mask << 16, y = first tile
  const int* tile_src;           // 8 source boxes per active tile (zero row for absent pairs)
  const int* blk_item_off;
  const int* blk_cols;           // target box per column, -1 = none
  const double* X;               // source expansions
  double* Out;                   // target expansions
  int add;                       // 0: Out = sum (M2L, M2M), 1: Out += sum (L2L)
};
// One launch = one SWEEP: a list of work units (block x split part x row block) of up to three batches, cut into
// phases (M2M level by level, then M2L, then L2L level by level) that depend on one another in a chain.  CTAs are
// dispatched in index order and units are listed phase by phase, so a unit that waits for the previous phase only
// ever waits for CTAs that are already running or done (the forward-progress argument of decoupled look-back).
struct SweepArgs {
  BlkArgs b[3];
  const int4* units;             // x = batch | phase << 4, y = block, z = split part | split << 8, w = scratch slot
  unsigned* phase_cnt;           // blocks of a phase whose output is written (zeroed before the launch)
  const int* phase_total;
  double* scratch;               // split > 1: partial accumulators
  unsigned* split_cnt;
  long long* trace;              // debug (FMMB_TRACE=1): globaltimer stamps of thread 0 of the first CTAs
};
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define FMMB_STAMP(i) do { if (a.trace && tid == 0 && blockIdx.x < 2048) a.trace[(size_t)blockIdx.x * 8 + (i)] = gtime(); } while (0)
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_inc(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

template <int P>
__global__ void __launch_bounds__(kThreads, 1)
trans_sweep_kernel(const SweepArgs a) {
  using C = Cfg<P>;
  constexpr int PP = C::PP, XS = C::XS, KC = C::KC, KB = C::KB, KB2 = C::KB2, TPW = C::TPW, LDB = C::LDB;
  extern __shared__ __align__(128) double stage[];          // 2 x ([kCols][LDB] source columns | T slice)
  __shared__ __align__(8) uint64_t full[2];
  __shared__ int s_last;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  const int rt = w % C::NRTW, cg = w / C::NRTW;
  const bool rows_live = rt < C::NRT;
  const int ui = blockIdx.x / C::RBLK, rb = blockIdx.x - ui * C::RBLK;
  const int4 u = a.units[ui];
  const int phase = u.x >> 4, blk = u.y, cs = u.z & 255, split = u.z >> 8;
  const BlkArgs& B = a.b[u.x & 15];
  int i0 = B.blk_item_off[blk], i1 = B.blk_item_off[blk + 1];
  {
    const int per = (i1 - i0 + split - 1) / split;
    i0 = min(i1, i0 + cs * per);
    i1 = min(i1, i0 + per);
  }
  const int nv = (i1 - i0) * KC;                            // virtual items: (item, k chunk)
  FMMB_STAMP(0);

  // stage buffers start finite (zero) where a copy does not reach
  if (C::NEED_ZERO)
    for (int i = tid; i < 2 * C::STAGE; i += kThreads) stage[i] = 0.0;
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic zero stores before async-proxy writes

  auto chunk_bytes = [&](int kc) -> unsigned {
    return (unsigned)((KC == 1 ? XS : (kc < KC - 1 ? 64 : XS - 64 * (KC - 1))) * sizeof(double));
  };
  // The operands of virtual item v.  The T_c slice (fragment-major, 32 KB at P = 8) arrives by 1-D TMA bulk copies,
  // one sixteenth per warp, completion in bytes on the stage's mbarrier.  The source expansions of the active columns
  // are scattered boxes: warp w copies columns w, w + 16, ... (at most 8) with 16-byte cp.async, lane = chunk;
  // srcv: lane j < 8 holds the source box of column w + 16 j.  (One bulk copy per column costs ~28 cycles per
  // request in the SM's TMA queue -- measured -- which is a quarter of an item.)
  auto produce = [&](int v, const int2 d, const int srcv) {
    const int s = v & 1, kc = KC == 1 ? 0 : v % KC;
    const unsigned mask = (unsigned)d.x >> 16;
    const int ncols = __popc(mask) * 8;
    const int chunks = (int)(chunk_bytes(kc) / 16);
    double* dst = stage + (size_t)s * C::STAGE;
    if (tid == 0) mbar_expect_tx(&full[s], (unsigned)(C::TSLICE * sizeof(double)));
    if (lane == 0) {
      constexpr int PART = C::TSLICE / kWarps;               // doubles; a multiple of 4
      bulk_g2s(dst + kCols * LDB + w * PART,
               B.T + (size_t)(d.x & 0xffff) * C::CLS_STRIDE + (size_t)(rb * KC + kc) * C::TSLICE + w * PART,
               (unsigned)(PART * sizeof(double)), &full[s]);
    }
#pragma unroll
    for (int j = 0; j < 128 / kWarps; ++j) {
      const int c = w + kWarps * j;
      if (c < ncols) {                                        // warp-uniform
        const int sb = __shfl_sync(0xffffffffu, srcv, j);
        if (lane < chunks) cp_async16(dst + c * LDB + 2 * lane, B.X + (size_t)sb * XS + kc * 64 + 2 * lane);
      }
    }
    cp_async_commit();
  };
  auto load_src = [&](const int2 d) -> int {
    const int c = w + kWarps * lane;
    return (lane < 128 / kWarps && c < __popc((unsigned)d.x >> 16) * 8) ? B.tile_src[(size_t)d.y * 8 + c] : 0;
  };
  auto item_of = [&](int v) { return i0 + (KC == 1 ? v : v / KC); };

  double acc[TPW][2];
#pragma unroll
  for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = 0.0;

  // descriptors run three virtual items ahead of the math, source indices two: all of this is plan data and is
  // fetched BEFORE the wait for the previous phase
  int2 d0 = make_int2(0, 0), d1 = d0, d2 = d0;
  int src0 = 0, src1 = 0;
  if (nv > 0) {
    d0 = B.items[item_of(0)];
    if (nv > 1) d1 = B.items[item_of(1)];
    if (nv > 2) d2 = B.items[item_of(2)];
    src0 = load_src(d0);
    if (nv > 1) src1 = load_src(d1);
  }
  // the expansions this unit reads are written by the previous phase of the same launch
  if (phase > 0) {
    if (tid == 0) {
      const unsigned need = (unsigned)a.phase_total[phase - 1] * C::RBLK;   // every row block of every block signals
      while (ld_acquire_gpu(a.phase_cnt + phase - 1) < need) __nanosleep(40);
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  __syncthreads();
  FMMB_STAMP(1);
  if (nv > 0) produce(0, d0, src0);

  for (int v = 0; v < nv; ++v) {
    cp_async_wait_all();                                     // this thread's column chunks of item v have landed
    __syncthreads();                                         // ... everyone's; and everyone is done with item v - 1
    if (v + 1 < nv) produce(v + 1, d1, src1);
    const int src2 = v + 2 < nv ? load_src(d2) : 0;
    const int2 d3 = v + 3 < nv ? B.items[item_of(v + 3)] : make_int2(0, 0);
    mbar_wait(&full[v & 1], (unsigned)(v >> 1) & 1u);
    if (v == 0) FMMB_STAMP(2);
    if (rows_live) {
      const unsigned mask = (unsigned)d0.x >> 16;
      const double* st = stage + (size_t)(v & 1) * C::STAGE;
      // The warp's 8 rows of T_c (fragment-major, 16 bytes per lane and k-step pair) are taken in two halves of the k
      // range: half the fragment registers, which leaves room to run the source-fragment loads from shared memory one
      // k-step pair ahead of the DMMAs that use them.  Column tiles are dealt round-robin to the NCG warp columns (the
      // tiles of a class spread over all of them) and taken two at a time with their k loops interleaved: independent
      // accumulator chains hide the 26-cycle latency of a dependent DMMA.
      constexpr int QH = KB2 > 4 ? (KB2 + 1) / 2 : KB2;      // k-step pairs per half
      const double2* As = reinterpret_cast<const double2*>(st + kCols * LDB) + rt * KB2 * 32 + lane;
      const double* Bw = st + (size_t)lr * LDB + lk;
      auto bcol = [&](int ct) { return Bw + (size_t)__popc(mask & ((1u << ct) - 1u)) * 8 * LDB; };
#pragma unroll
      for (int q0 = 0; q0 < KB2; q0 += QH) {
        double2 A2[QH];
#pragma unroll
        for (int q = 0; q < QH; ++q) A2[q] = q0 + q < KB2 ? As[(q0 + q) * 32] : make_double2(0.0, 0.0);
        auto one = [&](double (&c)[2], const double* Bs) {
          double bx = Bs[8 * q0], by = Bs[8 * q0 + 4];
#pragma unroll
          for (int q = 0; q < QH; ++q) {
            if (q0 + q >= KB2) break;
            const double cx = bx, cy = by;
            if (q + 1 < QH && q0 + q + 1 < KB2) { bx = Bs[8 * (q0 + q + 1)]; by = Bs[8 * (q0 + q + 1) + 4]; }
            dmma8x8x4(c[0], c[1], A2[q].x, cx);
            if (2 * (q0 + q) + 1 < KB) dmma8x8x4(c[0], c[1], A2[q].y, cy);
          }
        };
        auto two = [&](double (&c)[2], const double* Bs, double (&e)[2], const double* Es) {
          double bx = Bs[8 * q0], by = Bs[8 * q0 + 4], ex = Es[8 * q0], ey = Es[8 * q0 + 4];
#pragma unroll
          for (int q = 0; q < QH; ++q) {
            if (q0 + q >= KB2) break;
            const double cx = bx, cy = by, dx = ex, dy = ey;
            if (q + 1 < QH && q0 + q + 1 < KB2) {
              bx = Bs[8 * (q0 + q + 1)]; by = Bs[8 * (q0 + q + 1) + 4];
              ex = Es[8 * (q0 + q + 1)]; ey = Es[8 * (q0 + q + 1) + 4];
            }
            dmma8x8x4(c[0], c[1], A2[q].x, cx);
            dmma8x8x4(e[0], e[1], A2[q].x, dx);
            if (2 * (q0 + q) + 1 < KB) {
              dmma8x8x4(c[0], c[1], A2[q].y, cy);
              dmma8x8x4(e[0], e[1], A2[q].y, dy);
            }
          }
        };
        if (TPW == 1) {
          if ((mask >> cg) & 1u) one(acc[0], bcol(cg));
        } else {
#pragma unroll
          for (int t = 0; t < TPW; t += 2) {
            const int c0 = t * C::NCG + cg, c1 = (t + 1) * C::NCG + cg;
            const bool h0 = (mask >> c0) & 1u, h1 = (mask >> c1) & 1u;
            if (h0 && h1) two(acc[t], bcol(c0), acc[t + 1], bcol(c1));
            else if (h0) one(acc[t], bcol(c0));
            else if (h1) one(acc[t + 1], bcol(c1));
          }
        }
      }
    }
    d0 = d1; d1 = d2; d2 = d3; src1 = src2;
  }
  FMMB_STAMP(3);

  // ---- partial accumulators of a split block meet in scratch ([value][thread]: coalesced); the last CTA to arrive
  // adds them in split order
  if (split > 1) {
    // u.w = first partial buffer of the block (for one row block); the block's RBLK x split buffers are contiguous
    const size_t group = (size_t)u.w * C::RBLK + (size_t)rb * split;
    double* part = a.scratch + (group + cs) * kPartial + tid;
#pragma unroll
    for (int t = 0; t < TPW; ++t) { part[(2 * t) * kThreads] = acc[t][0]; part[(2 * t + 1) * kThreads] = acc[t][1]; }
    __syncthreads();
    if (tid == 0) {
      // one fence for the block (the barrier orders the other threads' stores before it)
      __threadfence();
      s_last = atomicAdd(a.split_cnt + (size_t)u.w * C::RBLK + rb, 1u) == (unsigned)split - 1;
      __threadfence();
    }
    __syncthreads();
    FMMB_STAMP(4);
    if (!s_last) return;
#pragma unroll
    for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = 0.0;
    for (int q = 0; q < split; ++q) {
      const double* pq = a.scratch + (group + q) * kPartial + tid;
      double v[TPW * 2];
#pragma unroll
      for (int t = 0; t < TPW * 2; ++t) v[t] = __ldcg(pq + t * kThreads);
#pragma unroll
      for (int t = 0; t < TPW; ++t) { acc[t][0] += v[2 * t]; acc[t][1] += v[2 * t + 1]; }
    }
    if (tid == 0) a.split_cnt[(size_t)u.w * C::RBLK + rb] = 0;
  }
  FMMB_STAMP(5);

  // ---- the block's expansions, written once
  const int row = rb * 64 + rt * 8 + lr;
  if (rows_live && row < PP) {
    const int* cols = B.blk_cols + (size_t)blk * kCols;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int cbase = (t * C::NCG + cg) * 8 + 2 * lk;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int tb = cols[cbase + i];
        if (tb < 0) continue;
        double* o = B.Out + (size_t)tb * XS + row;
        if (B.add) *o += acc[t][i];
        else {
          *o = acc[t][i];
          if (XS > PP && row == PP - 1) o[1] = 0.0;         // padding double of odd-sized expansions stays zero
        }
      }
    }
  }
  // ---- tell the next phase
  __syncthreads();
  if (tid == 0) red_release_gpu_inc(a.phase_cnt + phase);
  FMMB_STAMP(6);
}

template <int P>
void build_T_t(fmmb_plan* plan, BlkBatch& B, DevBuf<double>& T, cudaStream_t s) {
  const FragShape f = frag_shape<P>();
  T.resize((size_t)B.n_classes * f.cls_stride);
  T.zero(s);
  if (B.n_classes == 0) return;
  if (B.kind == 0)
    build_T_m2l_frag<<<(int)B.n_classes, 256, (size_t)4 * P * P * sizeof(double2), s>>>(P, f, B.class_vec.p, T.p);
  else if (B.kind == 1)
    build_T_probe_frag<1><<<(int)B.n_classes, 128, (size_t)(P * P + P * (P + 1) / 2) * sizeof(double2), s>>>(
        P, f, B.class_vec.p, T.p);
  else
    build_T_probe_frag<2><<<(int)B.n_classes, 128, (size_t)(P * P + P * (P + 1) / 2) * sizeof(double2), s>>>(
        P, f, B.class_vec.p, T.p);
  FMMB_CUDA(cudaGetLastError());
  ++plan->launches;
}

#define FMMB_FOR_P(P, CALL)                                                                                          \
  switch (P) {                                                                                                       \
    case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break;                  \
    case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break;                  \
    case 9: CALL(9); break; case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break;            \
    case 13: CALL(13); break; case 14: CALL(14); break; case 15: CALL(15); break; default: CALL(16); break;          \
  }

const double* ensure_T(fmmb_plan* plan, BlkBatch& B, int P, cudaStream_t s) {
  auto it = B.T.find(P);
  if (it != B.T.end()) return it->second->p;
  if (plan->capturing) throw StatusError{FMMB_ERR_INVALID, "translation matrices must exist before a graph capture"};
  DevBuf<double>* buf = new DevBuf<double>();
  B.T[P] = buf;
#define CALL(Q) build_T_t<Q>(plan, B, *buf, s)
  FMMB_FOR_P(P, CALL)
#undef CALL
  return buf->p;
}

}  // namespace


static void check_blk_batch(fmmb_plan* plan, const BlkBatch& B, const int* d_tgt, const int* d_src, int64_t n);

// Plan time: pairs (tgt[e], src[e]), e < n (device arrays) -> classes, blocks, items.  kind: 0 M2L, 1 M2M, 2 L2L.
void build_blk_batch(fmmb_plan* plan, BlkBatch& B, int kind, const int* d_tgt, const int* d_src, int64_t n) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  const int nb = T.nboxes;
  B.kind = kind;
  B.n_pairs = n; B.n_blocks = 0; B.n_items = 0; B.n_tiles = 0; B.n_classes = 0;
  B.level_blk_off.assign(T.nlevels + 1, 0);
  for (auto& kv : B.T) delete kv.second;
  B.T.clear();
  if (n <= 0) return;
  if (n >= (1ll << 31)) throw StatusError{FMMB_ERR_INVALID, "more than 2^31 translation pairs"};
  Temp tmp;
  auto scan_incl = [&](const int* in, int* out, int64_t cnt) {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, cnt, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceScan::InclusiveSum(t, bytes, in, out, cnt, s));
  };
  auto scan_excl = [&](const int* in, int* out, int64_t cnt) {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, cnt, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, cnt, s));
  };
  auto last_int = [&](const int* p, int64_t cnt) {
    int v = 0;
    FMMB_CUDA(cudaMemcpyAsync(&v, p + cnt - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    return v;
  };

  // 1. translation classes
  DevBuf<unsigned long long> k0, k1;
  DevBuf<int> i0, i1, head, scan, cid;
  k0.resize(n); k1.resize(n); i0.resize(n); i1.resize(n); head.resize(n); scan.resize(n); cid.resize(n);
  pair_class_keys<<<nblk(n, 256), 256, 0, s>>>(d_tgt, d_src, n, T.key.p, T.level.p, k0.p, i0.p);
  {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, i0.p, i1.p, n, 0, 36, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, k0.p, k1.p, i0.p, i1.p, n, 0, 36, s));
  }
  head_flags<<<nblk(n, 256), 256, 0, s>>>(k1.p, n, 0, head.p);
  scan_incl(head.p, scan.p, n);
  const int ncls = last_int(scan.p, n);
  if (ncls >= (1 << 16)) throw StatusError{FMMB_ERR_INVALID, "more than 65535 translation classes"};
  B.n_classes = ncls;
  B.class_vec.resize(ncls);
  scatter_class_ids<<<nblk(n, 256), 256, 0, s>>>(i1.p, head.p, scan.p, n, d_tgt, d_src, T.key.p, T.level.p,
                                                 make_double3(0.5 * T.cell[0], 0.5 * T.cell[1], 0.5 * T.cell[2]), cid.p,
                                                 B.class_vec.p);
  FMMB_CUDA(cudaGetLastError());

  // 2. blocks: targets of a level whose parents are consecutive (in the order of the parents that have targets)
  DevBuf<int> is_tgt, pflag, pr, box_blk, box_col;
  is_tgt.resize(nb); pflag.resize(nb + 1); pr.resize(nb + 1); box_blk.resize(nb); box_col.resize(nb);
  is_tgt.zero(s); pflag.zero(s);
  mark_targets_and_parents<<<nblk(n, 256), 256, 0, s>>>(d_tgt, n, T.parent.p, is_tgt.p, pflag.p);
  scan_excl(pflag.p, pr.p, nb + 1);
  std::vector<int> hpr = pr.to_host(s);
  int root_is_target = 0;
  FMMB_CUDA(cudaMemcpyAsync(&root_is_target, is_tgt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  FMMB_CUDA(cudaStreamSynchronize(s));
  if (T.nlevels > 15) throw StatusError{FMMB_ERR_TREE_DEPTH, "more than 15 levels"};
  LevelTable lt;
  std::memset(&lt, 0, sizeof lt);
  int nblocks = 0;
  for (int l = 0; l < T.nlevels; ++l) {
    B.level_blk_off[l] = nblocks;
    lt.blk_base[l] = nblocks;
    if (l == 0) { nblocks += root_is_target ? 1 : 0; continue; }
    const int plo = T.level_off[l - 1], phi = T.level_off[l];
    lt.pr_base[l] = hpr[plo];
    nblocks += (hpr[phi] - hpr[plo] + 15) / 16;
  }
  B.level_blk_off[T.nlevels] = nblocks;
  B.n_blocks = nblocks;
  if (nblocks >= (1 << 24)) throw StatusError{FMMB_ERR_INVALID, "too many translation blocks"};
  B.blk_cols.resize((size_t)nblocks * kCols);
  fill_int<<<nblk((int64_t)nblocks * kCols, 256), 256, 0, s>>>(B.blk_cols.p, (int64_t)nblocks * kCols, -1);
  place_targets<<<nblk(nb, 256), 256, 0, s>>>(is_tgt.p, nb, T.key.p, T.level.p, T.parent.p, pr.p, lt, box_blk.p,
                                              box_col.p, B.blk_cols.p);
  FMMB_CUDA(cudaGetLastError());

  // 3. pairs sorted by (block, class, column) -> items (block, class) with their active column tiles
  DevBuf<int> src_sorted, item_head, item_scan, tile_head, tile_scan;
  src_sorted.resize(n); item_head.resize(n); item_scan.resize(n); tile_head.resize(n); tile_scan.resize(n);
  pair_sort_keys<<<nblk(n, 256), 256, 0, s>>>(d_tgt, cid.p, box_blk.p, box_col.p, n, k0.p);
  {
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, d_src, src_sorted.p, n, 0, 51, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, k0.p, k1.p, d_src, src_sorted.p, n, 0, 51, s));
  }
  head_flags<<<nblk(n, 256), 256, 0, s>>>(k1.p, n, 7, item_head.p);
  head_flags<<<nblk(n, 256), 256, 0, s>>>(k1.p, n, 3, tile_head.p);
  scan_incl(item_head.p, item_scan.p, n);
  scan_incl(tile_head.p, tile_scan.p, n);
  const int n_items = last_int(item_scan.p, n), n_tiles = last_int(tile_scan.p, n);
  B.n_items = n_items; B.n_tiles = n_tiles;
  B.items.resize(n_items);
  B.items.zero(s);
  B.tile_src.resize((size_t)n_tiles * 8);
  fill_int<<<nblk((int64_t)n_tiles * 8, 256), 256, 0, s>>>(B.tile_src.p, (int64_t)n_tiles * 8, nb);   // nb = the zero row
  B.blk_item_off.resize(nblocks + 1);
  fill_int<<<nblk(nblocks + 1, 256), 256, 0, s>>>(B.blk_item_off.p, nblocks + 1, n_items);
  fill_items_tiles<<<nblk(n, 256), 256, 0, s>>>(k1.p, src_sorted.p, n, item_head.p, item_scan.p, tile_head.p,
                                                tile_scan.p, B.items.p, B.tile_src.p, B.blk_item_off.p);
  FMMB_CUDA(cudaGetLastError());

  // 4. launch order of a whole-batch launch: heavy blocks first
  {
    DevBuf<unsigned> w0, w1;
    DevBuf<int> id0;
    w0.resize(nblocks); w1.resize(nblocks); id0.resize(nblocks); B.blk_order.resize(nblocks);
    block_weights<<<nblk(nblocks, 256), 256, 0, s>>>(B.blk_item_off.p, B.items.p, nblocks, n_items, n_tiles, w0.p, id0.p);
    size_t bytes = 0;
    FMMB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, w0.p, w1.p, id0.p, B.blk_order.p, nblocks, 0, 32, s));
    void* t = tmp.get(bytes);
    FMMB_CUDA(cub::DeviceRadixSort::SortPairsDescending(t, bytes, w0.p, w1.p, id0.p, B.blk_order.p, nblocks, 0, 32, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
  }
  FMMB_CUDA(cudaStreamSynchronize(s));
  if (std::getenv("FMMB_SELF_CHECK")) check_blk_batch(plan, B, d_tgt, d_src, n);   // read per call: tests toggle it
}

// Host-side audit of a batch (FMMB_SELF_CHECK=1; compute-sanitizer is not available on every pool): every index the
// sweep kernel will dereference is inside its array, every pair of the input appears exactly once, masks and tile
// offsets are consistent.  Throws FMMB_ERR_INVALID with the first violation.
static void check_blk_batch(fmmb_plan* plan, const BlkBatch& B, const int* d_tgt, const int* d_src, int64_t n) {
  cudaStream_t s = plan->stream;
  const int nb = plan->tree.nboxes;
  auto fail = [&](const std::string& m) { throw StatusError{FMMB_ERR_INVALID, "blocked batch self-check: " + m}; };
  std::vector<int2> items = B.items.to_host(s);
  std::vector<int> tsrc = B.tile_src.to_host(s), off = B.blk_item_off.to_host(s), cols = B.blk_cols.to_host(s),
                   order = B.blk_order.to_host(s);
  std::vector<int> tg(n), sr(n);
  FMMB_CUDA(cudaMemcpy(tg.data(), d_tgt, n * sizeof(int), cudaMemcpyDeviceToHost));
  FMMB_CUDA(cudaMemcpy(sr.data(), d_src, n * sizeof(int), cudaMemcpyDeviceToHost));
  if ((int)off.size() != B.n_blocks + 1 || off[0] != 0 || off[B.n_blocks] != B.n_items) fail("block item offsets");
  if ((int64_t)tsrc.size() != B.n_tiles * 8 || (int)items.size() != B.n_items) fail("array sizes");
  std::vector<char> seen_blk(B.n_blocks, 0);
  for (int b : order) { if (b < 0 || b >= B.n_blocks || seen_blk[b]) fail("launch order is not a permutation"); seen_blk[b] = 1; }
  std::vector<long long> pair_key;
  pair_key.reserve(n);
  int64_t tile = 0;
  for (int b = 0; b < B.n_blocks; ++b) {
    if (off[b + 1] <= off[b]) fail("empty block");
    for (int c = 0; c < kCols; ++c) { const int t = cols[(size_t)b * kCols + c]; if (t < -1 || t >= nb) fail("column target out of range"); }
    for (int i = off[b]; i < off[b + 1]; ++i) {
      const unsigned mask = (unsigned)items[i].x >> 16;
      const int cls = items[i].x & 0xffff;
      if (cls >= B.n_classes || !mask) fail("item class / empty mask");
      if (items[i].y != tile) fail("tile offsets are not consecutive");
      int j = 0;
      for (int ct = 0; ct < 16; ++ct) {
        if (!((mask >> ct) & 1u)) continue;
        for (int k = 0; k < 8; ++k) {
          const int src = tsrc[(size_t)(tile + j) * 8 + k], t = cols[(size_t)b * kCols + ct * 8 + k];
          if (src < 0 || src > nb) fail("source box out of range");
          if (src < nb) {
            if (t < 0) fail("pair into an empty column");
            pair_key.push_back(((long long)t << 32) | (unsigned)src);
          }
        }
        ++j;
      }
      tile += j;
    }
  }
  if (tile != B.n_tiles) fail("tile count");
  std::vector<long long> want(n);
  for (int64_t e = 0; e < n; ++e) want[e] = ((long long)tg[e] << 32) | (unsigned)sr[e];
  std::sort(want.begin(), want.end());
  std::sort(pair_key.begin(), pair_key.end());
  if (want != pair_key) fail("the items do not hold exactly the input pairs");
}

// Plan time: the unit list of a sweep.  phases: (batch slot, first block, block count, heavy-first order) in
// dependency order; empty phases are dropped.
struct PhaseDesc { int slot, first, count; bool ordered; };
static void build_sweep(fmmb_plan* plan, Sweep& S, BlkBatch* b0, BlkBatch* b1, BlkBatch* b2, const int* modes,
                        const std::vector<PhaseDesc>& phases) {
  cudaStream_t s = plan->stream;
  int sms = 148;
  FMMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device));
  S.batch[0] = b0; S.batch[1] = b1; S.batch[2] = b2;
  for (int i = 0; i < 3; ++i) S.mode[i] = modes[i];
  std::vector<int4> units;
  std::vector<int> totals;
  int partials = 0;
  for (const PhaseDesc& ph : phases) {
    if (ph.count <= 0) continue;
    BlkBatch& B = *S.batch[ph.slot];
    // few blocks: split the item list of every block over several CTAs so that the phase fills the GPU
    int split = 1;
    while (split < 8 && ph.count * split * 2 <= sms) split *= 2;
    std::vector<int> order;
    if (ph.ordered) order = B.blk_order.to_host(s);
    const int phase = (int)totals.size();
    for (int i = 0; i < ph.count; ++i) {
      const int blk = ph.ordered ? order[ph.first + i] : ph.first + i;
      for (int cs = 0; cs < split; ++cs)
        units.push_back(make_int4(ph.slot | (phase << 4), blk, cs | (split << 8), split > 1 ? partials : -1));
      if (split > 1) partials += split;
    }
    totals.push_back(ph.count);
  }
  S.n_units = (int)units.size();
  S.n_phases = (int)totals.size();
  S.n_partials = partials;
  S.units.from_host(units.data(), units.size(), s);
  S.phase_total.from_host(totals.data(), totals.size(), s);
  S.phase_cnt.resize(std::max<size_t>(1, totals.size()));
  S.phase_cnt.zero(s);
  FMMB_CUDA(cudaStreamSynchronize(s));
  S.built = true;
}

template <int P>
static void launch_sweep_t(fmmb_plan* plan, Sweep& S, cudaStream_t s) {
  using C = Cfg<P>;
  SweepArgs a;
  std::memset(&a, 0, sizeof a);
  for (int i = 0; i < 3; ++i) {
    BlkBatch* B = S.batch[i];
    if (!B || B->n_blocks == 0) continue;
    BlkArgs& b = a.b[i];
    b.T = ensure_T(plan, *B, P, s);
    b.items = B->items.p; b.tile_src = B->tile_src.p; b.blk_item_off = B->blk_item_off.p; b.blk_cols = B->blk_cols.p;
    // mode 0: M -> M (M2M), 1: M -> L (M2L), 2: L += (L2L)
    b.X = S.mode[i] == 2 ? plan->L.p : plan->M.p;
    b.Out = S.mode[i] == 0 ? plan->M.p : plan->L.p;
    b.add = S.mode[i] == 2 ? 1 : 0;
  }
  if (S.n_partials > 0) {
    const size_t need = (size_t)S.n_partials * C::RBLK * kPartial;
    if (S.scratch.n < need || S.split_cnt.n < (size_t)S.n_partials * C::RBLK) {
      if (plan->capturing) throw StatusError{FMMB_ERR_INVALID, "split scratch must exist before a graph capture"};
      S.scratch.resize(need);
      S.split_cnt.resize((size_t)S.n_partials * C::RBLK);
      S.split_cnt.zero(s);
    }
  }
  a.units = S.units.p; a.phase_cnt = S.phase_cnt.p; a.phase_total = S.phase_total.p;
  a.scratch = S.scratch.p; a.split_cnt = S.split_cnt.p;
  FMMB_CUDA(cudaFuncSetAttribute(trans_sweep_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  FMMB_CUDA(cudaMemsetAsync(S.phase_cnt.p, 0, (size_t)S.n_phases * sizeof(unsigned), s));
  static const bool tracing = std::getenv("FMMB_TRACE") != nullptr;
  static long long* trace_buf = nullptr;
  const unsigned grid = (unsigned)S.n_units * C::RBLK;
  if (tracing && !plan->capturing) {
    if (!trace_buf) FMMB_CUDA(cudaMalloc(&trace_buf, 2048 * 8 * sizeof(long long)));
    FMMB_CUDA(cudaMemsetAsync(trace_buf, 0, 2048 * 8 * sizeof(long long), s));
    a.trace = trace_buf;
  }
  trans_sweep_kernel<P><<<grid, kThreads, C::SMEM, s>>>(a);
  FMMB_CUDA(cudaGetLastError());
  if (a.trace) {
    const size_t nc = std::min<size_t>(grid, 2048);
    std::vector<long long> h(nc * 8);
    std::vector<int4> hu = S.units.to_host(s);
    FMMB_CUDA(cudaMemcpy(h.data(), trace_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long t0 = h[0];
    for (size_t i = 0; i < h.size(); i += 8) if (h[i] && h[i] < t0) t0 = h[i];
    // per phase: first start, last dependency-wait end, last end
    std::vector<long long> first(S.n_phases, -1), waited(S.n_phases, 0), last(S.n_phases, 0);
    std::vector<int> cnt(S.n_phases, 0);
    for (size_t c = 0; c < nc; ++c) {
      const int ph = hu[c / C::RBLK].x >> 4;
      const long long b0 = h[c * 8] - t0, b1 = h[c * 8 + 1] - t0, b6 = h[c * 8 + 6] - t0;
      if (first[ph] < 0 || b0 < first[ph]) first[ph] = b0;
      waited[ph] = std::max(waited[ph], b1);
      last[ph] = std::max(last[ph], b6);
      ++cnt[ph];
    }
    fprintf(stderr, "TRACE sweep units %d phases %d:", S.n_units, S.n_phases);
    for (int ph = 0; ph < S.n_phases; ++ph)
      fprintf(stderr, " [ph %d n %d start %lld go %lld end %lld]", ph, cnt[ph], first[ph], waited[ph], last[ph]);
    fprintf(stderr, "\n");
  }
}

// One launch: every phase of the sweep at the plan's current order on plan->M / plan->L.
void run_sweep(fmmb_plan* plan, Sweep& S, cudaStream_t s) {
  if (S.n_units == 0) return;
  const int P = plan->p;
#define CALL(Q) launch_sweep_t<Q>(plan, S, s)
  FMMB_FOR_P(P, CALL)
#undef CALL
  ++plan->launches;
}

// The sweeps of a plan (built on first use, never inside a graph capture):
//   0 all:   M2M level by level, M2L, L2L level by level          (single GPU, or replicated upward pass)
//   1 up:    M2M level by level                                   (treecode)
//   2 own:   M2M of the parents inside this rank's range          (multi-GPU, before the multipole exchange)
//   3 rest:  M2M of the parents that straddle a cut, M2L, L2L     (multi-GPU, after the exchange)
//   4 strad: M2M of the straddling parents                        (multi-GPU treecode)
Sweep& plan_sweep(fmmb_plan* plan, int which) {
  Sweep& S = plan->sweeps[which];
  if (S.built) return S;
  if (plan->capturing) throw StatusError{FMMB_ERR_INVALID, "sweeps must exist before a graph capture"};
  Tree& T = plan->tree;
  std::vector<PhaseDesc> ph;
  const int modes[3] = {0, 1, 2};
  BlkBatch& up = which == 0 || which == 1 ? plan->b_m2m : (which == 2 ? plan->b_m2m_own : plan->b_m2m_strad);
  if (up.level_blk_off.size() == (size_t)T.nlevels + 1)
    for (int l = T.nlevels - 2; l >= 0; --l)
      ph.push_back(PhaseDesc{0, up.level_blk_off[l], up.level_blk_off[l + 1] - up.level_blk_off[l], false});
  if (which == 0 || which == 3) {
    ph.push_back(PhaseDesc{1, 0, plan->b_m2l.n_blocks, true});
    BlkBatch& dn = plan->b_l2l;
    if (dn.level_blk_off.size() == (size_t)T.nlevels + 1)
      for (int l = 1; l < T.nlevels; ++l)
        ph.push_back(PhaseDesc{2, dn.level_blk_off[l], dn.level_blk_off[l + 1] - dn.level_blk_off[l], false});
  }
  build_sweep(plan, S, &up, &plan->b_m2l, &plan->b_l2l, modes, ph);
  return S;
}

namespace {
__global__ void csr_targets(const int* __restrict__ off, int nb, int* __restrict__ tgt) {
  const int b = blockIdx.x;
  if (b >= nb) return;
  for (int e = off[b] + threadIdx.x; e < off[b + 1]; e += blockDim.x) tgt[e] = b;
}
}  // namespace

void blocked_init_tables() { upload_laplace_tables(); }

// Plan time: the far-field batches of this plan (this rank's share on a multi-GPU plan).
void build_blocked_batches(fmmb_plan* plan) {
  Tree& T = plan->tree;
  cudaStream_t s = plan->stream;
  const int nb = T.nboxes;
  if (T.n_lr >= (1ll << 31)) throw StatusError{FMMB_ERR_INVALID, "more than 2^31 M2L pairs"};
  // M2L: the target-major list restricted to the boxes that are targets on this rank
  if (T.n_lr_local > 0) {
    DevBuf<int> tgt;
    tgt.resize(T.n_lr_local);
    csr_targets<<<nb, 64, 0, s>>>(T.m2l_off.p, nb, tgt.p);
    FMMB_CUDA(cudaGetLastError());
    build_blk_batch(plan, plan->b_m2l, 0, tgt.p, T.m2l_src.p, T.n_lr_local);
  } else {
    build_blk_batch(plan, plan->b_m2l, 0, nullptr, nullptr, 0);
  }
  if (nb <= 1) return;
  std::vector<unsigned> par = T.parent.to_host(s);
  std::vector<unsigned char> hl = T.has_local.to_host(s), act = T.active.to_host(s);
  auto build = [&](BlkBatch& B, int kind, const std::vector<int>& tg, const std::vector<int>& sr) {
    DevBuf<int> dt, ds;
    dt.from_host(tg.data(), tg.size(), s);
    ds.from_host(sr.data(), sr.size(), s);
    build_blk_batch(plan, B, kind, dt.p, ds.p, (int64_t)tg.size());
    FMMB_CUDA(cudaStreamSynchronize(s));
  };
  std::vector<int> tg, sr;
  // M2M: child -> parent, only into parents whose multipole a matvec reads (M2L sources and what lies below them)
  const std::vector<unsigned char>& need = T.need_M_host;
  tg.clear(); sr.clear();
  for (int c = 1; c < nb; ++c) if (need[par[c]]) { tg.push_back((int)par[c]); sr.push_back(c); }
  build(plan->b_m2m, 1, tg, sr);
  if (T.nranks > 1) {
    std::vector<int> tg2, sr2;
    tg.clear(); sr.clear();
    for (int c = 1; c < nb; ++c) {
      if (!need[par[c]]) continue;
      const int o = T.box_owner[par[c]];
      if (o == T.rank) { tg.push_back((int)par[c]); sr.push_back(c); }           // parents inside my range
      else if (o < 0) { tg2.push_back((int)par[c]); sr2.push_back(c); }          // parents that straddle a cut
    }
    build(plan->b_m2m_own, 1, tg, sr);
    build(plan->b_m2m_strad, 1, tg2, sr2);
  }
  // L2L: parent -> child, into the boxes that are targets on this rank and whose parent carries a local expansion
  tg.clear(); sr.clear();
  for (int c = 1; c < nb; ++c)
    if (act[c] && hl[par[c]]) { tg.push_back(c); sr.push_back((int)par[c]); }
  build(plan->b_l2l, 2, tg, sr);
}

}  // namespace fmmb
