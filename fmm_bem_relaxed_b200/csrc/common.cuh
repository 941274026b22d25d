// csrc/common.cuh -- shared plumbing for the B200 FMM engine (device buffers, error handling,
// the plan object).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <map>
#include <functional>
#include <nvtx3/nvToolsExt.h>
#include "../../include/fmmb.h"

namespace fmmb {

void set_error(const std::string& msg);

struct CudaError {
  cudaError_t err;
  const char* what;
  const char* file;
  int line;
};

#define FMMB_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) throw ::fmmb::CudaError{e_, #call, __FILE__, __LINE__}; \
  } while (0)

// NVTX range for the host-side span of a plan-time step or of a matvec's launch sequence (timeline tools)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct StatusError {
  int status;
  std::string msg;
};

// Captured CUDA graphs hold raw device pointers.  Every DevBuf that FREES an allocation while a plan's matvec is
// being issued bumps that plan's counter (run_matvec points this thread-local at it); a changed counter tells
// run_matvec that graphs captured earlier may replay freed pointers, and they are dropped.
inline thread_local unsigned long long* g_realloc_counter = nullptr;
inline void note_realloc() { if (g_realloc_counter) ++*g_realloc_counter; }

// Plain device buffer.  resize() discards contents; grow() preserves them.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;   // elements allocated
  size_t n = 0;     // elements in use
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) { cudaFree(p); note_realloc(); }
    p = nullptr; cap = n = 0;
  }
  void resize(size_t count) {
    if (count > cap) {
      if (p) { cudaFree(p); note_realloc(); }
      p = nullptr;
      FMMB_CUDA(cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T)));
      cap = count ? count : 1;
    }
    n = count;
  }
  void grow(size_t count, cudaStream_t s) {
    if (count > cap) {
      size_t ncap = cap * 2 > count ? cap * 2 : count;
      T* q = nullptr;
      FMMB_CUDA(cudaMalloc((void**)&q, ncap * sizeof(T)));
      if (p && n) FMMB_CUDA(cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, s));
      if (p) { FMMB_CUDA(cudaStreamSynchronize(s)); cudaFree(p); note_realloc(); }
      p = q; cap = ncap;
    }
    n = count;
  }
  void zero(cudaStream_t s) { if (n) FMMB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  std::vector<T> to_host(cudaStream_t s) const {
    std::vector<T> h(n);
    if (n) {
      FMMB_CUDA(cudaMemcpyAsync(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
      FMMB_CUDA(cudaStreamSynchronize(s));
    }
    return h;
  }
  void from_host(const T* h, size_t count, cudaStream_t s) {
    resize(count);
    if (count) FMMB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};

// Octree + interaction lists, device resident (built by tree.cu).
struct Tree {
  int64_t n = 0;
  unsigned ncrit = 64;
  double theta = 0.5;
  double pmin[3], cell[3];
  int nboxes = 0, nlevels = 0, nleaves = 0;
  std::vector<int> level_off;        // host copy: boxes of level l are [level_off[l], level_off[l+1])

  DevBuf<double> pts_orig;           // the caller's points, original order (3n)
  DevBuf<unsigned> perm;             // tree index -> original index
  DevBuf<unsigned> iota;             // identity permutation (sharded calls: charges arrive in tree order)
  DevBuf<unsigned> code;             // Morton code, tree order
  DevBuf<double4> body;              // tree order: x, y, z, (charge slot written per matvec)
  // box table (SoA)
  DevBuf<unsigned> key, parent, cbegin, cend, bbegin, bend, level;
  DevBuf<double4> center;            // cx, cy, cz, side
  DevBuf<int> leaves;                // indices of leaf boxes, ascending
  // multi-GPU partition (rank owns the tree-order bodies [own_b0, own_b1), whole leaves)
  int rank = 0, nranks = 1;
  int64_t own_b0 = 0, own_b1 = 0;
  std::vector<int64_t> body_cuts;    // nranks + 1 body offsets, identical on every rank
  DevBuf<int> own_leaves;            // leaf boxes inside the owned range, ascending
  DevBuf<int> own_leaves_body;       // the same leaves in body order (consecutive leaves, consecutive body ranges)
  int n_own_leaves = 0;
  DevBuf<unsigned char> active;      // box intersects the owned range (is a target on this rank)
  // owned upward pass (used once a communicator exists): a box is "inside" a rank when all its bodies
  // belong to that rank; boxes that straddle a cut are recomputed by every rank after the exchange
  DevBuf<unsigned char> up_inside;   // inside MY range
  std::vector<int> box_owner;        // host: rank whose range holds the whole box, -1 = straddles a cut (all ranks agree)
  bool owned_upward = false;         // some rank owns a parent box: every rank runs the owned upward pass + exchange
  // straddling (non-leaf) boxes and, for each, its maximal descendants that lie inside one rank:
  // M[straddler] = sum of direct (multi-level) M2M translations of those descendants
  DevBuf<int> strad_box, strad_off, strad_desc, strad_pair_box;
  int n_strad = 0, n_strad_pairs = 0;
  DevBuf<double> strad_tmp;
  DevBuf<int> xchg_list;             // boxes inside rank q, concatenated rank-major
  std::vector<int> xchg_off;         // nranks + 1 offsets into xchg_list
  int xchg_max = 0;                  // largest per-rank box count (all-gather chunk)
  DevBuf<double> xchg_send, xchg_recv;
  int64_t n_lr_local = 0;            // M2L pairs whose target is active here (= n_lr on one GPU)
  DevBuf<unsigned char> has_local;   // box carries a local expansion (M2L target or descendant of one)
  DevBuf<unsigned char> need_M;      // box's multipole is read by a matvec (M2L source, or below one)
  std::vector<unsigned char> need_M_host;
  DevBuf<unsigned char> m2m_mask_all, m2m_mask_own;   // parents the batched M2M writes: needed [and inside my range]
  // M2L: reference-order pair list and target-major CSR (sources in list order per target)
  DevBuf<int2> lr;                   // (source, target) in LR_list order
  DevBuf<int> m2l_off, m2l_src;
  int64_t n_lr = 0;
  // P2P: target-major CSR of source leaf boxes in P2P_lists order
  DevBuf<int> p2p_off, p2p_src;
  int64_t n_p2p = 0, n_p2p_body_pairs = 0;
  DevBuf<int4> p2p_items;            // (target leaf, first target body, #targets <= 32, 0), leaf order
  int n_p2p_items = 0;
  DevBuf<int> p2p_run_off;           // per target box: its merged source body runs (laplace.cu: p2p_run_kernel)
  DevBuf<int2> p2p_runs;             // [begin, end) in tree-ordered bodies
  int n_p2p_runs = 0;
  DevBuf<int4> p2p_items_ext;        // per item: run range, own body range, close flag (p2p_pair2_kernel)
  DevBuf<unsigned char> p2p_close;   // per box: a source of another leaf lies within 1e-4 of one of its bodies
};

// One family of box-to-box translations (M2L, M2M or L2L) evaluated as class-batched GEMMs
// (built by m2l_classes.cu).  "slot" = position of a pair in the family's slot space: for M2L the
// index into the target-major CSR (Tree::m2l_src); for M2M/L2L the child box index.
struct TransBatch {
  int kind = 0;                      // 0 = M2L, 1 = M2M, 2 = L2L
  int64_t n_classes = 0;             // distinct translation vectors
  int64_t n_pairs = 0;               // pairs covered by the batched path
  int64_t n_res = 0;                 // M2L pairs left to the per-pair kernel
  int n_items = 0;                   // (class, <=128 pairs) GEMM tiles
  int built_p = 0;                   // order the matrices were built for (0 = none)
  DevBuf<double> T;                  // [class][k][row], leading dimension built_p^2
  DevBuf<double4> class_vec;         // representative translation vector per class
  DevBuf<int> slot_tgt, slot_src;    // target / source box of each slot
  const int* slot_src_p = nullptr;   // = slot_src.p, or Tree::m2l_src.p for M2L
  DevBuf<int> sorted_slot;           // slots ordered by class (stable: slot order inside a class)
  DevBuf<int> item_class, item_start, item_count;
  DevBuf<int4> item_desc;            // the same three per item in one word: x = class, y = start, z = count
  DevBuf<int> sorted_src;            // source box of sorted_slot[i] (one load instead of two dependent ones)
  std::vector<int> level_item_off;   // M2M/L2L: items whose target level is l
  DevBuf<unsigned char> batched;     // M2L, per slot: handled by the batched path
  DevBuf<int> res_off, res_src;      // M2L residual pairs, target-major CSR in list order
  DevBuf<int> res_boxes;             // target boxes that have residual pairs
  int n_res_boxes = 0;
  DevBuf<double> tmp;                // phase-1 output columns, [slot][p^2]
};

// One family of box-to-box translations evaluated output-stationary (trans_blocked.cu): targets of a level in
// blocks of <= 128 columns, per block the list of (class, active column tiles) items.
struct BlkBatch {
  int kind = 0;                      // 0 = M2L, 1 = M2M, 2 = L2L
  int n_blocks = 0, n_items = 0;
  int64_t n_pairs = 0, n_tiles = 0, n_classes = 0;
  DevBuf<int2> items;                // x = class | (tile mask << 16), y = first tile of the item
  DevBuf<int> tile_src;              // 8 source boxes per active tile (nboxes = the all-zero expansion)
  DevBuf<int> blk_item_off;          // n_blocks + 1
  DevBuf<int> blk_cols;              // n_blocks x 128: target box of a column, -1 = none
  DevBuf<int> blk_order;             // all blocks, heaviest first
  std::vector<int> level_blk_off;    // blocks whose targets are at level l: [l], [l + 1])
  DevBuf<double4> class_vec;         // translation vector of a class (target centre - source centre)
  std::map<int, DevBuf<double>*> T;  // per order: fragment-major translation matrices
  BlkBatch() {}
  BlkBatch(const BlkBatch&) = delete;
  BlkBatch& operator=(const BlkBatch&) = delete;
  ~BlkBatch() { for (auto& kv : T) delete kv.second; }
};

// One launch of trans_blocked.cu: work units of up to three batches in dependent phases.
struct Sweep {
  bool built = false;
  int n_units = 0, n_phases = 0, n_partials = 0;
  BlkBatch* batch[3] = {nullptr, nullptr, nullptr};
  int mode[3] = {0, 1, 2};           // per batch slot: 0 = M -> M, 1 = M -> L, 2 = L += L
  DevBuf<int4> units;
  DevBuf<int> phase_total;
  DevBuf<unsigned> phase_cnt, split_cnt;
  DevBuf<double> scratch;
};

struct BemData;
struct StokesData;
struct StokesBemData;
struct YukawaData;
struct GmresWorkspace;

struct LaplaceTables {
  int pmax = 0;
  DevBuf<double> pref;               // sqrt((n-|m|)!/(n+|m|)!), index n^2+n+m, n < 2*pmax
  DevBuf<double> anm;                // (-1)^n / sqrt((n-m)!(n+m)!)
};

}  // namespace fmmb

struct fmmb_plan {
  int device = 0;
  int kind = 0;
  int p = 5;
  int p_alloc = 0;                   // order the expansion buffers are sized for
  fmmb_options opts;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev[16];
  fmmb::Tree tree;
  fmmb::LaplaceTables tab;
  fmmb::TransBatch cls, m2m, l2l;    // batched M2L / M2M / L2L
  fmmb::TransBatch m2m_own;          // multi-GPU: M2M restricted to parents inside this rank's range
  // output-stationary translations (trans_blocked.cu): M2L; M2M of all / of the owned parents / of the parents
  // that straddle a partition cut; L2L
  fmmb::BlkBatch b_m2l, b_m2m, b_m2m_own, b_m2m_strad, b_l2l;
  fmmb::Sweep sweeps[5];             // plan_sweep(): all / up / own / rest / strad
  std::map<int, fmmb::DevBuf<double>*> m2l_coeff;  // per-order real M2L coefficient tables
  fmmb::DevBuf<double> M, L;         // box-major, real layout (laplace_ops.cuh), stride xstride(p)
  fmmb::DevBuf<double> charges;      // original order staging
  fmmb::DevBuf<double4> res_near, res_far;  // tree order
  fmmb::DevBuf<double4> res_tree;    // multi-GPU: near + far in tree order, all-gathered over NCCL
  fmmb::BemData* bem = nullptr;      // LaplaceSphericalBEM plans only
  fmmb::StokesData* stokes = nullptr;  // StokesSpherical plans only
  fmmb::StokesBemData* sbem = nullptr; // StokesSphericalBEM plans only
  fmmb::YukawaData* yukawa = nullptr;  // YukawaCartesian plans only
  fmmb::GmresWorkspace* gmres_ws = nullptr;  // fmmb_gmres scratch, kept between solves
  int charge_dim = 1, result_dim = 4;
  void* comm = nullptr;              // ncclComm_t once fmmb_plan_comm_init ran
  fmmb::DevBuf<int> xchg_off_dev;
  fmmb::DevBuf<double4> res_stage;   // padded all-gather staging for the result slices
  fmmb::DevBuf<double> gen_tree, gen_stage;  // finish_results(): tree-ordered results and the all-gather staging
  fmmb::DevBuf<long long> cuts_dev;
  fmmb::DevBuf<double> chg_stage, chg_send;  // sharded call: padded all-gather of the charge slices
  fmmb::DevBuf<double> q_tree;               // sharded call of the gather-by-permutation classes: all charges, tree order
  const double* sharded_q = nullptr;         // ... the vector their gather kernels read during the current call
  long long chg_chunk = 0;
  // multipole exchange through peer memory (comm.cu): the M allocation is exported once and never reallocated
  bool peer_alloc = false, peer_ready = false;
  fmmb::DevBuf<double*> peer_M;                       // per rank: base of its multipole array (own: local pointer)
  fmmb::DevBuf<unsigned long long*> peer_flags;       // per rank: its flag tail
  fmmb::DevBuf<unsigned long long> peer_state;        // [0] matvec counter, [1] push-kernel block counter
  std::vector<void*> peer_opened;
  unsigned long long* peer_flag_host = nullptr;       // pinned: timeout flag of the bounded flag waits, fetched per call
  std::function<void()> hook_after_owned_m2m;  // set by laplace_execute around laplace_translations
  std::function<void()> hook_after_m2l_gemm;   // set by laplace_execute: start the near field behind the M2L GEMM
  bool call_sharded = false;         // the current call is fmmb_plan_execute_sharded
  bool cuts_ready = false;
  bool xchg_off_ready = false;
  fmmb::DevBuf<double> results;      // original order staging, 4n
  fmmb::DevBuf<double> own_q, own_r; // fmmb_plan_execute_sharded_host: device staging of this rank's slices
  double phase_ms[FMMB_T_COUNT] = {0};
  bool timed = false;
  bool m2l_gemm_timed = false;
  bool graph_timed = false;
  bool overlap_p2p = true;
  int p2p_item_mode = 0;             // see p2p_fill_items (laplace.cu); BEM plans use 0
  int p2p_warps = 1;                 // warps per block of the near-field pair kernels
  int p2p_kernel = 2;                // 1 = merged source runs with prefetch (p2p_run_kernel), 0 = per source leaf
  int p2p_unroll = 4;
  int p2p_wps = 0;                   // > 0: persistent near-field blocks, that many one-warp blocks per SM (0 = plain grid).
                                     // Measured at N = 1M, P = 8: 8 per SM hides the whole far field under the near
                                     // field (3.31 ms) but the near field itself is then latency-bound per warp, so the
                                     // matvec is no faster than the plain grid (3.32 ms); fewer or more are slower.
  fmmb::DevBuf<unsigned> p2p_counter;
  int p2p_occ = 28;                  // resident one-warp blocks per SM the pair kernel is compiled for: 20 (96 registers,
                                     // no spill), 24, 28 (72 registers, 12 bytes of spill: fastest, 1.29 vs 1.35 ms at
                                     // N = 1M) or 32; same bits in every variant
  int bem_near_kernel = 1;           // cached BEM near field: 1 = eight warps per work item (bem_near_split_kernel), 0 = one
  int l2p_kernel = 1;                // 1 = four leaves per warp in body order (l2p_packed_kernel), 0 = one leaf per warp
  int p2m_kernel = 1;                // 1 = narrow transposition tile (p2m_cols_kernel), 0 = full tile (p2m_kernel)
  int p2p_order = 0;                 // one GPU, class-major engine.  0 = the near field starts with the upward pass and
                                     // fills whatever the far-field chain (higher stream priority) leaves free;
                                     // 1 = it starts when the M2L GEMM has finished (measured at N = 1M: 3.09 vs 2.84 ms:
                                     // the FP64 pipe idles under the upward pass)
  int m2l_reduce = 1;                // 1 = column reduction staged through shared memory by TMA bulk copies, m2l_reduce_bps
                                     // blocks per SM (leaves block slots, threads and registers to the near field);
                                     // 0 = a block per box with its loads in flight in registers
  int m2l_reduce_bps = 2;
  int p2p_defer = 1;                 // sharded plans with an owned upward pass: 1 = the near field starts behind the owned
                                     // M2M sweep (beside the multipole exchange), 0 = with the upward pass
  int p2p_newton = 0;                // 1 = Newton-only inverse root in the near-field pair kernel (p2p_kernel 3)
  int near_only = 0;                 // fmmb_options.near_only
  bool far_built_classes = false, far_built_blocked = false;   // which far-field structures exist (laplace_build_far)
  int p2p_chunk = 32, p2p_min_chunk = 8;  // targets per near-field work item (chosen at plan time)
  // CUDA graphs: one captured matvec per (order, charge pointer, result pointer)
  bool use_graph = true;
  bool graph_node_priority = true;   // instantiate graphs with cudaGraphInstantiateFlagUseNodePriority
  bool capturing = false;
  struct GraphKey {
    int p; const void* q; void* r; int mode;
    bool operator<(const GraphKey& o) const {
      if (p != o.p) return p < o.p;
      if (mode != o.mode) return mode < o.mode;
      if (q != o.q) return q < o.q;
      return r < o.r;
    }
  };
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::map<GraphKey, int> graph_seen;
  unsigned long long realloc_count = 0;   // device buffers freed while this plan's matvecs were issued (common.cuh)
  unsigned long long graphs_valid_at = 0; // value of realloc_count the cached graphs were captured under
  int launches = 0;                  // kernel launches of the last execute
};

namespace fmmb {
// tree.cu
void build_tree(fmmb_plan* plan, const double* points_host, int64_t n);
void restrict_p2p_to_self(fmmb_plan* plan);
void partition_ranges(const double* w, int64_t n, int nranks, int64_t* cuts);
void comm_unique_id(unsigned char* id);
void comm_init(fmmb_plan* plan, const unsigned char* id);
void comm_destroy(fmmb_plan* plan);
void allgather_results(fmmb_plan* plan, cudaStream_t s);
void finish_results(fmmb_plan* plan, const double* near, const double* far, int rd, double* d_results, cudaStream_t s);
void allgather_charges(fmmb_plan* plan, const double* d_own, cudaStream_t s);
const double* sharded_assemble_charges(fmmb_plan* plan, const double* d_own, cudaStream_t s);
// what a kernel class's charge-gather kernel reads: the caller's vector through the tree permutation, or -- in a
// sharded call -- the assembled tree-ordered vector through the identity
inline const unsigned* exec_perm(const fmmb_plan* plan) { return plan->call_sharded ? plan->tree.iota.p : plan->tree.perm.p; }
inline const double* exec_charges(const fmmb_plan* plan, const double* d_charges) {
  return plan->call_sharded ? plan->sharded_q : d_charges;
}
void exchange_multipoles(fmmb_plan* plan, cudaStream_t s);
void peer_export(fmmb_plan* plan, unsigned char* blob);
void peer_init(fmmb_plan* plan, const unsigned char* blobs);
void peer_close(fmmb_plan* plan);
void exchange_multipoles_peer(fmmb_plan* plan, cudaStream_t s);
void peer_exchange_charges(fmmb_plan* plan, const double* d_own, cudaStream_t s);
void peer_read_done(fmmb_plan* plan, cudaStream_t s);
void peer_flag_fetch(fmmb_plan* plan, cudaStream_t s);
void peer_check_timeout(fmmb_plan* plan);
// laplace.cu
void laplace_init_tables(fmmb_plan* plan);
void laplace_execute(fmmb_plan* plan, const double* d_charges, double* d_results);
void build_p2p_items(fmmb_plan* plan);
void laplace_translations(fmmb_plan* plan, cudaStream_t s);
void laplace_build_far(fmmb_plan* plan);
bool laplace_owned_upward(const fmmb_plan* plan);
void laplace_prepare_expansions(fmmb_plan* plan);
void stokes_prepare_expansions(fmmb_plan* plan);
void stokes_bem_prepare_expansions(fmmb_plan* plan);
// bem.cu
void bem_setup(fmmb_plan* plan, const double* verts_host, const int32_t* bc_host, int quad_k, double kappa = -1.0);
void bem_begin(fmmb_plan* plan, const double* d_charges, cudaStream_t s);
bool bem_set_active(const BemData* b, int set);
double* bem_res_near(BemData* b);
double* bem_res_far(BemData* b);
int bem_rule_points(const BemData* b);
void yukawa_bem_execute(fmmb_plan* plan, const double* d_charges, double* d_results);
void bem_execute(fmmb_plan* plan, const double* d_charges, double* d_results);
void bem_free(BemData* b);
void bem_direct(fmmb_plan* plan, const double* d_charges, int64_t nt, const double* d_tverts, const int* d_tbc,
                double* d_out, cudaStream_t s);
int64_t bem_nnz(const BemData* b);
void laplace_direct_raw(const double* d_spts, const double* d_q, int64_t ns, const double* d_tpts, int64_t nt,
                        double* d_out, cudaStream_t s);
void measure_fp64_peak(double* dfma, double* dmma);
// yukawa.cu
void yukawa_setup(fmmb_plan* plan, double kappa);
void yukawa_execute(fmmb_plan* plan, const double* d_charges, double* d_results);
void yukawa_free(YukawaData* d);
double yukawa_kappa(const YukawaData* d);
void yukawa_direct_raw(double kappa, const double* d_spts, const double* d_q, int64_t ns, const double* d_tpts,
                       int64_t nt, double* d_out, cudaStream_t s);
// gmres.cu
void gmres_free(GmresWorkspace* w);
void fgmres_solve(fmmb_plan* plan, fmmb_plan* pc_plan, const fmmb_solver_options* pc_opts, const double* b_host,
                  double* x_host, const fmmb_solver_options& o, fmmb_gmres_info* info, int32_t* p_sched,
                  double* res_hist, int cap);
void gmres_reserve(fmmb_plan* plan, double** z, double** w);   // solver workspace at plan construction (warm start)
void gmres_solve(fmmb_plan* plan, const double* b_host, double* x_host, const double* diag_host,
                 const fmmb_solver_options& o, fmmb_gmres_info* info, int32_t* p_sched, double* res_hist, int cap);
// stokes.cu
void stokes_setup(fmmb_plan* plan, bool stresslet);
void stokes_execute(fmmb_plan* plan, const double* d_charges, double* d_results);
void stokes_free(StokesData* d);
bool stokes_is_stresslet(const StokesData* d);
void stokes_direct_raw(bool stresslet, const double* d_spts, const double* d_q, int64_t ns, const double* d_tpts,
                       int64_t nt, double* d_out, cudaStream_t s);
// stokes_bem.cu
void stokes_bem_setup(fmmb_plan* plan, const double* verts_host, const int32_t* bc_host, int quad_k, int quad_kfine,
                      double mu);
void stokes_bem_execute(fmmb_plan* plan, const double* d_charges, double* d_results);
void stokes_bem_free(StokesBemData* d);
void stokes_bem_direct(fmmb_plan* plan, const double* d_charges, int64_t nt, const double* d_tverts, const int* d_tbc,
                       double* d_out, cudaStream_t s);
int64_t stokes_bem_nnz(const StokesBemData* d);
// trans_blocked.cu
void blocked_init_tables();
void build_blk_batch(fmmb_plan* plan, BlkBatch& B, int kind, const int* d_tgt, const int* d_src, int64_t n);
Sweep& plan_sweep(fmmb_plan* plan, int which);
void run_sweep(fmmb_plan* plan, Sweep& S, cudaStream_t s);
void build_blocked_batches(fmmb_plan* plan);
// m2l_classes.cu
void m2l_init_tables();
void build_m2l_classes(fmmb_plan* plan);
bool m2l_batched(fmmb_plan* plan, cudaStream_t s);
bool m2m_batched(fmmb_plan* plan, cudaStream_t s, bool owned_only = false);
bool l2l_batched(fmmb_plan* plan, cudaStream_t s);
}  // namespace fmmb
