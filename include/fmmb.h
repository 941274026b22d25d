/* include/fmmb.h -- C ABI of the B200 FMM matvec engine (libfmmb200.so).
 *
 * This is the drop-in boundary for the hot path of barbagroup/fmm-bem-relaxed:
 *     FMM_plan<Kernel>::FMM_plan(K, sources, opts)       reference include/FMM_plan.hpp:34-43
 *     FMM_plan<Kernel>::execute(charges) -> results       reference include/FMM_plan.hpp:75-90
 *     plan.kernel().set_p(p)                              reference kernel/LaplaceSpherical.hpp:119-128,
 *                                                         called per GMRES iteration, examples/BEM/GMRES.hpp:195-196
 * Everything below that boundary in the reference (include/executor/ *, include/tree/ *,
 * kernel/ *.hpp operator bodies) is replaced by CUDA kernels for sm_100a behind these
 * entry points.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative fmmb_status; fmmb_last_error()
 *     returns a thread-local message for the last failure on the calling thread.
 *   - "host" pointers are ordinary host memory (pinned memory makes copies faster),
 *     "device" pointers are CUDA device memory on the plan's device.
 *   - charges arrive and results leave in the caller's ORIGINAL body order, like
 *     std::vector<charge_type> / std::vector<result_type> in the reference
 *     (reference include/executor/ExecutorSingleTree.hpp:81-90).
 *   - one plan = one CUDA stream; calls on one plan must be serialised by the caller
 *     (the reference's plan is not re-entrant either: ExecutorSingleTree.hpp:153-160).
 *   - there is NO CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef FMMB_H_
#define FMMB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmmb_plan fmmb_plan;

typedef enum {
  FMMB_OK = 0,
  FMMB_ERR_INVALID = -1,      /* bad argument */
  FMMB_ERR_CUDA = -2,         /* a CUDA call failed (message has the CUDA error) */
  FMMB_ERR_TREE_DEPTH = -3,   /* input needs more than 10 octree levels; the reference loops
                                 forever on such input (include/tree/Octree.hpp:85-92,649) */
  FMMB_ERR_UNSUPPORTED = -4,  /* kernel kind / option not built yet */
  FMMB_ERR_NO_DEVICE = -5     /* no CUDA device: the engine has no CPU path */
} fmmb_status;

/* Kernel classes of the reference (kernel/ *.hpp) that a plan can be built for. */
typedef enum {
  FMMB_LAPLACE_SPHERICAL = 0,          /* kernel/LaplaceSpherical.hpp: charge 1, result 4 */
  FMMB_LAPLACE_SPHERICAL_BEM = 1,      /* kernel/LaplaceSphericalBEM.hpp: panels, charge 1, result 1 */
  FMMB_STOKES_SPHERICAL_STRESSLET = 2, /* kernel/StokesSpherical.hpp built with -DSTRESSLET: charge 6 (g, n),
                                          result 3 (serialrun_stresslet.cpp) */
  FMMB_YUKAWA_CARTESIAN = 3,           /* kernel/YukawaCartesian.hpp: charge 1, result 4, orders 1..16, kappa from
                                          fmmb_kernel_desc */
  FMMB_YUKAWA_CARTESIAN_BEM = 4,       /* kernel/YukawaCartesianBEM.hpp: panels, charge 1, result 1, kappa, quad_k, orders
                                          1..16.  Near field and treecode are pinned to the reference; its FMM
                                          evaluator is broken for this kernel, so the far field is validated against
                                          Direct only */
  FMMB_STOKES_SPHERICAL = 5,           /* kernel/StokesSpherical.hpp default build (Stokeslet): charge 3 (f), result 3 */
  FMMB_STOKES_SPHERICAL_BEM = 6        /* kernel/StokesSphericalBEM.hpp: panels (bc 0 = VELOCITY, 1 = TRACTION), charge 3,
                                          result 3; viscosity in fmmb_kernel_desc.kappa, quad_k and quad_kfine Gauss
                                          rules (examples/StokesBEM.cpp:211-214) */
} fmmb_kernel_kind;

/* Mirrors the kernel constructor arguments: LaplaceSpherical(int p) etc. */
typedef struct {
  int32_t kind;    /* fmmb_kernel_kind */
  int32_t p;       /* expansion order the kernel object was constructed with (1..FMMB_MAX_P) */
  double kappa;    /* Yukawa screening parameter; StokesSphericalBEM: the viscosity Mu (unused for Laplace) */
  int32_t quad_k;  /* BEM Gauss rule per panel: a key of the reference's table, 1, 3, 4, 7, 13, 17, 19, 25 or 79
                      (examples/BEM/GaussQuadrature.hpp:41-276; unused for point kernels) */
  int32_t quad_kfine; /* StokesSphericalBEM::set_Kfine: rule for panels closer than 2 sqrt(2 Area); 0 = the class
                         default 25 (kernel/StokesSphericalBEM.hpp:136).  Other kinds: must be 0. */
} fmmb_kernel_desc;

#define FMMB_MAX_P 16 /* SolverOptions::max_p default, examples/BEM/SolverOptions.hpp:23 */

/* Mirrors FMMOptions (reference include/FMMOptions.hpp:9-49). */
typedef enum { FMMB_EVAL_FMM = 0, FMMB_EVAL_TREECODE = 1 } fmmb_evaluator;

typedef struct {
  double theta;        /* FMMOptions::set_mac_theta, default 0.5 */
  uint32_t ncrit;      /* FMMOptions::set_max_per_box, default 64 */
  int32_t evaluator;   /* fmmb_evaluator; FMMB_EVAL_TREECODE (-eval TREE: M2P instead of M2L/L2L/L2P) is built for every
                          kind (for FMMB_YUKAWA_CARTESIAN_BEM it is the far-field path pinned to the reference, whose
                          FMM evaluator is broken for that kernel class) */
  int32_t device;      /* CUDA device ordinal, -1 = current device */
  int32_t m2l_mode;    /* far-field engine: 0 = auto (2 for orders <= 8, 3 above), 1 = per-pair kernels, 2 = class-major
                          DMMA GEMM + column reduction (orders <= 8), 3 = output-stationary fused sweep (all orders) */
  int32_t rank;        /* multi-GPU: this process's rank (0 when nranks <= 1) */
  int32_t nranks;      /* multi-GPU: number of ranks sharing the matvec; 0 or 1 = single GPU */
  int32_t near_only;   /* plans for preconditioners: 0 = full matvec; 1 = FMMOptions::local_evaluation, only the
                          near field of the traversal (reference include/executor/EvalLocal.hpp:12-72,
                          EvalLocalSparse.hpp); 2 = FMMOptions::block_diagonal, only leaf-with-itself blocks
                          (EvalDiagonalSparse.hpp:12-80).  LaplaceSpherical and the BEM kernel classes. */
  int32_t kernel_flags; /* FMMB_FLAG_* bits, 0 = the reference's behaviour as compiled (occupies what was tail padding:
                          size and offsets of the structure are unchanged) */
} fmmb_options;

/* FMMB_STOKES_SPHERICAL_BEM: evaluate the near-field entries as the reference's source text means them
 * (self terms, fine rule for close panels) instead of as the unmodified reference computes them when compiled, where
 * an expression template that outlives its operand makes every pair take the K-point rule
 * (kernel/StokesSphericalBEM.hpp:162-163,262-263; fmm_bem_relaxed_b200/hostcxx/stokes_bem_math.hpp has the details). */
#define FMMB_FLAG_STOKES_BEM_AS_WRITTEN 1
/* Every kind: skip the warm start of fmmb_plan_create.  By default construction also does what a first matvec would
 * otherwise do on the caller's clock -- expansion / scratch / staging buffers, the translation tables of the order,
 * one matvec on zero charges -- and, for the panel kernel classes (the operators of the relaxed solvers, which walk
 * down through the orders: reference examples/BEM/GMRES.hpp:195-196), prepares every order 1..p on the solver's
 * work vectors with its launch graph captured.  Single-GPU plans only; sharded plans are always cold. */
#define FMMB_FLAG_COLD_PLAN 2

/* Sources, host memory, borrowed for the duration of the call.
 *   points    3*n doubles, point-major (x0,y0,z0,x1,...): the positions the octree is built on.
 *             Point kernels: the sources themselves.  BEM kernels: the panel centres,
 *             static_cast<point_type>(panel) as the reference's tree sees them
 *             (kernel/LaplaceSphericalBEM.hpp:99); NULL = computed from the vertices.
 *   vertices  BEM kernels only: 9*n doubles, (p0, p1, p2) per panel as passed to Panel(p0,p1,p2)
 *             (kernel/LaplaceSphericalBEM.hpp:61-97).
 *   bc        BEM kernels only: n entries, 0 = Panel::POTENTIAL (Stokes: VELOCITY), 1 = Panel::NORMAL_DERIV (Stokes:
 *             TRACTION); NULL = all 0. */
typedef struct {
  int64_t n;
  const double* points;
  const double* vertices;
  const int32_t* bc;
} fmmb_sources;

/* Sizes a caller needs for fmmb_plan_get_tree and for recomputing work counts. */
typedef struct {
  int64_t n_bodies;
  int64_t n_boxes;
  int64_t n_leaves;
  int64_t n_levels;        /* Octree::levels(): max level + 1 ... see reference Octree.hpp:510-512 */
  int64_t n_m2l_pairs;     /* |LR_list|, reference EvalInteractionLazy.hpp:231 */
  int64_t n_p2p_box_pairs; /* sum |P2P_lists[b]|, reference EvalInteractionLazy.hpp:79 */
  int64_t n_p2p_body_pairs;
  int64_t n_m2l_classes;   /* distinct translation vectors handled by the batched M2L */
  int64_t n_m2l_pairs_batched;
  int64_t n_near_entries;  /* BEM: cached near-field matrix entries (= n_p2p_body_pairs) */
  int64_t own_body_begin;  /* multi-GPU: tree-order body range whose results this rank computes */
  int64_t own_body_end;
  int32_t p;               /* current expansion order */
  int32_t charge_dim;      /* doubles per charge (Laplace 1, Stokeslet and StokesSphericalBEM 3, stresslet 6) */
  int32_t result_dim;      /* doubles per result (Laplace 4: potential, fx, fy, fz; BEM 1; Stokes 3) */
  int32_t device;
} fmmb_plan_info;

/* Phase indices for fmmb_plan_phase_times (milliseconds, CUDA events, last execute). */
enum {
  FMMB_T_TOTAL = 0, FMMB_T_UPWARD = 1, FMMB_T_M2L = 2 /* GEMM + reduction */, FMMB_T_DOWNWARD = 3, FMMB_T_P2P = 4,
  FMMB_T_H2D = 5, FMMB_T_D2H = 6,
  FMMB_T_LAUNCHES = 7, /* number of kernel launches of the last execute (a count, not ms) */
  FMMB_T_M2L_GEMM = 8,  /* the batched M2L contraction kernel alone */
  FMMB_T_COUNT = 10
};

/* FMM_plan<K>(K, sources, opts): builds the octree and all interaction lists on the device.
 * Replaces reference include/FMM_plan.hpp:34-43 -> make_executor (executor/make_executor.hpp:67-78)
 * -> Octree ctor (tree/Octree.hpp:485-488) + EvalInteractionLazy ctor (:59-117). */
int fmmb_plan_create(const fmmb_kernel_desc* kernel, const fmmb_sources* sources,
                     const fmmb_options* options, fmmb_plan** out_plan);

/* plan.kernel().set_p(p): takes effect at the next execute.  1 <= p <= FMMB_MAX_P.
 * Replaces reference kernel/LaplaceSpherical.hpp:119-128. */
int fmmb_plan_set_p(fmmb_plan* plan, int p);

/* results = A * charges, host buffers, original order; blocks until results are written.
 * charges: n*charge_dim doubles; results: n*result_dim doubles (overwritten).
 * Replaces reference include/FMM_plan.hpp:75-90. */
int fmmb_plan_execute(fmmb_plan* plan, const double* charges_host, double* results_host);

/* Same with device buffers; asynchronous on the plan's stream (see fmmb_plan_stream /
 * fmmb_plan_sync).  This is what a device-resident GMRES feeds. */
int fmmb_plan_execute_device(fmmb_plan* plan, const double* charges_dev, double* results_dev);

/* Sharded matvec for a solver that keeps its vectors distributed ("results stay sharded by
 * target"): every rank passes the charges of ITS bodies and receives the results of ITS bodies, both as device
 * arrays in TREE order covering the plan's owned range [own_body_begin, own_body_end) of fmmb_plan_info
 * (tree index -> original index: perm of fmmb_plan_get_tree).  The one data exchange besides the multipoles is
 * an NCCL all-gather of the charge slices (8 bytes per body); no result collective, no permutation.
 * Works on a single-GPU plan too (the slice is then the whole tree-ordered vector).  Every kernel kind:
 * own_count * charge_dim doubles in, own_count * result_dim doubles out (LaplaceSpherical can exchange the slices
 * through peer memory, fmmb_plan_peer_init; the other kinds need fmmb_plan_comm_init).
 * Asynchronous on the plan's stream. */
int fmmb_plan_execute_sharded(fmmb_plan* plan, const double* charges_own_dev, double* results_own_dev);

/* The same call with HOST buffers (what a host-side distributed solver holds: the reference's GMRES keeps
 * std::vector's, examples/BEM/GMRES.hpp:142-252): copies this rank's charge slice to the device, runs the sharded
 * matvec, copies this rank's result slice back and blocks until it is written.  Per rank and matvec that is
 * own_count * charge_dim * 8 bytes up and own_count * result_dim * 8 bytes down -- 1/nranks of what
 * fmmb_plan_execute moves.  Pinned host memory makes the copies asynchronous to the host. */
int fmmb_plan_execute_sharded_host(fmmb_plan* plan, const double* charges_own_host, double* results_own_host);

/* Brute force reference sum on the GPU for accuracy checks:
 * Direct::matvec(K, sources, charges, targets, results), reference include/Direct.hpp:273-288.
 * targets: 3*nt doubles (host); results: nt*result_dim doubles (host).  Point kernels only
 * (Laplace, Stokeslet, stresslet). */
int fmmb_plan_direct(fmmb_plan* plan, const double* charges_host, int64_t nt,
                     const double* targets_host, double* results_host);

/* The same for the panel kernels (LaplaceSphericalBEM, YukawaCartesianBEM, StokesSphericalBEM): results[i] = sum over ALL
 * source panels j of the plan of K(t_i, s_j) q_j with K = the kernel class's operator()(target, source)
 * (kernel/LaplaceSphericalBEM.hpp:273-297, YukawaCartesianBEM.hpp:213-230, StokesSphericalBEM.hpp:377-390) -- what
 * examples/StokesBEM.cpp:377-380 and the accuracy checks of the BEM drivers call Direct::matvec for.
 * target_vertices: 9*nt doubles, (p0, p1, p2) per target panel (the kernels evaluate at its centre); target_bc: nt
 * entries like fmmb_sources.bc, NULL = all 0; results: nt*result_dim doubles (host). */
int fmmb_plan_direct_panels(fmmb_plan* plan, const double* charges_host, int64_t nt, const double* target_vertices_host,
                            const int32_t* target_bc_host, double* results_host);

/* Engine knobs that have no counterpart in the reference (all default to the fast setting):
 *   "overlap_p2p"  1 = near field runs on a second stream concurrently with the far field (default),
 *                  0 = every kernel on one stream, so per-kernel CUDA-event times are undisturbed
 *                      (used by bench.py for the roofline figures).
 *   "use_graph"    1 = from the second identical call on, a matvec is replayed as one CUDA graph (default);
 *                  per-kernel phase times are then unavailable (only FMMB_T_TOTAL).  0 = plain launches.
 *   "m2l_mode"     see fmmb_options.m2l_mode.
 *   "p2p_items"    near-field work decomposition of point kernels: 0 = chunks of <= 32 targets in leaf order
 *                  (default); 1 = chunks of 32 targets plus power-of-two pieces of the remainder, longest first.
 *   "p2p_kernel"   0 = one tile per source leaf; 1 = merged source runs, fixed 32-source tiles, prefetch; 2 = the same
 *                  with two targets per lane (default); 3 = 2 with the source tiles staged by 1-D TMA bulk copies.
 *   "p2p_occ"      resident warps per SM the default pair kernel is compiled for: 20, 24, 28 (default) or 32.
 *   "p2p_newton"   1 = Newton-only inverse root in p2p_kernel 3 (16 instead of 18 FP64 instructions per pair; the
 *                  near field is then accurate to ~1e-13 instead of round-off).  Default 0.
 *   "p2p_wps"      > 0: persistent near-field blocks, that many one-warp blocks per SM (default 0 = plain grid).
 *   "p2p_order"    one GPU, class-major far field: 0 (default) = the near field starts with the upward pass on the
 *                  low-priority stream; 1 = it starts behind the M2L GEMM (slower: measured 3.09 vs 2.84 ms at N = 1M).
 *   "m2l_reduce"   1 (default) = M2L column reduction staged through shared memory by TMA bulk copies, "m2l_reduce_bps"
 *                  (1..3, default 2) blocks of four warps per SM: leaves registers and block slots to the near field
 *                  beside it; 0 = a block per box, loads in flight in registers.  Same bits.
 *   "p2p_defer"    sharded plans with an owned upward pass: 1 (default) = the near field starts behind the owned M2M
 *                  sweep, beside the multipole exchange; 0 = with the upward pass (measured on 2 GPUs: 1.88 vs 1.59 ms).
 *   "graph_node_priority"  1 (default) = cached launch graphs keep the stream priorities of their kernels
 *                  (cudaGraphInstantiateFlagUseNodePriority): the far-field chain overtakes the near field's blocks.
 *   "p2m_kernel"   1 (default) = P2M with a narrow transposition tile at P >= 7 (p2m_cols_kernel), 0 = full tile.
 *   "l2p_kernel"   1 (default) = L2P with four leaves per warp in body order (l2p_packed_kernel), 0 = a leaf per warp.
 *   "bem_near_kernel"  cached BEM near field: 1 (default) = eight warps per work item (bem_near_split_kernel), 0 = one.
 *   "p2p_unroll"   pair-loop unroll of p2p_kernel 1: 4 (default) or 8.
 *   "p2p_warps"    warps per block of the near-field pair kernel: 1 (default), 2 or 4.
 *   (measured on B200 at N = 1M: all combinations within 4 %; see profiles/README.md) */
int fmmb_plan_set_option(fmmb_plan* plan, const char* name, int64_t value);

/* ---- multi-GPU (one process per GPU, single node) ---------------------------------------------
 * The matvec shards by TARGET leaves: every rank builds the same tree (replicated), owns a
 * Morton-contiguous range of leaves of equal estimated work (fmmb_partition_ranges), and computes
 * M2L / L2L / L2P / P2P only for its targets.  The one exchange step of a matvec is an all-gather of
 * the result slices over NCCL (NVLink); charges are replicated by the caller (a GMRES vector is).
 *   rank 0:       fmmb_comm_unique_id(id)         -> ship the 128 bytes to the other ranks
 *   every rank:   options.rank / options.nranks at fmmb_plan_create, then fmmb_plan_comm_init(plan, id)
 * Without fmmb_plan_comm_init a partitioned plan writes only its own slice of the results. */
int fmmb_comm_unique_id(unsigned char id[128]);
int fmmb_plan_comm_init(fmmb_plan* plan, const unsigned char id[128]);

/* Multipole exchange through peer memory instead of NCCL (optional, LaplaceSpherical plans, GPUs of one node with
 * NVLink / P2P access): after the owned upward pass every rank stores the multipoles of its boxes straight into
 * the peers' arrays from the producing side (one kernel, no pack / all-gather / unpack), ordered by flag vectors
 * that the peers write with system-scope release stores.
 *   every rank:  fmmb_plan_peer_export(plan, blob)            128 bytes (cudaIpcMemHandle of the multipole array)
 *   caller:      all-gather the blobs, ordered by rank        (any transport; bench.py uses torch.distributed)
 *   every rank:  fmmb_plan_peer_init(plan, blobs)             nranks * 128 bytes
 * With it, fmmb_plan_execute_sharded needs no NCCL communicator at all: the charge slices are stored into the
 * peers' tree-ordered charge vectors the same way.  (fmmb_plan_execute[_device] still all-gathers the result slices
 * over NCCL and needs fmmb_plan_comm_init.)  All ranks must run the same sequence of matvecs (as with any
 * collective). */
int fmmb_plan_peer_export(fmmb_plan* plan, unsigned char blob[128]);
int fmmb_plan_peer_init(fmmb_plan* plan, const unsigned char* blobs);

/* Host-only helper (no GPU needed): cut n non-negative work weights into nranks contiguous ranges of
 * nearly equal sum.  cuts has nranks+1 entries, cuts[0] = 0, cuts[nranks] = n; rank r owns
 * [cuts[r], cuts[r+1]). */
int fmmb_partition_ranges(const double* weights, int64_t n, int nranks, int64_t* cuts);

/* ---- device-resident relaxed GMRES (the caller of the hot path) ------------------------------------------------
 * Mirrors GMRES(plan, x, b, solver_options[, M]) of reference examples/BEM/GMRES.hpp:142-252 with
 * SolverOptions (examples/BEM/SolverOptions.hpp:9-38): restarted GMRES, modified Gram-Schmidt, Givens rotations,
 * convergence on the rotated residual estimate, and p = max(1, predict_p(|resid|)) set before every inner matvec.
 * The Krylov basis and all BLAS-1 work stay on the GPU; one host synchronisation per inner iteration.
 * BEM plans, single GPU: scalar unknowns (Laplace / Yukawa BEM) or Vec<3> unknowns solved on the flat array of
 * 3 n doubles exactly like examples/BEM/GMRES_Stokes.hpp:170-300 (StokesSphericalBEM; set p_min and p_offset = 1
 * for that header's order rule). */
typedef struct {
  double residual;       /* SolverOptions::residual (tolerance on |s[i+1]| / ||b||) */
  int32_t max_iters;     /* SolverOptions::max_iters */
  int32_t restart;       /* SolverOptions::restart */
  uint32_t max_p;        /* SolverOptions::max_p */
  int32_t variable_p;    /* SolverOptions::variable_p: 1 = relax the order with the residual, 0 = always max_p */
  int32_t relax_type;    /* 0 = BOURAS (default), 1 = SIMONCINI */
  int32_t verbose;       /* 1 = print the reference's progress lines ("it: 001, res: ..., fmm_req_p: ...") */
  uint32_t p_min;        /* SolverOptions::p_min: lower bound of the relaxed order (0 or 1 = the GMRES.hpp rule) */
  uint32_t p_offset;     /* subtracted from predict_p before the bound: 0 = GMRES.hpp:195 max(1, predict_p),
                            1 = GMRES_Stokes.hpp:229 max(p_min, predict_p - 1) */
} fmmb_solver_options;

typedef struct {
  int32_t iterations;    /* inner iterations performed */
  int32_t n_records;     /* entries that p_schedule / residuals would hold (may exceed the capacity passed in) */
  int32_t final_p;       /* expansion order the plan is left at (the reference leaves the kernel at the last relaxed p) */
  int32_t reserved;
  double final_residual;
} fmmb_gmres_info;

/* b, x: n * charge_dim doubles (host); x holds the initial guess on entry and the solution on return.
 * diag: NULL (identity) or n * charge_dim doubles d with M(v)_i = d_i v_i (Preconditioners::Diagonal, examples/BEM/Preconditioner.hpp:24-38).
 * p_schedule / residuals: optional arrays of `capacity` entries: order and |resid| of every inner iteration. */
int fmmb_gmres(fmmb_plan* plan, const double* b, double* x, const double* diag, const fmmb_solver_options* options,
               fmmb_gmres_info* info, int32_t* p_schedule, double* residuals, int32_t capacity);

/* Flexible GMRES, device resident (reference examples/BEM/GMRES_Stokes.hpp:319-431 FGMRES, driven by
 * examples/StokesBEM.cpp:309-323): the preconditioned vectors Z[j] are kept and the solution is updated from them.
 * pc_plan: NULL = identity, or a near-field-only plan over the same panels on the same device
 * (fmmb_options.near_only 1 = Preconditioners::LocalInnerSolver, LocalPC_Stokes.hpp:27-62; 2 =
 * Preconditioners::BlockDiagonal, BlockDiagonalPC_Stokes.hpp): every outer iteration solves pc_plan z = v from z = 0 by
 * GMRES with pc_options (the reference's: residual 1e-1, max_iters 1, variable_p 0, restart 50), also on the device.
 * The order rule of the outer iteration comes from options (FGMRES of GMRES_Stokes.hpp:375: p_min 5, p_offset 0). */
int fmmb_fgmres(fmmb_plan* plan, fmmb_plan* pc_plan, const fmmb_solver_options* pc_options, const double* b, double* x,
                const fmmb_solver_options* options, fmmb_gmres_info* info, int32_t* p_schedule, double* residuals,
                int32_t capacity);

int fmmb_plan_sync(fmmb_plan* plan);
void* fmmb_plan_stream(fmmb_plan* plan); /* cudaStream_t */

int fmmb_plan_get_info(fmmb_plan* plan, fmmb_plan_info* info);

/* Copies the tree and lists to host arrays (any pointer may be NULL to skip it):
 *   perm[n]        tree index -> original index (Octree::permute_, reference Octree.hpp:687-691)
 *   codes[n]       Morton code per tree-ordered body (Octree::mc_)
 *   boxes[8*nb]    key(with leaf bit 31), parent, child_begin, child_end (raw box_data fields,
 *                  reference Octree.hpp:196-210), body_begin, body_end, level, is_leaf
 *   geom[4*nb]     centre x,y,z and side length (reference Octree.hpp:243-248,334-355)
 *   m2l_pairs[2*n_m2l_pairs]  (source box, target box) in the reference's LR_list order
 *   p2p_off[nb+1], p2p_idx[n_p2p_box_pairs]  per TARGET box, the source boxes in P2P_lists order
 */
int fmmb_plan_get_tree(fmmb_plan* plan, uint32_t* perm, uint32_t* codes, uint32_t* boxes,
                       double* geom, int32_t* m2l_pairs, int32_t* p2p_off, int32_t* p2p_idx);

/* Multipole / local expansions of the last execute, box-major, nc = p(p+1)/2 complex numbers
 * (re,im interleaved) per box, packed index n(n+1)/2+m as in the reference
 * (kernel/LaplaceSpherical.hpp:193-198).  Host arrays of 2*nc*nb doubles; NULL skips. */
int fmmb_plan_get_expansions(fmmb_plan* plan, double* multipoles, double* locals);

int fmmb_plan_phase_times(fmmb_plan* plan, double* ms, int count);

void fmmb_plan_destroy(fmmb_plan* plan);

const char* fmmb_last_error(void);

/* Optional: create the CUDA context of `device` (-1 = current) and load this library's module now, so that a caller
 * can keep that process-level cost (0.2 - 1 s on a fresh process) out of what it times as plan construction; the
 * first fmmb_plan_create does the same implicitly.  FMMB_ERR_NO_DEVICE without a GPU. */
int fmmb_init(int32_t device);

/* Library build info: "fmmb200 <version> sm_100a". */
const char* fmmb_version(void);

/* Utility used by bench.py for the FP64 roofline: runs dependent-free loops of (a) DFMA and (b) FP64
 * tensor-core MMAs (mma.sync m8n8k4 f64 = DMMA.8x8x4) on every SM and returns the achieved FP64
 * TFLOP/s of each (2 flop per multiply-add).  Either pointer may be NULL. */
int fmmb_measure_fp64_peak(int device, double* tflops_dfma, double* tflops_dmma);

#ifdef __cplusplus
}
#endif
#endif /* FMMB_H_ */
