// csrc/capi.cu -- extern "C" entry points declared in include/fmmb.h.
// No torch, no C++ types across the boundary; every failure becomes a negative status plus a
// thread-local message (the reference prints and exit()s instead: include/executor/P2M.hpp:13-17).
#include "common.cuh"
#include <chrono>
#include <cstring>
#include <algorithm>
#include <new>

namespace fmmb {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return FMMB_OK;
  } catch (const CudaError& e) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e.err, cudaGetErrorString(e.err), e.what,
             e.file, e.line);
    set_error(buf);
    return FMMB_ERR_CUDA;
  } catch (const StatusError& e) {
    set_error(e.msg);
    return e.status;
  } catch (const std::bad_alloc&) {
    set_error("host allocation failed");
    return FMMB_ERR_INVALID;
  }
}

static void update_phase_times(fmmb_plan* plan) {
  if (plan->graph_timed && !plan->timed) {
    // graph launch: only the total is observable (per-kernel events are not recorded inside a capture)
    float t = 0;
    if (cudaEventElapsedTime(&t, plan->ev[0], plan->ev[5]) != cudaSuccess) { cudaGetLastError(); t = 0; }
    for (int i = 0; i < FMMB_T_COUNT; ++i) if (i != FMMB_T_LAUNCHES && i != FMMB_T_H2D && i != FMMB_T_D2H) plan->phase_ms[i] = 0;
    plan->phase_ms[FMMB_T_TOTAL] = t;
    plan->phase_ms[FMMB_T_LAUNCHES] = plan->launches;   // kernels inside the replayed graph (counted at its capture)
    return;
  }
  if (!plan->timed) return;
  auto ms = [&](int a, int b) {
    float t = 0;
    if (cudaEventElapsedTime(&t, plan->ev[a], plan->ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.0; }
    return (double)t;
  };
  if (std::getenv("FMMB_PRINT_TIMELINE"))   // where the phases lie on the clock of the matvec (plain launches only)
    std::fprintf(stderr, "timeline (ms after the first launch): far field %.3f | upward done %.3f | M2L GEMM %.3f .. %.3f | "
                 "M2L done %.3f | downward done %.3f | near field %.3f .. %.3f | results %.3f\n", ms(0, 12), ms(0, 2),
                 ms(0, 13), ms(0, 14), ms(0, 3), ms(0, 4), ms(0, 6), ms(0, 7), ms(0, 5));
  plan->phase_ms[FMMB_T_TOTAL] = ms(0, 5);
  plan->phase_ms[FMMB_T_UPWARD] = ms(12, 2);
  plan->phase_ms[FMMB_T_M2L] = ms(2, 3);
  plan->phase_ms[FMMB_T_DOWNWARD] = ms(3, 4);
  plan->phase_ms[FMMB_T_P2P] = ms(6, 7);
  plan->phase_ms[FMMB_T_LAUNCHES] = plan->launches;
  plan->phase_ms[FMMB_T_M2L_GEMM] = plan->m2l_gemm_timed ? ms(13, 14) : 0.0;
}
}  // namespace fmmb

using namespace fmmb;

extern "C" {

const char* fmmb_last_error(void) { return g_last_error.c_str(); }
const char* fmmb_version(void) { return "fmmb200 0.1 sm_100a"; }

static void warm_start(fmmb_plan* plan);   // after run_matvec

int fmmb_plan_create(const fmmb_kernel_desc* kernel, const fmmb_sources* sources, const fmmb_options* options,
                     fmmb_plan** out_plan) {
  if (!kernel || !sources || !out_plan) { set_error("null argument"); return FMMB_ERR_INVALID; }
  *out_plan = nullptr;
  const bool is_stokes = kernel->kind == FMMB_STOKES_SPHERICAL || kernel->kind == FMMB_STOKES_SPHERICAL_STRESSLET;
  const bool is_yukawa = kernel->kind == FMMB_YUKAWA_CARTESIAN || kernel->kind == FMMB_YUKAWA_CARTESIAN_BEM;
  const bool is_sbem = kernel->kind == FMMB_STOKES_SPHERICAL_BEM;
  if (kernel->kind != FMMB_LAPLACE_SPHERICAL && kernel->kind != FMMB_LAPLACE_SPHERICAL_BEM && !is_stokes && !is_yukawa &&
      !is_sbem) {
    set_error("built kernel kinds: FMMB_LAPLACE_SPHERICAL, FMMB_LAPLACE_SPHERICAL_BEM, FMMB_STOKES_SPHERICAL, "
              "FMMB_STOKES_SPHERICAL_STRESSLET, FMMB_YUKAWA_CARTESIAN, FMMB_YUKAWA_CARTESIAN_BEM, "
              "FMMB_STOKES_SPHERICAL_BEM");
    return FMMB_ERR_UNSUPPORTED;
  }
  const bool is_bem = kernel->kind == FMMB_LAPLACE_SPHERICAL_BEM || kernel->kind == FMMB_YUKAWA_CARTESIAN_BEM;
  const bool is_panel = is_bem || is_sbem;     // sources are triangular panels
  if (is_panel && !sources->vertices) { set_error("BEM kernels need the panel vertices"); return FMMB_ERR_INVALID; }
  if (kernel->p < 1 || kernel->p > FMMB_MAX_P) { set_error("expansion order must be in 1..16"); return FMMB_ERR_INVALID; }
  if (sources->n < 1 || (!sources->points && !(is_panel && sources->vertices))) {
    set_error("need at least one source point");
    return FMMB_ERR_INVALID;
  }
  fmmb_options opts;
  std::memset(&opts, 0, sizeof opts);
  opts.theta = 0.5; opts.ncrit = 64; opts.evaluator = FMMB_EVAL_FMM; opts.device = -1;
  if (options) opts = *options;
  if (opts.nranks < 1) { opts.nranks = 1; opts.rank = 0; }
  if (!(opts.theta > 0)) { set_error("theta must be positive"); return FMMB_ERR_INVALID; }
  if (opts.near_only < 0 || opts.near_only > 2) { set_error("near_only: 0, 1 or 2"); return FMMB_ERR_INVALID; }
  if (opts.near_only && (is_stokes || kernel->kind == FMMB_YUKAWA_CARTESIAN)) {
    set_error("near-field-only plans are built for LaplaceSpherical and the BEM kernel classes");
    return FMMB_ERR_UNSUPPORTED;
  }
  if (opts.ncrit < 1) { set_error("ncrit must be at least 1"); return FMMB_ERR_INVALID; }
  if (opts.evaluator != FMMB_EVAL_FMM && opts.evaluator != FMMB_EVAL_TREECODE) { set_error("unknown evaluator"); return FMMB_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: this engine has no CPU path");
    return FMMB_ERR_NO_DEVICE;
  }
  fmmb_plan* plan = new (std::nothrow) fmmb_plan();
  if (!plan) { set_error("host allocation failed"); return FMMB_ERR_INVALID; }
  int rc = guarded([&] {
    if (opts.device >= 0) FMMB_CUDA(cudaSetDevice(opts.device));
    FMMB_CUDA(cudaGetDevice(&plan->device));
    plan->opts = opts;
    plan->kind = kernel->kind;
    if (const char* g = std::getenv("FMMB_USE_GRAPH")) plan->use_graph = std::atoi(g) != 0;   // same as set_option
    if (const char* g = std::getenv("FMMB_M2L_MODE")) plan->opts.m2l_mode = std::atoi(g);     // experiments: drivers that build their own FMMOptions
    plan->p = kernel->p;
    // the far-field chain (many short dependent kernels and the collectives) outranks the near-field kernel
    // that runs beside it: its blocks take the SM slots first whenever both have work
    int prio_lo = 0, prio_hi = 0;
    FMMB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    FMMB_CUDA(cudaStreamCreateWithPriority(&plan->stream, cudaStreamNonBlocking, prio_hi));
    FMMB_CUDA(cudaStreamCreateWithPriority(&plan->stream2, cudaStreamNonBlocking, prio_lo));
    if (const char* e = std::getenv("FMMB_GRAPH_NODE_PRIORITY")) plan->graph_node_priority = std::atoi(e) != 0;
    if (const char* e = std::getenv("FMMB_M2L_REDUCE")) plan->m2l_reduce = std::atoi(e) != 0;
    for (auto& e : plan->ev) FMMB_CUDA(cudaEventCreate(&e));
    plan->charge_dim = is_stokes ? (kernel->kind == FMMB_STOKES_SPHERICAL_STRESSLET ? 6 : 3) : (is_sbem ? 3 : 1);
    plan->result_dim = is_bem ? 1 : (is_stokes || is_sbem ? 3 : 4);
    laplace_init_tables(plan);
    blocked_init_tables();
    std::vector<double> centres;
    const double* pts = sources->points;
    if (!pts) {
      // panel centres in the reference's operation order: ((p0 + p1) + p2) / 3
      centres.resize(3 * (size_t)sources->n);
      for (int64_t i = 0; i < sources->n; ++i)
        for (int k = 0; k < 3; ++k) {
          const double* v = sources->vertices + 9 * (size_t)i;
          centres[3 * (size_t)i + k] = ((v[k] + v[3 + k]) + v[6 + k]) / 3;
        }
      pts = centres.data();
    }
    // FMMB_PRINT_SETUP=1: host time of each construction step (stream synchronised) on stderr
    const bool print_setup = std::getenv("FMMB_PRINT_SETUP") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_step = now();
    auto step_done = [&](const char* what) {
      if (!print_setup) return;
      cudaStreamSynchronize(plan->stream);
      const double t = now();
      fprintf(stderr, "fmmb setup: %-46s %9.3f ms\n", what, (t - t_step) * 1e3);
      t_step = t;
    };
    step_done("streams, events, constant tables (+ CUDA context)");
    { NvtxRange r("fmmb: octree + dual traversal"); build_tree(plan, pts, sources->n); }
    step_done("octree + dual traversal");
    plan->near_only = opts.near_only;
    if (opts.near_only == 2) restrict_p2p_to_self(plan);
    NvtxRange r_far("fmmb: translation classes / near-field items / kernel setup");
    if (!opts.near_only) {
      // YukawaCartesian[BEM] runs its own translation kernels on the class tables of m2l_classes.cu; the Laplace
      // family runs the engine fmmb_options.m2l_mode selects (laplace_build_far)
      if (is_yukawa) { build_m2l_classes(plan); plan->far_built_classes = true; }
      else laplace_build_far(plan);
    }
    step_done("far-field structures (classes / blocks)");
    if (is_panel) plan->p2p_item_mode = 0;   // one cached near-field block per chunk of <= 32 targets
    build_p2p_items(plan);
    step_done("near-field work items");
    if (is_bem) bem_setup(plan, sources->vertices, sources->bc, kernel->quad_k,
                          kernel->kind == FMMB_YUKAWA_CARTESIAN_BEM ? kernel->kappa : -1.0);
    if (is_stokes) stokes_setup(plan, kernel->kind == FMMB_STOKES_SPHERICAL_STRESSLET);
    if (is_sbem) stokes_bem_setup(plan, sources->vertices, sources->bc, kernel->quad_k, kernel->quad_kfine, kernel->kappa);
    if (is_yukawa) yukawa_setup(plan, kernel->kappa);
    step_done("kernel class setup (panels, cached near field)");
    if (!(opts.kernel_flags & FMMB_FLAG_COLD_PLAN) && plan->tree.nranks <= 1) {
      warm_start(plan);
      step_done("warm start (buffers, tables, launch graphs)");
    }
  });
  if (rc != FMMB_OK) { fmmb_plan_destroy(plan); return rc; }
  *out_plan = plan;
  return FMMB_OK;
}

void fmmb_plan_destroy(fmmb_plan* plan) {
  if (!plan) return;
  cudaSetDevice(plan->device);
  if (plan->stream) cudaStreamSynchronize(plan->stream);
  if (plan->stream2) cudaStreamSynchronize(plan->stream2);
  for (auto& kv : plan->graphs) cudaGraphExecDestroy(kv.second);
  peer_close(plan);
  comm_destroy(plan);
  bem_free(plan->bem);
  stokes_free(plan->stokes);
  stokes_bem_free(plan->sbem);
  yukawa_free(plan->yukawa);
  gmres_free(plan->gmres_ws);
  for (auto& kv : plan->m2l_coeff) delete kv.second;
  for (auto& e : plan->ev) if (e) cudaEventDestroy(e);
  if (plan->stream) cudaStreamDestroy(plan->stream);
  if (plan->stream2) cudaStreamDestroy(plan->stream2);
  delete plan;
}

int fmmb_plan_set_p(fmmb_plan* plan, int p) {
  if (!plan) { set_error("null plan"); return FMMB_ERR_INVALID; }
  if (p < 1 || p > FMMB_MAX_P) { set_error("expansion order must be in 1..16"); return FMMB_ERR_INVALID; }
  if (plan->peer_alloc && p > 8) {
    // the exported multipole block holds 64 doubles per box with the flag vectors and the charge vector behind it
    set_error("plans with a peer-memory exchange run orders 1..8 (the exported multipole block is sized for P = 8)");
    return FMMB_ERR_UNSUPPORTED;
  }
  plan->p = p;
  return FMMB_OK;
}

namespace fmmb {
// One matvec on the plan stream.  The second time a (order, charges, results) combination is seen the
// kernel sequence -- including the second stream and the NCCL collectives -- is captured into a CUDA
// graph; from then on a matvec is a single graph launch (no per-kernel launch gaps).
static void drop_graphs(fmmb_plan* plan) {
  for (auto& kv : plan->graphs) cudaGraphExecDestroy(kv.second);
  plan->graphs.clear();
  plan->graph_seen.clear();
  plan->graphs_valid_at = plan->realloc_count;
}

static void run_matvec(fmmb_plan* plan, const double* q, double* r) {
  struct CountScope {                       // DevBuf frees during this call are charged to this plan
    unsigned long long* prev;
    explicit CountScope(fmmb_plan* p) : prev(g_realloc_counter) { g_realloc_counter = &p->realloc_count; }
    ~CountScope() { g_realloc_counter = prev; }
  } scope(plan);
  NvtxRange r_mv(plan->call_sharded ? "fmmb: sharded matvec" : "fmmb: matvec");
  auto direct = [&] {
    // sharded call of a class that gathers its charges through a permutation: assemble the tree-ordered vector first
    if (plan->call_sharded && plan->kind != FMMB_LAPLACE_SPHERICAL) plan->sharded_q = sharded_assemble_charges(plan, q, plan->stream);
    if (plan->bem && plan->yukawa) yukawa_bem_execute(plan, q, r);
    else if (plan->bem) bem_execute(plan, q, r);
    else if (plan->stokes) stokes_execute(plan, q, r);
    else if (plan->sbem) stokes_bem_execute(plan, q, r);
    else if (plan->yukawa) yukawa_execute(plan, q, r);
    else laplace_execute(plan, q, r);
  };
  if (!plan->use_graph || !plan->overlap_p2p) { direct(); return; }
  // The expansion arrays are laid out for ONE order at a time (row stride, the all-zero row behind the last box);
  // moving them to this call's order is host-driven work that a cached graph does not contain, so it happens here,
  // ahead of a replay as much as ahead of plain launches (Yukawa keeps per-order strides of its own and no zero row)
  if (plan->stokes) stokes_prepare_expansions(plan);
  else if (plan->sbem) stokes_bem_prepare_expansions(plan);
  else if (!plan->yukawa) laplace_prepare_expansions(plan);
  // a buffer that a cached graph points into was freed since the capture (a larger order came by): the graphs of
  // every order are stale.  They are rebuilt on the next two calls of their key.
  if (plan->graphs_valid_at != plan->realloc_count) {
    FMMB_CUDA(cudaStreamSynchronize(plan->stream));
    drop_graphs(plan);
  }
  fmmb_plan::GraphKey key{plan->p, q, r, plan->call_sharded ? 1 : 0};
  cudaStream_t s = plan->stream;
  auto it = plan->graphs.find(key);
  if (it == plan->graphs.end()) {
    if (plan->graph_seen[key]++ == 0) { direct(); return; }   // first call allocates buffers, builds tables
    cudaGraph_t g = nullptr;
    plan->capturing = true;
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    if (e == cudaSuccess) {
      try { direct(); } catch (...) { plan->capturing = false; cudaStreamEndCapture(s, &g); if (g) cudaGraphDestroy(g); plan->use_graph = false; throw; }
      e = cudaStreamEndCapture(s, &g);
    }
    plan->capturing = false;
    if (plan->graphs_valid_at != plan->realloc_count) {   // the captured launches already point at freed memory
      if (g) cudaGraphDestroy(g);
      drop_graphs(plan);
      direct();
      return;
    }
    if (e == cudaSuccess && g && std::getenv("FMMB_PRINT_GRAPH")) {
      size_t nn = 0;
      cudaGraphGetNodes(g, nullptr, &nn);
      std::vector<cudaGraphNode_t> nodes(nn);
      cudaGraphGetNodes(g, nodes.data(), &nn);
      for (size_t i = 0; i < nn; ++i) {
        cudaGraphNodeType ty;
        cudaGraphNodeGetType(nodes[i], &ty);
        if (ty != cudaGraphNodeTypeKernel) { std::fprintf(stderr, "node %zu type %d\n", i, (int)ty); continue; }
        cudaLaunchAttributeValue v{};
        cudaError_t ge = cudaGraphKernelNodeGetAttribute(nodes[i], cudaLaunchAttributePriority, &v);
        cudaKernelNodeParams kp{};
        cudaGraphKernelNodeGetParams(nodes[i], &kp);
        std::fprintf(stderr, "node %zu kernel grid %u block %u priority %d (%s)\n", i, kp.gridDim.x, kp.blockDim.x,
                     v.priority, cudaGetErrorName(ge));
      }
    }
    cudaGraphExec_t ex = nullptr;
    // per-node priorities: the far-field chain was captured from the high-priority stream, the near field from the
    // other one; without the flag every node runs at the priority of the launching stream
    if (e == cudaSuccess && g) e = cudaGraphInstantiate(&ex, g, plan->graph_node_priority ? cudaGraphInstantiateFlagUseNodePriority : 0);
    if (g) cudaGraphDestroy(g);
    if (e != cudaSuccess || !ex) {          // graphs are an optimisation: fall back to plain launches
      cudaGetLastError();
      plan->use_graph = false;
      direct();
      return;
    }
    it = plan->graphs.emplace(key, ex).first;
  }
  FMMB_CUDA(cudaEventRecord(plan->ev[0], s));
  FMMB_CUDA(cudaGraphLaunch(it->second, s));
  FMMB_CUDA(cudaEventRecord(plan->ev[5], s));
  plan->timed = false;
  plan->graph_timed = true;
}
}  // namespace fmmb

extern "C++" {
namespace fmmb {
void run_matvec_for_solver(fmmb_plan* plan, const double* q, double* r) { run_matvec(plan, q, r); }
}
}

}  // extern "C"

// Everything a first matvec would otherwise do on the caller's clock happens at construction: expansion and scratch
// buffers, the translation tables of the order, the host staging buffers.  A panel plan is the operator of a relaxed
// Krylov solve that walks down through the orders (GMRES.hpp:195-196), so for those every order 1..P is prepared on
// the solver's work vectors and its launch graph captured; the solve then starts on replays.  Single-GPU plans only
// (a sharded plan cannot run before its communicator exists).  FMMB_FLAG_COLD_PLAN skips all of it.
static void warm_start(fmmb_plan* plan) {
  const size_t n = (size_t)plan->tree.n;
  cudaStream_t s = plan->stream;
  plan->charges.resize(plan->charge_dim * n);
  plan->results.resize(plan->result_dim * n);
  plan->charges.zero(s);
  fmmb::run_matvec_for_solver(plan, plan->charges.p, plan->results.p);
  if ((plan->bem || plan->sbem) && !plan->near_only && plan->charge_dim == plan->result_dim) {
    double *z = nullptr, *w = nullptr;
    gmres_reserve(plan, &z, &w);
    const int P = plan->p;
    for (int p = P; p >= 1; --p) {
      plan->p = p;
      fmmb::run_matvec_for_solver(plan, z, w);      // buffers and tables of the order
      fmmb::run_matvec_for_solver(plan, z, w);      // captured
    }
    plan->p = P;
  }
  FMMB_CUDA(cudaStreamSynchronize(s));
  plan->timed = false;
  plan->graph_timed = false;
}

extern "C" {

int fmmb_gmres(fmmb_plan* plan, const double* b, double* x, const double* diag, const fmmb_solver_options* options,
               fmmb_gmres_info* info, int32_t* p_schedule, double* residuals, int32_t capacity) {
  if (!plan || !b || !x || !options) { set_error("null argument"); return FMMB_ERR_INVALID; }
  if (options->max_p < 1 || options->max_p > FMMB_MAX_P || !(options->residual > 0) || options->restart < 1) {
    set_error("solver options: residual > 0, restart >= 1, 1 <= max_p <= 16");
    return FMMB_ERR_INVALID;
  }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    gmres_solve(plan, b, x, diag, *options, info, p_schedule, residuals, capacity < 0 ? 0 : capacity);
  });
}

int fmmb_fgmres(fmmb_plan* plan, fmmb_plan* pc_plan, const fmmb_solver_options* pc_options, const double* b, double* x,
                const fmmb_solver_options* options, fmmb_gmres_info* info, int32_t* p_schedule, double* residuals,
                int32_t capacity) {
  if (!plan || !b || !x || !options) { set_error("null argument"); return FMMB_ERR_INVALID; }
  for (const fmmb_solver_options* o : {options, pc_plan ? pc_options : options}) {
    if (!o) { set_error("a preconditioner plan needs its solver options"); return FMMB_ERR_INVALID; }
    if (o->max_p < 1 || o->max_p > FMMB_MAX_P || !(o->residual > 0) || o->restart < 1) {
      set_error("solver options: residual > 0, restart >= 1, 1 <= max_p <= 16");
      return FMMB_ERR_INVALID;
    }
  }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    fgmres_solve(plan, pc_plan, pc_options, b, x, *options, info, p_schedule, residuals, capacity < 0 ? 0 : capacity);
  });
}

int fmmb_plan_execute_device(fmmb_plan* plan, const double* charges_dev, double* results_dev) {
  if (!plan || !charges_dev || !results_dev) { set_error("null argument"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    run_matvec(plan, charges_dev, results_dev);
  });
}

int fmmb_plan_execute_sharded(fmmb_plan* plan, const double* charges_own_dev, double* results_own_dev) {
  if (!plan || !charges_own_dev || !results_own_dev) { set_error("null argument"); return FMMB_ERR_INVALID; }
  if (plan->tree.nranks > 1 && !plan->comm && !plan->peer_ready) {
    set_error("call fmmb_plan_comm_init or fmmb_plan_peer_init first");
    return FMMB_ERR_INVALID;
  }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    plan->call_sharded = true;
    try { run_matvec(plan, charges_own_dev, results_own_dev); } catch (...) { plan->call_sharded = false; throw; }
    plan->call_sharded = false;
  });
}

int fmmb_plan_execute_sharded_host(fmmb_plan* plan, const double* charges_own_host, double* results_own_host) {
  if (!plan || !charges_own_host || !results_own_host) { set_error("null argument"); return FMMB_ERR_INVALID; }
  if (plan->tree.nranks > 1 && !plan->comm && !plan->peer_ready) {
    set_error("call fmmb_plan_comm_init or fmmb_plan_peer_init first");
    return FMMB_ERR_INVALID;
  }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    const size_t own = (size_t)(plan->tree.own_b1 - plan->tree.own_b0), cd = plan->charge_dim, rd = plan->result_dim;
    cudaStream_t s = plan->stream;
    // staging sized once for the slice (a rank may own nothing: keep the pointers valid)
    if (plan->own_q.n < std::max<size_t>(1, cd * own)) plan->own_q.resize(std::max<size_t>(1, cd * own));
    if (plan->own_r.n < std::max<size_t>(1, rd * own)) plan->own_r.resize(std::max<size_t>(1, rd * own));
    FMMB_CUDA(cudaEventRecord(plan->ev[8], s));
    if (own) FMMB_CUDA(cudaMemcpyAsync(plan->own_q.p, charges_own_host, cd * own * sizeof(double), cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaEventRecord(plan->ev[9], s));
    plan->call_sharded = true;
    try { run_matvec(plan, plan->own_q.p, plan->own_r.p); } catch (...) { plan->call_sharded = false; throw; }
    plan->call_sharded = false;
    FMMB_CUDA(cudaEventRecord(plan->ev[10], s));
    if (own) FMMB_CUDA(cudaMemcpyAsync(results_own_host, plan->own_r.p, rd * own * sizeof(double), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaEventRecord(plan->ev[11], s));
    peer_flag_fetch(plan, s);
    FMMB_CUDA(cudaStreamSynchronize(s));
    update_phase_times(plan);
    float t = 0;
    FMMB_CUDA(cudaEventElapsedTime(&t, plan->ev[8], plan->ev[9])); plan->phase_ms[FMMB_T_H2D] = t;
    FMMB_CUDA(cudaEventElapsedTime(&t, plan->ev[10], plan->ev[11])); plan->phase_ms[FMMB_T_D2H] = t;
    peer_check_timeout(plan);
  });
}

int fmmb_plan_execute(fmmb_plan* plan, const double* charges_host, double* results_host) {
  if (!plan || !charges_host || !results_host) { set_error("null argument"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    const int64_t n = plan->tree.n;
    cudaStream_t s = plan->stream;
    const size_t rd = plan->result_dim;
    const size_t cd = plan->charge_dim;
    plan->charges.resize(cd * (size_t)n);
    plan->results.resize(rd * (size_t)n);
    if (plan->tree.nranks > 1 && !plan->comm) plan->results.zero(s);   // only the owned slice gets written
    FMMB_CUDA(cudaEventRecord(plan->ev[8], s));
    FMMB_CUDA(cudaMemcpyAsync(plan->charges.p, charges_host, cd * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    FMMB_CUDA(cudaEventRecord(plan->ev[9], s));
    run_matvec(plan, plan->charges.p, plan->results.p);
    FMMB_CUDA(cudaEventRecord(plan->ev[10], s));
    FMMB_CUDA(cudaMemcpyAsync(results_host, plan->results.p, rd * (size_t)n * sizeof(double),
                              cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaEventRecord(plan->ev[11], s));
    FMMB_CUDA(cudaStreamSynchronize(s));
    update_phase_times(plan);
    float t = 0;
    FMMB_CUDA(cudaEventElapsedTime(&t, plan->ev[8], plan->ev[9])); plan->phase_ms[FMMB_T_H2D] = t;
    FMMB_CUDA(cudaEventElapsedTime(&t, plan->ev[10], plan->ev[11])); plan->phase_ms[FMMB_T_D2H] = t;
  });
}

int fmmb_plan_direct(fmmb_plan* plan, const double* charges_host, int64_t nt, const double* targets_host,
                     double* results_host) {
  if (!plan || !charges_host || !targets_host || !results_host || nt < 0) { set_error("bad argument"); return FMMB_ERR_INVALID; }
  if (plan->bem || plan->sbem) { set_error("fmmb_plan_direct is built for point kernels only"); return FMMB_ERR_UNSUPPORTED; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    cudaStream_t s = plan->stream;
    DevBuf<double> q, t, out;
    const size_t rd = plan->result_dim;
    q.from_host(charges_host, (size_t)plan->charge_dim * plan->tree.n, s);
    t.from_host(targets_host, 3 * (size_t)nt, s);
    out.resize(rd * (size_t)nt);
    if (nt && plan->yukawa)
      yukawa_direct_raw(yukawa_kappa(plan->yukawa), plan->tree.pts_orig.p, q.p, plan->tree.n, t.p, nt, out.p, s);
    else if (nt && plan->stokes)
      stokes_direct_raw(stokes_is_stresslet(plan->stokes), plan->tree.pts_orig.p, q.p, plan->tree.n, t.p, nt, out.p, s);
    else if (nt) laplace_direct_raw(plan->tree.pts_orig.p, q.p, plan->tree.n, t.p, nt, out.p, s);
    if (nt) FMMB_CUDA(cudaMemcpyAsync(results_host, out.p, rd * (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
  });
}

int fmmb_plan_direct_panels(fmmb_plan* plan, const double* charges_host, int64_t nt, const double* target_vertices_host,
                            const int32_t* target_bc_host, double* results_host) {
  if (!plan || !charges_host || !target_vertices_host || !results_host || nt < 0) { set_error("bad argument"); return FMMB_ERR_INVALID; }
  if (!plan->bem && !plan->sbem) { set_error("fmmb_plan_direct_panels is for the BEM kernel kinds (fmmb_plan_direct: point kernels)"); return FMMB_ERR_UNSUPPORTED; }
  for (int64_t i = 0; target_bc_host && i < nt; ++i)
    if (target_bc_host[i] != 0 && target_bc_host[i] != 1) { set_error("bc entries must be 0 or 1"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    cudaStream_t s = plan->stream;
    DevBuf<double> q, t, out;
    DevBuf<int> bc;
    const size_t rd = plan->result_dim;
    q.from_host(charges_host, (size_t)plan->charge_dim * plan->tree.n, s);
    t.from_host(target_vertices_host, 9 * (size_t)nt, s);
    if (target_bc_host) bc.from_host(target_bc_host, (size_t)nt, s);
    out.resize(rd * (size_t)nt);
    if (plan->sbem) stokes_bem_direct(plan, q.p, nt, t.p, target_bc_host ? bc.p : nullptr, out.p, s);
    else bem_direct(plan, q.p, nt, t.p, target_bc_host ? bc.p : nullptr, out.p, s);
    if (nt) FMMB_CUDA(cudaMemcpyAsync(results_host, out.p, rd * (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, s));
    FMMB_CUDA(cudaStreamSynchronize(s));
  });
}

int fmmb_plan_set_option(fmmb_plan* plan, const char* name, int64_t value) {
  if (!plan || !name) { set_error("null argument"); return FMMB_ERR_INVALID; }
  if (!std::strcmp(name, "overlap_p2p")) { plan->overlap_p2p = value != 0; return FMMB_OK; }
  if (!std::strcmp(name, "m2l_mode")) {
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);   // graphs captured with the other path are stale
      if (value < 0 || value > 3) throw StatusError{FMMB_ERR_INVALID, "m2l_mode: 0 (auto), 1 (per pair), 2 (class-major GEMM), 3 (blocked)"};
      plan->opts.m2l_mode = (int32_t)value;
      if (!plan->yukawa) laplace_build_far(plan);
    });
  }
  if (!std::strcmp(name, "use_graph")) { plan->use_graph = value != 0; return FMMB_OK; }
  if (!std::strcmp(name, "graph_node_priority")) {
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);
      plan->graph_node_priority = value != 0;
    });
  }
  if (!std::strcmp(name, "p2p_kernel") || !std::strcmp(name, "p2p_unroll")) {
    const bool kern = !std::strcmp(name, "p2p_kernel");
    if (kern ? (value < 0 || value > 3) : (value != 4 && value != 8)) {
      set_error("p2p_kernel: 0, 1, 2 or 3; p2p_unroll: 4 or 8");
      return FMMB_ERR_INVALID;
    }
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);
      if (kern) plan->p2p_kernel = (int)value; else plan->p2p_unroll = (int)value;
    });
  }
  if (!std::strcmp(name, "p2p_wps")) {
    if (value < 0 || value > 32) { set_error("p2p_wps: 0 (plain grid) .. 32 persistent near-field warps per SM"); return FMMB_ERR_INVALID; }
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);
      plan->p2p_wps = (int)value;
    });
  }
  if (!std::strcmp(name, "p2m_kernel") || !std::strcmp(name, "l2p_kernel") || !std::strcmp(name, "bem_near_kernel")) {
    int* which = !std::strcmp(name, "p2m_kernel") ? &plan->p2m_kernel
                 : (!std::strcmp(name, "l2p_kernel") ? &plan->l2p_kernel : &plan->bem_near_kernel);
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);
      *which = value != 0;
    });
  }
  if (!std::strcmp(name, "p2p_occ")) {
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);
      plan->p2p_occ = (int)value;
    });
  }
  if (!std::strcmp(name, "p2p_order") || !std::strcmp(name, "m2l_reduce") || !std::strcmp(name, "m2l_reduce_bps") ||
      !std::strcmp(name, "p2p_defer")) {
    const bool bps = !std::strcmp(name, "m2l_reduce_bps");
    int* which = !std::strcmp(name, "p2p_order") ? &plan->p2p_order
                 : (bps ? &plan->m2l_reduce_bps : (!std::strcmp(name, "p2p_defer") ? &plan->p2p_defer : &plan->m2l_reduce));
    if (bps ? (value < 1 || value > 3) : (value != 0 && value != 1)) {
      set_error("p2p_order / m2l_reduce / p2p_defer: 0 or 1; m2l_reduce_bps: 1..3 blocks per SM");
      return FMMB_ERR_INVALID;
    }
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);
      *which = (int)value;
    });
  }
  if (!std::strcmp(name, "p2p_newton")) {
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);
      plan->p2p_newton = value != 0;
    });
  }
  if (!std::strcmp(name, "p2p_warps") || !std::strcmp(name, "p2p_items")) {
    const bool items = !std::strcmp(name, "p2p_items");
    if (items && (plan->bem || plan->sbem)) { set_error("BEM plans keep one near-field block per chunk"); return FMMB_ERR_UNSUPPORTED; }
    if (items ? (value != 0 && value != 1) : (value != 1 && value != 2 && value != 4)) {
      set_error("p2p_items: 0 or 1; p2p_warps: 1, 2 or 4");
      return FMMB_ERR_INVALID;
    }
    return guarded([&] {
      FMMB_CUDA(cudaSetDevice(plan->device));
      FMMB_CUDA(cudaStreamSynchronize(plan->stream));
      drop_graphs(plan);   // captured launches are stale
      if (items) { plan->p2p_item_mode = (int)value; build_p2p_items(plan); }
      else plan->p2p_warps = (int)value;
    });
  }
  set_error(std::string("unknown option: ") + name);
  return FMMB_ERR_INVALID;
}

int fmmb_comm_unique_id(unsigned char id[128]) {
  if (!id) { set_error("null argument"); return FMMB_ERR_INVALID; }
  return guarded([&] { comm_unique_id(id); });
}

int fmmb_plan_comm_init(fmmb_plan* plan, const unsigned char id[128]) {
  if (!plan || !id) { set_error("null argument"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    FMMB_CUDA(cudaStreamSynchronize(plan->stream));
    drop_graphs(plan);   // graphs captured before the communicator existed took the single-rank path
    comm_init(plan, id);
  });
}

int fmmb_plan_peer_export(fmmb_plan* plan, unsigned char blob[128]) {
  if (!plan || !blob) { set_error("null argument"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    FMMB_CUDA(cudaStreamSynchronize(plan->stream));
    drop_graphs(plan);   // the multipole array moves: captured launches are stale
    peer_export(plan, blob);
  });
}

int fmmb_plan_peer_init(fmmb_plan* plan, const unsigned char* blobs) {
  if (!plan || !blobs) { set_error("null argument"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    drop_graphs(plan);
    peer_init(plan, blobs);
  });
}

int fmmb_partition_ranges(const double* weights, int64_t n, int nranks, int64_t* cuts) {
  if (!weights || !cuts || n < 0 || nranks < 1) { set_error("bad argument"); return FMMB_ERR_INVALID; }
  partition_ranges(weights, n, nranks, cuts);
  return FMMB_OK;
}

int fmmb_plan_sync(fmmb_plan* plan) {
  if (!plan) { set_error("null plan"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    peer_flag_fetch(plan, plan->stream);
    FMMB_CUDA(cudaStreamSynchronize(plan->stream));
    update_phase_times(plan);
    peer_check_timeout(plan);
  });
}

void* fmmb_plan_stream(fmmb_plan* plan) { return plan ? (void*)plan->stream : nullptr; }

int fmmb_plan_get_info(fmmb_plan* plan, fmmb_plan_info* info) {
  if (!plan || !info) { set_error("null argument"); return FMMB_ERR_INVALID; }
  std::memset(info, 0, sizeof *info);
  const Tree& T = plan->tree;
  info->n_bodies = T.n; info->n_boxes = T.nboxes; info->n_leaves = T.nleaves; info->n_levels = T.nlevels;
  info->n_m2l_pairs = T.n_lr; info->n_p2p_box_pairs = T.n_p2p; info->n_p2p_body_pairs = T.n_p2p_body_pairs;
  if (plan->far_built_classes) { info->n_m2l_classes = plan->cls.n_classes; info->n_m2l_pairs_batched = plan->cls.n_pairs; }
  else if (plan->far_built_blocked) { info->n_m2l_classes = plan->b_m2l.n_classes; info->n_m2l_pairs_batched = plan->b_m2l.n_pairs; }
  info->own_body_begin = T.own_b0; info->own_body_end = T.own_b1;
  info->n_near_entries = plan->bem ? bem_nnz(plan->bem) : (plan->sbem ? stokes_bem_nnz(plan->sbem) : 0);
  info->p = plan->p; info->charge_dim = plan->charge_dim; info->result_dim = plan->result_dim; info->device = plan->device;
  return FMMB_OK;
}

int fmmb_plan_get_tree(fmmb_plan* plan, uint32_t* perm, uint32_t* codes, uint32_t* boxes, double* geom,
                       int32_t* m2l_pairs, int32_t* p2p_off, int32_t* p2p_idx) {
  if (!plan) { set_error("null plan"); return FMMB_ERR_INVALID; }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    const Tree& T = plan->tree;
    cudaStream_t s = plan->stream;
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
      if (dst && bytes) FMMB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    };
    d2h(perm, T.perm.p, T.n * sizeof(unsigned));
    d2h(codes, T.code.p, T.n * sizeof(unsigned));
    d2h(m2l_pairs, T.lr.p, T.n_lr * sizeof(int2));
    d2h(p2p_off, T.p2p_off.p, (size_t)(T.nboxes + 1) * sizeof(int));
    d2h(p2p_idx, T.p2p_src.p, T.n_p2p * sizeof(int));
    FMMB_CUDA(cudaStreamSynchronize(s));
    const int nb = T.nboxes;
    if (boxes) {
      std::vector<unsigned> key = T.key.to_host(s), par = T.parent.to_host(s), cb = T.cbegin.to_host(s),
                            ce = T.cend.to_host(s), bb = T.bbegin.to_host(s), be = T.bend.to_host(s),
                            lv = T.level.to_host(s);
      for (int b = 0; b < nb; ++b) {
        uint32_t* r = boxes + 8 * (size_t)b;
        r[0] = key[b]; r[1] = par[b]; r[2] = cb[b]; r[3] = ce[b]; r[4] = bb[b]; r[5] = be[b]; r[6] = lv[b];
        r[7] = key[b] >> 31;
      }
    }
    if (geom) {
      std::vector<double4> c = T.center.to_host(s);
      for (int b = 0; b < nb; ++b) {
        geom[4 * (size_t)b] = c[b].x; geom[4 * (size_t)b + 1] = c[b].y; geom[4 * (size_t)b + 2] = c[b].z;
        geom[4 * (size_t)b + 3] = c[b].w;
      }
    }
  });
}

int fmmb_plan_get_expansions(fmmb_plan* plan, double* multipoles, double* locals) {
  if (!plan) { set_error("null plan"); return FMMB_ERR_INVALID; }
  if (plan->stokes || plan->yukawa || plan->sbem) {
    set_error("fmmb_plan_get_expansions returns LaplaceSpherical[BEM] expansions only");
    return FMMB_ERR_UNSUPPORTED;
  }
  return guarded([&] {
    FMMB_CUDA(cudaSetDevice(plan->device));
    cudaStream_t s = plan->stream;
    FMMB_CUDA(cudaStreamSynchronize(s));
    // device layout: P^2 reals per box (stride padded to even); API layout: packed complex
    const int P = plan->p, pp = P * P, xs = (pp + 1) & ~1, nc = P * (P + 1) / 2;
    const size_t nb = plan->tree.nboxes;
    auto convert = [&](const DevBuf<double>& X, double* out) {
      if (!out || X.n < nb * xs) return;
      std::vector<double> h(nb * xs);
      FMMB_CUDA(cudaMemcpyAsync(h.data(), X.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
      FMMB_CUDA(cudaStreamSynchronize(s));
      for (size_t b = 0; b < nb; ++b)
        for (int n = 0; n < P; ++n)
          for (int m = 0; m <= n; ++m) {
            size_t o = 2 * (b * nc + n * (n + 1) / 2 + m);
            out[o] = h[b * xs + n * n + n + m];
            out[o + 1] = m > 0 ? h[b * xs + n * n + n - m] : 0.0;
          }
    };
    convert(plan->M, multipoles);
    convert(plan->L, locals);
  });
}

int fmmb_plan_phase_times(fmmb_plan* plan, double* ms, int count) {
  if (!plan || !ms) { set_error("null argument"); return FMMB_ERR_INVALID; }
  for (int i = 0; i < count && i < FMMB_T_COUNT; ++i) ms[i] = plan->phase_ms[i];
  return FMMB_OK;
}

int fmmb_init(int32_t device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: this engine has no CPU path");
    return FMMB_ERR_NO_DEVICE;
  }
  return guarded([&] {
    if (device >= 0) FMMB_CUDA(cudaSetDevice(device));
    FMMB_CUDA(cudaFree(nullptr));          // context
    fmmb_plan tmp;                         // constant tables: the first use of the module loads it
    FMMB_CUDA(cudaGetDevice(&tmp.device));
    laplace_init_tables(&tmp);
    blocked_init_tables();
    FMMB_CUDA(cudaDeviceSynchronize());
  });
}

int fmmb_measure_fp64_peak(int device, double* tflops_dfma, double* tflops_dmma) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device");
    return FMMB_ERR_NO_DEVICE;
  }
  return guarded([&] {
    if (device >= 0) FMMB_CUDA(cudaSetDevice(device));
    double a = 0, b = 0;
    measure_fp64_peak(&a, &b);
    if (tflops_dfma) *tflops_dfma = a;
    if (tflops_dmma) *tflops_dmma = b;
  });
}

}  // extern "C"
