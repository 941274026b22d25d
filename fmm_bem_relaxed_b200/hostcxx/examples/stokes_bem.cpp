// StokesSphericalBEM, flow past the unit sphere: the problem of the reference's examples/StokesBEM.cpp (:229-279)
// solved with the device-resident relaxed GMRES (fmmb_gmres through GMRES_device; the reference's own driver and its
// GMRES_Stokes.hpp compile unchanged against these headers as bin/ref_StokesBEM).
//   stokes_bem -recursions 5 -p 8 -k 4 -kfine 19 -mu 1e-3 -pmin 5 -solver_tol 1e-5 [-fixed_p] [-as_written] [-check N]
//              [-fgmres [-local | -diagonal]]   the reference driver's flexible GMRES and its two inner-solve
//              preconditioners (examples/StokesBEM.cpp:309-323), device resident (fmmb_fgmres)
// First-kind equation for the traction t on the sphere moving with u = (1, 0, 0):  A t = b, b = (4 pi, 0, 0) on every
// panel (the reference overwrites its computed right-hand side with exactly this, :262-266), x0 = 0; the drag
// sum_j t_j[0] Area_j is compared with Stokes' law 6 pi mu.
// -check N also compares the first N rows of the GPU matvec with Direct::matvec of the host kernel class.
#include <FMM_plan.hpp>
#include <StokesSphericalBEM.hpp>
#include <Triangulation.hpp>
#include <GMRES.hpp>

#include <algorithm>
#include <memory>
#include <cmath>
#include <cstring>

int main(int argc, char** argv) {
  int recursions = 5, p = 8, k = 4, kfine = 19, check = 0;
  double mu = 1e-3;
  bool as_written = false, fgmres = false, pc_local = false, pc_diagonal = false;
  FMMOptions opts = get_options(argc, argv);
  opts.sparse_local = true;
  SolverOptions so;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-recursions")) recursions = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-p")) p = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-pmin")) so.p_min = (unsigned)atoi(argv[++i]);
    else if (!strcmp(argv[i], "-k")) k = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-kfine")) kfine = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-mu")) mu = atof(argv[++i]);
    else if (!strcmp(argv[i], "-solver_tol")) so.residual = atof(argv[++i]);
    else if (!strcmp(argv[i], "-fixed_p")) so.variable_p = false;
    else if (!strcmp(argv[i], "-as_written")) as_written = true;
    else if (!strcmp(argv[i], "-check")) check = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-fgmres")) fgmres = true;
    else if (!strcmp(argv[i], "-local")) pc_local = true;
    else if (!strcmp(argv[i], "-diagonal")) pc_diagonal = true;
  }
  if (pc_local || pc_diagonal) fgmres = true;     // the reference's driver: -local / -diagonal select FGMRES (StokesBEM.cpp:170-186)
  so.max_p = p;
  so.max_iters = so.restart = 100;
  typedef StokesSphericalBEM kernel_type;
  typedef kernel_type::charge_type charge_type;
  typedef kernel_type::result_type result_type;
  kernel_type K(p, k, mu);
  K.set_Kfine(kfine);
  K.near_field_as_written = as_written;
  std::vector<kernel_type::source_type> panels;
  Triangulation::UnitSphere(panels, recursions);
  const size_t n = panels.size();

  // "setup" is the plan (the reference times its right-hand-side plan there, examples/StokesBEM.cpp:265-283, and
  // overwrites that right-hand side with the constant below); the CUDA context + module load is reported by itself
  double tic = get_time();
  fmmb_init(opts.device);
  double context_time = get_time() - tic;
  tic = get_time();
  FMM_plan<kernel_type> plan(K, panels, opts);
  double setup = get_time() - tic;
  if (!plan.handle()) return 1;
  if (check > 0) {   // the GPU matvec against Direct::matvec with the host kernel class on the first rows
    check = std::min<int>(check, (int)n);
    std::vector<charge_type> q(n);
    for (size_t i = 0; i < n; ++i) q[i] = charge_type(1. + 0.1 * (i % 7), -0.5 + 0.05 * (i % 11), 0.25 * (i % 3));
    std::vector<result_type> r = plan.execute(q);
    if (r.empty()) return 1;
    std::vector<kernel_type::target_type> tg(panels.begin(), panels.begin() + check);
    std::vector<result_type> d(check, result_type(0.));
    Direct::matvec(K, panels.begin(), panels.end(), q.begin(), tg.begin(), tg.end(), d.begin());
    double e1 = 0, e2 = 0;
    for (int i = 0; i < check; ++i)
      for (int c = 0; c < 3; ++c) { e1 += (r[i][c] - d[i][c]) * (r[i][c] - d[i][c]); e2 += d[i][c] * d[i][c]; }
    printf("matvec vs Direct (first %d rows): %.3e\n", check, std::sqrt(e1 / e2));
  }
  std::vector<result_type> b(n, result_type(4 * M_PI, 0., 0.));
  std::vector<charge_type> x(n, charge_type(0., 0., 0.));
  tic = get_time();
  plan.kernel().set_p(p);
  GMRESReport rep;
  if (fgmres) {
    // like the reference's driver, the preconditioner's plan is built inside the timed solve (StokesBEM.cpp:305-329)
    std::unique_ptr<FMM_plan<kernel_type>> pc;
    if (pc_local || pc_diagonal) {
      FMMOptions po;                       // LocalPC_Stokes.hpp:8-18 / BlockDiagonalPC_Stokes.hpp local_options()
      po.lazy_evaluation = false;
      po.set_mac_theta(0.5);
      po.sparse_local = true;
      po.local_evaluation = pc_local && !pc_diagonal;
      po.block_diagonal = pc_diagonal;
      po.device = opts.device;
      pc.reset(new FMM_plan<kernel_type>(K, panels, po));
      if (!pc->handle()) return 1;
    }
    printf("Solver: FGMRES, Preconditioner: %s\n", pc_diagonal ? "Block-Diagonal" : (pc_local ? "Local Solve" : "Identity"));
    rep = FGMRES_device(plan, x, b, so, pc.get());
  } else {
    rep = GMRES_device(plan, x, b, so);
  }
  double solve = get_time() - tic;
  double fx = 0, fy = 0, fz = 0;
  for (size_t i = 0; i < n; ++i) { fx += x[i][0] * panels[i].Area; fy += x[i][1] * panels[i].Area; fz += x[i][2] * panels[i].Area; }
  const double exact = 6 * M_PI * mu;
  printf("panels: %zu, mu: %g\nTIMING:\n\tcontext : %.4es\n\tsetup : %.4es\n\tsolve : %.4es\n", n, mu, context_time, setup,
         solve);
  printf("iterations: %d, final residual: %.4e\n", rep.iterations, rep.final_residual);
  printf("Fx: %.5lf, analytical: %.4lg\nFy: %.4g, Fz: %.4g\nerror on a sphere: %.5e\n", fx, exact, fy, fz,
         std::fabs(exact - fx) / std::fabs(exact));
  return 0;
}
