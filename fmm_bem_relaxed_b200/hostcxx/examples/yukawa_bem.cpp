// YukawaCartesianBEM on the unit sphere (BASELINE config 3), in the manner of the reference's (stale, not a make
// target) examples/YukawaBEM.cpp: first-kind equation for the screened potential, relaxed GMRES on the GPU plan.
//   yukawa_bem -recursions 7 -p 8 -k 4 -kappa 1 -solver_tol 1e-6 [-fixed_p] [-host_gmres] [-check N]
// Manufactured problem: x_exact = du/dn of u = exp(-kappa r) / r on the unit sphere, -(kappa + 1) exp(-kappa);
// b = A_G x_exact is formed with the plan itself, then  A_G x = b  is solved from x = 0 and |x - x_exact| reported.
// -check N also compares the first N rows of the GPU matvec with Direct::matvec of the host kernel class.
#include <FMM_plan.hpp>
#include <YukawaCartesianBEM.hpp>
#include <Triangulation.hpp>
#include <GMRES.hpp>

#include <algorithm>
#include <cmath>
#include <cstring>

int main(int argc, char** argv) {
  int recursions = 5, p = 8, k = 4, check = 0;
  double kappa = 1.0;
  bool host_gmres = false;
  FMMOptions opts = get_options(argc, argv);
  SolverOptions so;
  so.residual = 1e-6;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-recursions")) recursions = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-p")) p = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-k")) k = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-kappa")) kappa = atof(argv[++i]);
    else if (!strcmp(argv[i], "-solver_tol")) so.residual = atof(argv[++i]);
    else if (!strcmp(argv[i], "-fixed_p")) so.variable_p = false;
    else if (!strcmp(argv[i], "-host_gmres")) host_gmres = true;
    else if (!strcmp(argv[i], "-check")) check = atoi(argv[++i]);
  }
  so.max_p = p;
  so.max_iters = so.restart = 200;
  typedef YukawaCartesianBEM kernel_type;
  kernel_type K(p, kappa, k);
  std::vector<kernel_type::source_type> panels;
  Triangulation::UnitSphere(panels, recursions);
  const size_t n = panels.size();

  double tic = get_time();
  FMM_plan<kernel_type> plan(K, panels, opts);
  const double dudn = -(kappa + 1) * std::exp(-kappa);
  std::vector<double> exact(n, dudn), x(n, 0.);
  std::vector<double> b = plan.execute(exact);
  double setup = get_time() - tic;
  if (b.empty()) return 1;
  if (check > 0) {   // the GPU matvec against Direct::matvec with the host kernel class on the first rows
    check = std::min<int>(check, (int)n);
    std::vector<kernel_type::target_type> tg(panels.begin(), panels.begin() + check);
    std::vector<double> d(check, 0.);
    Direct::matvec(K, panels.begin(), panels.end(), exact.begin(), tg.begin(), tg.end(), d.begin());
    double e1 = 0, e2 = 0;
    for (int i = 0; i < check; ++i) { e1 += (b[i] - d[i]) * (b[i] - d[i]); e2 += d[i] * d[i]; }
    printf("matvec vs Direct (first %d rows): %.3e\n", check, std::sqrt(e1 / e2));
  }
  tic = get_time();
  plan.kernel().set_p(p);
  GMRESReport rep = host_gmres ? GMRES(plan, x, b, so) : GMRES_device(plan, x, b, so);
  double solve = get_time() - tic;
  double e = 0;
  for (size_t i = 0; i < n; ++i) e += (x[i] - dudn) * (x[i] - dudn);
  printf("panels: %zu, kappa: %g\nTIMING:\n\tsetup : %.4es\n\tsolve : %.4es\n", n, kappa, setup, solve);
  printf("iterations: %d, final residual: %.4e, relative error of the solution: %.3e\n", rep.iterations,
         rep.final_residual, std::sqrt(e / n) / std::fabs(dudn));
  return 0;
}
