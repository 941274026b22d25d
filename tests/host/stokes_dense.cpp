// CPU harness: the flow-past-a-sphere problem of the reference's examples/StokesBEM.cpp (:229-279) solved with the
// REFERENCE's examples/BEM/GMRES_Stokes.hpp over a dense matvec built from the kernel class's operator() -- no FMM, no
// GPU.  Compiled twice by tests/test_host_logic.py:
//   -I fmm_bem_relaxed_b200/hostcxx -I <ref>/examples/BEM     -> the mirror class hostcxx/StokesSphericalBEM.hpp
//   -I <ref>/kernel -I <ref>/include -I <ref>/examples/BEM ... -> the reference's own kernel/StokesSphericalBEM.hpp
// Both must print the same lines; with 128 panels every pair is in the near field of the reference's FMM, so they are
// also the lines examples/StokesBEM.cpp itself prints for -recursions 3 (checked by the test).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <Vec.hpp>
#include <StokesSphericalBEM.hpp>
#include <Triangulation.hpp>
#include <SolverOptions.hpp>
#include <GMRES_Stokes.hpp>

struct DenseStokes {
  typedef StokesSphericalBEM kernel_type;
  typedef kernel_type::charge_type charge_type;
  typedef kernel_type::result_type result_type;
  kernel_type K;
  std::vector<kernel_type::source_type> panels;
  DenseStokes(const kernel_type& k, const std::vector<kernel_type::source_type>& p) : K(k), panels(p) {}
  kernel_type& kernel() { return K; }
  std::vector<result_type> execute(const std::vector<charge_type>& x) {
    std::vector<result_type> r(panels.size(), result_type(0.));
    for (size_t i = 0; i < panels.size(); ++i)
      for (size_t j = 0; j < panels.size(); ++j) r[i] += K(panels[i], panels[j]) * x[j];
    return r;
  }
};

int main(int argc, char** argv) {
  int recursions = 3, p = 8, k = 4;
  double mu = 1e-3;
  SolverOptions so;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-recursions")) recursions = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-k")) k = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-mu")) mu = atof(argv[++i]);
    else if (!strcmp(argv[i], "-solver_tol")) so.residual = atof(argv[++i]);
    else if (!strcmp(argv[i], "-fixed_p")) so.variable_p = false;
  }
  so.max_p = p; so.p_min = 5; so.max_iters = 100; so.restart = 100;
  typedef StokesSphericalBEM kernel_type;
  kernel_type K(p, k, mu);
  K.set_Kfine(19);
  std::vector<kernel_type::source_type> panels;
  Triangulation::UnitSphere(panels, recursions);
  DenseStokes MV(K, panels);
  std::vector<kernel_type::charge_type> x(panels.size(), kernel_type::charge_type(0., 0., 0.));
  std::vector<kernel_type::result_type> b(panels.size(), kernel_type::result_type(4 * M_PI, 0., 0.));
  GMRES(MV, x, b, so);
  double fx = 0;
  for (size_t i = 0; i < panels.size(); ++i) fx += x[i][0] * panels[i].Area;
  printf("Fx: %.5lf, analytical: %.4lg\n", fx, 6 * M_PI * mu);
  for (size_t i = 0; i < panels.size(); i += 17) printf("x[%zu] = %.12e %.12e %.12e\n", i, x[i][0], x[i][1], x[i][2]);
  return 0;
}
