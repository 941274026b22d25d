// oracle/ref_stokes_bem.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the UNMODIFIED reference Stokes BEM matvec: StokesSphericalBEM K(p, k, mu) + set_Kfine(kfine) as
// examples/StokesBEM.cpp:211-214 builds it; panels from Triangulation::UnitSphere
// (examples/BEM/Triangulation.hpp:105-121) or from a file of 9 doubles per panel; all panels VELOCITY (-bc 0, the
// solve of StokesBEM.cpp:279) or TRACTION (-bc 1, its right-hand side :260-264), or mixed (-bc 2: every third panel
// TRACTION); FMM_plan<StokesSphericalBEM>(K, panels, opts) with opts.sparse_local as StokesBEM.cpp:126 sets it; one
// plan.execute(charges).  Dumps panels, boundary conditions, charges, results and optionally Direct::matvec.
// Build: oracle/Makefile (g++ -fno-access-control, oracle/boost_shim).
//   ref_stokes_bem -recursions 4 -P 8 -K 4 -kfine 19 -mu 1e-3 -ncrit 64 -theta 0.5 -bc 0 [-rand] [-sparse 1]
//                  [-tree] [-direct] [-in file -n N] -dump prefix
// -selftest: prints K(t, t) (the Fata analytical self term of the single layer, examples/BEM/FataAnalytical.hpp:
//   414-713 through eval_velocity_integral, kernel/StokesSphericalBEM.hpp:264-275) for the first 64 panels.
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include <deque>
#include <string>
#include <iostream>
#include <algorithm>
#include <fstream>
#include <boost/numeric/ublas/vector.hpp>
using std::isnan;

#include <FMM_plan.hpp>
#include <StokesSphericalBEM.hpp>
#include <Triangulation.hpp>

typedef StokesSphericalBEM kernel_type;
typedef kernel_type::point_type point_type;
typedef kernel_type::source_type source_type;
typedef kernel_type::charge_type charge_type;
typedef kernel_type::result_type result_type;

static void dump(const std::string& path, const std::vector<double>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  if (!v.empty()) fwrite(v.data(), sizeof(double), v.size(), f);
  fclose(f);
}
static std::vector<double> flat(const std::vector<result_type>& v) {
  std::vector<double> o(3 * v.size());
  for (size_t i = 0; i < v.size(); ++i) for (int k = 0; k < 3; ++k) o[3 * i + k] = v[i][k];
  return o;
}

int main(int argc, char** argv) {
  int recursions = 4, P = 8, K = 4, kfine = 19, bc = 0, rnd = 0, sparse = 1, direct = 0, n_in = 0, reps = 1, selftest = 0, tree = 0;
  unsigned ncrit = 64;
  double theta = 0.5, mu = 1e-3;
  std::string dump_prefix, in_file;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-recursions")) recursions = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-P")) P = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-K")) K = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-kfine")) kfine = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-mu")) mu = atof(argv[++i]);
    else if (!strcmp(argv[i], "-ncrit")) ncrit = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-theta")) theta = atof(argv[++i]);
    else if (!strcmp(argv[i], "-bc")) bc = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-rand")) rnd = 1;
    else if (!strcmp(argv[i], "-sparse")) sparse = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-direct")) direct = 1;
    else if (!strcmp(argv[i], "-selftest")) selftest = 1;
    else if (!strcmp(argv[i], "-tree")) tree = 1;
    else if (!strcmp(argv[i], "-reps")) reps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-in")) in_file = argv[++i];
    else if (!strcmp(argv[i], "-n")) n_in = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-dump")) dump_prefix = argv[++i];
    else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
  }
  kernel_type Kn(P, K, mu);     // also sets the process-global quadrature order (BEMConfig)
  Kn.set_Kfine(kfine);
  std::vector<source_type> panels;
  if (in_file.empty()) {
    Triangulation::UnitSphere(panels, recursions);
  } else {
    std::vector<double> buf(9 * (size_t)n_in);
    FILE* f = fopen(in_file.c_str(), "rb");
    if (!f || fread(buf.data(), 8, buf.size(), f) != buf.size()) { fprintf(stderr, "cannot read %s\n", in_file.c_str()); return 2; }
    fclose(f);
    for (int i = 0; i < n_in; ++i)
      panels.push_back(source_type(point_type(buf[9 * i], buf[9 * i + 1], buf[9 * i + 2]),
                                   point_type(buf[9 * i + 3], buf[9 * i + 4], buf[9 * i + 5]),
                                   point_type(buf[9 * i + 6], buf[9 * i + 7], buf[9 * i + 8])));
  }
  const int n = (int)panels.size();
  for (int i = 0; i < n; ++i)
    if (bc == 1 || (bc == 2 && i % 3 == 1)) panels[i].switch_BC();
  if (selftest) {
    for (int i = 0; i < n && i < 64; ++i) {
      auto m = Kn.eval_velocity_integral(panels[i], panels[i]);
      printf("SELF %d", i);
      for (int k = 0; k < 9; ++k) printf(" %.17g", m.vals_[k]);
      printf("\n");
    }
    return 0;
  }
  std::vector<charge_type> charges(n, charge_type(1., 0., 0.));
  if (rnd) for (int i = 0; i < n; ++i) charges[i] = charge_type(drand48(), drand48(), drand48());

  FMMOptions opts;
  opts.set_mac_theta(theta);
  opts.set_max_per_box(ncrit);
  opts.sparse_local = sparse != 0;
  if (tree) opts.evaluator = FMMOptions::TREECODE;     // -tree (`StokesBEM -eval TREE`): M2P instead of M2L / L2L / L2P
  double t0 = get_time();
  FMM_plan<kernel_type> plan(Kn, panels, opts);
  double t_plan = get_time() - t0;
  std::vector<result_type> res;
  double best = 1e300;
  for (int r = 0; r < reps; ++r) {
    t0 = get_time();
    res = plan.execute(charges);
    best = std::min(best, get_time() - t0);
  }
  double sum = 0, wsum = 0;
  for (int i = 0; i < n; ++i) { sum += res[i][0] + res[i][1] + res[i][2]; wsum += res[i][0] * (i % 7 + 1); }
  double err = -1;
  std::vector<result_type> exact;
  if (direct) {
    exact.assign(n, result_type(0.));
    Direct::matvec(Kn, panels.begin(), panels.end(), charges.begin(), panels.begin(), panels.end(), exact.begin());
    double e1 = 0, e2 = 0;
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) { e1 += (res[i][k] - exact[i][k]) * (res[i][k] - exact[i][k]); e2 += exact[i][k] * exact[i][k]; }
    err = sqrt(e1 / e2);
  }
  printf("REF_JSON {\"n\": %d, \"P\": %d, \"K\": %d, \"kfine\": %d, \"mu\": %.17g, \"bc\": %d, \"ncrit\": %u, \"theta\": %.17g, "
         "\"plan_s\": %.6f, \"exec_s\": %.6f, \"sum\": %.17g, \"wsum\": %.17g, \"r0\": [%.17g, %.17g, %.17g], "
         "\"err_vs_direct\": %.6e}\n",
         n, P, K, kfine, mu, bc, ncrit, theta, t_plan, best, sum, wsum, res[0][0], res[0][1], res[0][2], err);
  if (!dump_prefix.empty()) {
    std::vector<double> verts(9 * (size_t)n), bcs(n);
    for (int i = 0; i < n; ++i) {
      bcs[i] = panels[i].BC == source_type::TRACTION ? 1 : 0;
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) verts[9 * (size_t)i + 3 * v + k] = panels[i].vertices[v][k];
    }
    dump(dump_prefix + ".verts.f64", verts);
    dump(dump_prefix + ".bc.f64", bcs);
    dump(dump_prefix + ".charges.f64", flat(charges));
    dump(dump_prefix + ".results.f64", flat(res));
    if (direct) dump(dump_prefix + ".direct.f64", flat(exact));
  }
  return 0;
}
