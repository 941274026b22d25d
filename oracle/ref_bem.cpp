// oracle/ref_bem.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the UNMODIFIED reference BEM matvec: LaplaceSphericalBEM K(p,k); panels from
// Triangulation::UnitSphere (reference examples/BEM/Triangulation.hpp:105-121) or from a file of
// 9 doubles per panel; FMM_plan<LaplaceSphericalBEM>(K, panels, opts) with opts.sparse_local as
// examples/LaplaceBEM.cpp:81 sets it; one plan.execute(charges).  Dumps panels, charges and results.
// Build: oracle/Makefile (g++ -fno-access-control, oracle/boost_shim).
//   ref_bem -recursions 4 -P 8 -K 4 -ncrit 64 -theta 0.5 -bc 0 [-rand] [-sparse 1] [-tree] [-direct] [-in file -n N] -dump prefix
// -tree: FMMOptions::TREECODE (what `LaplaceBEM -eval TREE` selects): M2P instead of M2L / L2L / L2P
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
#ifdef YUKAWA_BEM
// -DYUKAWA_BEM: kernel/YukawaCartesianBEM.hpp behind the arity adapter its executor needs (the shipped operators
// carry a trailing `unsigned p`); all arithmetic is the reference's.
#include <YukawaCartesianBEM.hpp>
#include <Triangulation.hpp>
static double g_kappa = 1.0;
class YukawaBEMAdapter : public YukawaCartesianBEM {
 public:
  YukawaBEMAdapter(int p, unsigned k) : YukawaCartesianBEM(p, g_kappa, k) {}
  kernel_value_type operator()(const source_type& t, const target_type& s) const { return YukawaCartesianBEM::operator()(t, s); }
  void init_multipole(multipole_type& M, const point_type& e, unsigned l) const { YukawaCartesianBEM::init_multipole(M, e, l); }
  void init_local(local_type& L, const point_type& e, unsigned l) const { YukawaCartesianBEM::init_local(L, e, l); }
  void P2M(const source_type& s, const charge_type& c, const point_type& ctr, multipole_type& M) const {
    YukawaCartesianBEM::P2M(s, c, ctr, M, (unsigned)P);
  }
  void M2M(const multipole_type& Ms, multipole_type& Mt, const point_type& t) const { YukawaCartesianBEM::M2M(Ms, Mt, t, (unsigned)P); }
  void M2P(const multipole_type& M, const point_type& ctr, const target_type& t, result_type& r) const {
    YukawaCartesianBEM::M2P(M, ctr, t, r, (unsigned)P);
  }
  void M2L(const multipole_type& Ms, local_type& Lt, const point_type& t) const { YukawaCartesianBEM::M2L(Ms, Lt, t, (unsigned)P); }
  void L2L(const local_type& Ls, local_type& Lt, const point_type& t) const { YukawaCartesianBEM::L2L(Ls, Lt, t, (unsigned)P); }
  void L2P(const local_type& L, const point_type& ctr, const target_type& t, result_type& r) const {
    YukawaCartesianBEM::L2P(L, ctr, t, r, (unsigned)P);
  }
};
typedef YukawaBEMAdapter kernel_type;
#else
#include <LaplaceSphericalBEM.hpp>
#include <Triangulation.hpp>

typedef LaplaceSphericalBEM kernel_type;
#endif
typedef kernel_type::point_type point_type;
typedef kernel_type::source_type source_type;
typedef kernel_type::charge_type charge_type;
typedef kernel_type::result_type result_type;

template <typename T>
static void dump(const std::string& path, const std::vector<T>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
  fclose(f);
}

int main(int argc, char** argv) {
  int recursions = 4, P = 8, K = 4, bc = 0, rnd = 0, sparse = 1, direct = 0, n_in = 0, reps = 1, tree = 0;
  unsigned ncrit = 64;
  double theta = 0.5;
  std::string dump_prefix, in_file;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-recursions")) recursions = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-P")) P = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-K")) K = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-ncrit")) ncrit = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-theta")) theta = atof(argv[++i]);
    else if (!strcmp(argv[i], "-bc")) bc = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-rand")) rnd = 1;
    else if (!strcmp(argv[i], "-sparse")) sparse = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-direct")) direct = 1;
    else if (!strcmp(argv[i], "-reps")) reps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-in")) in_file = argv[++i];
    else if (!strcmp(argv[i], "-n")) n_in = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-dump")) dump_prefix = argv[++i];
    else if (!strcmp(argv[i], "-tree")) tree = 1;
#ifdef YUKAWA_BEM
    else if (!strcmp(argv[i], "-kappa")) g_kappa = atof(argv[++i]);
#endif
    else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
  }
  kernel_type Kn(P, K);     // also sets the process-global quadrature order (BEMConfig)
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
  if (bc) for (auto& p : panels) p.switch_BC();
  std::vector<charge_type> charges(n, 1.0);
  if (rnd) for (int i = 0; i < n; ++i) charges[i] = drand48();

  FMMOptions opts;
  opts.set_mac_theta(theta);
  opts.set_max_per_box(ncrit);
  opts.sparse_local = sparse != 0;
  if (tree) opts.evaluator = FMMOptions::TREECODE;
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
  for (int i = 0; i < n; ++i) { sum += res[i]; wsum += res[i] * (i % 7 + 1); }
  double err = -1;
  std::vector<result_type> exact;
  if (direct) {
    exact.assign(n, 0.0);
    Direct::matvec(Kn, panels.begin(), panels.end(), charges.begin(), panels.begin(), panels.end(), exact.begin());
    double e1 = 0, e2 = 0;
    for (int i = 0; i < n; ++i) { e1 += (res[i] - exact[i]) * (res[i] - exact[i]); e2 += exact[i] * exact[i]; }
    err = sqrt(e1 / e2);
  }
  printf("REF_JSON {\"n\": %d, \"P\": %d, \"K\": %d, \"bc\": %d, \"ncrit\": %u, \"theta\": %.17g, \"plan_s\": %.6f, "
         "\"exec_s\": %.6f, \"sum\": %.17g, \"wsum\": %.17g, \"r0\": %.17g, \"err_vs_direct\": %.6e}\n",
         n, P, K, bc, ncrit, theta, t_plan, best, sum, wsum, res[0], err);
  if (!dump_prefix.empty()) {
    std::vector<double> verts(9 * (size_t)n);
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) verts[9 * (size_t)i + 3 * v + k] = panels[i].vertices[v][k];
    dump(dump_prefix + ".verts.f64", verts);
    dump(dump_prefix + ".charges.f64", charges);
    dump(dump_prefix + ".results.f64", res);
    if (direct) dump(dump_prefix + ".direct.f64", exact);
  }
  return 0;
}
