// oracle/ref_yukawa.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the reference's YukawaCartesian FMM (kernel/YukawaCartesian.hpp) through FMM_plan.
// As shipped the kernel is NOT callable by the reference's executor: every expansion operator carries a trailing
// `unsigned p` argument that include/KernelTraits.hpp:134-183 does not recognise, so
// `FMM_plan<YukawaCartesian>` prints "[W] Cannot use Kernel for FMM!".  The adapter below is the minimal glue the
// survey describes: it derives from the UNMODIFIED reference class and re-declares the operators with the arity
// the executor expects, forwarding p = P.  All arithmetic is the reference's.
// Inputs: glibc drand48 default state (N points, then N charges) or -in file (3N coordinates then N charges).
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
#include <boost/numeric/ublas/vector.hpp>
using std::isnan;

#include <FMM_plan.hpp>
#include <YukawaCartesian.hpp>

class YukawaAdapter : public YukawaCartesian {
 public:
  YukawaAdapter(int p, double kappa) : YukawaCartesian(p, kappa) {}
  // inherited members are invisible to the member-pointer SFINAE of KernelTraits.hpp: re-declare, forward
  kernel_value_type operator()(const point_type& t, const point_type& s) const { return YukawaCartesian::operator()(t, s); }
  void init_multipole(multipole_type& M, const point_type& extents, unsigned level) const {
    YukawaCartesian::init_multipole(M, extents, level);
  }
  void init_local(local_type& L, const point_type& extents, unsigned level) const {
    YukawaCartesian::init_local(L, extents, level);
  }
  void P2M(const source_type& s, const charge_type& c, const point_type& ctr, multipole_type& M) const {
    YukawaCartesian::P2M(s, c, ctr, M, (unsigned)P);
  }
  void M2M(const multipole_type& Ms, multipole_type& Mt, const point_type& t) const {
    YukawaCartesian::M2M(Ms, Mt, t, (unsigned)P);
  }
  void M2P(const multipole_type& M, const point_type& ctr, const target_type& t, result_type& r) const {
    YukawaCartesian::M2P(M, ctr, t, r, (unsigned)P);
  }
  void M2L(const multipole_type& Ms, local_type& Lt, const point_type& t) const {
    YukawaCartesian::M2L(Ms, Lt, t, (unsigned)P);
  }
  void L2L(const local_type& Ls, local_type& Lt, const point_type& t) const {
    YukawaCartesian::L2L(Ls, Lt, t, (unsigned)P);
  }
  void L2P(const local_type& L, const point_type& ctr, const target_type& t, result_type& r) const {
    YukawaCartesian::L2P(L, ctr, t, r, (unsigned)P);
  }
};

typedef YukawaAdapter kernel_type;
typedef kernel_type::point_type point_type;
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
  int N = 10000, P = 8, ndirect = 0, reps = 1;
  unsigned ncrit = 64;
  double theta = 0.5, kappa = 0.125;
  bool tree = false;
  std::string dump_prefix, in_file;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-N")) N = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-P")) P = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-ncrit")) ncrit = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-theta")) theta = atof(argv[++i]);
    else if (!strcmp(argv[i], "-kappa")) kappa = atof(argv[++i]);
    else if (!strcmp(argv[i], "-direct")) ndirect = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-reps")) reps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-tree")) tree = true;
    else if (!strcmp(argv[i], "-in")) in_file = argv[++i];
    else if (!strcmp(argv[i], "-dump")) dump_prefix = argv[++i];
    else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
  }
  std::vector<point_type> points(N);
  std::vector<charge_type> charges(N);
  if (in_file.empty()) {
    for (int k = 0; k < N; ++k) points[k] = point_type(drand48(), drand48(), drand48());
    for (int k = 0; k < N; ++k) charges[k] = drand48();
  } else {
    std::vector<double> buf(4 * (size_t)N);
    FILE* f = fopen(in_file.c_str(), "rb");
    if (!f || fread(buf.data(), 8, buf.size(), f) != buf.size()) { fprintf(stderr, "cannot read %s\n", in_file.c_str()); return 2; }
    fclose(f);
    for (int k = 0; k < N; ++k) {
      points[k] = point_type(buf[3 * k], buf[3 * k + 1], buf[3 * k + 2]);
      charges[k] = buf[3 * (size_t)N + k];
    }
  }
  kernel_type K(P, kappa);
  FMMOptions opts;
  opts.set_mac_theta(theta);
  opts.set_max_per_box(ncrit);
  if (tree) opts.evaluator = FMMOptions::TREECODE;
  double t0 = get_time();
  FMM_plan<kernel_type> plan(K, points, opts);
  double t_plan = get_time() - t0;
  std::vector<result_type> res;
  double best = 1e300;
  for (int r = 0; r < reps; ++r) {
    t0 = get_time();
    res = plan.execute(charges);
    best = std::min(best, get_time() - t0);
  }
  double pot = 0, fxw = 0;
  for (int k = 0; k < N; ++k) { pot += res[k][0]; fxw += res[k][1] * (k % 7 + 1); }
  double e_pot = -1, e_force = -1;
  if (ndirect > 0) {
    ndirect = std::min(ndirect, N);
    std::vector<point_type> tg(points.begin(), points.begin() + ndirect);
    std::vector<result_type> exact(ndirect);
    Direct::matvec(K, points, charges, tg, exact);
    double a = 0, b = 0, c = 0, d = 0;
    for (int k = 0; k < ndirect; ++k) {
      a += (res[k][0] - exact[k][0]) * (res[k][0] - exact[k][0]); b += exact[k][0] * exact[k][0];
      for (int j = 1; j < 4; ++j) { c += (res[k][j] - exact[k][j]) * (res[k][j] - exact[k][j]); d += exact[k][j] * exact[k][j]; }
    }
    e_pot = sqrt(a / b); e_force = sqrt(c / d);
  }
  printf("REF_JSON {\"N\": %d, \"P\": %d, \"ncrit\": %u, \"theta\": %.17g, \"kappa\": %.17g, \"treecode\": %d, \"plan_s\": %.6f, "
         "\"exec_s\": %.6f, \"pot\": %.17g, \"fxw\": %.17g, \"r0\": [%.17g, %.17g, %.17g, %.17g], \"err_pot\": %.6e, "
         "\"err_force\": %.6e}\n",
         N, P, ncrit, theta, kappa, (int)tree, t_plan, best, pot, fxw, res[0][0], res[0][1], res[0][2], res[0][3], e_pot, e_force);
  if (!dump_prefix.empty()) {
    std::vector<double> in(4 * (size_t)N), out(4 * (size_t)N);
    for (int k = 0; k < N; ++k) {
      for (int c = 0; c < 3; ++c) in[3 * (size_t)k + c] = points[k][c];
      in[3 * (size_t)N + k] = charges[k];
      for (int c = 0; c < 4; ++c) out[4 * (size_t)k + c] = res[k][c];
    }
    dump(dump_prefix + ".input.f64", in);
    dump(dump_prefix + ".results.f64", out);
  }
  return 0;
}
