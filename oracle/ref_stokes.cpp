// oracle/ref_stokes.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the reference's StokesSpherical FMM like serialrun_stresslet.cpp:98-128 does.
//   default build  : Stokeslet, charge Vec<3> (unmodified reference; compiles as shipped)
//   -DSTRESSLET    : stresslet, charge Vec<6> = (g, n).  The shipped reference does not compile in this mode
//                    (kernel/StokesSpherical.hpp:177-178 assigns complex to double; the sparse evaluators do not
//                    instantiate for a kernel without operator()), so oracle/Makefile builds this variant from a
//                    PATCHED TEMPORARY COPY with exactly the two patches SURVEY.md section 8(c) lists.  Parity for
//                    the stresslet is therefore stated against the patched reference.
// Inputs: glibc drand48 default state, N points then charges (3 draws per body; stresslet: n = (1,0,0)), or -in file
// (3N coordinates then CD*N charge entries).  Dumps charges and results; optional brute force with the kernel's own
// P2P / operator().
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
#include <StokesSpherical.hpp>

typedef StokesSpherical kernel_type;
typedef kernel_type::point_type point_type;
typedef kernel_type::charge_type charge_type;
typedef kernel_type::result_type result_type;
#ifdef STRESSLET
static const int CD = 6;
#else
static const int CD = 3;
#endif

template <typename T>
static void dump(const std::string& path, const std::vector<T>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
  fclose(f);
}

int main(int argc, char** argv) {
  int N = 10000, P = 8, ndirect = 0, reps = 1, tree = 0;
  unsigned ncrit = 64;
  double theta = 0.5;
  std::string dump_prefix, in_file;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-N")) N = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-P")) P = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-ncrit")) ncrit = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-theta")) theta = atof(argv[++i]);
    else if (!strcmp(argv[i], "-direct")) ndirect = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-reps")) reps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-in")) in_file = argv[++i];
    else if (!strcmp(argv[i], "-dump")) dump_prefix = argv[++i];
    else if (!strcmp(argv[i], "-tree")) tree = 1;
    else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
  }
  std::vector<point_type> points(N);
  std::vector<charge_type> charges(N);
  if (in_file.empty()) {
    for (int k = 0; k < N; ++k) points[k] = point_type(drand48(), drand48(), drand48());
    for (int k = 0; k < N; ++k) {
      charges[k][0] = drand48(); charges[k][1] = drand48(); charges[k][2] = drand48();
#ifdef STRESSLET
      charges[k][3] = 1.; charges[k][4] = 0.; charges[k][5] = 0.;
#endif
    }
  } else {
    std::vector<double> buf((3 + CD) * (size_t)N);
    FILE* f = fopen(in_file.c_str(), "rb");
    if (!f || fread(buf.data(), 8, buf.size(), f) != buf.size()) { fprintf(stderr, "cannot read %s\n", in_file.c_str()); return 2; }
    fclose(f);
    for (int k = 0; k < N; ++k) {
      points[k] = point_type(buf[3 * k], buf[3 * k + 1], buf[3 * k + 2]);
      for (int c = 0; c < CD; ++c) charges[k][c] = buf[3 * (size_t)N + CD * (size_t)k + c];
    }
  }
  kernel_type K(P);
  FMMOptions opts;
  opts.set_mac_theta(theta);
  opts.set_max_per_box(ncrit);
  if (tree) opts.evaluator = FMMOptions::TREECODE;     // -tree: M2P instead of M2L / L2L / L2P
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
  double sum[3] = {0, 0, 0};
  for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c) sum[c] += res[k][c];
  double err = -1;
  if (ndirect > 0) {
    ndirect = std::min(ndirect, N);
    std::vector<point_type> tg(points.begin(), points.begin() + ndirect);
    std::vector<result_type> exact(ndirect);
    Direct::matvec(K, points.begin(), points.end(), charges.begin(), tg.begin(), tg.end(), exact.begin());
    double e1 = 0, e2 = 0;
    for (int k = 0; k < ndirect; ++k)
      for (int c = 0; c < 3; ++c) { e1 += (res[k][c] - exact[k][c]) * (res[k][c] - exact[k][c]); e2 += exact[k][c] * exact[k][c]; }
    err = sqrt(e1 / e2);
  }
  printf("REF_JSON {\"N\": %d, \"P\": %d, \"ncrit\": %u, \"theta\": %.17g, \"stresslet\": %d, \"plan_s\": %.6f, \"exec_s\": %.6f, "
         "\"sum\": [%.17g, %.17g, %.17g], \"r0\": [%.17g, %.17g, %.17g], \"err_vs_direct\": %.6e}\n",
         N, P, ncrit, theta, CD == 6, t_plan, best, sum[0], sum[1], sum[2], res[0][0], res[0][1], res[0][2], err);
  if (!dump_prefix.empty()) {
    std::vector<double> in((3 + CD) * (size_t)N), out(3 * (size_t)N);
    for (int k = 0; k < N; ++k) {
      for (int c = 0; c < 3; ++c) { in[3 * (size_t)k + c] = points[k][c]; out[3 * (size_t)k + c] = res[k][c]; }
      for (int c = 0; c < CD; ++c) in[3 * (size_t)N + CD * (size_t)k + c] = charges[k][c];
    }
    dump(dump_prefix + ".input.f64", in);
    dump(dump_prefix + ".results.f64", out);
  }
  return 0;
}
