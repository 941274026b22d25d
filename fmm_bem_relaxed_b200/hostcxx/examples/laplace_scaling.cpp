// Own driver in the style of the reference's tests/scaling.cpp: N drand48 points and charges,
// FMM_plan<LaplaceSpherical>, three timed executes, error of the force against Direct::matvec on the
// first M targets.  Usage: laplace_scaling [N] [P] [ncrit] [M]
#include <FMM_plan.hpp>
#include <LaplaceSpherical.hpp>

#include <cmath>
#include <cstdlib>
#include <numeric>

int main(int argc, char** argv) {
  typedef LaplaceSpherical kernel_type;
  typedef kernel_type::point_type point_type;
  typedef kernel_type::charge_type charge_type;
  typedef kernel_type::result_type result_type;
  int N = argc > 1 ? atoi(argv[1]) : 100000, P = argc > 2 ? atoi(argv[2]) : 5;
  unsigned ncrit = argc > 3 ? atoi(argv[3]) : 64;
  int M = argc > 4 ? atoi(argv[4]) : 1000;
  if (M > N) M = N;
  kernel_type K(P);
  FMMOptions opts;
  opts.set_mac_theta(.5);
  opts.set_max_per_box(ncrit);

  std::vector<point_type> points(N);
  for (int k = 0; k < N; ++k) {
    double z = drand48(), y = drand48(), x = drand48();   // the order g++ evaluates the reference's call in
    points[k] = point_type(x, y, z);
  }
  std::vector<charge_type> charges(N);
  for (int k = 0; k < N; ++k) charges[k] = drand48();

  double tic = get_time();
  FMM_plan<kernel_type> plan(K, points, opts);
  double t_plan = get_time() - tic;
  std::vector<result_type> result;
  std::vector<double> timings(3);
  for (int i = 0; i < 3; ++i) {
    tic = get_time();
    result = plan.execute(charges);
    timings[i] = get_time() - tic;
  }
  if (result.empty()) return 1;
  std::cout << "plan construction time: " << t_plan << std::endl;
  std::cout << "FMM execution time: " << std::accumulate(timings.begin(), timings.end(), 0.0) / 3 << std::endl;

  std::vector<point_type> targets(points.begin(), points.begin() + M);
  std::vector<result_type> exact(M);
  Direct::matvec(K, points, charges, targets, exact);
  double e1 = 0, e2 = 0, p1 = 0, p2 = 0;
  for (int k = 0; k < M; ++k) {
    p1 += (result[k][0] - exact[k][0]) * (result[k][0] - exact[k][0]);
    p2 += exact[k][0] * exact[k][0];
    for (int m = 1; m < 4; ++m) {
      e1 += (result[k][m] - exact[k][m]) * (result[k][m] - exact[k][m]);
      e2 += exact[k][m] * exact[k][m];
    }
  }
  printf("the l2-norm of the relative error of Laplace potential: %.6e\n", sqrt(p1 / p2));
  printf("the l2-norm of the relative error of Laplace force: %.6e\n", sqrt(e1 / e2));
  printf("checksum pot %.17g\n", [&] { double s = 0; for (auto& r : result) s += r[0]; return s; }());
  return 0;
}
