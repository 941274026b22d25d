// CPU harness for the solver host logic: runs GMRES(MV, x, b, opts[, M]) on a small dense, well-conditioned system
// whose "kernel" records the expansion orders it is asked for.  Compiled twice by tests/test_host_logic.py:
//   -DUSE_REFERENCE -I /root/reference/examples/BEM   -> the reference's own examples/BEM/GMRES.hpp
//   (default)        -I fmm_bem_relaxed_b200/hostcxx   -> the mirror
// Both must print the same lines (iteration count, residuals, requested orders).  No GPU, no FMM.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <Vec.hpp>
#include <GMRES.hpp>

struct RecordingKernel {
  std::vector<int> orders;
  int p = 0;
  void set_p(int q) { p = q; orders.push_back(q); }
};

struct DenseMatvec {
  typedef double charge_type;
  typedef double result_type;
  int n;
  std::vector<double> A;
  RecordingKernel K;
  explicit DenseMatvec(int n_) : n(n_), A((size_t)n_ * n_) {
    // diagonally dominant, non-symmetric; the perturbation shrinks with the order the solver asked for, like an FMM
    // matvec whose accuracy follows p (so that relaxation has something to act on)
    srand48(7);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = (i == j ? 4.0 + 0.1 * (i % 5) : 0.0) + 0.5 * drand48() / std::sqrt((double)n);
  }
  RecordingKernel& kernel() { return K; }
  std::vector<double> execute(const std::vector<double>& x) {
    std::vector<double> y(n, 0.0);
    const double noise = K.p > 0 ? std::ldexp(1.0, -3 * K.p) : 0.0;
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int j = 0; j < n; ++j) s += A[(size_t)i * n + j] * x[j];
      y[i] = s * (1.0 + noise * ((i % 3) - 1));
    }
    return y;
  }
};

int main(int argc, char** argv) {
  int n = 200;
  SolverOptions opts;
  opts.residual = 1e-8;
  opts.max_p = 12;
  bool diag = false;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-n")) n = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-tol")) opts.residual = atof(argv[++i]);
    else if (!strcmp(argv[i], "-restart")) opts.restart = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-fixed_p")) opts.variable_p = false;
    else if (!strcmp(argv[i], "-max_p")) opts.max_p = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-diagonal")) diag = true;
  }
  DenseMatvec MV(n);
  std::vector<double> x(n, 0.0), b(n);
  for (int i = 0; i < n; ++i) b[i] = 1.0 + 0.01 * i;
  if (diag) {
    struct Jacobi {
      std::vector<double> r;
      void operator()(const std::vector<double>& v, std::vector<double>& z) const {
        z.resize(v.size());
        for (size_t i = 0; i < v.size(); ++i) z[i] = r[i] * v[i];
      }
    } M;
    for (int i = 0; i < n; ++i) M.r.push_back(1.0 / MV.A[(size_t)i * n + i]);
    GMRES(MV, x, b, opts, M);
  } else {
    GMRES(MV, x, b, opts);
  }
  printf("orders:");
  for (int p : MV.K.orders) printf(" %d", p);
  double s = 0;
  for (double v : x) s += v;
  printf("\nsum(x) = %.12e\n", s);
  return 0;
}
