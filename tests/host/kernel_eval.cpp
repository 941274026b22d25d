// CPU harness for the host-side kernel classes: evaluates K(t, s) of one kernel class on a fixed pseudo-random set of
// panels / points and prints every value with 17 digits.  Compiled by tests/test_host_logic.py against the mirror
// (fmm_bem_relaxed_b200/hostcxx) and against the reference's own kernel headers; the outputs must agree.
// The BEM panel integrals printed here are the same source (hostcxx/bem_math.hpp) the GPU near-field assembly runs.
//   -DKERNEL=1 LaplaceSpherical   2 LaplaceSphericalBEM   3 YukawaCartesian   4 YukawaCartesianBEM
//            5 StokesSpherical (Stokeslet: 3x3 kernel value)   6 StokesSphericalBEM (3x3 blocks, both layers; lines
//            of a panel with itself are tagged SELF: the mirror evaluates that closed form by another route)
//            7 the triangle Gauss rules behind every BEM class, all keys of the reference's table
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <Vec.hpp>
#if KERNEL == 1
#include <LaplaceSpherical.hpp>
typedef LaplaceSpherical kernel_type;
#elif KERNEL == 2
#include <LaplaceSphericalBEM.hpp>
typedef LaplaceSphericalBEM kernel_type;
#elif KERNEL == 3
#include <YukawaCartesian.hpp>
typedef YukawaCartesian kernel_type;
#elif KERNEL == 4
#include <YukawaCartesianBEM.hpp>
typedef YukawaCartesianBEM kernel_type;
#elif KERNEL == 6
#include <StokesSphericalBEM.hpp>
typedef StokesSphericalBEM kernel_type;
#elif KERNEL == 7
#if __has_include(<GaussQuadrature.hpp>)
#include <GaussQuadrature.hpp>     // the reference's table (examples/BEM)
#define REF_RULES 1
#else
#include <bem_math.hpp>            // the mirror's table
#endif
#include <LaplaceSpherical.hpp>
typedef LaplaceSpherical kernel_type;
#else
#include <StokesSpherical.hpp>
typedef StokesSpherical kernel_type;
#endif

static unsigned long long lcg = 12345;
static double rnd() { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(lcg >> 11) / 9007199254740992.0; }

int main() {
  typedef kernel_type::point_type point_type;
#if KERNEL == 1
  kernel_type K(5);
#elif KERNEL == 2
  kernel_type K(5, 4);
#elif KERNEL == 3
  kernel_type K(5, 0.75);
#elif KERNEL == 4
  kernel_type K(5, 0.75, 4);
#elif KERNEL == 6
  kernel_type K(5, 4, 1e-3);
  K.set_Kfine(19);
#ifdef AS_WRITTEN   // mirror only: the branches of the reference's source text (compared with the one-token-patched reference)
  K.near_field_as_written = true;
#endif
#else
  kernel_type K(5);
#endif
#if KERNEL == 7
  const int keys[] = {1, 3, 4, 7, 13, 17, 19, 25, 79};
  for (int k : keys) {
#ifdef REF_RULES
    GaussQuadrature<double> GQ;
    auto& pts = GQ.points(k);
    auto& w = GQ.weights(k);
    printf("rule %d: %zu points\n", k, pts.size());
    for (size_t i = 0; i < pts.size(); ++i) printf("%.17g %.17g %.17g %.17g\n", pts[i][0], pts[i][1], pts[i][2], w[i]);
#else
    const bem::Rule r = bem::make_rule(k);
    printf("rule %d: %zu points\n", k, (size_t)r.n);
    for (int i = 0; i < r.n; ++i) printf("%.17g %.17g %.17g %.17g\n", r.pt[i][0], r.pt[i][1], r.pt[i][2], r.w[i]);
#endif
  }
  (void)K;
#elif KERNEL == 6
  typedef kernel_type::Panel Panel;
  std::vector<Panel> pan;
  for (int i = 0; i < 40; ++i) {
    point_type c(0.5 * rnd(), 0.5 * rnd(), 0.1 * rnd());
    point_type a = c + point_type(0.03 * rnd(), 0.03 * rnd(), 0.01 * rnd());
    point_type b = c + point_type(-0.03 * rnd(), 0.03 * rnd(), 0.01 * rnd());
    pan.push_back(Panel(c, a, b));
  }
  for (int bc = 0; bc < 2; ++bc) {
    for (size_t i = 0; i < pan.size(); ++i)
      for (size_t j = 0; j < pan.size(); ++j) {
        auto v = K(pan[i], pan[j]);
        printf("%s%d %zu %zu", (i == j && bc == 0) ? "SELF " : "", bc, i, j);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) printf(" %.17g", (double)v(a, b));
        printf("\n");
      }
    for (auto& p : pan) p.switch_BC();
  }
#elif KERNEL == 2 || KERNEL == 4
  // small triangles on a patch: neighbours fall into the semi-analytical branch, distant ones into the Gauss branch
  typedef kernel_type::Panel Panel;
  std::vector<Panel> pan;
  for (int i = 0; i < 40; ++i) {
    point_type c(0.5 * rnd(), 0.5 * rnd(), 0.1 * rnd());
    point_type a = c + point_type(0.03 * rnd(), 0.03 * rnd(), 0.01 * rnd());
    point_type b = c + point_type(-0.03 * rnd(), 0.03 * rnd(), 0.01 * rnd());
    pan.push_back(Panel(c, a, b));
  }
  for (int bc = 0; bc < 2; ++bc) {
    for (auto& p : pan) p.switch_BC();
    for (size_t i = 0; i < pan.size(); ++i)
      for (size_t j = 0; j < pan.size(); ++j) printf("%d %zu %zu %.17g\n", bc, i, j, (double)K(pan[i], pan[j]));
  }
#else
  std::vector<point_type> pts;
  for (int i = 0; i < 60; ++i) pts.push_back(point_type(rnd(), rnd(), rnd()));
  pts.push_back(pts[3]);                        // a coincident pair: the kernels return 0 there
  for (size_t i = 0; i < pts.size(); ++i)
    for (size_t j = 0; j < pts.size(); ++j) {
      auto v = K(pts[i], pts[j]);
#if KERNEL == 5
      printf("%zu %zu", i, j);
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) printf(" %.17g", (double)v(a, b));
      printf("\n");
#else
      printf("%zu %zu %.17g %.17g %.17g %.17g\n", i, j, v[0], v[1], v[2], v[3]);
#endif
    }
#endif
  return 0;
}
