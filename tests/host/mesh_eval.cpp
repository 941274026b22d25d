// CPU harness: Triangulation::UnitSphere (octahedron subdivision) and the panel geometry built from it, printed with
// 17 digits.  Compiled against the mirror (hostcxx/Triangulation.hpp, LaplaceSphericalBEM.hpp) and against the
// reference's examples/BEM/Triangulation.hpp + kernel/LaplaceSphericalBEM.hpp; the outputs must be identical, so that a
// driver sees the same mesh, the same panel order and the same quadrature points.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <Vec.hpp>
#include <LaplaceSphericalBEM.hpp>
#include <Triangulation.hpp>

int main(int argc, char** argv) {
  int rec = argc > 1 ? atoi(argv[1]) : 4, k = argc > 2 ? atoi(argv[2]) : 4;
  LaplaceSphericalBEM K(5, k);
  std::vector<LaplaceSphericalBEM::Panel> panels;
  Triangulation::UnitSphere(panels, rec);
  printf("%zu\n", panels.size());
  for (auto& p : panels) {
    for (int v = 0; v < 3; ++v) printf("%.17g %.17g %.17g ", p.vertices[v][0], p.vertices[v][1], p.vertices[v][2]);
    printf("| %.17g %.17g %.17g | %.17g %.17g %.17g | %.17g |", p.center[0], p.center[1], p.center[2], p.normal[0],
           p.normal[1], p.normal[2], p.Area);
    for (auto& q : p.quad_points) printf(" %.17g %.17g %.17g", q[0], q[1], q[2]);
    printf("\n");
  }
  return 0;
}
