#pragma once
/** @file Triangulation.hpp
 * Unit sphere by octahedron subdivision: 8 * 4^(recursions-1) panels, same vertices and panel order as
 * the reference generator (examples/BEM/Triangulation.hpp:36-121) so that a driver sees the same mesh.
 * (The reference also writes test.vert / test.face as a side effect; this one does not.)
 */
#include <cstdio>
#include <vector>
#include <Vec.hpp>

namespace Triangulation {

template <typename PanelType>
void UnitSphere(std::vector<PanelType>& panels, unsigned recursions = 2) {
  typedef Vec<3, double> vertex;
  struct tri { vertex a, b, c; };
  const double ov[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  const int ot[8][3] = {{0, 4, 2}, {2, 4, 1}, {1, 4, 3}, {3, 4, 0}, {0, 2, 5}, {2, 1, 5}, {1, 3, 5}, {3, 0, 5}};
  std::vector<tri> t(8);
  auto vtx = [&](int i) { return vertex(ov[i][0], ov[i][1], ov[i][2]); };
  for (int i = 0; i < 8; ++i) t[i] = tri{vtx(ot[i][0]), vtx(ot[i][1]), vtx(ot[i][2])};
  for (unsigned r = 0; r + 1 < recursions; ++r) {
    std::vector<tri> next;
    next.reserve(t.size() * 4);
    for (const tri& s : t) {
      vertex a = (s.a + s.c) * 0.5, b = (s.a + s.b) * 0.5, c = (s.b + s.c) * 0.5;
      a /= norm(a); b /= norm(b); c /= norm(c);
      next.push_back(tri{s.a, b, a});
      next.push_back(tri{b, s.b, c});
      next.push_back(tri{a, b, c});
      next.push_back(tri{a, c, s.c});
    }
    t.swap(next);
  }
  printf("initialised %d triangles\n", (int)t.size());
  panels.clear();
  panels.reserve(t.size());
  for (const tri& s : t) panels.push_back(PanelType(s.a, s.b, s.c));
}

}  // namespace Triangulation
