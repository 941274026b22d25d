#pragma once
/** @file LaplaceSphericalBEM.hpp
 * Host-side kernel class with the reference's public surface
 * (reference kernel/LaplaceSphericalBEM.hpp:14-157,273-297): nested Panel (center, normal, vertices,
 * quad_points, Area, BC, switch_BC, conversion to point_type), LaplaceSphericalBEM(int p, unsigned k),
 * set_p, operator()(target, source).  The expansion operators run on the GPU behind FMM_plan
 * (csrc/bem.cu); operator() here serves Direct::matvec checks and the diagonal preconditioner and
 * shares its panel integrals with the device code (bem_math.hpp).
 */
#include <cmath>
#include <vector>
#include <Vec.hpp>

#include "bem_math.hpp"
#include "LaplaceSpherical.hpp"

class LaplaceSphericalBEM : public LaplaceSpherical {
 public:
  unsigned K;  //!< quadrature points per panel
  struct Panel;
  static constexpr unsigned dimension = LaplaceSpherical::dimension;
  typedef LaplaceSpherical::point_type point_type;
  typedef Panel source_type;
  typedef Panel target_type;
  typedef LaplaceSpherical::charge_type charge_type;
  typedef double kernel_value_type;
  typedef double result_type;
  typedef Panel panel_type;

  static constexpr int fmmb_kind = FMMB_LAPLACE_SPHERICAL_BEM;
  static constexpr int charge_dim = 1;
  static constexpr int result_dim = 1;

  //! Boundary element
  struct Panel {
    typedef enum { POTENTIAL, NORMAL_DERIV } BoundaryType;
    point_type center;
    point_type normal;
    std::vector<point_type> vertices;
    std::vector<point_type> quad_points;
    double Area;
    BoundaryType BC;

    Panel() : Area(0), BC(POTENTIAL) {}
    Panel(point_type p0, point_type p1, point_type p2) : BC(POTENTIAL) {
      vertices.resize(3);
      vertices[0] = p0; vertices[1] = p1; vertices[2] = p2;
      bem::Panel g;
      bem::make_panel(p0.data(), p1.data(), p2.data(), g);
      center = point_type(g.c[0], g.c[1], g.c[2]);
      normal = point_type(g.nrm[0], g.nrm[1], g.nrm[2]);
      Area = g.area;
      // quadrature points of the process-wide rule (the reference keeps K in a BEMConfig singleton)
      const bem::Rule r = bem::make_rule(global_K());
      quad_points.resize(r.n);
      for (int i = 0; i < r.n; ++i) {
        double q[3];
        bem::quad_point(g, r.pt[i], q);
        quad_points[i] = point_type(q[0], q[1], q[2]);
      }
    }
    operator point_type() const { return center; }
    void switch_BC(void) { BC = (BC == POTENTIAL) ? NORMAL_DERIV : POTENTIAL; }
  };

  LaplaceSphericalBEM() : LaplaceSphericalBEM(5, 3) {}
  LaplaceSphericalBEM(int p, unsigned k = 3) : LaplaceSpherical(p), K(k) { global_K() = (int)k; }

  /** K(t, s): the target's boundary condition picks G or dG/dn (reference :273-297) */
  kernel_value_type operator()(const source_type& t, const target_type& s) const {
    bem::Panel g;
    bem::make_panel(s.vertices[0].data(), s.vertices[1].data(), s.vertices[2].data(), g);
    return bem::kernel(t.BC == Panel::POTENTIAL ? 0 : 1, t.center.data(), g, bem::make_rule((int)K), bem::make_rule(17));
  }

  /** what FMM_plan ships through the C ABI for panel sources */
  static void pack_sources(const std::vector<source_type>& src, std::vector<double>& pts,
                           std::vector<double>& verts, std::vector<int32_t>& bc) {
    const size_t n = src.size();
    pts.resize(3 * n); verts.resize(9 * n); bc.resize(n);
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) pts[3 * i + k] = src[i].center[k];
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) verts[9 * i + 3 * v + k] = src[i].vertices[v][k];
      bc[i] = src[i].BC == Panel::POTENTIAL ? 0 : 1;
    }
  }
  int quad_k() const { return (int)K; }

 private:
  static int& global_K() {
    static int k = 3;
    return k;
  }
};
