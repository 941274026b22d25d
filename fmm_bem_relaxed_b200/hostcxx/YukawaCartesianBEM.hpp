#pragma once
/** @file YukawaCartesianBEM.hpp
 * Host-side kernel class with the reference's public surface (reference kernel/YukawaCartesianBEM.hpp:8-143,
 * 213-230): nested Panel, YukawaCartesianBEM(int p, double kappa, unsigned k), operator()(target, source).
 * Expansion operators and the cached near field run on the GPU behind FMM_plan (csrc/yukawa.cu, csrc/bem.cu);
 * operator() here serves Direct::matvec checks and the diagonal preconditioner and shares its panel integrals with
 * the device code (bem_math.hpp).  Parity status of the far field: DESIGN.md section 2.
 */
#include "LaplaceSphericalBEM.hpp"
#include "YukawaCartesian.hpp"

class YukawaCartesianBEM : public YukawaCartesian {
 public:
  unsigned K;  //!< quadrature points per panel
  typedef LaplaceSphericalBEM::Panel Panel;      // same boundary element: centre, normal, vertices, quad_points, Area, BC
  static constexpr unsigned dimension = YukawaCartesian::dimension;
  typedef YukawaCartesian::point_type point_type;
  typedef Panel source_type;
  typedef Panel target_type;
  typedef YukawaCartesian::charge_type charge_type;
  typedef double kernel_value_type;
  typedef double result_type;
  typedef Panel panel_type;

  static constexpr int fmmb_kind = FMMB_YUKAWA_CARTESIAN_BEM;
  static constexpr int charge_dim = 1;
  static constexpr int result_dim = 1;

  YukawaCartesianBEM() : YukawaCartesianBEM(5, 0.125, 3) {}
  YukawaCartesianBEM(int p, double kappa, unsigned k = 3) : YukawaCartesian(p, kappa), K(k) {
    LaplaceSphericalBEM(p, k);                   // sets the process-wide quadrature order Panel() reads (BEMConfig)
  }

  /** K(t, s): the target's boundary condition picks G or dG/dn (reference :213-230) */
  kernel_value_type operator()(const source_type& t, const target_type& s) const {
    bem::Panel g;
    bem::make_panel(s.vertices[0].data(), s.vertices[1].data(), s.vertices[2].data(), g);
    return bem::kernel_yk(t.BC == Panel::POTENTIAL ? 0 : 1, t.center.data(), g, bem::make_rule((int)K), Kappa);
  }

  static void pack_sources(const std::vector<source_type>& src, std::vector<double>& pts, std::vector<double>& verts,
                           std::vector<int32_t>& bc) {
    LaplaceSphericalBEM::pack_sources(src, pts, verts, bc);
  }
  int quad_k() const { return (int)K; }
};
