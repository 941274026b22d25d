#pragma once
/** @file StokesSphericalBEM.hpp
 * Host-side kernel class with the reference's public surface (reference kernel/StokesSphericalBEM.hpp:9-158,
 * 377-390): nested Panel (center, normal, vertices, quad_points, Area, BC = VELOCITY / TRACTION, switch_BC,
 * conversion to point_type), StokesSphericalBEM(int p, unsigned k, double mu), set_Kfine, set_p (inherited),
 * operator()(target, source) -> Mat3.  The expansion operators and the cached 3 x 3 block near field run on the
 * GPU behind FMM_plan (fmm_bem_relaxed_b200/csrc/stokes_bem.cu); operator() here serves Direct::matvec checks
 * (examples/StokesBEM.cpp:377-380) and shares its panel integrals with the device code (stokes_bem_math.hpp).
 * The reference's examples/StokesBEM.cpp and its solver stack (GMRES_Stokes.hpp, LocalPC_Stokes.hpp,
 * BlockDiagonalPC_Stokes.hpp) compile unchanged against this header (hostcxx/Makefile: bin/ref_StokesBEM).
 *
 * Like the reference this class is meant for the default (Stokeslet) build of StokesSpherical: do not define
 * STRESSLET.  stokeslet_str / stresslet_str are kept as members (the reference accumulates diagnostic sums into
 * them inside P2M, :420-423,447-454, and its driver prints them); here they stay zero.
 */
#include <cmath>
#include <complex>
#include <vector>

#include "stokes_bem_math.hpp"
#include "StokesSpherical.hpp"
#include "Mat3.hpp"

class StokesSphericalBEM : public StokesSpherical {
 public:
  typedef std::complex<double> complex;
  mutable complex stokeslet_str[4], stresslet_str[4];
  unsigned K;       //!< quadrature rule per panel
  unsigned K_fine;  //!< rule for panels closer than 2 sqrt(2 Area)
  double Mu;        //!< viscosity
  /** extension: false (default) = near-field entries as the unmodified reference computes them when compiled (K-point
   * rule for every pair); true = as its source text means them (self terms, fine rule).  See stokes_bem_math.hpp. */
  bool near_field_as_written = false;
  struct Panel;
  static constexpr unsigned dimension = StokesSpherical::dimension;
  typedef StokesSpherical::point_type point_type;
  typedef Panel source_type;
  typedef Panel target_type;
  typedef StokesSpherical::charge_type charge_type;
  typedef Mat3<real> kernel_value_type;
  typedef StokesSpherical::result_type result_type;
  typedef Panel panel_type;

  static constexpr int fmmb_kind = FMMB_STOKES_SPHERICAL_BEM;
  static constexpr int charge_dim = 3;
  static constexpr int result_dim = 3;

  //! Boundary element
  struct Panel {
    typedef enum { VELOCITY, TRACTION } BoundaryType;
    point_type center;
    point_type normal;
    std::vector<point_type> vertices;
    std::vector<point_type> quad_points;
    double Area;
    BoundaryType BC;

    Panel() : center(0), normal(0), Area(0), BC(VELOCITY) {}
    Panel(point_type p0, point_type p1, point_type p2) : BC(VELOCITY) {
      vertices.resize(3);
      vertices[0] = p0; vertices[1] = p1; vertices[2] = p2;
      bem::Panel g;
      bem::make_panel(p0.data(), p1.data(), p2.data(), g);
      center = point_type(g.c[0], g.c[1], g.c[2]);
      normal = point_type(g.nrm[0], g.nrm[1], g.nrm[2]);
      Area = g.area;
      // quadrature points of the process-wide rule (the reference keeps K in a BEMConfig singleton, :84-96)
      const bem::Rule r = bem::make_rule(global_K());
      quad_points.resize(r.n);
      for (int i = 0; i < r.n; ++i) {
        double q[3];
        bem::quad_point(g, r.pt[i], q);
        quad_points[i] = point_type(q[0], q[1], q[2]);
      }
    }
    operator point_type() const { return center; }
    void switch_BC(void) { BC = (BC == VELOCITY) ? TRACTION : VELOCITY; }
  };

  StokesSphericalBEM() : StokesSphericalBEM(5, 3, 1e-3) {}
  StokesSphericalBEM(int p, unsigned k) : StokesSphericalBEM(p, k, 1e-3) {}
  StokesSphericalBEM(int p, unsigned k, double mu) : StokesSpherical(p), K(k), K_fine(25), Mu(mu) {
    global_K() = (int)k;
    for (unsigned i = 0; i < 4; ++i) stresslet_str[i] = stokeslet_str[i] = 0.;
  }
  void set_Kfine(unsigned k) { K_fine = k; }

  kernel_value_type eval_velocity_integral(const source_type& source, const target_type& target) const {
    return entry(0, source, target);
  }
  kernel_value_type eval_traction_integral(const source_type& source, const target_type& target) const {
    return entry(1, source, target);
  }
  /** K(t, s): the target's boundary condition picks the single or the double layer (reference :377-390) */
  kernel_value_type operator()(const target_type& t, const source_type& s) const {
    return entry(t.BC == Panel::VELOCITY ? 0 : 1, s, t);
  }

  /** what FMM_plan ships through the C ABI for panel sources */
  static void pack_sources(const std::vector<source_type>& src, std::vector<double>& pts, std::vector<double>& verts,
                           std::vector<int32_t>& bc) {
    const size_t n = src.size();
    pts.resize(3 * n); verts.resize(9 * n); bc.resize(n);
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) pts[3 * i + k] = src[i].center[k];
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) verts[9 * i + 3 * v + k] = src[i].vertices[v][k];
      bc[i] = src[i].BC == Panel::VELOCITY ? 0 : 1;
    }
  }
  int quad_k() const { return (int)K; }
  int quad_kfine() const { return (int)K_fine; }
  double kappa() const { return Mu; }   // fmmb_kernel_desc.kappa carries the viscosity for this kernel class
  int kernel_flags() const { return near_field_as_written ? FMMB_FLAG_STOKES_BEM_AS_WRITTEN : 0; }

 private:
  kernel_value_type entry(int layer, const source_type& s, const target_type& t) const {
    bem::Panel g;
    bem::make_panel(s.vertices[0].data(), s.vertices[1].data(), s.vertices[2].data(), g);
    double m[9];
    bem::stokes_kernel(layer, t.center.data(), g, bem::make_rule((int)K), bem::make_rule((int)K_fine), Mu,
                       near_field_as_written, m);
    return kernel_value_type(m, m + 9);
  }
  static int& global_K() {
    static int k = 3;
    return k;
  }
};
