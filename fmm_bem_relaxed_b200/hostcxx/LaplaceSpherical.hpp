#pragma once
/** @file LaplaceSpherical.hpp
 * Host-side kernel class with the reference's public surface
 * (reference kernel/LaplaceSpherical.hpp:54-81,119-128,153-176): typedefs, LaplaceSpherical(int p),
 * set_p, operator()(t, s), transpose.  The expansion operators (P2M ... L2P) are NOT here: they run as
 * sm_100a kernels behind FMM_plan (fmm_bem_relaxed_b200/csrc/laplace.cu, m2l_classes.cu).
 *
 * K(t,s) = 1/|s-t| (potential), (s-t)/|s-t|^3 (force)
 */
#include <cmath>
#include <cstdint>
#include <vector>
#include <Vec.hpp>

#include "../../include/fmmb.h"

class LaplaceSpherical {
  int P;

 public:
  typedef double real;
  static constexpr unsigned dimension = 3;
  typedef Vec<dimension, real> point_type;
  typedef point_type source_type;
  typedef point_type target_type;
  typedef real charge_type;
  typedef Vec<4, real> kernel_value_type;
  typedef Vec<4, real> result_type;

  //! what fmmb_plan_create needs to know about this kernel class
  static constexpr int fmmb_kind = FMMB_LAPLACE_SPHERICAL;
  static constexpr int charge_dim = 1;
  static constexpr int result_dim = 4;

  LaplaceSpherical() : LaplaceSpherical(5) {}
  LaplaceSpherical(int p) : P(p) {}

  /** Change the expansion order; takes effect at the next FMM_plan::execute */
  void set_p(int p) { P = p; }
  int order() const { return P; }
  double kappa() const { return 0.0; }

  /** Kernel evaluation K(t,s), same operation order as the reference (:153-162) */
  kernel_value_type operator()(const target_type& t, const source_type& s) const {
    point_type dist = s - t;
    real R2 = normSq(dist);
    real invR2 = 1.0 / R2;
    if (R2 < 1e-8) invR2 = 0;
    real invR = std::sqrt(invR2);
    dist *= invR2 * invR;
    return kernel_value_type(invR, dist[0], dist[1], dist[2]);
  }
  kernel_value_type transpose(const kernel_value_type& kst) const {
    return kernel_value_type(kst[0], -kst[1], -kst[2], -kst[3]);
  }

  /** what FMM_plan ships through the C ABI for point sources */
  static void pack_sources(const std::vector<source_type>& src, std::vector<double>& pts, std::vector<double>& verts,
                           std::vector<int32_t>& bc) {
    pts.resize(3 * src.size());
    verts.clear(); bc.clear();
    for (size_t i = 0; i < src.size(); ++i) { pts[3 * i] = src[i][0]; pts[3 * i + 1] = src[i][1]; pts[3 * i + 2] = src[i][2]; }
  }
  int quad_k() const { return 0; }
  int quad_kfine() const { return 0; }
  int kernel_flags() const { return 0; }
};
