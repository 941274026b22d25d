#pragma once
/** @file YukawaCartesian.hpp
 * Host-side kernel class with the reference's public surface (reference kernel/YukawaCartesian.hpp:14-159):
 * typedefs, YukawaCartesian(int p, double kappa = 0.125), operator()(t, s).  The Cartesian Taylor expansion
 * operators run as sm_100a kernels behind FMM_plan (fmm_bem_relaxed_b200/csrc/yukawa.cu); unlike the shipped
 * reference class, whose operators carry a trailing `unsigned p` that its own executor cannot supply
 * (SURVEY.md F7), this one IS usable through FMM_plan, and set_p(p) means "the full order-p expansion".
 *
 * K(t,s) = exp(-kappa |t-s|) / |t-s|                              (potential)
 *          -(kappa |t-s| + 1) exp(-kappa |t-s|) (t-s) / |t-s|^3   (force)
 */
#include <cmath>
#include <cstdint>
#include <vector>
#include <Vec.hpp>

#include "../../include/fmmb.h"

class YukawaCartesian {
 protected:
  int P;
  double Kappa;

 public:
  typedef double real;
  static constexpr unsigned dimension = 3;
  typedef Vec<dimension, real> point_type;
  typedef point_type source_type;
  typedef point_type target_type;
  typedef real charge_type;
  typedef Vec<4, real> kernel_value_type;
  typedef Vec<4, real> result_type;

  static constexpr int fmmb_kind = FMMB_YUKAWA_CARTESIAN;
  static constexpr int charge_dim = 1;
  static constexpr int result_dim = 4;

  YukawaCartesian() : YukawaCartesian(4, 0.125) {}
  YukawaCartesian(int p, double kappa = 0.125) : P(p), Kappa(kappa) {}

  /** extension: change the expansion order (1..10); takes effect at the next FMM_plan::execute */
  void set_p(int p) { P = p; }
  int order() const { return P; }
  double kappa() const { return Kappa; }
  int quad_k() const { return 0; }
  int quad_kfine() const { return 0; }
  int kernel_flags() const { return 0; }

  /** Kernel evaluation K(t,s), same operation order as the reference (:148-159) */
  kernel_value_type operator()(const point_type& t, const point_type& s) const {
    point_type dist = t - s;
    real r2 = normSq(dist);
    real r = std::sqrt(r2);
    real invR2 = 1.0 / r2;
    real invR = 1.0 / r;
    if (r < 1e-8) { invR = 0; invR2 = 0; }
    real pot = std::exp(-Kappa * r) * invR;
    dist *= pot * (Kappa * r + 1) * invR2;
    return kernel_value_type(pot, -dist[0], -dist[1], -dist[2]);
  }

  static void pack_sources(const std::vector<source_type>& src, std::vector<double>& pts, std::vector<double>& verts,
                           std::vector<int32_t>& bc) {
    pts.resize(3 * src.size());
    verts.clear(); bc.clear();
    for (size_t i = 0; i < src.size(); ++i) { pts[3 * i] = src[i][0]; pts[3 * i + 1] = src[i][1]; pts[3 * i + 2] = src[i][2]; }
  }
};
