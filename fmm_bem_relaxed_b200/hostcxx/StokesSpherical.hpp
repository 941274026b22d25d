#pragma once
/** @file StokesSpherical.hpp
 * Host-side kernel class with the reference's public surface (reference kernel/StokesSpherical.hpp:11-116):
 * typedefs, StokesSpherical(int p), set_p (inherited), and the pair rule the drivers use for accuracy checks --
 * operator()(t, s) -> Mat3 for the Stokeslet, the vector P2P(...) for the stresslet.  Like the reference the
 * flavour is chosen at compile time: define STRESSLET before including this header
 * (reference serialrun_stresslet.cpp:8-9).  The expansion operators run as sm_100a kernels behind FMM_plan
 * (fmm_bem_relaxed_b200/csrc/stokes.cu).
 */
#include "LaplaceSpherical.hpp"
#include "Mat3.hpp"

#include <iostream>

class StokesSpherical : public LaplaceSpherical {
 public:
  typedef LaplaceSpherical::point_type point_type;
  typedef LaplaceSpherical::source_type source_type;
  typedef LaplaceSpherical::target_type target_type;
#ifdef STRESSLET
  //! { g1, g2, g3, n1, n2, n3 }
  typedef Vec<6, LaplaceSpherical::charge_type> charge_type;
  static constexpr int fmmb_kind = FMMB_STOKES_SPHERICAL_STRESSLET;
  static constexpr int charge_dim = 6;
#else
  //! { f1, f2, f3 }
  typedef Vec<3, LaplaceSpherical::charge_type> charge_type;
  static constexpr int fmmb_kind = FMMB_STOKES_SPHERICAL;
  static constexpr int charge_dim = 3;
#endif
  typedef Mat3<real> kernel_value_type;
  typedef Vec<3, real> result_type;
  static constexpr int result_dim = 3;

  StokesSpherical() : StokesSpherical(5) {}
  StokesSpherical(int p) : LaplaceSpherical(p) {
#ifdef STRESSLET
    std::cout << "Stresslet calculation" << std::endl;
#endif
  }

#ifndef STRESSLET
  /** Stokeslet K(t,s): (I r^2 + d d^T) / r^3, zero for r^2 < 1e-8 */
  kernel_value_type operator()(const target_type& t, const source_type& s) const {
    point_type d = s - t;
    real r2 = normSq(d);
    real invR2 = 1.0 / r2;
    if (r2 < 1e-8) invR2 = 0;
    real invR3 = invR2 * std::sqrt(invR2);
    kernel_value_type r(0.);
    for (int i = 0; i < 3; ++i)       // symmetric: the lower-index component is multiplied first, as the reference does
      for (int j = i + 1; j < 3; ++j) r(i, j) = r(j, i) = invR3 * d[i] * d[j];
    for (int i = 0; i < 3; ++i) r(i, i) = invR3 * (r2 + d[i] * d[i]);
    return r;
  }
#else
  /** Stresslet, vector form: u_i += (d.n) d_i (d.g) / r^5 with d = t - s, zero for r^2 < 1e-8 */
  template <typename SourceIter, typename ChargeIter, typename TargetIter, typename ResultIter>
  void P2P(SourceIter s_first, SourceIter s_last, ChargeIter c_first, TargetIter t_first, TargetIter t_last,
           ResultIter r_first) const {
    for (; t_first != t_last; ++t_first, ++r_first) {
      SourceIter s = s_first;
      ChargeIter c = c_first;
      for (; s != s_last; ++s, ++c) {
        point_type d = *t_first - *s;
        real r2 = normSq(d);
        real inv = 1. / r2;
        if (r2 < 1e-8) inv = 0;
        const charge_type& g = *c;
        real dn = d[0] * g[3] + d[1] * g[4] + d[2] * g[5];
        real H = std::sqrt(inv) * inv;
        H *= dn * inv;
        real dg = d[0] * g[0] + d[1] * g[1] + d[2] * g[2];
        for (int i = 0; i < 3; ++i) (*r_first)[i] += H * d[i] * dg;
      }
    }
  }
#endif
};
