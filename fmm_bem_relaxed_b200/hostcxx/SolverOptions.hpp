#pragma once
/** @file SolverOptions.hpp
 * Solver configuration and the per-iteration expansion-order rule of the relaxed GMRES
 * (reference examples/BEM/SolverOptions.hpp:9-38): Bouras-Fraysse relaxation
 *   p = min(ceil(-log2(min(residual_tol / min(eps, 1), 1))), max_p).
 */
#include <algorithm>
#include <cmath>

struct SolverOptions {
  double residual;
  int max_iters, restart;
  unsigned max_p, p_min;
  bool variable_p;
  enum relaxation_type { SIMONCINI, BOURAS };
  relaxation_type relax_type;

  SolverOptions(double r, int m_iters, unsigned p)
      : residual(r), max_iters(m_iters), restart(50), max_p(p), p_min(5), variable_p(false), relax_type(BOURAS) {}
  SolverOptions()
      : residual(1e-5), max_iters(500), restart(500), max_p(16), p_min(5), variable_p(true), relax_type(BOURAS) {}

  unsigned predict_p(double eps) const {
    if (!variable_p) return max_p;
    if (relax_type == BOURAS) {
      double alpha = 1. / std::min(eps, 1.);
      double nu = std::min(alpha * residual, 1.);
      return std::min((unsigned)std::ceil(-std::log2(nu)), max_p);
    }
    return std::min((unsigned)std::ceil(-std::log2(eps)), max_p);
  }
};
