#pragma once
/** @file GMRES.hpp
 * Restarted GMRES with per-iteration relaxation of the FMM expansion order -- the caller of the hot
 * path.  Same algorithm, call signature and progress lines as the reference
 * (examples/BEM/GMRES.hpp:120-252): modified Gram-Schmidt, Givens rotations, convergence on the
 * rotated residual estimate |s[i+1]| / ||b||; before every matvec
 *     p = max(1, opts.predict_p(|resid|));  MV.kernel().set_p(p);
 * so a plan whose kernel honours set_p (FMM_plan) runs cheaper matvecs as the residual falls.
 * Vectors are std::vector<double> (scalar charge / result types).
 */
#include <cmath>
#include <cstdio>
#include <vector>

#include "Preconditioner.hpp"
#include "SolverOptions.hpp"
#include "../../include/fmmb.h"

namespace gmres_detail {
inline double dot(const std::vector<double>& a, const std::vector<double>& b) {
  double r = 0;
  for (size_t i = 0; i < a.size(); ++i) r += a[i] * b[i];
  return r;
}
inline double nrm2(const std::vector<double>& a) { return std::sqrt(dot(a, a)); }
inline void axpy(const std::vector<double>& x, std::vector<double>& y, double a) {
  for (size_t i = 0; i < x.size(); ++i) y[i] = a * x[i] + y[i];
}
inline void apply_rotation(double& dx, double& dy, double cs, double sn) {
  double t = cs * dx + sn * dy;
  dy = -sn * dx + cs * dy;
  dx = t;
}
inline void make_rotation(double dx, double dy, double& cs, double& sn) {
  if (dy == 0.0) { cs = 1.0; sn = 0.0; }
  else if (std::fabs(dy) > std::fabs(dx)) { double t = dx / dy; sn = 1.0 / std::sqrt(1.0 + t * t); cs = t * sn; }
  else { double t = dy / dx; cs = 1.0 / std::sqrt(1.0 + t * t); sn = t * cs; }
}
}  // namespace gmres_detail

/** Iteration record for callers that want more than the printed lines */
struct GMRESReport {
  int iterations = 0;
  double final_residual = 0;
  std::vector<int> p_schedule;        // expansion order requested for every inner iteration
  std::vector<double> residuals;      // |resid| after every inner iteration
};

template <typename Matvec, typename Preconditioner>
GMRESReport GMRES(Matvec& MV, std::vector<typename Matvec::charge_type>& x,
                  std::vector<typename Matvec::result_type>& b, const SolverOptions& opts,
                  const Preconditioner& M, bool output = true) {
  using namespace gmres_detail;
  const int R = opts.restart, n = (int)x.size();
  GMRESReport rep;
  std::vector<std::vector<double>> V(1, std::vector<double>(n));
  std::vector<std::vector<double>> H;   // H[j] = column j (R+1 entries used up to j+1)
  std::vector<double> s(R + 1), cs(R), sn(R), w, z(n);
  const double normb = nrm2(b);
  auto& K = MV.kernel();
  double resid = 0;
  int i, iter = 0;
  do {
    w = MV.execute(x);               // uses whatever order the kernel currently has (reference :174-175)
    axpy(b, w, -1.);
    const double beta = nrm2(w);
    V.assign(1, w);
    for (auto& v : V[0]) v *= -1. / beta;
    H.clear();
    s.assign(R + 1, 0.0);
    s[0] = beta;
    i = -1;
    resid = s[0] / normb;
    do {
      ++i; ++iter;
      const int p = (int)std::max(1u, opts.predict_p(std::fabs(resid)));
      K.set_p(p);
      M(V[i], z);
      w = MV.execute(z);
      H.push_back(std::vector<double>(i + 2));
      for (int k = 0; k <= i; ++k) {
        H[i][k] = dot(w, V[k]);
        axpy(V[k], w, -H[i][k]);
      }
      H[i][i + 1] = nrm2(w);
      V.push_back(w);
      for (auto& v : V[i + 1]) v *= 1. / H[i][i + 1];
      for (int k = 0; k < i; ++k) apply_rotation(H[i][k], H[i][k + 1], cs[k], sn[k]);
      make_rotation(H[i][i], H[i][i + 1], cs[i], sn[i]);
      apply_rotation(H[i][i], H[i][i + 1], cs[i], sn[i]);
      apply_rotation(s[i], s[i + 1], cs[i], sn[i]);
      resid = s[i + 1] / normb;
      rep.p_schedule.push_back(p);
      rep.residuals.push_back(std::fabs(resid));
      if (std::fabs(resid) < opts.residual) break;
      if (output) printf("it: %03d, res: %.3e, fmm_req_p: %01d\n", iter, std::fabs(resid), p);
    } while (i + 1 < R && i + 1 <= opts.max_iters && std::fabs(resid) > opts.residual);
    for (int j = i; j >= 0; --j) {
      s[j] /= H[j][j];
      for (int k = j - 1; k >= 0; --k) s[k] -= H[j][k] * s[j];
    }
    for (int j = 0; j <= i; ++j) {
      M(V[j], z);
      axpy(z, x, s[j]);
    }
    if (output && iter % 10 == 0) printf("it: %04d, residual: %.3e\n", iter, std::fabs(resid));
  } while (std::fabs(resid) > opts.residual && iter < opts.max_iters);
  if (output) printf("Final residual: %.4e, after %d iterations\n", std::fabs(resid), iter);
  rep.iterations = iter;
  rep.final_residual = std::fabs(resid);
  return rep;
}

template <typename Matvec>
GMRESReport GMRES(Matvec& MV, std::vector<typename Matvec::charge_type>& x,
                  std::vector<typename Matvec::result_type>& b, const SolverOptions& opts, bool output = true) {
  return GMRES(MV, x, b, opts, Preconditioners::Identity(), output);
}

/** The same solver, device resident (fmmb_gmres in include/fmmb.h): Krylov basis and BLAS-1 work stay on the GPU and
 * the matvec is fed from device vectors; iteration counts and orders follow the same rule.  `diag`: optional
 * reciprocals for the diagonal preconditioner, empty = identity.  Available when MV is an FMM_plan. */
template <typename Matvec>
GMRESReport GMRES_device(Matvec& MV, std::vector<double>& x, std::vector<double>& b, const SolverOptions& opts,
                         const std::vector<double>& diag = std::vector<double>(), bool output = true) {
  GMRESReport rep;
  fmmb_solver_options so = {opts.residual, opts.max_iters, opts.restart, opts.max_p, opts.variable_p ? 1 : 0,
                            opts.relax_type == SolverOptions::BOURAS ? 0 : 1, output ? 1 : 0};
  fmmb_gmres_info info = {};
  const int cap = 4096;
  std::vector<int32_t> ps(cap);
  std::vector<double> rs(cap);
  if (fmmb_plan_set_p(MV.handle(), MV.kernel().order()) != FMMB_OK ||
      fmmb_gmres(MV.handle(), b.data(), x.data(), diag.empty() ? nullptr : diag.data(), &so, &info, ps.data(), rs.data(),
                 cap) != FMMB_OK) {
    fprintf(stderr, "[E]: GMRES_device: %s\n", fmmb_last_error());
    return rep;
  }
  MV.kernel().set_p(info.final_p);       // the reference leaves the kernel at the last relaxed order
  rep.iterations = info.iterations;
  rep.final_residual = info.final_residual;
  const int k = info.n_records < cap ? info.n_records : cap;
  rep.p_schedule.assign(ps.begin(), ps.begin() + k);
  rep.residuals.assign(rs.begin(), rs.begin() + k);
  return rep;
}

/** Vec<3> unknowns (StokesSphericalBEM): GMRES(plan, x, b, opts) of reference examples/BEM/GMRES_Stokes.hpp:170-300 --
 * the same iteration on the flat array of 3 n doubles, with that header's order rule
 * p = max(p_min, predict_p(|resid|) - 1) (:229) -- device resident. */
template <typename Matvec>
GMRESReport GMRES_device(Matvec& MV, std::vector<Vec<3, double>>& x, std::vector<Vec<3, double>>& b,
                         const SolverOptions& opts, bool output = true) {
  GMRESReport rep;
  static_assert(sizeof(Vec<3, double>) == 3 * sizeof(double), "Vec<3,double> must be packed");
  fmmb_solver_options so = {opts.residual, opts.max_iters, opts.restart, opts.max_p, opts.variable_p ? 1 : 0,
                            opts.relax_type == SolverOptions::BOURAS ? 0 : 1, output ? 1 : 0, opts.p_min, 1u};
  fmmb_gmres_info info = {};
  const int cap = 4096;
  std::vector<int32_t> ps(cap);
  std::vector<double> rs(cap);
  if (x.size() != b.size() || fmmb_plan_set_p(MV.handle(), MV.kernel().order()) != FMMB_OK ||
      fmmb_gmres(MV.handle(), reinterpret_cast<const double*>(b.data()), reinterpret_cast<double*>(x.data()), nullptr, &so,
                 &info, ps.data(), rs.data(), cap) != FMMB_OK) {
    fprintf(stderr, "[E]: GMRES_device: %s\n", x.size() != b.size() ? "x.size() != b.size()" : fmmb_last_error());
    return rep;
  }
  MV.kernel().set_p(info.final_p);
  rep.iterations = info.iterations;
  rep.final_residual = info.final_residual;
  const int k = info.n_records < cap ? info.n_records : cap;
  rep.p_schedule.assign(ps.begin(), ps.begin() + k);
  rep.residuals.assign(rs.begin(), rs.begin() + k);
  return rep;
}

/** FGMRES(plan, x, b, opts[, M]) of reference examples/BEM/GMRES_Stokes.hpp:297-431 on Vec<3> unknowns, device resident
 * (fmmb_fgmres): flexible GMRES with that function's order rule p = max(5, predict_p(|resid|)) (:375).  `pc`: nullptr for
 * the identity, or the near-field-only plan a Preconditioners::LocalInnerSolver / BlockDiagonal would own
 * (FMMOptions::local_evaluation / block_diagonal); the inner solves then run with those classes' options
 * (LocalPC_Stokes.hpp:53-57: residual 1e-1, max_iters 1, fixed order, restart 50), on the device as well. */
template <typename Matvec>
GMRESReport FGMRES_device(Matvec& MV, std::vector<Vec<3, double>>& x, std::vector<Vec<3, double>>& b,
                          const SolverOptions& opts, Matvec* pc = nullptr, bool output = true) {
  GMRESReport rep;
  static_assert(sizeof(Vec<3, double>) == 3 * sizeof(double), "Vec<3,double> must be packed");
  fmmb_solver_options so = {opts.residual, opts.max_iters, opts.restart, opts.max_p, opts.variable_p ? 1 : 0,
                            opts.relax_type == SolverOptions::BOURAS ? 0 : 1, output ? 1 : 0, 5u, 0u};
  fmmb_solver_options in = {1e-1, 1, 50, opts.max_p, 0, 0, 0, 0u, 0u};
  fmmb_gmres_info info = {};
  const int cap = 4096;
  std::vector<int32_t> ps(cap);
  std::vector<double> rs(cap);
  if (x.size() != b.size() || fmmb_plan_set_p(MV.handle(), MV.kernel().order()) != FMMB_OK ||
      fmmb_fgmres(MV.handle(), pc ? pc->handle() : nullptr, pc ? &in : nullptr, reinterpret_cast<const double*>(b.data()),
                  reinterpret_cast<double*>(x.data()), &so, &info, ps.data(), rs.data(), cap) != FMMB_OK) {
    fprintf(stderr, "[E]: FGMRES_device: %s\n", x.size() != b.size() ? "x.size() != b.size()" : fmmb_last_error());
    return rep;
  }
  MV.kernel().set_p(info.final_p);
  rep.iterations = info.iterations;
  rep.final_residual = info.final_residual;
  const int k = info.n_records < cap ? info.n_records : cap;
  rep.p_schedule.assign(ps.begin(), ps.begin() + k);
  rep.residuals.assign(rs.begin(), rs.begin() + k);
  return rep;
}
