#pragma once
/** @file stokes_bem_math.hpp
 * Near-field panel integrals of StokesSphericalBEM, written once for the host (kernel class operator(), used by
 * Direct::matvec checks) and for the device (near-field assembly in csrc/stokes_bem.cu).  A matrix entry is a
 * 3 x 3 block (the reference's Mat3<real>), row-major in m[9].
 *
 * Follows reference kernel/StokesSphericalBEM.hpp:160-255 (eval_traction_integral), :257-375
 * (eval_velocity_integral), :377-390 (operator()) and the self term of the single layer that the reference takes
 * from examples/BEM/FataAnalytical.hpp:414-713 + :272-341.
 */
#include "bem_math.hpp"

namespace bem {

/** Single-layer self term: the integral over the panel seen from its own centre, as the reference evaluates it.
 *
 * With the panel frame e1 = (v1 - v0)/|v1 - v0|, e2 in the panel plane, e3 = e1 x e2 (FataAnalytical.hpp:156-187),
 * edge i running from vertex i to vertex i+1 with unit tangent t_i and in-plane normal n_i = e3 x t_i, and for a
 * vertex y: p = (y - x).t_i, q_i = (y_i - x).n_i, rho = |y - x|,
 *     omega = sum_i q_i log((p_start + rho_start) / (p_end + rho_end))            (:547-551; = int 1/r dS)
 *     K     = omega I + sum_i (rho_i - rho_{i+1}) [ t_x t_y (e1 e1' - e2 e2') + t_y^2 (e1 e2' + e2 e1') ]
 * (t_x, t_y = components of t_i in the frame).  The second sum is what :300-317 leaves of the r r'/r^3 moments when
 * the edge logarithms chi_i are zero, and the self branch (:547-551) never sets them -- so this is the reference's
 * value, not the exact integral of (I/r + r r'/r^3), which has extra q_i chi_i terms.  Kept for parity.
 * The geometry is taken from dot products instead of the reference's acos / sin / cos chain; agreement 1e-15. */
BEM_HD void stokes_self_term(const Panel& s, double* m) {
  const double* y[3] = {s.v[0], s.v[1], s.v[2]};
  const double* x = s.c;
  double a[3], b[3], e1[3], e2[3], e3[3];
  for (int k = 0; k < 3; ++k) { a[k] = y[1][k] - y[0][k]; b[k] = y[2][k] - y[0][k]; }
  const double na2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], na = sqrt(na2);
  const double al = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / na2;
  for (int k = 0; k < 3; ++k) e2[k] = b[k] - al * a[k];
  const double nb = norm3(e2);
  for (int k = 0; k < 3; ++k) { e1[k] = a[k] / na; e2[k] = e2[k] / nb; }
  cross3(e1, e2, e3);
  double rho[3], d[3][3];
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) d[i][k] = y[i][k] - x[k];
    rho[i] = norm3(d[i]);
  }
  double omega = 0, sxy = 0, syy = 0;
  for (int i = 0; i < 3; ++i) {
    const int j = (i + 1) % 3;
    double t[3], n[3];
    for (int k = 0; k < 3; ++k) t[k] = y[j][k] - y[i][k];
    const double tl = norm3(t);
    for (int k = 0; k < 3; ++k) t[k] /= tl;
    cross3(e3, t, n);
    const double ps = d[i][0] * t[0] + d[i][1] * t[1] + d[i][2] * t[2];
    const double pe = d[j][0] * t[0] + d[j][1] * t[1] + d[j][2] * t[2];
    const double q = d[i][0] * n[0] + d[i][1] * n[1] + d[i][2] * n[2];
    omega += q * log((ps + rho[i]) / (pe + rho[j]));
    const double tx = t[0] * e1[0] + t[1] * e1[1] + t[2] * e1[2];
    const double ty = t[0] * e2[0] + t[1] * e2[1] + t[2] * e2[2];
    sxy += (rho[i] - rho[j]) * tx * ty;
    syy += (rho[i] - rho[j]) * ty * ty;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      m[3 * r + c] = (r == c ? omega : 0.0) + sxy * (e1[r] * e1[c] - e2[r] * e2[c]) + syy * (e1[r] * e2[c] + e2[r] * e1[c]);
}

/** Single layer (eval_velocity_integral, :257-375): sum_i w_i A (r^2 I + d d') / r^3, d = t - q_i, scaled by
 * 1 / (2 mu); a quadrature point closer than 1e-4 contributes nothing (:353).
 *
 * as_written = false (default everywhere): the K-point rule for EVERY pair, the panel itself included.  This is what
 * the unmodified reference computes when built with g++: lines :262-263 (and :162-163 of the double layer) read
 *     auto dist = static_cast<point_type>(target) - source.center;   auto d = norm(dist);
 * and the expression template `dist` refers to the temporary point, destroyed at the end of its statement, so `d`
 * comes from dead stack (undefined behaviour) and, as compiled, never selects the self term or the fine rule.
 * as_written = true: the branches as the source text means them -- the panel itself through the self term above,
 * panels with sqrt(2A)/dist >= 0.5 through the fine rule; pinned against the reference with that one declaration
 * changed to point_type.  What the dead stack holds depends on the translation unit and its
 * optimisation level: the reference's drivers built by the oracle recipe behave as described (bit-identical matvecs on
 * every fixture), other programs around the same header need not.  On the unit sphere (2 048 panels, tol 1e-5) the
 * as-compiled entries converge in 10 iterations to a drag 0.8 % low, the as-written ones in 22 iterations to a drag
 * 0.4 % high. */
BEM_HD void stokes_velocity_entry(const Panel& s, const double* t, const Rule& rule, const Rule& fine, double mu,
                                  bool as_written, double* m) {
  const double dc[3] = {t[0] - s.c[0], t[1] - s.c[1], t[2] - s.c[2]};
  const double dist = norm3(dc);
  const double scale = 1. / 2 / mu;
  if (as_written && dist < 1e-8) {
    stokes_self_term(s, m);
    for (int e = 0; e < 9; ++e) m[e] *= scale;
    return;
  }
  const Rule& g = (as_written && sqrt(2 * s.area) / dist >= 0.5) ? fine : rule;
  for (int e = 0; e < 9; ++e) m[e] = 0.0;
  for (int i = 0; i < g.n; ++i) {
    double q[3];
    quad_point(s, g.pt[i], q);
    const double dx = t[0] - q[0], dy = t[1] - q[1], dz = t[2] - q[2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    double inv2 = 1. / r2;
    if (r2 < 1e-8) inv2 = 0;
    const double inv3 = inv2 * sqrt(inv2);
    const double f = g.w[i] * s.area * inv3;
    m[0] += f * (r2 + dx * dx); m[1] += f * (dx * dy); m[2] += f * (dx * dz);
    m[3] += f * (dx * dy); m[4] += f * (r2 + dy * dy); m[5] += f * (dy * dz);
    m[6] += f * (dx * dz); m[7] += f * (dy * dz); m[8] += f * (r2 + dz * dz);
  }
  for (int e = 0; e < 9; ++e) m[e] *= scale;
}

/** Double layer (eval_traction_integral, :160-255): -3 sum_i w_i A (d.n) d d' / r^5, d = t - q_i.  As written the
 * panel itself contributes 2 pi I and close panels use the fine rule; as compiled (see above) every pair takes the
 * K-point rule, which leaves rounding noise on the diagonal (d.n = 0 in the panel plane). */
BEM_HD void stokes_traction_entry(const Panel& s, const double* t, const Rule& rule, const Rule& fine, bool as_written,
                                  double* m) {
  const double dc[3] = {t[0] - s.c[0], t[1] - s.c[1], t[2] - s.c[2]};
  const double dist = norm3(dc);
  for (int e = 0; e < 9; ++e) m[e] = 0.0;
  if (as_written && fabs(dist) < 1e-8) {
    m[0] = m[4] = m[8] = 2 * M_PI;
    return;
  }
  const Rule& g = (as_written && sqrt(2 * s.area) / dist >= 0.5) ? fine : rule;
  for (int i = 0; i < g.n; ++i) {
    double q[3];
    quad_point(s, g.pt[i], q);
    const double dx = t[0] - q[0], dy = t[1] - q[1], dz = t[2] - q[2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    double inv2 = 1. / r2;
    if (r2 < 1e-8) inv2 = 0;
    const double inv5 = inv2 * inv2 * sqrt(inv2);
    const double dn = dx * s.nrm[0] + dy * s.nrm[1] + dz * s.nrm[2];
    const double f = g.w[i] * s.area * dn * inv5;
    m[0] += f * (dx * dx); m[1] += f * (dx * dy); m[2] += f * (dx * dz);
    m[3] += f * (dx * dy); m[4] += f * (dy * dy); m[5] += f * (dy * dz);
    m[6] += f * (dx * dz); m[7] += f * (dy * dz); m[8] += f * (dz * dz);
  }
  for (int e = 0; e < 9; ++e) m[e] *= -3;
}

/** K(t, s) (operator(), :377-390): the TARGET's boundary condition picks the layer.
 * bc 0 = VELOCITY (single layer), 1 = TRACTION (double layer). */
BEM_HD void stokes_kernel(int target_bc, const double* target_centre, const Panel& s, const Rule& rule, const Rule& fine,
                          double mu, bool as_written, double* m) {
  if (target_bc == 0) stokes_velocity_entry(s, target_centre, rule, fine, mu, as_written, m);
  else stokes_traction_entry(s, target_centre, rule, fine, as_written, m);
}

}  // namespace bem
