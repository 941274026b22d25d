#pragma once
/** @file bem_math.hpp
 * Panel geometry and near-field panel integrals of LaplaceSphericalBEM, written once for the host
 * (kernel class operator(), used by Direct::matvec checks and the diagonal preconditioner) and for the
 * device (near-field assembly in csrc/bem.cu).
 *
 * Follows reference kernel/LaplaceSphericalBEM.hpp:61-97 (Panel), :159-264 (eval_G / eval_dGdn),
 * :273-297 (operator()), examples/BEM/SemiAnalytical.hpp:13-203 (Laplace branch) and the Gauss rules
 * 1, 3, 4 and "17" (16 points) of examples/BEM/GaussQuadrature.hpp:27-31,86-116.
 */
#include <cmath>

#if defined(__CUDACC__)
#define BEM_HD __host__ __device__ inline
#else
#define BEM_HD inline
#endif

namespace bem {

struct Panel {
  double v[3][3];   // vertices
  double c[3];      // centre
  double nrm[3];    // normal (cross(p2-p0, p1-p0) normalised, the reference's orientation)
  double area;
};

struct Rule {
  int n;
  double pt[16][3];
  double w[16];
};

/** Triangle Gauss rule with k points; k = 7 aliases the 4-point rule like the reference does. */
inline Rule make_rule(int k) {
  Rule r = {};
  if (k == 1) {
    r.n = 1; r.pt[0][0] = r.pt[0][1] = r.pt[0][2] = 1. / 3; r.w[0] = 1.;
  } else if (k == 3) {
    const double p[3][3] = {{0.5, 0.5, 0.}, {0., 0.5, 0.5}, {0.5, 0., 0.5}};
    r.n = 3;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) r.pt[i][j] = p[i][j]; r.w[i] = 1. / 3; }
  } else if (k == 17) {
    const double a = 1 / 3., b1 = 0.081414823414554, b2 = 0.459292588292723, c1 = 0.658861384496480,
                 c2 = 0.170569307751760, d1 = 0.898905543365938, d2 = 0.050547228317031,
                 e1 = 0.008394777409958, e2 = 0.263112829634638, e3 = 0.728492392955404;
    const double wa = 0.144315607677787, wb = 0.095091634267285, wc = 0.103217370534718,
                 wd = 0.032458497623198, we = 0.027230314174435;
    const double p[16][3] = {{a, a, a}, {b1, b2, b2}, {b2, b1, b2}, {b2, b2, b1}, {c1, c2, c2}, {c2, c1, c2},
                             {c2, c2, c1}, {d1, d2, d2}, {d2, d1, d2}, {d2, d2, d1}, {e1, e2, e3}, {e1, e3, e2},
                             {e2, e1, e3}, {e2, e3, e1}, {e3, e1, e2}, {e3, e2, e1}};
    const double w[16] = {wa, wb, wb, wb, wc, wc, wc, wd, wd, wd, we, we, we, we, we, we};
    r.n = 16;
    for (int i = 0; i < 16; ++i) { for (int j = 0; j < 3; ++j) r.pt[i][j] = p[i][j]; r.w[i] = w[i]; }
  } else {
    const double p[4][3] = {{1. / 3, 1. / 3, 1. / 3}, {.6, .2, .2}, {.2, .6, .2}, {.2, .2, .6}};
    const double w[4] = {-27. / 48, 25. / 48, 25. / 48, 25. / 48};
    r.n = 4;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 3; ++j) r.pt[i][j] = p[i][j]; r.w[i] = w[i]; }
  }
  return r;
}
inline bool rule_supported(int k) { return k == 1 || k == 3 || k == 4 || k == 7; }

BEM_HD double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
BEM_HD void cross3(const double* u, const double* v, double* o) {
  o[0] = u[1] * v[2] - u[2] * v[1];
  o[1] = u[2] * v[0] - u[0] * v[2];
  o[2] = u[0] * v[1] - u[1] * v[0];
}
BEM_HD void matvec3(const double* m, const double* x, double* o) {
  o[0] = m[0] * x[0] + m[1] * x[1] + m[2] * x[2];
  o[1] = m[3] * x[0] + m[4] * x[1] + m[5] * x[2];
  o[2] = m[6] * x[0] + m[7] * x[1] + m[8] * x[2];
}

/** Panel(p0, p1, p2): centre, area, normal in the reference's operation order */
BEM_HD void make_panel(const double* p0, const double* p1, const double* p2, Panel& p) {
  double L0[3], L1[3];
  for (int k = 0; k < 3; ++k) {
    p.v[0][k] = p0[k]; p.v[1][k] = p1[k]; p.v[2][k] = p2[k];
    p.c[k] = ((p0[k] + p1[k]) + p2[k]) / 3;
    L0[k] = p2[k] - p0[k];
    L1[k] = p1[k] - p0[k];
  }
  double c[3] = {L0[1] * L1[2] - L0[2] * L1[1], -(L0[0] * L1[2] - L0[2] * L1[0]), L0[0] * L1[1] - L0[1] * L1[0]};
  p.area = 0.5 * norm3(c);
  for (int k = 0; k < 3; ++k) p.nrm[k] = c[k] / 2 / p.area;
}
/** i-th quadrature point: barycentric combination of the vertices */
BEM_HD void quad_point(const Panel& p, const double* bary, double* q) {
  for (int k = 0; k < 3; ++k) q[k] = p.v[0][k] * bary[0] + p.v[1][k] * bary[1] + p.v[2][k] * bary[2];
}

/** 5-point Gauss line integral in the polar angle (SemiAnalytical.hpp:13-71, Laplace branch) */
BEM_HD void line_int(double& G, double& dGdn, double z, double x, double v1, double v2) {
  const double theta1 = atan2(v1, x), theta2 = atan2(v2, x);
  const double dtheta = theta2 - theta1, thetam = (theta2 + theta1) / 2;
  const double absZ = fabs(z), signZ = absZ < 1e-10 ? 0 : z / absZ;
  const double xk[5] = {-9.06179846e-01, -5.38469310e-01, 1.78162900e-17, 9.06179846e-01, 5.38469310e-01};
  const double wk[5] = {0.23692689, 0.47862867, 0.56888889, 0.23692689, 0.47862867};
  for (int i = 0; i < 5; ++i) {
    const double thetak = dtheta / 2 * xk[i] + thetam;
    const double Rtheta = x / cos(thetak);
    const double R = sqrt(Rtheta * Rtheta + z * z);
    G += wk[i] * (R - absZ) * dtheta / 2;
    dGdn += wk[i] * (z / R - signZ) * dtheta / 2;
  }
}
/** Contribution of one panel edge (SemiAnalytical.hpp:81-145) */
BEM_HD void int_side(double& G, double& dGdn, const double* v1, const double* v2, double p) {
  const double v21[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
  const double L21 = norm3(v21);
  const double v21u[3] = {v21[0] / L21, v21[1] / L21, v21[2] / L21};
  const double unit[3] = {0, 0, 1};
  double orthog[3], rot[9], v1new[3], v2new[3];
  cross3(unit, v21u, orthog);
  for (int i = 0; i < 3; ++i) { rot[i * 3] = orthog[i]; rot[i * 3 + 1] = v21u[i]; rot[i * 3 + 2] = unit[i]; }
  matvec3(rot, v1, v1new);
  if (v1new[0] < 0) {
    for (int i = 0; i < 9; ++i) rot[i] = -rot[i];
    rot[8] = 1.;
    matvec3(rot, v1, v1new);
  }
  matvec3(rot, v2, v2new);
  const double x = v1new[0];
  if ((v1new[1] > 0 && v2new[1] < 0) || (v1new[1] < 0 && v2new[1] > 0)) {
    double G1 = 0, d1 = 0, G2 = 0, d2 = 0;
    line_int(G1, d1, p, x, 0, v1new[1]);
    line_int(G2, d2, p, x, v2new[1], 0);
    G += G1 + G2;
    dGdn += d1 + d2;
  } else {
    double G1 = 0, d1 = 0;
    line_int(G1, d1, p, x, v1new[1], v2new[1]);
    G -= G1;
    dGdn -= d1;
  }
}
/** Semi-analytical integral of 1/r over the panel seen from x (SemiAnalytical.hpp:148-203) */
BEM_HD double semi_analytical_G(const Panel& s, const double* x) {
  double xp[3], y1p[3], y2p[3];
  const double y0p[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) { xp[k] = x[k] - s.v[0][k]; y1p[k] = s.v[1][k] - s.v[0][k]; y2p[k] = s.v[2][k] - s.v[0][k]; }
  double X[3] = {y1p[0], y1p[1], y1p[2]}, Y[3], Z[3];
  cross3(y1p, y2p, Z);
  const double Xn = norm3(X), Zn = norm3(Z);
  for (int k = 0; k < 3; ++k) { X[k] /= Xn; Z[k] /= Zn; }
  cross3(Z, X, Y);
  const double rot[9] = {X[0], X[1], X[2], Y[0], Y[1], Y[2], Z[0], Z[1], Z[2]};
  double p0[3], p1[3], p2[3], xpl[3], f0[3], f1[3], f2[3];
  matvec3(rot, y0p, p0); matvec3(rot, y1p, p1); matvec3(rot, y2p, p2); matvec3(rot, xp, xpl);
  for (int k = 0; k < 3; ++k) { f0[k] = p0[k] - xpl[k]; f1[k] = p1[k] - xpl[k]; f2[k] = p2[k] - xpl[k]; }
  f0[2] = p0[2]; f1[2] = p1[2]; f2[2] = p2[2];
  double G = 0, dGdn = 0;
  int_side(G, dGdn, f0, f1, xpl[2]);
  int_side(G, dGdn, f1, f2, xpl[2]);
  int_side(G, dGdn, f2, f0, xpl[2]);
  return G;
}

/** int G over the source panel seen from t (eval_G, LaplaceSphericalBEM.hpp:159-205) */
BEM_HD double eval_G(const Panel& s, const double* t, const Rule& rule) {
  const double d[3] = {t[0] - s.c[0], t[1] - s.c[1], t[2] - s.c[2]};
  const double dist = norm3(d);
  if (sqrt(2 * s.area) / dist >= 0.5) return semi_analytical_G(s, t);
  double r = 0;
  for (int i = 0; i < rule.n; ++i) {
    double q[3];
    quad_point(s, rule.pt[i], q);
    const double e[3] = {t[0] - q[0], t[1] - q[1], t[2] - q[2]};
    r += rule.w[i] * s.area / norm3(e);
  }
  return r;
}
/** int dG/dn over the source panel seen from t (eval_dGdn, LaplaceSphericalBEM.hpp:208-264) */
BEM_HD double eval_dGdn(const Panel& s, const double* t, const Rule& rule, const Rule& fine) {
  const double d[3] = {t[0] - s.c[0], t[1] - s.c[1], t[2] - s.c[2]};
  const double dist = norm3(d);
  if (dist < 1e-8) return 2 * M_PI;
  const Rule& g = (sqrt(2 * s.area) / dist >= 0.5) ? fine : rule;
  double r = 0;
  for (int i = 0; i < g.n; ++i) {
    double q[3];
    quad_point(s, g.pt[i], q);
    const double dx[3] = {q[0] - t[0], q[1] - t[1], q[2] - t[2]};
    const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    const double r3 = r2 * sqrt(r2);
    r += g.w[i] * s.area * (dx[0] * s.nrm[0] + dx[1] * s.nrm[1] + dx[2] * s.nrm[2]) / r3;
  }
  return r;
}
// ---- YukawaCartesianBEM near field (reference kernel/YukawaCartesianBEM.hpp:145-204 with the YUKAWA branch of
// ---- examples/BEM/SemiAnalytical.hpp:13-203): exp(-kappa r) / r and its normal derivative ----------------------
BEM_HD void line_int_yk(double& G, double& dGdn, double z, double x, double v1, double v2, double kappa) {
  const double theta1 = atan2(v1, x), theta2 = atan2(v2, x);
  const double dtheta = theta2 - theta1, thetam = (theta2 + theta1) / 2;
  const double absZ = fabs(z), signZ = absZ < 1e-10 ? 0 : z / absZ;
  const double expKz = exp(-kappa * absZ);
  const double xk[5] = {-9.06179846e-01, -5.38469310e-01, 1.78162900e-17, 9.06179846e-01, 5.38469310e-01};
  const double wk[5] = {0.23692689, 0.47862867, 0.56888889, 0.23692689, 0.47862867};
  for (int i = 0; i < 5; ++i) {
    const double thetak = dtheta / 2 * xk[i] + thetam;
    const double Rtheta = x / cos(thetak);
    const double R = sqrt(Rtheta * Rtheta + z * z);
    const double expKr = exp(-kappa * R);
    if (kappa > 1e-10) {
      G += -wk[i] * (expKr - expKz) / kappa * dtheta / 2;
      dGdn += wk[i] * (z / R * expKr - expKz * signZ) * dtheta / 2;
    } else {                                          // devolves to Laplace
      G += wk[i] * (R - absZ) * dtheta / 2;
      dGdn += wk[i] * (z / R - signZ) * dtheta / 2;
    }
  }
}
BEM_HD void int_side_yk(double& G, double& dGdn, const double* v1, const double* v2, double p, double kappa) {
  const double v21[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
  const double L21 = norm3(v21);
  const double v21u[3] = {v21[0] / L21, v21[1] / L21, v21[2] / L21};
  const double unit[3] = {0, 0, 1};
  double orthog[3], rot[9], v1new[3], v2new[3];
  cross3(unit, v21u, orthog);
  for (int i = 0; i < 3; ++i) { rot[i * 3] = orthog[i]; rot[i * 3 + 1] = v21u[i]; rot[i * 3 + 2] = unit[i]; }
  matvec3(rot, v1, v1new);
  if (v1new[0] < 0) {
    for (int i = 0; i < 9; ++i) rot[i] = -rot[i];
    rot[8] = 1.;
    matvec3(rot, v1, v1new);
  }
  matvec3(rot, v2, v2new);
  const double x = v1new[0];
  if ((v1new[1] > 0 && v2new[1] < 0) || (v1new[1] < 0 && v2new[1] > 0)) {
    double G1 = 0, d1 = 0, G2 = 0, d2 = 0;
    line_int_yk(G1, d1, p, x, 0, v1new[1], kappa);
    line_int_yk(G2, d2, p, x, v2new[1], 0, kappa);
    G += G1 + G2;
    dGdn += d1 + d2;
  } else {
    double G1 = 0, d1 = 0;
    line_int_yk(G1, d1, p, x, v1new[1], v2new[1], kappa);
    G -= G1;
    dGdn -= d1;
  }
}
BEM_HD void semi_analytical_yk(const Panel& s, const double* x, double kappa, double& G, double& dGdn) {
  double xp[3], y1p[3], y2p[3];
  const double y0p[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) { xp[k] = x[k] - s.v[0][k]; y1p[k] = s.v[1][k] - s.v[0][k]; y2p[k] = s.v[2][k] - s.v[0][k]; }
  double X[3] = {y1p[0], y1p[1], y1p[2]}, Y[3], Z[3];
  cross3(y1p, y2p, Z);
  const double Xn = norm3(X), Zn = norm3(Z);
  for (int k = 0; k < 3; ++k) { X[k] /= Xn; Z[k] /= Zn; }
  cross3(Z, X, Y);
  const double rot[9] = {X[0], X[1], X[2], Y[0], Y[1], Y[2], Z[0], Z[1], Z[2]};
  double p0[3], p1[3], p2[3], xpl[3], f0[3], f1[3], f2[3];
  matvec3(rot, y0p, p0); matvec3(rot, y1p, p1); matvec3(rot, y2p, p2); matvec3(rot, xp, xpl);
  for (int k = 0; k < 3; ++k) { f0[k] = p0[k] - xpl[k]; f1[k] = p1[k] - xpl[k]; f2[k] = p2[k] - xpl[k]; }
  f0[2] = p0[2]; f1[2] = p1[2]; f2[2] = p2[2];
  G = 0; dGdn = 0;
  int_side_yk(G, dGdn, f0, f1, xpl[2], kappa);
  int_side_yk(G, dGdn, f1, f2, xpl[2], kappa);
  int_side_yk(G, dGdn, f2, f0, xpl[2], kappa);
}
BEM_HD double eval_G_yk(const Panel& s, const double* t, const Rule& rule, double kappa) {
  const double d[3] = {t[0] - s.c[0], t[1] - s.c[1], t[2] - s.c[2]};
  const double dist = norm3(d);
  if (sqrt(2 * s.area) / dist >= 0.5) {
    double G, dGdn;
    semi_analytical_yk(s, t, kappa, G, dGdn);
    return G;
  }
  double r = 0;
  for (int i = 0; i < rule.n; ++i) {
    double q[3];
    quad_point(s, rule.pt[i], q);
    const double e[3] = {t[0] - q[0], t[1] - q[1], t[2] - q[2]};
    const double dd = norm3(e);
    const double inv = dd < 1e-8 ? 0. : 1. / dd;
    r += rule.w[i] * s.area * exp(-kappa * dd) * inv;
  }
  return r;
}
BEM_HD double eval_dGdn_yk(const Panel& s, const double* t, const Rule& rule, double kappa) {
  const double d[3] = {t[0] - s.c[0], t[1] - s.c[1], t[2] - s.c[2]};
  const double dist = norm3(d);
  if (dist < 1e-8) return 2 * M_PI;
  if (sqrt(2 * s.area) / dist >= 0.5) {
    double G, dGdn;
    semi_analytical_yk(s, t, kappa, G, dGdn);
    return -dGdn;
  }
  double res = 0;
  for (int i = 0; i < rule.n; ++i) {
    double q[3];
    quad_point(s, rule.pt[i], q);
    const double dx[3] = {t[0] - q[0], t[1] - q[1], t[2] - q[2]};
    const double r = norm3(dx);
    double inv_r = 1. / r, inv_r2 = inv_r * inv_r;
    if (r < 1e-8) { inv_r = 0.; inv_r2 = 0.; }
    const double f = exp(-kappa * r) * inv_r * (kappa * r + 1) * inv_r2;
    res += rule.w[i] * s.area * (-(dx[0] * f) * s.nrm[0] - (dx[1] * f) * s.nrm[1] - (dx[2] * f) * s.nrm[2]);
  }
  return res;
}
/** YukawaCartesianBEM::operator() (:213-230) */
BEM_HD double kernel_yk(int target_bc, const double* target_centre, const Panel& s, const Rule& rule, double kappa) {
  return target_bc == 0 ? eval_G_yk(s, target_centre, rule, kappa) : eval_dGdn_yk(s, target_centre, rule, kappa);
}

/** K(t, s): the TARGET's boundary condition picks the kernel (operator(), :273-297).
 * bc 0 = POTENTIAL (G), 1 = NORMAL_DERIV (dG/dn). */
BEM_HD double kernel(int target_bc, const double* target_centre, const Panel& s, const Rule& rule, const Rule& fine) {
  return target_bc == 0 ? eval_G(s, target_centre, rule) : eval_dGdn(s, target_centre, rule, fine);
}

}  // namespace bem
