#pragma once
/** @file bem_math.hpp
 * Panel geometry and near-field panel integrals of LaplaceSphericalBEM, written once for the host
 * (kernel class operator(), used by Direct::matvec checks and the diagonal preconditioner) and for the
 * device (near-field assembly in csrc/bem.cu).
 *
 * Follows reference kernel/LaplaceSphericalBEM.hpp:61-97 (Panel), :159-264 (eval_G / eval_dGdn),
 * :273-297 (operator()), examples/BEM/SemiAnalytical.hpp:13-203 (Laplace branch) and the triangle Gauss rules
 * of examples/BEM/GaussQuadrature.hpp:27-276 (every key of its table: 1, 3, 4, 7, 13, 17, 19, 25, 79).
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

constexpr int kMaxRulePoints = 79;

struct Rule {
  int n;
  double pt[kMaxRulePoints][3];
  double w[kMaxRulePoints];
};

namespace detail {
/** Symmetric triangle rules are stored by orbit: 1 point (a, a, a); 3 points (p, q, q), (q, p, q), (q, q, p);
 * 6 points (p, q, r), (p, r, q), (q, p, r), (q, r, p), (r, p, q), (r, q, p) -- the order in which the reference lists
 * them (examples/BEM/GaussQuadrature.hpp:62-276), which fixes the summation order of every panel integral. */
struct Orbit { int kind; double p, q, r, w; };

inline void expand_orbits(const Orbit* o, int count, Rule& rule) {
  int n = 0;
  for (int i = 0; i < count; ++i) {
    const double p = o[i].p, q = o[i].q, r = o[i].r;
    if (o[i].kind == 1) {
      const double t[1][3] = {{p, p, p}};
      for (int k = 0; k < 3; ++k) rule.pt[n][k] = t[0][k];
      rule.w[n++] = o[i].w;
    } else if (o[i].kind == 3) {
      const double t[3][3] = {{p, q, q}, {q, p, q}, {q, q, p}};
      for (int j = 0; j < 3; ++j) { for (int k = 0; k < 3; ++k) rule.pt[n][k] = t[j][k]; rule.w[n++] = o[i].w; }
    } else {
      const double t[6][3] = {{p, q, r}, {p, r, q}, {q, p, r}, {q, r, p}, {r, p, q}, {r, q, p}};
      for (int j = 0; j < 6; ++j) { for (int k = 0; k < 3; ++k) rule.pt[n][k] = t[j][k]; rule.w[n++] = o[i].w; }
    }
  }
  rule.n = n;
}
}  // namespace detail

/** Triangle Gauss rule filed under k in the reference's table (GaussQuadrature.hpp:41-276): k = 1, 3, 4, 13, 19, 25,
 * 79 points; k = 7 aliases the 4-point rule and the rule filed under 17 has 16 points, both like the reference. */
inline Rule make_rule(int k) {
  using detail::Orbit;
  Rule r = {};
  const double a = 1 / 3.;
  if (k == 1) {
    r.n = 1; r.pt[0][0] = r.pt[0][1] = r.pt[0][2] = 1. / 3; r.w[0] = 1.;
  } else if (k == 3) {
    const double p[3][3] = {{0.5, 0.5, 0.}, {0., 0.5, 0.5}, {0.5, 0., 0.5}};
    r.n = 3;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) r.pt[i][j] = p[i][j]; r.w[i] = 1. / 3; }
  } else if (k == 13) {
    const Orbit o[] = {{1, a, 0, 0, -0.149570044467682},
                       {3, 0.479308067841920, 0.260345966079040, 0, 0.175615257433208},
                       {3, 0.869739794195568, 0.065130102902216, 0, 0.053347235608838},
                       {6, 0.048690315425316, 0.312865496004874, 0.638444188569810, 0.077113760890257}};
    detail::expand_orbits(o, 4, r);
  } else if (k == 17) {
    const Orbit o[] = {{1, a, 0, 0, 0.144315607677787},
                       {3, 0.081414823414554, 0.459292588292723, 0, 0.095091634267285},
                       {3, 0.658861384496480, 0.170569307751760, 0, 0.103217370534718},
                       {3, 0.898905543365938, 0.050547228317031, 0, 0.032458497623198},
                       {6, 0.008394777409958, 0.263112829634638, 0.728492392955404, 0.027230314174435}};
    detail::expand_orbits(o, 5, r);
  } else if (k == 19) {
    const Orbit o[] = {{1, a, 0, 0, 0.097135796282799},
                       {3, 0.020634961602525, 0.489682519198738, 0, 0.031334700227139},
                       {3, 0.125820817014127, 0.437089591492937, 0, 0.077827541004774},
                       {3, 0.623592928761935, 0.188203535619033, 0, 0.079647738927210},
                       {3, 0.910540973211095, 0.044729513394453, 0, 0.025577675658698},
                       {6, 0.036838412054736, 0.221962989160766, 0.741198598784498, 0.043283539377289}};
    detail::expand_orbits(o, 6, r);
  } else if (k == 25) {
    const Orbit o[] = {{1, a, 0, 0, 0.090817990382754},
                       {3, 0.028844733232685, 0.485577633383657, 0, 0.036725957756467},
                       {3, 0.781036849029926, 0.109481575485037, 0, 0.045321059435528},
                       {6, 0.141707219414880, 0.307939838764121, 0.550352941820999, 0.072757916845420},
                       {6, 0.025003534762686, 0.246672560639903, 0.728323904597411, 0.028327242531057},
                       {6, 0.009540815400299, 0.066803251012200, 0.923655933587500, 0.009421666963733}};
    detail::expand_orbits(o, 6, r);
  } else if (k == 79) {
    const Orbit o[] = {{1, a, 0, 0, 0.033057055541624},
                       {3, -0.001900928704400, 0.500950464352200, 0, 0.000867019185663},
                       {3, 0.023574084130543, 0.488212957934729, 0, 0.011660052716448},
                       {3, 0.089726636099435, 0.455136681950283, 0, 0.022876936356421},
                       {3, 0.196007481363421, 0.401996259318289, 0, 0.030448982673938},
                       {3, 0.488214180481157, 0.255892909759421, 0, 0.030624891725355},
                       {3, 0.647023488009788, 0.176488255995106, 0, 0.024368057676800},
                       {3, 0.791658289326483, 0.104170855336758, 0, 0.015997432032024},
                       {3, 0.893862072318140, 0.053068963840930, 0, 0.007698301815602},
                       {3, 0.916762569607942, 0.041618715196029, 0, -0.000632060497488},
                       {3, 0.976836157186356, 0.011581921406822, 0, 0.001751134301193},
                       {6, 0.048741583664839, 0.344855770229001, 0.606402646106160, 0.016465839189576},
                       {6, 0.006314115948605, 0.377843269594854, 0.615842614456541, 0.004839033540485},
                       {6, 0.134316520547348, 0.306635479062357, 0.559048000390295, 0.025804906534650},
                       {6, 0.013973893962392, 0.249419362774742, 0.736606743262866, 0.008471091054441},
                       {6, 0.075549132909764, 0.212775724802802, 0.711675142287434, 0.018354914106280},
                       {6, -0.008368153208227, 0.146965436053239, 0.861402717154987, 0.000704404677908},
                       {6, 0.026686063258714, 0.137726978828923, 0.835586957912363, 0.010112684927462},
                       {6, 0.010547719294141, 0.059696109149007, 0.929756171556853, 0.003573909385950}};
    detail::expand_orbits(o, 19, r);
  } else {
    const double p[4][3] = {{1. / 3, 1. / 3, 1. / 3}, {.6, .2, .2}, {.2, .6, .2}, {.2, .2, .6}};
    const double w[4] = {-27. / 48, 25. / 48, 25. / 48, 25. / 48};
    r.n = 4;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 3; ++j) r.pt[i][j] = p[i][j]; r.w[i] = w[i]; }
  }
  return r;
}
/** the keys of the reference's rule table (GaussQuadrature.hpp:41-276); any other k makes the reference exit */
inline bool rule_supported(int k) {
  return k == 1 || k == 3 || k == 4 || k == 7 || k == 13 || k == 17 || k == 19 || k == 25 || k == 79;
}

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
