// csrc/laplace_tables.cuh -- factorial tables of LaplaceSpherical::precompute
// (reference kernel/LaplaceSpherical.hpp:87-104) in constant memory.  Each translation unit that
// includes this header gets its own copy and must call upload_laplace_tables() once per device.
#pragma once
#include "common.cuh"
#include <cmath>

namespace fmmb {
// sqrt((n-|m|)!/(n+|m|)!) and (-1)^n/sqrt((n-m)!(n+m)!), index n^2+n+m, n < 2*FMMB_MAX_P.
// The reference carries an extra factor 1/EPS in Anm that cancels in every use; dropped here.
static __constant__ double c_pref[4 * FMMB_MAX_P * FMMB_MAX_P];
static __constant__ double c_anm[4 * FMMB_MAX_P * FMMB_MAX_P];
static __constant__ double c_rcp[2 * FMMB_MAX_P + 2];   // 1/k

static inline void upload_laplace_tables() {
  const int top = 2 * FMMB_MAX_P;
  std::vector<double> pref(top * top), anm(top * top);
  for (int n = 0; n < top; ++n)
    for (int m = -n; m <= n; ++m) {
      int nm = n * n + n + m, am = std::abs(m);
      double fnmm = 1, fnpm = 1, fnma = 1, fnpa = 1;
      for (int i = 1; i <= n - m; ++i) fnmm *= i;
      for (int i = 1; i <= n + m; ++i) fnpm *= i;
      for (int i = 1; i <= n - am; ++i) fnma *= i;
      for (int i = 1; i <= n + am; ++i) fnpa *= i;
      pref[nm] = std::sqrt(fnma / fnpa);
      anm[nm] = ((n & 1) ? -1.0 : 1.0) / std::sqrt(fnmm * fnpm);
    }
  FMMB_CUDA(cudaMemcpyToSymbol(c_pref, pref.data(), pref.size() * sizeof(double)));
  FMMB_CUDA(cudaMemcpyToSymbol(c_anm, anm.data(), anm.size() * sizeof(double)));
  double rcp[2 * FMMB_MAX_P + 2];
  rcp[0] = 0.0;
  for (int k = 1; k < 2 * FMMB_MAX_P + 2; ++k) rcp[k] = 1.0 / k;
  FMMB_CUDA(cudaMemcpyToSymbol(c_rcp, rcp, sizeof rcp));
}
}  // namespace fmmb
