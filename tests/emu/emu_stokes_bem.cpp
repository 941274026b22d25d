// tests/emu/emu_stokes_bem.cpp -- TEST INFRASTRUCTURE ONLY (built and run by tests/test_cuda_emulation.py, CPU).
// Executes the StokesSphericalBEM kernels of fmm_bem_relaxed_b200/csrc/stokes_bem.cu -- the shipped source lines, cut
// out by tests/emu/extract_kernels.py -- under the lock-step emulation of tests/emu/cuda_emu.hpp:
//   near  <file>   sbem_setup_kernel, sbem_count_kernel, sbem_assemble_kernel, sbem_gather, sbem_near_kernel on a tree
//                  and near-field lists written by the test (from the oracle); prints nothing, writes the near-field
//                  result in ORIGINAL order to <file>.out
//   pipeline <file>  stokes_bem_setup + stokes_bem_execute, the HOST functions of csrc/stokes_bem.cu as written (launch
//                  configurations checked, translations stubbed: the test mesh has no far-field pairs); writes <file>.out
//   gmres <file>   gmres_solve of csrc/gmres.cu as written (fmmb_gmres on Vec<3> unknowns) over the emulated
//                  stokes_bem_execute: the sphere problem of examples/StokesBEM.cpp; writes the solution to <file>.out
//   direct <file>  sbem_direct_kernel / bem_direct_kernel (fmmb_plan_direct_panels) on panels and targets written by the
//                  test; writes the sums to <file>.out
//   far            sbem_p2m_kernel<0|1> against stokes_p2m_kernel<false|true> of csrc/stokes.cu (hardware-verified
//                  this round) fed with one point source per (panel, quadrature point), and sbem_l2p_kernel against
//                  stokes_l2p_kernel at the panel centres; prints the largest relative differences
//   m2p            bem_m2p_kernel<0|1> (treecode of LaplaceSphericalBEM, csrc/bem.cu) against m2p_kernel of
//                  csrc/laplace.cu (hardware-verified) on a toy tree with random multipoles
//   stokes_m2p     stokes_m2p_kernel (csrc/stokes.cu) and sbem_m2p_kernel (csrc/stokes_bem.cu), the treecode evaluators of
//                  the Stokes classes, against four runs of m2p_kernel combined on the host
//   bem_rules      bem_p2m_kernel<0|1> (csrc/bem.cu) with the 13-, 19-, 25- and 79-point rules against the sum of its
//                  own one-point-rule runs (the K = 1 path is hardware-verified)
//   ykm2p          yk_bem_m2p_kernel<0|1> (treecode of YukawaCartesianBEM, csrc/yukawa.cu) against yk_table_kernel
//                  (the table builder of the hardware-verified M2L) + a host dot product
#include "cuda_emu.hpp"
#include "../../fmm_bem_relaxed_b200/csrc/common.cuh"
#include "../../fmm_bem_relaxed_b200/csrc/laplace_ops.cuh"
#include "../../fmm_bem_relaxed_b200/hostcxx/stokes_bem_math.hpp"
#include <random>

#define asm(...) ((void)0)   // the PTX rsqrt of the Stokes pair kernel (compiled, never launched here)
namespace fmmb {
namespace emu_stokes {
#include "stokes_kernels.inc"
}
// csrc/stokes_bem.cu whole, kernels AND host functions (launches rewritten to emu::launch_cfg), at fmmb scope so that
// fmmb_plan::sbem points at its StokesBemData
#include "sbem_whole.inc"
namespace emu_sbem = ::fmmb;
namespace emu_whole = ::fmmb;
// csrc/gmres.cu whole (its anonymous-namespace helper nblk renamed: stokes_bem.cu has one of that name at this scope)
#define nblk gm_nblk
#include "gmres_whole.inc"
#undef nblk
namespace emu_trans {          // the per-pair translation kernels of csrc/laplace.cu (the path behind P > 8 and m2l_mode 1)
using namespace ops;
#include "lap_trans.inc"
}
namespace emu_m2p {            // treecode: the point kernel of csrc/laplace.cu and the panel kernel of csrc/bem.cu
using namespace ops;
#include "lap_m2p.inc"
#include "bem_m2p.inc"
}
namespace emu_bem {            // LaplaceSphericalBEM kernels of csrc/bem.cu
#include "bem_kernels.inc"
}
namespace emu_yk {             // YukawaCartesian[BEM] kernels of csrc/yukawa.cu
const bem::Panel* bem_panels(const BemData* b);
const int* bem_bc(const BemData* b);
#include "yukawa_kernels.inc"
}
}  // namespace fmmb
#undef asm

using namespace fmmb;

static std::vector<char> slurp(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> b(n);
  if (fread(b.data(), 1, n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
  fclose(f);
  return b;
}
template <class T> static const T* take(const char*& p, size_t n) { const T* r = (const T*)p; p += n * sizeof(T); return r; }
static int nblocks(long n, int t) { return (int)((n + t - 1) / t); }

// ---- near field --------------------------------------------------------------------------------------------------
// file layout (little endian): int64 n, nb, ni, ne; int32 K, kfine, as_written, pad; double mu;
// verts[9n] f64, bc[n] i32, perm[n] u32, bb[nb] u32, be[nb] u32, off[nb+1] i32, src[ne] i32, items[4 ni] i32, q[3n] f64
static int run_near(const char* path) {
  std::vector<char> buf = slurp(path);
  const char* p = buf.data();
  const long long* hd = take<long long>(p, 4);
  const long n = hd[0], nb = hd[1], ni = hd[2], ne = hd[3];
  const int* ip = take<int>(p, 4);
  const int K = ip[0], kfine = ip[1], as_written = ip[2];
  const double mu = *take<double>(p, 1);
  const double* verts = take<double>(p, 9 * n);
  const int* bc = take<int>(p, n);
  const unsigned* perm = take<unsigned>(p, n);
  const unsigned* bb = take<unsigned>(p, nb);
  const unsigned* be = take<unsigned>(p, nb);
  const int* off = take<int>(p, nb + 1);
  const int* src = take<int>(p, ne);
  const int4* items = (const int4*)take<int>(p, 4 * ni);
  const double* q = take<double>(p, 3 * n);

  using namespace emu_sbem;
  c_srule = bem::make_rule(K);
  c_sfine = bem::make_rule(kfine);
  std::vector<bem::Panel> pan(n);
  std::vector<int> bct(n);
  emu::launch(dim3(nblocks(n, 128)), dim3(128), [&] { sbem_setup_kernel(verts, bc, perm, n, pan.data(), bct.data()); });
  std::vector<long long> cnt(ni + 1), base(ni + 1, 0);
  emu::launch(dim3(nblocks(ni + 1, 128)), dim3(128), [&] { sbem_count_kernel(items, (int)ni, bb, be, off, src, cnt.data()); });
  for (long i = 0; i < ni; ++i) base[i + 1] = base[i] + cnt[i];
  std::vector<double> val(kSbemEntries * (size_t)base[ni], -7.0), chg(3 * n), res(3 * n, -7.0);
  emu::launch(dim3(nblocks(ni, kSbemWarps)), dim3(32 * kSbemWarps), [&] {
    sbem_assemble_kernel(items, (int)ni, bb, be, off, src, pan.data(), bct.data(), base.data(), mu, as_written != 0, val.data());
  });
  emu::launch(dim3(nblocks(3 * n, 256)), dim3(256), [&] { sbem_gather(q, perm, n, chg.data()); });
  emu::launch(dim3(nblocks(ni, kSbemWarps)), dim3(32 * kSbemWarps), [&] {
    sbem_near_kernel(items, (int)ni, bb, be, off, src, chg.data(), base.data(), val.data(), res.data());
  });
  std::vector<double> out(3 * n);
  for (long i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) out[3 * (size_t)perm[i] + c] = res[3 * i + c];
  std::string o = std::string(path) + ".out";
  FILE* f = fopen(o.c_str(), "wb");
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  printf("near: n %ld items %ld pairs %lld\n", n, ni, base[ni]);
  return 0;
}

// ---- host sequencing and launch configurations of csrc/stokes_bem.cu: stokes_bem_setup + stokes_bem_execute run as
// written (DevBuf on host memory, every launch with its own grid / block / dynamic shared size, guard bytes behind the
// shared segment) on a tree written by the test.  The test mesh has no far-field pairs, so the translations -- the one
// thing not emulated -- are a stub; the far-field kernels still launch with their real configurations.
namespace fmmb {
static int g_translation_calls = 0;
// csrc/laplace.cu::laplace_translations, its per-pair path (the one behind P > 8 and m2l_mode = 1), with the emulated
// kernels and the launch configurations of :943-981: M2M level sweep, M2L per target box, L2L level sweep, on
// plan->M / plan->L.  The class-batched DMMA path is not emulated (mma.sync PTX).  No far-field pairs: nothing to do.
void laplace_translations(fmmb_plan* plan, cudaStream_t) {
  ++g_translation_calls;
  Tree& T = plan->tree;
  if (T.n_lr == 0) return;
  const int P = plan->p, nc = P * (P + 1) / 2, pp = P * P, nb = T.nboxes;
  const size_t sh_mm = (size_t)(pp + nc) * sizeof(double2);
  for (int l = T.nlevels - 2; l >= 0; --l) {
    const int lo = T.level_off[l], hi = T.level_off[l + 1];
    emu::launch_cfg(hi - lo, 64, sh_mm, [&] {
      emu_trans::m2m_kernel(lo, hi, nullptr, T.key.p, T.cbegin.p, T.cend.p, T.center.p, P, plan->M.p);
    });
  }
  if (plan->opts.evaluator == FMMB_EVAL_TREECODE) return;
  static std::map<int, std::vector<double>> coeff;
  std::vector<double>& C = coeff[P];
  if (C.empty()) {
    C.resize((size_t)nc * pp);
    emu::launch_cfg((int)((C.size() + 255) / 256), 256, 0, [&] { emu_trans::m2l_coeff_kernel(P, C.data()); });
  }
  int threads = 128;
  while (threads < nc) threads += 32;
  size_t sh = (size_t)(5 * pp) * sizeof(double2);
  const size_t red = (size_t)(threads / nc) * nc * sizeof(double2);
  if (red > sh) sh = red;
  emu::launch_cfg(nb, threads, sh, [&] {
    emu_trans::m2l_pair_kernel(nb, nullptr, T.m2l_off.p, T.m2l_src.p, T.center.p, P, C.data(), plan->M.p, plan->L.p, 0);
  });
  for (int l = 1; l < T.nlevels; ++l) {
    const int lo = T.level_off[l], hi = T.level_off[l + 1];
    emu::launch_cfg(hi - lo, 64, sh_mm, [&] {
      emu_trans::l2l_kernel(lo, hi, T.parent.p, T.has_local.p, T.center.p, P, plan->L.p);
    });
  }
}
void run_matvec_for_solver(fmmb_plan* plan, const double* q, double* r) { stokes_bem_execute(plan, q, r); }   // capi.cu
void finish_results(fmmb_plan* plan, const double* near, const double* far, int rd, double* d_results, cudaStream_t) {
  Tree& T = plan->tree;                                   // csrc/comm.cu: gen_combine_scatter on one rank
  for (int64_t i = T.own_b0; i < T.own_b1; ++i)
    for (int c = 0; c < rd; ++c) d_results[(size_t)T.perm.p[i] * rd + c] = near[i * rd + c] + far[i * rd + c];
}
}  // namespace fmmb

// same file layout as `near`, followed by geom[4 nb] f64 (box centre x, y, z, side), parent[nb] u32, leaf[nb] u8 padded
// to 4-byte entries (i32), P (i32), treecode (i32)
static int run_pipeline(const char* path, bool solve = false) {
  std::vector<char> buf = slurp(path);
  const char* p = buf.data();
  const long long* hd = take<long long>(p, 4);
  const long n = hd[0], nb = hd[1], ni = hd[2], ne = hd[3];
  const int* ip = take<int>(p, 4);
  const int K = ip[0], kfine = ip[1], as_written = ip[2];
  const double mu = *take<double>(p, 1);
  const double* verts = take<double>(p, 9 * n);
  const int* bc = take<int>(p, n);
  const unsigned* perm = take<unsigned>(p, n);
  const unsigned* bb = take<unsigned>(p, nb);
  const unsigned* be = take<unsigned>(p, nb);
  const int* off = take<int>(p, nb + 1);
  const int* src = take<int>(p, ne);
  const int4* items = (const int4*)take<int>(p, 4 * ni);
  const double* q = take<double>(p, 3 * n);
  const double* geom = take<double>(p, 4 * nb);
  const unsigned* parent = take<unsigned>(p, nb);
  const int* leaf = take<int>(p, nb);
  const int P = *take<int>(p, 1), treecode = *take<int>(p, 1);
  // optional far-field section: int64 n_lr, int32 nlevels, pad; key[nb] u32, cbegin[nb] u32, cend[nb] u32,
  // level_off[nlevels + 1] i32, m2l_off[nb + 1] i32, m2l_src[n_lr] i32, has_local[nb] i32
  long n_lr = 0;
  int nlevels = 0;
  const unsigned *key = nullptr, *cbegin = nullptr, *cend = nullptr;
  const int *level_off = nullptr, *m2l_off = nullptr, *m2l_src = nullptr, *has_local_i = nullptr;
  if (p < buf.data() + buf.size()) {
    n_lr = (long)*take<long long>(p, 1);
    nlevels = take<int>(p, 2)[0];
    key = take<unsigned>(p, nb); cbegin = take<unsigned>(p, nb); cend = take<unsigned>(p, nb);
    level_off = take<int>(p, nlevels + 1);
    m2l_off = take<int>(p, nb + 1); m2l_src = take<int>(p, n_lr);
    has_local_i = take<int>(p, nb);
  }

  upload_laplace_tables();
  fmmb_plan plan;
  plan.bem_near_kernel = 0;   // the one-warp-per-item kernel: the split kernel's warp shuffles are not emulated
  plan.kind = FMMB_STOKES_SPHERICAL_BEM;
  plan.p = P;
  std::memset(&plan.opts, 0, sizeof plan.opts);
  plan.opts.kernel_flags = as_written ? FMMB_FLAG_STOKES_BEM_AS_WRITTEN : 0;
  plan.opts.evaluator = treecode ? FMMB_EVAL_TREECODE : FMMB_EVAL_FMM;
  plan.charge_dim = plan.result_dim = 3;
  Tree& T = plan.tree;
  T.n = n; T.nboxes = (int)nb; T.own_b0 = 0; T.own_b1 = n;
  T.perm.from_host(perm, n, nullptr);
  T.bbegin.from_host(bb, nb, nullptr); T.bend.from_host(be, nb, nullptr);
  T.parent.from_host(parent, nb, nullptr);
  T.p2p_off.from_host(off, nb + 1, nullptr); T.p2p_src.from_host(src, ne, nullptr);
  T.p2p_items.from_host(items, ni, nullptr); T.n_p2p_items = (int)ni;
  std::vector<double4> cen(nb);
  std::vector<int> leaves;
  for (long b = 0; b < nb; ++b) { cen[b] = make_double4(geom[4 * b], geom[4 * b + 1], geom[4 * b + 2], geom[4 * b + 3]); if (leaf[b]) leaves.push_back((int)b); }
  T.center.from_host(cen.data(), nb, nullptr);
  T.leaves.from_host(leaves.data(), leaves.size(), nullptr); T.nleaves = (int)leaves.size();
  T.own_leaves.from_host(leaves.data(), leaves.size(), nullptr); T.n_own_leaves = (int)leaves.size();
  std::vector<unsigned char> hl(nb, 0);
  std::vector<int> zoff(nb + 1, 0);
  if (n_lr > 0) {
    for (long b = 0; b < nb; ++b) hl[b] = (unsigned char)has_local_i[b];
    T.n_lr = n_lr; T.nlevels = nlevels;
    T.level_off.assign(level_off, level_off + nlevels + 1);
    T.key.from_host(key, nb, nullptr); T.cbegin.from_host(cbegin, nb, nullptr); T.cend.from_host(cend, nb, nullptr);
    T.m2l_off.from_host(m2l_off, nb + 1, nullptr); T.m2l_src.from_host(m2l_src, n_lr, nullptr);
  } else {
    T.m2l_off.from_host(zoff.data(), nb + 1, nullptr); T.m2l_src.resize(1);
  }
  T.has_local.from_host(hl.data(), nb, nullptr);

  emu_whole::stokes_bem_setup(&plan, verts, bc, K, kfine, mu);
  if (solve) {
    // fmmb_gmres on Vec<3> unknowns (csrc/gmres.cu as written): the sphere problem of examples/StokesBEM.cpp,
    // b = (4 pi, 0, 0) per panel, x0 = 0, order rule of GMRES_Stokes.hpp:229
    plan.bem = nullptr;
    std::vector<double> b(3 * n), x(3 * n, 0.0), hist(256);
    for (long i = 0; i < n; ++i) { b[3 * i] = 4 * M_PI; b[3 * i + 1] = b[3 * i + 2] = 0.0; }
    fmmb_solver_options so = {1e-5, 100, 100, 8u, 1, 0, 0, 5u, 1u};
    fmmb_gmres_info info = {};
    std::vector<int32_t> ps(256);
    gmres_solve(&plan, b.data(), x.data(), nullptr, so, &info, ps.data(), hist.data(), 256);
    std::string o = std::string(path) + ".out";
    FILE* f = fopen(o.c_str(), "wb");
    fwrite(x.data(), 8, x.size(), f);
    fclose(f);
    printf("gmres: iterations %d final_residual %.6e final_p %d guard_failures %ld schedule", info.iterations, info.final_residual,
           info.final_p, emu::guard_failures);
    for (int k = 0; k < info.n_records && k < 256; ++k) printf(" %d", ps[k]);
    printf("\nresiduals");
    for (int k = 0; k < info.n_records && k < 256; ++k) printf(" %.6e", hist[k]);
    printf("\n");
    gmres_free(plan.gmres_ws);
    emu_whole::stokes_bem_free(plan.sbem);
    return 0;
  }
  std::vector<double> out(3 * n, -11.0), out2(3 * n, -12.0);
  emu_whole::stokes_bem_execute(&plan, q, out.data());
  bool same = true;
  if (n_lr == 0) {                                                  // second call: no reallocation, same result
    emu_whole::stokes_bem_execute(&plan, q, out2.data());           // (skipped with a far field: the emulated
    same = out == out2;                                             // translations dominate the run time)
  }
  std::string o = std::string(path) + ".out";
  FILE* f = fopen(o.c_str(), "wb");
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  printf("pipeline: n %ld launches %ld (reported %d) translation_calls %d guard_failures %ld repeatable %d nnz %lld\n", n, emu::launches,
         plan.launches, g_translation_calls, emu::guard_failures, (int)same, (long long)emu_whole::stokes_bem_nnz(plan.sbem));
  emu_whole::stokes_bem_free(plan.sbem);
  return 0;
}

// ---- Direct::matvec kernels: sbem_direct_kernel (csrc/stokes_bem.cu) and bem_direct_kernel (csrc/bem.cu) ----------
// file layout: int64 n, nt; int32 K, kfine, as_written, kind (0 Stokes BEM, 1 Laplace BEM, 2 Yukawa BEM); double mu_or_kappa;
// verts[9n] f64 (tree order = file order), q[cd n] f64, tverts[9 nt] f64, tbc[nt] i32.  Writes <file>.out (rd nt doubles).
static int run_direct(const char* path) {
  std::vector<char> buf = slurp(path);
  const char* p = buf.data();
  const long long* hd = take<long long>(p, 2);
  const long n = hd[0], nt = hd[1];
  const int* ip = take<int>(p, 4);
  const int K = ip[0], kfine = ip[1], as_written = ip[2], kind = ip[3];
  const double par = *take<double>(p, 1);
  const double* verts = take<double>(p, 9 * n);
  const int cd = kind == 0 ? 3 : 1;
  const double* q = take<double>(p, cd * n);
  const double* tverts = take<double>(p, 9 * nt);
  const int* tbc = take<int>(p, nt);
  std::vector<bem::Panel> pan(n);
  for (long i = 0; i < n; ++i) bem::make_panel(verts + 9 * i, verts + 9 * i + 3, verts + 9 * i + 6, pan[i]);
  std::vector<double> out(cd * nt, -5.0);
  if (kind == 0) {
    emu_sbem::c_srule = bem::make_rule(K);
    emu_sbem::c_sfine = bem::make_rule(kfine);
    emu::launch(dim3((unsigned)nt), dim3(128), [&] {
      emu_sbem::sbem_direct_kernel(pan.data(), q, n, tverts, tbc, par, as_written != 0, out.data());
    });
  } else {
    emu_bem::c_rule = bem::make_rule(K == 7 ? 4 : K);
    emu_bem::c_fine = bem::make_rule(17);
    emu::launch(dim3((unsigned)nt), dim3(128), [&] {
      emu_bem::bem_direct_kernel(pan.data(), q, n, tverts, tbc, kind == 2 ? par : -1.0, out.data());
    });
  }
  std::string o = std::string(path) + ".out";
  FILE* f = fopen(o.c_str(), "wb");
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  printf("direct: n %ld nt %ld kind %d\n", n, nt, kind);
  return 0;
}

// ---- far field ---------------------------------------------------------------------------------------------------
static double rel_diff(const std::vector<double>& a, const std::vector<double>& b) {
  double e = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) { e = std::max(e, std::fabs(a[i] - b[i])); s = std::max(s, std::fabs(b[i])); }
  return e / s;
}

static int run_far() {
  upload_laplace_tables();
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(0., 1.);
  // three "leaves" of 37, 1 and 64 small panels inside unit-ish boxes; box 1 has no local expansion
  const int nleaf = 3, sizes[nleaf] = {37, 1, 64};
  std::vector<unsigned> bb(nleaf), be(nleaf);
  std::vector<double4> center(nleaf);
  std::vector<int> leaves = {0, 1, 2};
  std::vector<unsigned char> has_local = {1, 0, 1};
  int n = 0;
  for (int b = 0; b < nleaf; ++b) { bb[b] = n; n += sizes[b]; be[b] = n; center[b] = make_double4(0.5 + b, 0.5, 0.25 * b, 1.0); }
  std::vector<bem::Panel> pan(n);
  std::vector<int> bc(n);
  std::vector<double> chg(3 * n);
  for (int b = 0; b < nleaf; ++b)
    for (unsigned i = bb[b]; i < be[b]; ++i) {
      double c[3] = {center[b].x - 0.45 + 0.9 * U(rng), center[b].y - 0.45 + 0.9 * U(rng), center[b].z - 0.45 + 0.9 * U(rng)};
      double v[9];
      for (int k = 0; k < 9; ++k) v[k] = c[k % 3] + 0.04 * (U(rng) - 0.5);
      bem::make_panel(v, v + 3, v + 6, pan[i]);
      bc[i] = (i % 3 == 1) ? 1 : 0;
      for (int k = 0; k < 3; ++k) chg[3 * i + k] = U(rng) - 0.3;
    }
  double worst_p2m = 0, worst_l2p = 0, max_m = 0, max_u = 0;
  const int keys[] = {1, 4, 13};
  for (int P : {3, 8, 11})
    for (int key : keys) {
      const bem::Rule rule = bem::make_rule(key);
      emu_sbem::c_srule = rule;
      const int K = rule.n, xs = ops::xstride(P), pp = P * P;
      const int warps = pp <= 64 ? 4 : 1;
      for (int group = 0; group < 2; ++group) {
        // the kernel under test
        std::vector<double> M[4], R[4];
        for (int s = 0; s < 4; ++s) { M[s].assign((size_t)nleaf * xs, 0.0); R[s].assign((size_t)nleaf * xs, 0.0); }
        emu::launch(dim3(nblocks(nleaf, warps), 4), dim3(32 * warps), [&] {
          if (group == 0) emu_sbem::sbem_p2m_kernel<0>(leaves.data(), nleaf, bb.data(), be.data(), center.data(), pan.data(),
                                                       bc.data(), chg.data(), P, M[0].data(), M[1].data(), M[2].data(), M[3].data());
          else emu_sbem::sbem_p2m_kernel<1>(leaves.data(), nleaf, bb.data(), be.data(), center.data(), pan.data(), bc.data(),
                                            chg.data(), P, M[0].data(), M[1].data(), M[2].data(), M[3].data());
        });
        // comparator: the point-source kernel of csrc/stokes.cu on one body per (panel of this group, quadrature point)
        const int REC = group == 0 ? 6 : 9;
        std::vector<double> srcrec;
        std::vector<unsigned> vb(nleaf), ve(nleaf);
        unsigned m = 0;
        for (int b = 0; b < nleaf; ++b) {
          vb[b] = m;
          for (unsigned i = bb[b]; i < be[b]; ++i) {
            if (bc[i] != group) continue;
            for (int k = 0; k < K; ++k) {
              double qp[3];
              bem::quad_point(pan[i], rule.pt[k], qp);
              const double wa = pan[i].area * rule.w[k];
              for (int c = 0; c < 3; ++c) srcrec.push_back(qp[c]);
              for (int c = 0; c < 3; ++c) srcrec.push_back(wa * chg[3 * i + c]);
              if (group == 1) for (int c = 0; c < 3; ++c) srcrec.push_back(pan[i].nrm[c]);
              ++m;
            }
          }
          ve[b] = m;
        }
        emu::launch(dim3(nblocks(nleaf, warps), 4), dim3(32 * warps), [&] {
          if (group == 0) emu_stokes::stokes_p2m_kernel<false>(leaves.data(), nleaf, vb.data(), ve.data(), center.data(),
                                                               srcrec.data(), P, R[0].data(), R[1].data(), R[2].data(), R[3].data());
          else emu_stokes::stokes_p2m_kernel<true>(leaves.data(), nleaf, vb.data(), ve.data(), center.data(), srcrec.data(), P,
                                                   R[0].data(), R[1].data(), R[2].data(), R[3].data());
        });
        for (int s = 0; s < 4; ++s) {
          worst_p2m = std::max(worst_p2m, rel_diff(M[s], R[s]));
          for (double x : M[s]) max_m = std::max(max_m, std::fabs(x));
        }
        (void)REC;
      }
      // L2P: random local expansions; the BEM kernel writes only the targets of its group, scaled
      std::vector<double> L[4];
      for (int s = 0; s < 4; ++s) { L[s].resize((size_t)nleaf * xs); for (auto& x : L[s]) x = U(rng) - 0.5; }
      std::vector<double4> body(n);
      for (int i = 0; i < n; ++i) body[i] = make_double4(pan[i].c[0], pan[i].c[1], pan[i].c[2], 0.0);
      const int nc = P * (P + 1) / 2;
      if ((size_t)4 * 4 * nc * sizeof(double2) > sizeof emu::dyn_shared) return 4;
      for (int group = 0; group < 2; ++group) {
        const double scale = group == 0 ? 1. / 2 / 0.037 : 0.5;
        std::vector<double> got(3 * n, 0.0), want(3 * n, -1.0);
        emu::launch(dim3(nblocks(nleaf, 4)), dim3(128), [&] {
          emu_sbem::sbem_l2p_kernel(leaves.data(), nleaf, bb.data(), be.data(), center.data(), has_local.data(), pan.data(),
                                    bc.data(), group, P, L[0].data(), L[1].data(), L[2].data(), L[3].data(), scale, got.data());
        });
        emu::launch(dim3(nblocks(nleaf, 4)), dim3(128), [&] {
          emu_stokes::stokes_l2p_kernel(leaves.data(), nleaf, bb.data(), be.data(), center.data(), has_local.data(), body.data(),
                                        P, L[0].data(), L[1].data(), L[2].data(), L[3].data(), scale, want.data());
        });
        for (int i = 0; i < n; ++i)
          if (bc[i] != group) for (int c = 0; c < 3; ++c) want[3 * i + c] = 0.0;     // untouched by the BEM kernel
        worst_l2p = std::max(worst_l2p, rel_diff(got, want));
        for (double x : got) max_u = std::max(max_u, std::fabs(x));
      }
    }
  printf("far: p2m %.3e l2p %.3e max_multipole %.3e max_velocity %.3e\n", worst_p2m, worst_l2p, max_m, max_u);
  return 0;
}

// ---- treecode: bem_m2p_kernel<SET> (csrc/bem.cu) against m2p_kernel (csrc/laplace.cu, hardware-verified) ----------
// a three-level toy tree: root 0, children 1..2 (box 2 is a leaf), leaves 3..4 under box 1; interaction lists of random
// source boxes for every box; random multipoles.  The panel kernel must add (set 0) or subtract (set 1) the potential
// component of the point kernel at the centres of the panels of its set, and leave the others alone.
static int run_m2p() {
  upload_laplace_tables();
  std::mt19937_64 rng(13);
  std::uniform_real_distribution<double> U(0., 1.);
  const int nb = 5;
  std::vector<unsigned> parent = {0, 0, 0, 1, 1}, bb = {0, 0, 70, 0, 33}, be = {103, 70, 103, 33, 70};
  std::vector<int> leaves = {2, 3, 4};
  std::vector<int> off = {0, 0, 2, 5, 6, 9}, src = {2, 4, 1, 3, 4, 2, 2, 1, 3};   // far boxes (any box with a multipole)
  std::vector<double4> center(nb);
  for (int b = 0; b < nb; ++b) center[b] = make_double4(3.0 * b, -2.0 + b, 1.5 * b, 1.0);
  const int n = 103;
  std::vector<bem::Panel> pan(n);
  std::vector<int> bc(n);
  std::vector<double4> body(n);
  for (int i = 0; i < n; ++i) {
    double c[3] = {20 + U(rng), 20 + U(rng), 20 + U(rng)}, v[9];     // far from every source box
    for (int k = 0; k < 9; ++k) v[k] = c[k % 3] + 0.04 * (U(rng) - 0.5);
    bem::make_panel(v, v + 3, v + 6, pan[i]);
    bc[i] = i % 2;
    body[i] = make_double4(pan[i].c[0], pan[i].c[1], pan[i].c[2], 0.0);
  }
  double worst = 0, biggest = 0;
  for (int P : {2, 8, 13}) {
    const int xs = ops::xstride(P), nc = P * (P + 1) / 2;
    std::vector<double> M((size_t)nb * xs);
    for (auto& x : M) x = U(rng) - 0.5;
    std::vector<double4> ref(n, make_double4(0, 0, 0, 0));
    emu::launch(dim3(nblocks(3, 4)), dim3(128), [&] {
      emu_m2p::m2p_kernel(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(), center.data(),
                          body.data(), P, M.data(), ref.data());
    });
    for (int set = 0; set < 2; ++set) {
      std::vector<double> got(n, 0.25), want(n, 0.25);
      emu::launch(dim3(nblocks(3, 4)), dim3(128), [&] {
        if (set == 0) emu_m2p::bem_m2p_kernel<0>(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(),
                                                 center.data(), pan.data(), bc.data(), P, M.data(), got.data());
        else emu_m2p::bem_m2p_kernel<1>(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(),
                                        center.data(), pan.data(), bc.data(), P, M.data(), got.data());
      });
      for (int i = 0; i < n; ++i) if (bc[i] == set) want[i] += set == 0 ? ref[i].x : -ref[i].x;
      for (int i = 0; i < n; ++i) biggest = std::max(biggest, std::fabs(ref[i].x));
      worst = std::max(worst, rel_diff(got, want));
    }
    (void)nc;
  }
  printf("m2p: %.3e max_potential %.3e\n", worst, biggest);
  return 0;
}

// ---- Gauss rules above 4 points in the LaplaceSphericalBEM P2M: bem_p2m_kernel<SET> with the K-point rule against the
// sum of K runs of the same kernel with the one-point rules (pt_k, w_k) -- the one-point path is the K = 1 rule that is
// green on hardware; what is new with K = 13, 19, 25, 79 is the (panel, quadrature point) -> lane mapping.
static int run_bem_rules() {
  upload_laplace_tables();
  std::mt19937_64 rng(29);
  std::uniform_real_distribution<double> U(0., 1.);
  const int nleaf = 2, sizes[nleaf] = {41, 7};
  std::vector<unsigned> bb(nleaf), be(nleaf);
  std::vector<double4> center(nleaf);
  std::vector<int> leaves = {0, 1};
  int n = 0;
  for (int b = 0; b < nleaf; ++b) { bb[b] = n; n += sizes[b]; be[b] = n; center[b] = make_double4(0.5 + b, 0.5, 0.25 * b, 1.0); }
  std::vector<bem::Panel> pan(n);
  std::vector<int> bc(n);
  std::vector<double4> body(n);
  for (int b = 0; b < nleaf; ++b)
    for (unsigned i = bb[b]; i < be[b]; ++i) {
      double c[3] = {center[b].x - 0.45 + 0.9 * U(rng), center[b].y - 0.45 + 0.9 * U(rng), center[b].z - 0.45 + 0.9 * U(rng)}, v[9];
      for (int k = 0; k < 9; ++k) v[k] = c[k % 3] + 0.04 * (U(rng) - 0.5);
      bem::make_panel(v, v + 3, v + 6, pan[i]);
      bc[i] = i % 2;
      body[i] = make_double4(pan[i].c[0], pan[i].c[1], pan[i].c[2], U(rng) - 0.3);     // .w = charge
    }
  double worst = 0, biggest = 0;
  for (int P : {4, 8})
    for (int key : {13, 19, 25, 79}) {
      const bem::Rule rule = bem::make_rule(key);
      const int xs = ops::xstride(P), pp = P * P, warps = pp <= 64 ? 4 : 1;
      for (int set = 0; set < 2; ++set) {
        auto run = [&](const bem::Rule& r, std::vector<double>& M) {
          emu_bem::c_rule = r;
          M.assign((size_t)nleaf * xs, 0.0);
          emu::launch(dim3(nblocks(nleaf, warps)), dim3(32 * warps), [&] {
            if (set == 0) emu_bem::bem_p2m_kernel<0>(leaves.data(), nleaf, bb.data(), be.data(), center.data(), body.data(),
                                                     pan.data(), bc.data(), P, M.data());
            else emu_bem::bem_p2m_kernel<1>(leaves.data(), nleaf, bb.data(), be.data(), center.data(), body.data(), pan.data(),
                                            bc.data(), P, M.data());
          });
        };
        std::vector<double> got, part, want((size_t)nleaf * xs, 0.0);
        run(rule, got);
        for (int k = 0; k < rule.n; ++k) {
          bem::Rule one = {};
          one.n = 1;
          for (int c = 0; c < 3; ++c) one.pt[0][c] = rule.pt[k][c];
          one.w[0] = rule.w[k];
          run(one, part);
          for (size_t t = 0; t < want.size(); ++t) want[t] += part[t];
        }
        for (double x : want) biggest = std::max(biggest, std::fabs(x));
        worst = std::max(worst, rel_diff(got, want));
      }
    }
  printf("bem_rules: %.3e max_multipole %.3e\n", worst, biggest);
  return 0;
}

// ---- Stokes treecode: stokes_m2p_kernel (csrc/stokes.cu) and sbem_m2p_kernel (csrc/stokes_bem.cu) against four runs
// of the Laplace point treecode kernel m2p_kernel (csrc/laplace.cu, hardware-verified) -- one per expansion set, its
// potential and Cartesian gradient combined on the host as StokesSpherical.hpp:207-291 prescribes:
//   u_k = scale ( pot_k + sum_{s<3} (-x_s) grad_s[k] + grad_3[k] )
static int run_stokes_m2p() {
  upload_laplace_tables();
  std::mt19937_64 rng(17);
  std::uniform_real_distribution<double> U(0., 1.);
  const int nb = 5;
  std::vector<unsigned> parent = {0, 0, 0, 1, 1}, bb = {0, 0, 70, 0, 33}, be = {103, 70, 103, 33, 70};
  std::vector<int> leaves = {2, 3, 4};
  std::vector<int> off = {0, 0, 2, 5, 6, 9}, src = {2, 4, 1, 3, 4, 2, 2, 1, 3};
  std::vector<double4> center(nb);
  for (int b = 0; b < nb; ++b) center[b] = make_double4(3.0 * b, -2.0 + b, 1.5 * b, 1.0);
  const int n = 103;
  std::vector<bem::Panel> pan(n);
  std::vector<int> bc(n);
  std::vector<double4> body(n);
  for (int i = 0; i < n; ++i) {
    double c[3] = {20 + U(rng), 20 + U(rng), 20 + U(rng)}, v[9];
    for (int k = 0; k < 9; ++k) v[k] = c[k % 3] + 0.04 * (U(rng) - 0.5);
    bem::make_panel(v, v + 3, v + 6, pan[i]);
    bc[i] = (i % 3 == 1) ? 1 : 0;
    body[i] = make_double4(pan[i].c[0], pan[i].c[1], pan[i].c[2], 0.0);
  }
  double worst_pt = 0, worst_bem = 0, biggest = 0;
  for (int P : {2, 8, 13}) {
    const int xs = ops::xstride(P), nc = P * (P + 1) / 2;
    std::vector<double> M[4];
    std::vector<double4> lap[4];
    for (int s4 = 0; s4 < 4; ++s4) {
      M[s4].resize((size_t)nb * xs);
      for (auto& x : M[s4]) x = U(rng) - 0.5;
      lap[s4].assign(n, make_double4(0, 0, 0, 0));
      emu::launch(dim3(nblocks(3, 4)), dim3(128), [&] {
        emu_m2p::m2p_kernel(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(), center.data(),
                            body.data(), P, M[s4].data(), lap[s4].data());
      });
    }
    const double scale = 1. / 6;
    std::vector<double> want(3 * n);
    for (int i = 0; i < n; ++i) {
      const double x[3] = {body[i].x, body[i].y, body[i].z};
      const double g[4][3] = {{lap[0][i].y, lap[0][i].z, lap[0][i].w}, {lap[1][i].y, lap[1][i].z, lap[1][i].w},
                              {lap[2][i].y, lap[2][i].z, lap[2][i].w}, {lap[3][i].y, lap[3][i].z, lap[3][i].w}};
      const double pot[3] = {lap[0][i].x, lap[1][i].x, lap[2][i].x};
      for (int k = 0; k < 3; ++k)
        want[3 * i + k] = scale * (pot[k] - x[0] * g[0][k] - x[1] * g[1][k] - x[2] * g[2][k] + g[3][k]);
    }
    for (double x : want) biggest = std::max(biggest, std::fabs(x));
    if ((size_t)4 * 4 * nc * sizeof(double2) > sizeof emu::dyn_shared) return 4;
    std::vector<double> got(3 * n, -3.0);
    emu::launch(dim3(nblocks(3, 4)), dim3(128), [&] {
      emu_stokes::stokes_m2p_kernel(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(),
                                    center.data(), body.data(), P, M[0].data(), M[1].data(), M[2].data(), M[3].data(), scale,
                                    got.data());
    });
    worst_pt = std::max(worst_pt, rel_diff(got, want));
    for (int group = 0; group < 2; ++group) {
      std::vector<double> gb(3 * n, 0.0), wb(3 * n, 0.0);
      emu::launch(dim3(nblocks(3, 4)), dim3(128), [&] {
        emu_sbem::sbem_m2p_kernel(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(), center.data(),
                                  pan.data(), bc.data(), group, P, M[0].data(), M[1].data(), M[2].data(), M[3].data(), scale,
                                  gb.data());
      });
      for (int i = 0; i < n; ++i) if (bc[i] == group) for (int k = 0; k < 3; ++k) wb[3 * i + k] = want[3 * i + k];
      worst_bem = std::max(worst_bem, rel_diff(gb, wb));
    }
  }
  printf("stokes_m2p: point %.3e bem %.3e max_velocity %.3e\n", worst_pt, worst_bem, biggest);
  return 0;
}

// ---- Yukawa BEM treecode: yk_bem_m2p_kernel<SET> (lane-private Taylor tables) against yk_table_kernel (the block-
// cooperative table builder behind the hardware-verified M2L) + a host dot product with the multipoles --------------
static int run_ykm2p() {
  using namespace emu_yk;
  upload_tables();
  std::mt19937_64 rng(21);
  std::uniform_real_distribution<double> U(0., 1.);
  const int nb = 5;
  std::vector<unsigned> parent = {0, 0, 0, 1, 1}, bb = {0, 0, 40, 0, 17}, be = {57, 40, 57, 17, 40};
  std::vector<int> leaves = {2, 3, 4};
  std::vector<int> off = {0, 0, 2, 5, 6, 9}, src = {2, 4, 1, 3, 4, 2, 2, 1, 3};
  std::vector<double4> center(nb);
  for (int b = 0; b < nb; ++b) center[b] = make_double4(0.7 * b, -1.0 + 0.5 * b, 0.4 * b, 1.0);
  const int n = 57;
  std::vector<bem::Panel> pan(n);
  std::vector<int> bc(n);
  for (int i = 0; i < n; ++i) {
    double c[3] = {6 + U(rng), 5 + U(rng), 4 + U(rng)}, v[9];
    for (int k = 0; k < 9; ++k) v[k] = c[k % 3] + 0.04 * (U(rng) - 0.5);
    bem::make_panel(v, v + 3, v + 6, pan[i]);
    bc[i] = i % 2;
  }
  auto leaf_of = [&](int i) { for (int l : leaves) if ((unsigned)i >= bb[l] && (unsigned)i < be[l]) return l; return -1; };
  double worst = 0, biggest = 0, worst_pt = 0;
  const double kappa = 0.35;
  for (int P : {1, 4, 8, 10}) {
    const int nt = yk_terms(P);
    std::vector<double> M((size_t)nb * nt);
    for (auto& x : M) x = U(rng) - 0.5;
    // expected: one table per (target, accepted box)
    std::vector<double4> vec;
    std::vector<int> pair_t, pair_b;
    for (int i = 0; i < n; ++i)
      for (int a = leaf_of(i);; a = (int)parent[a]) {
        for (int e = off[a]; e < off[a + 1]; ++e) {
          const double4 c = center[src[e]];
          vec.push_back(make_double4(pan[i].c[0] - c.x, pan[i].c[1] - c.y, pan[i].c[2] - c.z, 0));
          pair_t.push_back(i); pair_b.push_back(src[e]);
        }
        if (a == 0) break;
      }
    std::vector<double> table(vec.size() * (size_t)nt);
    emu::launch(dim3((unsigned)vec.size()), dim3(128), [&] { yk_table_kernel<10>(P, kappa, vec.data(), table.data()); });
    std::vector<double> phi(n, 0.0);
    for (size_t k = 0; k < vec.size(); ++k) {
      double v = 0;
      for (int t = 0; t < nt; ++t) v += table[k * nt + t] * M[(size_t)pair_b[k] * nt + t];
      phi[pair_t[k]] += v;
    }
    // the point kernel: potential and gradient (ax_m f_m = a'_{m + e_x} for |m| < P)
    {
      std::vector<double> w4(4 * n, 0.0);
      for (size_t k = 0; k < vec.size(); ++k) {
        const double* tab = &table[k * nt];
        const double* Mb = &M[(size_t)pair_b[k] * nt];
        double* r = &w4[4 * pair_t[k]];
        for (int t = 0; t < nt; ++t) {
          const int ii = c_yI[P][t], jj = c_yJ[P][t], kk = c_yK[P][t];
          r[0] += tab[t] * Mb[t];
          if (ii + jj + kk < P) {
            r[1] += tab[yk_idx(P, ii + 1, jj, kk)] * Mb[t];
            r[2] += tab[yk_idx(P, ii, jj + 1, kk)] * Mb[t];
            r[3] += tab[yk_idx(P, ii, jj, kk + 1)] * Mb[t];
          }
        }
      }
      std::vector<double4> body(n), g4(n, make_double4(-9, -9, -9, -9));
      for (int i = 0; i < n; ++i) body[i] = make_double4(pan[i].c[0], pan[i].c[1], pan[i].c[2], 1.0);
      emu::launch(dim3(nblocks(3, 4)), dim3(128), [&] {
        yk_m2p_kernel<10>(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(), center.data(), body.data(),
                      P, kappa, M.data(), g4.data());
      });
      std::vector<double> got4(4 * n);
      for (int i = 0; i < n; ++i) { got4[4 * i] = g4[i].x; got4[4 * i + 1] = g4[i].y; got4[4 * i + 2] = g4[i].z; got4[4 * i + 3] = g4[i].w; }
      worst_pt = std::max(worst_pt, rel_diff(got4, w4));
    }
    for (int set = 0; set < 2; ++set) {
      std::vector<double> got(n, 0.125), want(n, 0.125);
      emu::launch(dim3(nblocks(3, 4)), dim3(128), [&] {
        if (set == 0) yk_bem_m2p_kernel<0, 10>(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(),
                                           center.data(), pan.data(), bc.data(), P, kappa, M.data(), got.data());
        else yk_bem_m2p_kernel<1, 10>(leaves.data(), 3, bb.data(), be.data(), parent.data(), off.data(), src.data(), center.data(),
                                  pan.data(), bc.data(), P, kappa, M.data(), got.data());
      });
      for (int i = 0; i < n; ++i) if (bc[i] == set) want[i] += set == 0 ? phi[i] : -phi[i];
      for (int i = 0; i < n; ++i) { got[i] -= 0.125; want[i] -= 0.125; biggest = std::max(biggest, std::fabs(want[i])); }
      worst = std::max(worst, rel_diff(got, want));
    }
  }
  printf("ykm2p: %.3e max_potential %.3e point %.3e\n", worst, biggest, worst_pt);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "ykm2p")) return run_ykm2p();
  if (argc >= 2 && !strcmp(argv[1], "bem_rules")) return run_bem_rules();
  if (argc >= 2 && !strcmp(argv[1], "stokes_m2p")) return run_stokes_m2p();
  if (argc >= 2 && !strcmp(argv[1], "m2p")) return run_m2p();
  if (argc >= 3 && !strcmp(argv[1], "near")) return run_near(argv[2]);
  if (argc >= 3 && !strcmp(argv[1], "direct")) return run_direct(argv[2]);
  if (argc >= 3 && !strcmp(argv[1], "pipeline")) return run_pipeline(argv[2]);
  if (argc >= 3 && !strcmp(argv[1], "gmres")) return run_pipeline(argv[2], true);
  if (argc >= 2 && !strcmp(argv[1], "far")) return run_far();
  fprintf(stderr, "usage: emu_stokes_bem near <file> | direct <file> | pipeline <file> | gmres <file> | far | m2p | ykm2p | stokes_m2p | bem_rules\n");
  return 2;
}
