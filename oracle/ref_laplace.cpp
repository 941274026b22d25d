// oracle/ref_laplace.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference (barbagroup/fmm-bem-relaxed, headers read from
// $(REF)=/root/reference at build time, Boost replaced by oracle/boost_shim) exactly the
// way the reference's own tests/scaling.cpp:14-74 does: LaplaceSpherical K(P); FMM_plan;
// plan.execute(charges); optional Direct::matvec check.  On top of that it dumps the
// reference's internal state (permutation, box table, interaction lists, expansions) so
// that oracle/port and the CUDA engine can be compared array by array.
//
// Private members of the reference's plan/executor/evaluator are read by compiling this one
// file with g++ -fno-access-control; the reference sources are not edited or copied.
//
// Build: see oracle/Makefile (output: oracle/_ref/ref_laplace).
// Run  : ref_laplace -N 10000 -P 5 -ncrit 64 -theta 0.5 [-reps 3] [-direct M] [-in file]
//                    [-dump prefix] [-threads T]
//   inputs : glibc drand48() default seed, N points (x,y,z) then N charges
//            (reference tests/scaling.cpp:29-38), or a raw file of 4N doubles
//            (3N coordinates point-major, then N charges).
//   stdout : one "REF_JSON {...}" line; the reference's own "P2P: .. M2L .." lines pass through.
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include <deque>
#include <set>
#include <unordered_set>
#include <list>
#include <functional>
#include <complex>
#include <string>
#include <iostream>
#include <iomanip>
#include <algorithm>
#include <utility>
#include <type_traits>
#include <iterator>
#include <memory>
#include <sys/time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <boost/numeric/ublas/vector.hpp>
#include <boost/numeric/ublas/matrix_sparse.hpp>
#include <boost/iterator/iterator_adaptor.hpp>
#include <boost/iterator/transform_iterator.hpp>
using std::isnan;

#include <FMM_plan.hpp>
#include <LaplaceSpherical.hpp>

typedef LaplaceSpherical kernel_type;
typedef kernel_type::point_type point_type;
typedef kernel_type::charge_type charge_type;
typedef kernel_type::result_type result_type;
typedef FMM_plan<kernel_type> plan_type;
typedef plan_type::executor_type executor_type;
typedef EvalInteractionLazy<executor_type, true> lazy_type;

template <typename T>
static void dump(const std::string& path, const std::vector<T>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
  fclose(f);
}

int main(int argc, char** argv) {
  int N = 10000, P = 5, reps = 1, ndirect = 0, threads = 0, lazy = 1, treecode = 0;
  unsigned ncrit = 64;
  double theta = 0.5;
  std::string dump_prefix, in_file;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-N")) N = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-P")) P = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-ncrit")) ncrit = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-theta")) theta = atof(argv[++i]);
    else if (!strcmp(argv[i], "-reps")) reps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-direct")) ndirect = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-threads")) threads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-lazy")) lazy = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-tree")) treecode = 1;
    else if (!strcmp(argv[i], "-dump")) dump_prefix = argv[++i];
    else if (!strcmp(argv[i], "-in")) in_file = argv[++i];
    else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
  }
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
  threads = omp_get_max_threads();
#else
  threads = 1;
#endif

  std::vector<point_type> points(N);
  std::vector<charge_type> charges(N);
  if (in_file.empty()) {
    for (int k = 0; k < N; ++k) {
      // same expression as reference tests/scaling.cpp:32 -- with g++ the three calls are
      // evaluated right to left, i.e. the FIRST draw of each triple lands in z.
      points[k] = point_type(drand48(), drand48(), drand48());
    }
    for (int k = 0; k < N; ++k) charges[k] = drand48();
  } else {
    FILE* f = fopen(in_file.c_str(), "rb");
    if (!f) { perror(in_file.c_str()); return 2; }
    std::vector<double> buf(4 * (size_t)N);
    if (fread(buf.data(), sizeof(double), buf.size(), f) != buf.size()) {
      fprintf(stderr, "short read on %s\n", in_file.c_str()); return 2;
    }
    fclose(f);
    for (int k = 0; k < N; ++k) {
      points[k] = point_type(buf[3 * k], buf[3 * k + 1], buf[3 * k + 2]);
      charges[k] = buf[3 * (size_t)N + k];
    }
  }

  kernel_type K(P);
  FMMOptions opts;
  opts.set_mac_theta(theta);
  opts.set_max_per_box(ncrit);
  opts.lazy_evaluation = lazy != 0;
  if (treecode) opts.evaluator = FMMOptions::TREECODE;   // -eval TREE (FMMOptions.hpp:86-92)

  double t0 = get_time();
  plan_type plan(K, points, opts);
  double t_plan = get_time() - t0;

  std::vector<result_type> result(N);
  std::vector<double> times(reps);
  for (int r = 0; r < reps; ++r) {
    double tic = get_time();
    result = plan.execute(charges);
    times[r] = get_time() - tic;
  }
  double best = *std::min_element(times.begin(), times.end());
  double mean = std::accumulate(times.begin(), times.end(), 0.0) / reps;

  // checksums as defined in SURVEY.md section 8(c)
  double pot = 0, fxw = 0;
  for (int k = 0; k < N; ++k) { pot += result[k][0]; fxw += result[k][1] * (k % 7 + 1); }

  // accuracy vs Direct on the first ndirect targets (reference include/Direct.hpp:99-125)
  double err_pot = -1, err_force = -1;
  if (ndirect > 0) {
    ndirect = std::min(ndirect, N);
    std::vector<point_type> tgt(points.begin(), points.begin() + ndirect);
    std::vector<result_type> exact(ndirect);
    Direct::matvec(K, points, charges, tgt, exact);
    double e1 = 0, e2 = 0, f1 = 0, f2 = 0;
    for (int k = 0; k < ndirect; ++k) {
      e1 += (result[k][0] - exact[k][0]) * (result[k][0] - exact[k][0]);
      e2 += exact[k][0] * exact[k][0];
      for (int m = 1; m < 4; ++m) {
        f1 += (result[k][m] - exact[k][m]) * (result[k][m] - exact[k][m]);
        f2 += exact[k][m] * exact[k][m];
      }
    }
    err_pot = sqrt(e1 / e2);
    err_force = sqrt(f1 / f2);
  }

  auto& ex = *plan.executor_;
  auto& tree = ex.source_tree_;
  unsigned nboxes = tree.boxes();
  size_t n_lr = 0, n_p2p = 0, n_m2m = 0, n_l2l = 0, n_p2m = 0, n_l2p = 0;
  lazy_type* ev = nullptr;
  if (lazy && !ex.evals_.evals_.empty())
    ev = dynamic_cast<lazy_type*>(ex.evals_.evals_[0]);
  if (ev) {
    n_lr = ev->LR_list.size();
    for (auto& l : ev->P2P_lists) n_p2p += l.size();
    n_m2m = ev->M2M_list.size(); n_l2l = ev->L2L_list.size();
    n_p2m = ev->P2M_list.size(); n_l2p = ev->L2P_list.size();
  }

  printf("REF_JSON {\"N\": %d, \"P\": %d, \"ncrit\": %u, \"theta\": %.17g, \"threads\": %d, "
         "\"plan_s\": %.6f, \"best_s\": %.6f, \"mean_s\": %.6f, \"reps\": %d, "
         "\"pot\": %.17g, \"fxw\": %.17g, \"r0\": [%.17g, %.17g, %.17g, %.17g], "
         "\"err_pot\": %.6e, \"err_force\": %.6e, \"boxes\": %u, \"levels\": %u, "
         "\"lr_pairs\": %zu, \"p2p_pairs\": %zu, \"m2m\": %zu, \"l2l\": %zu, \"p2m\": %zu, \"l2p\": %zu}\n",
         N, P, ncrit, theta, threads, t_plan, best, mean, reps, pot, fxw,
         result[0][0], result[0][1], result[0][2], result[0][3], err_pot, err_force,
         nboxes, tree.levels(), n_lr, n_p2p, n_m2m, n_l2l, n_p2m, n_l2p);

  if (!dump_prefix.empty()) {
    std::vector<double> res(4 * (size_t)N);
    for (int k = 0; k < N; ++k) for (int m = 0; m < 4; ++m) res[4 * (size_t)k + m] = result[k][m];
    dump(dump_prefix + ".results.f64", res);
    std::vector<double> in(4 * (size_t)N);
    for (int k = 0; k < N; ++k) {
      for (int m = 0; m < 3; ++m) in[3 * (size_t)k + m] = points[k][m];
      in[3 * (size_t)N + k] = charges[k];
    }
    dump(dump_prefix + ".input.f64", in);
    // permutation + morton codes in tree order
    std::vector<unsigned> perm(N), codes(N);
    for (auto it = tree.body_begin(); it != tree.body_end(); ++it) {
      perm[it->index()] = it->number();
      codes[it->index()] = it->morton_index();
    }
    dump(dump_prefix + ".perm.u32", perm);
    dump(dump_prefix + ".codes.u32", codes);
    // box table: key(with leaf bit), parent, child_begin, child_end (raw), body_begin, body_end, level, is_leaf
    std::vector<unsigned> boxes(8 * (size_t)nboxes);
    std::vector<double> geom(4 * (size_t)nboxes);
    for (unsigned b = 0; b < nboxes; ++b) {
      auto box = tree.box(b);
      auto& d = tree.box_data_[b];
      boxes[8 * b + 0] = d.key_; boxes[8 * b + 1] = d.parent_;
      boxes[8 * b + 2] = d.child_begin_; boxes[8 * b + 3] = d.child_end_;
      boxes[8 * b + 4] = box.body_begin()->index();
      boxes[8 * b + 5] = (--box.body_end())->index() + 1;
      boxes[8 * b + 6] = box.level(); boxes[8 * b + 7] = box.is_leaf();
      point_type c = box.center();
      geom[4 * b + 0] = c[0]; geom[4 * b + 1] = c[1]; geom[4 * b + 2] = c[2];
      geom[4 * b + 3] = box.side_length();
    }
    dump(dump_prefix + ".boxes.u32", boxes);
    dump(dump_prefix + ".geom.f64", geom);
    if (ev) {
      std::vector<int> lr(2 * n_lr);
      for (size_t i = 0; i < n_lr; ++i) { lr[2 * i] = ev->LR_list[i].first; lr[2 * i + 1] = ev->LR_list[i].second; }
      dump(dump_prefix + ".lr.i32", lr);
      std::vector<int> off(nboxes + 1, 0), idx;
      for (unsigned b = 0; b < nboxes; ++b) {
        off[b] = idx.size();
        idx.insert(idx.end(), ev->P2P_lists[b].begin(), ev->P2P_lists[b].end());
      }
      off[nboxes] = idx.size();
      dump(dump_prefix + ".p2p_off.i32", off);
      dump(dump_prefix + ".p2p_idx.i32", idx);
      std::vector<int> m2m(2 * n_m2m), l2l(2 * n_l2l);
      for (size_t i = 0; i < n_m2m; ++i) { m2m[2 * i] = ev->M2M_list[i].first; m2m[2 * i + 1] = ev->M2M_list[i].second; }
      for (size_t i = 0; i < n_l2l; ++i) { l2l[2 * i] = ev->L2L_list[i].first; l2l[2 * i + 1] = ev->L2L_list[i].second; }
      dump(dump_prefix + ".m2m.i32", m2m);
      dump(dump_prefix + ".l2l.i32", l2l);
      dump(dump_prefix + ".p2m.i32", ev->P2M_list);
      dump(dump_prefix + ".l2p.i32", ev->L2P_list);
    }
    // expansions after the last execute: nc complex per box (zeros where never initialised)
    int nc = P * (P + 1) / 2;
    std::vector<double> Mx(2 * (size_t)nc * nboxes, 0.0), Lx(2 * (size_t)nc * nboxes, 0.0);
    for (unsigned b = 0; b < nboxes; ++b) {
      auto& M = ex.M_[b].M;
      for (size_t i = 0; i < M.size() && i < (size_t)nc; ++i) {
        Mx[2 * ((size_t)b * nc + i)] = M[i].real(); Mx[2 * ((size_t)b * nc + i) + 1] = M[i].imag();
      }
      if (b < ex.L_.size()) {
        auto& L = ex.L_[b];
        for (size_t i = 0; i < L.size() && i < (size_t)nc; ++i) {
          Lx[2 * ((size_t)b * nc + i)] = L[i].real(); Lx[2 * ((size_t)b * nc + i) + 1] = L[i].imag();
        }
      }
    }
    dump(dump_prefix + ".M.f64", Mx);
    dump(dump_prefix + ".L.f64", Lx);
  }
  return 0;
}
