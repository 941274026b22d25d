#pragma once
/** @file FMM_plan.hpp
 * FMM_plan<Kernel> with the reference's surface (reference include/FMM_plan.hpp:15-128):
 *   FMM_plan(const Kernel&, const std::vector<source_type>&, FMMOptions&)
 *   std::vector<result_type> execute(const std::vector<charge_type>&)
 *   kernel_type& kernel();   FMMOptions& options();
 * Everything below it -- octree, interaction lists, P2M/M2M/M2L/L2L/L2P/P2P -- runs on the GPU
 * through the C ABI of include/fmmb.h (libfmmb200.so).  The plan copies the kernel and the options
 * like the reference does; kernel().set_p(p) is picked up by the next execute.
 * Differences on purpose: move-only (the reference's implicit copy would double free);
 * no per-matvec printf; errors are reported on stderr and leave an empty result like the
 * reference's "[E]" path (:79-82).
 */
#include <cstdio>
#include <iostream>
#include <vector>
#include <cstdlib>

#include "Vec.hpp"
#include "Mat3.hpp"       // the reference's FMM_plan.hpp brings it in through its executors (include/Matvec.hpp:11)
#include "FMMOptions.hpp"
#include "Direct.hpp"
#include <timing.hpp>   // <>: a build that puts the reference's examples/BEM first gets that one, not both
#include "../../include/fmmb.h"

template <class Kernel>
class FMM_plan {
 public:
  typedef Kernel kernel_type;
  typedef typename kernel_type::point_type point_type;
  typedef typename kernel_type::source_type source_type;
  typedef typename kernel_type::target_type target_type;
  typedef typename kernel_type::charge_type charge_type;
  typedef typename kernel_type::result_type result_type;

  FMM_plan(const kernel_type& k, const std::vector<source_type>& source, FMMOptions& opts)
      : plan_(nullptr), K(k), opts_(opts), n_(source.size()), sources_(source) {
    // sources -> plain arrays: positions (panel centres for BEM kernels), panel vertices, boundary conditions
    std::vector<double> pts, verts;
    std::vector<int32_t> bc;
    Kernel::pack_sources(source, pts, verts, bc);
    fmmb_kernel_desc kd = {Kernel::fmmb_kind, K.order(), K.kappa(), K.quad_k(), K.quad_kfine()};
    fmmb_sources src = {(int64_t)n_, pts.data(), verts.empty() ? nullptr : verts.data(),
                        bc.empty() ? nullptr : bc.data()};
    fmmb_options fo = {};
    fo.theta = opts_.MAC().theta_;
    fo.ncrit = opts_.max_per_box();
    fo.evaluator = opts_.evaluator == FMMOptions::FMM ? FMMB_EVAL_FMM : FMMB_EVAL_TREECODE;
    fo.device = opts_.device;
    fo.near_only = opts_.block_diagonal ? 2 : (opts_.local_evaluation ? 1 : 0);   // plans for preconditioners
    fo.kernel_flags = K.kernel_flags() | (opts_.cold_plan ? FMMB_FLAG_COLD_PLAN : 0);
    if (fmmb_plan_create(&kd, &src, &fo, &plan_) != FMMB_OK) {
      std::cerr << "[E]: FMM_plan: " << fmmb_last_error() << "\n";
      plan_ = nullptr;
    }
  }
  FMM_plan(const FMM_plan&) = delete;
  FMM_plan& operator=(const FMM_plan&) = delete;
  FMM_plan(FMM_plan&& o) : plan_(o.plan_), K(o.K), opts_(o.opts_), n_(o.n_), sources_(std::move(o.sources_)) {
    o.plan_ = nullptr;
  }
  ~FMM_plan() { fmmb_plan_destroy(plan_); }

  kernel_type& kernel() { return K; }
  const kernel_type& kernel() const { return K; }
  FMMOptions& options() { return opts_; }

  /** results = A * charges, original body order */
  std::vector<result_type> execute(const std::vector<charge_type>& charges) {
    if (!plan_) {
      printf("[E]: Executor not initialised -- returning..\n");
      return std::vector<result_type>(0);
    }
    static_assert(sizeof(result_type) == Kernel::result_dim * sizeof(double), "result_type must be packed doubles");
    static_assert(sizeof(charge_type) == Kernel::charge_dim * sizeof(double), "charge_type must be packed doubles");
    std::vector<result_type> results(charges.size());
    if (charges.size() != n_ || fmmb_plan_set_p(plan_, K.order()) != FMMB_OK ||
        fmmb_plan_execute(plan_, reinterpret_cast<const double*>(charges.data()),
                          reinterpret_cast<double*>(results.data())) != FMMB_OK) {
      std::cerr << "[E]: FMM_plan::execute: "
                << (charges.size() != n_ ? "charges.size() != sources.size()" : fmmb_last_error()) << "\n";
      return std::vector<result_type>(0);
    }
    // the reference prints its two dominant phases after every execute (include/executor/EvalInteractionLazy.hpp:152);
    // here that line is opt-in (FMMB_PRINT_PHASES=1) and carries device times -- both phases when the matvec ran
    // as plain launches, the whole matvec when it was one CUDA-graph replay
    static const bool print_phases = std::getenv("FMMB_PRINT_PHASES") != nullptr;
    if (print_phases) {
      double ms[FMMB_T_COUNT] = {0};
      fmmb_plan_info info;
      if (fmmb_plan_phase_times(plan_, ms, FMMB_T_COUNT) == FMMB_OK && fmmb_plan_get_info(plan_, &info) == FMMB_OK) {
        if (ms[FMMB_T_P2P] > 0 || ms[FMMB_T_M2L] > 0)
          printf("P2P: %.4gs, M2L (%d): %.4gs\n", ms[FMMB_T_P2P] * 1e-3, (int)info.n_m2l_pairs, ms[FMMB_T_M2L] * 1e-3);
        else
          printf("matvec (CUDA graph): %.4gs, M2L pairs %d\n", ms[FMMB_T_TOTAL] * 1e-3, (int)info.n_m2l_pairs);
      }
    }
    return results;
  }

  /** The plan's copies of the sources in TREE order (reference include/FMM_plan.hpp:100-107; used by
   * Preconditioners::Diagonal, examples/LaplaceBEM.cpp:241-244) */
  typedef typename std::vector<source_type>::iterator body_source_iterator;
  body_source_iterator source_begin() { tree_order(); return tree_sources_.begin(); }
  body_source_iterator source_end() { tree_order(); return tree_sources_.end(); }

  /** extension: the C handle, for fmmb_plan_get_info / phase times */
  fmmb_plan* handle() { return plan_; }

 private:
  void tree_order() {
    if (!tree_sources_.empty() || !plan_) return;
    std::vector<uint32_t> perm(n_);
    if (fmmb_plan_get_tree(plan_, perm.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) != FMMB_OK) return;
    tree_sources_.reserve(n_);
    for (size_t i = 0; i < n_; ++i) tree_sources_.push_back(sources_[perm[i]]);
  }
  fmmb_plan* plan_;
  kernel_type K;
  FMMOptions opts_;
  size_t n_;
  std::vector<source_type> sources_, tree_sources_;
};
