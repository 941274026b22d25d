#pragma once
/** @file FMMOptions.hpp
 * Same class, fields, setters and command-line parser as the reference
 * (reference include/FMMOptions.hpp:9-106).  Only the options the GPU plan honours change
 * behaviour; the rest are kept so existing drivers compile and run unchanged.
 */
#include "Vec.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

class FMMOptions {
 public:
  bool lazy_evaluation;   // accepted; the GPU plan always replays precomputed lists
  bool local_evaluation;  // near-field-only plan (preconditioners): honoured
  bool sparse_local;      // accepted; the BEM near field is always cached on the GPU
  bool block_diagonal;    // leaf-with-itself blocks only (preconditioners): honoured

  enum EvalType { FMM, TREECODE };
  EvalType evaluator;

  struct DefaultMAC {
    double theta_;
    DefaultMAC(double theta) : theta_(theta) {}
    template <typename BOX>
    bool operator()(const BOX& b1, const BOX& b2) const {
      double r0_normSq = normSq(b1.center() - b2.center());
      double rhs = (b1.radius() + b2.radius()) / theta_;
      return r0_normSq > rhs * rhs;
    }
  };

  DefaultMAC MAC_;
  unsigned NCRIT_;
  bool printTree;
  int device;  // extension: CUDA device ordinal, -1 = current
  bool cold_plan;  // extension: FMMB_FLAG_COLD_PLAN -- no warm start at construction (include/fmmb.h)

  FMMOptions()
      : lazy_evaluation(true), local_evaluation(false), sparse_local(false), block_diagonal(false),
        evaluator(FMM), MAC_(DefaultMAC(0.5)), NCRIT_(64), printTree(false), device(-1), cold_plan(false) {}

  void set_mac_theta(double theta) { MAC_ = DefaultMAC(theta); }
  DefaultMAC MAC() { return MAC_; }
  void set_max_per_box(unsigned ncrit) { NCRIT_ = ncrit; }
  unsigned max_per_box() const { return NCRIT_; }
  void print_tree(bool v) { printTree = v; }
  bool print_tree() const { return printTree; }
};

/** Get the FMMOptions from command line arguments (same flags as the reference) */
inline FMMOptions get_options(int argc, char** argv) {
  FMMOptions opts = FMMOptions();
  for (int i = 1; i < argc; ++i) {
    if (strcmp(argv[i], "-theta") == 0) {
      i++;
      opts.set_mac_theta((double)atof(argv[i]));
    } else if (strcmp(argv[i], "-eval") == 0) {
      i++;
      if (strcmp(argv[i], "FMM") == 0) opts.evaluator = FMMOptions::FMM;
      else if (strcmp(argv[i], "TREE") == 0) opts.evaluator = FMMOptions::TREECODE;
      else printf("[W]: Unknown evaluator type: \"%s\"\n", argv[i]);
    } else if (strcmp(argv[i], "-lazy_eval") == 0) {
      opts.lazy_evaluation = true;
    } else if (strcmp(argv[i], "-ncrit") == 0) {
      i++;
      opts.set_max_per_box((unsigned)atoi(argv[i]));
    } else if (strcmp(argv[i], "-printtree") == 0) {
      opts.print_tree(true);
    }
  }
  return opts;
}
