#pragma once
/** @file Preconditioner.hpp
 * Identity and diagonal preconditioners (reference examples/BEM/Preconditioner.hpp:8-41). */
#include <vector>

namespace Preconditioners {

class Identity {
 public:
  template <typename VecType>
  void operator()(const VecType& x, VecType& y) const {
    y = x;
  }
};

/** divides by the panel self interactions K(s_i, s_i), in the order the sources are handed in */
template <typename ValueType>
class Diagonal {
 public:
  template <typename Kernel, typename SourceIter>
  Diagonal(Kernel& K, SourceIter first, SourceIter last) {
    for (; first != last; ++first) recip_.push_back(1. / K(*first, *first));
  }
  template <typename VecType>
  void operator()(const VecType& x, VecType& y) const {
    for (size_t i = 0; i < x.size(); ++i) y[i] = recip_[i] * x[i];
  }

 private:
  std::vector<ValueType> recip_;
};

}  // namespace Preconditioners
