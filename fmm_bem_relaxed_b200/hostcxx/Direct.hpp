#pragma once
/** @file Direct.hpp
 * Direct::matvec overloads the reference's drivers call for accuracy checks
 * (reference include/Direct.hpp:232-303): brute-force sums with the kernel's operator().
 * This is the host-side CHECK, kept on the CPU like in the reference; the FMM near field itself runs
 * on the GPU inside FMM_plan::execute.
 */
#include <cassert>
#include <vector>

class Direct {
 public:
  /** r_i += sum_j K(t_i, s_j) c_j */
  template <typename Kernel, typename SourceIter, typename ChargeIter, typename TargetIter, typename ResultIter>
  inline static void matvec(const Kernel& K, SourceIter s_first, SourceIter s_last, ChargeIter c_first,
                            TargetIter t_first, TargetIter t_last, ResultIter r_first) {
    for (; t_first != t_last; ++t_first, ++r_first) {
      SourceIter s = s_first;
      ChargeIter c = c_first;
      for (; s != s_last; ++s, ++c) *r_first += K(*t_first, *s) * (*c);
    }
  }
  template <typename Kernel>
  inline static void matvec(const Kernel& K, const std::vector<typename Kernel::source_type>& s,
                            const std::vector<typename Kernel::charge_type>& c,
                            const std::vector<typename Kernel::target_type>& t,
                            std::vector<typename Kernel::result_type>& r) {
    assert(s.size() == c.size());
    assert(t.size() == r.size());
    const long nt = (long)t.size();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < nt; ++i)
      for (size_t j = 0; j < s.size(); ++j) r[i] += K(t[i], s[j]) * c[j];
  }
  /** sources == targets */
  template <typename Kernel>
  inline static void matvec(const Kernel& K, const std::vector<typename Kernel::source_type>& p,
                            const std::vector<typename Kernel::charge_type>& c,
                            std::vector<typename Kernel::result_type>& r) {
    matvec(K, p, c, p, r);
  }
};
