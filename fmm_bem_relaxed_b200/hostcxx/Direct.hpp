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
  // kernels that define a vector P2P (the stresslet) are evaluated through it, the others through operator(),
  // like the reference's trait dispatch (reference include/Direct.hpp:30-97)
  template <typename Kernel, typename SI, typename CI, typename TI, typename RI>
  static auto eval(const Kernel& K, SI s0, SI s1, CI c0, TI t0, TI t1, RI r0, int)
      -> decltype(K.P2P(s0, s1, c0, t0, t1, r0), void()) {
    K.P2P(s0, s1, c0, t0, t1, r0);
  }
  template <typename Kernel, typename SI, typename CI, typename TI, typename RI>
  static void eval(const Kernel& K, SI s0, SI s1, CI c0, TI t0, TI t1, RI r0, long) {
    for (; t0 != t1; ++t0, ++r0) {
      SI s = s0;
      CI c = c0;
      for (; s != s1; ++s, ++c) *r0 += K(*t0, *s) * (*c);
    }
  }

 public:
  /** r_i += sum_j K(t_i, s_j) c_j */
  template <typename Kernel, typename SourceIter, typename ChargeIter, typename TargetIter, typename ResultIter>
  inline static void matvec(const Kernel& K, SourceIter s_first, SourceIter s_last, ChargeIter c_first,
                            TargetIter t_first, TargetIter t_last, ResultIter r_first) {
    eval(K, s_first, s_last, c_first, t_first, t_last, r_first, 0);
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
