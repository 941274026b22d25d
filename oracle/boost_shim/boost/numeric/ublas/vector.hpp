// Minimal stand-in for <boost/numeric/ublas/vector.hpp>.
//
// TEST INFRASTRUCTURE ONLY. Boost is not installed in this image; the reference
// (barbagroup/fmm-bem-relaxed) only uses a small subset of uBLAS: a dense vector on a
// pluggable storage array, lazy element-wise expressions, and a few reductions
// (reference include/Vec.hpp:11-13,18-20,175-177,340-444). This header provides exactly that
// subset, written from the uBLAS public interface, so the UNMODIFIED reference headers
// compile into oracle/_ref/. Nothing under fmm_bem_relaxed_b200/ includes this file.
//
// Arithmetic conventions that parity depends on:
//   inner_prod / norm_2: plain left-to-right accumulation starting from T(0).
#pragma once
#include <cstddef>
#include <cmath>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <iostream>
#include <iterator>
#include <deque>
#include <string>
#include <memory>
#include <algorithm>
#include <type_traits>
#include <complex>

#ifndef BOOST_UBLAS_INLINE
#define BOOST_UBLAS_INLINE inline
#endif
#ifndef BOOST_UBLAS_CHECK
#define BOOST_UBLAS_CHECK(cond, exc)
#endif

namespace boost {
template <bool B, class T = void> struct enable_if_c { typedef T type; };
template <class T> struct enable_if_c<false, T> {};
template <class Cond, class T = void> struct enable_if : enable_if_c<Cond::value, T> {};
template <class From, class To> struct is_convertible : std::is_convertible<From, To> {};

namespace numeric { namespace ublas {

struct bad_size {};
struct bad_index {};

template <class E> class storage_array {};

// Heap storage used by ublas::vector<T> when no storage is named.
template <class T, class ALLOC = std::allocator<T> >
class unbounded_array : public storage_array<unbounded_array<T, ALLOC> > {
 public:
  typedef std::size_t size_type;
  typedef std::ptrdiff_t difference_type;
  typedef T value_type;
  typedef const T& const_reference;
  typedef T& reference;
  typedef const T* const_iterator;
  typedef T* iterator;
  unbounded_array() : n_(0), d_(nullptr) {}
  explicit unbounded_array(size_type n) : n_(n), d_(n ? new T[n]() : nullptr) {}
  unbounded_array(size_type n, const T& v) : n_(n), d_(n ? new T[n] : nullptr) {
    std::fill(d_, d_ + n_, v);
  }
  unbounded_array(const unbounded_array& o) : n_(o.n_), d_(o.n_ ? new T[o.n_] : nullptr) {
    std::copy(o.d_, o.d_ + n_, d_);
  }
  ~unbounded_array() { delete[] d_; }
  unbounded_array& operator=(const unbounded_array& o) {
    if (this != &o) {
      if (n_ != o.n_) { delete[] d_; n_ = o.n_; d_ = n_ ? new T[n_] : nullptr; }
      std::copy(o.d_, o.d_ + n_, d_);
    }
    return *this;
  }
  void resize(size_type n) {
    if (n == n_) return;
    T* d = n ? new T[n]() : nullptr;
    std::copy(d_, d_ + std::min(n, n_), d);
    delete[] d_; d_ = d; n_ = n;
  }
  size_type size() const { return n_; }
  const_reference operator[](size_type i) const { return d_[i]; }
  reference operator[](size_type i) { return d_[i]; }
  const_iterator begin() const { return d_; }
  const_iterator end() const { return d_ + n_; }
  iterator begin() { return d_; }
  iterator end() { return d_ + n_; }
 private:
  size_type n_;
  T* d_;
};

// CRTP roots
template <class E> class vector_expression {
 public:
  typedef E expression_type;
  const E& operator()() const { return *static_cast<const E*>(this); }
  E& operator()() { return *static_cast<E*>(this); }
};
template <class C> class vector_container : public vector_expression<C> {
 public:
  typedef C container_type;
  const C& operator()() const { return *static_cast<const C*>(this); }
  C& operator()() { return *static_cast<C*>(this); }
};

// Result-type promotion for mixed scalar arithmetic
template <class A, class B> struct promote_traits {
  typedef decltype(std::declval<A>() + std::declval<B>()) promote_type;
};

template <class T1, class T2> struct scalar_plus {
  typedef typename promote_traits<T1, T2>::promote_type result_type;
  static result_type apply(const T1& a, const T2& b) { return a + b; }
};
template <class T1, class T2> struct scalar_minus {
  typedef typename promote_traits<T1, T2>::promote_type result_type;
  static result_type apply(const T1& a, const T2& b) { return a - b; }
};
template <class T1, class T2> struct scalar_multiplies {
  typedef typename promote_traits<T1, T2>::promote_type result_type;
  static result_type apply(const T1& a, const T2& b) { return a * b; }
};
template <class T1, class T2> struct scalar_divides {
  typedef typename promote_traits<T1, T2>::promote_type result_type;
  static result_type apply(const T1& a, const T2& b) { return a / b; }
};
template <class T> struct scalar_negate {
  typedef T result_type;
  static result_type apply(const T& a) { return -a; }
};

// Lazy nodes: sub-expressions by const&, scalars by value (uBLAS closure semantics)
template <class E, class F>
class vector_unary : public vector_expression<vector_unary<E, F> > {
 public:
  typedef typename F::result_type value_type;
  typedef std::size_t size_type;
  explicit vector_unary(const E& e) : e_(e) {}
  size_type size() const { return e_.size(); }
  value_type operator()(size_type i) const { return F::apply(e_(i)); }
  value_type operator[](size_type i) const { return F::apply(e_(i)); }
 private:
  const E& e_;
};
template <class E1, class E2, class F>
class vector_binary : public vector_expression<vector_binary<E1, E2, F> > {
 public:
  typedef typename F::result_type value_type;
  typedef std::size_t size_type;
  vector_binary(const E1& a, const E2& b) : a_(a), b_(b) {}
  size_type size() const { return a_.size(); }
  value_type operator()(size_type i) const { return F::apply(a_(i), b_(i)); }
  value_type operator[](size_type i) const { return F::apply(a_(i), b_(i)); }
 private:
  const E1& a_;
  const E2& b_;
};
template <class T1, class E2, class F>
class vector_binary_scalar1 : public vector_expression<vector_binary_scalar1<T1, E2, F> > {
 public:
  typedef typename F::result_type value_type;
  typedef std::size_t size_type;
  vector_binary_scalar1(const T1& s, const E2& e) : s_(s), e_(e) {}
  size_type size() const { return e_.size(); }
  value_type operator()(size_type i) const { return F::apply(s_, e_(i)); }
  value_type operator[](size_type i) const { return F::apply(s_, e_(i)); }
 private:
  typename std::remove_const<T1>::type s_;
  const E2& e_;
};
template <class E1, class T2, class F>
class vector_binary_scalar2 : public vector_expression<vector_binary_scalar2<E1, T2, F> > {
 public:
  typedef typename F::result_type value_type;
  typedef std::size_t size_type;
  vector_binary_scalar2(const E1& e, const T2& s) : e_(e), s_(s) {}
  size_type size() const { return e_.size(); }
  value_type operator()(size_type i) const { return F::apply(e_(i), s_); }
  value_type operator[](size_type i) const { return F::apply(e_(i), s_); }
 private:
  const E1& e_;
  typename std::remove_const<T2>::type s_;
};

// The traits only NAME the node type, so overload SFINAE never instantiates a functor.
template <class E, class F> struct vector_unary_traits {
  typedef vector_unary<E, F> expression_type;
  typedef expression_type result_type;
};
template <class E1, class E2, class F> struct vector_binary_traits {
  typedef vector_binary<E1, E2, F> expression_type;
  typedef expression_type result_type;
};
template <class T1, class E2, class F> struct vector_binary_scalar1_traits {
  typedef vector_binary_scalar1<T1, E2, F> expression_type;
  typedef expression_type result_type;
};
template <class E1, class T2, class F> struct vector_binary_scalar2_traits {
  typedef vector_binary_scalar2<E1, T2, F> expression_type;
  typedef expression_type result_type;
};

// Dense vector on storage A
template <class T, class A = unbounded_array<T> >
class vector : public vector_container<vector<T, A> > {
 public:
  typedef T value_type;
  typedef A array_type;
  typedef std::size_t size_type;
  typedef std::ptrdiff_t difference_type;
  typedef T& reference;
  typedef const T& const_reference;
  typedef typename A::iterator iterator;
  typedef typename A::const_iterator const_iterator;

  vector() : data_() {}
  explicit vector(size_type n) : data_(n) {}
  vector(size_type n, const value_type& v) : data_(n, v) {}
  vector(const vector& v) : data_(v.data_) {}
  template <class A2>
  vector(const vector<T, A2>& v) : data_(v.size()) {
    for (size_type i = 0; i < v.size(); ++i) data_[i] = v[i];
  }
  template <class AE>
  vector(const vector_expression<AE>& ae) : data_(ae().size()) {
    for (size_type i = 0; i < data_.size(); ++i) data_[i] = ae()(i);
  }

  size_type size() const { return data_.size(); }
  const array_type& data() const { return data_; }
  array_type& data() { return data_; }
  void resize(size_type n, bool = true) { data_.resize(n); }

  const_reference operator()(size_type i) const { return data_[i]; }
  reference operator()(size_type i) { return data_[i]; }
  const_reference operator[](size_type i) const { return data_[i]; }
  reference operator[](size_type i) { return data_[i]; }

  const_iterator begin() const { return data_.begin(); }
  const_iterator end() const { return data_.end(); }
  iterator begin() { return data_.begin(); }
  iterator end() { return data_.end(); }

  vector& operator=(const vector& v) { data_ = v.data_; return *this; }
  template <class A2>
  vector& operator=(const vector<T, A2>& v) {
    for (size_type i = 0; i < size(); ++i) data_[i] = v[i];
    return *this;
  }
  template <class C>
  vector& operator=(const vector_container<C>& v) {
    for (size_type i = 0; i < size(); ++i) data_[i] = v()(i);
    return *this;
  }
  template <class AE>
  vector& operator=(const vector_expression<AE>& ae) {
    // evaluate into a temporary first: the expression may alias *this
    vector tmp(ae);
    data_ = tmp.data_;
    return *this;
  }
  template <class AE>
  vector& operator+=(const vector_expression<AE>& ae) {
    vector tmp(ae);
    for (size_type i = 0; i < size(); ++i) data_[i] += tmp.data_[i];
    return *this;
  }
  template <class AE>
  vector& operator-=(const vector_expression<AE>& ae) {
    vector tmp(ae);
    for (size_type i = 0; i < size(); ++i) data_[i] -= tmp.data_[i];
    return *this;
  }
  template <class S>
  typename std::enable_if<std::is_convertible<S, T>::value, vector&>::type
  operator*=(const S& s) {
    for (size_type i = 0; i < size(); ++i) data_[i] *= s;
    return *this;
  }
  template <class S>
  typename std::enable_if<std::is_convertible<S, T>::value, vector&>::type
  operator/=(const S& s) {
    for (size_type i = 0; i < size(); ++i) data_[i] /= s;
    return *this;
  }
 private:
  array_type data_;
};

// Native uBLAS operators
template <class E1, class E2>
typename vector_binary_traits<E1, E2,
    scalar_plus<typename E1::value_type, typename E2::value_type> >::result_type
operator+(const vector_expression<E1>& a, const vector_expression<E2>& b) {
  return typename vector_binary_traits<E1, E2,
      scalar_plus<typename E1::value_type, typename E2::value_type> >::expression_type(a(), b());
}
template <class E1, class E2>
typename vector_binary_traits<E1, E2,
    scalar_minus<typename E1::value_type, typename E2::value_type> >::result_type
operator-(const vector_expression<E1>& a, const vector_expression<E2>& b) {
  return typename vector_binary_traits<E1, E2,
      scalar_minus<typename E1::value_type, typename E2::value_type> >::expression_type(a(), b());
}
template <class E>
typename vector_unary_traits<E, scalar_negate<typename E::value_type> >::result_type
operator-(const vector_expression<E>& a) {
  return typename vector_unary_traits<E,
      scalar_negate<typename E::value_type> >::expression_type(a());
}
template <class T1, class E2>
typename boost::enable_if<boost::is_convertible<T1, typename E2::value_type>,
    typename vector_binary_scalar1_traits<const T1, E2,
        scalar_multiplies<T1, typename E2::value_type> >::result_type>::type
operator*(const T1& s, const vector_expression<E2>& e) {
  return typename vector_binary_scalar1_traits<const T1, E2,
      scalar_multiplies<T1, typename E2::value_type> >::expression_type(s, e());
}
template <class E1, class T2>
typename boost::enable_if<boost::is_convertible<T2, typename E1::value_type>,
    typename vector_binary_scalar2_traits<E1, const T2,
        scalar_multiplies<typename E1::value_type, T2> >::result_type>::type
operator*(const vector_expression<E1>& e, const T2& s) {
  return typename vector_binary_scalar2_traits<E1, const T2,
      scalar_multiplies<typename E1::value_type, T2> >::expression_type(e(), s);
}
template <class E1, class T2>
typename boost::enable_if<boost::is_convertible<T2, typename E1::value_type>,
    typename vector_binary_scalar2_traits<E1, const T2,
        scalar_divides<typename E1::value_type, T2> >::result_type>::type
operator/(const vector_expression<E1>& e, const T2& s) {
  return typename vector_binary_scalar2_traits<E1, const T2,
      scalar_divides<typename E1::value_type, T2> >::expression_type(e(), s);
}

// Reductions: left-to-right from zero
template <class E1, class E2>
typename promote_traits<typename E1::value_type, typename E2::value_type>::promote_type
inner_prod(const vector_expression<E1>& a, const vector_expression<E2>& b) {
  typedef typename promote_traits<typename E1::value_type,
                                  typename E2::value_type>::promote_type R;
  R t = R();
  const std::size_t n = a().size();
  for (std::size_t i = 0; i < n; ++i) t += a()(i) * b()(i);
  return t;
}
template <class E>
typename E::value_type norm_1(const vector_expression<E>& a) {
  typedef typename E::value_type R;
  R t = R();
  for (std::size_t i = 0; i < a().size(); ++i) t += std::abs(a()(i));
  return t;
}
template <class E>
typename E::value_type norm_2(const vector_expression<E>& a) {
  typedef typename E::value_type R;
  R t = R();
  for (std::size_t i = 0; i < a().size(); ++i) { R u = std::abs(a()(i)); t += u * u; }
  return std::sqrt(t);
}
template <class E>
typename E::value_type norm_inf(const vector_expression<E>& a) {
  typedef typename E::value_type R;
  R t = R();
  for (std::size_t i = 0; i < a().size(); ++i) { R u = std::abs(a()(i)); if (u > t) t = u; }
  return t;
}

}}}  // namespace boost::numeric::ublas
