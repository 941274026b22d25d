#pragma once
/** @file Vec.hpp
 * Fixed-size value vector with the public surface the reference's drivers use
 * (reference include/Vec.hpp:175-333: Vec<N,T>, element access, arithmetic, norm/normSq/dot,
 * stream output).  Own implementation on a plain array -- no Boost.
 * Quirk kept on purpose: Vec<N,T>(size_type) is the ZERO vector, not a fill.
 */
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <type_traits>

template <std::size_t N, typename T>
class Vec {
  T a_[N];

 public:
  typedef T value_type;
  typedef std::size_t size_type;
  static constexpr size_type dimension = N;

  Vec() { for (size_type i = 0; i < N; ++i) a_[i] = T(); }
  /** One value per coordinate, e.g. Vec<3,double>(x, y, z) */
  template <typename... Arg, typename std::enable_if<sizeof...(Arg) == N && (N > 1), int>::type = 0>
  explicit Vec(Arg... args) : a_{static_cast<T>(args)...} {}
  /** Reference semantics: a single integral argument is a SIZE, the vector is zero */
  template <typename S, typename std::enable_if<N == 1 || std::is_arithmetic<S>::value, int>::type = 0>
  explicit Vec(S) { for (size_type i = 0; i < N; ++i) a_[i] = T(); }

  size_type size() const { return N; }
  T& operator[](size_type i) { return a_[i]; }
  const T& operator[](size_type i) const { return a_[i]; }
  T* begin() { return a_; }
  T* end() { return a_ + N; }
  const T* begin() const { return a_; }
  const T* end() const { return a_ + N; }
  const T* data() const { return a_; }
  T* data() { return a_; }

  Vec& operator+=(const Vec& b) { for (size_type i = 0; i < N; ++i) a_[i] += b.a_[i]; return *this; }
  Vec& operator-=(const Vec& b) { for (size_type i = 0; i < N; ++i) a_[i] -= b.a_[i]; return *this; }
  Vec& operator*=(const T& s) { for (size_type i = 0; i < N; ++i) a_[i] *= s; return *this; }
  Vec& operator/=(const T& s) { for (size_type i = 0; i < N; ++i) a_[i] /= s; return *this; }
  Vec operator-() const { Vec r; for (size_type i = 0; i < N; ++i) r.a_[i] = -a_[i]; return r; }
};

template <std::size_t N, typename T> Vec<N, T> operator+(Vec<N, T> a, const Vec<N, T>& b) { return a += b; }
template <std::size_t N, typename T> Vec<N, T> operator-(Vec<N, T> a, const Vec<N, T>& b) { return a -= b; }
template <std::size_t N, typename T> Vec<N, T> operator*(Vec<N, T> a, const T& s) { return a *= s; }
template <std::size_t N, typename T> Vec<N, T> operator*(const T& s, Vec<N, T> a) { return a *= s; }
template <std::size_t N, typename T> Vec<N, T> operator/(Vec<N, T> a, const T& s) { return a /= s; }
/** scalars of another arithmetic type (v / 4, 2 * v): converted like the built-in promotion would */
template <std::size_t N, typename T, typename S,
          typename std::enable_if<std::is_arithmetic<S>::value && !std::is_same<S, T>::value, int>::type = 0>
Vec<N, T> operator*(Vec<N, T> a, const S& s) { return a *= static_cast<T>(s); }
template <std::size_t N, typename T, typename S,
          typename std::enable_if<std::is_arithmetic<S>::value && !std::is_same<S, T>::value, int>::type = 0>
Vec<N, T> operator*(const S& s, Vec<N, T> a) { return a *= static_cast<T>(s); }
template <std::size_t N, typename T, typename S,
          typename std::enable_if<std::is_arithmetic<S>::value && !std::is_same<S, T>::value, int>::type = 0>
Vec<N, T> operator/(Vec<N, T> a, const S& s) { return a /= static_cast<T>(s); }
template <std::size_t N, typename T> Vec<N, T> operator+(Vec<N, T> a, const T& s) {
  for (std::size_t i = 0; i < N; ++i) a[i] += s;
  return a;
}
template <std::size_t N, typename T> Vec<N, T> operator-(Vec<N, T> a, const T& s) {
  for (std::size_t i = 0; i < N; ++i) a[i] -= s;
  return a;
}
/** Element-wise product / quotient (reference include/Vec.hpp:408-444) */
template <std::size_t N, typename T> Vec<N, T> operator*(Vec<N, T> a, const Vec<N, T>& b) {
  for (std::size_t i = 0; i < N; ++i) a[i] *= b[i];
  return a;
}
template <std::size_t N, typename T> Vec<N, T> operator/(Vec<N, T> a, const Vec<N, T>& b) {
  for (std::size_t i = 0; i < N; ++i) a[i] /= b[i];
  return a;
}
template <std::size_t N, typename T> bool operator==(const Vec<N, T>& a, const Vec<N, T>& b) {
  return std::equal(a.begin(), a.end(), b.begin());
}
template <std::size_t N, typename T> bool operator!=(const Vec<N, T>& a, const Vec<N, T>& b) { return !(a == b); }
template <std::size_t N, typename T> std::ostream& operator<<(std::ostream& s, const Vec<N, T>& v) {
  s << "(";
  for (std::size_t i = 0; i < N; ++i) s << v[i] << (i + 1 < N ? ", " : "");
  return s << ")";
}
/** Left-to-right sums starting from zero, like uBLAS inner_prod / norm_2 */
template <std::size_t N, typename T> T inner_prod(const Vec<N, T>& a, const Vec<N, T>& b) {
  T t = T();
  for (std::size_t i = 0; i < N; ++i) t += a[i] * b[i];
  return t;
}
template <std::size_t N, typename T> T dot(const Vec<N, T>& a, const Vec<N, T>& b) { return inner_prod(a, b); }
template <std::size_t N, typename T> T normSq(const Vec<N, T>& a) { return inner_prod(a, a); }
template <std::size_t N, typename T> T norm(const Vec<N, T>& a) { return std::sqrt(normSq(a)); }
template <std::size_t N, typename T> T norm_2(const Vec<N, T>& a) { return norm(a); }
template <std::size_t N, typename T> T norm_1(const Vec<N, T>& a) {
  T t = T();
  for (std::size_t i = 0; i < N; ++i) t += std::abs(a[i]);
  return t;
}
template <std::size_t N, typename T> T norm_inf(const Vec<N, T>& a) {
  T t = T();
  for (std::size_t i = 0; i < N; ++i) t = std::max(t, (T)std::abs(a[i]));
  return t;
}
