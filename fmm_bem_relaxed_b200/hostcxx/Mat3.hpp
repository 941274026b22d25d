#pragma once
/** @file Mat3.hpp
 * 3x3 matrix, row major, with the public surface the reference's drivers and mesh generators use
 * (reference include/Mat3.hpp:7-90): construction from a fill value, element access, negation, sums, products with a
 * Vec<3> and with a scalar (operator* and multiply()).  Own implementation on a plain array.
 */
#include "Vec.hpp"

template <typename T>
class Mat3 {
  T v_[9];

 public:
  Mat3() { for (int i = 0; i < 9; ++i) v_[i] = T(); }
  Mat3(double x) { for (int i = 0; i < 9; ++i) v_[i] = (T)x; }
  template <typename Iter>
  Mat3(Iter first, Iter last) { int i = 0; for (; first != last && i < 9; ++first, ++i) v_[i] = *first; for (; i < 9; ++i) v_[i] = T(); }

  const T& operator()(unsigned i, unsigned j) const { return v_[3 * i + j]; }
  T& operator()(unsigned i, unsigned j) { return v_[3 * i + j]; }

  Mat3 operator-() const { Mat3 r; for (int i = 0; i < 9; ++i) r.v_[i] = -v_[i]; return r; }
  Mat3& operator+=(const Mat3& m) { for (int i = 0; i < 9; ++i) v_[i] += m.v_[i]; return *this; }
  Mat3 operator+(const Mat3& m) const { Mat3 r(*this); return r += m; }

  Vec<3, T> multiply(const Vec<3, T>& x) const {
    return Vec<3, T>(v_[0] * x[0] + v_[1] * x[1] + v_[2] * x[2], v_[3] * x[0] + v_[4] * x[1] + v_[5] * x[2],
                     v_[6] * x[0] + v_[7] * x[1] + v_[8] * x[2]);
  }
  Mat3 multiply(double s) const { Mat3 r; for (int i = 0; i < 9; ++i) r.v_[i] = v_[i] * s; return r; }
  Vec<3, T> operator*(const Vec<3, T>& x) const { return multiply(x); }
  Mat3 operator*(double s) const { return multiply(s); }
};
