// Minimal stand-in for <boost/numeric/ublas/matrix_sparse.hpp> (TEST INFRASTRUCTURE ONLY).
// The reference fills a compressed_matrix row by row with in-order push_back and then reads
// the three CSR arrays directly (reference include/executor/EvalP2P.hpp:82-94,
// include/Matvec.hpp:14-33). Only that interface exists here.
#pragma once
#include <vector>
#include <cstddef>
#include "vector.hpp"

namespace boost { namespace numeric { namespace ublas {

template <class T>
class compressed_matrix {
 public:
  typedef T value_type;
  typedef std::size_t size_type;
  compressed_matrix() : rows_(0), cols_(0), filled_row_(0) { offsets_.assign(1, 0); }
  compressed_matrix(size_type rows, size_type cols, size_type nnz = 0)
      : rows_(rows), cols_(cols), filled_row_(0) {
    offsets_.assign(rows + 1, 0);
    indices_.reserve(nnz);
    values_.reserve(nnz);
  }
  // Elements must arrive in row-major order (as the reference guarantees).
  void push_back(size_type i, size_type j, const T& v) {
    while (filled_row_ < i) { ++filled_row_; offsets_[filled_row_] = indices_.size(); }
    indices_.push_back(j);
    values_.push_back(v);
  }
  size_type size1() const { return rows_; }
  size_type size2() const { return cols_; }
  size_type nnz() const { return indices_.size(); }
  const std::vector<size_type>& index1_data() const { fix(); return offsets_; }
  const std::vector<size_type>& index2_data() const { return indices_; }
  const std::vector<T>& value_data() const { return values_; }
 private:
  // offsets of the row being filled and of every later (empty) row end at nnz
  void fix() const {
    for (size_type r = filled_row_ + 1; r <= rows_; ++r) offsets_[r] = indices_.size();
  }
  size_type rows_, cols_;
  mutable size_type filled_row_;
  mutable std::vector<size_type> offsets_;
  std::vector<size_type> indices_;
  std::vector<T> values_;
};

}}}
