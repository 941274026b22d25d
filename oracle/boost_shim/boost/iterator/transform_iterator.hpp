// Minimal stand-in for <boost/iterator/transform_iterator.hpp> (TEST INFRASTRUCTURE ONLY).
// reference = F::result_type, so a functor that returns a real reference (reference
// include/executor/ExecutorSingleTree.hpp:81-90) lets writes go through the iterator.
#pragma once
#include <iterator>
#include <cstddef>
#include <type_traits>

namespace boost {

template <class F, class It>
class transform_iterator {
 public:
  typedef typename F::result_type reference;
  typedef typename std::remove_cv<typename std::remove_reference<reference>::type>::type value_type;
  typedef std::ptrdiff_t difference_type;
  typedef typename std::add_pointer<typename std::remove_reference<reference>::type>::type pointer;
  typedef std::random_access_iterator_tag iterator_category;

  transform_iterator() : it_(), f_() {}
  transform_iterator(const It& it, const F& f) : it_(it), f_(f) {}
  const It& base() const { return it_; }
  reference operator*() const { return f_(*it_); }
  pointer operator->() const { return &f_(*it_); }
  reference operator[](difference_type n) const { return f_(*(it_ + n)); }
  transform_iterator& operator++() { ++it_; return *this; }
  transform_iterator operator++(int) { transform_iterator t(*this); ++it_; return t; }
  transform_iterator& operator--() { --it_; return *this; }
  transform_iterator operator--(int) { transform_iterator t(*this); --it_; return t; }
  transform_iterator& operator+=(difference_type n) { it_ += n; return *this; }
  transform_iterator& operator-=(difference_type n) { it_ -= n; return *this; }
  friend transform_iterator operator+(const transform_iterator& a, difference_type n) {
    transform_iterator t(a); t += n; return t;
  }
  friend transform_iterator operator-(const transform_iterator& a, difference_type n) {
    transform_iterator t(a); t -= n; return t;
  }
  friend difference_type operator-(const transform_iterator& a, const transform_iterator& b) {
    return a.it_ - b.it_;
  }
  friend bool operator==(const transform_iterator& a, const transform_iterator& b) { return a.it_ == b.it_; }
  friend bool operator!=(const transform_iterator& a, const transform_iterator& b) { return !(a.it_ == b.it_); }
  friend bool operator<(const transform_iterator& a, const transform_iterator& b) { return a.it_ < b.it_; }
 private:
  It it_;
  F f_;
};

template <class F, class It>
transform_iterator<F, It> make_transform_iterator(It it, F f) {
  return transform_iterator<F, It>(it, f);
}

}  // namespace boost
