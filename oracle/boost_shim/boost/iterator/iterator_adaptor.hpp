// Minimal stand-in for <boost/iterator/iterator_adaptor.hpp> (TEST INFRASTRUCTURE ONLY).
// The reference's octree iterators adapt an unsigned index and dereference to a Box/Body
// returned BY VALUE (reference include/tree/Octree.hpp:420-475).
#pragma once
#include <iterator>
#include <cstddef>
#include <type_traits>

namespace boost {

struct use_default {};

class iterator_core_access {
 public:
  template <class It>
  static typename It::reference dereference(const It& it) { return it.dereference(); }
};

namespace detail {
template <class T> struct arrow_proxy {
  T value;
  explicit arrow_proxy(const T& v) : value(v) {}
  T* operator->() { return &value; }
};
}

template <class Derived, class Base, class Value, class Category, class Reference,
          class Difference = std::ptrdiff_t>
class iterator_adaptor {
 public:
  typedef Value value_type;
  typedef Reference reference;
  typedef Difference difference_type;
  typedef Category iterator_category;
  typedef typename std::conditional<std::is_reference<Reference>::value,
      typename std::add_pointer<typename std::remove_reference<Reference>::type>::type,
      detail::arrow_proxy<typename std::remove_const<Value>::type> >::type pointer;
  typedef iterator_adaptor iterator_adaptor_;

  iterator_adaptor() : base_() {}
  explicit iterator_adaptor(const Base& b) : base_(b) {}

  const Base& base() const { return base_; }

  reference operator*() const { return iterator_core_access::dereference(derived()); }
  pointer operator->() const { return arrow(std::is_reference<Reference>()); }
  reference operator[](difference_type n) const { Derived t(derived()); t += n; return *t; }

  Derived& operator++() { ++base_; return derived(); }
  Derived operator++(int) { Derived t(derived()); ++base_; return t; }
  Derived& operator--() { --base_; return derived(); }
  Derived operator--(int) { Derived t(derived()); --base_; return t; }
  Derived& operator+=(difference_type n) { base_ += n; return derived(); }
  Derived& operator-=(difference_type n) { base_ -= n; return derived(); }
  friend Derived operator+(const Derived& a, difference_type n) { Derived t(a); t += n; return t; }
  friend Derived operator+(difference_type n, const Derived& a) { Derived t(a); t += n; return t; }
  friend Derived operator-(const Derived& a, difference_type n) { Derived t(a); t -= n; return t; }
  friend difference_type operator-(const Derived& a, const Derived& b) {
    return difference_type(a.base()) - difference_type(b.base());
  }
  // non-template friends: these must win over std::rel_ops (KernelTraits.hpp:8)
  friend bool operator==(const Derived& a, const Derived& b) { return a.base() == b.base(); }
  friend bool operator!=(const Derived& a, const Derived& b) { return !(a.base() == b.base()); }
  friend bool operator<(const Derived& a, const Derived& b) { return a.base() < b.base(); }
  friend bool operator>(const Derived& a, const Derived& b) { return b.base() < a.base(); }
  friend bool operator<=(const Derived& a, const Derived& b) { return !(b.base() < a.base()); }
  friend bool operator>=(const Derived& a, const Derived& b) { return !(a.base() < b.base()); }

 protected:
  const Base& base_reference() const { return base_; }
  Base& base_reference() { return base_; }

 private:
  Derived& derived() { return *static_cast<Derived*>(this); }
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  pointer arrow(std::true_type) const { return &**this; }
  pointer arrow(std::false_type) const { return pointer(**this); }
  Base base_;
};

}  // namespace boost
