// boost/iterator_adaptors.hpp -- stand-in: the part of boost::iterator_facade that bvh::Tree's node iterator uses
// (forward traversal through increment / equal / dereference of the derived class).
#pragma once
#include <cstddef>
#include <iterator>
namespace boost {
struct forward_traversal_tag {};
class iterator_core_access {
 public:
  template <typename It> static void increment(It& it) { it.increment(); }
  template <typename It> static bool equal(const It& a, const It& b) { return a.equal(b); }
  template <typename It> static typename It::reference dereference(const It& it) { return it.dereference(); }
};
template <typename Derived, typename Value, typename Traversal>
class iterator_facade {
 public:
  typedef Value value_type;
  typedef Value& reference;
  typedef Value* pointer;
  typedef std::ptrdiff_t difference_type;
  typedef std::forward_iterator_tag iterator_category;
  reference operator*() const { return iterator_core_access::dereference(derived()); }
  pointer operator->() const { return &iterator_core_access::dereference(derived()); }
  Derived& operator++() { iterator_core_access::increment(derived()); return derived(); }
  Derived operator++(int) { Derived t(derived()); iterator_core_access::increment(derived()); return t; }
  friend bool operator==(const Derived& a, const Derived& b) { return iterator_core_access::equal(a, b); }
  friend bool operator!=(const Derived& a, const Derived& b) { return !iterator_core_access::equal(a, b); }
 private:
  Derived& derived() { return *static_cast<Derived*>(this); }
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
};
}  // namespace boost
