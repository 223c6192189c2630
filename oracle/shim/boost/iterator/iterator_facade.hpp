// Minimal stand-in for boost::iterator_facade (Boost is not installed in this image and is
// not vendored by the reference).  TEST INFRASTRUCTURE ONLY: it exists so that the reference's
// own headers (/root/reference/geomc/linalg/...) compile into oracle/_ref.  No hot-path
// arithmetic lives behind this interface; it is iterator plumbing only.
#pragma once
#include <cstddef>
#include <iterator>
#include <memory>
#include <type_traits>

namespace boost {

struct forward_traversal_tag {};
struct bidirectional_traversal_tag {};
struct random_access_traversal_tag {};

class iterator_core_access {
public:
    template <class I> static decltype(auto) dereference(const I& i) { return i.dereference(); }
    template <class I> static void increment(I& i) { i.increment(); }
    template <class I> static void decrement(I& i) { i.decrement(); }
    template <class I, class D> static void advance(I& i, D n) { i.advance(n); }
    template <class I, class J> static bool equal(const I& a, const J& b) { return a.equal(b); }
    template <class I, class J> static auto distance_from(const I& a, const J& b) {
        return a.distance_to(b);
    }
};

namespace facade_detail {
template <class Tag> struct to_std_category { using type = Tag; };
template <> struct to_std_category<forward_traversal_tag> { using type = std::forward_iterator_tag; };
template <> struct to_std_category<bidirectional_traversal_tag> {
    using type = std::bidirectional_iterator_tag;
};
template <> struct to_std_category<random_access_traversal_tag> {
    using type = std::random_access_iterator_tag;
};
}  // namespace facade_detail

template <class Derived, class Value, class Category, class Reference = Value&,
          class Difference = std::ptrdiff_t>
class iterator_facade {
    Derived& self() { return *static_cast<Derived*>(this); }
    const Derived& self() const { return *static_cast<const Derived*>(this); }

public:
    using value_type = std::remove_cv_t<Value>;
    using reference = Reference;
    using pointer = std::add_pointer_t<std::remove_reference_t<Reference>>;
    using difference_type = Difference;
    using iterator_category = typename facade_detail::to_std_category<Category>::type;

    reference operator*() const { return iterator_core_access::dereference(self()); }
    pointer operator->() const { return std::addressof(*self()); }
    reference operator[](difference_type n) const {
        Derived t(self());
        iterator_core_access::advance(t, n);
        return *t;
    }
    Derived& operator++() { iterator_core_access::increment(self()); return self(); }
    Derived operator++(int) { Derived t(self()); ++*this; return t; }
    Derived& operator--() { iterator_core_access::decrement(self()); return self(); }
    Derived operator--(int) { Derived t(self()); --*this; return t; }
    Derived& operator+=(difference_type n) { iterator_core_access::advance(self(), n); return self(); }
    Derived& operator-=(difference_type n) { iterator_core_access::advance(self(), -n); return self(); }
    Derived operator-(difference_type n) const { Derived t(self()); t -= n; return t; }

    friend Derived operator+(const Derived& a, difference_type n) { Derived t(a); t += n; return t; }
    friend Derived operator+(difference_type n, const Derived& a) { Derived t(a); t += n; return t; }
};

// comparison / distance between (possibly different, interoperable) facade-derived iterators
template <class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2>
bool operator==(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    return iterator_core_access::equal(static_cast<const D1&>(a), static_cast<const D2&>(b));
}
template <class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2>
bool operator!=(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    return !(a == b);
}
template <class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2>
auto operator-(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    // a - b == distance from b to a
    return iterator_core_access::distance_from(static_cast<const D2&>(b), static_cast<const D1&>(a));
}
template <class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2>
bool operator<(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    return (a - b) < 0;
}
template <class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2>
bool operator>(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    return (a - b) > 0;
}
template <class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2>
bool operator<=(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    return (a - b) <= 0;
}
template <class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2>
bool operator>=(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    return (a - b) >= 0;
}

}  // namespace boost
