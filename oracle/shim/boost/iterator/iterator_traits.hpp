// Stand-in for <boost/iterator/iterator_traits.hpp>; see iterator_facade.hpp. Test infrastructure only.
#pragma once
#include <iterator>
namespace boost {
template <class I> struct iterator_value      { using type = typename std::iterator_traits<I>::value_type; };
template <class I> struct iterator_reference  { using type = typename std::iterator_traits<I>::reference; };
template <class I> struct iterator_pointer    { using type = typename std::iterator_traits<I>::pointer; };
template <class I> struct iterator_difference { using type = typename std::iterator_traits<I>::difference_type; };
template <class I> struct iterator_category   { using type = typename std::iterator_traits<I>::iterator_category; };
}  // namespace boost
