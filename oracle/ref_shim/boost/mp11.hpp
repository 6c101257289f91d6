// Stand-in for the four boost::mp11 facilities the ACTS seeding sources use
// (mp_list, mp_product, mp_at_c, mp_for_each).  TEST INFRASTRUCTURE, see
// Eigen/Core in this directory.  Written from the documented interface.
#pragma once
#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace boost::mp11 {

template <typename... T>
struct mp_list {};

namespace shim {
// concatenation of lists
template <typename... L>
struct concat;
template <>
struct concat<> { using type = mp_list<>; };
template <typename... A>
struct concat<mp_list<A...>> { using type = mp_list<A...>; };
template <typename... A, typename... B, typename... Rest>
struct concat<mp_list<A...>, mp_list<B...>, Rest...> { using type = typename concat<mp_list<A..., B...>, Rest...>::type; };

// cartesian product, first list varying slowest (mp_product's order)
template <template <typename...> class F, typename Prefix, typename... Lists>
struct product;
template <template <typename...> class F, typename... P>
struct product<F, mp_list<P...>> { using type = mp_list<F<P...>>; };
template <template <typename...> class F, typename... P, typename... H, typename... Rest>
struct product<F, mp_list<P...>, mp_list<H...>, Rest...> {
  using type = typename concat<typename product<F, mp_list<P..., H>, Rest...>::type...>::type;
};

template <typename L, std::size_t I>
struct at_c;
template <typename... T, std::size_t I>
struct at_c<mp_list<T...>, I> { using type = std::tuple_element_t<I, std::tuple<T...>>; };
}  // namespace shim

template <template <typename...> class F, typename... Lists>
using mp_product = typename shim::product<F, mp_list<>, Lists...>::type;

template <typename L, std::size_t I>
using mp_at_c = typename shim::at_c<L, I>::type;

namespace shim {
template <typename... T, typename Fn>
constexpr Fn for_each_impl(mp_list<T...>, Fn&& f) {
  (f(T{}), ...);
  return std::forward<Fn>(f);
}
}  // namespace shim

template <typename L, typename Fn>
constexpr Fn mp_for_each(Fn&& f) {
  return shim::for_each_impl(L{}, std::forward<Fn>(f));
}

}  // namespace boost::mp11
