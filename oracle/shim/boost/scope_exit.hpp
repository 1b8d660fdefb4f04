/* oracle/shim/boost/scope_exit.hpp — TEST INFRASTRUCTURE ONLY.  BOOST_SCOPE_EXIT_TPL / _END as used by
 * core/io/LASPersistence.h (run a block when the enclosing scope ends), built on a lambda guard. */
#pragma once
#include <utility>
namespace shim_scope_exit {
template<typename F>
struct Guard
{
  F f;
  explicit Guard(F&& fn)
    : f(std::move(fn))
  {}
  Guard(Guard&& o)
    : f(std::move(o.f))
  {}
  ~Guard() { f(); }
};
template<typename F>
Guard<F>
make(F&& f)
{
  return Guard<F>(std::forward<F>(f));
}
} // namespace shim_scope_exit
#define SHIM_SCOPE_EXIT_CAT2(a, b) a##b
#define SHIM_SCOPE_EXIT_CAT(a, b) SHIM_SCOPE_EXIT_CAT2(a, b)
#define BOOST_SCOPE_EXIT_TPL(...) auto SHIM_SCOPE_EXIT_CAT(shim_scope_exit_guard_, __LINE__) = ::shim_scope_exit::make([&]()
#define BOOST_SCOPE_EXIT(...) BOOST_SCOPE_EXIT_TPL(__VA_ARGS__)
#define BOOST_SCOPE_EXIT_END );
