// stand-in for taskflow v3.7.0 (schwarzwald/util/CMakeLists.txt:3-7): only what the adapter touches
#pragma once
#include <functional>
#include <utility>
#include <vector>
namespace tf {
struct Task
{
  Task& name(const char*) { return *this; }
  template<typename... T> Task& precede(T&&...) { return *this; }
};
struct Subflow;
struct Taskflow
{
  template<typename F> Task emplace(F&& f)
  {
    work.emplace_back(std::forward<F>(f));
    return Task{};
  }
  std::vector<std::function<void()>> work;
};
}
