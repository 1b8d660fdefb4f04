#pragma once
#include <ostream>
#include <cstdint>
#include <stdint.h>
namespace boost { namespace units {
template<typename Unit, typename T = double> struct quantity {
  T v{};
  quantity() = default;
  explicit quantity(T x) : v(x) {}
  T value() const { return v; }
  quantity operator+(const quantity& o) const { return quantity(v + o.v); }
  quantity& operator+=(const quantity& o) { v += o.v; return *this; }
  quantity operator-(const quantity& o) const { return quantity(v - o.v); }
  quantity operator*(T s) const { return quantity(v * s); }
  bool operator<(const quantity& o) const { return v < o.v; }
  bool operator>(const quantity& o) const { return v > o.v; }
  bool operator<=(const quantity& o) const { return v <= o.v; }
  bool operator>=(const quantity& o) const { return v >= o.v; }
  bool operator==(const quantity& o) const { return v == o.v; }
  friend std::ostream& operator<<(std::ostream& o, const quantity& q) { return o << q.v; }
};
template<typename Unit, typename T> quantity<Unit, T> operator*(T s, const quantity<Unit, T>& q) { return quantity<Unit, T>(q.v * s); }
} }
