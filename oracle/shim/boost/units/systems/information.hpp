#pragma once
#include "../quantity.hpp"
#include <cstddef>
namespace boost { namespace units { namespace information {
namespace hu { namespace byte { struct info {}; } }
struct byte_unit {};
static const byte_unit byte{};
static const byte_unit bytes{};
using byte_quantity = quantity<hu::byte::info>;
inline byte_quantity operator*(std::size_t v, byte_unit) { return byte_quantity((double)v); }
inline byte_quantity operator*(int v, byte_unit) { return byte_quantity((double)v); }
inline byte_quantity operator*(double v, byte_unit) { return byte_quantity(v); }
} }
namespace units {
inline information::byte_quantity operator*(const information::byte_quantity& q, std::size_t s) { return information::byte_quantity(q.v * (double)s); }
inline information::byte_quantity operator*(std::size_t s, const information::byte_quantity& q) { return information::byte_quantity(q.v * (double)s); }
inline information::byte_quantity operator*(int s, const information::byte_quantity& q) { return information::byte_quantity(q.v * (double)s); }
} }
