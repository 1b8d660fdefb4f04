#pragma once
#include <map>
namespace boost { namespace container { template<typename K, typename V> using flat_map = std::map<K, V>; } }
