// stand-in for core/io/PointsPersistence.h:11-74: same persist_points template, one in-memory sink
#pragma once
#include "datastructures/PointBuffer.h"
#include "math/AABB.h"
#include <map>
#include <string>
#include <vector>
struct PointsPersistence
{
  template<typename Iter>
  void persist_points(Iter points_begin, Iter points_end, const AABB& bounds, const std::string& node_name)
  {
    auto& v = nodes[node_name];
    v.clear();
    for (; points_begin != points_end; ++points_begin)
      v.push_back((*points_begin).position());
    (void)bounds;
  }
  bool node_exists(const std::string& node_name) const { return nodes.count(node_name) != 0; }
  bool is_lossless() const { return true; }
  std::map<std::string, std::vector<Vector3<double>>> nodes;
};
