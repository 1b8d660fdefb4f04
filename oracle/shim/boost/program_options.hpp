#pragma once
#include <any>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost {
using any = std::any;
namespace program_options {
struct validation_error : std::runtime_error {
  enum kind_t { invalid_option_value = 1 };
  validation_error(kind_t, const std::string& v = "") : std::runtime_error("invalid option value " + v) {}
};
namespace validators {
template<typename T> inline void check_first_occurrence(const T&) {}
inline const std::string& get_single_string(const std::vector<std::string>& v) { static const std::string e; return v.empty() ? e : v[0]; }
}
} }
