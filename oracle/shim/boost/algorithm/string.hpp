#pragma once
#include <algorithm>
#include <cctype>
#include <string>
#include <string_view>
#include <vector>
namespace boost {
struct is_any_of_pred { std::string chars; bool operator()(char c) const { return chars.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(const std::string& s) { return { s }; }
template<typename Out, typename Pred> inline void split(Out& out, std::string_view in, Pred p) {
  out.clear(); std::string cur;
  for (char c : in) { if (p(c)) { out.push_back(cur); cur.clear(); } else cur.push_back(c); }
  out.push_back(cur);
}
inline std::string to_lower_copy(std::string s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); }); return s; }
inline std::string to_upper_copy(std::string s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::toupper(c); }); return s; }
inline bool iequals(const std::string& a, const std::string& b) { return to_lower_copy(a) == to_lower_copy(b); }
}
namespace boost { namespace algorithm { using boost::split; using boost::is_any_of; using boost::to_lower_copy; } }
