// shim: boost::format used by the reference only to build exception / log messages
#pragma once
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
public:
  explicit format(const std::string& f) : _fmt(f) {}
  template<typename T> format& operator%(const T& v) { std::ostringstream ss; ss << v; _args.push_back(ss.str()); return *this; }
  std::string str() const {
    std::string out;
    for (size_t i = 0; i < _fmt.size(); ++i) {
      if (_fmt[i] == '%' && i + 1 < _fmt.size()) {
        size_t j = i + 1; size_t idx = 0; bool digits = false;
        while (j < _fmt.size() && _fmt[j] >= '0' && _fmt[j] <= '9') { idx = idx * 10 + (_fmt[j] - '0'); ++j; digits = true; }
        if (digits && j < _fmt.size() && _fmt[j] == '%') { if (idx >= 1 && idx <= _args.size()) out += _args[idx - 1]; i = j; continue; }
      }
      out += _fmt[i];
    }
    return out;
  }
  friend std::ostream& operator<<(std::ostream& o, const format& f) { return o << f.str(); }
private:
  std::string _fmt; std::vector<std::string> _args;
};
inline std::string str(const format& f) { return f.str(); }
}
