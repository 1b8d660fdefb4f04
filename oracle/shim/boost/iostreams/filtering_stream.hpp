#pragma once
#include <istream>
#include <ostream>
#include <sstream>
namespace boost { namespace iostreams {
struct filtering_ostream : std::ostringstream {};
struct filtering_istream : std::istringstream {};
} }
