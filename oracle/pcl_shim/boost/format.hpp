// oracle/pcl_shim — minimal stand-in for boost::format (only the LM debug print uses it, lsq_registration_impl.hpp:152-157).
#pragma once
#include <ostream>
#include <sstream>
#include <string>
namespace boost {
class format {
public:
  explicit format(const std::string& f) { s_ << f; }
  format(const format& o) { s_ << o.s_.str(); }
  template <typename T>
  format& operator%(const T& v) {
    s_ << ' ' << v;
    return *this;
  }
  std::string str() const { return s_.str(); }

private:
  std::ostringstream s_;
};
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}  // namespace boost
