#pragma once
#include <sstream>
#include <string>
namespace util {
template <class M> std::string eigenToString(const M &m) {
    std::stringstream ss;
    for (int r = 0; r < m.rows(); ++r) for (int c = 0; c < m.cols(); ++c) ss << m(r, c) << (c + 1 < m.cols() ? "," : ";");
    return ss.str();
}
}  // namespace util
