#pragma once
#include <string>
namespace util {
inline std::string removeFileSuffix(const std::string &path) {
    const auto dot = path.find_last_of('.');
    const auto slash = path.find_last_of('/');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return path;
    return path.substr(0, dot);
}
}  // namespace util
