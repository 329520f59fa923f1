// stand-in for ../tracker/track.hpp (absent parent-project header): the VIO tracker's feature record
#pragma once
#include <array>
namespace tracker {
struct Feature {
    struct Point { float x, y; };
    int id = -1;
    std::array<Point, 2> points{};   // points[0]: first camera (orb_extractor.cpp:90)
    float depth = -1;                // keyframe.cpp:56
};
}  // namespace tracker
