// Minimal stand-in for <opencv2/core.hpp> so that the reference's header-only leaves
// (openvslam/match_base.h, match_angle_checker.h, trigonometric.h) compile from where they lie.
// Only cvRound / cvFloor are needed; both restate opencv2/core/fast_math.hpp (SSE2 path:
// cvtss2si / cvtsd2si, i.e. round-half-to-even).
#pragma once
#include <cmath>
#include <cstdint>
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(int v) { return v; }
