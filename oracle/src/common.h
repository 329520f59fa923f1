// ORACLE (test infrastructure only) -- shared helpers.  See ../orb_oracle.h.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../orb_oracle.h"

namespace orc {

// cvRound: SSE2 cvtss2si / cvtsd2si == round-half-to-even in the default rounding mode.
static inline int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_round(double v) { return (int)lrint(v); }
// cvFloor (opencv2/core/fast_math.hpp)
static inline int cv_floor(float v) { int i = (int)v; return i - (i > v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }
static inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

constexpr int PATCH_RADIUS = 19;      // StaticSettings::ORB_PATCH_RADIUS   (static_settings.hpp:14)
constexpr int FAST_PATCH_SIZE = 31;   // StaticSettings::ORB_FAST_PATCH_SIZE (static_settings.hpp:15)
constexpr int HALF_PATCH = FAST_PATCH_SIZE / 2;  // 15

struct Geometry {
    int levels;
    float scale[ORC_MAX_LEVELS];
    int w[ORC_MAX_LEVELS], h[ORC_MAX_LEVELS];
    size_t off[ORC_MAX_LEVELS + 1];  // tight-packed plane offsets
    int budget[ORC_MAX_LEVELS];
};
Geometry make_geometry(const orc_params &p);

struct Kp { int x, y, resp; };
std::vector<Kp> cv_fast(const uint8_t *img, int w, int h, int stride, int thr);
std::vector<Kp> detect_level(const uint8_t *img, int w, int h, int stride, int budget,
                             int ini_thr, int min_thr, std::vector<Kp> *cands);
std::vector<int> distribute(const std::vector<Kp> &cands, int area_w, int area_h, int budget);

}  // namespace orc
