// Host-side static tables of the context: pyramid geometry, per-level keypoint budgets, detection
// cell grids and the fixed-point linear-resize taps.  All of it is tiny scalar arithmetic that must
// reproduce the x86 float/double results of the reference exactly, so it is done once on the host
// at sg_create and uploaded; the kernels only consume integers.
//
// Follows static_settings.cpp:9-15 (float scale products), :39-60 (budget split),
// image_pyramid.cpp:76-78 (std::round of width / double(scale)) and the cv::resize INTER_LINEAR
// coefficient generation of OpenCV imgproc (un-vendored dependency of the reference).
#include <algorithm>
#include <cmath>
#include "ctx.h"

namespace sg {

static inline int round_half_even(float v) { return (int)lrintf(v); }       // cvRound
static inline int floor_int(float v) { int i = (int)v; return i - (i > v); }  // cvFloor
static inline int16_t to_short(int v) { return (int16_t)std::min(32767, std::max(-32768, v)); }

void make_geometry(const sg_params &p, std::vector<Level> &lv) {
    lv.assign(p.levels, Level{});
    float s = 1.0f;
    for (int l = 0; l < p.levels; ++l) {
        if (l > 0) s = p.scale_factor * s;   // float product, as calc_scale_factors does
        lv[l].scale = s;
        const double ds = s;
        lv[l].w = l == 0 ? p.width : (int)std::round(p.width * 1.0 / ds);
        lv[l].h = l == 0 ? p.height : (int)std::round(p.height * 1.0 / ds);
    }
    // budget per level: geometric series, remainder to the last level
    double desired = p.max_keypoints * (1.0 - 1.0 / p.scale_factor)
                     / (1.0 - std::pow(1.0 / p.scale_factor, (double)p.levels));
    long total = 0;
    for (int l = 0; l + 1 < p.levels; ++l) {
        lv[l].budget = (int)std::round(desired);
        total += lv[l].budget;
        desired *= 1.0 / p.scale_factor;
    }
    lv[p.levels - 1].budget = (int)std::max<long>(p.max_keypoints - total, 0);

    for (int l = 0; l < p.levels; ++l) {
        Level &L = lv[l];
        L.pitch = (L.w + 127) & ~127;                       // 128-B rows: aligned vector stores, TMA strides
        L.frame_stride = (size_t)L.pitch * L.h;
        L.area_w = L.w - 2 * PATCH_RADIUS;
        L.area_h = L.h - 2 * PATCH_RADIUS;
        const int ew = L.w - 2 * EVAL_ORIGIN, eh = L.h - 2 * EVAL_ORIGIN;  // evaluated interior
        L.cells_x = ew > 0 ? (ew + CELL - 1) / CELL : 0;
        L.cells_y = eh > 0 ? (eh + CELL - 1) / CELL : 0;
        if (L.cells_x == 0 || L.cells_y == 0) L.cells_x = L.cells_y = 0;
        if (L.area_w > 0 && L.area_h > 0) {
            const double ratio = (double)L.area_w / L.area_h;
            if (ratio > 1) { L.init_nx = (int)std::round(ratio); L.init_ny = 1; }
            else           { L.init_nx = 1; L.init_ny = (int)std::round(1 / ratio); }
        }
        // a whole round runs only while nodes + 3*pool <= budget, the partial round stops at the first
        // division reaching the budget (<= budget + 2); the unconditional first round makes <= 4*initial
        L.node_cap = std::max(L.budget + 3, 4 * L.init_nx * L.init_ny) + 1;
    }
}

bool is_area2x(int sw, int sh, int dw, int dh) {
    const double sx = 1. / ((double)dw / sw), sy = 1. / ((double)dh / sh);
    const int ix = (int)lrint(sx), iy = (int)lrint(sy);
    return std::abs(sx - ix) < 2.220446049250313e-16 && std::abs(sy - iy) < 2.220446049250313e-16
           && ix == 2 && iy == 2;
}

void make_resize_taps(int src, int dst, std::vector<ResizeTap> &taps, bool horizontal) {
    taps.resize(dst);
    const double scale = 1. / ((double)dst / src);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = floor_int(f);
        f -= s;
        ResizeTap t;
        if (horizontal) {
            // the x direction zeroes the fraction at the borders
            if (s < 0) { f = 0; s = 0; }
            if (s >= src - 1) { f = 0; s = src - 1; }
            t.s0 = s;
            t.s1 = std::min(s + 1, src - 1);
        } else {
            // the y direction keeps the fraction and clips the row indices instead
            t.s0 = std::min(std::max(s, 0), src - 1);
            t.s1 = std::min(std::max(s + 1, 0), src - 1);
        }
        t.a0 = to_short(round_half_even((1.f - f) * 2048));
        t.a1 = to_short(round_half_even(f * 2048));
        taps[d] = t;
    }
}

}  // namespace sg
