// ORACLE (test infrastructure only): pyramid geometry, cv::resize(INTER_LINEAR) and
// cv::GaussianBlur(7x7, sigma 2) restated for 8-bit single-channel images.
// Follows image_pyramid.cpp:68-86 and static_settings.cpp:9-60 of the reference; the OpenCV
// arithmetic (un-vendored dependency, version unpinned by the reference) is restated from the
// published algorithm of OpenCV 4.x imgproc (resize.cpp: HResizeLinear/VResizeLinear fixed point,
// smooth: fixed-point 8.8 bit-exact Gaussian) and pinned against cv2 4.13.0 in tests/.
#include "common.h"

namespace orc {

Geometry make_geometry(const orc_params &p) {
    Geometry g{};
    g.levels = p.levels;
    // static_settings.cpp:9-15: float products
    g.scale[0] = 1.0f;
    for (int l = 1; l < p.levels; ++l) g.scale[l] = p.scale_factor * g.scale[l - 1];
    g.off[0] = 0;
    for (int l = 0; l < p.levels; ++l) {
        // image_pyramid.cpp:76-78: double scale, std::round
        const double scale = g.scale[l];
        g.w[l] = l == 0 ? p.width : (int)std::round(p.width * 1.0 / scale);
        g.h[l] = l == 0 ? p.height : (int)std::round(p.height * 1.0 / scale);
        g.off[l + 1] = g.off[l] + (size_t)g.w[l] * g.h[l];
    }
    // static_settings.cpp:39-60
    double desired = p.max_keypoints * (1.0 - 1.0 / p.scale_factor)
                     / (1.0 - std::pow(1.0 / p.scale_factor, static_cast<double>(p.levels)));
    unsigned total = 0;
    for (int l = 0; l < p.levels - 1; ++l) {
        g.budget[l] = (int)(size_t)std::round(desired);
        total += g.budget[l];
        desired *= 1.0 / p.scale_factor;
    }
    g.budget[p.levels - 1] = std::max(p.max_keypoints - (int)total, 0);
    return g;
}

}  // namespace orc

using namespace orc;

extern "C" void orc_level_geometry(const orc_params *p, float *scales, int *widths, int *heights) {
    Geometry g = make_geometry(*p);
    for (int l = 0; l < p->levels; ++l) {
        if (scales) scales[l] = g.scale[l];
        if (widths) widths[l] = g.w[l];
        if (heights) heights[l] = g.h[l];
    }
}

extern "C" void orc_level_budgets(const orc_params *p, int *budgets) {
    Geometry g = make_geometry(*p);
    for (int l = 0; l < p->levels; ++l) budgets[l] = g.budget[l];
}

extern "C" void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                                     uint8_t *dst, int dw, int dh, int dstride) {
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;

    if (dw == sw && dh == sh) {  // cv::resize copies when sizes agree
        for (int y = 0; y < sh; ++y) memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, sw);
        return;
    }
    // cv::resize silently switches INTER_LINEAR to INTER_AREA for an exact 2x decimation.
    {
        const int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y);  // saturate_cast<int>
        const bool is_area_fast = std::abs(scale_x - iscale_x) < 2.220446049250313e-16
                                  && std::abs(scale_y - iscale_y) < 2.220446049250313e-16;
        if (is_area_fast && iscale_x == 2 && iscale_y == 2) {
            for (int y = 0; y < dh; ++y)
                for (int x = 0; x < dw; ++x) {
                    const uint8_t *s0 = src + (size_t)(2 * y) * sstride + 2 * x;
                    const uint8_t *s1 = s0 + sstride;
                    dst[(size_t)y * dstride + x] = (uint8_t)((s0[0] + s0[1] + s1[0] + s1[1] + 2) >> 2);
                }
            return;
        }
    }

    std::vector<int> xofs(dw);
    std::vector<short> ialpha(2 * dw);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = sat_short(cv_round((1.f - fx) * 2048));
        ialpha[2 * dx + 1] = sat_short(cv_round(fx * 2048));
    }
    std::vector<int> row0(dw), row1(dw);
    auto hresize = [&](int sy, std::vector<int> &row) {
        const uint8_t *S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx];
            const int s1 = sx + 1 < sw ? S[sx + 1] : 0;  // alpha1 == 0 there
            row[dx] = S[sx] * ialpha[2 * dx] + s1 * ialpha[2 * dx + 1];
        }
    };
    auto clip = [](int v, int lo, int hi) { return v >= lo ? (v < hi ? v : hi - 1) : lo; };
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        const short b0 = sat_short(cv_round((1.f - fy) * 2048));
        const short b1 = sat_short(cv_round(fy * 2048));
        hresize(clip(sy, 0, sh), row0);
        hresize(clip(sy + 1, 0, sh), row1);
        uint8_t *D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx)
            D[dx] = sat_u8((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

extern "C" void orc_gaussian7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
    // 8.8 fixed-point kernel of getGaussianKernel(7, 2) as OpenCV's bit-exact 8-bit path uses it.
    static const int kq[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *S = src + (size_t)y * sstride;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int k = 0; k < 7; ++k) acc += kq[k] * S[reflect101(x + k - 3, w)];
            tmp[(size_t)y * w + x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; ++k) acc += (uint32_t)kq[k] * tmp[(size_t)reflect101(y + k - 3, h) * w + x];
            D[x] = sat_u8((int)((acc + 32768u) >> 16));
        }
    }
}

extern "C" void orc_pyramid(const orc_params *p, const uint8_t *img, int stride, uint8_t *pyr, uint8_t *blur) {
    Geometry g = make_geometry(*p);
    for (int y = 0; y < g.h[0]; ++y) memcpy(pyr + (size_t)y * g.w[0], img + (size_t)y * stride, g.w[0]);
    for (int l = 1; l < g.levels; ++l)  // chained: level l from level l-1 (image_pyramid.cpp:79)
        orc_resize_linear_u8(pyr + g.off[l - 1], g.w[l - 1], g.h[l - 1], g.w[l - 1],
                             pyr + g.off[l], g.w[l], g.h[l], g.w[l]);
    if (blur)
        for (int l = 0; l < g.levels; ++l)
            orc_gaussian7_u8(pyr + g.off[l], g.w[l], g.h[l], g.w[l], blur + g.off[l], g.w[l]);
}
