// ORACLE (test infrastructure only): intensity-centroid orientation and rotated-BRIEF descriptor.
// Follows orb_extractor.cpp:174-186 (u_max_), :245-275 (ic_angle), :284-352 (compute_orb_descriptor,
// non-SSE macro branch :326-331 -- USE_SSE_ORB is not defined by CMakeLists.txt), and
// openvslam/trigonometric.h:17-46.  cv::fastAtan2 (OpenCV core, mathfuncs_core: atan_f32 scalar
// path) is restated and pinned against cv2.fastAtan2 4.13.0.
#include "common.h"
#include <cfloat>

namespace orc {

static const int8_t PATTERN[1024] = {
#include "orb_pattern.inc"
};

struct UMax {
    int v[HALF_PATCH + 1];
    UMax() {  // orb_extractor.cpp:174-186
        const unsigned vmax = (unsigned)std::floor(HALF_PATCH * std::sqrt(2.0) / 2 + 1);
        const unsigned vmin = (unsigned)std::ceil(HALF_PATCH * std::sqrt(2.0) / 2);
        for (unsigned i = 0; i <= vmax; ++i)
            v[i] = (int)std::round(std::sqrt((double)HALF_PATCH * HALF_PATCH - (double)i * i));
        for (unsigned i = HALF_PATCH, v0 = 0; vmin <= i; --i) {
            while (v[v0] == v[v0 + 1]) ++v0;
            v[i] = (int)v0;
            ++v0;
        }
    }
};
static const UMax UMAX;

static inline float fast_atan2_deg(float y, float x) {
    static const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    static const float p1 = 0.9997878412794807f * scale;
    static const float p3 = -0.3258083974640975f * scale;
    static const float p5 = 0.1555786518463281f * scale;
    static const float p7 = -0.04432655554792128f * scale;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

static inline void ic_moments(const uint8_t *img, int stride, int x, int y, int &m_10, int &m_01) {
    m_01 = 0; m_10 = 0;
    const uint8_t *center = img + (size_t)y * stride + x;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
        unsigned v_sum = 0;  // unsigned in the reference; wraps exactly like int32
        const int d = UMAX.v[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * stride];
            const int val_minus = center[u - v * stride];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
}

// openvslam/trigonometric.h
static constexpr float PI_ = 3.14159265358979f;
static constexpr float PI_2_ = PI_ / 2.0f;
static constexpr float TWO_PI_ = 2.0f * PI_;
static constexpr float INV_TWO_PI_ = 1.0f / TWO_PI_;
static constexpr float THREE_PI_2_ = 3.0f * PI_2_;
static inline float poly_cos(float v) {
    constexpr float c1 = 0.99940307f, c2 = -0.49558072f, c3 = 0.03679168f;
    const float v2 = v * v;
    return c1 + v2 * (c2 + c3 * v2);
}
static inline float util_cos(float v) {
    v = v - cv_floor(v * INV_TWO_PI_) * TWO_PI_;
    v = (0.0f < v) ? v : -v;
    if (v < PI_2_) return poly_cos(v);
    if (v < PI_) return -poly_cos(PI_ - v);
    if (v < THREE_PI_2_) return -poly_cos(v - PI_);
    return poly_cos(TWO_PI_ - v);
}
static inline float util_sin(float v) { return util_cos(PI_2_ - v); }

static void descriptor(const uint8_t *img, int stride, int x, int y, float angle_deg, uint32_t *out) {
    uint8_t *desc = reinterpret_cast<uint8_t *>(out);
    const float angle = angle_deg * M_PI / 180.0;  // double math, narrowed (orb_extractor.cpp:286)
    const float cos_angle = util_cos(angle);
    const float sin_angle = util_sin(angle);
    const uint8_t *center = img + (size_t)y * stride + x;
    auto value = [&](int idx) -> int {
        const float px = PATTERN[idx], py = PATTERN[idx + 1];
        return center[cv_round(px * sin_angle + py * cos_angle) * stride
                      + cv_round(px * cos_angle - py * sin_angle)];
    };
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const int p = (8 * i + k) * 4;
            val |= (value(p) < value(p + 2)) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

}  // namespace orc

using namespace orc;

extern "C" float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
extern "C" void orc_ic_moments(const uint8_t *img, int stride, int x, int y, int *m10, int *m01) {
    ic_moments(img, stride, x, y, *m10, *m01);
}
extern "C" float orc_ic_angle(const uint8_t *img, int stride, int x, int y) {
    int m10, m01;
    ic_moments(img, stride, x, y, m10, m01);
    return fast_atan2_deg((float)m01, (float)m10);
}
extern "C" float orc_util_cos(float v) { return util_cos(v); }
extern "C" float orc_util_sin(float v) { return util_sin(v); }
extern "C" void orc_descriptor(const uint8_t *blurred, int stride, int x, int y, float angle_deg, uint32_t *desc8) {
    descriptor(blurred, stride, x, y, angle_deg, desc8);
}
extern "C" void orc_umax(int *umax16) { for (int i = 0; i <= HALF_PATCH; ++i) umax16[i] = UMAX.v[i]; }
extern "C" void orc_pattern(float *out1024) { for (int i = 0; i < 1024; ++i) out1024[i] = PATTERN[i]; }
