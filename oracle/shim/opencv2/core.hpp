// Minimal stand-in for <opencv2/core.hpp> (TEST INFRASTRUCTURE ONLY) so that reference sources compile from
// where they lie (oracle/Makefile target `ref`).  OpenCV's C++ headers are not in this image.
//   cvRound / cvFloor restate opencv2/core/fast_math.hpp (SSE2 path: cvtss2si / cvtsd2si, round-half-to-even).
//   cv::Mat is a reference-counted 8-bit matrix with just the members the compiled reference files touch.
//   cv::resize / cv::GaussianBlur / cv::fastAtan2 are DECLARED here and defined in oracle/ref_slam.cpp on top of
//   the oracle's restatements, which are pinned bit-for-bit against cv2 4.13 (tests/golden/golden_cv2.npz).
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(int v) { return v; }

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0

namespace cv {
struct Point { int x = 0, y = 0; Point() {} Point(int x, int y) : x(x), y(y) {} };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x, float y) : x(x), y(y) {} };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x, int y, int w, int h) : x(x), y(y), width(w), height(h) {}
    Rect(Point p, Size s) : x(p.x), y(p.y), width(s.width), height(s.height) {}
    bool contains(const Point2f &p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
template <class T, int N> struct Vec {
    T v[N];
    Vec() : v{} {}
    Vec(T a, T b, T c) : v{a, b, c} {}
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
};
using Vec3b = Vec<uchar, 3>;
using Vec4b = Vec<uchar, 4>;
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };

enum { INTER_LINEAR = 1, BORDER_REFLECT_101 = 4 };

struct MatStep {
    size_t v = 0;
    operator size_t() const { return v; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    MatStep step;
    Mat() {}
    Mat(int r, int c, int /*type*/) { create(r, c); }
    Mat(const Mat &m, const Rect &roi) : rows(roi.height), cols(roi.width), data(m.data + roi.y * m.step.v + roi.x), step(m.step), buf(m.buf) {}
    // external data, not owned
    Mat(int r, int c, int /*type*/, void *ptr, size_t stride) : rows(r), cols(c), data((uchar *)ptr) { step.v = stride; }
    void create(int r, int c) {
        if (r == rows && c == cols && buf && step.v == (size_t)c) return;
        rows = r; cols = c; step.v = (size_t)c;
        buf = std::shared_ptr<uchar>(new uchar[(size_t)r * c + 16], std::default_delete<uchar[]>());
        data = buf.get();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int channels() const { return 1; }
    Size size() const { return Size(cols, rows); }
    size_t step1() const { return step.v; }
    template <class T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
    template <class T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
    template <class T> const T &at(Point2f p) const { return at<T>((int)p.y, (int)p.x); }
    void copyTo(Mat &dst) const {
        dst.create(rows, cols);
        for (int r = 0; r < rows; ++r) std::memcpy(dst.data + (size_t)r * dst.step.v, data + (size_t)r * step.v, (size_t)cols);
    }
    void copyTo(Mat &&dst) const {   // copy into a region-of-interest header
        for (int r = 0; r < rows; ++r) std::memcpy(dst.data + (size_t)r * dst.step.v, data + (size_t)r * step.v, (size_t)cols);
    }
private:
    std::shared_ptr<uchar> buf;
};

float fastAtan2(float y, float x);
void resize(const Mat &src, Mat &dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void vconcat(const Mat &a, const Mat &b, Mat &dst);
}  // namespace cv
