// Small fixed-size vectors / matrices for the host adapters (the reference uses Eigen, which is not part of this
// repository's dependencies): just the operations the matcher adapters of slam_matchers.hpp and the keyframe geometry
// in slam_frontend.cpp need, evaluated eagerly.  Sums of 3 or 4 products follow Eigen's unrolled reduction order
// (halves: x0 + (x1 + x2), (x0 + x1) + (x2 + x3)) so that results agree with the reference built on Eigen.
#pragma once
#include <cmath>
#include <cstddef>
#include <type_traits>
#include <utility>

namespace slam {
namespace la {
namespace detail {
template <class T> inline T tree_sum(const T *v, int n) {
    if (n == 1) return v[0];
    const int h = n / 2;
    return tree_sum(v, h) + tree_sum(v + h, n - h);
}
}  // namespace detail

template <class T, int R, int C> struct Matrix;

template <class T, int R, int C> struct CommaInit {
    Matrix<T, R, C> *m;
    int i;
    CommaInit &operator,(const T &v) { m->d[i++] = v; return *this; }
};

template <class T, int R, int C> struct Matrix {
    static_assert(R > 0 && C > 0, "fixed sizes only");
    T d[R * C];   // row-major: (r, c) -> d[r * C + c]

    Matrix() : d{} {}
    template <int N = R * C, class = std::enable_if_t<N == 2>> Matrix(T a, T b) : d{a, b} {}
    template <int N = R * C, class = std::enable_if_t<N == 3>> Matrix(T a, T b, T c) : d{a, b, c} {}
    // row vector <-> column vector (Eigen transposes vectors implicitly on assignment)
    template <int R2, int C2, class = std::enable_if_t<(R2 != R) && R2 * C2 == R * C && (R == 1 || C == 1) && (R2 == 1 || C2 == 1)>>
    Matrix(const Matrix<T, R2, C2> &o) { for (int i = 0; i < R * C; ++i) d[i] = o.d[i]; }

    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }

    T &operator()(int r, int c) { return d[r * C + c]; }
    const T &operator()(int r, int c) const { return d[r * C + c]; }
    T &operator()(int i) { return d[i]; }
    const T &operator()(int i) const { return d[i]; }
    T &operator[](int i) { return d[i]; }
    const T &operator[](int i) const { return d[i]; }
    const T &x() const { return d[0]; }
    const T &y() const { return d[1]; }
    const T &z() const { return d[2]; }
    T &x() { return d[0]; }
    T &y() { return d[1]; }
    T &z() { return d[2]; }
    int rows() const { return R; }
    int cols() const { return C; }

    CommaInit<T, R, C> operator<<(const T &v) { d[0] = v; return CommaInit<T, R, C>{this, 1}; }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> t;
        for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c);
        return t;
    }
    template <int BR, int BC> Matrix<T, BR, BC> block(int r0, int c0) const {
        Matrix<T, BR, BC> b;
        for (int r = 0; r < BR; ++r) for (int c = 0; c < BC; ++c) b(r, c) = (*this)(r0 + r, c0 + c);
        return b;
    }
    template <int BR, int BC> Matrix<T, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    template <class U> Matrix<U, R, C> cast() const {
        Matrix<U, R, C> o;
        for (int i = 0; i < R * C; ++i) o.d[i] = static_cast<U>(d[i]);
        return o;
    }

    Matrix operator-() const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = -d[i]; return o; }
    Matrix operator+(const Matrix &b) const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] + b.d[i]; return o; }
    Matrix operator-(const Matrix &b) const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] - b.d[i]; return o; }
    Matrix &operator+=(const Matrix &b) { for (int i = 0; i < R * C; ++i) d[i] += b.d[i]; return *this; }
    Matrix &operator-=(const Matrix &b) { for (int i = 0; i < R * C; ++i) d[i] -= b.d[i]; return *this; }
    template <class U, class = std::enable_if_t<std::is_arithmetic<U>::value>> Matrix operator*(U s) const {
        const T t = static_cast<T>(s); Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] * t; return o;
    }
    template <class U, class = std::enable_if_t<std::is_arithmetic<U>::value>> Matrix operator/(U s) const {
        const T t = static_cast<T>(s); Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] / t; return o;
    }
    // coefficient-based lazy product, each coefficient a tree-ordered sum (see the header note)
    template <int C2> Matrix<T, R, C2> operator*(const Matrix<T, C, C2> &b) const {
        Matrix<T, R, C2> o;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C2; ++c) {
                T p[C];
                for (int k = 0; k < C; ++k) p[k] = (*this)(r, k) * b(k, c);
                o(r, c) = detail::tree_sum(p, C);
            }
        return o;
    }
    T dot(const Matrix &b) const {
        T p[R * C];
        for (int i = 0; i < R * C; ++i) p[i] = d[i] * b.d[i];
        return detail::tree_sum(p, R * C);
    }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const {
        const T z = squaredNorm();
        if (z > T(0)) return *this / std::sqrt(z);
        return *this;
    }
    bool isZero(T prec = T(1e-12)) const {
        for (int i = 0; i < R * C; ++i) if (std::abs(d[i]) > prec) return false;
        return true;
    }
    // general inverse (Gauss-Jordan with partial pivoting); Eigen's 4x4 kernel rounds differently, which the
    // oracle/_ref tests do not depend on (they record the projected queries the reference code produced)
    Matrix inverse() const {
        static_assert(R == C, "square");
        Matrix a = *this, inv = Identity();
        for (int col = 0; col < R; ++col) {
            int piv = col;
            for (int r = col + 1; r < R; ++r) if (std::abs(a(r, col)) > std::abs(a(piv, col))) piv = r;
            if (piv != col) for (int c = 0; c < C; ++c) { std::swap(a(col, c), a(piv, c)); std::swap(inv(col, c), inv(piv, c)); }
            const T s = T(1) / a(col, col);
            for (int c = 0; c < C; ++c) { a(col, c) *= s; inv(col, c) *= s; }
            for (int r = 0; r < R; ++r) {
                if (r == col) continue;
                const T f = a(r, col);
                if (f == T(0)) continue;
                for (int c = 0; c < C; ++c) { a(r, c) -= f * a(col, c); inv(r, c) -= f * inv(col, c); }
            }
        }
        return inv;
    }
};

template <class U, class T, int R, int C, class = std::enable_if_t<std::is_arithmetic<U>::value>>
inline Matrix<T, R, C> operator*(U s, const Matrix<T, R, C> &m) { return m * s; }

using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector2f = Matrix<float, 2, 1>;
using Vector3f = Matrix<float, 3, 1>;
}  // namespace la
}  // namespace slam
