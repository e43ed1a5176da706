// oracle/ref_mock/eigenmini.hpp — TEST INFRASTRUCTURE ONLY (never included by the product).
//
// Stand-in for <Eigen/Core> + <Eigen/Geometry> (Eigen is not in this image), holding only what the reference's
// src/Event/EventConversion.cc touches: small fixed-size double matrices with the comma initialiser, products, sums and
// scalar scaling, and AngleAxisd <-> rotation matrix through a quaternion.  The rotation conversions follow the published
// Eigen 3.3 Geometry algorithms (Quaternion from a matrix: the trace / largest-diagonal branches; AngleAxis from a quaternion:
// 2*atan2(|vec|, |w|); AngleAxis::toRotationMatrix: Rodrigues in Eigen's operation order).  Products are summed left to
// right; Eigen's unrolled reductions may associate differently, a last-bit effect in DOUBLE that is far inside the float
// tolerance of this path (event-frame intensities <= 1e-4 of peak).
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>

namespace Eigen {

template <class S, int R, int C> class Matrix;

template <class S, int R, int C> class CommaInit {
public:
    CommaInit(Matrix<S, R, C>& m, S first) : m_(m), i_(0) { put(first); }
    template <class T> CommaInit& operator,(const T& v) { put((S)v); return *this; }
private:
    void put(S v) { m_.d[i_ / C][i_ % C] = v; ++i_; }   // row-major fill, as Eigen's comma initialiser
    Matrix<S, R, C>& m_;
    int i_;
};

template <class S, int R, int C> class Matrix {
public:
    S d[R][C];
    Matrix() { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) d[i][j] = S(0); }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); i++) m.d[i][i] = S(1); return m; }
    S& operator()(int i, int j) { return d[i][j]; }
    const S& operator()(int i, int j) const { return d[i][j]; }
    S& operator()(int i) { return R == 1 ? d[0][i] : d[i][0]; }
    const S& operator()(int i) const { return R == 1 ? d[0][i] : d[i][0]; }
    S& operator[](int i) { return (*this)(i); }
    const S& operator[](int i) const { return (*this)(i); }
    S& x() { return (*this)(0); } S& y() { return (*this)(1); } S& z() { return (*this)(2); }
    const S& x() const { return (*this)(0); } const S& y() const { return (*this)(1); } const S& z() const { return (*this)(2); }
    template <class T> CommaInit<S, R, C> operator<<(const T& v) { return CommaInit<S, R, C>(*this, (S)v); }
    Matrix operator-() const { Matrix r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r.d[i][j] = -d[i][j]; return r; }
    Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r.d[i][j] = d[i][j] + o.d[i][j]; return r; }
    Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r.d[i][j] = d[i][j] - o.d[i][j]; return r; }
    template <class T> Matrix& operator*=(const T& s) { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) d[i][j] *= (S)s; return *this; }
    template <int K> Matrix<S, R, K> operator*(const Matrix<S, C, K>& o) const {
        Matrix<S, R, K> r;
        for (int i = 0; i < R; i++) for (int k = 0; k < K; k++) { S a = d[i][0] * o.d[0][k]; for (int j = 1; j < C; j++) a += d[i][j] * o.d[j][k]; r.d[i][k] = a; }
        return r;
    }
    Matrix<S, C, R> transpose() const { Matrix<S, C, R> r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r.d[j][i] = d[i][j]; return r; }
    S squaredNorm() const { S a = 0; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) a += d[i][j] * d[i][j]; return a; }
    S norm() const { return std::sqrt(squaredNorm()); }
    S trace() const { S a = 0; for (int i = 0; i < (R < C ? R : C); i++) a += d[i][i]; return a; }
};
// scalar * matrix and matrix * scalar for any arithmetic scalar (the reference mixes float and double)
template <class S, int R, int C> Matrix<S, R, C> operator*(const Matrix<S, R, C>& m, double s) { Matrix<S, R, C> r(m); r *= s; return r; }
template <class S, int R, int C> Matrix<S, R, C> operator*(double s, const Matrix<S, R, C>& m) { Matrix<S, R, C> r(m); r *= s; return r; }
template <class S, int R, int C> Matrix<S, R, C> operator/(const Matrix<S, R, C>& m, double s) {
    Matrix<S, R, C> r(m); for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r.d[i][j] = m.d[i][j] / (S)s; return r;
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

class Quaterniond {
public:
    double q[4];   // x, y, z, w
    Quaterniond() { q[0] = q[1] = q[2] = 0; q[3] = 1; }
    Quaterniond(double w, double x, double y, double z) { q[0] = x; q[1] = y; q[2] = z; q[3] = w; }
    explicit Quaterniond(const Matrix3d& m) {   // Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3>
        double t = m.trace();
        if (t > 0.0) {
            t = std::sqrt(t + 1.0);
            q[3] = 0.5 * t;
            t = 0.5 / t;
            q[0] = (m(2, 1) - m(1, 2)) * t;
            q[1] = (m(0, 2) - m(2, 0)) * t;
            q[2] = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
            q[i] = 0.5 * t;
            t = 0.5 / t;
            q[3] = (m(k, j) - m(j, k)) * t;
            q[j] = (m(j, i) + m(i, j)) * t;
            q[k] = (m(k, i) + m(i, k)) * t;
        }
    }
    double x() const { return q[0]; } double y() const { return q[1]; } double z() const { return q[2]; } double w() const { return q[3]; }
    Matrix3d toRotationMatrix() const {   // QuaternionBase::toRotationMatrix
        Matrix3d r;
        const double tx = 2 * x(), ty = 2 * y(), tz = 2 * z();
        const double twx = tx * w(), twy = ty * w(), twz = tz * w();
        const double txx = tx * x(), txy = ty * x(), txz = tz * x();
        const double tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
        return r;
    }
};

class AngleAxisd {
public:
    AngleAxisd() : angle_(0) { axis_ << 1, 0, 0; }
    AngleAxisd(double angle, const Vector3d& axis) : angle_(angle), axis_(axis) {}
    explicit AngleAxisd(const Matrix3d& m) { *this = fromQuaternion(Quaterniond(m)); }   // AngleAxis::fromRotationMatrix
    double angle() const { return angle_; }
    const Vector3d& axis() const { return axis_; }
    Matrix3d toRotationMatrix() const {   // Eigen/src/Geometry/AngleAxis.h
        Matrix3d res;
        Vector3d sin_axis = axis_ * std::sin(angle_);
        const double c = std::cos(angle_);
        Vector3d cos1_axis = axis_ * (1.0 - c);
        double tmp;
        tmp = cos1_axis.x() * axis_.y(); res(0, 1) = tmp - sin_axis.z(); res(1, 0) = tmp + sin_axis.z();
        tmp = cos1_axis.x() * axis_.z(); res(0, 2) = tmp + sin_axis.y(); res(2, 0) = tmp - sin_axis.y();
        tmp = cos1_axis.y() * axis_.z(); res(1, 2) = tmp - sin_axis.x(); res(2, 1) = tmp + sin_axis.x();
        for (int i = 0; i < 3; i++) res(i, i) = cos1_axis(i) * axis_(i) + c;
        return res;
    }
    operator Matrix3d() const { return toRotationMatrix(); }   // RotationBase -> matrix assignment
private:
    static AngleAxisd fromQuaternion(const Quaterniond& q) {   // AngleAxis::operator=(const QuaternionBase&), Eigen 3.3
        AngleAxisd r;
        double n = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z());
        if (n != 0.0) {
            r.angle_ = 2.0 * std::atan2(n, std::fabs(q.w()));
            if (q.w() < 0) n = -n;
            r.axis_(0) = q.x() / n; r.axis_(1) = q.y() / n; r.axis_(2) = q.z() / n;
        } else {
            r.angle_ = 0.0;
            r.axis_ << 1, 0, 0;
        }
        return r;
    }
    double angle_;
    Vector3d axis_;
};

}  // namespace Eigen
