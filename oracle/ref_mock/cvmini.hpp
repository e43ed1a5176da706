// oracle/ref_mock/cvmini.hpp — TEST INFRASTRUCTURE ONLY (never included by the product).
//
// A header-only stand-in for the slice of the OpenCV C++ API that the reference's hot-path sources use, so that
// /root/reference/src/ORBextractor.cc and /root/reference/src/Event/EventConversion.cc compile UNMODIFIED, from where
// they lie, into oracle/_ref/libref.so (recipe: oracle/Makefile, target `ref`).  OpenCV's C++ headers do not exist in
// this image; the reference pins OpenCV 3.4.1 (build_eorb_slam.sh:136-139).
//
// What is the reference's own code and what is not, in a _ref build:
//   * every line of ORBextractor.cc / EventConversion.cc (grid loop, FAST fallback, DistributeOctTree and its pointer sort,
//     DivideNode, IC_Angle, computeOrbDescriptor, operator() assembly, ComputePyramid, the event splat loops, running
//     min/max, normalizeImage ...) is the reference's, compiled as is;
//   * the OpenCV primitives it calls are NOT OpenCV here: cv::resize / copyMakeBorder / GaussianBlur / FAST / fastAtan2
//     forward to the oracle's primitives (oracle/orb_oracle.cc), which tests/golden/make_golden.py pins bit-exactly to the
//     cv2 4.13.0 wheel; cv::Mat, KeyPoint, Point_, convertTo, meanStdDev are restated below from OpenCV's documented semantics.
// Anything the reference does not use is absent on purpose; unsupported arguments abort loudly.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../oracle.h"

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CVMINI_FAIL(msg) do { std::fprintf(stderr, "cvmini: unsupported: %s (%s:%d)\n", msg, __FILE__, __LINE__); std::abort(); } while (0)

// fast_math.hpp: cvRound = round-half-even (SSE cvtsd2si / cvtss2si), cvFloor / cvCeil
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(int v) { return v; }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(int v) { return v; }

namespace cv {
typedef ::uchar uchar;

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { NORM_MINMAX = 32 };

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
// types.hpp: a.x = saturate_cast<_Tp>(a.x * b); for float that is the plain float product
template <class T> static inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <class T> static inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <class T> static inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <class T> static inline Point_<T> operator*(const Point_<T>& a, int b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <class T> static inline Point_<T> operator*(const Point_<T>& a, float b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <class T> static inline Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }

template <class T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
typedef Size_<int> Size2i;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <class T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

class Mat {
public:
    int flags, rows, cols;
    uchar* data;
    size_t step;
    uchar* datastart;   // start of the allocation `data` points into (isSubmatrix / locateROI)

    Mat() : flags(0), rows(0), cols(0), data(nullptr), step(0), datastart(nullptr) {}
    Mat(int r, int c, int t) : Mat() { create(r, c, t); }
    Mat(Size s, int t) : Mat() { create(s.height, s.width, t); }
    Mat(int r, int c, int t, void* d, size_t s = 0)
        : flags(t), rows(r), cols(c), data((uchar*)d), step(s ? s : (size_t)c * esz(t)), datastart((uchar*)d) {}

    static size_t esz(int t) {
        switch (t) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2; case CV_32S: case CV_32F: return 4; case CV_64F: return 8; }
        CVMINI_FAIL("Mat type");
    }
    int type() const { return flags; }
    int depth() const { return flags; }
    int channels() const { return 1; }
    size_t elemSize() const { return esz(flags); }
    size_t elemSize1() const { return esz(flags); }
    size_t step1() const { return step / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    bool isSubmatrix() const { return buf_ && (data != datastart || step != (size_t)cols * elemSize() || bufRows_ != rows); }

    // Mat::create keeps the buffer when shape and type already match (the reference relies on it: resize() and
    // copyMakeBorder() write THROUGH the ROI headers ComputePyramid sets up, ORBextractor.cc:1247-1262)
    void create(int r, int c, int t) {
        if (data && rows == r && cols == c && flags == t) return;
        release();
        flags = t; rows = r; cols = c; step = (size_t)c * esz(t);
        size_t n = (size_t)r * step;
        if (n == 0) return;
        // malloc, not operator new: keeps Mat pixels out of the bump arena (ref_api.cc).  A zeroed guard band of 32 rows + 4 KiB
        // on either side makes the reference's out-of-bounds descriptor taps (edgeTh < 19, ORBextractor.cc:119-124 on the
        // borderless clone of :1141) deterministic instead of a fault; rows that took such a tap are excluded from the goldens.
        const size_t guard = 32 * step + 4096;
        uchar* p = (uchar*)std::calloc(n + 2 * guard, 1);
        if (!p) CVMINI_FAIL("out of memory");
        buf_.reset(p, std::free);
        data = datastart = p + guard; bufRows_ = r;
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    void release() { buf_.reset(); data = datastart = nullptr; rows = cols = 0; step = 0; bufRows_ = 0; }

    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); for (int y = 0; y < r; y++) std::memset(m.data + y * m.step, 0, (size_t)c * esz(t)); return m; }
    static Mat zeros(Size s, int t) { return zeros(s.height, s.width, t); }

    Mat clone() const { Mat m; copyToMat(m); return m; }
    void copyToMat(Mat& m) const {
        m.create(rows, cols, flags);
        for (int y = 0; y < rows; y++) std::memmove(m.data + y * m.step, data + y * step, (size_t)cols * elemSize());
    }
    inline void copyTo(OutputArray dst) const;
    inline void convertTo(OutputArray dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat t() const {
        if (flags != CV_32F) CVMINI_FAIL("Mat::t");
        Mat r(cols, rows, CV_32F);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) r.at<float>(x, y) = at<float>(y, x);
        return r;
    }
    inline double dot(const Mat& o) const;
    inline Mat inv() const;
    Mat mul(const Mat& o) const {
        if (flags != CV_32F || o.flags != CV_32F || rows != o.rows || cols != o.cols) CVMINI_FAIL("Mat::mul");
        Mat r(rows, cols, CV_32F);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) r.at<float>(y, x) = at<float>(y, x) * o.at<float>(y, x);
        return r;
    }

    Mat roi(int y0, int y1, int x0, int x1) const {
        if (y0 < 0 || y1 > rows || x0 < 0 || x1 > cols || y0 > y1 || x0 > x1) CVMINI_FAIL("ROI out of range");
        Mat m(*this);
        m.rows = y1 - y0; m.cols = x1 - x0; m.data = data + (size_t)y0 * step + (size_t)x0 * elemSize();
        return m;
    }
    Mat rowRange(int a, int b) const { return roi(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return roi(0, rows, a, b); }
    Mat row(int y) const { return roi(y, y + 1, 0, cols); }
    Mat col(int x) const { return roi(0, rows, x, x + 1); }
    Mat operator()(const Rect& r) const { return roi(r.y, r.y + r.height, r.x, r.x + r.width); }

    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <class T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <class T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    template <class T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }
    template <class T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }

private:
    std::shared_ptr<uchar> buf_;
    int bufRows_ = 0;
};
// ---- the little matrix algebra the matcher functions use on 3x3 / 3x1 / 4x4 CV_32F matrices (Rcw * x3Dw + tcw, -Rcw.t() * tcw,
// cv::norm): cv::gemm accumulates 32-bit products in double and rounds once (GEMMSingleMul<float, double>)
static inline Mat operator*(const Mat& a, const Mat& b) {
    if (a.type() != CV_32F || b.type() != CV_32F || a.cols != b.rows) CVMINI_FAIL("Mat * Mat");
    Mat r(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double acc = 0;
            for (int k = 0; k < a.cols; k++) acc += (double)a.at<float>(i, k) * (double)b.at<float>(k, j);
            r.at<float>(i, j) = (float)acc;
        }
    return r;
}
static inline Mat matAddSub(const Mat& a, const Mat& b, float sb) {
    if (a.type() != CV_32F || b.type() != CV_32F || a.rows != b.rows || a.cols != b.cols) CVMINI_FAIL("Mat +/- Mat");
    Mat r(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) r.at<float>(i, j) = a.at<float>(i, j) + sb * b.at<float>(i, j);
    return r;
}
static inline Mat operator+(const Mat& a, const Mat& b) { return matAddSub(a, b, 1.f); }
static inline Mat operator-(const Mat& a, const Mat& b) { return matAddSub(a, b, -1.f); }
static inline Mat operator-(const Mat& a) {
    if (a.type() != CV_32F) CVMINI_FAIL("-Mat");
    Mat r(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) r.at<float>(i, j) = -a.at<float>(i, j);
    return r;
}
// 3 x 3 CV_32F inverse the way cv::invert's small-matrix path forms it: determinant and cofactors in double, one rounding per element
static inline Mat matInv3(const Mat& m) {
    if (m.type() != CV_32F || m.rows != 3 || m.cols != 3) CVMINI_FAIL("Mat::inv (3x3 CV_32F only)");
    auto S = [&](int y, int x) { return (double)m.at<float>(y, x); };
    double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) + S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    Mat r = Mat::zeros(3, 3, CV_32F);
    if (d == 0.) return r;
    d = 1. / d;
    r.at<float>(0, 0) = (float)((S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d);
    r.at<float>(0, 1) = (float)((S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d);
    r.at<float>(0, 2) = (float)((S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d);
    r.at<float>(1, 0) = (float)((S(1, 2) * S(2, 0) - S(1, 0) * S(2, 2)) * d);
    r.at<float>(1, 1) = (float)((S(0, 0) * S(2, 2) - S(0, 2) * S(2, 0)) * d);
    r.at<float>(1, 2) = (float)((S(0, 2) * S(1, 0) - S(0, 0) * S(1, 2)) * d);
    r.at<float>(2, 0) = (float)((S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)) * d);
    r.at<float>(2, 1) = (float)((S(0, 1) * S(2, 0) - S(0, 0) * S(2, 1)) * d);
    r.at<float>(2, 2) = (float)((S(0, 0) * S(1, 1) - S(0, 1) * S(1, 0)) * d);
    return r;
}
// cv::Mat_<float>(r, c) << a, b, c ... (Pinhole::SkewSymmetricMatrix): row-major fill
template <class T> struct MatCommaInit_ {
    Mat m; int idx = 0;
    explicit MatCommaInit_(const Mat& mm) : m(mm) {}
    template <class V> MatCommaInit_& operator,(V v) { m.at<T>(idx / m.cols, idx % m.cols) = (T)v; idx++; return *this; }
    operator Mat() const { return m; }
};
template <class T> struct Mat_ : public Mat {
    Mat_(int r, int c) : Mat(r, c, CV_32F) { static_assert(sizeof(T) == 4, "Mat_<float> only"); }
    template <class V> MatCommaInit_<T> operator<<(V v) { MatCommaInit_<T> ci(*this); ci, v; return ci; }
};
// scalar scaling and the dot product of the keyframe-side searches (sRcw / scw, s12 * R12, (1.0 / s12) * R12.t(), row.dot(row),
// PO.dot(Pn): ORBmatcher.cc:491-494, :538, :1757-1759): element * double factor rounded once, dot accumulated in double like cv::Mat::dot.
// The pin only calls these with factor 1 and identity rotations, where every formula OpenCV could use gives the same floats.
static inline Mat matScale(const Mat& a, double f) {
    if (a.type() != CV_32F) CVMINI_FAIL("Mat * scalar");
    Mat r(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) r.at<float>(i, j) = (float)((double)a.at<float>(i, j) * f);
    return r;
}
static inline Mat operator*(double f, const Mat& a) { return matScale(a, f); }
static inline Mat operator*(const Mat& a, double f) { return matScale(a, f); }
static inline Mat operator/(const Mat& a, double f) { return matScale(a, 1.0 / f); }
static inline double matDot(const Mat& a, const Mat& b) {
    if (a.type() != CV_32F || b.type() != CV_32F || a.rows != b.rows || a.cols != b.cols) CVMINI_FAIL("Mat::dot");
    double s = 0;
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) s += (double)a.at<float>(i, j) * (double)b.at<float>(i, j);
    return s;
}
inline double Mat::dot(const Mat& o) const { return matDot(*this, o); }
inline Mat Mat::inv() const { return matInv3(*this); }
static inline double norm(const Mat& a) {
    if (a.type() != CV_32F) CVMINI_FAIL("norm");
    double s = 0;
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) s += (double)a.at<float>(i, j) * (double)a.at<float>(i, j);
    return std::sqrt(s);
}

static inline std::ostream& operator<<(std::ostream& os, const Mat& m) { return os << "Mat(" << m.rows << "x" << m.cols << ")"; }

class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
protected:
    const Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() : _InputArray(), o_(nullptr) {}
    _OutputArray(Mat& m) : _InputArray(m), o_(&m) {}
    _OutputArray(const Mat& m) : _InputArray(m), o_(const_cast<Mat*>(&m)) {}   // a temporary header (desc.row(i)): fixed size
    void create(int r, int c, int t) const { if (o_) o_->create(r, c, t); }
    void create(Size s, int t) const { create(s.height, s.width, t); }
    void release() const { if (o_) o_->release(); }
    Mat getMat() const { return o_ ? *o_ : Mat(); }
    Mat& getMatRef() const { return *o_; }
    bool needed() const { return o_ != nullptr; }
private:
    Mat* o_;
};
static inline InputArray noArray() { static _OutputArray none; return none; }

inline void Mat::copyTo(OutputArray dst) const {
    if (empty()) { dst.release(); return; }
    dst.create(rows, cols, flags);
    Mat d = dst.getMat();
    if (d.rows != rows || d.cols != cols) CVMINI_FAIL("copyTo into a fixed-size header of another size");
    for (int y = 0; y < rows; y++) std::memmove(d.data + y * d.step, data + y * step, (size_t)cols * elemSize());
}

// convertTo 32F -> 8U with scale: cvtScale_<float, uchar, float>: saturate_cast<uchar>(src*(float)alpha + (float)beta),
// saturate_cast<uchar>(float) = saturate(cvRound(v)).  A type-changing convertTo allocates a fresh destination even when
// source and destination are the same Mat object (normalizeImage(image, image, ...), EventConversion.cc:68-73).
inline void Mat::convertTo(OutputArray dst, int rtype, double alpha, double beta) const {
    Mat src = *this;   // keeps the source buffer alive across the re-create
    if (rtype < 0) rtype = flags;
    if (src.flags == CV_32F && rtype == CV_8U) {
        Mat out(rows, cols, CV_8U);
        const float a = (float)alpha, b = (float)beta;
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                float v = src.at<float>(y, x) * a;
                v = v + b;
                int i = (int)lrintf(v);
                out.at<uchar>(y, x) = (uchar)(i < 0 ? 0 : (i > 255 ? 255 : i));
            }
        dst.getMatRef() = out;
        return;
    }
    if (src.flags == rtype && alpha == 1 && beta == 0) { Mat out = src.clone(); dst.getMatRef() = out; return; }
    CVMINI_FAIL("convertTo combination");
}

// ---- the five OpenCV primitives ORBextractor.cc calls: forwarded to the cv2-pinned oracle primitives ----
static inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    Mat src = _src.getMat();
    if (src.type() != CV_8U || interpolation != INTER_LINEAR || fx != 0 || fy != 0 || dsize.width <= 0 || dsize.height <= 0) CVMINI_FAIL("resize arguments");
    _dst.create(dsize.height, dsize.width, CV_8U);
    Mat dst = _dst.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType,
                                  const Scalar& = Scalar()) {
    Mat src = _src.getMat();
    const int bt = borderType & ~BORDER_ISOLATED;
    if (src.type() != CV_8U || bt != BORDER_REFLECT_101 || top != bottom || top != left || top != right) CVMINI_FAIL("copyMakeBorder arguments");
    // without BORDER_ISOLATED OpenCV reads real pixels around a sub-matrix source; the reference's only non-isolated call
    // (ORBextractor.cc:1258, level 0) passes the caller's whole image, so a sub-matrix there is out of this mock's contract
    if (!(borderType & BORDER_ISOLATED) && src.isSubmatrix()) CVMINI_FAIL("non-isolated copyMakeBorder of a sub-matrix");
    Mat keep = src.clone();   // the reference calls it with src = an ROI of dst (in place)
    _dst.create(src.rows + 2 * top, src.cols + 2 * top, CV_8U);
    Mat dst = _dst.getMat();
    orc_copy_make_border_reflect101(keep.data, keep.cols, keep.rows, keep.step, dst.data, top, dst.step);
}

struct GaussTap { std::vector<Mat>* sink = nullptr; };
inline GaussTap& gaussTap() { static thread_local GaussTap t; return t; }

static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
    Mat src = _src.getMat();
    if (src.type() != CV_8U || ksize.width != 5 || ksize.height != 5 || sigmaX != 2.0 || sigmaY != 2.0 || borderType != BORDER_REFLECT_101 || src.isSubmatrix())
        CVMINI_FAIL("GaussianBlur arguments (only 5x5, sigma 2, REFLECT_101, whole 8-bit matrix)");
    Mat keep = src.clone();   // in place in the reference
    _dst.create(src.rows, src.cols, CV_8U);
    Mat dst = _dst.getMat();
    orc_gauss5x5_s2_u8(keep.data, keep.cols, keep.rows, keep.step, dst.data, dst.step);
    if (gaussTap().sink) gaussTap().sink->push_back(dst.clone());
}

// per-thread tap of the FAST calls of one extraction (read by oracle/ref_api.cc; not part of OpenCV)
struct FastTap { int calls = 0, calls_nonempty = 0; long candidates = 0; std::vector<int>* thresholds = nullptr; };
inline FastTap& fastTap() { static thread_local FastTap t; return t; }

static inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    Mat img = _img.getMat();
    if (img.type() != CV_8U) CVMINI_FAIL("FAST image type");
    keypoints.clear();
    FastTap& tap = fastTap();
    tap.calls++;
    if (tap.thresholds) tap.thresholds->push_back(threshold);
    if (img.rows < 7 || img.cols < 7) return;
    const int cap = img.rows * img.cols;
    std::vector<int> xs(cap), ys(cap), sc(cap);
    const int n = orc_fast9_16(img.data, img.cols, img.rows, img.step, threshold, nonmaxSuppression ? 1 : 0, xs.data(), ys.data(), sc.data(), cap);
    keypoints.reserve(n);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));   // fast.cpp: KeyPoint(j, i-1, 7.f, -1, score)
    tap.candidates += n;
    tap.calls_nonempty += n > 0;
}

struct KeyPointsFilter {
    // keypoint.cpp: keep the n best by response plus everything tied with the n-th (only ComputeKeyPointsOld, dead code, calls it)
    static void retainBest(std::vector<KeyPoint>& kps, int n) {
        if (n < 0 || (int)kps.size() <= n) return;
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        const float amb = kps[n - 1].response;
        auto e = std::partition(kps.begin() + n, kps.end(), [amb](const KeyPoint& k) { return k.response >= amb; });
        kps.resize(e - kps.begin());
    }
};

// ---- statistics used by EventConversion.cc (mean.cpp / meanStdDev: double accumulation of sum and square sum) ----
static inline void meanStdDev(InputArray _src, Scalar& mean, Scalar& stddev) {
    Mat src = _src.getMat();
    if (src.type() != CV_32F) CVMINI_FAIL("meanStdDev type");
    double s = 0, sq = 0;
    for (int y = 0; y < src.rows; y++) {
        const float* p = src.ptr<float>(y);
        for (int x = 0; x < src.cols; x++) { double v = p[x]; s += v; sq += v * v; }
    }
    const double n = (double)src.total();
    const double m = n > 0 ? s / n : 0;
    double var = n > 0 ? sq / n - m * m : 0;
    mean = Scalar(m);
    stddev = Scalar(std::sqrt(var > 0 ? var : 0));
}
static inline Scalar mean(InputArray _src) {
    Mat src = _src.getMat();
    if (src.type() != CV_32F) CVMINI_FAIL("mean type");
    double s = 0;
    for (int y = 0; y < src.rows; y++) { const float* p = src.ptr<float>(y); for (int x = 0; x < src.cols; x++) s += p[x]; }
    return Scalar(src.total() ? s / (double)src.total() : 0);
}

class FileStorage;   // only named in declarations the hot path never reaches (EventData.h:122)
}  // namespace cv

// glog stand-in: LOG(sev) << ...  is swallowed
struct CvMiniNullLog { template <class T> CvMiniNullLog& operator<<(const T&) { return *this; } };
#ifndef LOG
#define LOG(sev) CvMiniNullLog()
#endif
#ifndef DLOG
#define DLOG(sev) CvMiniNullLog()
#endif
