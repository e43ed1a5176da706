// Minimal stand-in for the OpenCV core types the reference's hot-path signatures use.  ONLY for compiling the
// shims in environments without OpenCV headers (this container); a real build uses the real <opencv2/core/core.hpp>.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_32FC1 5

namespace cv {
typedef unsigned char uchar;
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point3f { float x = 0, y = 0, z = 0; Point3f() {} Point3f(float a, float b, float c) : x(a), y(b), z(c) {} };
struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
class Mat {
public:
    int rows = 0, cols = 0, flags = 0;
    uchar* data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void* d, size_t s = 0) : rows(r), cols(c), flags(t), data((uchar*)d), step(s ? s : (size_t)c * esz(t)) {}
    static size_t esz(int t) { return t == CV_32F ? 4 : 1; }
    int type() const { return flags; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    size_t elemSize() const { return esz(flags); }
    void create(int r, int c, int t) {
        rows = r; cols = c; flags = t; step = (size_t)c * esz(t);
        buf_.reset(new std::vector<uchar>((size_t)r * step)); data = buf_->data();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; }
    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); if (m.data) std::memset(m.data, 0, (size_t)r * m.step); return m; }
    Mat clone() const { Mat m(rows, cols, flags); for (int y = 0; y < rows; y++) std::memcpy(m.data + y * m.step, data + y * step, (size_t)cols * elemSize()); return m; }
    template <class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <class T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <class T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    Mat row(int y) const { return Mat(1, cols, flags, data + (size_t)y * step, step); }
private:
    std::shared_ptr<std::vector<uchar>> buf_;
};
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
    _OutputArray(Mat& m) : _InputArray(m), o_(&m) {}
    void create(int r, int c, int t) const { o_->create(r, c, t); }
    void release() const { o_->release(); }
    Mat getMat() const { return *o_; }
private:
    Mat* o_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
}  // namespace cv
