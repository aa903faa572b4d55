// Minimal OpenCV header shim -- TEST INFRASTRUCTURE ONLY.
//
// Just enough of the cv:: surface for the reference's src/orbextractor.cpp to compile
// UNMODIFIED (token census in SURVEY.md 8c).  OpenCV itself is a third-party dependency of the
// reference (pinned 3.3.1, Dockerfile:29-31) whose source is not in the reference tree; the five
// pixel primitives are therefore implemented by the C oracle (oracle/orb_oracle.c), each of which
// is checked bit-for-bit against cv2 4.13 in tests/test_oracle_primitives.py.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_ &operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt(), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatZeros { int rows, cols, type; };

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar *data;
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(Size sz, int t) { alloc(sz.height, sz.width, t); }
    Mat(int r, int c, int t) { alloc(r, c, t); }
    Mat(int r, int c, int, void *ext, size_t stp) : rows(r), cols(c), step(stp), data((uchar *)ext) {}
    static MatZeros zeros(int r, int c, int t) { return MatZeros{r, c, t}; }
    Mat(const MatZeros &z) : rows(0), cols(0), step(0), data(nullptr) { alloc(z.rows, z.cols, z.type); }   // vector<uchar> value-initialises: zeros
    // Mat = MatExpr(zeros): OpenCV's create() keeps an existing buffer of the same size and zero-fills it in place
    Mat &operator=(const MatZeros &z) {
        if (rows != z.rows || cols != z.cols || tp != z.type || !data) alloc(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; r++) memset(data + (size_t)r * step, 0, (size_t)cols * esz());
        return *this;
    }
    void create(int r, int c, int t) { if (r != rows || c != cols || t != tp || !data) alloc(r, c, t); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    Mat operator()(const Rect &r) const { Mat m(*this); m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    void copyTo(Mat &dst) const { dst.create(rows, cols, tp); for (int r = 0; r < rows; r++) memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * esz()); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * esz(); m.cols = b - a; return m; }
    Mat col(int c) const { return colRange(c, c + 1); }
    void copyTo(const Mat &dst) const { Mat d(dst); copyTo(d); }                 // a row / range header of an existing matrix: written in place
    Mat reshape(int) const { return *this; }                                     // channel count only (orbframe.cpp:466-468); no element moves
    void convertTo(Mat &dst, int t) const;                                        // CV_8U -> CV_32F, dst may alias *this
    static Mat ones(int r, int c, int t);
    static Mat eye(int r, int c, int t);                                          // orbkeyframe.cpp:82
    Mat t() const;
    double dot(const Mat &b) const;
    template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    Mat clone() const { Mat m; m.alloc(rows, cols, tp); for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz()); return m; }
    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + c * sizeof(T)); }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    size_t step1() const { return step; }
    int type() const { return tp; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
private:
    std::shared_ptr<std::vector<uchar>> buf;
    int tp = CV_8UC1;                                    // CV_8U (1 byte) or CV_32F (4 bytes) elements
    size_t esz() const { return tp == CV_32F ? 4 : 1; }
    void alloc(int r, int c, int t = CV_8UC1) { tp = t; buf = std::make_shared<std::vector<uchar>>((size_t)r * c * esz()); data = buf->data(); rows = r; cols = c; step = (size_t)c * esz(); }
};

inline void Mat::convertTo(Mat &dst, int t) const
{
    assert(t == CV_32F);
    Mat o(rows, cols, CV_32F);
    for (int r = 0; r < rows; r++) for (int k = 0; k < cols; k++) o.ptr<float>(r)[k] = tp == CV_32F ? ptr<float>(r)[k] : (float)ptr(r)[k];
    dst = o;
}
inline Mat Mat::ones(int r, int c, int t)
{
    assert(t == CV_32F);
    Mat o(r, c, CV_32F);
    for (int i = 0; i < r; i++) for (int k = 0; k < c; k++) o.ptr<float>(i)[k] = 1.0f;
    return o;
}
inline Mat Mat::eye(int r, int c, int t)
{
    assert(t == CV_32F);
    Mat o(r, c, CV_32F);
    for (int i = 0; i < r; i++) for (int k = 0; k < c; k++) o.ptr<float>(i)[k] = i == k ? 1.0f : 0.0f;
    return o;
}
inline Mat Mat::t() const
{
    assert(tp == CV_32F);
    Mat o(cols, rows, CV_32F);
    for (int r = 0; r < rows; r++) for (int k = 0; k < cols; k++) o.ptr<float>(k)[r] = ptr<float>(r)[k];
    return o;
}
inline double Mat::dot(const Mat &b) const
{
    assert(tp == CV_32F && b.tp == CV_32F && rows == b.rows && cols == b.cols);
    double s = 0;
    for (int r = 0; r < rows; r++) for (int k = 0; k < cols; k++) s += (double)ptr<float>(r)[k] * b.ptr<float>(r)[k];
    return s;
}

// the little CV_32F arithmetic the reference's map-point / frame code uses (orbmappoint.cpp, orbframe.cpp): float matrices
inline Mat operator-(const Mat &a, const Mat &b)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    Mat c(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++) for (int k = 0; k < a.cols; k++) c.ptr<float>(r)[k] = a.ptr<float>(r)[k] - b.ptr<float>(r)[k];
    return c;
}
inline Mat operator+(const Mat &a, const Mat &b)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    Mat c(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++) for (int k = 0; k < a.cols; k++) c.ptr<float>(r)[k] = a.ptr<float>(r)[k] + b.ptr<float>(r)[k];
    return c;
}
inline Mat operator/(const Mat &a, double d)
{
    assert(a.type() == CV_32F);
    Mat c(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++) for (int k = 0; k < a.cols; k++) c.ptr<float>(r)[k] = (float)(a.ptr<float>(r)[k] / d);
    return c;
}
inline Mat operator*(const Mat &a, const Mat &b)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.cols == b.rows);
    Mat c(a.rows, b.cols, CV_32F);
    for (int r = 0; r < a.rows; r++) for (int k = 0; k < b.cols; k++) {
        float s = 0;
        for (int j = 0; j < a.cols; j++) s += a.ptr<float>(r)[j] * b.ptr<float>(j)[k];
        c.ptr<float>(r)[k] = s;
    }
    return c;
}
inline Mat operator*(double f, const Mat &a)
{
    assert(a.type() == CV_32F);
    Mat c(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++) for (int k = 0; k < a.cols; k++) c.ptr<float>(r)[k] = (float)(a.ptr<float>(r)[k] * f);
    return c;
}
inline Mat operator*(const Mat &a, double f) { return f * a; }
inline Mat operator-(const Mat &a) { return -1.0 * a; }
enum { NORM_L1 = 2 };
// cv::norm(a, b, NORM_L1) on CV_32F: OpenCV accumulates |a-b| in double (normDiffL1_<float, double>)
inline double norm(const Mat &a, const Mat &b, int kind)
{
    assert(kind == NORM_L1 && a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    double s = 0;
    for (int r = 0; r < a.rows; r++) for (int k = 0; k < a.cols; k++) s += std::fabs((double)a.ptr<float>(r)[k] - (double)b.ptr<float>(r)[k]);
    return s;
}
// (cv::Mat_<float>(r, c) << a, b, c): the comma initialiser of orbframe.cpp:739
template <typename T> struct Mat_ : Mat { Mat_(int r, int c) : Mat(r, c, CV_32F) {} };
template <typename T> struct MatCommaInit_ {
    Mat m; int n;
    MatCommaInit_ &operator,(T v) { m.ptr<T>(n / m.cols)[n % m.cols] = v; n++; return *this; }
    operator Mat() const { return m; }
};
template <typename T> inline MatCommaInit_<T> operator<<(const Mat_<T> &m, T v) { MatCommaInit_<T> c{m, 0}; return (c, v); }
inline double norm(const Mat &a)
{
    assert(a.type() == CV_32F);
    double s = 0;
    for (int r = 0; r < a.rows; r++) for (int k = 0; k < a.cols; k++) s += (double)a.ptr<float>(r)[k] * a.ptr<float>(r)[k];
    return std::sqrt(s);
}

class _InputArray {
public:
    _InputArray(const Mat &m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }
private:
    const Mat *m_;
};
typedef const _InputArray &InputArray;

class _OutputArray {
public:
    _OutputArray(Mat &m) : m_(&m) {}
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    Mat getMat() const { return *m_; }
    void release() const { m_->release(); }
private:
    Mat *m_;
};
typedef const _OutputArray &OutputArray;

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

// implemented in ref_glue.cpp on top of the C oracle's primitives
void resize(const Mat &src, Mat &dst, Size dsize, double fx, double fy, int interpolation);
void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType);
void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression);
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int borderType);
float fastAtan2(float y, float x);

} // namespace cv
